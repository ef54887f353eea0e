"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import snickery_b200
from snickery_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "snk_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(snk_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_and_loads():
    assert os.path.exists(engine.LIB_PATH), "run `python -m snickery_b200.build`"
    lib = engine.load_library()
    assert lib.snk_version() >= 100


def test_every_header_symbol_is_exported():
    lib = ctypes.CDLL(engine.LIB_PATH)
    names = header_functions()
    assert len(names) >= 19
    for name in names:
        assert hasattr(lib, name), "include/snk_b200.h declares %s but the library does not export it" % name
    assert sorted(engine.EXPORTED_SYMBOLS) == names


def test_no_torch_or_oracle_in_product():
    # the product path must not import the oracle, and the ABI must not carry torch types
    for dirpath, _, files in os.walk(os.path.join(ROOT, "snickery_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "snk_b200.h")).read(), flags=re.S)
    assert "torch" not in hdr and "at::" not in hdr and "Tensor" not in hdr


@pytest.mark.skipif(engine.load_library().snk_device_count() > 0, reason="GPU present")
def test_fails_loudly_without_gpu():
    with pytest.raises(snickery_b200.EngineError, match="no CPU fallback"):
        snickery_b200.UnitDatabase(np.zeros((10, 3), np.float32), np.zeros((11, 2), np.float32))
    with pytest.raises(snickery_b200.EngineError):
        snickery_b200.GpuKDTree(np.zeros((10, 3)))


def test_argument_validation_happens_before_device_work():
    with pytest.raises(ValueError):
        snickery_b200.UnitDatabase(np.zeros((10, 3), np.float32), np.zeros((10, 2), np.float32))  # Jc needs N+1 rows
