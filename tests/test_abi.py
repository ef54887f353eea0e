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


def _function_bodies(src):
    """name -> body of every top-level C function definition in src (brace matching; strings in this code base hold no braces)."""
    import re
    out = {}
    for m in re.finditer(r"^(?:static\s+|extern\s+\"C\"\s+)?(?:int|bool|void|float)\s+(\w+)\s*\([^;{]*\)\s*\{", src, re.M):
        depth, i = 1, m.end()
        while depth and i < len(src):
            depth += {"{": 1, "}": -1}.get(src[i], 0)
            i += 1
        out[m.group(1)] = src[m.end():i]
    return out


def test_dev_entry_points_do_not_synchronise():
    """include/snk_b200.h promises that `_dev` entry points only enqueue on the caller's stream.  Walk the call graph of every
    `_dev` function through the engine's own functions (everything but the `*_finish` calls, which are where waiting belongs)
    and make sure no blocking CUDA call is reachable."""
    import glob
    import re
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "snickery_b200", "csrc")
    bodies = {}
    for path in glob.glob(os.path.join(csrc, "*.cu")):
        bodies.update(_function_bodies(open(path).read()))
    roots = [n for n in bodies if n.endswith("_dev")]
    assert {"snk_knn_dev", "snk_greedy_batch_dev", "snk_join_viterbi_batch_dev", "snk_acoustic_viterbi_batch_dev",
            "snk_knn_sharded_dev", "snk_greedy_sharded_batch_dev"} <= set(roots)
    blocking = re.compile(r"cudaStreamSynchronize|cudaDeviceSynchronize|cudaEventSynchronize|cudaMemcpy\s*\(|cudaMemcpy2D\s*\(|cudaFree\s*\(")
    seen, todo = set(), list(roots)
    while todo:
        name = todo.pop()
        if name in seen or name.endswith("_finish"):
            continue
        seen.add(name)
        body = bodies[name]
        hit = blocking.search(body)
        # two documented exceptions (include/snk_b200.h): the staging ring waits on an event only when it wraps around an
        # upload that is still pending eight uploads later, and a grow-only workspace is reallocated the first time a call
        # needs more than any call before it
        assert hit is None or name in ("snk_upload_async", "snk_buf_reserve", "take_flags"), "%s reaches %s" % (name, hit.group(0))
        todo += [c for c in set(re.findall(r"\b(\w+)\s*\(", body)) if c in bodies]
    assert len(seen) > 25


def _build_c_host(tmp_path):
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "c_host")
    libdir = os.path.dirname(engine.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_host.c"), "-L", libdir, "-lsnk_b200", "-Wl,-rpath," + libdir,
                           "-lm", "-o", exe])
    return exe


def test_header_is_c_and_a_c_host_links(tmp_path):
    """The boundary is a C ABI: include/snk_b200.h must compile as C99 and a plain C program (examples/c_host.c) must link
    against the library with nothing but -lsnk_b200.  Without a GPU it reports that and exits 2 (no compute, no fallback)."""
    import subprocess
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                           os.path.join(ROOT, "include", "snk_b200.h")])
    exe = _build_c_host(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode in (0, 2), (r.returncode, r.stdout, r.stderr)
    if r.returncode == 2:
        assert "no CUDA device" in r.stderr or "snk_db_create" in r.stderr


@pytest.mark.gpu
def test_c_host_recovers_the_identity_path(tmp_path):
    """The same C program on a B200: database frames in, consecutive units out at distance 0 (the reference's own
    assertion, synth_simple.py:909-928), through snk_db_create / snk_db_set_weights / snk_greedy_batch from C."""
    import subprocess
    r = subprocess.run([_build_c_host(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-500:])
    assert "identity path recovered" in r.stdout
