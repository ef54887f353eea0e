"""snickery_b200 -- B200-native unit-selection search (k-NN -> join tiles -> greedy / Viterbi).

The CUDA engine lives in csrc/ behind the C ABI of include/snk_b200.h; `engine` binds it with
ctypes, `kdtree` and `synth` mirror the reference's call sites.  No CPU fallback exists.
"""
from . import engine  # noqa: F401
from .engine import EngineError, FrameStore, UnitDatabase  # noqa: F401
from .kdtree import GpuKDTree, GpuStashableKDTree  # noqa: F401
from .synth import Synthesiser  # noqa: F401

__version__ = "0.1.0"
