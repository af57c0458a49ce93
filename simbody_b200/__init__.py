"""simbody_b200 -- B200-native batched forward dynamics for tree-topology multibody systems.

The product is the C-ABI shared library `libsbk.so` (include/sbk.h: hand-written FP64 CUDA for
sm_100a + C++ host code).  This package is the thin Python view of that ABI used by the tests
and by bench.py; it holds no compute of its own and has no CPU fallback.
"""
from .capi import SbkError, load_library, library_path  # noqa: F401
from .engine import BatchedMatter, Topology, model_text  # noqa: F401

__all__ = ["SbkError", "load_library", "library_path", "BatchedMatter", "Topology", "model_text"]
