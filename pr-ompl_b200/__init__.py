"""pr-ompl_b200: B200-native (sm_100a) hot path behind pr-ompl's RRTConnect and NNF planners.

The product is the C-ABI shared library `libprgpu.so` (include/prgpu.h, sources in csrc/) and
the C++ OMPL-plugin mirror in plugin/.  This Python package is a thin ctypes binding used by
tests and bench.py; it computes nothing itself and raises if the CUDA library is missing.
"""
from . import synth, sharded  # noqa: F401
from .capi import (KnnIndex, World, PrgpuError, build, lib, lib_path, topk_merge_device,  # noqa: F401
                   INVALID_INDEX, MODE_OMPL_DISCRETE, MODE_NNF_ALPHA, RrtcBatch, RrtcParams, Nnf, Comm)
