"""sampler_b200 -- B200-native Gibbs sampling hot path for DimmWitted factor graphs.

The product is native: CUDA sweep kernels for sm_100a + a C++ host behind the C ABI of
include/dwg.h (libdwg.so, `dw` CLI).  This Python package is only the ctypes stub over
that ABI (api.py), the build recipe (_build.py) and input tooling (graphio.py, graphgen.py).
"""
from . import graphio  # noqa: F401
from .api import (DwgError, Graph, Sampler, SCHED_AUTO, SCHED_CHROMATIC, SCHED_HOGWILD, SCHED_SEQUENTIAL,  # noqa: F401
                  draw, dw_main, load, logadd, release_cached_memory)
