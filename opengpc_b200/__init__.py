"""opengpc_b200 -- B200 (sm_100a) implementation of openGPC's Global Patch Collider inference
path (preprocessImage x2 + rectifiedMatch, the window samples/sparsematch.cpp:45-52 times).

Layout: csrc/ (hand-written CUDA kernels + the C ABI of include/gpc_b200.h), capi.py (ctypes
binding used by tests and bench.py), synth.py (synthetic stereo pairs), build.py (nvcc build).
The drop-in C++ API mirroring gpc::inference::Forest is in include/gpc/.
"""
from .capi import (CORR_DTYPE, MATCHER_AUTO, MATCHER_ROWS_GENERAL, MATCHER_SORT, Context, GpcError, Pool, GpcForest, GpcSettings, SUPPORT_DTYPE, load_library, make_forest, make_settings,
                   read_forest, read_forest_tests, sparsematch_settings)

__all__ = ["CORR_DTYPE", "MATCHER_AUTO", "MATCHER_ROWS_GENERAL", "MATCHER_SORT", "Context", "GpcError", "Pool", "GpcForest", "GpcSettings", "SUPPORT_DTYPE", "load_library", "make_forest",
           "make_settings", "read_forest", "read_forest_tests", "sparsematch_settings"]
