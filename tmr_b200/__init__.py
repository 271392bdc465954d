"""tmr_b200 -- B200-native TMROctForest hot path (refine / balance /
createNodes / createInterpolation) behind the reference's class API.

The only compute backend is the CUDA library tmr_b200/lib/libtmr_b200.so
(built by `__graft_entry__.build()`); there is NO CPU fallback: loading fails
loudly if the library is missing.
"""
import ctypes
import os

from . import _capi

_LIB = None
LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")
LIB_PATH = os.path.join(LIB_DIR, "libtmr_b200.so")


def load_library():
    """Load and bind the product library (CUDA).  Raises if it is not built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "tmr_b200: %s is not built; run `python -c 'import "
                "__graft_entry__ as g; g.build()'` (there is no CPU fallback)"
                % LIB_PATH)
        _LIB = _capi.bind(ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL))
    return _LIB


from .forest import (  # noqa: E402
    BERNSTEIN_POINTS,
    GAUSS_LOBATTO_POINTS,
    UNIFORM_POINTS,
    OctantArray,
    OctForest,
    VecInterp,
)
