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
        _LIB = _capi.bind(ctypes.CDLL(LIB_PATH))
    return _LIB


def require_gpu():
    """Create the process-wide CUDA context now; raise if there is no GPU."""
    lib = load_library()
    lib.tmr_b200_context.restype = ctypes.c_void_p
    if not lib.tmr_b200_context():
        raise RuntimeError(
            "tmr_b200: no usable CUDA device (the library has no CPU fallback)")
    return lib


def use_stream(stream_handle):
    """Run all forest work on this cudaStream_t (e.g. torch's current stream:
    torch.cuda.current_stream().cuda_stream).  Call before the first forest."""
    lib = load_library()
    lib.tmr_b200_use_stream.argtypes = [ctypes.c_void_p]
    lib.tmr_b200_use_stream(ctypes.c_void_p(stream_handle))


from .forest import (  # noqa: E402
    BERNSTEIN_POINTS,
    GAUSS_LOBATTO_POINTS,
    UNIFORM_POINTS,
    OctantArray,
    OctForest,
    VecInterp,
)
