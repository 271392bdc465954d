"""TEST INFRASTRUCTURE ONLY -- loads oracle/_ref/libtmr_ref.so (the unmodified
reference sources + shims, built by oracle/Makefile) and binds the same
tmr_capi.h signatures the product library exports.  Only tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke() may import
this module."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libtmr_ref.so")


def available():
    return os.path.exists(REF_LIB)


def load():
    import sys

    root = os.path.dirname(HERE)
    if root not in sys.path:
        sys.path.insert(0, root)
    from tmr_b200 import _capi

    lib = _capi.bind(ctypes.CDLL(REF_LIB))
    for name in ("shim_world_begin", "shim_attach", "shim_world_end"):
        getattr(lib, name)
    lib.shim_world_begin.argtypes = [ctypes.c_int]
    lib.shim_attach.argtypes = [ctypes.c_int]
    return lib
