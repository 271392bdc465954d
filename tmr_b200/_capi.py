"""ctypes signatures for include/tmr_capi.h.

`bind(cdll)` decorates a loaded shared library with argument/return types.  It
is library-agnostic on purpose: the product package binds
tmr_b200/lib/libtmr_b200.so (CUDA); the test-suite binds the oracle build of
the same C file against the reference sources (oracle/_ref/libtmr_ref.so).
"""
import ctypes as C

import numpy as np

# byte-identical to TMROctant (reference src/TMROctant.h:49-53)
OCT_DTYPE = np.dtype(
    [
        ("block", "<i4"),
        ("x", "<i4"),
        ("y", "<i4"),
        ("z", "<i4"),
        ("tag", "<i4"),
        ("level", "<i2"),
        ("info", "<i2"),
    ],
    align=False,
)
assert OCT_DTYPE.itemsize == 24

P = C.c_void_p
I = C.c_int
PI = C.POINTER(C.c_int)
PPI = C.POINTER(C.POINTER(C.c_int))
PPD = C.POINTER(C.POINTER(C.c_double))

SIGNATURES = {
    "tmrc_backend": (C.c_char_p, []),
    "tmrc_forest_create": (P, [I, I]),
    "tmrc_forest_destroy": (None, [P]),
    "tmrc_set_connectivity": (None, [P, I, P, I]),
    "tmrc_set_mesh_order": (None, [P, I, I]),
    "tmrc_get_mesh_order": (I, [P]),
    "tmrc_get_interp_type": (I, [P]),
    "tmrc_repartition": (None, [P, I]),
    "tmrc_create_trees": (None, [P, I]),
    "tmrc_create_random_trees": (None, [P, I, I, I]),
    "tmrc_duplicate": (P, [P]),
    "tmrc_coarsen": (P, [P]),
    "tmrc_refine": (None, [P, P, I, I]),
    "tmrc_balance": (None, [P, I]),
    "tmrc_create_nodes": (None, [P]),
    "tmrc_num_octants": (I, [P]),
    "tmrc_get_octants": (None, [P, P]),
    "tmrc_write_octants": (None, [P, P, I]),
    "tmrc_get_node_conn": (None, [P, PPI, PI, PI]),
    "tmrc_get_dep_node_conn": (I, [P, PPI, PPI, PPD]),
    "tmrc_get_node_numbers": (I, [P, PPI]),
    "tmrc_get_owned_node_range": (I, [P, PPI]),
    "tmrc_get_ext_pre_offset": (I, [P]),
    "tmrc_get_local_node_number": (I, [P, I]),
    "tmrc_get_interp_knots": (I, [P, PPD]),
    "tmrc_eval_interp": (None, [P] + [P] * 11),
    "tmrc_get_connectivity": (None, [P, PI, PI, PI, PI, PPI, PPI, PPI, PPI]),
    "tmrc_get_inverse_connectivity": (None, [P, PPI, PPI, PPI, PPI, PPI, PPI]),
    "tmrc_transform_nodes": (None, [P, P, I, I, P, P]),
    "tmrc_find_enclosing": (None, [P, I, P, P, I, P, P]),
    "tmrc_distribute_octants": (I, [P, P, I, I, I, I, P, I, P, P]),
    "tmrc_send_octants": (I, [P, P, I, P, P, I, P, I]),
    "tmrc_interp_create": (P, []),
    "tmrc_interp_destroy": (None, [P]),
    "tmrc_create_interpolation": (None, [P, P, P]),
    "tmrc_interp_get": (None, [P, PI, PI, PPI, PPI, PPI, PPD]),
    "tmrc_array_sort": (I, [P, I, I]),
    "tmrc_array_contains": (None, [P, I, I, P, I, I, P]),
    "tmrc_array_merge": (I, [P, I, P, I, I, P, I]),
    "tmrc_queue_exercise": (I, [P, I, I, P, P]),
    "tmrc_hash_exercise": (I, [P, I, I, P, P, I]),
    "tmrc_forest_create_self": (P, [I, I]),
}


def bind(lib):
    """Attach restype/argtypes for every tmr_capi.h entry point."""
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return lib


def as_int_array(ptr, n):
    """Copy n ints out of a borrowed C pointer (None-safe)."""
    if n <= 0 or not ptr:
        return np.zeros(0, dtype=np.int32)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(np.int32, copy=True)


def view_array(ptr, n, dtype):
    """Borrowed numpy view of n items behind a C pointer: NO copy; valid as long
    as the owner keeps the array (None-safe)."""
    if n <= 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,))


def as_double_array(ptr, n):
    if n <= 0 or not ptr:
        return np.zeros(0, dtype=np.float64)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(np.float64, copy=True)
