"""The hand-written device primitives against numpy: onesweep radix sort
(keys and pairs, partial bit ranges, stability) and the chained scan."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def prim(gpu_lib):
    lib = gpu_lib
    lib.tmr_b200_context.restype = ctypes.c_void_p
    ctx = ctypes.c_void_p(lib.tmr_b200_context())
    lib.tmrgpu_test_radix_sort.argtypes = [ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.c_int, ctypes.c_int]
    lib.tmrgpu_test_scan.argtypes = [ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_int64, ctypes.c_void_p,
                                     ctypes.c_void_p]
    return lib, ctx


@pytest.mark.parametrize("n", [1, 2, 31, 4096, 4097, 100003, 3_000_000])
@pytest.mark.parametrize("bits", [(0, 64), (0, 8), (5, 44), (3, 20), (0, 33)])
def test_radix_sort_keys(prim, n, bits):
    lib, ctx = prim
    lo, hi = bits
    rng = np.random.default_rng(n + lo + hi)
    keys = rng.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n, dtype=np.uint64)
    mask = np.uint64(((1 << (hi - lo)) - 1) << lo) if hi - lo < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
    order = np.argsort(keys & mask, kind="stable")
    expect = keys[order]
    got = keys.copy()
    assert lib.tmrgpu_test_radix_sort(ctx, got.ctypes.data, None, n, lo, hi) == 0
    assert np.array_equal(got, expect)


@pytest.mark.parametrize("n", [5, 4096, 70001, 2_500_000])
def test_radix_sort_pairs_stable(prim, n):
    lib, ctx = prim
    rng = np.random.default_rng(n)
    # few distinct keys => long equal runs => stability is observable
    keys = rng.integers(0, 37, n, dtype=np.uint64) << np.uint64(9)
    vals = np.arange(n, dtype=np.uint32)
    order = np.argsort(keys, kind="stable")
    k, v = keys.copy(), vals.copy()
    assert lib.tmrgpu_test_radix_sort(ctx, k.ctypes.data, v.ctypes.data, n, 0, 20) == 0
    assert np.array_equal(k, keys[order])
    assert np.array_equal(v, vals[order])


def test_radix_sort_sorted_and_constant(prim):
    lib, ctx = prim
    n = 1_000_000
    for keys in (np.arange(n, dtype=np.uint64), np.full(n, 12345, dtype=np.uint64),
                 np.arange(n, dtype=np.uint64)[::-1].copy()):
        got = keys.copy()
        assert lib.tmrgpu_test_radix_sort(ctx, got.ctypes.data, None, n, 0, 24) == 0
        assert np.array_equal(got, np.sort(keys))


@pytest.mark.parametrize("n", [1, 7, 2048, 2049, 1_000_001, 20_000_000])
def test_scan_counts(prim, n):
    lib, ctx = prim
    rng = np.random.default_rng(n)
    c = rng.integers(0, 9, n, dtype=np.uint32)
    out = np.zeros(n, dtype=np.uint32)
    total = ctypes.c_uint64(0)
    assert lib.tmrgpu_test_scan(ctx, c.ctypes.data, n, out.ctypes.data, ctypes.byref(total)) == 0
    ex = np.cumsum(c, dtype=np.uint64) - c
    assert total.value == int(c.sum(dtype=np.uint64))
    assert np.array_equal(out, ex.astype(np.uint32))


def test_allocation_failure_is_contained(prim, ref_lib):
    """A device allocation that fails in the middle of an operation (injected)
    must not launch a kernel on the missing buffer -- an illegal address would
    destroy the process's CUDA context -- the operation reports an error, and
    the NEXT operation on the same context runs normally and gives the
    reference's result."""
    import util
    from tmr_b200.forest import OctForest

    lib, ctx = prim
    lib.tmrgpu_test_fail_alloc.argtypes = [ctypes.c_void_p, ctypes.c_long]
    conn = util.box_conn()
    want = util.node_results(util.build_forest(ref_lib, conn, 1, 2, 30, 1, 2))
    for nth in (1, 2, 5, 9, 17, 33):
        f = OctForest(order=2, lib=lib)
        f.setConnectivity(conn)
        f.createTrees(1)
        o = f.getOctants().as_array()
        f.refine(util.synth_flags(o, 2024, 30))
        lib.tmrgpu_test_fail_alloc(ctx, nth)
        f.balance(1)          # hits the injected failure (prints an error)
        f.createNodes()       # may hit it too if balance allocated less
        lib.tmrgpu_test_fail_alloc(ctx, 0)
        # the context is still alive and clean: a fresh forest gives the right answer
        got = util.node_results(util.build_forest(lib, conn, 1, 2, 30, 1, 2))
        util.assert_nodes_equal(want, got, "after injected failure %d" % nth)
