"""Full-size checks on the GPU (BASELINE.json configs[1], "C2": 8x8x8 trees,
createTrees(3), 4 hash-driven passes pct 35 -> 86,278,900 octants).

The oracle needs minutes and 17 GB for this size, so the test does not run it.
It checks
 * the fingerprints the unmodified reference produced for this exact recipe at
   survey time (tests/golden/fingerprints.json: octant count, record checksum,
   owned nodes), and
 * size-independent properties of the result: every tree is covered exactly
   once (sum of 8^-level per tree == 1), the array is strictly Morton-sorted,
   balance and a zero-flag refine are idempotent, connectivity indices are in
   range, dep_ptr is a prefix sum and every dependent stencil sums to 1.
"""
import ctypes
import json
import os

import numpy as np
import pytest

import util
from tmr_b200.forest import OctForest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
P, I64 = ctypes.c_void_p, ctypes.c_int64


def _bind(lib):
    lib.tmr_b200_context.restype = P
    lib.tmr_b200_device_forest.restype = P
    lib.tmr_b200_device_forest.argtypes = [P]
    lib.tmrgpu_count.restype = I64
    lib.tmrgpu_count.argtypes = [P]
    lib.tmrgpu_synth_flags.argtypes = [P, ctypes.c_uint64, ctypes.c_int, P]
    lib.tmrgpu_refine_device.argtypes = [P, P, ctypes.c_int, ctypes.c_int]
    lib.tmrgpu_balance.argtypes = [P, ctypes.c_int]
    lib.tmrgpu_dev_alloc.argtypes = [P, I64, ctypes.POINTER(P)]
    lib.tmrgpu_dev_free.argtypes = [P, P]
    lib.tmrgpu_checksum.argtypes = [P, ctypes.POINTER(ctypes.c_uint64)]
    return P(lib.tmr_b200_context())


def _checksum(lib, dev):
    c = ctypes.c_uint64(0)
    assert lib.tmrgpu_checksum(dev, ctypes.byref(c)) == 0
    return "%016x" % c.value


def test_c2_full_size(gpu_lib):
    lib = gpu_lib
    ctx = _bind(lib)
    with open(os.path.join(GOLD, "fingerprints.json")) as fh:
        fp = json.load(fh)["C2"]
    f = OctForest(order=2, lib=lib)
    f.setConnectivity(util.structured_conn(8))
    f.createTrees(3)
    dev = P(lib.tmr_b200_device_forest(f._ptr))
    for p in range(4):
        n = lib.tmrgpu_count(dev)
        buf = P()
        assert lib.tmrgpu_dev_alloc(ctx, 4 * n, ctypes.byref(buf)) == 0
        lib.tmrgpu_synth_flags(dev, 2024 + p, 35, buf)
        assert lib.tmrgpu_refine_device(dev, buf, 0, 30) == 0
        assert lib.tmrgpu_balance(dev, 0) == 0
        lib.tmrgpu_dev_free(ctx, buf)

    # fingerprints of the unmodified reference for this recipe
    n = lib.tmrgpu_count(dev)
    assert n == fp["octants"]
    assert _checksum(lib, dev) == fp["checksum"]

    # idempotence: a balanced forest stays as it is
    assert lib.tmrgpu_balance(dev, 0) == 0
    assert lib.tmrgpu_count(dev) == n and _checksum(lib, dev) == fp["checksum"]
    zero = f.duplicate()
    zdev = P(lib.tmr_b200_device_forest(zero._ptr))
    buf = P()
    assert lib.tmrgpu_dev_alloc(ctx, 4 * n, ctypes.byref(buf)) == 0
    lib.tmrgpu_synth_flags(zdev, 1, 0, buf)  # pct 0: every flag is 0
    assert lib.tmrgpu_refine_device(zdev, buf, 0, 30) == 0
    lib.tmrgpu_dev_free(ctx, buf)
    assert lib.tmrgpu_count(zdev) == n and _checksum(lib, zdev) == fp["checksum"]
    del zero

    # every tree covered exactly once; strictly sorted in the reference's order
    octs = f.getOctants().as_array()
    lv = octs["level"].astype(np.int64)
    lmax = int(lv.max())
    vol = np.left_shift(np.int64(1), 3 * (lmax - lv))
    per_tree = np.bincount(octs["block"], weights=vol.astype(np.float64), minlength=512)
    assert np.all(per_tree == float(8 ** lmax))
    assert np.array_equal(octs["tag"], np.arange(n, dtype=np.int32))
    blk = octs["block"].astype(np.int64)
    assert np.all(np.diff(blk) >= 0)
    # x-major Morton order inside a tree (reference src/TMROctant.cpp:171-204):
    # compare consecutive records through the most significant differing bit
    same = np.diff(blk) == 0
    a, b = octs[:-1], octs[1:]
    dx = (a["x"] ^ b["x"]).astype(np.int64)
    dy = (a["y"] ^ b["y"]).astype(np.int64)
    dz = (a["z"] ^ b["z"]).astype(np.int64)
    top = np.maximum(np.maximum(dx, dy), dz)
    # x wins ties with y and z, y wins ties with z: the deciding axis is the
    # first whose xor has the same leading bit as the maximum
    def lead(v):
        return np.where(v > 0, np.floor(np.log2(np.maximum(v, 1))).astype(np.int64), -1)
    lt = lead(top)
    use_x = lead(dx) == lt
    use_y = ~use_x & (lead(dy) == lt)
    less = np.where(use_x, a["x"] < b["x"], np.where(use_y, a["y"] < b["y"], a["z"] < b["z"]))
    assert np.all(less[same & (top > 0)])
    assert not np.any(same & (top == 0))  # no two leaves share an anchor
    del octs, a, b, dx, dy, dz, top, lt, use_x, use_y, less, vol, lv

    # nodes
    f.createNodes()
    assert f.getNumOwnedNodes() == fp["owned_nodes"]
    conn = f.getMeshConn()
    ptr, dconn, w = f.getDepNodeConn()
    nd = len(ptr) - 1
    assert conn.shape[0] == n
    assert int(conn.min()) >= -nd and int(conn.max()) < fp["owned_nodes"]
    assert ptr[0] == 0 and ptr[-1] == len(dconn) == len(w)
    assert np.all(np.diff(ptr) > 0)
    assert int(dconn.min()) >= 0 and int(dconn.max()) < fp["owned_nodes"]
    sums = np.add.reduceat(w, ptr[:-1])
    assert np.abs(sums - 1.0).max() < 1e-14
