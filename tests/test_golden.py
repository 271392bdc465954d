"""Committed golden vectors (tests/golden/*.npz, generated from the oracle by
tests/golden/make_golden.py) and the fingerprints pinned in BASELINE.md.

 * the oracle must reproduce them (pins the oracle build itself),
 * the kernel bodies (emu, CPU) and the CUDA library (gpu) must reproduce them
   without the oracle present -- this is what runs when /root/reference is
   absent."""
import glob
import json
import os

import numpy as np
import pytest

import util
from tmr_b200.forest import OctForest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(glob.glob(os.path.join(GOLD, "*.npz")))


@pytest.fixture(params=["ref", "emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def lib(request):
    return request.getfixturevalue(request.param + "_lib")


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p)[:-4] for p in FILES])
def test_golden_case(path, lib):
    g = np.load(path)
    level, passes, pct, corner, order = [int(v) for v in g["params"]]
    interp = int(g["interp_type"]) if "interp_type" in g.files else 1
    rec = []
    f = util.build_forest(lib, g["block_conn"], level, passes, pct, corner, order,
                          interp=interp, record=rec)
    assert [len(r[1]) for r in rec] == list(g["counts"])
    assert [util.checksum(r[1]) for r in rec] == [int(c) for c in g["checksums"]]
    res = util.node_results(f)
    util.assert_octants_equal(g["octants"], res["octants"], "octants")
    assert np.array_equal(res["conn"], g["conn"])
    assert np.array_equal(res["node_numbers"], g["node_numbers"])
    assert np.array_equal(res["dep"][0], g["dep_ptr"])
    assert np.array_equal(res["dep"][1], g["dep_conn"])
    np.testing.assert_allclose(res["dep"][2], g["dep_weights"], rtol=1e-12, atol=0)
    coarse = f.coarsen() if order == 2 else f.duplicate()
    if order == 2:
        coarse.balance(1)
    else:
        coarse.setMeshOrder(2, interp)
    rows, rowp, cols, vals = f.createInterpolation(coarse).get()
    assert np.array_equal(rows, g["interp_rows"])
    assert np.array_equal(rowp, g["interp_rowp"])
    assert np.array_equal(cols, g["interp_cols"])
    np.testing.assert_allclose(vals, g["interp_vals"], rtol=1e-12, atol=1e-300)


def test_pinned_fingerprint_c1(lib):
    """1 tree, createTrees(4), 4 passes pct 30 ->
    1,027,916 octants, checksum 55e9487c98c7a2ff, 652,025 owned nodes,
    908,576 dependent nodes with 2,441,728 stencil entries."""
    with open(os.path.join(GOLD, "fingerprints.json")) as fh:
        fp = json.load(fh)["C1"]
    rec = []
    f = util.build_forest(lib, util.single_conn(), 4, 4, 30, 0, 2, record=rec)
    counts = [len(r[1]) for r in rec if r[0].startswith("balance")]
    assert counts == fp["counts"]
    assert "%016x" % util.checksum(rec[-1][1]) == fp["checksum"]
    f.createNodes()
    ptr, conn, w = f.getDepNodeConn()
    assert f.getNumOwnedNodes() == fp["owned_nodes"]
    assert len(ptr) - 1 == fp["dep_nodes"]
    assert len(conn) == fp["dep_nnz"]
    assert len(f.getNodeNumbers()) == fp["local_nodes"]
    # size-independent structure: every dependent stencil sums to 1
    sums = np.add.reduceat(w, ptr[:-1])
    assert np.abs(sums - 1.0).max() < 1e-14


def test_pinned_fingerprint_c4_1050_trees(lib):
    """BASELINE.json configs[3] connectivity in full: the 5x5x6 lattice of
    7-tree butterfly cells (1050 trees, every face orientation id, irregular
    edge valence), createTrees(1), 2 passes pct 30, balance(1), order 2 --
    fingerprint computed on the oracle by tests/golden/make_golden.py."""
    with open(os.path.join(GOLD, "fingerprints.json")) as fh:
        fp = json.load(fh)["C4_1050_trees"]
    conn = util.butterfly_conn(5, 5, 6)
    assert len(conn) == fp["trees"]
    rec = []
    f = util.build_forest(lib, conn, 1, 2, 30, 1, 2, record=rec)
    assert [len(r[1]) for r in rec if r[0].startswith("balance")] == fp["counts"]
    res = util.node_results(f)
    assert "%016x" % util.checksum(res["octants"]) == fp["checksum"]
    assert f.getNumOwnedNodes() == fp["owned_nodes"]
    assert len(res["dep"][0]) - 1 == fp["dep_nodes"]
    assert len(res["dep"][1]) == fp["dep_nnz"]
    assert len(res["node_numbers"]) == fp["local_nodes"]
    cs = int(np.sum(res["conn"].astype(np.int64).ravel() *
                    (np.arange(res["conn"].size) % 1000003 + 1))) & 0xFFFFFFFFFFFFFFFF
    assert "%016x" % cs == fp["conn_checksum"]
