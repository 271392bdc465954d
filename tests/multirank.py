"""Helpers to run one body on R ranks: oracle thread-ranks (MPI shim), the
test-only emulation (threads + in-process communicator), or real GPUs (one
process per GPU under torchrun -- see tests/multi_gpu_check.py)."""
import threading

import numpy as np

import util
from tmr_b200 import dist
from tmr_b200.forest import OctForest


def run_thread_ranks(lib, nranks, body, is_reference):
    out = [None] * nranks
    errs = []
    if is_reference:
        lib.shim_world_begin(nranks)
        uid = None
    else:
        uid = dist.make_unique_id(lib)

    def th(r):
        try:
            if is_reference:
                lib.shim_attach(r)
            else:
                dist.init_world(lib, r, nranks, uid)
            out[r] = body(lib, r)
        except Exception as e:  # noqa: BLE001
            import traceback

            traceback.print_exc()
            errs.append(e)

    threads = [threading.Thread(target=th, args=(r,)) for r in range(nranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if is_reference:
        lib.shim_world_end()
    assert not errs, errs
    return out


def adapt_body(conn, level, passes, pct, corner, order, repartition=True, seed=2024,
               with_interp=False, interp=1):
    """createTrees -> repartition -> passes x {refine, balance, repartition}
    -> createNodes; returns (per-stage octant arrays, node results)."""

    def body(lib, rank):
        f = OctForest(order=order, interp=interp, lib=lib)
        f.setConnectivity(conn)
        f.createTrees(level)
        if repartition:
            f.repartition()
        rec = []
        for p in range(passes):
            o = f.getOctants().as_array()
            f.refine(util.synth_flags(o, seed + p, pct))
            f.balance(corner)
            rec.append(f.getOctants().as_array().copy())
            if repartition:
                f.repartition()
                rec.append(f.getOctants().as_array().copy())
        res = util.node_results(f)
        if with_interp:
            # TopOptUtils-style coarse level (reference tmr/TopOptUtils.py:79-99)
            if order > 2:
                coarse = f.duplicate()
                coarse.setMeshOrder(order - 1, interp)
            else:
                coarse = f.coarsen()
                coarse.balance(1)
            if with_interp == "repartitioned":
                # a coarse partition that is NOT aligned with the fine one, so
                # that some rows must be computed by another rank
                coarse.repartition()
            rows, d = util.interp_rows(f.createInterpolation(coarse))
            res["interp"] = {r: (c.copy(), w.copy()) for r, (c, w) in d.items()}
        return rec, res

    return body


CASES = [
    # name, conn, level, passes, pct, corner, order, ranks, repartition
    ("box7_r2", "box7", 1, 2, 30, 0, 2, 2, True),
    ("box7_r3_corner", "box7", 1, 3, 30, 1, 2, 3, True),
    ("single_r4", "single", 2, 3, 30, 0, 2, 4, True),
    ("connector15_r2_order3", "connector15", 1, 2, 30, 0, 3, 2, True),
    ("butterfly2_r4", "butterfly2", 1, 2, 30, 1, 2, 4, True),
    ("grid2_r8", "grid2", 1, 2, 35, 0, 2, 8, True),
    ("box7_r2_norepart", "box7", 1, 2, 30, 0, 2, 2, False),
    ("single_r3_more_ranks_than_trees", "single", 1, 3, 40, 0, 2, 3, True),
]


def compare_rank_results(a, b, what):
    for r, ((ra, na), (rb, nb)) in enumerate(zip(a, b)):
        assert len(ra) == len(rb)
        for k, (x, y) in enumerate(zip(ra, rb)):
            util.assert_octants_equal(x, y, "%s rank %d stage %d" % (what, r, k))
        util.assert_nodes_equal(na, nb, "%s rank %d nodes" % (what, r))
        if "interp" in na:
            # rows a rank emits (its own + those it computes for other ranks);
            # the order of received rows is unspecified in the reference (qsort)
            ia, ib = na["interp"], nb["interp"]
            assert sorted(ia) == sorted(ib), "%s rank %d interp rows differ" % (what, r)
            for row in ia:
                assert (ia[row][0] == ib[row][0]).all(), (what, r, row)
                assert abs(ia[row][1] - ib[row][1]).max() <= 1e-12 * abs(ia[row][1]).max()


def find_enclosing_body(conn, queries):
    """every rank asks findEnclosing about every record of `queries` (element
    records of the whole forest with info = a local node index): hits name the
    local element, misses the owner rank of the node's position (reference
    src/TMROctForest.cpp:6348-6372)"""
    knots = np.array([-1.0, 1.0])

    def body(lib, rank):
        f = OctForest(order=2, lib=lib)
        f.setConnectivity(conn)
        f.createTrees(1)
        f.repartition()
        for p in range(2):
            o = f.getOctants().as_array()
            f.refine(util.synth_flags(o, 2024 + p, 30))
            f.balance(0)
            f.repartition()
        f.createNodes()
        idx, own = f.findEnclosing(2, knots, queries)
        return idx.copy(), own.copy()

    return body


def find_enclosing_queries(ref_lib, conn):
    octs = util.build_forest(ref_lib, conn, 1, 2, 30, 0).getOctants().as_array().copy()
    octs["info"] = np.random.default_rng(5).integers(0, 8, len(octs))
    return octs


def compare_find_enclosing(a, b, ranks):
    misses = 0
    for r in range(ranks):
        hit = a[r][0] >= 0
        assert np.array_equal(hit, b[r][0] >= 0), "rank %d: hit/miss differs" % r
        assert np.array_equal(a[r][0][hit], b[r][0][hit]), "rank %d: element index" % r
        assert np.array_equal(a[r][1], b[r][1]), "rank %d: mpi_owner" % r
        misses += int((~hit).sum())
        assert set(np.unique(b[r][1][~hit]).tolist()) <= set(range(ranks)) - {r}
    assert misses > 0
    return misses
