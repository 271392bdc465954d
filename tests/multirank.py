"""Helpers to run one body on R ranks: oracle thread-ranks (MPI shim), the
test-only emulation (threads + in-process communicator), or real GPUs (one
process per GPU under torchrun -- see tests/multi_gpu_check.py)."""
import threading

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


def adapt_body(conn, level, passes, pct, corner, order, repartition=True, seed=2024):
    """createTrees -> repartition -> passes x {refine, balance, repartition}
    -> createNodes; returns (per-stage octant arrays, node results)."""

    def body(lib, rank):
        f = OctForest(order=order, lib=lib)
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
        return rec, util.node_results(f)

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
