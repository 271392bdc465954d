"""N>1 on the CPU: the multi-rank exchange logic (SFC partition, owner lookup,
all-to-all-v routing, ghost layer, node ownership and numbering) run as
thread-ranks of the test-only emulation and compared, rank by rank and
bit for bit, with the reference's own MPI path run as thread-ranks of the
oracle.  The GPU counterpart is tests/multi_gpu_check.py (torchrun)."""
import numpy as np
import pytest

import multirank
import util


@pytest.mark.parametrize("case", multirank.CASES, ids=[c[0] for c in multirank.CASES])
def test_multirank_matches_reference(case, emu_lib, ref_lib):
    name, conn_name, level, passes, pct, corner, order, ranks, repart = case
    conn = util.CONNS[conn_name]()
    body = multirank.adapt_body(conn, level, passes, pct, corner, order, repart)
    a = multirank.run_thread_ranks(ref_lib, ranks, body, True)
    b = multirank.run_thread_ranks(emu_lib, ranks, body, False)
    multirank.compare_rank_results(a, b, name)


@pytest.mark.parametrize("order,ranks,mode", [(2, 2, True), (2, 4, True), (3, 3, True),
                                              (2, 3, "repartitioned"),
                                              (3, 2, "repartitioned")])
def test_multirank_interpolation(order, ranks, mode, emu_lib, ref_lib):
    """createInterpolation with rows whose enclosing coarse element lives on
    another rank (reference src/TMROctForest.cpp:6699-6783)."""
    conn = util.box_conn()
    body = multirank.adapt_body(conn, 1, 3, 30, 1, order, True, with_interp=mode)
    a = multirank.run_thread_ranks(ref_lib, ranks, body, True)
    b = multirank.run_thread_ranks(emu_lib, ranks, body, False)
    multirank.compare_rank_results(a, b, "interp o%d r%d" % (order, ranks))
    assert sum(len(x[1]["interp"]) for x in b) > 0


def test_multirank_interpolation_uneven_depths(emu_lib, ref_lib):
    """A forest that is never repartitioned keeps whole trees per rank, so the
    ranks' deepest levels differ; reading the octants after createNodes makes
    the host class upload them again.  The keys a rank ships for remote
    prolongation rows must still be at the depth every rank agreed on in
    createNodes (found by the seeded fuzz runs: grid2, 5 ranks)."""
    conn = util.CONNS["grid2"]()
    body = multirank.adapt_body(conn, 1, 3, 30, 1, 2, False, seed=738575,
                                with_interp="repartitioned")
    a = multirank.run_thread_ranks(ref_lib, 5, body, True)
    b = multirank.run_thread_ranks(emu_lib, 5, body, False)
    multirank.compare_rank_results(a, b, "uneven depths")


def test_multirank_repartition_when_only_some_ranks_hold_info(emu_lib, ref_lib):
    """16 elements on 8 ranks: after createNodes and a read of the octants some
    ranks hold an `info` array and some do not; repartitioning the duplicate must
    still run the SAME collectives on every rank (it used to exchange `info` only
    where an array existed -- a deadlock under NCCL; the emulated communicator
    aborts when the ranks' collective sequences diverge).  Found by the fuzz
    runs."""
    conn = util.CONNS["rectangle"]()
    body = multirank.adapt_body(conn, 1, 1, 15, 0, 3, True, seed=673287,
                                with_interp="repartitioned")
    a = multirank.run_thread_ranks(ref_lib, 8, body, True)
    b = multirank.run_thread_ranks(emu_lib, 8, body, False)
    multirank.compare_rank_results(a, b, "partial info")


@pytest.mark.parametrize("ranks", [2, 3])
def test_multirank_bernstein_order3(ranks, emu_lib, ref_lib):
    """Labelled node keys (order-3 Bernstein points) through the multi-rank
    node ownership, external numbering and remote interpolation rows."""
    conn = util.box_conn()
    body = multirank.adapt_body(conn, 1, 2, 30, 1, 3, True, with_interp="repartitioned",
                                interp=2)
    a = multirank.run_thread_ranks(ref_lib, ranks, body, True)
    b = multirank.run_thread_ranks(emu_lib, ranks, body, False)
    multirank.compare_rank_results(a, b, "bernstein3_r%d" % ranks)


@pytest.mark.parametrize("ranks,interp", [(2, 1), (3, 2)])
def test_multirank_order4(ranks, interp, emu_lib, ref_lib):
    """Order 4 on several ranks: entity owners expanded to nodes, external
    numbers requested per (entity, sub-node) (reference :4139-4232), remote
    prolongation rows."""
    conn = util.box_conn()
    body = multirank.adapt_body(conn, 1, 2, 30, 1, 4, True, with_interp="repartitioned",
                                interp=interp)
    a = multirank.run_thread_ranks(ref_lib, ranks, body, True)
    b = multirank.run_thread_ranks(emu_lib, ranks, body, False)
    multirank.compare_rank_results(a, b, "order4_r%d" % ranks)


def test_public_distribute_octants(emu_lib, ref_lib):
    """distributeOctants / sendOctants as external callers use them (reference
    src/topology/TMR_TACSTopoCreator.cpp:166-217): route a sorted octant list
    to the owners of the positions and send a reply back."""
    conn = util.box_conn()
    ranks = 3

    def body(lib, rank):
        from tmr_b200.forest import OctForest, array_sort

        f = OctForest(lib=lib)
        f.setConnectivity(conn)
        f.createTrees(2)
        f.repartition()
        rng = np.random.default_rng(100 + rank)
        mine = util.random_octants(rng, 200, 7, 4)
        mine["tag"] = rank
        mine = array_sort(lib, mine, 0)
        got, optr, rptr = f.distributeOctants(mine, ranks)
        got2, _, _ = f.distributeOctants(mine, ranks, include_local=1)
        back = f.sendOctants(got, rptr, optr)
        return got, optr, rptr, got2, back

    a = multirank.run_thread_ranks(ref_lib, ranks, body, True)
    b = multirank.run_thread_ranks(emu_lib, ranks, body, False)
    for r in range(ranks):
        for k in (1, 2):
            assert np.array_equal(a[r][k], b[r][k]), (r, k)
        for k in (0, 3):
            util.assert_octants_equal(a[r][k], b[r][k], "rank %d list %d" % (r, k))
        # the reply lands in the slots of the original send intervals
        ra, rb = a[r][4], b[r][4]
        assert len(ra) == len(rb)
        optr = a[r][1]
        for i in range(ranks):
            if i != r:
                util.assert_octants_equal(ra[optr[i]:optr[i + 1]], rb[optr[i]:optr[i + 1]],
                                          "reply rank %d from %d" % (r, i))


def test_rank_count_invariance(emu_lib):
    """The balanced octant set does not depend on the number of ranks."""
    conn = util.box_conn()
    sums = []
    for ranks in (1, 2, 5):
        body = multirank.adapt_body(conn, 1, 3, 30, 1, 2, True)
        out = multirank.run_thread_ranks(emu_lib, ranks, body, False)
        allocts = np.concatenate([o[0][-1] for o in out])
        sums.append((len(allocts), util.checksum(allocts)))
    assert sums[0] == sums[1] == sums[2]


def test_partition_counts():
    from tmr_b200 import dist

    assert dist.partition_counts(10, 4) == [3, 3, 2, 2]
    assert dist.partition_counts(10, 4, 2) == [5, 5, 0, 0]
    assert dist.partition_counts(3, 8) == [1, 1, 1, 0, 0, 0, 0, 0]
    assert sum(dist.partition_counts(86278900, 8)) == 86278900


def test_serial_forest_inside_multirank_job(emu_lib, ref_lib):
    """A TMROctForest built on MPI_COMM_SELF while the process belongs to a
    multi-rank world is not partitioned: every rank gets the full single-rank
    result (reference src/TMROctForest.cpp:331-337 takes rank and size from
    the communicator it is given)."""
    from tmr_b200.forest import OctForest

    conn = util.box_conn()

    def body(lib, rank):
        # a partitioned forest first, so the world communicator is in use
        w = OctForest(order=2, lib=lib)
        w.setConnectivity(conn)
        w.createTrees(1)
        w.repartition()
        f = OctForest(order=2, lib=lib, comm_self=True)
        f.setConnectivity(conn)
        f.createTrees(1)
        f.repartition()
        rec = []
        for p in range(2):
            o = f.getOctants().as_array()
            f.refine(util.synth_flags(o, 2024 + p, 30))
            f.balance(1)
            rec.append(f.getOctants().as_array().copy())
        res = util.node_results(f)
        c = f.coarsen()
        c.balance(1)
        rows, d = util.interp_rows(f.createInterpolation(c))
        res["interp"] = {r: (cc.copy(), ww.copy()) for r, (cc, ww) in d.items()}
        return rec, res

    one = multirank.run_thread_ranks(ref_lib, 1, body, True)
    for ranks in (2, 3):
        got = multirank.run_thread_ranks(emu_lib, ranks, body, False)
        multirank.compare_rank_results(one * ranks, got, "serial in %d-rank job" % ranks)


@pytest.mark.parametrize("ranks", [2, 3])
def test_find_enclosing_names_the_owner_of_a_miss(ranks, emu_lib, ref_lib):
    """findEnclosing for nodes this rank does NOT hold: no element, and
    mpi_owner = the rank owning the node's position (reference
    src/TMROctForest.cpp:6348-6372; src/topology/TMR_TACSTopoCreator.cpp:166-217
    routes on that value).  Every rank asks about every element corner of the
    whole forest."""
    conn = util.box_conn()
    body = multirank.find_enclosing_body(conn, multirank.find_enclosing_queries(ref_lib, conn))
    a = multirank.run_thread_ranks(ref_lib, ranks, body, True)
    b = multirank.run_thread_ranks(emu_lib, ranks, body, False)
    multirank.compare_find_enclosing(a, b, ranks)


@pytest.mark.parametrize("ranks", [2, 3])
def test_multirank_device_views(ranks, emu_lib):
    """Several ranks: the node arrays rebuilt on request from the slot state
    (node numbers by local node, tmrgpu_assembler_views) agree with what the
    getters return, which test_multirank_matches_reference pins to the oracle."""
    conn = util.box_conn()

    def body(lib, rank):
        f = multirank.OctForest(order=2, lib=lib)
        f.setConnectivity(conn)
        f.createTrees(1)
        f.repartition()
        for p in range(2):
            f.refine(util.synth_flags(f.getOctants().as_array(), 2024 + p, 30))
            f.balance(1)
            f.repartition()
        res = util.node_results(f)
        v = f.assemblerViews()
        assert np.array_equal(np.sort(v["node_numbers"]), res["node_numbers"])
        assert np.array_equal(v["conn"].reshape(-1, 8), res["conn"].reshape(-1, 8))
        assert np.array_equal(v["dep_ptr"], res["dep"][0])
        assert np.array_equal(v["dep_conn"], res["dep"][1])
        return len(v["node_numbers"])

    out = multirank.run_thread_ranks(emu_lib, ranks, body, False)
    assert all(n > 0 for n in out)


@pytest.mark.parametrize("ranks", [2, 3])
def test_multirank_node_locations(ranks, emu_lib, ref_lib):
    """getPoints on several ranks (evaluateNodeLocations, reference
    src/TMROctForest.cpp:5524-5675): each rank's X array, in the order of its
    own sorted node numbers, against the same rank of the reference."""
    conn = util.box_conn()
    n = int(np.max(conn)) + 1
    xpts = (np.random.default_rng(3).uniform(-1, 1, (n, 3)) +
            2.5 * np.arange(n)[:, None] * np.array([1.0, -0.4, 0.25]))

    def body(lib, rank):
        f = multirank.OctForest(order=2, lib=lib)
        f.setTrilinearTopology(conn, xpts)
        f.createTrees(1)
        f.repartition()
        for p in range(2):
            f.refine(util.synth_flags(f.getOctants().as_array(), 2024 + p, 30))
            f.balance(0)
            f.repartition()
        f.createNodes()
        return f.getNodeNumbers().copy(), f.getPoints()

    a = multirank.run_thread_ranks(ref_lib, ranks, body, True)
    b = multirank.run_thread_ranks(emu_lib, ranks, body, False)
    for r in range(ranks):
        assert np.array_equal(a[r][0], b[r][0]), r
        assert len(a[r][1]) == len(a[r][0]) > 0
        assert np.array_equal(a[r][1], b[r][1]), r



@pytest.mark.parametrize("ranks,order", [(2, 2), (3, 3)])
def test_multirank_name_queries(ranks, order, emu_lib, ref_lib):
    """getOctsWithName / getNodesWithName answer for the LOCAL octants and
    nodes of each rank (reference src/TMROctForest.cpp:5747-5862, 5882-6203)."""
    import random
    conn = util.box_conn()
    xpts = np.random.default_rng(3).uniform(-1, 1, (int(np.max(conn)) + 1, 3))

    def body(lib, rank):
        f = multirank.OctForest(order=order, lib=lib)
        f.setTrilinearTopology(conn, xpts)
        c = f.getConnectivity()
        r = random.Random(5)
        for kind, count in ((0, c["nnodes"]), (1, c["nedges"]), (2, c["nfaces"]),
                            (3, c["nblocks"])):
            for i in range(count):
                name = r.choice([None, "fixed", "free"])
                if name:
                    f.setEntityName(kind, i, name)
        f.createTrees(1)
        f.repartition()
        for p in range(2):
            f.refine(util.synth_flags(f.getOctants().as_array(), 2024 + p, 30))
            f.balance(1)
            f.repartition()
        f.createNodes()
        return (f.getOctsWithName("fixed"), f.getOctsWithName("free"),
                f.getNodesWithName("fixed"), f.getNodesWithName("free"))

    a = multirank.run_thread_ranks(ref_lib, ranks, body, True)
    b = multirank.run_thread_ranks(emu_lib, ranks, body, False)
    total = 0
    for r in range(ranks):
        for x, y in zip(a[r], b[r]):
            assert np.array_equal(x, y), r
        total += len(b[r][2])
    assert total > 0
