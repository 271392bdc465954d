"""Seeded random configurations (super-mesh, start level, passes, refinement
percentage, corner balance, mesh order, rank count) of the adaptation pipeline:
the test-only emulation of the kernel bodies against the oracle, every array
bit for bit (weights 1e-12).  The fixed cases of test_parity / test_multirank
were chosen by hand; these were not."""
import random

import numpy as np
import pytest

import multirank
import util

NAMES = ["single", "box7", "connector15", "grid2", "butterfly2", "rectangle"]


def _single_rank_cases(n, seed):
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        cn = rng.choice(NAMES)
        level = rng.choice([0, 1, 1, 2])
        passes = rng.choice([1, 2, 3])
        if cn in ("butterfly2", "grid2") and level == 2 and passes == 3:
            passes = 2
        out.append((cn, level, passes, rng.choice([10, 30, 50, 70]), rng.choice([0, 1]),
                    rng.choice([2, 2, 2, 3, 4]), rng.randrange(1, 10 ** 6)))
    return out


def _multi_rank_cases(n, seed):
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        cn = rng.choice(["box7", "connector15", "grid2", "butterfly2"])
        out.append((cn, 1, rng.choice([1, 2]), rng.choice([20, 40, 60]), rng.choice([0, 1]),
                    rng.choice([2, 2, 3]), rng.choice([2, 3, 4]), rng.choice([True, True, False]),
                    rng.randrange(1, 10 ** 6)))
    return out


@pytest.mark.parametrize("case", _single_rank_cases(12, 7), ids=lambda c: "-".join(map(str, c)))
def test_fuzz_single_rank(case, emu_lib, ref_lib):
    cn, level, passes, pct, corner, order, seed = case
    conn = util.CONNS[cn]()
    res = []
    for lib in (ref_lib, emu_lib):
        f = util.build_forest(lib, conn, level, passes, pct, corner, order, seed=seed)
        r = util.node_results(f)
        coarse = f.coarsen() if order == 2 else f.duplicate()
        if order == 2:
            coarse.balance(1)
        else:
            coarse.setMeshOrder(order - 1)
        res.append((r, f.createInterpolation(coarse).get()))
    util.assert_nodes_equal(res[0][0], res[1][0], "fuzz %s" % (case,))
    for a, b in zip(res[0][1][:3], res[1][1][:3]):
        assert np.array_equal(a, b)
    np.testing.assert_allclose(res[1][1][3], res[0][1][3], rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("case", _multi_rank_cases(5, 11), ids=lambda c: "-".join(map(str, c)))
def test_fuzz_multi_rank(case, emu_lib, ref_lib):
    cn, level, passes, pct, corner, order, ranks, repart, seed = case
    conn = util.CONNS[cn]()
    body = multirank.adapt_body(conn, level, passes, pct, corner, order, repart, seed=seed,
                                with_interp=True)
    a = multirank.run_thread_ranks(ref_lib, ranks, body, True)
    b = multirank.run_thread_ranks(emu_lib, ranks, body, False)
    multirank.compare_rank_results(a, b, "fuzz %s" % (case,))


def _apply_step(f, kind, pct, seed, corner):
    """One randomly chosen forest operation; returns the forest to go on with."""
    o = f.getOctants().as_array()
    if kind == "refine":
        f.refine(util.synth_flags(o, seed, pct))
        f.balance(corner)
    elif kind == "refine_neg":  # coarsen the flagged ones, refine a few others
        fl = util.synth_flags(o, seed, pct)
        fl = np.where(fl > 0, -1, np.where(util.synth_flags(o, seed + 1, 20) > 0, 1, 0))
        f.refine(fl.astype(np.int32))
        f.balance(corner)
    elif kind == "refine_clamp":  # level limits (reference :1798-1816)
        fl = np.where(util.synth_flags(o, seed, pct) > 0, 2, -2).astype(np.int32)
        f.refine(fl, 1, 4)
        f.balance(corner)
    elif kind == "coarsen":
        f = f.coarsen()
        f.balance(corner)
    elif kind == "dup":
        f = f.duplicate()
    elif kind == "repart":
        f.repartition()
    return f


def _sequence_cases(n, seed, multi):
    rng = random.Random(seed)
    kinds = ["refine", "refine", "refine_neg", "coarsen", "dup", "refine_clamp"]
    if multi:
        kinds += ["repart", "repart"]
    out = []
    for _ in range(n):
        cn = rng.choice(["box7", "connector15", "grid2", "butterfly2"] if multi else NAMES)
        steps = tuple((rng.choice(kinds), rng.choice([15, 35, 60]), rng.randrange(1, 10 ** 6))
                      for _ in range(rng.choice([2, 3, 4])))
        out.append((cn, rng.choice([1, 2]) if multi else rng.choice([0, 1, 2]),
                    rng.choice([2, 2, 3, 4]), rng.choice([0, 1]),
                    rng.choice([2, 3, 4, 6]) if multi else 1, steps))
    return out


def _case_id(c):
    return "-".join([c[0]] + [str(x) for x in c[1:5]] + [s[0] for s in c[5]])


@pytest.mark.parametrize("case", _sequence_cases(6, 101, False), ids=_case_id)
def test_fuzz_operation_sequences(case, emu_lib, ref_lib):
    """Random chains of refine (positive, negative, clamped) / coarsen /
    duplicate with a balance after each, every intermediate octant array and
    the final node arrays against the oracle."""
    cn, level, order, corner, _, steps = case
    conn = util.CONNS[cn]()
    res = []
    for lib in (ref_lib, emu_lib):
        f = util.build_forest(lib, conn, level, 0, 0, corner, order)
        stages = []
        for kind, pct, seed in steps:
            if f.getOctants().as_array().shape[0] > 40000 and kind.startswith("refine"):
                kind = "coarsen"
            f = _apply_step(f, kind, pct, seed, corner)
            stages.append(f.getOctants().as_array().copy())
        res.append((stages, util.node_results(f)))
    for k, (a, b) in enumerate(zip(res[0][0], res[1][0])):
        util.assert_octants_equal(a, b, "%s stage %d" % (case, k))
    util.assert_nodes_equal(res[0][1], res[1][1], "sequence %s" % (case,))


@pytest.mark.parametrize("case", _sequence_cases(4, 21, True), ids=_case_id)
def test_fuzz_operation_sequences_multi_rank(case, emu_lib, ref_lib):
    cn, level, order, corner, ranks, steps = case
    conn = util.CONNS[cn]()

    def body(lib, rank):
        from tmr_b200.forest import OctForest
        f = OctForest(order=order, lib=lib)
        f.setConnectivity(conn)
        f.createTrees(level)
        f.repartition()
        stages = []
        for kind, pct, seed in steps:
            f = _apply_step(f, kind, pct, seed, corner)
            stages.append(f.getOctants().as_array().copy())
        return stages, util.node_results(f)

    a = multirank.run_thread_ranks(ref_lib, ranks, body, True)
    b = multirank.run_thread_ranks(emu_lib, ranks, body, False)
    multirank.compare_rank_results(a, b, "sequence %s" % (case,))


def _interp_cases(n, seed):
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        fo = rng.choice([2, 3, 4, 5])
        co = min(fo, rng.choice([2, 3, 4]))
        it = rng.choice([1, 1, 2])
        if it == 2 and fo - co > 1:
            co = fo - 1
        out.append((rng.choice(NAMES), rng.choice([1, 2]), fo, co, it,
                    rng.randrange(1, 10 ** 6), rng.randrange(1, 10 ** 6),
                    rng.choice([20, 40, 60]), rng.choice([10, 30, 50]),
                    rng.choice(["unrelated", "coarser2", "same"])))
    return out


@pytest.mark.parametrize("case", _interp_cases(6, 11), ids=lambda c: "-".join(map(str, c)))
def test_fuzz_interpolation_pairs(case, emu_lib, ref_lib):
    """createInterpolation between forests that are NOT one coarsening apart:
    an independently refined 'coarse' forest, two coarsenings, or the same
    elements at a lower order (any order pair the reference accepts)."""
    cn, level, fo, co, it, sa, sb, pa, pb, mode = case
    conn = util.CONNS[cn]()
    res = []
    for lib in (ref_lib, emu_lib):
        f = util.build_forest(lib, conn, level, 2, pa, 1, fo, seed=sa)
        f.setMeshOrder(fo, it)
        if mode == "unrelated":
            c = util.build_forest(lib, conn, level, 1, pb, 1, co, seed=sb)
        elif mode == "coarser2":
            c = f.coarsen()
            c.balance(1)
            c = c.coarsen()
            c.balance(1)
        else:
            c = f.duplicate()
        c.setMeshOrder(co, it)
        res.append(f.createInterpolation(c).get())
    for a, b in zip(res[0][:3], res[1][:3]):
        assert np.array_equal(a, b)
    np.testing.assert_allclose(res[1][3], res[0][3], rtol=1e-12, atol=1e-300)


def _scrambled_cases(n, seed):
    rng = random.Random(seed)
    out = []
    for k in range(n):
        ranks = rng.choice([1, 1, 2, 3])
        out.append((rng.choice([(2, 2, 2), (3, 2, 1), (2, 2, 1), (3, 3, 1), (2, 1, 1)]),
                    rng.choice([1, 2]) if ranks > 1 else rng.choice([0, 1, 2]),
                    rng.choice([1, 2]), rng.choice([20, 40, 60]), rng.choice([0, 1]),
                    rng.choice([2, 2, 3, 4]), rng.randrange(1, 10 ** 6), ranks,
                    rng.randrange(1, 10 ** 6)))
    return out


@pytest.mark.parametrize("case", _scrambled_cases(8, 5), ids=lambda c: "-".join(map(str, c)))
def test_fuzz_scrambled_tree_orientations(case, emu_lib, ref_lib):
    """Boxes of trees with randomly permuted and flipped local axes (all 48
    cube symmetries, left-handed trees included): the inter-tree transforms of
    balance and of the node construction across every face / edge pairing."""
    dims, level, passes, pct, corner, order, seed, ranks, conn_seed = case
    conn = util.scrambled_conn(*dims, random.Random(conn_seed))
    if ranks == 1:
        res = []
        for lib in (ref_lib, emu_lib):
            f = util.build_forest(lib, conn, level, passes, pct, corner, order, seed=seed)
            r = util.node_results(f)
            coarse = f.coarsen() if order == 2 else f.duplicate()
            if order == 2:
                coarse.balance(1)
            else:
                coarse.setMeshOrder(order - 1)
            res.append((f.getOctants().as_array().copy(), r,
                        f.createInterpolation(coarse).get()))
        util.assert_octants_equal(res[0][0], res[1][0], "scrambled %s" % (case,))
        util.assert_nodes_equal(res[0][1], res[1][1], "scrambled %s" % (case,))
        for a, b in zip(res[0][2][:3], res[1][2][:3]):
            assert np.array_equal(a, b)
        np.testing.assert_allclose(res[1][2][3], res[0][2][3], rtol=1e-12, atol=1e-300)
    else:
        body = multirank.adapt_body(conn, level, passes, pct, corner, order, True, seed=seed,
                                    with_interp="repartitioned")
        a = multirank.run_thread_ranks(ref_lib, ranks, body, True)
        b = multirank.run_thread_ranks(emu_lib, ranks, body, False)
        multirank.compare_rank_results(a, b, "scrambled %s" % (case,))


def _query_cases(n, seed):
    rng = random.Random(seed)
    return [(rng.choice([(2, 2, 1), (2, 1, 1), (2, 2, 2), (3, 2, 1)]), rng.choice([2, 3, 4]),
             rng.choice([1, 2, 3, 4]), rng.choice([0, 1]), rng.randrange(1, 10 ** 6),
             rng.randrange(1, 10 ** 6), rng.choice([1, 2]), rng.choice([0, 1, 2]),
             rng.choice([1, 2, 3]), rng.choice([True, False]), rng.randrange(1, 10 ** 6))
            for _ in range(n)]


@pytest.mark.parametrize("case", _query_cases(5, 8), ids=lambda c: "-".join(map(str, c)))
def test_fuzz_find_enclosing(case, emu_lib, ref_lib):
    """findEnclosing asked about the nodes of an UNRELATED forest (other
    levels, other refinement): hits name the same local element, misses the
    same owner rank (reference :6228-6377)."""
    from tmr_b200.forest import OctForest
    dims, order, ranks, corner, sa, sb, la, lb, passes_b, lobatto, conn_seed = case
    conn = util.scrambled_conn(*dims, random.Random(conn_seed))
    kn = (-np.cos(np.pi * np.arange(order) / (order - 1)) if lobatto
          else np.linspace(-1, 1, order))
    kn[0], kn[-1] = -1.0, 1.0
    q = util.build_forest(ref_lib, conn, lb, passes_b, 40, corner,
                          seed=sb).getOctants().as_array().copy()
    q["info"] = np.random.default_rng(sb).integers(0, order ** 3, len(q))

    def body(lib, rank):
        f = OctForest(order=order, lib=lib)
        f.setConnectivity(conn)
        f.createTrees(la)
        f.repartition()
        for p in range(2):
            f.refine(util.synth_flags(f.getOctants().as_array(), sa + p, 35))
            f.balance(corner)
            f.repartition()
        idx, own = f.findEnclosing(order, kn, q)
        return idx.copy(), own.copy()

    a = multirank.run_thread_ranks(ref_lib, ranks, body, True)
    b = multirank.run_thread_ranks(emu_lib, ranks, body, False)
    for r in range(ranks):
        hit = a[r][0] >= 0
        assert np.array_equal(hit, b[r][0] >= 0), "rank %d: hit/miss" % r
        assert np.array_equal(a[r][0][hit], b[r][0][hit]), "rank %d: element" % r
        if ranks > 1:
            assert np.array_equal(a[r][1][~hit], b[r][1][~hit]), "rank %d: owner" % r


def _location_cases(n, seed):
    rng = random.Random(seed)
    return [(rng.choice([(2, 2, 1), (2, 1, 1), (2, 2, 2), (3, 2, 1)]), rng.choice([2, 3, 4, 5]),
             rng.choice([0, 1]), rng.choice([1, 1, 2, 3]), rng.choice([0, 1]),
             rng.randrange(1, 10 ** 6), rng.randrange(1, 10 ** 6)) for _ in range(n)]


@pytest.mark.parametrize("case", _location_cases(5, 9), ids=lambda c: "-".join(map(str, c)))
def test_fuzz_node_locations(case, emu_lib, ref_lib):
    """getPoints over randomly placed trilinear trees (re-oriented, some of
    them inverted): bit-equal locations, one and several ranks."""
    from tmr_b200.forest import OctForest
    dims, order, it, ranks, corner, seed, conn_seed = case
    conn = util.scrambled_conn(*dims, random.Random(conn_seed))
    xpts = np.random.default_rng(seed).normal(size=(int(conn.max()) + 1, 3)) * 3.0

    def body(lib, rank):
        f = OctForest(order=order, interp=it, lib=lib)
        f.setTrilinearTopology(conn, xpts)
        f.createTrees(1)
        if ranks > 1:
            f.repartition()
        for p in range(2):
            f.refine(util.synth_flags(f.getOctants().as_array(), seed + p, 30))
            f.balance(corner)
            if ranks > 1:
                f.repartition()
        f.createNodes()
        return f.getNodeNumbers().copy(), f.getPoints().copy()

    if ranks == 1:
        a, b = [body(ref_lib, 0)], [body(emu_lib, 0)]
    else:
        a = multirank.run_thread_ranks(ref_lib, ranks, body, True)
        b = multirank.run_thread_ranks(emu_lib, ranks, body, False)
    for r in range(ranks):
        assert np.array_equal(a[r][0], b[r][0])
        assert np.array_equal(a[r][1], b[r][1])


def _sparse_rank_cases(n, seed):
    rng = random.Random(seed)
    return [(rng.choice(["single", "rectangle", "single", "box7"]), rng.choice([0, 0, 1]),
             rng.choice([2, 2, 3, 4]), rng.choice([0, 1]), rng.choice([3, 4, 5, 8]),
             rng.choice([1, 2, 3]), rng.choice([20, 50, 80]), rng.randrange(1, 10 ** 6),
             rng.choice(["always", "never", "end", "first"])) for _ in range(n)]


@pytest.mark.parametrize("case", _sparse_rank_cases(6, 31), ids=lambda c: "-".join(map(str, c)))
def test_fuzz_more_ranks_than_elements(case, emu_lib, ref_lib):
    """Ranks that hold no element, own no node or receive nothing: forests of one
    or two trees at level 0/1 on up to 8 ranks, repartitioned always / never /
    once, through refine, balance, createNodes and a same-mesh prolongation."""
    from tmr_b200.forest import OctForest
    cn, level, order, corner, ranks, passes, pct, seed, repart = case
    conn = util.CONNS[cn]()

    def body(lib, rank):
        f = OctForest(order=order, lib=lib)
        f.setConnectivity(conn)
        f.createTrees(level)
        if repart in ("always", "first"):
            f.repartition()
        stages = []
        for p in range(passes):
            f.refine(util.synth_flags(f.getOctants().as_array(), seed + p, pct))
            f.balance(corner)
            if repart == "always":
                f.repartition()
            stages.append(f.getOctants().as_array().copy())
        if repart == "end":
            f.repartition()
        res = util.node_results(f)
        c = f.duplicate()
        if order > 2:
            c.setMeshOrder(order - 1)
        rows, d = util.interp_rows(f.createInterpolation(c))
        res["interp"] = {k: (cc.copy(), w.copy()) for k, (cc, w) in d.items()}
        return stages, res

    a = multirank.run_thread_ranks(ref_lib, ranks, body, True)
    b = multirank.run_thread_ranks(emu_lib, ranks, body, False)
    multirank.compare_rank_results(a, b, "sparse ranks %s" % (case,))


def _max_rank_cases(n, seed):
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        ranks = rng.choice([2, 3, 4, 6])
        out.append((rng.choice(["box7", "grid2", "connector15", "butterfly2"]), ranks,
                    rng.choice([2, 3]), rng.choice([0, 1]), rng.randrange(1, 10 ** 6),
                    tuple(rng.choice([0, 1, 2, ranks, ranks - 1, ranks + 3]) for _ in range(3))))
    return out


@pytest.mark.parametrize("case", _max_rank_cases(4, 12), ids=lambda c: "-".join(map(str, c)))
def test_fuzz_repartition_max_rank(case, emu_lib, ref_lib):
    """repartition(max_rank) (reference :1922-2088): the elements are dealt to
    the first max_rank ranks only; values below 1 and above the rank count
    included, the node construction afterwards on ranks left empty."""
    from tmr_b200.forest import OctForest
    cn, ranks, order, corner, seed, max_rank = case
    conn = util.CONNS[cn]()

    def body(lib, rank):
        f = OctForest(order=order, lib=lib)
        f.setConnectivity(conn)
        f.createTrees(1)
        f.repartition(max_rank[0])
        stages = []
        for p in range(2):
            f.refine(util.synth_flags(f.getOctants().as_array(), seed + p, 35))
            f.balance(corner)
            f.repartition(max_rank[1 + p])
            stages.append(f.getOctants().as_array().copy())
        return stages, util.node_results(f)

    a = multirank.run_thread_ranks(ref_lib, ranks, body, True)
    b = multirank.run_thread_ranks(emu_lib, ranks, body, False)
    multirank.compare_rank_results(a, b, "max_rank %s" % (case,))


def _name_cases(n, seed):
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        ranks = rng.choice([1, 1, 2, 3])
        out.append((rng.choice([(2, 2, 1), (2, 1, 1), (2, 2, 2), (1, 1, 1)]), rng.randrange(1, 10 ** 6),
                    rng.choice([2, 3, 4]), ranks, rng.choice([1, 2]) if ranks > 1 else rng.choice([0, 1, 2]),
                    rng.choice([0, 1, 2]), rng.randrange(1, 10 ** 6), rng.randrange(1, 10 ** 6),
                    rng.choice([(None, "a", "b"), (None, None, None, "a"), ("a", "b"), ("a",)])))
    return out


@pytest.mark.parametrize("case", _name_cases(5, 141), ids=lambda c: "-".join(map(str, c[:7])))
def test_fuzz_name_queries(case, emu_lib, ref_lib):
    """getOctsWithName / getNodesWithName with random names on every vertex,
    edge, face and volume of re-oriented tree boxes: root octants (level 0, no
    refinement), sparse and dense naming, a name nothing carries; 1-3 ranks."""
    from tmr_b200.forest import OctForest
    dims, conn_seed, order, ranks, level, passes, seed, name_seed, names = case
    conn = util.scrambled_conn(*dims, random.Random(conn_seed))
    xpts = np.random.default_rng(seed).normal(size=(int(conn.max()) + 1, 3))

    def body(lib, rank):
        f = OctForest(order=order, lib=lib)
        f.setTrilinearTopology(conn, xpts)
        c = f.getConnectivity()
        r = random.Random(name_seed)
        for kind, count in ((0, c["nnodes"]), (1, c["nedges"]), (2, c["nfaces"]), (3, c["nblocks"])):
            for i in range(count):
                name = r.choice(names)
                if name:
                    f.setEntityName(kind, i, name)
        f.createTrees(level)
        if ranks > 1:
            f.repartition()
        for p in range(passes):
            f.refine(util.synth_flags(f.getOctants().as_array(), seed + p, 30))
            f.balance(1)
            if ranks > 1:
                f.repartition()
        f.createNodes()
        return ([f.getOctsWithName(x) for x in ("a", "b", "c")] +
                [f.getNodesWithName(x) for x in ("a", "b", "c")])

    if ranks == 1:
        a, b = [body(ref_lib, 0)], [body(emu_lib, 0)]
    else:
        a = multirank.run_thread_ranks(ref_lib, ranks, body, True)
        b = multirank.run_thread_ranks(emu_lib, ranks, body, False)
    for r in range(ranks):
        for x, y in zip(a[r], b[r]):
            assert np.array_equal(x, y), r
