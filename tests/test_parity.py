"""Parity of the kernel bodies against the oracle (the unmodified reference).

Every test runs twice through the `impl` fixture:
  * "emu": test-only serial emulation of the kernel bodies (CPU, checks logic)
  * "gpu": the product library on a B200 (marker: gpu) -- the parity tests
           proper, through the reference-facing class API / C-ABI.
Bar: bit-exact octants, info, conn, node numbers, dependent CSR pattern;
weights within 1e-12 relative (north_star).
"""
import numpy as np
import pytest

import util
from tmr_b200 import _capi
from tmr_b200.forest import OctForest, array_contains, array_sort


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def impl(request):
    return request.getfixturevalue(request.param + "_lib")


CASES = [
    # name, conn, level, passes, pct, corner, order
    ("single", "single", 2, 3, 30, 0, 2),
    ("single_corner", "single", 1, 4, 40, 1, 2),
    ("rectangle", "rectangle", 1, 3, 30, 0, 2),
    ("box7", "box7", 1, 3, 30, 0, 2),
    ("box7_corner", "box7", 2, 2, 30, 1, 2),
    ("connector15", "connector15", 1, 3, 30, 0, 2),
    ("grid2", "grid2", 1, 3, 35, 0, 2),
    ("butterfly2", "butterfly2", 1, 2, 30, 1, 2),
    ("box7_order3", "box7", 1, 2, 30, 1, 3),
    ("single_order3", "single", 2, 2, 30, 0, 3),
    ("connector15_order3", "connector15", 0, 3, 40, 0, 3),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_refine_balance_nodes(case, impl, ref_lib):
    _, conn_name, level, passes, pct, corner, order = case
    conn = util.CONNS[conn_name]()
    rec_ref, rec_new = [], []
    f_ref = util.build_forest(ref_lib, conn, level, passes, pct, corner, order,
                              record=rec_ref)
    f_new = util.build_forest(impl, conn, level, passes, pct, corner, order,
                              record=rec_new)
    for (name, a), (_, b) in zip(rec_ref, rec_new):
        util.assert_octants_equal(a, b, name)
    util.assert_nodes_equal(util.node_results(f_ref), util.node_results(f_new),
                            case[0])


@pytest.mark.parametrize("conn_name", ["box7", "connector15"])
def test_bernstein_order2(conn_name, impl, ref_lib):
    """TMR_BERNSTEIN_POINTS at order 2 (reference eval_bernstein_weights /
    bernstein_shape_functions, src/TMRInterpolation.h:164-183,309-322):
    nodes, dependent weights and the prolongation against the oracle."""
    conn = util.CONNS[conn_name]()
    res = []
    for lib in (ref_lib, impl):
        f = util.build_forest(lib, conn, 1, 3, 30, 1, 2, interp=2)
        assert f.getInterpType() == 2
        nodes = util.node_results(f)
        coarse = f.coarsen()
        coarse.balance(1)
        res.append((nodes, f.createInterpolation(coarse)))
    util.assert_nodes_equal(res[0][0], res[1][0], "bernstein order 2")
    util.assert_interp_equal(res[0][1], res[1][1], "bernstein order 2 interp")


@pytest.mark.parametrize("conn_name", ["single", "box7", "connector15"])
def test_bernstein_order3(conn_name, impl, ref_lib):
    """Order-3 Bernstein points: corner / edge / face / block node labels
    (reference initLabel src/TMROctForest.cpp:6798-6811), untrimmed dependent
    ranges (:3755-3761), subdivision weights (:5364-5374, :5453-5473) and the
    order 3 -> 2 and 2 -> 2 prolongations (:6434-6500)."""
    conn = util.CONNS[conn_name]()
    res = []
    for lib in (ref_lib, impl):
        f = util.build_forest(lib, conn, 1, 2, 30, 1, 3, interp=2)
        nodes = util.node_results(f)
        low = f.duplicate()
        low.setMeshOrder(2, 2)
        v32 = f.createInterpolation(low)
        coarse = low.coarsen()
        coarse.balance(1)
        v22 = low.createInterpolation(coarse)
        res.append((nodes, v32, v22))
    util.assert_nodes_equal(res[0][0], res[1][0], "bernstein order 3")
    util.assert_interp_equal(res[0][1], res[1][1], "bernstein 3 -> 2")
    util.assert_interp_equal(res[0][2], res[1][2], "bernstein 2 -> 2")


@pytest.mark.parametrize("interp", [1, 2], ids=["gauss_lobatto", "bernstein"])
@pytest.mark.parametrize("conn_name", ["single", "box7", "connector15"])
def test_order4(conn_name, interp, impl, ref_lib):
    """Order 4: (order-2)^dim nodes per edge / face / block entity, ordered
    through edge reversal and face orientation (reference createLocalConn
    src/TMROctForest.cpp:4660-4867, getEdgeNodes/getFaceNodes :4886-5146), and
    the 4 -> 3 -> 2 prolongations (examples/interp/interp.py uses order 4)."""
    conn = util.CONNS[conn_name]()
    res = []
    for lib in (ref_lib, impl):
        f = util.build_forest(lib, conn, 1, 2, 30, 1, 4, interp=interp)
        nodes = util.node_results(f)
        o3 = f.duplicate()
        o3.setMeshOrder(3, interp)
        v43 = f.createInterpolation(o3)
        o2 = o3.duplicate()
        o2.setMeshOrder(2, interp)
        v32 = o3.createInterpolation(o2)
        res.append((nodes, v43, v32))
    util.assert_nodes_equal(res[0][0], res[1][0], "order 4")
    util.assert_interp_equal(res[0][1], res[1][1], "order 4 -> 3")
    util.assert_interp_equal(res[0][2], res[1][2], "order 3 -> 2")


@pytest.mark.parametrize("interp", [0, 1, 2], ids=["uniform", "gauss_lobatto", "bernstein"])
@pytest.mark.parametrize("order", [2, 3, 4])
def test_eval_interp(order, interp, impl, ref_lib):
    """evalInterp: shape functions and their first / second derivatives
    (reference src/TMROctForest.cpp:1508-1620)."""
    pts = [(0.0, 0.0, 0.0), (-1.0, 1.0, 0.25), (0.3, -0.7, 0.9)]
    a = OctForest(order=order, interp=interp, lib=ref_lib)
    b = OctForest(order=order, interp=interp, lib=impl)
    for pt in pts:
        for der in (0, 1, 2):
            ra, rb = a.evalInterp(pt, der), b.evalInterp(pt, der)
            ra, rb = (ra, rb) if der else ([ra], [rb])
            for x, y in zip(ra, rb):
                np.testing.assert_allclose(y, x, rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("order,interp,conn_name", [(5, 1, "box7"), (5, 2, "connector15"),
                                                    (6, 1, "connector15"),
                                                    (8, 1, "single"), (7, 0, "box7"),
                                                    (10, 1, "single")])
def test_high_order(order, interp, conn_name, impl, ref_lib):
    """Orders 5..16 (kMaxOrder = the reference's MAX_ORDER): same entity machinery as order 4, prolongation
    to the next lower order.  Bernstein points stop at order 5: the reference's
    eval_bernstein_weights has no table beyond it (src/TMRInterpolation.h:309-455)
    and leaves the weights unset."""
    conn = util.CONNS[conn_name]()
    res = []
    for lib in (ref_lib, impl):
        level = 1 if (conn_name == "single" and order < 9) else 0
        f = util.build_forest(lib, conn, level, 2, 40 if order >= 9 else 30, 1,
                              order, interp=interp)
        nodes = util.node_results(f)
        low = f.duplicate()
        low.setMeshOrder(order - 1, interp)
        res.append((nodes, f.createInterpolation(low)))
    util.assert_nodes_equal(res[0][0], res[1][0], "order %d" % order)
    util.assert_interp_equal(res[0][1], res[1][1], "order %d -> %d" % (order, order - 1))


def test_order16_nodes(impl, ref_lib):
    """The reference's MAX_ORDER: 14 nodes per edge, 196 per face, 2744 per
    block entity, on a tree with hanging faces and edges (nodes and dependent
    CSR; the prolongation at this order takes the reference itself minutes)."""
    res = []
    for lib in (ref_lib, impl):
        f = util.build_forest(lib, util.single_conn(), 1, 1, 30, 1, 16, interp=1)
        res.append(util.node_results(f))
    assert len(res[0]["dep"][0]) > 1
    util.assert_nodes_equal(res[0], res[1], "order 16")


def test_deep_corner_refinement(impl, ref_lib):
    """A tree refined 13 levels deep into one corner: the leaf bitmap of the
    hanging-node kernel covers levels 0..9 within its budget, so the probes at
    parent levels 10-12 take the key-search path; the node keys use 14+1 bits
    per axis."""
    res = []
    for lib in (ref_lib, impl):
        f = OctForest(order=2, lib=lib)
        f.setConnectivity(util.single_conn())
        f.createTrees(1)
        for _ in range(12):
            flags = np.zeros(f.getNumOctants(), dtype=np.int32)
            flags[0] = 1
            f.refine(flags)
            f.balance(1)
        res.append(util.node_results(f))
    assert int(res[0]["octants"]["level"].max()) == 13
    util.assert_octants_equal(res[0]["octants"], res[1]["octants"], "deep corner")
    util.assert_nodes_equal(res[0], res[1], "deep corner")


def test_connectivity_tables(impl, ref_lib):
    """setConnectivity derives identical edge/face numbering, inverse maps,
    orientation ids (reference src/TMROctForest.cpp:558-1143)."""
    for name, make in util.CONNS.items():
        conn = make()
        a = OctForest(lib=ref_lib)
        b = OctForest(lib=impl)
        a.setConnectivity(conn)
        b.setConnectivity(conn)
        ca, cb = a.getConnectivity(), b.getConnectivity()
        for k in ca:
            assert np.array_equal(ca[k], cb[k]), (name, k)
        ia, ib = a.getInverseConnectivity(), b.getInverseConnectivity()
        for k in ia:
            assert np.array_equal(ia[k], ib[k]), (name, k)


def test_refine_negative_and_multilevel(impl, ref_lib):
    """coarsening flags, multi-level refinement, min/max clamps
    (reference src/TMROctForest.cpp:2197-2296)."""
    conn = util.box_conn()
    rng = np.random.default_rng(7)
    fa = util.build_forest(ref_lib, conn, 2, 1, 30, 0)
    fb = util.build_forest(impl, conn, 2, 1, 30, 0)
    for rnd in range(3):
        n = fa.getNumOctants()
        assert n == fb.getNumOctants()
        flags = rng.integers(-2, 3, n).astype(np.int32)
        fa.refine(flags, 1, 5)
        fb.refine(flags, 1, 5)
        util.assert_octants_equal(fa.getOctants().as_array(),
                                  fb.getOctants().as_array(), "refine round %d" % rnd)
        fa.balance(rnd % 2)
        fb.balance(rnd % 2)
        util.assert_octants_equal(fa.getOctants().as_array(),
                                  fb.getOctants().as_array(), "balance round %d" % rnd)
    util.assert_nodes_equal(util.node_results(fa), util.node_results(fb), "neg")


def test_refine_null_everywhere(impl, ref_lib):
    conn = util.rectangle_conn()
    fa = util.build_forest(ref_lib, conn, 1, 1, 50, 0)
    fb = util.build_forest(impl, conn, 1, 1, 50, 0)
    fa.refine(None, 0, 4)
    fb.refine(None, 0, 4)
    util.assert_octants_equal(fa.getOctants().as_array(), fb.getOctants().as_array())
    fa.balance(1)
    fb.balance(1)
    util.assert_octants_equal(fa.getOctants().as_array(), fb.getOctants().as_array())


def test_info_survives_refine(impl, ref_lib):
    """octants kept verbatim by refine() keep the hanging `info` createNodes()
    wrote (hash stores the record as is, reference :2199,2227)."""
    conn = util.box_conn()
    fa = util.build_forest(ref_lib, conn, 1, 2, 30, 0)
    fb = util.build_forest(impl, conn, 1, 2, 30, 0)
    fa.createNodes()
    fb.createNodes()
    octs = fa.getOctants().as_array()
    assert octs["info"].any()
    flags = util.synth_flags(octs, 99, 20)
    fa.refine(flags)
    fb.refine(flags)
    util.assert_octants_equal(fa.getOctants().as_array(), fb.getOctants().as_array())


def test_stale_nodes_after_balance(impl, ref_lib):
    """balance() does not invalidate node data; createNodes() is then a no-op
    (reference :4071-4075) -- replicated, not fixed."""
    conn = util.single_conn()
    fa = util.build_forest(ref_lib, conn, 2, 1, 30, 0)
    fb = util.build_forest(impl, conn, 2, 1, 30, 0)
    ra, rb = util.node_results(fa), util.node_results(fb)
    for f in (fa, fb):
        f.balance(1)
        f.createNodes()
    assert np.array_equal(fa.getMeshConn(), ra["conn"])
    assert np.array_equal(fb.getMeshConn(), rb["conn"])


def test_coarsen_duplicate(impl, ref_lib):
    conn = util.box_conn()
    fa = util.build_forest(ref_lib, conn, 2, 2, 30, 1)
    fb = util.build_forest(impl, conn, 2, 2, 30, 1)
    da, db = fa.duplicate(), fb.duplicate()
    util.assert_octants_equal(da.getOctants().as_array(), db.getOctants().as_array())
    ca, cb = fa.coarsen(), fb.coarsen()
    util.assert_octants_equal(ca.getOctants().as_array(), cb.getOctants().as_array(),
                              "coarsen")
    ca.balance(1)
    cb.balance(1)
    util.assert_octants_equal(ca.getOctants().as_array(), cb.getOctants().as_array(),
                              "coarsen+balance")


def _hierarchy(lib, conn, order):
    """TopOptUtils-style multigrid hierarchy (reference tmr/TopOptUtils.py:79-99)."""
    f0 = util.build_forest(lib, conn, 1, 2, 30, 1, order=order)
    forests = [f0]
    o = order
    for _ in range(3):
        prev = forests[-1]
        if o > 2:
            nxt = prev.duplicate()
            o -= 1
            nxt.setMeshOrder(o)
        else:
            nxt = prev.coarsen()
            nxt.balance(1)
        forests.append(nxt)
    interps = []
    for k in range(len(forests) - 1):
        interps.append(forests[k].createInterpolation(forests[k + 1]))
    return forests, interps


@pytest.mark.parametrize("order", [2, 3])
def test_create_interpolation(order, impl, ref_lib):
    conn = util.box_conn()
    fa, ia = _hierarchy(ref_lib, conn, order)
    fb, ib = _hierarchy(impl, conn, order)
    for k, (va, vb) in enumerate(zip(ia, ib)):
        util.assert_interp_equal(va, vb, "level %d" % k)
        rows, d = util.interp_rows(vb)
        assert len(rows) == fb[k].getNumOwnedNodes()
        for r, (c, w) in d.items():
            assert abs(w.sum() - 1.0) < 1e-13


def test_find_enclosing_and_transform(impl, ref_lib):
    conn = util.box_conn()
    fa = util.build_forest(ref_lib, conn, 1, 2, 30, 0)
    fb = util.build_forest(impl, conn, 1, 2, 30, 0)
    fine_a = util.build_forest(ref_lib, conn, 1, 3, 30, 0)
    octs = fine_a.getOctants().as_array().copy()
    rng = np.random.default_rng(3)
    for order in (2, 3):
        octs["info"] = rng.integers(0, order ** 3, len(octs))
        knots = np.array([-1.0, 1.0]) if order == 2 else np.array([-1.0, 0.0, 1.0])
        ia, _ = fa.findEnclosing(order, knots, octs)
        ib, _ = fb.findEnclosing(order, knots, octs)
        assert np.array_equal(ia, ib)
        assert (ia >= 0).all()
    # transformNode on every element corner
    nodes = np.repeat(fine_a.getOctants().as_array(), 8)
    h = (1 << (30 - nodes["level"].astype(np.int64)))
    c = np.tile(np.arange(8), len(nodes) // 8)
    nodes["x"] = nodes["x"] + h * (c & 1)
    nodes["y"] = nodes["y"] + h * ((c >> 1) & 1)
    nodes["z"] = nodes["z"] + h * (c >> 2)
    for edge_dir in (-1, 0, 1, 2):
        ra, reva, fida = fa.transformNodes(nodes, edge_dir)
        rb, revb, fidb = fb.transformNodes(nodes, edge_dir)
        for fld in ("block", "x", "y", "z"):
            assert np.array_equal(ra[fld], rb[fld])
        assert np.array_equal(reva, revb)
        assert np.array_equal(fida, fidb)


@pytest.mark.parametrize("node_mode", [0, 1])
def test_octant_array_sort_contains(node_mode, impl, ref_lib):
    """TMROctantArray::sort / contains on arbitrary arrays incl. duplicates,
    nodes at 2^30-1 and negative coordinates (reference src/TMROctant.cpp:357-424)."""
    rng = np.random.default_rng(11 + node_mode)
    for n in (0, 1, 2, 33, 1000, 5000):
        rec = util.random_octants(rng, n, 5, 6)
        if n >= 33:
            rec[: n // 4] = rec[n // 4: 2 * (n // 4)]  # duplicates
            rec["level"][: n // 8] = 6
            rec["x"][5] = (1 << 30) - 1
            rec["y"][6] = -(1 << 24)
        if node_mode:
            rec["info"] = rng.integers(0, 3, n)
        a = array_sort(ref_lib, rec, node_mode)
        b = array_sort(impl, rec, node_mode)
        assert len(a) == len(b)
        key = ("block", "x", "y", "z", "info") if node_mode else ("block", "x", "y", "z", "level")
        for fld in key:
            assert np.array_equal(a[fld], b[fld]), (n, fld)
        if n:
            q = np.concatenate([a[:: max(1, len(a) // 50)], util.random_octants(rng, 20, 5, 6)])
            for use_pos in (0, 1):
                ia = array_contains(ref_lib, a, q, node_mode, use_pos)
                ib = array_contains(impl, b, q, node_mode, use_pos)
                assert np.array_equal(ia >= 0, ib >= 0)
                assert np.array_equal(ia, ib)


def test_empty_and_level0(impl, ref_lib):
    """level-0 forests, balance/createNodes on trivial inputs."""
    for conn in (util.single_conn(), util.box_conn()):
        fa = OctForest(lib=ref_lib)
        fb = OctForest(lib=impl)
        for f in (fa, fb):
            f.setConnectivity(conn)
            f.createTrees(0)
            f.balance(1)
        util.assert_octants_equal(fa.getOctants().as_array(), fb.getOctants().as_array())
        util.assert_nodes_equal(util.node_results(fa), util.node_results(fb), "level0")
        n = fa.getNumOctants()
        flags = np.zeros(n, dtype=np.int32)
        flags[0] = 3
        for f in (fa, fb):
            f.refine(flags)
            f.balance(0)
        util.assert_nodes_equal(util.node_results(fa), util.node_results(fb), "one deep")


def test_create_random_trees(impl, ref_lib):
    """createRandomTrees draws from libc rand() in the reference's order
    (reference src/TMROctForest.cpp:1863-1881): the same seed gives the same
    forest, and the usual pipeline on top of it matches."""
    import ctypes

    libc = ctypes.CDLL("libc.so.6")
    conn = util.connector_conn()
    results = []
    for lib in (ref_lib, impl):
        libc.srand(12345)
        f = OctForest(lib=lib)
        f.setConnectivity(conn)
        f.createRandomTrees(15, 0, 6)
        first = f.getOctants().as_array().copy()
        f.balance(1)
        results.append((first, util.node_results(f)))
    util.assert_octants_equal(results[0][0], results[1][0], "random trees")
    util.assert_nodes_equal(results[0][1], results[1][1], "random trees nodes")


def test_unbalanced_input_to_create_nodes(impl, ref_lib):
    """createNodes on a forest that was refined but not balanced: the reference
    produces *something* deterministic from its probes; so must we."""
    conn = util.single_conn()
    fa = OctForest(lib=ref_lib)
    fb = OctForest(lib=impl)
    for f in (fa, fb):
        f.setConnectivity(conn)
        f.createTrees(2)
        f.balance(0)
    util.assert_nodes_equal(util.node_results(fa), util.node_results(fb), "uniform")


def test_write_through_octant_array(impl, ref_lib):
    """Python's OctantArray.__setitem__ writes into the forest's array
    (reference tmr/TMR.pyx:3303-3317): tags and info written by the caller are
    what the next getOctants() shows, and octants written by the caller are what
    the next operation works on."""
    conn = util.box_conn()
    outs = []
    for lib in (ref_lib, impl):
        f = util.build_forest(lib, conn, 1, 1, 30, 0)
        rec = f.getOctants().as_array().copy()
        rec["tag"] = 1000 + np.arange(len(rec))
        f.writeOctants(rec)
        again = f.getOctants().as_array().copy()
        f.createNodes()  # only rewrites info
        after = f.getOctants().as_array().copy()
        outs.append((again, after))
    util.assert_octants_equal(outs[0][0], outs[1][0], "write-through")
    util.assert_octants_equal(outs[0][1], outs[1][1], "write-through + createNodes")
    assert (outs[1][1]["tag"] >= 1000).all()


def _oct_ptr(a):
    return a.ctypes.data if len(a) else None


@pytest.mark.parametrize("node_mode", [0])
def test_octant_array_merge(node_mode, impl, ref_lib):
    """TMROctantArray::merge (reference src/TMROctant.cpp:429-509): set union,
    `this` keeps its entry on ties, unsorted inputs are sorted first.  Element
    arrays only: the reference sorts node arrays with compareNode but merges
    with compare(), which overruns its own buffer on node arrays (and nothing
    in the reference calls merge at all)."""
    rng = np.random.default_rng(5 + node_mode)
    for na, nb in ((0, 0), (0, 7), (9, 0), (40, 40), (500, 333)):
        a = util.random_octants(rng, na, 3, 4)
        b = util.random_octants(rng, nb, 3, 4)
        if na >= 40 and nb >= 40:
            b[: nb // 3] = a[: nb // 3]  # common entries
            b["tag"][: nb // 3] = 77     # ... told apart by the tag: a's survive
        if node_mode:
            a["info"] = rng.integers(0, 3, na)
            b["info"] = rng.integers(0, 3, nb)
        res = []
        for lib in (ref_lib, impl):
            out = np.zeros(na + nb + 1, dtype=_capi.OCT_DTYPE)
            n = lib.tmrc_array_merge(_oct_ptr(a), na, _oct_ptr(b), nb, node_mode,
                                     out.ctypes.data, len(out))
            res.append(out[:n].copy())
        assert len(res[0]) == len(res[1]), (na, nb)
        for fld in ("block", "x", "y", "z", "level", "tag"):
            assert np.array_equal(res[0][fld], res[1][fld]), (na, nb, fld)


def test_octant_queue(impl, ref_lib):
    """TMROctantQueue push / pop / length / toArray: FIFO (reference
    src/TMROctant.cpp:514-590)."""
    rng = np.random.default_rng(3)
    for n, npop in ((0, 0), (1, 1), (50, 17), (50, 50), (20, 30)):
        rec = util.random_octants(rng, n, 2, 5)
        res = []
        for lib in (ref_lib, impl):
            popped = np.zeros(max(npop, 1), dtype=_capi.OCT_DTYPE)
            rest = np.zeros(max(n, 1), dtype=_capi.OCT_DTYPE)
            left = lib.tmrc_queue_exercise(_oct_ptr(rec), n, npop, popped.ctypes.data,
                                           rest.ctypes.data)
            res.append((left, popped[: min(n, npop)].copy(), rest[:left].copy()))
        assert res[0][0] == res[1][0] == max(0, n - npop)
        assert res[0][1].tobytes() == res[1][1].tobytes()
        assert res[0][2].tobytes() == res[1][2].tobytes()
        assert res[1][1].tobytes() == rec[: min(n, npop)].tobytes()


@pytest.mark.parametrize("node_mode", [0, 1])
def test_octant_hash(node_mode, impl, ref_lib):
    """TMROctantHash::addOctant / toArray (reference src/TMROctant.cpp:599-798):
    which insertions are new, and the unique set (toArray order is a hash detail
    every caller sorts away, so sets are compared).  50 000 entries cross the
    reference's rehash threshold (10 x 4095)."""
    rng = np.random.default_rng(17 + node_mode)
    for n in (0, 1, 300, 50000):
        rec = util.random_octants(rng, n, 3, 5 if n < 1000 else 8)
        if n >= 300:
            rec[n // 2:] = rec[: n - n // 2]  # every entry twice
            if node_mode:
                rec["info"] = rng.integers(0, 2, n)
            else:
                rec["level"][: n // 10] += 1  # same anchor, other level: distinct
        res = []
        for lib in (ref_lib, impl):
            added = np.zeros(max(n, 1), dtype=np.int32)
            out = np.zeros(max(n, 1), dtype=_capi.OCT_DTYPE)
            m = lib.tmrc_hash_exercise(_oct_ptr(rec), n, node_mode, added.ctypes.data,
                                       out.ctypes.data, len(out))
            o = out[:m]
            key = np.lexsort((o["info"] if node_mode else o["level"], o["z"], o["y"],
                              o["x"], o["block"]))
            res.append((added[:n].copy(), o[key].copy()))
        assert np.array_equal(res[0][0], res[1][0])
        assert len(res[0][1]) == len(res[1][1])
        for fld in ("block", "x", "y", "z", "level", "info", "tag"):
            assert np.array_equal(res[0][1][fld], res[1][1][fld]), fld


@pytest.mark.parametrize("order", [2, 3])
def test_device_views_hold_what_the_getters_return(order, impl, ref_lib):
    """the device-resident arrays a GPU assembler would read
    (tmrgpu_assembler_views: elem_ptr/conn/dependent CSR in the shape createTACS
    hands to TACSAssembler, reference src/TMR_TACSCreator.cpp:332-461) equal the
    arrays the host getters return -- which are compared with the oracle -- and
    the bulk prolongation hand-off (createInterpolationCSR) equals the per-row
    addInterp stream of the reference."""
    conn = util.box_conn()
    fa = util.build_forest(ref_lib, conn, 1, 2, 30, 1, order)
    fb = util.build_forest(impl, conn, 1, 2, 30, 1, order)
    ra, rb = util.node_results(fa), util.node_results(fb)
    util.assert_nodes_equal(ra, rb, "device views")
    v = fb.assemblerViews()
    npe = order ** 3
    assert v["num_elements"] == len(rb["octants"]) and v["order"] == order
    assert v["num_owned_nodes"] == fa.getNumOwnedNodes()
    assert np.array_equal(v["elem_ptr"], npe * np.arange(len(rb["octants"]) + 1))
    assert np.array_equal(v["conn"].reshape(-1, npe), ra["conn"].reshape(-1, npe))
    assert np.array_equal(v["dep_ptr"], ra["dep"][0])
    assert np.array_equal(v["dep_conn"], ra["dep"][1])
    np.testing.assert_allclose(v["dep_weights"], ra["dep"][2], rtol=1e-12, atol=0)
    assert np.array_equal(np.sort(v["node_numbers"]), ra["node_numbers"])
    # bulk prolongation vs the reference's addInterp stream
    ca = fa.coarsen() if order == 2 else fa.duplicate()
    cb = fb.coarsen() if order == 2 else fb.duplicate()
    if order == 2:
        ca.balance(1)
        cb.balance(1)
    else:
        ca.setMeshOrder(2)
        cb.setMeshOrder(2)
    rows_a, rowp_a, cols_a, vals_a = fa.createInterpolation(ca).get()
    rows_b, rowp_b, cols_b, vals_b = fb.createInterpolationCSR(cb)
    assert np.array_equal(rows_a, rows_b) and np.array_equal(rowp_a, rowp_b)
    assert np.array_equal(cols_a, cols_b)
    np.testing.assert_allclose(vals_b, vals_a, rtol=1e-12, atol=1e-300)


def _warped_points(conn, seed=11):
    """physical locations for the super-mesh nodes: the 7-tree box layout of
    reference examples/parallel/octant_test.cpp:47-50 is only topological here;
    any injective placement will do for a trilinear geometry"""
    n = int(np.max(conn)) + 1
    rng = np.random.default_rng(seed)
    return rng.uniform(-1.0, 1.0, (n, 3)) + 3.0 * np.arange(n)[:, None] * np.array([1.0, 0.3, -0.2])


@pytest.mark.parametrize("order,conn_name", [(2, "box7"), (3, "box7"), (4, "connector15"),
                                             (2, "butterfly2")])
def test_node_locations(order, conn_name, impl, ref_lib):
    """getPoints with a topology = evaluateNodeLocations
    (reference src/TMROctForest.cpp:5524-5675): every local node evaluated
    through the FIRST element and slot that reference it.  The oracle runs the
    unmodified reference loop over a stand-in topology of trilinear volumes
    (oracle/shim/stubs.cpp); the drop-in evaluates the same volumes on the
    device.  Bit-exact: both sides evaluate the identical trilinear expression
    and the library is built without fused multiply-adds."""
    conn = util.CONNS[conn_name]()
    xpts = _warped_points(conn)
    res = []
    for lib in (ref_lib, impl):
        f = OctForest(order=order, lib=lib)
        f.setTrilinearTopology(conn, xpts)
        f.createTrees(1)
        for p in range(2):
            f.refine(util.synth_flags(f.getOctants().as_array(), 2024 + p, 30))
            f.balance(1)
        f.createNodes()
        res.append((f.getNodeNumbers().copy(), f.getPoints(), f.getMeshConn().copy()))
    (na, xa, ca), (nb, xb, cb) = res
    assert np.array_equal(na, nb) and np.array_equal(ca, cb)
    assert xa.shape == xb.shape and len(xa) == len(na) and len(xa) > 100
    assert np.abs(xa).max() > 1.0
    assert np.array_equal(xa, xb)


def test_node_locations_without_topology_are_zero(impl, ref_lib):
    """no topology: X is all zeros in the reference (:5526-5527) and here"""
    conn = util.box_conn()
    for lib in (ref_lib, impl):
        f = util.build_forest(lib, conn, 1, 1, 30, 0)
        x = f.getPoints()
        assert len(x) == len(f.getNodeNumbers()) and not x.any()


@pytest.mark.parametrize("order,conn_name", [(2, "box7"), (3, "box7"), (2, "connector15")])
def test_create_nodes_after_bare_refine(order, conn_name, impl, ref_lib):
    """createNodes on what refine() leaves behind -- not a complete tree: a
    refined octant is replaced by ONE representative child with the same anchor
    (reference :2169-2329) -- without a balance in between.  Eight consecutive
    keys whose first and last are siblings 0 and 7 are then NOT necessarily a
    family; the node construction must fall back from the (leaf, slot) scheme
    and check every member."""
    conn = util.CONNS[conn_name]()
    res = []
    for lib in (ref_lib, impl):
        f = OctForest(order=order, lib=lib)
        f.setConnectivity(conn)
        f.createTrees(1)
        f.refine(util.synth_flags(f.getOctants().as_array(), 2024, 30))
        res.append(util.node_results(f))
    util.assert_nodes_equal(res[0], res[1], "bare refine")


def _name_entities(f, seed, choices=(None, "clamped", "load")):
    """Give every vertex / edge / face / volume of the stand-in topology a
    random name (or none), identically for every backend."""
    import random
    c = f.getConnectivity()
    r = random.Random(seed)
    for kind, count in ((OctForest.VERTEX, c["nnodes"]), (OctForest.EDGE, c["nedges"]),
                        (OctForest.FACE, c["nfaces"]), (OctForest.VOLUME, c["nblocks"])):
        for i in range(count):
            name = r.choice(choices)
            if name:
                f.setEntityName(kind, i, name)


@pytest.mark.parametrize("order,conn_name,level", [(2, "box7", 1), (3, "box7", 1),
                                                   (4, "connector15", 0), (2, "butterfly2", 1),
                                                   (2, "single", 0)])
def test_name_queries(order, conn_name, level, impl, ref_lib):
    """getOctsWithName / getNodesWithName (reference src/TMROctForest.cpp:
    5747-5862, 5882-6203), the calls boundary conditions are applied through:
    the unmodified reference over a stand-in topology with nameable entities
    against the drop-in, including a level-0 forest (root octants touch all six
    faces) and a name nothing carries."""
    conn = util.CONNS[conn_name]()
    xpts = _warped_points(conn)
    res = []
    for lib in (ref_lib, impl):
        f = OctForest(order=order, lib=lib)
        f.setTrilinearTopology(conn, xpts)
        _name_entities(f, 7)
        f.createTrees(level)
        octs0 = [f.getOctsWithName(n) for n in ("clamped", "load")]
        for p in range(2 if level else 1):
            f.refine(util.synth_flags(f.getOctants().as_array(), 2024 + p, 30))
            f.balance(1)
        f.createNodes()
        res.append(octs0 + [f.getOctsWithName(n) for n in ("clamped", "load", "nobody")]
                   + [f.getNodesWithName(n) for n in ("clamped", "load", "nobody")])
    for k, (a, b) in enumerate(zip(*res)):
        assert np.array_equal(a, b), "query %d differs" % k
    assert len(res[1][2]) > 0 and len(res[1][5]) > 0 and len(res[1][7]) == 0


def test_name_queries_unnamed(impl, ref_lib):
    """a NULL name asks for the entities that carry no name (:5785, :5991)"""
    conn = util.box_conn()
    res = []
    for lib in (ref_lib, impl):
        f = OctForest(order=2, lib=lib)
        f.setTrilinearTopology(conn, _warped_points(conn))
        f.createTrees(1)
        f.refine(util.synth_flags(f.getOctants().as_array(), 9, 30))
        f.balance(0)
        f.createNodes()
        res.append((f.getOctsWithName(None), f.getNodesWithName(None)))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert len(res[1][0]) == len(f.getOctants().as_array())  # every volume is unnamed


def test_text_writers(impl, ref_lib, tmp_path):
    """writeToVTK / writeToTecplot (super-mesh) and writeForestToVTK (every
    octant as a brick) (reference src/TMROctForest.cpp:1149-1384): the files
    of the unmodified reference, byte for byte."""
    conn = util.box_conn()
    xpts = _warped_points(conn)
    out = []
    for tag, lib in (("ref", ref_lib), ("impl", impl)):
        f = OctForest(order=2, lib=lib)
        f.setTrilinearTopology(conn, xpts)
        f.createTrees(1)
        f.refine(util.synth_flags(f.getOctants().as_array(), 3, 30))
        f.balance(1)
        files = []
        for name, write in (("mesh.vtk", f.writeToVTK), ("mesh.dat", f.writeToTecplot),
                            ("forest.vtk", f.writeForestToVTK)):
            path = tmp_path / (tag + "_" + name)
            write(path)
            files.append(path.read_bytes())
        out.append(files)
    for a, b in zip(*out):
        assert len(a) > 200 and a == b
