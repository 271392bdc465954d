"""Shared helpers for the parity tests: super-mesh fixtures, the deterministic
refinement recipe of SURVEY.md section 8(d), and result comparison."""
import numpy as np

from tmr_b200 import _capi
from tmr_b200.forest import OctForest

U64 = np.uint64


def _sm(x):
    x = x + U64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> U64(30))) * U64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> U64(27))) * U64(0x94D049BB133111EB)
    return x ^ (x >> U64(31))


def record_hash(octs, seed):
    """splitmix64 chain over (block, x, y, z, level) -- SURVEY.md 8(d)."""
    with np.errstate(over="ignore"):
        k = _sm(U64(seed) ^ octs["block"].astype(np.uint32).astype(U64))
        for f in ("x", "y", "z"):
            k = _sm(k ^ octs[f].astype(np.uint32).astype(U64))
        k = _sm(k ^ octs["level"].astype(np.uint16).astype(U64))
    return k


def checksum(octs):
    with np.errstate(over="ignore"):
        return int(record_hash(octs, 0).sum(dtype=U64))


def synth_flags(octs, seed, pct):
    return (record_hash(octs, seed) % U64(100) < U64(pct)).astype(np.int32)


# ---- super-mesh fixtures ---------------------------------------------------
def single_conn():
    return np.arange(8, dtype=np.int32).reshape(1, 8)


def rectangle_conn():
    # reference examples/parallel/octant_test.cpp:29-30
    return np.array([0, 1, 3, 4, 6, 7, 9, 10, 8, 11, 2, 5, 7, 10, 1, 4],
                    dtype=np.int32).reshape(2, 8)


def box_conn():
    # reference examples/parallel/octant_test.cpp:52-55 (7 trees, all 8 face
    # orientation ids occur)
    return np.array(
        [0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 0, 1, 9, 11, 4, 5, 5, 11, 1, 10, 7, 15,
         3, 14, 7, 15, 3, 14, 6, 13, 2, 12, 9, 13, 4, 6, 8, 12, 0, 2, 10, 14,
         8, 12, 1, 3, 0, 2, 4, 5, 6, 7, 9, 11, 13, 15],
        dtype=np.int32).reshape(7, 8)


def connector_conn():
    # reference examples/parallel/octant_test.cpp:83-90 (15 trees)
    return np.array(
        [0, 1, 2, 3, 26, 27, 28, 29, 0, 2, 8, 4, 26, 28, 34, 30, 3, 1, 5, 9,
         29, 27, 31, 35, 4, 5, 6, 7, 30, 31, 32, 33, 6, 7, 10, 11, 32, 33, 36,
         37, 8, 4, 10, 6, 34, 30, 36, 32, 7, 5, 11, 9, 33, 31, 37, 35, 10, 11,
         12, 13, 36, 37, 38, 39, 12, 13, 18, 19, 38, 39, 44, 45, 14, 12, 16,
         18, 40, 38, 42, 44, 13, 15, 19, 17, 39, 41, 45, 43, 14, 16, 20, 22,
         40, 42, 46, 48, 16, 18, 22, 24, 42, 44, 48, 50, 19, 17, 25, 23, 45,
         43, 51, 49, 17, 15, 23, 21, 43, 41, 49, 47],
        dtype=np.int32).reshape(15, 8)


def structured_conn(nb, nby=None, nbz=None):
    """nb x nby x nbz box of trees (nb^3 by default), node id
    i + (nb+1)(j + (nby+1)k), corners in bit order."""
    nby = nb if nby is None else nby
    nbz = nb if nbz is None else nbz
    nx1, ny1 = nb + 1, nby + 1
    i, j, k = np.meshgrid(np.arange(nb), np.arange(nby), np.arange(nbz), indexing="ij")
    # block index runs x fastest
    order = np.lexsort((i.ravel(), j.ravel(), k.ravel()))
    i, j, k = i.ravel()[order], j.ravel()[order], k.ravel()[order]
    conn = np.zeros((nb * nby * nbz, 8), dtype=np.int32)
    for c in range(8):
        conn[:, c] = (i + (c & 1)) + nx1 * ((j + ((c >> 1) & 1)) + ny1 * (k + (c >> 2)))
    return conn


def butterfly_conn(cx, cy, cz):
    """Tile the 7-tree butterfly cell on a lattice:
    irregular valence and all 8 face orientations."""
    cell = box_conn()
    offs = {8: (0, 0, 0), 9: (0, 0, 1), 10: (1, 0, 0), 11: (1, 0, 1),
            12: (0, 1, 0), 13: (0, 1, 1), 14: (1, 1, 0), 15: (1, 1, 1)}
    nlat = (cx + 1) * (cy + 1) * (cz + 1)
    out = []
    c = 0
    for k in range(cz):
        for j in range(cy):
            for i in range(cx):
                for blk in cell:
                    row = []
                    for v in blk:
                        if v < 8:
                            row.append(nlat + 8 * c + v)
                        else:
                            di, dj, dk = offs[int(v)]
                            row.append((i + di) + (cx + 1) * ((j + dj) + (cy + 1) * (k + dk)))
                    out.append(row)
                c += 1
    return np.array(out, dtype=np.int32)


def scrambled_conn(nx, ny, nz, rng):
    """nx x ny x nz box of trees whose local corner numbering is, tree by tree,
    one of the 48 symmetries of the cube (axis permutation and flips), in a
    shuffled tree order: every relative orientation of two trees across a face
    or an edge occurs."""
    import itertools
    sym = [(p, f) for p in itertools.permutations(range(3)) for f in range(8)]
    conn = structured_conn(nx, ny, nz)
    out = np.zeros_like(conn)
    for b in range(conn.shape[0]):
        p, f = rng.choice(sym)
        for c in range(8):
            bits = [(c >> a) & 1 for a in range(3)]
            nb = [bits[p[a]] ^ ((f >> a) & 1) for a in range(3)]
            out[b, nb[0] | (nb[1] << 1) | (nb[2] << 2)] = conn[b, c]
    perm = list(range(conn.shape[0]))
    rng.shuffle(perm)
    return np.ascontiguousarray(out[perm])


CONNS = {
    "single": single_conn,
    "rectangle": rectangle_conn,
    "box7": box_conn,
    "connector15": connector_conn,
    "grid2": lambda: structured_conn(2),
    "butterfly2": lambda: butterfly_conn(2, 2, 2),
    # BASELINE configs[3]: 1050 trees, all 8 face orientations, edge valence 2/3/4/8
    "butterfly556": lambda: butterfly_conn(5, 5, 6),
}


# ---- pipeline ---------------------------------------------------------------
def build_forest(lib, conn, level, passes, pct, corner, order=2, seed=2024,
                 interp=1, record=None):
    """createTrees(level) then `passes` x {hash-driven refine, balance}."""
    f = OctForest(order=order, interp=interp, lib=lib)
    f.setConnectivity(conn)
    f.createTrees(level)
    for p in range(passes):
        octs = f.getOctants().as_array()
        f.refine(synth_flags(octs, seed + p, pct))
        if record is not None:
            record.append(("refine%d" % p, f.getOctants().as_array().copy()))
        f.balance(corner)
        if record is not None:
            record.append(("balance%d" % p, f.getOctants().as_array().copy()))
    return f


def node_results(f):
    f.createNodes()
    return {
        "octants": f.getOctants().as_array().copy(),
        "conn": f.getMeshConn(),
        "dep": f.getDepNodeConn(),
        "node_numbers": f.getNodeNumbers(),
        "node_range": f.getNodeRange(),
        "ext_pre": f.getExtPreOffset(),
        "num_owned": f.getNumOwnedNodes(),
    }


def assert_octants_equal(a, b, what=""):
    assert len(a) == len(b), "%s: %d vs %d octants" % (what, len(a), len(b))
    for fld in a.dtype.names:
        bad = np.nonzero(a[fld] != b[fld])[0]
        assert len(bad) == 0, "%s: field %s differs at %s" % (what, fld, bad[:8])


def assert_nodes_equal(a, b, what="", rtol=1e-12):
    """Bit-exact integers; weights within rtol (they are bit-equal in
    practice: same operation order, no FMA contraction)."""
    assert_octants_equal(a["octants"], b["octants"], what + " octants/info")
    assert np.array_equal(a["conn"], b["conn"]), what + ": conn differs"
    assert np.array_equal(a["node_numbers"], b["node_numbers"]), what + ": node numbers"
    assert np.array_equal(a["node_range"], b["node_range"]), what + ": node_range"
    assert a["num_owned"] == b["num_owned"], what + ": owned count"
    if a["num_owned"] > 0:
        # a rank that owns no node: the reference's bsearch for node_range[rank]
        # finds nothing and ext_pre_offset is the difference to a NULL pointer
        # (src/TMROctForest.cpp:4260-4264); here it is the lower bound
        assert a["ext_pre"] == b["ext_pre"], what + ": ext_pre_offset"
    pa, ca, wa = a["dep"]
    pb, cb, wb = b["dep"]
    assert np.array_equal(pa, pb), what + ": dep_ptr differs"
    assert np.array_equal(ca, cb), what + ": dep_conn differs"
    np.testing.assert_allclose(wa, wb, rtol=rtol, atol=0, err_msg=what + ": dep_weights")


def interp_rows(vec):
    """createInterpolation output as {row: (cols, vals)} plus the call order."""
    rows, rowp, cols, vals = vec.get()
    d = {}
    for r in range(len(rows)):
        d[int(rows[r])] = (cols[rowp[r]:rowp[r + 1]], vals[rowp[r]:rowp[r + 1]])
    return rows, d


def assert_interp_equal(va, vb, what="", rtol=1e-12):
    ra, da = interp_rows(va)
    rb, db = interp_rows(vb)
    assert np.array_equal(ra, rb), what + ": row emission order differs"
    for r in da:
        ca, wa = da[r]
        cb, wb = db[r]
        assert np.array_equal(ca, cb), "%s: row %d columns differ" % (what, r)
        np.testing.assert_allclose(wa, wb, rtol=rtol, atol=1e-300,
                                   err_msg="%s: row %d weights" % (what, r))


def random_octants(rng, n, nblocks, max_level):
    """Arbitrary (possibly overlapping / duplicated) octants for sort tests."""
    rec = np.zeros(n, dtype=_capi.OCT_DTYPE)
    lev = rng.integers(0, max_level + 1, n)
    rec["level"] = lev
    rec["block"] = rng.integers(0, nblocks, n)
    for f in ("x", "y", "z"):
        h = (1 << (30 - lev)).astype(np.int64)
        rec[f] = (rng.integers(0, 1 << 30, n) // h) * h
    rec["tag"] = np.arange(n)
    return rec
