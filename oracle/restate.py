"""TEST INFRASTRUCTURE ONLY -- a small pure-Python restatement of the
reference's octant algorithms, written from the reference's formulation (hash
of 0-sibling families + work queue), NOT from the GPU formulation, so that it
is an independent check of both.  Each function cites the reference code it
restates.  Pinned against oracle/_ref and tests/golden by
tests/test_restatement.py.  Pure Python loops: small forests only.

Only tests/ may import this module.
"""
HMAX = 1 << 30


def h_of(level):
    return 1 << (30 - level)


def morton_key(o):
    """Total order of reference src/TMROctant.cpp:171-204: block, then Morton
    with x most significant, then level.  o = (block, x, y, z, level)."""
    b, x, y, z, lev = o
    m = 0
    for bit in range(29, -1, -1):
        m = (m << 3) | (((x >> bit) & 1) << 2) | (((y >> bit) & 1) << 1) | ((z >> bit) & 1)
    return (b, m, lev)


def sort_unique(octs):
    """TMROctantArray::sort, element mode (src/TMROctant.cpp:357-399): sort,
    then keep the LAST (finest) of every run of equal position."""
    s = sorted(set(octs), key=morton_key)
    out = []
    for o in s:
        if out and out[-1][:4] == o[:4]:
            out[-1] = o
        else:
            out.append(o)
    return out


def sibling0(o):
    """getSibling(0) (src/TMROctant.cpp:42-58), two's-complement safe."""
    b, x, y, z, lev = o
    h = h_of(lev)
    return (b, x - h if x & h else x, y - h if y & h else y, z - h if z & h else z, lev)


def parent(o):
    b, x, y, z, lev = o
    if lev == 0:
        return o
    h = h_of(lev)
    return (b, x & ~h, y & ~h, z & ~h, lev - 1)


def refine(octs, flags, min_level=0, max_level=30):
    """TMROctForest::refine (src/TMROctForest.cpp:2169-2329), one rank."""
    out = set()
    for o, r in zip(octs, flags):
        b, x, y, z, lev = o
        if r == 0 or (r > 0 and lev >= max_level) or (r < 0 and lev <= min_level):
            out.add(o)
        elif r < 0:
            nl = max(lev + r, min_level)
            h = h_of(nl)
            out.add((b, x - x % h, y - y % h, z - z % h, nl))
        else:
            nl = min(lev + r, max_level)
            ref = 1 << (nl - lev - 1)
            h = h_of(nl)
            for ii in range(ref):
                for jj in range(ref):
                    for kk in range(ref):
                        out.add((b, x + 2 * ii * h, y + 2 * jj * h, z + 2 * kk * h, nl))
    return sort_unique(out)


# orientation ids (src/TMROctForest.cpp:79-143): get = to owner, set = from owner
def _get_face(fid, M, x, y):
    return [(x, y), (M - y, x), (M - x, M - y), (y, M - x), (y, x), (x, M - y),
            (M - y, M - x), (M - x, y)][fid]


def _set_face(fid, M, u, v):
    return [(u, v), (v, M - u), (M - u, M - v), (M - v, u), (v, u), (u, M - v),
            (M - v, M - u), (M - u, v)][fid]


EDGE_NODES = [(0, 1), (2, 3), (4, 5), (6, 7), (0, 2), (1, 3), (4, 6), (5, 7),
              (0, 4), (1, 5), (2, 6), (3, 7)]


def _face_images(T, q, face):
    """addFaceNeighbors (src/TMROctForest.cpp:2525-2590); q at level L, box 2h."""
    b, x, y, z, lev = q
    h2 = 2 * h_of(lev)
    M = HMAX - h2
    fno = T["block_face_conn"][6 * b + face]
    fid = T["block_face_ids"][6 * b + face]
    ab = (y, z) if face < 2 else ((x, z) if face < 4 else (x, y))
    u, v = _get_face(fid, M, *ab)
    out = []
    for ip in range(T["face_block_ptr"][fno], T["face_block_ptr"][fno + 1]):
        adj, af = divmod(int(T["face_block_conn"][ip]), 6)
        if adj == b:
            continue
        a2, b2 = _set_face(T["block_face_ids"][6 * adj + af], M, u, v)
        n = M * (af % 2)
        out.append((adj,) + ((n, a2, b2) if af < 2 else ((a2, n, b2) if af < 4 else (a2, b2, n))) + (lev,))
    return out


def _edge_images(T, q, e):
    """addEdgeNeighbors (src/TMROctForest.cpp:2608-2686)."""
    b, x, y, z, lev = q
    M = HMAX - 2 * h_of(lev)
    eno = T["block_edge_conn"][12 * b + e]
    u = x if e < 4 else (y if e < 8 else z)
    n1, n2 = (T["block_conn"][8 * b + c] for c in EDGE_NODES[e])
    out = []
    for ip in range(T["edge_block_ptr"][eno], T["edge_block_ptr"][eno + 1]):
        adj, ae = divmod(int(T["edge_block_conn"][ip]), 12)
        if adj == b:
            continue
        m1, m2 = (T["block_conn"][8 * adj + c] for c in EDGE_NODES[ae])
        uu = M - u if (n1 == m2 and n2 == m1) else u
        s = ae % 4
        t1, t2 = M * (s % 2), M * (s // 2)
        out.append((adj,) + ((uu, t1, t2) if ae < 4 else ((t1, uu, t2) if ae < 8 else (t1, t2, uu))) + (lev,))
    return out


def _corner_images(T, q, c):
    """addCornerNeighbors (src/TMROctForest.cpp:2704-2745)."""
    b, x, y, z, lev = q
    M = HMAX - 2 * h_of(lev)
    node = T["block_conn"][8 * b + c]
    out = []
    for ip in range(T["node_block_ptr"][node], T["node_block_ptr"][node + 1]):
        adj, ac = divmod(int(T["node_block_conn"][ip]), 8)
        if adj != b:
            out.append((adj, M * (ac % 2), M * ((ac % 4) // 2), M * (ac // 4), lev))
    return out


def balance(octs, T, corner):
    """TMROctForest::balance + balanceOctant (src/TMROctForest.cpp:2763-3089),
    one rank: hash of 0-siblings, queue ripple, sibling expansion, sort/uniq."""
    seen, queue = set(), []

    def add(q):
        if q not in seen:
            seen.add(q)
            queue.append(q)

    for o in octs:
        add(sibling0(o))
    dirs = [(dx, dy, dz) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)
            if (dx, dy, dz) != (0, 0, 0) and (corner or abs(dx) + abs(dy) + abs(dz) < 3)]
    while queue:
        o = queue.pop()
        if o[4] <= 1:
            continue
        p = parent(o)
        b, x, y, z, lev = p
        h = h_of(lev)
        for dx, dy, dz in dirs:
            q = sibling0((b, x + dx * h, y + dy * h, z + dz * h, lev))
            ex, ey, ez = [not (0 <= c < HMAX) for c in q[1:4]]
            nout = ex + ey + ez
            if nout == 0:
                add(q)
            elif nout == 1:
                face = (0 if q[1] < 0 else 1) if ex else ((2 if q[2] < 0 else 3) if ey else (4 if q[3] < 0 else 5))
                for im in _face_images(T, q, face):
                    add(im)
            elif nout == 2:
                if ey and ez:
                    e = (0 if q[2] < 0 else 1) + (0 if q[3] < 0 else 2)
                elif ex and ez:
                    e = (4 if q[1] < 0 else 5) + (0 if q[3] < 0 else 2)
                else:
                    e = (8 if q[1] < 0 else 9) + (0 if q[2] < 0 else 2)
                for im in _edge_images(T, q, e):
                    add(im)
            else:
                c = (0 if q[1] < 0 else 1) + (0 if q[2] < 0 else 2) + (0 if q[3] < 0 else 4)
                for im in _corner_images(T, q, c):
                    add(im)
    out = set()
    for f in seen:
        b, x, y, z, lev = f
        if lev == 0:
            out.add(f)
        else:
            h = h_of(lev)
            for c in range(8):
                out.add((b, x + h * (c & 1), y + h * ((c >> 1) & 1), z + h * (c >> 2), lev))
    return sort_unique(out)


def to_tuples(records):
    return [(int(r["block"]), int(r["x"]), int(r["y"]), int(r["z"]), int(r["level"]))
            for r in records]
