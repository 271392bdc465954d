"""Python mirror of the reference's Cython `OctForest` / `OctantArray` classes
(reference tmr/TMR.pyx:3270-3790) on top of include/tmr_capi.h.

Method names, argument meaning and return shapes follow TMR.pyx so that a
script written against `tmr.TMR.OctForest` reads the same here:

    forest = OctForest(order=2)
    forest.setConnectivity(conn)          # (nblocks, 8) int32
    forest.createTrees(4)
    forest.refine(flags); forest.balance(0); forest.createNodes()
    conn = forest.getMeshConn(); ptr, dconn, w = forest.getDepNodeConn()

Differences forced by the missing toolchain (no mpi4py / tacs in this image):
the `comm` argument is dropped (the library is bound to the process-wide rank
set up by `tmr_b200.dist`), and `createInterpolation` fills a `VecInterp`
recorder instead of a TACS `VecInterp`.
"""
import ctypes as C

import numpy as np

from . import _capi

UNIFORM_POINTS = 0
GAUSS_LOBATTO_POINTS = 1
BERNSTEIN_POINTS = 2


class OctantArray:
    """Snapshot of a forest's local octants (structured numpy array of the
    reference's 24-byte records).  Mirrors reference tmr/TMR.pyx:3270-3332."""

    def __init__(self, records):
        self.records = records

    def __len__(self):
        return len(self.records)

    def __getitem__(self, k):
        return self.records[k]

    def as_array(self):
        return self.records


class VecInterp:
    """Recorder for the (row, cols, weights) stream createInterpolation emits
    through TACSBVecInterp::addInterp (reference src/TMROctForest.cpp:6683)."""

    def __init__(self, lib):
        self._lib = lib
        self._ptr = lib.tmrc_interp_create()

    def __del__(self):
        if getattr(self, "_ptr", None):
            self._lib.tmrc_interp_destroy(self._ptr)
            self._ptr = None

    def get(self):
        """rows (call order), rowp, cols, vals as numpy arrays."""
        nrows, nnz = C.c_int(0), C.c_int(0)
        rows, rowp, cols = (C.POINTER(C.c_int)() for _ in range(3))
        vals = C.POINTER(C.c_double)()
        self._lib.tmrc_interp_get(
            self._ptr, C.byref(nrows), C.byref(nnz), C.byref(rows), C.byref(rowp),
            C.byref(cols), C.byref(vals))
        return (
            _capi.as_int_array(rows, nrows.value),
            _capi.as_int_array(rowp, nrows.value + 1),
            _capi.as_int_array(cols, nnz.value),
            _capi.as_double_array(vals, nnz.value),
        )


class OctForest:
    """Forest of octrees behind the TMROctForest C++ class
    (reference src/TMROctForest.h:46-181)."""

    def __init__(self, order=2, interp=GAUSS_LOBATTO_POINTS, lib=None, _ptr=None,
                 comm_self=False):
        """comm_self: construct on MPI_COMM_SELF (an unpartitioned forest, also
        inside a multi-rank job) instead of the world communicator."""
        if lib is None:
            from . import load_library

            lib = load_library()
        self._lib = lib
        create = lib.tmrc_forest_create_self if comm_self else lib.tmrc_forest_create
        self._ptr = _ptr if _ptr is not None else create(order, interp)
        if not self._ptr:
            raise RuntimeError("TMROctForest construction failed")

    def __del__(self):
        if getattr(self, "_ptr", None):
            self._lib.tmrc_forest_destroy(self._ptr)
            self._ptr = None

    # -- configuration ---------------------------------------------------
    def setMeshOrder(self, order, interp=GAUSS_LOBATTO_POINTS):
        self._lib.tmrc_set_mesh_order(self._ptr, order, interp)

    def getMeshOrder(self):
        return self._lib.tmrc_get_mesh_order(self._ptr)

    def getInterpType(self):
        return self._lib.tmrc_get_interp_type(self._ptr)

    def setConnectivity(self, conn):
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        if conn.ndim != 2 or conn.shape[1] != 8:
            raise ValueError("conn must have shape (nblocks, 8)")
        num_nodes = int(conn.max()) + 1
        self._lib.tmrc_set_connectivity(
            self._ptr, num_nodes, conn.ctypes.data, conn.shape[0])

    # -- octant operations -----------------------------------------------
    def repartition(self, max_rank=-1):
        self._lib.tmrc_repartition(self._ptr, max_rank)

    def createTrees(self, depth=0):
        self._lib.tmrc_create_trees(self._ptr, depth)

    def createRandomTrees(self, nrand=10, min_lev=0, max_lev=8):
        self._lib.tmrc_create_random_trees(self._ptr, nrand, min_lev, max_lev)

    def refine(self, refine=None, min_lev=0, max_lev=30):
        if refine is None:
            self._lib.tmrc_refine(self._ptr, None, min_lev, max_lev)
            return
        refine = np.ascontiguousarray(refine, dtype=np.int32)
        n = self._lib.tmrc_num_octants(self._ptr)
        if refine.shape[0] != n:
            raise ValueError(
                "refine array has %d entries, forest has %d octants"
                % (refine.shape[0], n))
        self._lib.tmrc_refine(self._ptr, refine.ctypes.data, min_lev, max_lev)

    def duplicate(self):
        return OctForest(lib=self._lib, _ptr=self._lib.tmrc_duplicate(self._ptr))

    def coarsen(self):
        return OctForest(lib=self._lib, _ptr=self._lib.tmrc_coarsen(self._ptr))

    def balance(self, btype=0):
        self._lib.tmrc_balance(self._ptr, btype)

    def createNodes(self):
        self._lib.tmrc_create_nodes(self._ptr)

    # -- results -----------------------------------------------------------
    def getNumOctants(self):
        return self._lib.tmrc_num_octants(self._ptr)

    def getOctants(self):
        n = self._lib.tmrc_num_octants(self._ptr)
        out = np.zeros(n, dtype=_capi.OCT_DTYPE)
        if n:
            self._lib.tmrc_get_octants(self._ptr, out.ctypes.data)
        return OctantArray(out)

    def writeOctants(self, records):
        """Bulk form of OctantArray.__setitem__ (reference tmr/TMR.pyx:3303)."""
        records = np.ascontiguousarray(records, dtype=_capi.OCT_DTYPE)
        self._lib.tmrc_write_octants(self._ptr, records.ctypes.data, len(records))

    def getNodeRange(self):
        ptr = C.POINTER(C.c_int)()
        size = self._lib.tmrc_get_owned_node_range(self._ptr, C.byref(ptr))
        return _capi.as_int_array(ptr, size + 1)

    def getMeshConn(self):
        ptr = C.POINTER(C.c_int)()
        nelems, nowned = C.c_int(0), C.c_int(0)
        self._lib.tmrc_get_node_conn(
            self._ptr, C.byref(ptr), C.byref(nelems), C.byref(nowned))
        order = self.getMeshOrder()
        npe = order ** 3
        return _capi.as_int_array(ptr, nelems.value * npe).reshape(-1, npe)

    def getNumOwnedNodes(self):
        ptr = C.POINTER(C.c_int)()
        nelems, nowned = C.c_int(0), C.c_int(0)
        self._lib.tmrc_get_node_conn(
            self._ptr, C.byref(ptr), C.byref(nelems), C.byref(nowned))
        return nowned.value

    def getDepNodeConn(self):
        ptr, conn = C.POINTER(C.c_int)(), C.POINTER(C.c_int)()
        w = C.POINTER(C.c_double)()
        ndep = self._lib.tmrc_get_dep_node_conn(
            self._ptr, C.byref(ptr), C.byref(conn), C.byref(w))
        p = _capi.as_int_array(ptr, ndep + 1) if ndep > 0 else np.zeros(1, np.int32)
        nnz = int(p[-1]) if ndep > 0 else 0
        return p, _capi.as_int_array(conn, nnz), _capi.as_double_array(w, nnz)

    def getNodeNumbers(self):
        ptr = C.POINTER(C.c_int)()
        n = self._lib.tmrc_get_node_numbers(self._ptr, C.byref(ptr))
        return _capi.as_int_array(ptr, n)

    def getExtPreOffset(self):
        return self._lib.tmrc_get_ext_pre_offset(self._ptr)

    def getInterpKnots(self):
        ptr = C.POINTER(C.c_double)()
        n = self._lib.tmrc_get_interp_knots(self._ptr, C.byref(ptr))
        return _capi.as_double_array(ptr, n)

    def evalInterp(self, pt, derivatives=0):
        """Shape functions at pt (3 values in [-1,1]); derivatives=1 adds the
        three first derivatives, 2 the six second derivatives as well
        (reference TMROctForest::evalInterp, tmr/TMR.pyx)."""
        npe = self.getMeshOrder() ** 3
        nout = {0: 1, 1: 4, 2: 10}[derivatives]
        out = [np.zeros(npe) for _ in range(nout)]
        pt = np.ascontiguousarray(pt, dtype=np.float64)
        ptrs = [a.ctypes.data for a in out] + [None] * (10 - nout)
        self._lib.tmrc_eval_interp(self._ptr, pt.ctypes.data, *ptrs)
        return out[0] if derivatives == 0 else out

    def getConnectivity(self):
        nb, nf, ne, nn = (C.c_int(0) for _ in range(4))
        bc, bfc, bec, bfi = (C.POINTER(C.c_int)() for _ in range(4))
        self._lib.tmrc_get_connectivity(
            self._ptr, C.byref(nb), C.byref(nf), C.byref(ne), C.byref(nn),
            C.byref(bc), C.byref(bfc), C.byref(bec), C.byref(bfi))
        return {
            "nblocks": nb.value, "nfaces": nf.value, "nedges": ne.value,
            "nnodes": nn.value,
            "block_conn": _capi.as_int_array(bc, 8 * nb.value),
            "block_face_conn": _capi.as_int_array(bfc, 6 * nb.value),
            "block_edge_conn": _capi.as_int_array(bec, 12 * nb.value),
            "block_face_ids": _capi.as_int_array(bfi, 6 * nb.value),
        }

    def getInverseConnectivity(self):
        c = self.getConnectivity()
        nbc, nbp, ebc, ebp, fbc, fbp = (C.POINTER(C.c_int)() for _ in range(6))
        self._lib.tmrc_get_inverse_connectivity(
            self._ptr, C.byref(nbc), C.byref(nbp), C.byref(ebc), C.byref(ebp),
            C.byref(fbc), C.byref(fbp))
        out = {}
        for name, conn, ptr, n in (("node", nbc, nbp, c["nnodes"]),
                                   ("edge", ebc, ebp, c["nedges"]),
                                   ("face", fbc, fbp, c["nfaces"])):
            p = _capi.as_int_array(ptr, n + 1)
            out[name + "_block_ptr"] = p
            out[name + "_block_conn"] = _capi.as_int_array(conn, int(p[-1]) if n else 0)
        return out

    def transformNodes(self, records, edge_dir=-1):
        rec = np.array(records, dtype=_capi.OCT_DTYPE, copy=True)
        n = len(rec)
        rev = np.zeros(n, dtype=np.int32)
        fid = np.zeros(n, dtype=np.int32)
        self._lib.tmrc_transform_nodes(
            self._ptr, rec.ctypes.data, n, edge_dir, rev.ctypes.data, fid.ctypes.data)
        return rec, rev, fid

    def findEnclosing(self, order, knots, records):
        rec = np.ascontiguousarray(records, dtype=_capi.OCT_DTYPE)
        knots = np.ascontiguousarray(knots, dtype=np.float64)
        n = len(rec)
        idx = np.zeros(n, dtype=np.int32)
        own = np.zeros(n, dtype=np.int32)
        self._lib.tmrc_find_enclosing(
            self._ptr, order, knots.ctypes.data, rec.ctypes.data, n,
            idx.ctypes.data, own.ctypes.data)
        return idx, own

    def distributeOctants(self, records, nranks, use_tags=0, include_local=0,
                          use_node_index=0, cap=None):
        """TMROctForest::distributeOctants on a host list; returns (received
        records, oct_ptr, recv_ptr)."""
        rec = np.ascontiguousarray(records, dtype=_capi.OCT_DTYPE)
        cap = cap if cap is not None else max(16, 64 * len(rec) + 1024)
        out = np.zeros(cap, dtype=_capi.OCT_DTYPE)
        optr = np.zeros(nranks + 1, dtype=np.int32)
        rptr = np.zeros(nranks + 1, dtype=np.int32)
        n = self._lib.tmrc_distribute_octants(
            self._ptr, rec.ctypes.data, len(rec), use_tags, include_local,
            use_node_index, out.ctypes.data, cap, optr.ctypes.data, rptr.ctypes.data)
        return out[:n].copy(), optr, rptr

    def sendOctants(self, records, oct_ptr, recv_ptr, use_node_index=0, cap=None):
        rec = np.ascontiguousarray(records, dtype=_capi.OCT_DTYPE)
        optr = np.ascontiguousarray(oct_ptr, dtype=np.int32)
        rptr = np.ascontiguousarray(recv_ptr, dtype=np.int32)
        cap = cap if cap is not None else int(rptr[-1]) + 16
        out = np.zeros(cap, dtype=_capi.OCT_DTYPE)
        n = self._lib.tmrc_send_octants(
            self._ptr, rec.ctypes.data, len(rec), optr.ctypes.data, rptr.ctypes.data,
            use_node_index, out.ctypes.data, cap)
        return out[:n].copy()

    def createInterpolation(self, coarse, vec=None):
        if vec is None:
            vec = VecInterp(self._lib)
        self._lib.tmrc_create_interpolation(self._ptr, coarse._ptr, vec._ptr)
        return vec

    def setTrilinearTopology(self, block_conn, xpts):
        """Attach a topology of trilinear hexahedra (one per tree) through the
        super-mesh node locations `xpts` (tmrc_set_trilinear_topology): gives
        the forest a geometry without the CAD layer, so that getPoints() runs
        evaluateNodeLocations (reference src/TMROctForest.cpp:5524-5675)."""
        conn = np.ascontiguousarray(block_conn, dtype=np.int32)
        x = np.ascontiguousarray(xpts, dtype=np.float64)
        self._lib.tmrc_set_trilinear_topology.argtypes = [
            C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        rc = self._lib.tmrc_set_trilinear_topology(
            self._ptr, len(x), conn.ctypes.data, len(conn), x.ctypes.data)
        if rc != 0:
            raise RuntimeError("tmrc_set_trilinear_topology failed")

    def getPoints(self):
        """Node locations, one row per local node in the order of
        getNodeNumbers() (reference getPoints :1476-1481)."""
        ptr = C.POINTER(C.c_double)()
        self._lib.tmrc_get_points.restype = C.c_int
        self._lib.tmrc_get_points.argtypes = [C.c_void_p, C.c_void_p]
        n = self._lib.tmrc_get_points(self._ptr, C.byref(ptr))
        if n <= 0 or not ptr:
            return np.zeros((0, 3))
        return _capi.as_double_array(ptr, 3 * n).reshape(n, 3).copy()

    # ---- name queries -----------------------------------------------------------
    VERTEX, EDGE, FACE, VOLUME = 0, 1, 2, 3

    def setEntityName(self, kind, index, name):
        """Name a vertex / edge / face / volume of the topology attached by
        setTrilinearTopology (index in the numbering of getConnectivity)."""
        self._lib.tmrc_set_entity_name.restype = C.c_int
        self._lib.tmrc_set_entity_name.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
        raw = None if name is None else name.encode()
        if self._lib.tmrc_set_entity_name(self._ptr, kind, index, raw) != 0:
            raise ValueError("no entity of kind %d with index %d" % (kind, index))

    def getOctsWithName(self, name):
        """Local octants of a named volume, or touching a named tree face
        (info = face index) (reference getOctsWithName :5747-5862)."""
        f = self._lib.tmrc_get_octs_with_name
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
        raw = None if name is None else name.encode()
        n = f(self._ptr, raw, None, 0)  # count, then fetch
        if n < 0:
            raise RuntimeError("getOctsWithName: no topology or no octants")
        out = np.zeros(max(n, 1), dtype=_capi.OCT_DTYPE)
        n = f(self._ptr, raw, out.ctypes.data, n)
        return out[:n].copy()

    def getNodesWithName(self, name):
        """Sorted unique numbers of the local nodes on named vertices, edges
        and faces (reference getNodesWithName :5882-6203)."""
        f = self._lib.tmrc_get_nodes_with_name
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
        raw = None if name is None else name.encode()
        n = f(self._ptr, raw, None, 0)
        if n < 0:
            raise RuntimeError("getNodesWithName: no topology or no nodes")
        out = np.zeros(max(n, 1), dtype=np.int32)
        n = f(self._ptr, raw, out.ctypes.data, n)
        return out[:n].copy()

    # ---- text writers (reference :1149-1384) -------------------------------------
    def _write(self, which, filename):
        self._lib.tmrc_write.restype = None
        self._lib.tmrc_write.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        self._lib.tmrc_write(self._ptr, which, str(filename).encode())

    def writeToVTK(self, filename):
        self._write(0, filename)

    def writeToTecplot(self, filename):
        self._write(1, filename)

    def writeForestToVTK(self, filename):
        self._write(2, filename)

    # ---- B200 extensions (include/tmr_b200_ext.h, include/tmrgpu.h) ----------
    def createInterpolationCSR(self, coarse):
        """The whole prolongation in one hand-off: (rows, rowp, cols, vals),
        rows in the order createInterpolation() emits its addInterp calls
        (tmr_b200_create_interpolation_csr; not available on the oracle)."""
        lib = self._lib
        rows, rowp, cols = (C.POINTER(C.c_int)() for _ in range(3))
        vals = C.POINTER(C.c_double)()
        nnz = C.c_int(0)
        lib.tmr_b200_create_interpolation_csr.restype = C.c_int
        lib.tmr_b200_create_interpolation_csr.argtypes = [C.c_void_p, C.c_void_p] + [C.c_void_p] * 5
        n = lib.tmr_b200_create_interpolation_csr(
            self._ptr, coarse._ptr, C.byref(rows), C.byref(rowp), C.byref(cols),
            C.byref(vals), C.byref(nnz))
        # borrowed views into arrays owned by this forest (valid until its next
        # createInterpolationCSR / destruction): no copy of what can be 18 GB
        return (_capi.view_array(rows, n, np.int32), _capi.view_array(rowp, n + 1, np.int32),
                _capi.view_array(cols, nnz.value, np.int32),
                _capi.view_array(vals, nnz.value, np.float64))

    def assemblerViews(self):
        """Copies of the DEVICE arrays tmrgpu_assembler_views exposes (the
        arrays createTACS hands to TACSAssembler, reference
        src/TMR_TACSCreator.cpp:332-461), for checking them against the host
        getters; a GPU consumer would use the pointers where they are."""
        lib = self._lib

        class View(C.Structure):
            _fields_ = [(n, C.c_int64) for n in ("num_elements", "num_owned_nodes",
                                                 "num_dep_nodes", "num_local_nodes", "dep_nnz")]
            _fields_ += [("order", C.c_int)]
            _fields_ += [(n, C.c_void_p) for n in ("elem_ptr", "conn", "dep_ptr", "dep_conn",
                                                   "dep_weights", "node_numbers")]

        lib.tmr_b200_device_forest.restype = C.c_void_p
        lib.tmr_b200_device_forest.argtypes = [C.c_void_p]
        lib.tmr_b200_context.restype = C.c_void_p
        dev = C.c_void_p(lib.tmr_b200_device_forest(self._ptr))
        ctx = C.c_void_p(lib.tmr_b200_context())
        v = View()
        lib.tmrgpu_assembler_views.argtypes = [C.c_void_p, C.c_void_p]
        if lib.tmrgpu_assembler_views(dev, C.byref(v)) != 0:
            raise RuntimeError("tmr_b200: no node data on the device (call createNodes)")
        lib.tmrgpu_copy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]

        def pull(ptr, n, dtype):
            out = np.zeros(n, dtype=dtype)
            if n:
                lib.tmrgpu_copy_d2h(ctx, out.ctypes.data, ptr, out.nbytes)
            return out

        npe = v.order ** 3
        return {
            "num_elements": v.num_elements, "num_owned_nodes": v.num_owned_nodes,
            "num_dep_nodes": v.num_dep_nodes, "order": v.order,
            "elem_ptr": pull(v.elem_ptr, v.num_elements + 1, np.int32),
            "conn": pull(v.conn, v.num_elements * npe, np.int32),
            "dep_ptr": pull(v.dep_ptr, v.num_dep_nodes + 1, np.int32),
            "dep_conn": pull(v.dep_conn, v.dep_nnz, np.int32),
            "dep_weights": pull(v.dep_weights, v.dep_nnz, np.float64),
            "node_numbers": pull(v.node_numbers, v.num_local_nodes, np.int32),
        }


def array_sort(lib, records, use_node_index=0):
    """TMROctantArray::sort (reference src/TMROctant.cpp:357-399)."""
    rec = np.array(records, dtype=_capi.OCT_DTYPE, copy=True)
    n = lib.tmrc_array_sort(rec.ctypes.data, len(rec), use_node_index)
    return rec[:n].copy()


def array_contains(lib, records, queries, use_node_index=0, use_position=0):
    """TMROctantArray::contains (reference src/TMROctant.cpp:404-424)."""
    rec = np.ascontiguousarray(records, dtype=_capi.OCT_DTYPE)
    q = np.ascontiguousarray(queries, dtype=_capi.OCT_DTYPE)
    out = np.zeros(len(q), dtype=np.int32)
    lib.tmrc_array_contains(rec.ctypes.data, len(rec), use_node_index,
                            q.ctypes.data, len(q), use_position, out.ctypes.data)
    return out
