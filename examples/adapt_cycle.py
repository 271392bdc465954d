#!/usr/bin/env python
"""The adaptation cycle of reference examples/parallel/octant_test.cpp:243-281
and tmr/TopOptUtils.py:69-110 with the B200 drop-in: build a forest on the
7-tree "box" super-mesh, refine by a flag array, balance, create nodes, coarsen
a multigrid level and build the prolongation.

    python examples/adapt_cycle.py            # needs a CUDA device
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import tmr_b200  # noqa: E402
import util  # noqa: E402


def main():
    tmr_b200.require_gpu()
    forest = tmr_b200.OctForest(order=2)
    forest.setConnectivity(util.box_conn())       # (nblocks, 8) int32
    forest.createTrees(3)
    for sweep in range(3):
        octs = forest.getOctants().as_array()
        # refine the octants whose centre lies inside a sphere (any int32 flags work)
        h = 1 << (30 - octs["level"].astype(np.int64))
        c = np.stack([octs[k] + h // 2 for k in ("x", "y", "z")], axis=1) / float(1 << 30)
        flags = (np.linalg.norm(c - 0.5, axis=1) < 0.35).astype(np.int32)
        forest.refine(flags)
        forest.balance(1)
        print("sweep %d: %d octants" % (sweep, forest.getNumOctants()))
    forest.createNodes()
    conn = forest.getMeshConn()
    ptr, dconn, w = forest.getDepNodeConn()
    print("elements %d, owned nodes %d, dependent nodes %d (%d stencil entries)"
          % (len(conn), forest.getNumOwnedNodes(), len(ptr) - 1, len(dconn)))
    coarse = forest.coarsen()
    coarse.balance(1)
    rows, rowp, cols, vals = forest.createInterpolation(coarse).get()
    print("prolongation: %d rows, %d non-zeros, max |row sum - 1| = %.1e"
          % (len(rows), len(cols), np.abs(np.add.reduceat(vals, rowp[:-1]) - 1).max()))

    # ---- B200 extensions (include/tmr_b200_ext.h, include/tmrgpu.h) ------------------
    # the whole prolongation in one hand-off instead of one addInterp call per row
    rows2, rowp2, cols2, vals2 = forest.createInterpolationCSR(coarse)
    assert np.array_equal(rows, rows2) and np.array_equal(cols, cols2)
    # the arrays createTACS hands to TACSAssembler, as they sit on the device
    view = forest.assemblerViews()
    assert np.array_equal(view["conn"].reshape(conn.shape), conn)
    print("device views: elem_ptr[-1] = %d, %d dependent stencil entries"
          % (view["elem_ptr"][-1], len(view["dep_conn"])))
    # a geometry without the CAD layer: one trilinear hexahedron per tree
    block_conn = util.box_conn()
    nodes = int(block_conn.max()) + 1
    xpts = np.random.default_rng(0).uniform(-1.0, 1.0, (nodes, 3)) + 4.0 * np.arange(nodes)[:, None]
    geo = tmr_b200.OctForest(order=2)
    geo.setTrilinearTopology(block_conn, xpts)
    geo.createTrees(2)
    geo.createNodes()
    X = geo.getPoints()           # evaluateNodeLocations on the GPU
    print("node locations: %d points, bounding box %s .. %s"
          % (len(X), np.round(X.min(axis=0), 2), np.round(X.max(axis=0), 2)))
    # boundary conditions go through names: clamp the nodes of one tree face
    face = int(geo.getConnectivity()["block_face_conn"][0])
    geo.setEntityName(tmr_b200.OctForest.FACE, face, "clamped")
    print("face %d named 'clamped': %d octants touch it, %d nodes lie on it"
          % (face, len(geo.getOctsWithName("clamped")), len(geo.getNodesWithName("clamped"))))


if __name__ == "__main__":
    main()
