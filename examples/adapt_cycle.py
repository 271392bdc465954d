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


if __name__ == "__main__":
    main()
