"""Generate the committed golden vectors from the oracle (the unmodified
reference sources built by oracle/Makefile).  Run in the build container:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; tests there compare against these files
(and against oracle/_ref, which travels with the snapshot as a built .so)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import util  # noqa: E402
from oracle import ref_loader  # noqa: E402

# name, conn, level, passes, pct, corner, order, interpolation type
# (1 = Gauss-Lobatto, 2 = Bernstein points)
CASES = [
    ("single_l2_p3_c0_o2", "single", 2, 3, 30, 0, 2, 1),
    ("box7_l2_p2_c1_o2", "box7", 2, 2, 30, 1, 2, 1),
    ("box7_l1_p2_c1_o3", "box7", 1, 2, 30, 1, 3, 1),
    ("connector15_l1_p3_c0_o2", "connector15", 1, 3, 30, 0, 2, 1),
    ("butterfly2_l1_p2_c1_o2", "butterfly2", 1, 2, 30, 1, 2, 1),
    ("box7_l1_p2_c1_o3_bernstein", "box7", 1, 2, 30, 1, 3, 2),
    ("connector15_l0_p2_c1_o4", "connector15", 0, 2, 30, 1, 4, 1),
]


def main():
    lib = ref_loader.load()
    for name, conn_name, level, passes, pct, corner, order, interp in CASES:
        conn = util.CONNS[conn_name]()
        rec = []
        f = util.build_forest(lib, conn, level, passes, pct, corner, order,
                              interp=interp, record=rec)
        res = util.node_results(f)
        coarse = f.coarsen() if order == 2 else f.duplicate()
        if order == 2:
            coarse.balance(1)
        else:
            coarse.setMeshOrder(2, interp)
        vec = f.createInterpolation(coarse)
        rows, rowp, cols, vals = vec.get()
        out = {
            "block_conn": conn,
            "params": np.array([level, passes, pct, corner, order]),
            "interp_type": np.array(interp),
            "counts": np.array([len(r[1]) for r in rec]),
            "checksums": np.array([util.checksum(r[1]) for r in rec], dtype=np.uint64),
            "checksum": np.uint64(util.checksum(res["octants"])),
            "octants": res["octants"],
            "conn": res["conn"],
            "node_numbers": res["node_numbers"],
            "dep_ptr": res["dep"][0], "dep_conn": res["dep"][1], "dep_weights": res["dep"][2],
            "interp_rows": rows, "interp_rowp": rowp, "interp_cols": cols, "interp_vals": vals,
        }
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, len(res["octants"]), "octants", os.path.getsize(path) // 1024, "KB")
    # fingerprints pinned at survey time on the same oracle
    fp = {
        "C1": {"recipe": "nb=1, createTrees(4), 4 passes pct=30, balance(0), order 2",
               "counts": [13077, 54097, 243965, 1027916], "checksum": "55e9487c98c7a2ff",
               "owned_nodes": 652025, "dep_nodes": 908576, "dep_nnz": 2441728,
               "local_nodes": 1560601},
        "C2": {"recipe": "nb=8, createTrees(3), 4 passes pct=35, balance(0), order 2",
               "octants": 86278900, "checksum": "da1d7223d950ef5c", "owned_nodes": 53774081},
        "C2_pct30": {"counts_after_pass_2_3_4": [3387273, 15408464, 64740026],
                     "checksums": ["d80cc85220bc9c0e", "9bd697a577179487", "bae29ce1fadfa5d9"],
                     "owned_nodes": 39894697},
    }
    # C4 (BASELINE configs[3]): the full 1050-tree butterfly lattice with all 8
    # face orientations and edge valence 2/3/4/8, computed here on the oracle
    conn = util.butterfly_conn(5, 5, 6)
    rec = []
    f = util.build_forest(lib, conn, 1, 2, 30, 1, 2, record=rec)
    res = util.node_results(f)
    fp["C4_1050_trees"] = {
        "recipe": "butterfly_conn(5,5,6) = 1050 trees, createTrees(1), 2 passes pct=30, "
                  "balance(1), order 2",
        "trees": int(len(conn)),
        "counts": [len(r[1]) for r in rec if r[0].startswith("balance")],
        "checksum": "%016x" % util.checksum(res["octants"]),
        "owned_nodes": int(f.getNumOwnedNodes()),
        "dep_nodes": int(len(res["dep"][0]) - 1),
        "dep_nnz": int(len(res["dep"][1])),
        "local_nodes": int(len(res["node_numbers"])),
        "conn_checksum": "%016x" % (int(np.sum(res["conn"].astype(np.int64).ravel() *
                                               (np.arange(res["conn"].size) % 1000003 + 1)))
                                    & 0xFFFFFFFFFFFFFFFF),
    }
    with open(os.path.join(HERE, "fingerprints.json"), "w") as fh:
        json.dump(fp, fh, indent=1)


if __name__ == "__main__":
    main()
