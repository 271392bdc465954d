#!/usr/bin/env python3
"""Time-boxed fuzz of the CUDA library against the oracle (no torch import, so
it starts in seconds): random operation sequences and randomly re-oriented
tree boxes, the same generators as tests/test_fuzz.py, until --seconds are up.

  gpurun --timeout 60 -- 'timeout 50 python tools/gpu/quick_fuzz.py --seconds 35 > gpurun_out/quick_fuzz.log 2>&1'
"""
import argparse
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import util  # noqa: E402
import test_fuzz  # noqa: E402
import tmr_b200  # noqa: E402
from oracle import ref_loader  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=35.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--emu", action="store_true",
                    help="dry run of this script on the test-only emulation (no GPU)")
    a = ap.parse_args()
    t0 = time.time()
    if a.emu:
        import ctypes
        from tmr_b200 import _capi
        gpu = _capi.bind(ctypes.CDLL(os.path.join(ROOT, "tests/emu/_build/libtmr_emu.so")))
    else:
        gpu = tmr_b200.require_gpu()
    ref = ref_loader.load()
    rng = random.Random(a.seed)
    ok = bad = 0
    while time.time() - t0 < a.seconds:
        dims = rng.choice([(2, 2, 2), (3, 2, 1), (2, 2, 1), (3, 3, 1)])
        conn = util.scrambled_conn(*dims, rng)
        level, order, corner = rng.choice([0, 1, 2]), rng.choice([2, 2, 3, 4]), rng.choice([0, 1])
        steps = [(rng.choice(["refine", "refine", "refine_neg", "coarsen", "dup", "refine_clamp"]),
                  rng.choice([15, 35, 60]), rng.randrange(1, 10 ** 6)) for _ in range(rng.choice([2, 3]))]
        case = (dims, level, order, corner, steps)
        try:
            res = []
            for lib in (ref, gpu):
                f = util.build_forest(lib, conn, level, 0, 0, corner, order)
                for kind, pct, seed in steps:
                    if f.getOctants().as_array().shape[0] > 20000 and kind.startswith("refine"):
                        kind = "coarsen"
                    f = test_fuzz._apply_step(f, kind, pct, seed, corner)
                r = util.node_results(f)
                coarse = f.coarsen() if order == 2 else f.duplicate()
                if order == 2:
                    coarse.balance(1)
                else:
                    coarse.setMeshOrder(order - 1)
                res.append((f.getOctants().as_array().copy(), r,
                            f.createInterpolation(coarse).get()))
            util.assert_octants_equal(res[0][0], res[1][0], "octants")
            util.assert_nodes_equal(res[0][1], res[1][1], "nodes")
            for x, y in zip(res[0][2][:3], res[1][2][:3]):
                assert np.array_equal(x, y), "interp structure"
            np.testing.assert_allclose(res[1][2][3], res[0][2][3], rtol=1e-12, atol=1e-300)
            ok += 1
        except Exception as e:  # noqa: BLE001
            bad += 1
            print("MISMATCH", case, conn.tolist(), str(e)[:300].replace("\n", " "), flush=True)
    print("quick_fuzz: %d cases equal, %d differ, %.1f s" % (ok, bad, time.time() - t0), flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
