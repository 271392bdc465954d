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
