"""The pure-Python restatement (oracle/restate.py, written from the reference's
hash+queue formulation) against the compiled reference and the goldens, and
the kernel bodies against the restatement: two independent derivations of
refine / balance must agree."""
import glob
import os

import numpy as np
import pytest

import util
from oracle import restate
from tmr_b200.forest import OctForest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def tables_of(forest):
    t = forest.getConnectivity()
    t.update(forest.getInverseConnectivity())
    return t


@pytest.mark.parametrize("conn_name,level,passes,corner", [
    ("single", 1, 3, 0), ("box7", 1, 2, 1), ("connector15", 0, 3, 0), ("rectangle", 1, 2, 1)])
def test_restatement_matches_reference_and_kernels(conn_name, level, passes, corner,
                                                   ref_lib, emu_lib):
    conn = util.CONNS[conn_name]()
    fr = OctForest(lib=ref_lib)
    fe = OctForest(lib=emu_lib)
    for f in (fr, fe):
        f.setConnectivity(conn)
        f.createTrees(level)
    T = tables_of(fr)
    octs = restate.to_tuples(fr.getOctants().as_array())
    rng = np.random.default_rng(5)
    for p in range(passes):
        rec = fr.getOctants().as_array()
        flags = util.synth_flags(rec, 2024 + p, 35)
        if p == 1:
            flags = rng.integers(-1, 3, len(rec)).astype(np.int32)
        octs = restate.refine(octs, flags.tolist())
        for f in (fr, fe):
            f.refine(flags)
        assert octs == restate.to_tuples(fr.getOctants().as_array())
        assert octs == restate.to_tuples(fe.getOctants().as_array())
        octs = restate.balance(octs, T, corner)
        for f in (fr, fe):
            f.balance(corner)
        assert octs == restate.to_tuples(fr.getOctants().as_array())
        assert octs == restate.to_tuples(fe.getOctants().as_array())


def test_restatement_matches_golden():
    g = np.load(os.path.join(GOLD, "box7_l2_p2_c1_o2.npz"))
    level, passes, pct, corner, order = [int(v) for v in g["params"]]
    # the restatement needs the tables: derive them with the emulated host class
    import ctypes
    from tmr_b200 import _capi

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    emu = _capi.bind(ctypes.CDLL(os.path.join(root, "tests", "emu", "_build", "libtmr_emu.so")))
    f = OctForest(lib=emu)
    f.setConnectivity(g["block_conn"])
    f.createTrees(level)
    T = tables_of(f)
    octs = restate.to_tuples(f.getOctants().as_array())
    for p in range(passes):
        rec = np.zeros(len(octs), dtype=_capi.OCT_DTYPE)
        for i, o in enumerate(octs):
            rec[i] = (o[0], o[1], o[2], o[3], i, o[4], 0)
        octs = restate.refine(octs, util.synth_flags(rec, 2024 + p, pct).tolist())
        octs = restate.balance(octs, T, corner)
    assert octs == restate.to_tuples(g["octants"])
