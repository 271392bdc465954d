#!/usr/bin/env python
"""Secondary measurement (not the driver's headline): the multigrid hierarchy of
BASELINE.json configs[2] ("C3") -- 8 trees, createTrees(5), hash-driven passes,
level 0 at order 3, level 1 = duplicate at order 2, deeper levels = coarsen +
balance(1) -- timing createNodes and createInterpolation per level pair in
rows/s, next to the unmodified reference on a bounded sample.

    python bench_interp.py [--passes P] [--level L] [--cpu-passes Q]
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def hierarchy(lib, level, passes, pct, sync, device_flags, device_interp=None, api=True):
    import util
    from tmr_b200.forest import OctForest

    f = OctForest(order=3, lib=lib)
    f.setConnectivity(util.structured_conn(2))
    f.createTrees(level)
    for p in range(passes):
        if device_flags is not None:
            device_flags(f, 2024 + p, pct)
        else:
            f.refine(util.synth_flags(f.getOctants().as_array(), 2024 + p, pct))
        f.balance(1)
    forests, times = [f], []
    order = 3
    for lev in range(4):
        prev = forests[-1]
        if order > 2:
            nxt = prev.duplicate()
            order -= 1
            nxt.setMeshOrder(order)
        else:
            nxt = prev.coarsen()
            nxt.balance(1)
        forests.append(nxt)
    out = []
    for k in range(len(forests) - 1):
        fine, coarse = forests[k], forests[k + 1]
        sync()
        t0 = time.perf_counter()
        fine.createNodes()
        sync()
        t1 = time.perf_counter()
        coarse.createNodes()
        sync()
        t2 = time.perf_counter()
        t_nodes_coarse = t2 - t1
        t_dev = None
        if device_interp is not None:
            # device-resident CSR only (no D2H, no per-row addInterp hand-off)
            device_interp(fine, coarse)
            sync()
            t_a = time.perf_counter()
            device_interp(fine, coarse)
            sync()
            t_dev = time.perf_counter() - t_a
            t2 = time.perf_counter()
        t_bulk = None
        if (api == "bulk" or api is True) and hasattr(lib, "tmr_b200_create_interpolation_csr"):
            # the whole CSR in one hand-off (tmr_b200_create_interpolation_csr)
            tb = time.perf_counter()
            rows, rowp, cols, vals = fine.createInterpolationCSR(coarse)
            sync()
            t_bulk = time.perf_counter() - tb
            t2 = time.perf_counter()
        if api is True:
            # the reference's own hand-off: one addInterp call per row
            vec = fine.createInterpolation(coarse)
            sync()
            t3 = time.perf_counter()
            rows, rowp, cols, vals = vec.get()
        elif api == "bulk" and t_bulk is not None:
            t3 = t2
        else:
            t3 = t2
            a, b = ctypes.c_int64(0), ctypes.c_int64(0)
            lib.tmrgpu_create_interp(ctypes.c_void_p(lib.tmr_b200_device_forest(fine._ptr)),
                                     ctypes.c_void_p(lib.tmr_b200_device_forest(coarse._ptr)),
                                     ctypes.byref(a), ctypes.byref(b))
            rows, cols = np.zeros(a.value, np.int8), np.zeros(b.value, np.int8)
            rowp, vals = np.zeros(1, np.int64), np.zeros(0)
        sums = np.add.reduceat(vals, rowp[:-1]) if (api and len(rows)) else np.zeros(0)
        if api and len(rows):
            rowp = np.asarray(rowp)
        out.append({
            "level": k, "fine_order": fine.getMeshOrder(), "coarse_order": coarse.getMeshOrder(),
            "fine_octants": fine.getNumOctants(), "coarse_octants": coarse.getNumOctants(),
            "rows": int(len(rows)), "nnz": int(len(cols)),
            "create_nodes_fine_s": t1 - t0, "create_nodes_coarse_s": t_nodes_coarse,
            "create_interp_s": (t3 - t2) if api is True else None,
            "rows_per_s": (len(rows) / max(t3 - t2, 1e-9)) if api is True else None,
            "device_csr_s": t_dev, "api_bulk_csr_s": t_bulk,
            "api_bulk_rows_per_s": (len(rows) / t_bulk) if t_bulk else None,
            "device_rows_per_s": (len(rows) / t_dev) if t_dev else None,
            "max_rowsum_err": float(np.abs(sums - 1).max()) if len(sums) else None,
        })
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=5)
    ap.add_argument("--passes", type=int, default=3)
    ap.add_argument("--pct", type=int, default=30)
    ap.add_argument("--cpu-level", type=int, default=3)
    ap.add_argument("--cpu-passes", type=int, default=2)
    ap.add_argument("--profile-out", default=None,
                    help="per-kernel CUDA-event times (ms) of the whole hierarchy build go here")
    ap.add_argument("--no-api", action="store_true",
                    help="skip the per-row addInterp hand-off timing (device CSR only)")
    ap.add_argument("--bulk-only", action="store_true",
                    help="time the one-call CSR hand-off (createInterpolationCSR) but not the "
                         "per-row addInterp loop")
    args = ap.parse_args()
    import torch

    import tmr_b200

    lib = tmr_b200.require_gpu()
    P = ctypes.c_void_p
    lib.tmr_b200_context.restype = P
    lib.tmr_b200_device_forest.restype = P
    lib.tmr_b200_device_forest.argtypes = [P]
    lib.tmrgpu_count.restype = ctypes.c_int64
    lib.tmrgpu_count.argtypes = [P]
    lib.tmrgpu_dev_alloc.argtypes = [P, ctypes.c_int64, ctypes.POINTER(P)]
    lib.tmrgpu_dev_free.argtypes = [P, P]
    lib.tmrgpu_synth_flags.argtypes = [P, ctypes.c_uint64, ctypes.c_int, P]
    lib.tmrgpu_refine_device.argtypes = [P, P, ctypes.c_int, ctypes.c_int]
    lib.tmrgpu_ctx_sync.argtypes = [P]
    ctx = P(lib.tmr_b200_context())

    def dev_flags(f, seed, pct):
        dev = P(lib.tmr_b200_device_forest(f._ptr))
        n = lib.tmrgpu_count(dev)
        buf = P()
        lib.tmrgpu_dev_alloc(ctx, 4 * max(n, 1), ctypes.byref(buf))
        lib.tmrgpu_synth_flags(dev, seed, pct, buf)
        lib.tmrgpu_refine_device(dev, buf, 0, 30)
        lib.tmrgpu_dev_free(ctx, buf)

    def sync():
        lib.tmrgpu_ctx_sync(ctx)

    lib.tmrgpu_create_interp.argtypes = [P, P, ctypes.POINTER(ctypes.c_int64),
                                         ctypes.POINTER(ctypes.c_int64)]

    def dev_interp(fine, coarse):
        a, b = ctypes.c_int64(0), ctypes.c_int64(0)
        lib.tmrgpu_create_interp(P(lib.tmr_b200_device_forest(fine._ptr)),
                                 P(lib.tmr_b200_device_forest(coarse._ptr)),
                                 ctypes.byref(a), ctypes.byref(b))

    hierarchy(lib, 2, 2, args.pct, sync, dev_flags)  # warm-up (allocator, kernels)
    if args.profile_out:
        lib.tmrgpu_profile_enable.argtypes = [P, ctypes.c_int]
        lib.tmrgpu_profile_reset.argtypes = [P]
        lib.tmrgpu_profile_json.argtypes = [P, ctypes.c_char_p, ctypes.c_int]
        lib.tmrgpu_profile_reset(ctx)
        lib.tmrgpu_profile_enable(ctx, 1)
    gpu = hierarchy(lib, args.level, args.passes, args.pct, sync, dev_flags, dev_interp,
                    api=("bulk" if args.bulk_only else (not args.no_api)))
    if args.profile_out:
        buf = ctypes.create_string_buffer(1 << 16)
        lib.tmrgpu_profile_json(ctx, buf, len(buf))
        lib.tmrgpu_profile_enable(ctx, 0)
        prof = json.loads(buf.value.decode())
        with open(args.profile_out, "w") as fh:
            json.dump(dict(sorted(prof.items(), key=lambda kv: -kv[1]["ms"])), fh, indent=1)
    result = {"workload": "C3 hierarchy: 8 trees, createTrees(%d), %d passes pct %d, balance(1)"
                          % (args.level, args.passes, args.pct), "b200": gpu}
    from oracle import ref_loader

    if ref_loader.available():
        ref = ref_loader.load()
        cpu = hierarchy(ref, args.cpu_level, args.cpu_passes, args.pct, lambda: None, None)
        result["reference_cpu_1_core"] = {
            "sample": "same hierarchy at createTrees(%d), %d passes" % (args.cpu_level, args.cpu_passes),
            "levels": cpu}
    print(json.dumps(result, indent=1))


if __name__ == "__main__":
    main()
