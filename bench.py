#!/usr/bin/env python
"""bench.py -- octants/s of one adaptation cycle refine -> balance ->
createNodes (BASELINE.json metric) on synthetic random-refinement forests.

Workload at N=1 (BASELINE.json configs[1], recipe SURVEY.md 8(d) "C2"): 8x8x8
trees, createTrees(3), four hash-driven passes at pct=35; the timed step is the
LAST cycle (refine(flags) + balance(0) + createNodes(order 2)) which ends at
86,278,900 octants.  Every step restarts from a device copy of the pre-step
forest (the copy is inside the timed region; it is < 0.1% of a step).

  value     whole-job octants/s with the flags already in HBM
  e2e       the same cycle through the reference-facing TMROctForest API with
            HOST flags in, and conn / node numbers / dependent CSR read back
  roofline  dominant kernel, algorithmic bytes / CUDA-event time, vs the
            measured HBM copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the unmodified reference C++ (oracle/_ref)
            on the box's host cores, threads-as-ranks, on a bounded sample.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "octants/sec for refine+balance+createNodes"
UNIT = "octants/s"
FULL = dict(nb=8, level=3, passes=4, pct=35, order=2, corner=0, seed=2024)
# bounded CPU sample of the same recipe (two passes instead of four)
CPU_SAMPLE = dict(nb=8, level=3, passes=2, pct=35, order=2, corner=0, seed=2024)
# BASELINE.json configs[0] (SURVEY.md 8(d) "C1"): one tree, createTrees(4), four
# passes pct=30 -> 1,027,916 octants.  Small enough for the reference to run in
# FULL, so both arms time it: the one same-input GPU/CPU comparison.
C1 = dict(nb=1, level=4, passes=4, pct=30, order=2, corner=0, seed=2024)
# fingerprints of the reference on C2 / C1 (pinned by the oracle)
C2_PIN = dict(octants=86278900, checksum="da1d7223d950ef5c", owned_nodes=53774081)
# the ~1e9-octant forest of BASELINE configs[4] (8x8x96 trees): fingerprint of the
# same forest built on 4 GPUs (profiles/bench_r02_1B_n4.json)
NORTH_STAR_PIN = dict(octants=1037532543, checksum="b263cb130799478f", owned_nodes=644833973)
C1_PIN = dict(octants=1027916, checksum="55e9487c98c7a2ff", owned_nodes=652025,
              dep_nodes=908576, dep_nnz=2441728)


def workload_name(cfg):
    return ("%dx%dx%d-tree box, createTrees(%d), %d passes pct=%d, last cycle "
            "refine+balance(%d)+createNodes(order %d)" %
            (cfg["nb"], cfg["nb"], cfg["nb"], cfg["level"], cfg["passes"],
             cfg["pct"], cfg["corner"], cfg["order"]))


# --------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [t.strip() for t in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------
# reference arm (CPU): the unmodified reference through oracle/_ref
# --------------------------------------------------------------------------
def run_reference_cycle(cfg, nranks, steps, warmup):
    """Time the last adaptation cycle of `cfg` on `nranks` thread-ranks of the
    oracle.  Returns (octants/s, final octants, per-step seconds)."""
    import util
    from oracle import ref_loader
    from tmr_b200.forest import OctForest

    lib = ref_loader.load()
    conn = util.structured_conn(cfg["nb"])
    total = steps + warmup
    times = [[0.0] * total for _ in range(nranks)]
    finals = [0] * nranks
    lib.shim_world_begin(nranks)
    barrier = threading.Barrier(nranks)

    def body(rank):
        lib.shim_attach(rank)
        f = OctForest(order=cfg["order"], lib=lib)
        f.setConnectivity(conn)
        f.createTrees(cfg["level"])
        f.repartition()
        for p in range(cfg["passes"] - 1):
            o = f.getOctants().as_array()
            f.refine(util.synth_flags(o, cfg["seed"] + p, cfg["pct"]))
            f.balance(cfg["corner"])
            f.repartition()
        o = f.getOctants().as_array()
        flags = util.synth_flags(o, cfg["seed"] + cfg["passes"] - 1, cfg["pct"])
        for s in range(total):
            work = f.duplicate()
            barrier.wait()
            t0 = time.perf_counter()
            work.refine(flags)
            work.balance(cfg["corner"])
            work.createNodes()
            barrier.wait()
            times[rank][s] = time.perf_counter() - t0
            finals[rank] = work.getNumOctants()
            del work

    threads = [threading.Thread(target=body, args=(r,)) for r in range(nranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    lib.shim_world_end()
    per_step = [max(times[r][s] for r in range(nranks)) for s in range(warmup, total)]
    n_final = sum(finals)
    return n_final * len(per_step) / sum(per_step), n_final, per_step


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_loader

    if not ref_loader.available():
        emit({"impl": "reference", "unavailable":
              "oracle/_ref/libtmr_ref.so missing (built by __graft_entry__.build() where /root/reference exists)"})
        return
    cores = min(os.cpu_count() or 1, 8)
    cfg = CPU_SAMPLE
    value, n_final, per_step = run_reference_cycle(cfg, cores, args.steps, args.warmup)
    sample = ("bounded sample: %s -> %d octants, %d thread-ranks (MPI shim), each step = that whole cycle"
              % (workload_name(cfg), n_final, cores))
    # the same-input comparison: configs[0] (C1) in full, at `cores` ranks and at 1 rank
    c1_v, c1_n, c1_t = run_reference_cycle(C1, cores, 2, 1)
    c1_v1, _, c1_t1 = run_reference_cycle(C1, 1, 1, 0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(per_step)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
        "config": {"workload": workload_name(FULL), "timed_sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores,
                         "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "same_config": {"workload": workload_name(C1) + " (BASELINE configs[0], in full)",
                        "octants": int(c1_n), "value": c1_v, "unit": UNIT, "cores": cores,
                        "ms_per_step": 1e3 * float(np.mean(c1_t)),
                        "value_1_rank": c1_v1, "ms_per_step_1_rank": 1e3 * c1_t1[0]},
    }
    emit(line)


# --------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------
def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def emit(line):
    """Print the ONE JSON line on the real stdout (everything else -- NCCL's
    version banner, library chatter -- was redirected to stderr)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


# keep stdout clean for the driver: route fd 1 to stderr until the final line
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--passes", type=int, default=FULL["passes"],
                    help="refinement passes of the recipe (4 = the named ~86M config)")
    ap.add_argument("--pct", type=int, default=None,
                    help="refinement percentage of the recipe (default: 35 for c2, 30 for c4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-north-star", action="store_true",
                    help="N=8: skip the additional 1e9-octant measurement (BASELINE configs[4])")
    ap.add_argument("--no-parity", action="store_true",
                    help="skip the parity gate that runs before the timed region")
    ap.add_argument("--profiler-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed region (for ncu --profile-from-start off)")
    ap.add_argument("--profile-out", default=None,
                    help="write the per-kernel CUDA-event times of the timed region (ms per step) here")
    ap.add_argument("--workload", default="c2", choices=["c2", "c4"],
                    help="c2: 8x8x8N-tree box (BASELINE configs[1], the headline); "
                         "c4: 5x5x6 tiled 7-tree butterfly cells = 1050 trees with all 8 "
                         "face orientations, balance(1) (BASELINE configs[3])")
    ap.add_argument("--nbz-per-gpu", type=int, default=8,
                    help="c2: trees along z per GPU (8 = 86 M octants per GPU; 12 = 130 M, "
                         "i.e. >1e9 octants on 8 GPUs, BASELINE configs[4])")
    ap.add_argument("--strong", action="store_true",
                    help="N>1: split the SAME 8x8x8-tree forest over the GPUs "
                         "(default: weak scaling, 8x8x8N trees)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_main(args)
        return

    import torch
    import torch.distributed as dist

    import util
    import tmr_b200
    from tmr_b200.forest import OctForest

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    from tmr_b200 import dist as tdist

    numa_cpus = tdist.bind_to_gpu_numa(local, world)  # page-locked mirrors on the GPU's NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ["TMR_B200_DEVICE"] = str(local)
    stream = torch.cuda.current_stream()
    tmr_b200.use_stream(stream.cuda_stream)
    lib = tmr_b200.require_gpu()
    if world > 1:
        tdist.init_from_torch(lib)

    P, I64 = ctypes.c_void_p, ctypes.c_int64
    lib.tmr_b200_context.restype = P
    lib.tmr_b200_device_forest.restype = P
    lib.tmr_b200_device_forest.argtypes = [P]
    for name, argt in [
        ("tmrgpu_synth_flags", [P, ctypes.c_uint64, ctypes.c_int, P]),
        ("tmrgpu_refine_device", [P, P, ctypes.c_int, ctypes.c_int]),
        ("tmrgpu_balance", [P, ctypes.c_int]),
        ("tmrgpu_create_nodes", [P, ctypes.c_int, ctypes.c_int, P]),
        ("tmrgpu_duplicate", [P, P]),
        ("tmrgpu_dev_alloc", [P, I64, ctypes.POINTER(P)]),
        ("tmrgpu_dev_free", [P, P]),
        ("tmrgpu_checksum", [P, ctypes.POINTER(ctypes.c_uint64)]),
        ("tmrgpu_node_sizes", [P, ctypes.POINTER(I64)]),
        ("tmrgpu_profile_enable", [P, ctypes.c_int]),
        ("tmrgpu_profile_reset", [P]),
        ("tmrgpu_profile_json", [P, ctypes.c_char_p, ctypes.c_int]),
        ("tmrgpu_count", [P]),
        ("tmrgpu_repartition", [P, ctypes.c_int]),
        ("tmrgpu_copy_d2h", [P, P, P, I64]),
        ("tmrgpu_copy_h2d", [P, P, P, I64]),
        ("tmrgpu_set_node_prefetch", [P, ctypes.c_int]),
    ]:
        getattr(lib, name).argtypes = argt
    lib.tmrgpu_count.restype = I64
    lib.tmrgpu_launch_count.restype = ctypes.c_long
    lib.tmrgpu_launch_count.argtypes = [P]
    ctx = P(lib.tmr_b200_context())

    # ---- parity gate (before anything is timed) -----------------------------------
    # N>1: every multi-rank case of tests/multi_gpu_check.py, rank by rank against
    # the oracle at the same rank count; then the C2 forest split over the N GPUs
    # must reproduce the reference's fingerprint.  A mismatch ends the run.
    parity = {"cases": 0, "ok": True}
    if world > 1 and not args.no_parity:
        import multi_gpu_check

        ncases, nfail, nunchecked = multi_gpu_check.check_cases(lib, rank, world, verbose=True)
        parity = {"cases": ncases, "ok": nfail == 0, "failures": nfail,
                  "without_oracle": nunchecked,
                  "what": "octants per stage, conn, node numbers, node_range, dependent CSR, "
                          "prolongation rows of every rank vs the same rank of the reference "
                          "at %d ranks" % world}
        if nfail:
            if rank == 0:
                emit({"metric": METRIC, "value": None, "parity": parity,
                      "error": "multi-GPU parity mismatch"})
            dist.destroy_process_group()
            sys.exit(1)

    cfg = dict(FULL)
    cfg["passes"] = args.passes
    if args.workload == "c4":
        cfg["corner"] = 1
        cfg["pct"] = 30
    if args.pct is not None:
        cfg["pct"] = args.pct
    knots = (ctypes.c_double * 2)(-1.0, 1.0)

    # ---- build-up: everything before the timed cycle, all on the device ----
    # N>1: ONE forest partitioned along the Morton curve over the N GPUs
    # (NCCL exchanges inside refine/balance/repartition/createNodes).  Weak
    # scaling stacks N copies of the 8x8x8-tree box along z (8x8x8N trees);
    # --strong splits the same 8x8x8 box.
    nbz = args.nbz_per_gpu * (1 if (args.strong or world == 1) else world)
    if args.workload == "c4":
        block_conn = util.butterfly_conn(5, 5, 6 * (1 if (args.strong or world == 1) else world))
    else:
        block_conn = util.structured_conn(cfg["nb"], cfg["nb"], nbz)

    def synth(dev, seed, pct):
        n = lib.tmrgpu_count(dev)
        buf = P()
        lib.tmrgpu_dev_alloc(ctx, 4 * max(n, 1), ctypes.byref(buf))
        lib.tmrgpu_synth_flags(dev, seed, pct, buf)
        return buf, n

    def build_base(conn, c):
        """everything before the timed cycle, all on the device"""
        b = OctForest(order=c["order"], lib=lib)
        b.setConnectivity(conn)
        b.createTrees(c["level"])
        dev = P(lib.tmr_b200_device_forest(b._ptr))
        if world > 1:
            assert lib.tmrgpu_repartition(dev, -1) == 0
        for p in range(c["passes"] - 1):
            buf, _ = synth(dev, c["seed"] + p, c["pct"])
            assert lib.tmrgpu_refine_device(dev, buf, 0, 30) == 0
            assert lib.tmrgpu_balance(dev, c["corner"]) == 0
            if world > 1:
                assert lib.tmrgpu_repartition(dev, -1) == 0
            lib.tmrgpu_dev_free(ctx, buf)
        flags, n_in = synth(dev, c["seed"] + c["passes"] - 1, c["pct"])
        return b, dev, flags, n_in

    def cycle(b, flags, c):
        work = b.duplicate()
        wdev = P(lib.tmr_b200_device_forest(work._ptr))
        assert lib.tmrgpu_refine_device(wdev, flags, 0, 30) == 0
        assert lib.tmrgpu_balance(wdev, c["corner"]) == 0
        if world > 1:
            assert lib.tmrgpu_repartition(wdev, -1) == 0
        assert lib.tmrgpu_create_nodes(wdev, c["order"], 1, knots) == 0
        return work, wdev

    def fingerprint(wdev):
        """global (octants, checksum, owned nodes, dependent nodes, stencil entries)"""
        sz = (I64 * 6)()
        lib.tmrgpu_node_sizes(wdev, sz)
        cs_ = ctypes.c_uint64(0)
        lib.tmrgpu_checksum(wdev, ctypes.byref(cs_))
        v = torch.tensor([lib.tmrgpu_count(wdev), cs_.value & 0xFFFFFFFF, cs_.value >> 32,
                          sz[3], sz[2], sz[4]], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
        v = [int(x) for x in v.tolist()]
        return {"octants": v[0],
                "checksum": "%016x" % ((v[1] + (v[2] << 32)) & 0xFFFFFFFFFFFFFFFF),
                "owned_nodes": v[3], "dep_nodes": v[4], "dep_nnz": v[5]}

    def timed_cycles(b, flags, c, steps, warmup):
        for _ in range(warmup):
            w, _ = cycle(b, flags, c)
            del w
        barrier()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        w = None
        for _ in range(steps):
            w = None
            w = cycle(b, flags, c)
        z.record(stream)
        barrier()
        fp = fingerprint(w[1])
        return a.elapsed_time(z) / steps, fp

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N>1: the SAME C2 forest split over the N GPUs must give the reference's result
    if world > 1 and not args.no_parity and not args.strong:
        sb, sdev, sflags, _ = build_base(util.structured_conn(8), FULL)
        sw, swdev = cycle(sb, sflags, FULL)
        fp = fingerprint(swdev)
        ok = all(fp[k] == C2_PIN[k] for k in C2_PIN)
        parity["c2_split_over_%d_gpus" % world] = dict(fp, matches_reference=ok)
        parity["ok"] = parity["ok"] and ok
        del sw, sb
        lib.tmrgpu_dev_free(ctx, sflags)
        if not ok:
            if rank == 0:
                emit({"metric": METRIC, "value": None, "parity": parity,
                      "error": "C2 fingerprint mismatch on %d GPUs" % world})
            dist.destroy_process_group()
            sys.exit(1)

    base, bdev, d_flags, e_in = build_base(block_conn, cfg)
    # host copy of the flags (pinned): the e2e arm uploads it every step
    h_flags_t = torch.empty(max(e_in, 1), dtype=torch.int32).pin_memory()
    h_flags = h_flags_t.numpy()[:e_in]
    lib.tmrgpu_copy_d2h(ctx, h_flags.ctypes.data, d_flags, 4 * e_in)

    def step_device():
        """device-resident cycle: flags already in HBM"""
        return cycle(base, d_flags, cfg)

    def step_e2e(marks=None):
        """reference-facing API: host flags in; conn, node numbers and the
        dependent-node CSR out (everything createTACS reads, reference
        src/TMR_TACSCreator.cpp:332-461)"""
        def mark(name):
            if marks is not None:
                marks.append((name, time.perf_counter()))
        mark("start")
        work = base.duplicate()
        lib.tmrc_refine(work._ptr, h_flags.ctypes.data, 0, 30)  # H2D inside
        lib.tmrc_balance(work._ptr, cfg["corner"])
        if world > 1:
            lib.tmrc_repartition(work._ptr, -1)
        mark("refine+balance (host returns)")
        lib.tmrc_create_nodes(work._ptr)
        mark("createNodes (host returns)")
        cptr = ctypes.POINTER(ctypes.c_int)()
        ne, no = ctypes.c_int(0), ctypes.c_int(0)
        lib.tmrc_get_node_conn(work._ptr, ctypes.byref(cptr), ctypes.byref(ne),
                               ctypes.byref(no))
        mark("getNodeConn")
        p1, p2 = ctypes.POINTER(ctypes.c_int)(), ctypes.POINTER(ctypes.c_int)()
        p3 = ctypes.POINTER(ctypes.c_double)()
        nd_ = lib.tmrc_get_dep_node_conn(work._ptr, ctypes.byref(p1), ctypes.byref(p2),
                                         ctypes.byref(p3))
        mark("getDepNodeConn")
        p4 = ctypes.POINTER(ctypes.c_int)()
        nn_ = lib.tmrc_get_node_numbers(work._ptr, ctypes.byref(p4))
        mark("getNodeNumbers")
        # touch the last word of every array: the copies have landed
        probe = 0
        if ne.value:
            probe += cptr[ne.value * 8 - 1] + p4[nn_ - 1] + p1[nd_]
        return work, ne.value, probe

    # ---- device-resident arm -------------------------------------------------
    lib.tmrgpu_profile_enable(ctx, 0)
    for _ in range(args.warmup):
        w, wdev = step_device()
        del w
    barrier()
    lib.tmrgpu_profile_reset(ctx)
    lib.tmrgpu_profile_enable(ctx, 1)
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if args.profiler_range:
        torch.cuda.profiler.start()
    ev0.record(stream)
    last = None
    for _ in range(args.steps):
        last = None  # free the previous step's forest
        last = step_device()
    ev1.record(stream)
    barrier()
    if args.profiler_range:
        torch.cuda.profiler.stop()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = lib.tmrgpu_launch_count(ctx)
    lib.tmrgpu_sync_count.restype = ctypes.c_long
    lib.tmrgpu_sync_count.argtypes = [P]
    host_syncs = lib.tmrgpu_sync_count(ctx)
    buf = ctypes.create_string_buffer(1 << 16)
    lib.tmrgpu_profile_json(ctx, buf, len(buf))
    prof = json.loads(buf.value.decode())
    lib.tmrgpu_profile_enable(ctx, 0)
    if args.profile_out and rank == 0:
        with open(args.profile_out, "w") as fh:
            json.dump({k: {"launches_per_step": v["launches"] / args.steps,
                           "ms_per_step": v["ms"] / args.steps}
                       for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}, fh, indent=1)
    work, wdev = last
    fp_final = fingerprint(wdev)
    pinned = (args.workload == "c2" and cfg["passes"] == FULL["passes"] and
              cfg["pct"] == FULL["pct"] and
              (world == 1 or args.strong) and args.nbz_per_gpu == 8)
    if pinned:
        ok = all(fp_final[k] == C2_PIN[k] for k in C2_PIN)
        parity["timed_result_matches_reference_c2"] = ok
        parity["ok"] = parity["ok"] and ok
    e_final = lib.tmrgpu_count(wdev)
    sizes = (I64 * 6)()
    lib.tmrgpu_node_sizes(wdev, sizes)
    lib.tmrgpu_node_candidates.restype = I64
    lib.tmrgpu_node_candidates.argtypes = [P]
    n_cand = int(lib.tmrgpu_node_candidates(wdev))
    csum = ctypes.c_uint64(0)
    lib.tmrgpu_checksum(wdev, ctypes.byref(csum))
    del work, last

    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(e_final), float(e_in)], dtype=torch.float64, device="cuda")
    # 64-bit wrap-around checksum: reduce as two 32-bit halves
    cs = torch.tensor([csum.value & 0xFFFFFFFF, csum.value >> 32], dtype=torch.int64,
                      device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(cs, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    total_octants = float(tot[0].item())
    total_in = float(tot[1].item())
    lo, hi = int(cs[0].item()), int(cs[1].item())
    global_checksum = (lo + (hi << 32)) & 0xFFFFFFFFFFFFFFFF
    value = total_octants * args.steps / (ms * 1e-3)

    # per phase of one cycle: synchronised wall time, kernel launches, blocking
    # host round trips (rank 0's view; the phases are collective)
    phases = {}
    if True:
        lib.tmrgpu_profile_reset(ctx)

        def phase(name, fn):
            barrier()
            l0, s0 = lib.tmrgpu_launch_count(ctx), lib.tmrgpu_sync_count(ctx)
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            phases[name] = {"ms": round(1e3 * (time.perf_counter() - t0), 3),
                            "launches": int(lib.tmrgpu_launch_count(ctx) - l0),
                            "host_round_trips": int(lib.tmrgpu_sync_count(ctx) - s0)}

        pw = base.duplicate()
        pdev = P(lib.tmr_b200_device_forest(pw._ptr))
        phase("refine", lambda: lib.tmrgpu_refine_device(pdev, d_flags, 0, 30))
        phase("balance", lambda: lib.tmrgpu_balance(pdev, cfg["corner"]))
        if world > 1:
            phase("repartition", lambda: lib.tmrgpu_repartition(pdev, -1))
        phase("createNodes", lambda: lib.tmrgpu_create_nodes(pdev, cfg["order"], 1, knots))
        del pw

    # the same cycle with the H2D of the flags inside the timed region
    def step_device_h2d():
        lib.tmrgpu_copy_h2d(ctx, d_flags, h_flags.ctypes.data, 4 * e_in)
        return cycle(base, d_flags, cfg)

    w = step_device_h2d()
    del w
    barrier()
    a_, z_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a_.record(stream)
    for _ in range(args.steps):
        w = None
        w = step_device_h2d()
    z_.record(stream)
    barrier()
    del w
    th = torch.tensor([a_.elapsed_time(z_)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(th, op=dist.ReduceOp.MAX)
    value_h2d = total_octants * args.steps / (float(th.item()) * 1e-3)

    # BASELINE configs[0] (C1) in full on the GPU: the reference arm times the same
    # input in full, so this pair is a same-config comparison
    same_config = None
    if world == 1 and not args.no_parity:
        b1, d1, f1, _ = build_base(util.structured_conn(1), C1)
        c1_ms, c1_fp = timed_cycles(b1, f1, C1, max(args.steps, 5), 2)
        c1_ok = all(c1_fp[k] == C1_PIN[k] for k in C1_PIN)
        parity["c1_matches_reference"] = c1_ok
        parity["ok"] = parity["ok"] and c1_ok
        same_config = {"workload": workload_name(C1) + " (BASELINE configs[0], in full)",
                       "octants": c1_fp["octants"], "ms_per_step": c1_ms,
                       "value": c1_fp["octants"] / (c1_ms * 1e-3), "unit": UNIT,
                       "fingerprint": c1_fp}
        del b1
        lib.tmrgpu_dev_free(ctx, f1)

    # BASELINE configs[4] / north_star: the ~1e9-octant forest (8x8x96 trees, 130 M
    # octants per GPU) on 8 GPUs, measured inside the same run so that the driver's
    # N=8 line carries it; its fingerprint must equal the one recorded from the
    # same forest on 4 GPUs (profiles/scaling_r02.md)
    north_star = None
    ns_world = int(os.environ.get("TMR_B200_NORTH_STAR_WORLD", "8"))  # (test hook)
    if (world == ns_world and world > 1 and args.workload == "c2" and not args.strong
            and args.nbz_per_gpu == 8 and cfg["passes"] == FULL["passes"]
            and cfg["pct"] == FULL["pct"] and not args.no_north_star):
        nsb, _, nsf, _ = build_base(util.structured_conn(8, 8, 12 * world), cfg)
        ns_ms, ns_fp = timed_cycles(nsb, nsf, cfg, 3, 2)
        tt = torch.tensor([ns_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ns_ms = float(tt.item())
        ns_ok = all(ns_fp[k] == NORTH_STAR_PIN[k] for k in NORTH_STAR_PIN)
        if world == 8:
            parity["north_star_matches_4_gpu_run"] = ns_ok
            parity["ok"] = parity["ok"] and ns_ok
        north_star = {"workload": "8x8x%d-tree box (BASELINE configs[4]), same recipe, one forest over %d GPUs"
                                  % (12 * world, world),
                      "octants": ns_fp["octants"], "ms_per_step": ns_ms,
                      "value": ns_fp["octants"] / (ns_ms * 1e-3), "unit": UNIT,
                      "fingerprint": ns_fp, "fingerprint_on_4_gpus": NORTH_STAR_PIN,
                      "frac_of_aggregate_hbm_peak_by_compulsory_bytes":
                          134.0 * ns_fp["octants"] / (ns_ms * 1e-3) / 1e9 / (world * load_peaks()[0])}
        del nsb
        lib.tmrgpu_dev_free(ctx, nsf)

    # ---- end-to-end arm --------------------------------------------------------
    # node arrays are copied to page-locked host memory on a second stream as
    # soon as each is final (conn after the renumbering, the dependent CSR at
    # the end): the read-back overlaps the rest of createNodes
    lib.tmrgpu_set_node_prefetch(bdev, 7)
    for _ in range(max(1, args.warmup - 1)):
        w = step_e2e()
        del w
    barrier()
    lib.tmrgpu_bus_bytes.restype = I64
    lib.tmrgpu_bus_bytes.argtypes = [P, ctypes.c_int]
    lib.tmrgpu_profile_reset(ctx)  # zero the bus byte counters
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        w = step_e2e()
        del w
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    bus_d2h = int(lib.tmrgpu_bus_bytes(ctx, 0)) // args.steps
    bus_h2d = int(lib.tmrgpu_bus_bytes(ctx, 1)) // args.steps
    e2e_ms = max(e0.elapsed_time(e1), wall * 1e3)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_octants * args.steps / (float(t.item()) * 1e-3)
    # where one end-to-end step spends its wall time (host clock, rank 0)
    marks = []
    w = step_e2e(marks)
    del w
    barrier()
    e2e_breakdown = {marks[k][0]: round(1e3 * (marks[k][1] - marks[k - 1][1]), 2)
                     for k in range(1, len(marks))}
    lib.tmrgpu_set_node_prefetch(bdev, 0)
    npe = cfg["order"] ** 3
    # bytes the host arrays handed out hold (conn, sorted numbers, dependent CSR);
    # what crossed the bus is counted by the library (dep_ptr / dep_weights travel
    # as 2-byte stencil codes, one rank's sorted numbers are a range)
    d2h_arrays = 4 * (sizes[0] * npe + sizes[1] + sizes[2] + 1 + sizes[4]) + 8 * sizes[4]

    # ---- roofline of the dominant kernel -------------------------------------------
    peak, peak_src = load_peaks()
    dom = max(prof.items(), key=lambda kv: kv[1]["ms"]) if prof else (None, None)
    roof = None
    kernel_share = {k: round(v["ms"] / max(ms, 1e-9), 4) for k, v in
                    sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]}
    if dom[0]:
        name, st = dom
        # algorithmic bytes per launch of the dominant kernel
        n_pairs = sizes[0] * npe
        n_slot_nodes = sizes[1]
        alg = {"radix_pass_pairs[nodes]": 2 * 12 * n_cand,
               "radix_pass_keys[nodes]": 2 * 8 * n_cand,
               "radix_pass_keys[leaves]": 2 * 8 * e_final,
               "radix_hist[nodes]": 8 * n_cand,
               "nodes_candidates": 8 * sizes[0] + 8 * n_cand,
               "nodes_unique_scatter_conn": 8 * n_cand + 4 * n_pairs + 8 * sizes[1],
               # keys in; (leaf, slot) of every corner out; slot masks read+written;
               # one 16-byte rank entry per 64 finest cells of the forest
               "nodes_slot_locate": (8 + 4 * npe + npe + 8) * sizes[0]
                                    + 16 * (len(block_conn) << (3 * 7)) // 64,
               "nodes_slot_resolve": (2 * 4 * npe + npe + 8) * sizes[0],
               "nodes_slot_keys": 16 * sizes[0] + 8 * n_slot_nodes,
               "nodes_conn_remap": 2 * 4 * n_pairs + 4 * sizes[1],
               "nodes_dep_winner": (8 + 2 + 4 * npe) * sizes[0] + 16 * sizes[2],
               "nodes_dep_fill": 12 * sizes[4] + 16 * sizes[2],
               "nodes_hanging_info": 10 * sizes[0]}.get(name)
        avg_ms = st["ms"] / st["launches"]
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")) as fh:
                ent = json.load(fh)["kernels"].get(name)
            # only valid for the launch shape it was captured on (the C2 cycle)
            if ent and pinned:
                traffic = ent["dram_bytes_per_launch"]
        except Exception:
            traffic = None
        if alg:
            ach = alg / (avg_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak,
                    "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "peak_source": peak_src, "avg_launch_ms": avg_ms,
                    "launches_per_step": st["launches"] / args.steps,
                    "algorithmic_bytes_per_launch": alg}
        else:
            roof = {"bound": "hbm", "kernel": name, "achieved": None, "peak": peak,
                    "unit": "GB/s", "frac": None, "traffic": None,
                    "avg_launch_ms": avg_ms}

    # whole-cycle compulsory traffic (SURVEY.md 8(d) B_alg) as a fraction of HBM peak
    b_alg = (28 * e_in + 24 * e_in + 24 * e_in + 24 * e_final + 24 * e_final
             + 4 * npe * e_final + 4 * sizes[1] + 4 * (sizes[2] + 1) + 12 * sizes[4])
    cycle_frac = b_alg * args.steps / (ms * 1e-3) / 1e9 / peak

    # whole cycle, measured: DRAM bytes of every kernel of one C2 cycle from the
    # committed ncu capture (profiles/ncu_traffic_r02.json, tools/ncu_cycle.py),
    # weighted with the kernel times measured live above
    cycle_dram = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")) as fh:
            cap = json.load(fh)
        if pinned:
            tot_b = sum(k["dram_bytes_per_step"] for k in cap["kernels"].values())
            kern_ms = sum(v["ms"] for v in prof.values()) / args.steps
            cycle_dram = {
                "measured_dram_bytes_per_cycle": int(tot_b),
                "amplification_vs_compulsory": tot_b / b_alg,
                "dram_gbs_over_kernel_time": tot_b / (kern_ms * 1e-3) / 1e9,
                "frac_of_hbm_peak_over_kernel_time": tot_b / (kern_ms * 1e-3) / 1e9 / peak,
                "frac_of_hbm_peak_over_step_time": tot_b / (ms / args.steps * 1e-3) / 1e9 / peak,
                "source": "ncu dram__bytes_read+write of every kernel of one cycle (%s), "
                          "times measured live" % cap.get("captured", "profiles/")}
    except Exception:
        cycle_dram = None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import ref_loader

        if ref_loader.available():
            cores = min(os.cpu_count() or 1, 8)
            v, nfin, per = run_reference_cycle(CPU_SAMPLE, cores, 1, 0)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": "%s -> %d octants, 1 cycle in %.1f s on %d thread-ranks of the unmodified reference"
                             % (workload_name(CPU_SAMPLE), nfin, per[0], cores)}
        else:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                   "sample": "oracle/_ref not built on this box"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if (args.strong and world > 1) else "weak",
            "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(cfg) if args.workload == "c2" else
                       "%d-tree tiled butterfly forest (all 8 face orientations, irregular valence), createTrees(%d), %d passes pct=%d, last cycle refine+balance(1)+createNodes(order 2)"
                       % (len(block_conn), cfg["level"], cfg["passes"], cfg["pct"]),
                       "arithmetic": "u64 Morton keys, int32 connectivity, f64 stencil weights",
                       "trees": int(len(block_conn)),
                       "octants_in": int(e_in), "octants_out_per_gpu": int(e_final),
                       "local_nodes": int(sizes[1]), "dep_nodes": int(sizes[2]),
                       "dep_nnz": int(sizes[4]), "node_candidates_sorted": n_cand, "checksum": "%016x" % global_checksum,
                       "octants_out_total": int(total_octants), "octants_in_total": int(total_in),
                       "parallelism": "1 GPU" if world == 1 else
                       "%d GPUs, one forest SFC-partitioned, NCCL all-to-all-v (%s: %dx%dx%d trees); cycle includes repartition()"
                       % (world, "strong" if args.strong else "weak", cfg["nb"], cfg["nb"], nbz),
                       "l2": "inputs larger than L2 (%.0f MB of keys per pass)" % (8e-6 * e_final)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bus_h2d,
                    "d2h_bytes_per_step": bus_d2h, "host_array_bytes_per_step": int(d2h_arrays),
                    "ms_per_step": float(t.item()) / args.steps,
                    "host_ms_per_call": e2e_breakdown,
                    "what": "TMROctForest API: refine(host flags) + balance + createNodes + "
                            "getNodeConn + getDepNodeConn + getNodeNumbers into host arrays; conn and "
                            "dep_conn are copied on a second stream as soon as they are final, "
                            "dep_ptr/dep_weights cross the bus as 2-byte stencil codes and are rebuilt "
                            "by host threads, the sorted node numbers are ranges written on the host "
                            "(d2h_bytes_per_step = what crossed the bus, host_array_bytes_per_step = "
                            "what the caller holds afterwards)"},
            "value_with_flags_h2d": value_h2d,
            "parity": parity,
            "same_config": same_config,
            "north_star_1e9_octants": north_star,
            "fingerprint": fp_final,
            "numa_bound_cpus": numa_cpus,
            "gpu_launches": int(launches),
            "kernel_launches_per_step": launches / args.steps,
            "blocking_host_round_trips_per_step": host_syncs / args.steps,
            "kernel_ms_per_step": sum(v["ms"] for v in prof.values()) / args.steps,
            "phases_of_one_cycle": phases,
            "clocks": clocks,
            "roofline": roof,
            "cycle_compulsory_bytes": int(b_alg),
            "cycle_frac_of_hbm_peak": cycle_frac,
            "cycle_dram_measured": cycle_dram,
            "kernel_share_of_step": kernel_share,
            "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    if not parity["ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
