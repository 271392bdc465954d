#!/usr/bin/env python3
"""Turn the ncu launch list of ONE C2 cycle into profiles/ncu_traffic_r02.json
and a markdown table.

  TMR_B200_LAUNCH_LOG=gpurun_out/launchlog.txt ncu --metrics \
      gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none --profile-from-start off --csv --log-file gpurun_out/l.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --profiler-range
  python tools/ncu_cycle.py gpurun_out/l.csv gpurun_out/launchlog.txt profiles/ncu_traffic_r02.json profiles/launches_r02.md

The launch log (written by prof_begin, prim_cuda.cu) names the profiled bracket
every kernel launch belongs to, so the DRAM bytes ncu measured per launch can be
summed under the same names bench.py times with CUDA events."""
import collections
import csv
import json
import re
import sys


def main():
    csv_path, log_path, out_json, out_md = sys.argv[1:5]
    rows = list(csv.DictReader(l for l in open(csv_path) if l.startswith('"')))
    launches = collections.OrderedDict()
    for r in rows:
        d = launches.setdefault(int(r["ID"]), {"fn": r["Kernel Name"], "grid": r["Grid Size"],
                                               "block": r["Block Size"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    # bracket names: "<launch index at bracket start> <name>"; the log holds the
    # brackets of the timed cycle only (profiling is enabled there only)
    starts = []
    for ln in open(log_path):
        p = ln.split(None, 1)
        if len(p) == 2:
            starts.append((int(p[0]), p[1].strip()))
    ids = sorted(launches)
    name_of = {}
    for k, (idx, name) in enumerate(starts):
        end = starts[k + 1][0] if k + 1 < len(starts) else len(ids) + 64
        for i in range(idx, end):
            name_of[i] = name
    # read-back kernels (copy_d2h of a few bytes into the mailbox, the SM copy
    # lane) are launched outside the brackets and not counted by launch_count:
    # they get their own name and do not advance the bracket position
    uncounted = re.compile(r"small_copy_kernel|sm_copy_kernel")
    kern = collections.OrderedDict()
    bracket_pos = 0
    md = ["| # | bracket (bench.py name) | kernel | grid x block | ms (ncu: cold, serialised) | DRAM read MB | DRAM write MB | DRAM GB/s |",
          "|---|---|---|---|---:|---:|---:|---:|"]
    tt = tr = tw = 0.0
    for pos, i in enumerate(ids):
        d = launches[i]
        if uncounted.search(d["fn"]):
            name = "(read-back store into page-locked memory)"
        else:
            name = name_of.get(bracket_pos, "(unnamed)")
            bracket_pos += 1
        t = d["gpu__time_duration.sum"] / 1e6
        r_, w_ = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
        k = kern.setdefault(name, {"launches_per_step": 0, "dram_bytes_per_step": 0.0,
                                   "dram_read_per_step": 0.0, "dram_write_per_step": 0.0,
                                   "ncu_ms_per_step": 0.0})
        k["launches_per_step"] += 1
        k["dram_bytes_per_step"] += r_ + w_
        k["dram_read_per_step"] += r_
        k["dram_write_per_step"] += w_
        k["ncu_ms_per_step"] += t
        fn = re.sub(r"\(.*$", "", re.sub(r"^void ", "", d["fn"]))
        md.append("| %d | %s | `%s` | %s x %s | %.3f | %.1f | %.1f | %.0f |" %
                  (pos, name, fn, d["grid"], d["block"], t, r_ / 1e6, w_ / 1e6,
                   (r_ + w_) / t / 1e6 if t > 0 else 0))
        tt += t
        tr += r_
        tw += w_
    for k in kern.values():
        k["dram_bytes_per_launch"] = k["dram_bytes_per_step"] / k["launches_per_step"]
    md.append("| | | **total** | | %.3f | %.1f | %.1f | %.0f |" % (tt, tr / 1e6, tw / 1e6, (tr + tw) / tt / 1e6))
    json.dump({"captured": out_md, "cycle_ncu_ms": tt, "cycle_dram_bytes": tr + tw,
               "kernels": kern}, open(out_json, "w"), indent=1)
    open(out_md, "w").write(
        "# One C2 cycle (86,278,900 octants, 1 B200): every kernel launch with its DRAM traffic\n\n"
        "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
        "--clock-control none --profile-from-start off` around `bench.py --steps 1 --profiler-range`.\n"
        "Times are ncu's (cold cache, serialised); bench.py weights the DRAM bytes with live CUDA-event times.\n\n"
        + "\n".join(md) + "\n")
    print("%d launches, %.3f ms, %.2f GB DRAM, %.0f GB/s" % (len(ids), tt, (tr + tw) / 1e9, (tr + tw) / tt / 1e6))


if __name__ == "__main__":
    main()
