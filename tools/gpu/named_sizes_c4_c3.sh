cd $GRAFT_REPO_ROOT
# BASELINE configs[3] (C4): 1050 trees, ~200 M octants, balance(1) + createNodes
timeout 900 python bench.py --workload c4 --pct 35 --steps 3 --warmup 3 --no-cpu-baseline --no-parity --profile-out gpurun_out/r2_kt_c4.json > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err
tail -2 gpurun_out/r2_bench_c4.err; cut -c1-400 gpurun_out/r2_bench_c4.json
# BASELINE configs[2] (C3): 5-level hierarchy, ~50 M finest octants, createInterpolation orders 3->2->2..
timeout 1200 python bench_interp.py --level 5 --passes 4 --pct 27 --bulk-only --cpu-level 3 --cpu-passes 2 > gpurun_out/r2_interp_c3.json 2> gpurun_out/r2_interp_c3.err
tail -3 gpurun_out/r2_interp_c3.err; head -c 1500 gpurun_out/r2_interp_c3.json
