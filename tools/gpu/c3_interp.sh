cd $GRAFT_REPO_ROOT
timeout 1200 python bench_interp.py --level 5 --passes 4 --pct 27 --bulk-only --cpu-level 3 --cpu-passes 2 > gpurun_out/r2_interp_c3.json 2> gpurun_out/r2_interp_c3.err
tail -3 gpurun_out/r2_interp_c3.err; head -c 900 gpurun_out/r2_interp_c3.json
