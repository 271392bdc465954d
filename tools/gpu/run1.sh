set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_pytest1.log
TMR_B200_NODES_VERBOSE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2_kt1.json > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
TMR_B200_NODES=sort timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2_kt1_sort.json > gpurun_out/r2_bench1_sort.json 2> gpurun_out/r2_bench1_sort.err
tail -3 gpurun_out/r2_pytest1.log; cat gpurun_out/r2_bench1.json | cut -c1-400
