set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest7.log
TMR_B200_NODES_VERBOSE=1 timeout 600 python bench.py --steps 5 --warmup 3 --profile-out gpurun_out/r2_kt7.json > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches7.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --profiler-range > gpurun_out/r2_ncu7.log 2>&1
tail -5 gpurun_out/r2_pytest7.log; cut -c1-1500 gpurun_out/r2_bench7.json; tail -5 gpurun_out/r2_bench7.err
