set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_pytest9.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2_kt9.json > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err
tail -3 gpurun_out/r2_pytest9.log; tail -5 gpurun_out/r2_bench9.err; cut -c1-300 gpurun_out/r2_bench9.json
