set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_pytest5.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2_kt5.json > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
bash tools/gpu/ncu_kernel.sh block3_kernel r2_locate5
tail -3 gpurun_out/r2_pytest5.log
