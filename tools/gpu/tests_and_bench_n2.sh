set -x
cd $GRAFT_REPO_ROOT
N=${1:-2}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_pytest_all_n$N.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --profile-out gpurun_out/r2_kt_n$N.json > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -3 gpurun_out/r2_pytest_all_n$N.log; tail -5 gpurun_out/r2_bench_n$N.err; cut -c1-300 gpurun_out/r2_bench_n$N.json
