cd $GRAFT_REPO_ROOT
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --profile-out gpurun_out/r2_kt_n2.json > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -3 gpurun_out/r2_bench_n2.err | cut -c1-200
