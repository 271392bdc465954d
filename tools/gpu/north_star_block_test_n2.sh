cd $GRAFT_REPO_ROOT
TMR_B200_NORTH_STAR_WORLD=2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 --no-parity > gpurun_out/r2_bench_ns_test.json 2> gpurun_out/r2_bench_ns_test.err
tail -3 gpurun_out/r2_bench_ns_test.err; python -c "
import json; b=json.load(open('gpurun_out/r2_bench_ns_test.json')); print(b['ms_per_step'], b['north_star_1e9_octants'], b['parity'])"
