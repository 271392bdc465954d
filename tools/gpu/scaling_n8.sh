set -x
cd $GRAFT_REPO_ROOT
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# N = 8: parity gate, weak-scaling cycle, and the 1e9-octant forest (north_star_1e9_octants in the line)
timeout 900 $T --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --profile-out gpurun_out/r2_kt_n8.json > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
grep "multi-gpu" gpurun_out/r2_bench_n8.err | cut -c1-140; cut -c1-200 gpurun_out/r2_bench_n8.json
timeout 600 $T --nproc-per-node 4 --master-port 29524 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
cut -c1-200 gpurun_out/r2_bench_n4.json
timeout 600 $T --nproc-per-node 2 --master-port 29525 bench.py --gpus 2 --steps 5 --warmup 3 --profile-out gpurun_out/r2_kt_n2.json > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
cut -c1-200 gpurun_out/r2_bench_n2.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_samebox.json 2> /dev/null
cut -c1-200 gpurun_out/r2_bench_n1_samebox.json
