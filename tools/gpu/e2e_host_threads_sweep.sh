cd $GRAFT_REPO_ROOT
nproc; numactl -H 2>/dev/null | head -5; nvidia-smi topo -m 2>/dev/null | head -6
for T in 2 4 8 16; do
TMR_B200_HOST_THREADS=$T timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2_bench10_t$T.json 2> /dev/null
python -c "
import json; b=json.load(open('gpurun_out/r2_bench10_t$T.json')); print($T, b['e2e']['ms_per_step'], b['e2e']['host_ms_per_call'])"
done
