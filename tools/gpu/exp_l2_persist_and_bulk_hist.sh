cd $GRAFT_REPO_ROOT
for M in 1 0; do
TMR_B200_L2_PERSIST=$M timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --profile-out gpurun_out/r2_kt16_l2$M.json > gpurun_out/r2_bench16_l2$M.json 2> /dev/null
python -c "
import json; b=json.load(open('gpurun_out/r2_bench16_l2$M.json')); k=json.load(open('gpurun_out/r2_kt16_l2$M.json')); print('l2persist=$M', b['ms_per_step'], 'hanging', k['nodes_hanging_info']['ms_per_step'], 'locate', k['nodes_slot_locate']['ms_per_step'], b['fingerprint']['checksum'])"
done
for H in tile bulk; do
echo "hist=$H"; TMR_B200_HIST=$H timeout 300 python tools/bench_radix.py 200000000 33 2>&1 | tail -2 | cut -c1-600
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
