cd $GRAFT_REPO_ROOT
for M in block elem; do
TMR_B200_HANGING=$M timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --profile-out gpurun_out/r2_kt11_$M.json > gpurun_out/r2_bench11_$M.json 2> /dev/null
python -c "
import json; b=json.load(open('gpurun_out/r2_bench11_$M.json')); k=json.load(open('gpurun_out/r2_kt11_$M.json')); print('$M', b['ms_per_step'], k['nodes_hanging_info'], b['fingerprint']['checksum'], b['parity'])"
done
