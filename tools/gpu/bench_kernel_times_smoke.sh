cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --profile-out gpurun_out/r2_kt17.json > gpurun_out/r2_bench17.json 2> /dev/null
python -c "
import json; b=json.load(open('gpurun_out/r2_bench17.json')); k=json.load(open('gpurun_out/r2_kt17.json')); print(b['ms_per_step'], b['fingerprint']['checksum'], b['parity'], b['e2e']['ms_per_step']); [print('%-26s %.3f'%(n,v['ms_per_step'])) for n,v in list(k.items())[:8]]"
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2
