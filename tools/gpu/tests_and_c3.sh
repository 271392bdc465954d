cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/gpu/c3_interp.sh > /dev/null 2>&1
python -c "
import json
c=json.load(open('gpurun_out/r2_interp_c3.json'))
for l in c['b200']: print(l['rows'], l['nnz'], 'device', round(l['device_csr_s'],4), 'bulk', round(l['api_bulk_csr_s'],3), l['max_rowsum_err'])"
