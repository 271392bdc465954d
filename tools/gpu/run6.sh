set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_pytest6.log
tail -5 gpurun_out/r2_pytest6.log
