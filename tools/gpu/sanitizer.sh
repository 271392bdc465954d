cd $GRAFT_REPO_ROOT
( echo "# compute-sanitizer --tool memcheck, round-2 kernels (1 B200)"; date -u
  echo "## smoke()"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke 2>&1 | tail -6
  echo "## pytest -m gpu -k 'refine_balance_nodes or node_locations or device_views or deep_corner'"
  timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity.py -m gpu -x -q -k "refine_balance_nodes or node_locations or device_views or deep_corner" 2>&1 | tail -6
  echo "## racecheck, smoke()"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke 2>&1 | tail -6
) > gpurun_out/sanitizer_r02.txt 2>&1
cat gpurun_out/sanitizer_r02.txt
