cd $GRAFT_REPO_ROOT
timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_primitives_gpu.py -m gpu -x -q -k allocation_failure 2>&1 | grep -vE "^=========\s+at |Host Frame|^=========\s+in /" | head -60
