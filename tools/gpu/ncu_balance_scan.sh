cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k 'regex:MapClosureFn|MapCountFn|SlotCount2Fn|DepLenFn|RefineCountFn' -f -o gpurun_out/r2_balance_scan python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --profiler-range > gpurun_out/r2_balance_scan.log 2>&1
ls -la gpurun_out/r2_balance_scan.ncu-rep
