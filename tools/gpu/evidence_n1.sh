set -x
cd $GRAFT_REPO_ROOT
# (1) one cycle: every launch with DRAM bytes + the bracket names
TMR_B200_LAUNCH_LOG=gpurun_out/r2_launchlog14.txt timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches14.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --profiler-range > gpurun_out/r2_ncu14.log 2>&1
# (2) full captures of the four longest kernels of the cycle
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k 'regex:NodeSlotFn|HangingFn|HangingBoundaryFn|DepFillFn|SlotResolveWin2Fn|MapFillFn' -f -o gpurun_out/r2_top4b python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --profiler-range > gpurun_out/r2_top4b.log 2>&1
# (3) the driver's commands
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
timeout 900 python bench.py --profile-out gpurun_out/r2_kt_final_n1.json > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/r2_bench_final_n1.err
ls -la gpurun_out | tail -12
