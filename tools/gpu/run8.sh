set -x
cd $GRAFT_REPO_ROOT
# one cycle: launch list with DRAM bytes + bracket names
TMR_B200_LAUNCH_LOG=gpurun_out/r2_launchlog8.txt timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches8.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --profiler-range > gpurun_out/r2_ncu8.log 2>&1
# full captures of the four longest kernels of the cycle
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:block3_kernel|HangingFn|DepFillFn|DepWinnerFn' -f -o gpurun_out/r2_top4 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --profiler-range > gpurun_out/r2_top4.log 2>&1
ls -la gpurun_out
