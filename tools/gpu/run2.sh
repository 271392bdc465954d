set -x
cd $GRAFT_REPO_ROOT
for B in 22 24; do
TMR_B200_IXBITS=$B timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2_kt2_ix$B.json > gpurun_out/r2_bench2_ix$B.json 2> gpurun_out/r2_bench2_ix$B.err
done
bash tools/gpu/ncu_kernel.sh block3_kernel r2_locate2
