# usage: ncu_kernel.sh <kernel-regex> <out-name> [extra bench args]
set -x
cd $GRAFT_REPO_ROOT
K=$1; O=$2; shift 2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/$O python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/$O.log 2>&1
ls -la gpurun_out/
