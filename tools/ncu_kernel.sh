#!/bin/bash
# ncu --set full capture of one kernel (run under gpurun)
#   tools/ncu_kernel.sh <tag> <kernel-regex> <skip> <count> <workload> <nstruct>
TAG=$1; KR=$2; SK=${3:-0}; CN=${4:-1}; WL=${5:-c2}; NS=${6:-2000}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$KR -s $SK -c $CN -f -o gpurun_out/$TAG \
    python tools/e2e_breakdown.py $WL $NS > gpurun_out/$TAG.log 2>&1
ncu -i gpurun_out/$TAG.ncu-rep --page raw --csv > gpurun_out/$TAG.raw.csv 2>/dev/null
echo done $TAG
