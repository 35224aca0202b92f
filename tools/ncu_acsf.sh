#!/bin/bash
# ncu --set full capture of the ACSF value kernel on a reduced C2 workload (run under gpurun)
#   tools/ncu_acsf.sh <tag> [workload] [nstruct] [kernel-regex]
TAG=${1:-acsf}; WL=${2:-c2}; NS=${3:-2000}; KR=${4:-k_acsf}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$KR -s 1 -c 1 -f -o gpurun_out/$TAG \
    python tools/e2e_breakdown.py $WL $NS > gpurun_out/$TAG.log 2>&1
ncu -i gpurun_out/$TAG.ncu-rep --page raw --csv > gpurun_out/$TAG.raw.csv 2>/dev/null
echo done
