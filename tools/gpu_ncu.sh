#!/bin/bash
# ncu --set full capture (source pages) of kernels matching a regex on the C2 / C3 workload:
#   gpurun --timeout 900 -- 'bash tools/gpu_ncu.sh <tag> <kernel-regex> <skip> <count> [workload] [nstruct]'
TAG=$1; KR=$2; SK=${3:-0}; CN=${4:-1}; WL=${5:-c2}; NS=${6:-10000}
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KR -s $SK -c $CN -f -o $O/$TAG \
    python tools/e2e_breakdown.py $WL $NS > $O/$TAG.log 2>&1; echo "ncu rc=$?"
ncu -i $O/$TAG.ncu-rep --page raw --csv > $O/$TAG.raw.csv 2>/dev/null
ncu -i $O/$TAG.ncu-rep --page source --csv > $O/$TAG.source.csv 2>/dev/null
rm -f $O/$TAG.ncu-rep   # the csv pages are what is read here; the merged output is limited to 64 MiB
ls -la $O/$TAG.*
