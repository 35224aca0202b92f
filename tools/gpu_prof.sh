#!/bin/bash
# One 1-GPU call: e2e breakdown + ncu --set full captures (source pages) of the hot kernels.
#   gpurun --timeout 900 -- 'bash tools/gpu_prof.sh <tag>'
TAG=${1:-r01p}
O=gpurun_out; mkdir -p $O
timeout 200 python tools/e2e_breakdown.py c2 > $O/${TAG}_e2e_c2.txt 2>&1; echo "e2e rc=$?"
# full-size C2 launch of the ACSF value kernel: dram traffic per launch for bench.py's roofline.traffic
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_acsf -s 1 -c 1 -f -o $O/${TAG}_acsf_full \
    python tools/e2e_breakdown.py c2 10000 > $O/${TAG}_acsf_full.log 2>&1
ncu -i $O/${TAG}_acsf_full.ncu-rep --page raw --csv > $O/${TAG}_acsf_full.raw.csv 2>/dev/null
ncu -i $O/${TAG}_acsf_full.ncu-rep --page source --csv > $O/${TAG}_acsf_full.source.csv 2>/dev/null; echo "ncu acsf rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_bpnn -s 2 -c 2 -f -o $O/${TAG}_bpnn \
    python tools/e2e_breakdown.py c2 10000 > $O/${TAG}_bpnn.log 2>&1
ncu -i $O/${TAG}_bpnn.ncu-rep --page raw --csv > $O/${TAG}_bpnn.raw.csv 2>/dev/null
ncu -i $O/${TAG}_bpnn.ncu-rep --page source --csv > $O/${TAG}_bpnn.source.csv 2>/dev/null; echo "ncu bpnn rc=$?"
rm -f $O/${TAG}_bpnn.ncu-rep   # keep the merged output small; the csv pages are what is read
cat $O/${TAG}_e2e_c2.txt | tail -2
