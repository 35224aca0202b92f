#!/bin/bash
# ncu --set full capture of one launch of each socket-step kernel (192-atom cell): raw + source pages as csv
#   gpurun --timeout 900 -- 'bash tools/gpu_ncu_sock.sh <tag> "<kernel regex> ..."'
TAG=$1; shift
O=gpurun_out; mkdir -p $O
for KR in "$@"; do
  N=$(echo $KR | tr -c 'a-zA-Z0-9_' '_')
  FNETGPU_GRAPHS=0 FNETGPU_PDL=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$KR -s 40 -c 1 -f -o $O/${TAG}_$N \
      python tools/sock_profile.py 5 > $O/${TAG}_$N.log 2>&1; echo "ncu $KR rc=$?"
  ncu -i $O/${TAG}_$N.ncu-rep --page raw --csv > $O/${TAG}_$N.raw.csv 2>/dev/null
  ncu -i $O/${TAG}_$N.ncu-rep --page source --csv > $O/${TAG}_$N.source.csv 2>/dev/null
  rm -f $O/${TAG}_$N.ncu-rep
done
ls -la $O/${TAG}_*
