#!/bin/bash
# A/B of several builds of the library on the same box:  tools/ab.sh [workload] [nstruct]
#   fortnet_b200/libfnetgpu.so and every fortnet_b200/libfnetgpu_ab*.so (FNETGPU_LIB override)
WL=${1:-c2}; NS=${2:-10000}
for rep in 1 2; do
  for L in fortnet_b200/libfnetgpu.so fortnet_b200/libfnetgpu_ab*.so; do
    echo "== $L"; FNETGPU_LIB=$PWD/$L python tools/e2e_breakdown.py $WL $NS | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['total_ms'], d['kernels_ms'])"
  done
done
