#!/bin/bash
# force path A/B: parity tests of the force kernels, then the C4 line (bench.py's c4_force_prediction) per library build
TAG=${1:-c4ab}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "force or socket or smoke" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log
for L in fortnet_b200/libfnetgpu.so fortnet_b200/libfnetgpu_ab*.so; do
  b=$(basename $L .so)
  FNETGPU_LIB=$PWD/$L timeout 300 python -c "
import json, bench, fortnet_b200 as fb
print(json.dumps(bench.c4_forces(fb)))" > $O/${TAG}_${b}.json 2> $O/${TAG}_${b}.err; echo "$b rc=$?"; cut -c1-400 $O/${TAG}_${b}.json; tail -2 $O/${TAG}_${b}.err
done
