#!/bin/bash
# A/B of several builds of the library: quick parity subset with the default build, then C2 / C3 bench per build
#   gpurun -- 'bash tools/gpu_ab2.sh <tag> "<pytest -k>"'
TAG=${1:-ab}; KEXPR=$2
O=gpurun_out; mkdir -p $O
if [ -n "$KEXPR" ]; then
timeout 900 python -m pytest tests -m gpu -q -x -k "$KEXPR" > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 $O/${TAG}_pytest_gpu.log
fi
for L in fortnet_b200/libfnetgpu.so fortnet_b200/libfnetgpu_ab*.so; do
  b=$(basename $L .so)
  for wl in c2 c3; do
    FNETGPU_LIB=$PWD/$L timeout 300 python bench.py --workload $wl --no-cpu-baseline > $O/${TAG}_${b}_$wl.json 2> $O/${TAG}_${b}_$wl.err
    python - $O/${TAG}_${b}_$wl.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g ms %.4g e2e_ms %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]),
          {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[1], "unreadable", e)
    print(open(sys.argv[1].replace(".json", ".err")).read()[-800:])
PY
  done
done
