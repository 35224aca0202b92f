#!/bin/bash
# end-to-end A/B of library builds on one box, alternating (A B A B) to cancel warm-up / order effects:
#   gpurun -- 'bash tools/gpu_e2e_ab.sh <tag>'    (fortnet_b200/libfnetgpu.so and every libfnetgpu_ab*.so)
TAG=${1:-e2eab}; O=gpurun_out; mkdir -p $O
for rep in 1 2; do
  for L in fortnet_b200/libfnetgpu.so fortnet_b200/libfnetgpu_ab*.so; do
    b=$(basename $L .so)
    FNETGPU_LIB=$PWD/$L timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-c4 > $O/${TAG}_${b}_$rep.json 2> $O/${TAG}_${b}_$rep.err
    python - $O/${TAG}_${b}_$rep.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
st = sorted(d["e2e"]["ms_steps_rank0"])
print(sys.argv[1], "device %.4f ms  e2e mean %.4f  median %.4f  min %.4f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], st[len(st) // 2], st[0]))
PY
  done
done
