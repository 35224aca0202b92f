#!/bin/bash
# Round-2 iteration call: selected GPU tests, then C2 / C3 bench lines (lean and generic ACSF kernel).
#   gpurun --timeout 1500 -- 'bash tools/gpu_r2.sh <tag> "<pytest -k expr or empty for all>"'
TAG=${1:-r02}; KEXPR=$2
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/${TAG}_gpu.txt 2>&1
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x -k "$KEXPR" > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
else
  timeout 1200 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
fi
tail -25 $O/${TAG}_pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err; echo "bench c2 rc=$?"
FNETGPU_ACSF_KERNEL=generic timeout 300 python bench.py --no-cpu-baseline > $O/${TAG}_bench_c2_generic.json 2> $O/${TAG}_bench_c2_generic.err; echo "bench c2 generic rc=$?"
timeout 300 python bench.py --workload c3 --no-cpu-baseline > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
FNETGPU_ACSF_KERNEL=generic timeout 300 python bench.py --workload c3 --no-cpu-baseline > $O/${TAG}_bench_c3_generic.json 2> $O/${TAG}_bench_c3_generic.err; echo "bench c3 generic rc=$?"
for f in c2 c2_generic c3 c3_generic; do python - $O/${TAG}_bench_$f.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g ms %.4g e2e %.4g e2e_ms %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]),
          {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[1], "unreadable", e)
    print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
done
