#!/bin/bash
# quick GPU iteration: a pytest -k subset, then C2 / C3 bench lines
#   gpurun --timeout 900 -- 'bash tools/gpu_quick.sh <tag> "<pytest -k expr>" [workloads]'
TAG=${1:-q}; KEXPR=$2; WLS=${3:-"c2 c3"}
O=gpurun_out; mkdir -p $O
if [ -n "$KEXPR" ]; then
timeout 900 python -m pytest tests -m gpu -q -x -k "$KEXPR" > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 $O/${TAG}_pytest_gpu.log
fi
for wl in $WLS; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline > $O/${TAG}_bench_$wl.json 2> $O/${TAG}_bench_$wl.err; echo "bench $wl rc=$?"
  python - $O/${TAG}_bench_$wl.json <<'PY'
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
