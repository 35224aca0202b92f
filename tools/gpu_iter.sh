#!/bin/bash
# One 1-GPU iteration call: GPU tests, C2 bench (auto path and cell-list path), C3 bench, e2e breakdown.
#   gpurun --timeout 1200 -- 'bash tools/gpu_iter.sh <tag> [pytest-args]'
TAG=${1:-it}; shift
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x "$@" > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 $O/${TAG}_pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err; echo "bench c2 rc=$?"
FNETGPU_ACSF_PATH=cells FNETGPU_MLP=legacy timeout 300 python bench.py --no-cpu-baseline > $O/${TAG}_bench_c2_cells.json 2> $O/${TAG}_bench_c2_cells.err; echo "bench c2 cells+legacy rc=$?"
FNETGPU_MLP=nofuse timeout 300 python bench.py --no-cpu-baseline > $O/${TAG}_bench_c2_nofuse.json 2> $O/${TAG}_bench_c2_nofuse.err; echo "bench c2 nofuse rc=$?"
timeout 300 python bench.py --workload c3 --no-cpu-baseline > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
timeout 200 python tools/e2e_breakdown.py c2 > $O/${TAG}_e2e_c2.txt 2>&1
for f in c2 c2_nofuse c2_cells c3; do python - $O/${TAG}_bench_$f.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g ms %.4g e2e %.4g e2e_ms %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]),
          {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
tail -1 $O/${TAG}_e2e_c2.txt
timeout 200 python tools/md_latency.py 500 > $O/${TAG}_md_latency.txt 2>&1; cat $O/${TAG}_md_latency.txt
