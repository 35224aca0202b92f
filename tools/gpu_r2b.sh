#!/bin/bash
# lean-kernel A/B: selected tests, then C2 / C3 bench lines for each G (central atoms per warp)
TAG=${1:-r02b}; KEXPR=$2
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "$KEXPR" > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 $O/${TAG}_pytest_gpu.log
for g in 0 1 2 4; do
  FNETGPU_LEAN_G=$g timeout 300 python bench.py --no-cpu-baseline > $O/${TAG}_bench_c2_g$g.json 2> $O/${TAG}_bench_c2_g$g.err; echo "bench c2 G=$g rc=$?"
done
for g in 0 1 2; do
  FNETGPU_LEAN_G=$g timeout 300 python bench.py --workload c3 --no-cpu-baseline > $O/${TAG}_bench_c3_g$g.json 2> $O/${TAG}_bench_c3_g$g.err; echo "bench c3 G=$g rc=$?"
done
for f in c2_g0 c2_g1 c2_g2 c2_g4 c3_g0 c3_g1 c3_g2; do python - $O/${TAG}_bench_$f.json <<'PY'
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
