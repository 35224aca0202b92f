O=gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-c4 $EXTRA > $O/r03x_$tag.json 2> $O/r03x_$tag.err; python - $O/r03x_$tag.json $tag <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms %.4g" % d["ms_per_step"], {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()}, d.get("grad_launch"))
except Exception as e:
    print(sys.argv[2], "unreadable", e)
PY
}
EXTRA="--workload c3"
run c3_base X=1
run c3_g1 FNETGPU_LEAN_G=1
run c3_apw64 FNETGPU_ACSF_ATOMS_PER_WARP=64
run c3_apw16 FNETGPU_ACSF_ATOMS_PER_WARP=16
run c3_cs6 FNETGPU_MLP_CLUSTER=6
EXTRA=""
run c2_g2 FNETGPU_LEAN_G=2
