#!/bin/bash
# N-GPU call (round 2): weak scaling of C2 and C3 at N = 1, 2, 4, 8 and STRONG scaling of C3 (20k structures split over N),
# launched the way the driver does.   gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_scale2.sh <tag>'
TAG=${1:-r02scale}
O=gpurun_out; mkdir -p $O
NG=$(nvidia-smi -L | wc -l)
run() {  # name N extra-args
  local name=$1 N=$2; shift 2
  if [ $N -eq 1 ]; then
    timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-c4 "$@" > $O/${TAG}_${name}_n1.json 2> $O/${TAG}_${name}_n1.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
        bench.py --gpus $N --steps 20 --warmup 5 --no-c4 "$@" > $O/${TAG}_${name}_n$N.json 2> $O/${TAG}_${name}_n$N.err
  fi
  echo "$name n$N rc=$?"
}
for N in 1 2 4 8; do
  [ $N -le $NG ] || continue
  run c2weak $N
  run c3weak $N --workload c3
  run c3strong $N --workload c3 --scaling strong
done
python - $O $TAG <<'PY'
import json, sys, glob, os
O, TAG = sys.argv[1:3]
for name in ("c2weak", "c3weak", "c3strong"):
    base = None
    for N in (1, 2, 4, 8):
        f = os.path.join(O, "%s_%s_n%d.json" % (TAG, name, N))
        try:
            d = json.loads(open(f).read().strip().splitlines()[-1])
        except Exception as e:
            print(name, N, "unreadable", e); continue
        if N == 1: base = d["value"]
        eff = d["value"] / (base * N) if (base and name != "c3strong") else (d["value"] / (base * N) if base else None)
        print(name, "N=%d value %.4g atoms/s ms %.4g e2e %.4g eff %.3f" % (N, d["value"], d["ms_per_step"], d["e2e"]["value"], eff or 0))
PY
