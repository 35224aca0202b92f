#!/bin/bash
# N-GPU call: bench lines at N = 1, 2, 4, 8 (as many as the box has), launched the way the driver does.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_scale.sh <tag>'
TAG=${1:-scale}
O=gpurun_out; mkdir -p $O
NG=$(nvidia-smi -L | wc -l)
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_n1.json 2> $O/${TAG}_n1.err; echo "n1 rc=$?"
for N in 2 4 8; do
  [ $N -le $NG ] || continue
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 5 --warmup 3 > $O/${TAG}_n$N.json 2> $O/${TAG}_n$N.err; echo "n$N rc=$?"
done
N=$NG
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29600 \
    bench.py --gpus $N --workload c3 --steps 3 --warmup 3 > $O/${TAG}_c3_n$N.json 2> $O/${TAG}_c3_n$N.err; echo "c3 n$N rc=$?"
for f in $O/${TAG}_n*.json $O/${TAG}_c3_n*.json; do python - $f <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n_gpus %d value %.4g ms %.4g e2e %.4g" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
