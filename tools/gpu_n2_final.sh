#!/bin/bash
# 2-GPU call for the final build: the tests that need two GPUs (NCCL world-2 gradient, single-process multi-GPU driver) and
# N = 1 / 2 bench lines of C2 (weak) and C3 (weak, strong), launched the way the driver does.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_n2_final.sh <tag>'
TAG=${1:-r02n2}
O=gpurun_out; mkdir -p $O
[ -n "$SKIP_TESTS" ] || timeout 600 python -m pytest tests -m gpu -q -k "nccl or multi_gpu or sharded or mg" > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 $O/${TAG}_pytest_gpu.log
run() { local name=$1 N=$2; shift 2
  if [ $N -eq 1 ]; then timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-c4 "$@" > $O/${TAG}_${name}_n1.json 2> $O/${TAG}_${name}_n1.err
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
        bench.py --gpus $N --steps 20 --warmup 5 --no-c4 "$@" > $O/${TAG}_${name}_n$N.json 2> $O/${TAG}_${name}_n$N.err; fi
  echo "$name n$N rc=$?"; python - $O/${TAG}_${name}_n$N.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   value %.4g atoms/s  ms %.4g  e2e %.4g (%.4g ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print("   unreadable", e)
PY
}
for N in ${NS:-1 2}; do run c2weak $N; run c3weak $N --workload c3; run c3strong $N --workload c3 --scaling strong; done
