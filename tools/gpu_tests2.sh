#!/bin/bash
# 2-GPU call: full GPU test suite (incl. the NCCL world-2 test) + N=1/N=2 bench lines
TAG=${1:-r01n2}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --durations=15 > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err; echo "bench n2 rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench n1 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_c3_n2.json 2> $O/${TAG}_bench_c3_n2.err; echo "bench c3 n2 rc=$?"
cat $O/${TAG}_bench_n1.json | cut -c1-300; cat $O/${TAG}_bench_c3_n2.json | cut -c1-300
tail -30 $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_bench_n2.json | cut -c1-400; tail -5 $O/${TAG}_bench_n2.err
