#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench lines, ncu launch list + --set full captures, measured peaks.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
TAG=${1:-r01}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 python bench.py > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err; echo "bench c2 rc=$?"
timeout 400 python bench.py --workload c3 --no-cpu-baseline > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "bench ref rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_launches_c2.log 2>&1; echo "launch list rc=$?"
for K in k_acsf:1:1 k_bpnn_mma:2:1; do
  C=${K##*:}; R=${K%:*}; S=${R##*:}; K=${R%%:*}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -f -o $O/${TAG}_$K \
      python tools/e2e_breakdown.py c2 10000 > $O/${TAG}_$K.log 2>&1
  ncu -i $O/${TAG}_$K.ncu-rep --page raw --csv > $O/${TAG}_$K.raw.csv 2>/dev/null
  echo "ncu $K rc=$?"
done
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/peaks tools/peaks.cu && timeout 120 /tmp/peaks > $O/${TAG}_peaks.txt 2>&1; echo "peaks rc=$?"
tail -3 $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_smoke.log | tail -2; cat $O/${TAG}_bench_c2.json; cat $O/${TAG}_peaks.txt
