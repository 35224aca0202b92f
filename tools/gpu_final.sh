#!/bin/bash
# One gpurun call for the round's final build: parity suite, smoke, bench lines (C2 with the CPU baseline and C4, C3,
# precision 32, the reference arm), ncu launch list of the bench command, ncu --set full captures of the four hot
# kernels (raw + source pages as csv), MD-step latency, measured peaks.
#   gpurun --timeout 1800 -- 'bash tools/gpu_final.sh <tag>'
TAG=${1:-r02final}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 500 python bench.py > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err; echo "bench c2 rc=$?"
timeout 400 python bench.py --workload c3 --no-cpu-baseline --no-c4 > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
timeout 300 python bench.py --precision 32 --no-cpu-baseline --no-c4 > $O/${TAG}_bench_c2_fp32.json 2> $O/${TAG}_bench_c2_fp32.err; echo "bench c2 fp32 rc=$?"
timeout 300 python bench.py --workload c3 --precision 32 --no-cpu-baseline --no-c4 > $O/${TAG}_bench_c3_fp32.json 2> $O/${TAG}_bench_c3_fp32.err; echo "bench c3 fp32 rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "bench ref rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c4 > $O/${TAG}_launches_c2.log 2>&1; echo "launch list rc=$?"
bash tools/gpu_ncu.sh ${TAG}_lean_c2 k_acsf_lean 1 1 c2 10000
bash tools/gpu_ncu.sh ${TAG}_lean_c3 k_acsf_lean 1 1 c3 2000
bash tools/gpu_ncu.sh ${TAG}_mma_c2 k_bpnn_mma 1 1 c2 10000
bash tools/gpu_ncu.sh ${TAG}_mma_c3 k_bpnn_mma 1 1 c3 2000
timeout 200 python tools/md_latency.py 2000 > $O/${TAG}_md_latency.txt 2>&1; echo "md latency rc=$?"
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/peaks tools/peaks.cu && timeout 120 /tmp/peaks > $O/${TAG}_peaks.txt 2>&1; echo "peaks rc=$?"
tail -3 $O/${TAG}_pytest_gpu.log; tail -2 $O/${TAG}_smoke.log; cat $O/${TAG}_bench_c2.json | cut -c1-600; tail -5 $O/${TAG}_md_latency.txt
