#!/bin/bash
# Full-size default bench (BASELINE configs[2], 2 M reads) at N=1 and the reference arm.   usage: gpu_session_full.sh <tag> [steps] [warmup]
tag=${1:-f}; steps=${2:-5}; warm=${3:-3}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/${tag}_gpu.txt 2>&1
nproc >> gpurun_out/${tag}_gpu.txt; free -g | head -2 >> gpurun_out/${tag}_gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/${tag}_gpu.txt
t0=$(date +%s.%N)
timeout 1500 python bench.py --steps $steps --warmup $warm > gpurun_out/${tag}_bench_full.json 2> gpurun_out/${tag}_bench_full.err
t1=$(date +%s.%N); echo "bench_full wall_s $(echo "$t1 - $t0" | bc)" | tee gpurun_out/${tag}_wall.txt
tail -c 500 gpurun_out/${tag}_bench_full.err; head -c 700 gpurun_out/${tag}_bench_full.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
t2=$(date +%s.%N); echo "bench_ref wall_s $(echo "$t2 - $t1" | bc)" | tee -a gpurun_out/${tag}_wall.txt
tail -c 300 gpurun_out/${tag}_bench_ref.err; head -c 400 gpurun_out/${tag}_bench_ref.json
python scripts/upstream_probe.py > gpurun_out/${tag}_upstream_probe.json 2>&1
