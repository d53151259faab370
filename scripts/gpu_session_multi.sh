#!/bin/bash
# Multi-GPU box session: usage scripts/gpu_session_multi.sh <tag> <ngpus> [reads]
tag=${1:-m}; n=${2:-2}; reads=${3:-2000000}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/${tag}_gpu.txt 2>&1
nproc >> gpurun_out/${tag}_gpu.txt; free -g | head -2 >> gpurun_out/${tag}_gpu.txt
nvidia-smi topo -m >> gpurun_out/${tag}_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_gpu or segment" 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
tail -3 gpurun_out/${tag}_tests.log
timeout 900 python -m pytest tests/test_cli.py -m gpu -q 2>&1 | tail -5 >> gpurun_out/${tag}_tests.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --reads $reads --steps 5 --warmup 3 \
    > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err
tail -c 400 gpurun_out/${tag}_bench_n${n}.err; head -c 600 gpurun_out/${tag}_bench_n${n}.json
timeout 900 python scripts/cli_e2e.py --config 3 --reads 400000 --gpus $n > gpurun_out/${tag}_cli_e2e.json 2> gpurun_out/${tag}_cli_e2e.err; tail -c 300 gpurun_out/${tag}_cli_e2e.err; head -c 900 gpurun_out/${tag}_cli_e2e.json
lscpu | head -25 >> gpurun_out/${tag}_gpu.txt; lspci -tv 2>/dev/null | head -80 >> gpurun_out/${tag}_gpu.txt
if [ -x scripts/microbench/h2d ]; then timeout 300 scripts/microbench/h2d > gpurun_out/${tag}_h2d.json 2> gpurun_out/${tag}_h2d.err; fi
