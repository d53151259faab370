#!/bin/bash
# One GPU-box session: parity tests, then the bench lines that matter, everything logged under gpurun_out/.
# usage: scripts/gpu_session.sh [tag]
tag=${1:-s}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/${tag}_gpu.txt 2>&1
nproc >> gpurun_out/${tag}_gpu.txt; free -g | head -2 >> gpurun_out/${tag}_gpu.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_cli.py tests/test_abi.py tests/test_bench_contract.py -m gpu -q --maxfail=12 2>&1 | tail -150 > gpurun_out/${tag}_tests.log
tail -5 gpurun_out/${tag}_tests.log
timeout 300 python bench.py --config 2 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err
tail -c 600 gpurun_out/${tag}_bench_c2.err
timeout 600 python bench.py --reads 200000 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_c3_200k.json 2> gpurun_out/${tag}_bench_c3_200k.err
tail -c 600 gpurun_out/${tag}_bench_c3_200k.err
head -c 1500 gpurun_out/${tag}_bench_c3_200k.json
