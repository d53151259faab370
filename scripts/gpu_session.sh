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
if [ "$2" = "prof" ]; then
  # launch list of two bench steps (shares of the step), then full captures of the scan kernel in both formats
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
      python bench.py --config 2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-packed --check 0 > gpurun_out/${tag}_ncu_launch.log 2>&1
  # value region of config 2: packed-resident launches come first (warm-up 3 + 2 steps = 5 calls x 2 sub-batches), then ASCII-resident
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_scan_minimizers -s 6 -c 1 -o gpurun_out/${tag}_scan_packed -f \
      python bench.py --config 2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-packed --check 0 > gpurun_out/${tag}_ncu_p.log 2>&1
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_scan_minimizers -s 16 -c 1 -o gpurun_out/${tag}_scan_ascii -f \
      python bench.py --config 2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-packed --check 0 > gpurun_out/${tag}_ncu_a.log 2>&1
  ls -la gpurun_out/
fi
