#!/bin/bash
# the bench lines of the round on one B200 (both arms, configs 3 / 2 / 4) and the launch list of the `value` region
tag=${1:-ln}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt 2>&1; nproc >> gpurun_out/${tag}_gpu.txt
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_full.json 2> gpurun_out/${tag}_bench_full.err; head -c 250 gpurun_out/${tag}_bench_full.json; echo
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; head -c 150 gpurun_out/${tag}_bench_ref.json; echo
timeout 600 python bench.py --config 2 --steps 10 --warmup 3 > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err
timeout 900 python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --reads 600000 --steps 2 --warmup 3 --value-only > gpurun_out/${tag}_ncu_launch.log 2>&1
tail -2 gpurun_out/${tag}_ncu_launch.log
