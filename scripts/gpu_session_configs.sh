#!/bin/bash
# BASELINE configs 4 and 5 at their stated sizes on one B200 (config 5's full grid on the 310 Mbp cut, every point checked
# against the CPU oracle; its three 1-D sweeps on the full 3.1 Gbp genome).   usage: gpu_session_configs.sh <tag>
tag=${1:-c}
mkdir -p gpurun_out
timeout 1200 python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_config4.json 2> gpurun_out/${tag}_bench_config4.err; head -c 300 gpurun_out/${tag}_bench_config4.json; echo
timeout 1500 python scripts/sweep_config5.py --scale 10 --reads 20000 --full-grid --parity-all 5000 > gpurun_out/${tag}_config5_fullgrid_parity.csv 2> gpurun_out/${tag}_config5_fullgrid.err; tail -3 gpurun_out/${tag}_config5_fullgrid_parity.csv
timeout 1200 python scripts/sweep_config5.py --reads 100000 > gpurun_out/${tag}_config5_sweep.csv 2> gpurun_out/${tag}_config5_sweep.err; tail -3 gpurun_out/${tag}_config5_sweep.csv
