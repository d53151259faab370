#!/bin/bash
# On-the-fly host packing (mq_set_host_threads): parity tests, then the default bench (config 3, full size).
tag=${1:-hy}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt 2>&1
nproc >> gpurun_out/${tag}_gpu.txt; lscpu | grep -E "Model name|Socket|NUMA|L3" >> gpurun_out/${tag}_gpu.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "host_threads or multi_gpu_context or test_map" 2>&1 | tail -5 > gpurun_out/${tag}_tests.log; tail -2 gpurun_out/${tag}_tests.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_full.json 2> gpurun_out/${tag}_bench_full.err; head -c 300 gpurun_out/${tag}_bench_full.json; echo
python - gpurun_out/${tag}_bench_full.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ("e2e","e2e_ascii_link_only","e2e_prepacked","e2e_packed"):
    print(k, {x:d[k].get(x) for x in ("value","gbp_per_s","h2d_bytes_per_step","bases_packed_on_host_fraction","stage_ms_last_step_rank0","pack_gb_per_s_rank0")})
print("value", d["value"], "parity", d["parity"])
PY
timeout 600 python bench.py --config 2 --steps 10 --warmup 3 > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err; head -c 200 gpurun_out/${tag}_bench_c2.json; echo
