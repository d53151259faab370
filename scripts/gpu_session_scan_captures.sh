#!/bin/bash
# ncu --set full of the two scan instantiations with the final library (512 Mi-base launches of config 2), raw + source pages as CSV
tag=${1:-cap}
mkdir -p gpurun_out
export MQ_SUB_BASES_RESIDENT=536870912
B2="python bench.py --config 2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-packed --check 0"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_scan_minimizers -s 3 -c 1 -o gpurun_out/${tag}_scan_packed -f $B2 > gpurun_out/${tag}_ncu_sp.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_scan_minimizers -s 13 -c 1 -o gpurun_out/${tag}_scan_ascii -f $B2 > gpurun_out/${tag}_ncu_sa.log 2>&1
for r in scan_packed scan_ascii; do ncu -i gpurun_out/${tag}_$r.ncu-rep --page raw --csv > gpurun_out/${tag}_${r}_ncu_raw.csv 2>/dev/null; done
ncu -i gpurun_out/${tag}_scan_packed.ncu-rep --page source --csv > gpurun_out/${tag}_scan_packed_source.csv 2>/dev/null
python - gpurun_out/${tag}_scan_packed_ncu_raw.csv gpurun_out/${tag}_scan_ascii_ncu_raw.csv <<'PY'
import csv,sys
for f in sys.argv[1:]:
    r=list(csv.reader(open(f))); h=r[0]; v=r[2]
    g=lambda k: v[h.index(k)]
    print(f, g("Kernel Name")[:40], g("launch__grid_size"), g("gpu__time_duration.sum"), g("smsp__inst_executed.sum"), g("dram__bytes_read.sum"), g("dram__bytes_write.sum"), g("smsp__issue_active.avg.pct_of_peak_sustained_active"))
PY
