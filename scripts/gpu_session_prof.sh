#!/bin/bash
# ncu evidence of one round: launch list (shares of the step) + full captures of the scan (both formats), probe and insert
# kernels at human scale (3.1 Gbp index).   usage: gpu_session_prof.sh <tag>
tag=${1:-p}
mkdir -p gpurun_out
B="python bench.py --reads 100000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-packed --check 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_launch.log 2>&1
# packed-resident value region: warm-up 3 + 2 steps, 3 sub-batches each (ramp 64 M, then 512 M ...): launch 7 is a 512 Mbase one
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_scan_minimizers -s 31 -c 1 -o gpurun_out/${tag}_scan_packed -f $B > gpurun_out/${tag}_ncu_sp.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_probe_match -s 7 -c 1 -o gpurun_out/${tag}_probe -f $B > gpurun_out/${tag}_ncu_pr.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_insert_kminmers -c 1 -o gpurun_out/${tag}_insert -f $B > gpurun_out/${tag}_ncu_in.log 2>&1
ls -la gpurun_out | head -30
