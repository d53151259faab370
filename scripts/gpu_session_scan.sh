#!/bin/bash
# scan-kernel work: seeding parity tests (both formats), the A/B of build/ab variants on config 2, a light ncu pass
tag=${1:-sc}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "minimizers or giant or split or kminmers or fuzz or non_acgt or edge_reads" 2>&1 | tail -15 > gpurun_out/${tag}_tests.log; tail -15 gpurun_out/${tag}_tests.log
bash scripts/ab_scan.sh ${tag}
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,sm__warps_active.avg.per_cycle_active,launch__grid_size,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
for lib in build/ab/libmq_*.so; do n=$(basename $lib .so); n=${n#libmq_}
MQ_LIB=$PWD/$lib timeout 600 ncu --metrics $M --clock-control none -k regex:k_scan_minimizers -s 4 -c 8 --csv --log-file gpurun_out/${tag}_ncu_$n.csv python bench.py --config 2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-packed --check 0 > /dev/null 2>&1
done
