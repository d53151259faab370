#!/bin/bash
# Round-end evidence on ONE B200: GPU parity tests, compute-sanitizer (4 tools), both bench arms at full size, the ncu
# launch list and full captures of the kernels DESIGN.md / profiles/README.md quote.  Outputs land in gpurun_out/<tag>_*.
tag=${1:-ev}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt 2>&1
nproc >> gpurun_out/${tag}_gpu.txt; free -g | head -2 >> gpurun_out/${tag}_gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/${tag}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/${tag}_tests.log; tail -1 gpurun_out/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_full.json 2> gpurun_out/${tag}_bench_full.err; head -c 300 gpurun_out/${tag}_bench_full.json; echo
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; head -c 200 gpurun_out/${tag}_bench_ref.json; echo
timeout 600 python bench.py --config 2 --steps 10 --warmup 3 > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1200 compute-sanitizer --tool $tool --log-file gpurun_out/${tag}_sanitizer_$tool.log python scripts/sanitize_smoke.py > gpurun_out/${tag}_san_$tool.out 2>&1
  tail -1 gpurun_out/${tag}_san_$tool.out; tail -2 gpurun_out/${tag}_sanitizer_$tool.log
done
B3="python bench.py --reads 200000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-packed --check 0"
B2="python bench.py --config 2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-packed --check 0"
# launch list of the `value` region alone (index build + 5 packed-resident steps of 200,000 reads on config 3)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${tag}_launches.csv $B3 --value-only > gpurun_out/${tag}_ncu_launch.log 2>&1
# config 2 scan launches: 0-2 index builds (ASCII, packed, packed), then per resident call one sub-batch (1 Gbp < 2 Gi bases);
# MQ_SUB_BASES_RESIDENT=512Mi cuts a call into 512 Mi bases + the rest, the launch size of the committed captures:
# launch 3 = the first 512 Mi-base packed one; the ASCII-resident region starts at launch 13 with the 512 Mi-base ASCII one
export MQ_SUB_BASES_RESIDENT=536870912
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_scan_minimizers -s 3 -c 1 -o gpurun_out/${tag}_scan_packed -f $B2 > gpurun_out/${tag}_ncu_sp.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_scan_minimizers -s 13 -c 1 -o gpurun_out/${tag}_scan_ascii -f $B2 > gpurun_out/${tag}_ncu_sa.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_probe_match -s 7 -c 1 -o gpurun_out/${tag}_probe -f $B3 > gpurun_out/${tag}_ncu_pr.log 2>&1
MQ_NO_BLOOM=1 timeout 900 ncu --set full --clock-control none -k regex:k_probe_match -s 7 -c 1 -o gpurun_out/${tag}_probe_nobloom -f $B3 > gpurun_out/${tag}_ncu_prn.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_insert_kminmers -c 1 -o gpurun_out/${tag}_insert -f $B3 > gpurun_out/${tag}_ncu_in.log 2>&1
unset MQ_SUB_BASES_RESIDENT
for r in scan_packed scan_ascii probe probe_nobloom insert; do ncu -i gpurun_out/${tag}_$r.ncu-rep --page raw --csv > gpurun_out/${tag}_${r}_ncu_raw.csv 2>/dev/null; done
ncu -i gpurun_out/${tag}_scan_packed.ncu-rep --page source --csv > gpurun_out/${tag}_scan_packed_source.csv 2>/dev/null
timeout 600 python scripts/cli_e2e.py --config 3 --reads 200000 > gpurun_out/${tag}_cli_e2e.json 2> gpurun_out/${tag}_cli_e2e.err
ls -la gpurun_out | head -50
