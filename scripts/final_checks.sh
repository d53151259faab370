# Round-end evidence run on one B200: parity tests, compute-sanitizer (4 tools), bench (both arms), ncu launch list and
# one full capture of the scan kernel.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; tail -2 gpurun_out/final_tests.log
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/$tool.log python scripts/sanitize_smoke.py > gpurun_out/san_$tool.out 2>&1
  tail -1 gpurun_out/san_$tool.out; tail -2 gpurun_out/$tool.log
done
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; tail -c 300 gpurun_out/bench_ref_final.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_scan_minimizers_v3 -s 3 -c 1 -o gpurun_out/prof_scan_v3_final python bench.py --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/ncu_final.log 2>&1
ls -la gpurun_out/prof_scan_v3_final.ncu-rep
