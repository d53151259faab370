#!/bin/bash
# last pass of the round on one B200 with the final library: every GPU test, smoke, the four sanitizer tools, the default bench
tag=${1:-fz}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/${tag}_tests.log; tail -1 gpurun_out/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1200 compute-sanitizer --tool $tool --log-file gpurun_out/${tag}_sanitizer_$tool.log python scripts/sanitize_smoke.py > gpurun_out/${tag}_san_$tool.out 2>&1
  tail -1 gpurun_out/${tag}_san_$tool.out; tail -1 gpurun_out/${tag}_sanitizer_$tool.log
done
timeout 1500 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err; head -c 200 gpurun_out/${tag}_bench_default.json; echo
python - gpurun_out/${tag}_bench_default.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.2fM e2e %.2fM link %.2fM" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e_ascii_link_only"]["value"]/1e6), {k: d["e2e"][k] for k in ("host_threads_per_rank","calibration_ms_per_step","bases_packed_on_host_fraction")}, d["parity"]["paths_identical"], d["parity"]["hits_identical"], d["parity"]["checked_reads"])
PY
