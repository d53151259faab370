# A/B helper: full GPU parity tests with the in-tree library (unless SKIP_TESTS=1), then bench.py (and the scan
# parity tests) once per variant library under build/ (same ABI, different compile-time options).
# PADS="0 3500" additionally sweeps MQ_SCAN_PAD (extra dynamic smem per CTA = lower occupancy) with the in-tree library.
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_ab.log 2>&1; tail -3 gpurun_out/t_ab.log; fi
for pad in ${PADS:-0}; do
  MQ_SCAN_PAD=$pad MQ_DEBUG=1 timeout 300 python bench.py --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/ab_default_pad$pad.json 2> gpurun_out/ab_default_pad$pad.err
done
for f in build/libmq_*.so; do
  [ -e "$f" ] || continue
  v=$(basename $f .so)
  MQ_DEBUG=1 MQ_LIB=$PWD/$f timeout 300 python bench.py --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  if [ -z "$SKIP_TESTS" ]; then MQ_LIB=$PWD/$f timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "minimizers or segment" > gpurun_out/t_$v.log 2>&1; tail -1 gpurun_out/t_$v.log; fi
done
for f in gpurun_out/ab_*.json; do python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms_last_step"].items()})
except Exception as e: print("$f", "ERR", e)
PY
done
grep -h "\[mq\]" gpurun_out/ab_*.err | sort | uniq -c
