# A/B helper: parity tests with the in-tree library, then bench.py once per scan-kernel generation / variant library.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_ab.log 2>&1; tail -5 gpurun_out/t_ab.log
MQ_SCAN_V2=1 timeout 300 python bench.py --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/ab_v2.json 2> gpurun_out/ab_v2.err
timeout 300 python bench.py --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/ab_v3.json 2> gpurun_out/ab_v3.err
for f in build/libmq_*.so; do
  v=$(basename $f .so)
  MQ_LIB=$PWD/$f timeout 300 python bench.py --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
done
for f in gpurun_out/ab_*.json; do python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["config"]["stage_ms_last_step"].items()} if "stage_ms_last_step" in d.get("config",{}) else {k: round(v,3) for k,v in d.get("stage_ms_last_step",{}).items()})
except Exception as e: print("$f", "ERR", e)
PY
done
