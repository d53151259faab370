set -x
mkdir -p gpurun_out
./build/pipes > gpurun_out/pipes.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_rot2.log 2>&1; tail -3 gpurun_out/t_rot2.log
for v in base rot0 rot1 rot3 rot4; do
  MQ_LIB=$PWD/build/libmq_$v.so timeout 300 python bench.py --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
done
timeout 300 python bench.py --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/ab_rot2.json 2> gpurun_out/ab_rot2.err
for v in base rot0 rot1 rot2 rot3 rot4; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1])
    print("$v", d["value"], d["ms_per_step"], d["config"].get("stage_ms_last_step"))
except Exception as e: print("$v", "ERR", e)
PY
done
cat gpurun_out/pipes.txt
