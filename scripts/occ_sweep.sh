mkdir -p gpurun_out
for pad in 0 1536 4096 8192 16384; do
  MQ_SCAN_PAD=$pad timeout 300 python bench.py --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/occ_$pad.json 2> gpurun_out/occ_$pad.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/occ_$pad.json").read().strip().splitlines()[-1])
print("pad $pad", round(d["ms_per_step"],3), round(d["stage_ms_last_step"]["scan_kernel"],3))
PY
done
