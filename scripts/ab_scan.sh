#!/bin/bash
# A/B of scan-kernel build variants (build/ab/libmq_<name>.so, same ABI; see the loop in DESIGN.md section 5 "knobs"):
# config 2 with the reads resident, scan-kernel milliseconds per step and the in-run path parity flag.
tag=${1:-ab}
out=gpurun_out/${tag}_scan_variants.txt; : > $out
for lib in build/ab/libmq_*.so; do
  n=$(basename $lib .so); n=${n#libmq_}
  MQ_LIB=$PWD/$lib timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e-packed --check 0 2> gpurun_out/${tag}_$n.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('$n', 'value %.2f M reads/s' % (d['value']/1e6), 'ascii %.2f' % (d['value_ascii']/1e6), 'scan_kernel ms/step packed %.4f ascii %.4f' % (r['avg_launch_ms']*r['launches']/d['steps'], r['ascii_kernel']['avg_launch_ms']*r['launches']/d['steps']), 'stages', d['stage_ms_last_step_rank0'], 'paths_identical', d['parity']['paths_identical'])
" >> $out 2>&1
done
cat $out
