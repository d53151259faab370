#!/bin/bash
# A/B of alternative builds of the library (same ABI) against the shipped one.  usage: gpu_session_ab.sh <tag> <lib>...
tag=$1; shift
mkdir -p gpurun_out
for lib in default "$@"; do
  name=$(basename $lib .so)
  if [ "$lib" != "default" ]; then export MQ_LIB=$PWD/$lib; else unset MQ_LIB; fi
  if [ "$lib" != "default" ]; then
    timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "minimizers or packed or config3" 2>&1 | tail -3 > gpurun_out/${tag}_${name}_tests.log; tail -1 gpurun_out/${tag}_${name}_tests.log
  fi
  for rep in 1 2; do
    timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e-packed --check 2000 > gpurun_out/${tag}_${name}_c2_$rep.json 2> gpurun_out/${tag}_${name}_c2_$rep.err
    timeout 600 python bench.py --reads 200000 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e-packed --check 2000 > gpurun_out/${tag}_${name}_c3_$rep.json 2> gpurun_out/${tag}_${name}_c3_$rep.err
  done
done
