#!/bin/bash
# bench.py at N GPUs (torchrun), nothing else.  usage: gpu_session_bench_n.sh <tag> <ngpus> [steps]
tag=${1:-b}; n=${2:-4}; steps=${3:-5}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $n --steps $steps --warmup 3 \
    > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err
tail -c 300 gpurun_out/${tag}_bench_n${n}.err; head -c 500 gpurun_out/${tag}_bench_n${n}.json
