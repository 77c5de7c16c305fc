#!/usr/bin/env bash
# usage: tools/gpu_multi.sh N tag : bench.py under torchrun on N GPUs of one box
n=$1; tag=$2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${n}gpu_$tag.json 2> gpurun_out/bench_${n}gpu_$tag.err
tail -c 400 gpurun_out/bench_${n}gpu_$tag.json
