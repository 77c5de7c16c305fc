#!/usr/bin/env bash
# usage: tools/gpu_multi8.sh tag : on an 8-GPU box -- bench.py under torchrun on 8 GPUs, then C5
# at its full size sharded over the 8 GPUs in one process
tag=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_8gpu_$tag.json 2> gpurun_out/bench_8gpu_$tag.err
tail -c 300 gpurun_out/bench_8gpu_$tag.json
timeout 600 python tools/run_c5_sharded.py 65536 > gpurun_out/c5_sharded_$tag.json 2> gpurun_out/c5_sharded_$tag.err
cat gpurun_out/c5_sharded_$tag.json; tail -3 gpurun_out/c5_sharded_$tag.err
