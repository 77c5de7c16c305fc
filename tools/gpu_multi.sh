#!/usr/bin/env bash
# usage: tools/gpu_multi.sh N tag : bench.py under torchrun on N GPUs of one box (the weak-scaling
# line with the strong-scaling leg of the 1 M batch inside it)
n=$1; tag=$2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${n}gpu_$tag.json 2> gpurun_out/bench_${n}gpu_$tag.err
python - <<PY
import json
l = json.loads(open("gpurun_out/bench_${n}gpu_$tag.json").read().strip().splitlines()[-1])
s = l.get("strong_1M", {})
print("weak", l["value"], "e2e", l["e2e"]["value"], "ms", l["ms_per_step"])
print("strong", {k: s.get(k) for k in ("value", "e2e", "ms_per_step", "efficiency_vs_n1", "e2e_efficiency_vs_n1", "bits_equal_rank_local_result", "unavailable")})
PY
