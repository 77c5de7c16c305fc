#!/usr/bin/env bash
# usage: tools/gpu_final_profile.sh tag : the measurements profiles/ is built from, each pass
# only after the plain command has exited 0 (B200_PROFILING.md):
#   1. plain bench (the numbers)            -> gpurun_out/bench_$tag.json
#   2. ncu launch list of the same command  -> gpurun_out/launches_$tag.csv
#   3. one ncu --set full capture of the stepping kernel at the full batch size
tag=$1
mkdir -p gpurun_out
set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || exit 1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_$tag.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
  > gpurun_out/ncu_launches_$tag.log 2>&1
timeout 1500 ncu --set full --import-source on --clock-control none -k regex:"dopri_tpt|dopri_dde1" -s 1 -c 1 \
  -f -o gpurun_out/prof_${tag}_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline \
  > gpurun_out/ncu_full_$tag.log 2>&1
