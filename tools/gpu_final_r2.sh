#!/usr/bin/env bash
# usage: tools/gpu_final_r2.sh tag : everything profiles/ is built from on one GPU
tag=$1
tools/gpu_final_profile.sh $tag
python tools/bench_configs.py > gpurun_out/secondary_$tag.json 2> gpurun_out/secondary_$tag.err
tail -2 gpurun_out/secondary_$tag.err
