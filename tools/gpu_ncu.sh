#!/usr/bin/env bash
# usage: tools/gpu_ncu.sh tag [n_traj] [lib] : one ncu --set full capture of the dopri stepping kernel (second launch)
tag=$1; ntraj=${2:-1048576}; lib=${3:-}
mkdir -p gpurun_out
if [ -n "$lib" ]; then export DDE_B200_LIB=$PWD/$lib; fi
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"dopri_tpt|dopri_dde1" -s 1 -c 1 \
  -f -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 1 --no-cpu-baseline --n-traj $ntraj \
  > gpurun_out/ncu_$tag.log 2>&1
