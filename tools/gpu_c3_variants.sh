#!/usr/bin/env bash
# usage: tools/gpu_c3_variants.sh lib1.so lib2.so ... : C3 Lorenz timing of each library variant
for lib in "$@"; do
  echo "== $lib"
  DDE_B200_LIB=$PWD/$lib python tools/run_c3.py 1048576
done
