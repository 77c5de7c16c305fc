#!/usr/bin/env bash
# usage: tools/gpu_variants.sh tag lib1.so lib2.so ... : smoke check (bit-compare with the
# oracle) + short bench of each library variant
tag=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  echo "== $lib" >> gpurun_out/variants_$tag.log
  DDE_B200_LIB=$PWD/$lib timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/variants_$tag.log 2>&1 || echo "SMOKE FAILED" >> gpurun_out/variants_$tag.log
  DDE_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline >> gpurun_out/variants_$tag.log 2>&1
done
grep -o '"value": [0-9.]*\|"e2e": {"value": [0-9.]*\|^== .*\|smoke ok\|SMOKE FAILED\|Error.*' gpurun_out/variants_$tag.log
