#!/usr/bin/env bash
# usage: tools/gpu_quick.sh tag : smoke check (bit-compare with the oracle) + short bench of the in-tree library
tag=$1
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/quick_$tag.log 2>&1 || echo "SMOKE FAILED" >> gpurun_out/quick_$tag.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline >> gpurun_out/quick_$tag.log 2>&1
grep -o '"value": [0-9.]*\|"e2e": {"value": [0-9.]*\|smoke ok\|SMOKE FAILED\|Error.*' gpurun_out/quick_$tag.log
