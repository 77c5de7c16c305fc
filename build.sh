#!/usr/bin/env bash
# Build libdde_b200.so (CUDA kernels + C-ABI) in-tree for sm_100a.
# -fmad=false: no FMA contraction -- bit-parity with the reference depends on it.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false
       -Xcompiler -fPIC,-mfma,-ffp-contract=off,-Wall,-Wno-unused-function
       -Iinclude -Iinclude/dde_b200 -Idde_b200/models ${DDE_NVCC_EXTRA:-})
# DDE_OUT / DDE_TAG: build an experimental variant next to the product library
OUT=${DDE_OUT:-dde_b200/libdde_b200.so}
TAG=${DDE_TAG:-}
mkdir -p build
rm -f build/capi$TAG.o build/models_builtin$TAG.o build/models_coop$TAG.o
"$NVCC" "${FLAGS[@]}" -c dde_b200/csrc/capi.cu -o build/capi$TAG.o 2> build/nvcc_capi$TAG.log &
"$NVCC" "${FLAGS[@]}" -Xptxas -v -c dde_b200/csrc/models_builtin.cu -o build/models_builtin$TAG.o 2> build/ptxas_models$TAG.log &
"$NVCC" "${FLAGS[@]}" -Xptxas -v -c dde_b200/csrc/models_coop.cu -o build/models_coop$TAG.o 2> build/ptxas_coop$TAG.log &
wait
grep -v "offline compilation" build/nvcc_capi$TAG.log | grep -i -B2 -A6 "error\|warning" || true
grep -i -B2 -A6 "error" build/ptxas_models$TAG.log || true
grep -i -B2 -A6 "error" build/ptxas_coop$TAG.log || true
test -f build/capi$TAG.o -a -f build/models_builtin$TAG.o -a -f build/models_coop$TAG.o || { echo "build failed" >&2; exit 1; }
"$NVCC" -shared -o "$OUT" build/capi$TAG.o build/models_builtin$TAG.o build/models_coop$TAG.o -lcudart
echo "built $OUT"
# variant objects are not worth keeping (they would travel to the GPU box)
if [ -n "$TAG" ]; then rm -f build/capi$TAG.o build/models_builtin$TAG.o build/models_coop$TAG.o; fi
