#!/usr/bin/env bash
# Build libdde_b200.so (CUDA kernels + C-ABI) in-tree for sm_100a.
# -fmad=false: no FMA contraction -- bit-parity with the reference depends on it.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false
       -Xcompiler -fPIC,-mfma,-ffp-contract=off,-Wall,-Wno-unused-function
       -Iinclude -Idde_b200/csrc ${DDE_NVCC_EXTRA:-})
mkdir -p build
"$NVCC" "${FLAGS[@]}" -c dde_b200/csrc/capi.cu -o build/capi.o &
"$NVCC" "${FLAGS[@]}" -Xptxas -v -c dde_b200/csrc/models_builtin.cu -o build/models_builtin.o 2> build/ptxas_models.log &
wait
"$NVCC" -shared -o dde_b200/libdde_b200.so build/capi.o build/models_builtin.o -lcudart
echo "built dde_b200/libdde_b200.so"
