#!/usr/bin/env bash
# TEST INFRASTRUCTURE: the reference's fixture models as an out-of-tree model library,
# tests/models/libdde_test_models.so, built against include/ only (no source of the product
# library) -- the recipe INTEGRATION.md gives for a user's model library.
# -fmad=false: no FMA contraction -- bit-parity with the reference depends on it.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false
       -Xcompiler -fPIC,-mfma,-ffp-contract=off,-Wall,-Wno-unused-function -I../../include -I.)
mkdir -p ../../build
"$NVCC" "${FLAGS[@]}" -c test_models.cu -o ../../build/test_models.o 2> ../../build/nvcc_test_models.log &
"$NVCC" "${FLAGS[@]}" -c test_models_coop.cu -o ../../build/test_models_coop.o 2> ../../build/nvcc_test_models_coop.log &
wait
grep -i -B2 -A6 "error" ../../build/nvcc_test_models.log ../../build/nvcc_test_models_coop.log || true
test -f ../../build/test_models.o -a -f ../../build/test_models_coop.o || { echo "build failed" >&2; exit 1; }
# dde_register_model is resolved when the library is loaded next to libdde_b200.so
"$NVCC" -shared -o libdde_test_models.so ../../build/test_models.o ../../build/test_models_coop.o -lcudart
rm -f ../../build/test_models.o ../../build/test_models_coop.o
echo "built tests/models/libdde_test_models.so"
