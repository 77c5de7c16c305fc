#!/usr/bin/env bash
# TEST INFRASTRUCTURE: the reference's own example model, /root/reference/inst/examples/seir.c,
# compiled as a dde_b200 model library with nothing but the DDE_MODEL_FN decoration injected
# (tools/decorate_model.py; no reference source is copied into the repo: the decorated file
# lives in build/ only).  Output: tests/models/_ref/libdde_ref_example.so (git-ignored;
# travels to the GPU box like oracle/_ref).  Needs the reference tree.
set -euo pipefail
cd "$(dirname "$0")"
REF=${REF:-/root/reference}
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
test -f "$REF/inst/examples/seir.c" || { echo "no reference tree at $REF: keeping prebuilt tests/models/_ref"; exit 0; }
mkdir -p ../../build _ref
python ../../tools/decorate_model.py "$REF/inst/examples/seir.c" ../../build/ref_example_seir.c
cat > ../../build/ref_example.cu <<'EOT'
#include <dde_b200/dde_model.cuh>
#include "ref_example_seir.c"
DDE_REGISTER_DOPRI_OUTPUT(seir_reference_example, seir, 4, -1, seir_output, 1)
EOT
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false \
  -Xcompiler -fPIC,-mfma,-ffp-contract=off,-Wno-unused-function -I../../include \
  -shared -o _ref/libdde_ref_example.so ../../build/ref_example.cu -lcudart 2> ../../build/nvcc_ref_example.log \
  || { cat ../../build/nvcc_ref_example.log >&2; exit 1; }
rm -f ../../build/ref_example_seir.c ../../build/ref_example.cu
echo "built tests/models/_ref/libdde_ref_example.so from $REF/inst/examples/seir.c"
