#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY: the device code of include/dde_b200 compiled by g++ and run
# one lane at a time (tests/emu/cuda_emu.h).  Output: tests/emu/libdde_emu.so.
set -euo pipefail
cd "$(dirname "$0")"
g++ -O2 -g -std=c++17 -fPIC -mfma -ffp-contract=off -Wall -Wno-unused-function \
    -Wno-unused-variable -Wno-unused-but-set-variable -Wno-maybe-uninitialized -Wno-unknown-pragmas -Wno-sign-compare \
    ${DDE_EMU_EXTRA:-} -I. -I../../include -I../../include/dde_b200 -I../../dde_b200/models \
    -fvisibility=hidden -Wl,-Bsymbolic -shared -o ${DDE_EMU_OUT:-libdde_emu.so} emu_driver.cpp -lm
echo "built tests/emu/${DDE_EMU_OUT:-libdde_emu.so}"
