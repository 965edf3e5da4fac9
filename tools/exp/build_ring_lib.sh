#!/bin/bash
# Builds the EXPERIMENTAL library tools/exp/bin/libndconv_cuda_ring.so (-DNDCONV_EXP_RING): workspace slots taken modulo
# NDCONV_EXP_RING_TILES, and NDCONV_EXP_FLAGS (1 = col_pass_tma_kres without workspace loads, 2 = without stores).  Its results are wrong by
# construction; it exists for the timing experiments of DESIGN.md section 9 / 3.3 (tools/gpu_r3h.sh, gpu_r3l.sh, gpu_r3n.sh, gpu_r3v.sh copy it
# over the product library on the scratch GPU box only).  Never loaded by the product or the tests.
set -e
cd "$(dirname "$0")/../../ndarray-conv_b200/csrc"
mkdir -p ../../tools/exp/bin
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=true -Xcompiler -fPIC,-O2,-Wall,-Wno-unknown-pragmas -shared \
     --expt-relaxed-constexpr -DNDCONV_EXP_RING -o ../../tools/exp/bin/libndconv_cuda_ring.so api.cu host_logic.cpp
ls -la ../../tools/exp/bin/libndconv_cuda_ring.so
