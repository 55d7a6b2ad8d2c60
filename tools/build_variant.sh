#!/bin/bash
# build_variant.sh TAG "-DRAMCU_SPLIT_MASK=3 ..." : tools/variants/libramcu_TAG.so with kernels_b40 rebuilt under the
# given macros (development aid for A/B runs on the GPU: RAMCU_LIB=tools/variants/libramcu_TAG.so python tools/run_one.py)
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/variants
C=repeatafterme_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 128 $2 -c $C/kernels_b40.cu -o tools/variants/kernels_b40_$1.o
nvcc -gencode arch=compute_100a,code=sm_100a --shared -o tools/variants/libramcu_$1.so $C/ramcu.o $C/kernels_b14.o $C/kernels_b20.o \
  tools/variants/kernels_b40_$1.o $C/kernels_b60.o $C/kernels_b80.o $C/kernels_b100.o $C/kernels_generic.o
rm tools/variants/kernels_b40_$1.o
