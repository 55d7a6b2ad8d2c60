#!/bin/bash
# build_variant.sh TAG "-DMACRO=... ..." : tools/variants/libramcu_TAG.so with every kernel unit rebuilt under the given
# macros (development aid for A/B runs on the GPU: RAMCU_LIB=tools/variants/libramcu_TAG.so python tools/run_one.py)
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/variants
C=repeatafterme_b200/csrc
UNITS="kernels_b14 kernels_b20 kernels_b40 kernels_b60 kernels_b80 kernels_b100 kernels_generic"
for u in $UNITS; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 128 $2 -c $C/$u.cu -o tools/variants/${u}_$1.o &
done
wait
OBJS=""; for u in $UNITS; do OBJS="$OBJS tools/variants/${u}_$1.o"; done
nvcc -gencode arch=compute_100a,code=sm_100a --shared -o tools/variants/libramcu_$1.so $C/ramcu.o $OBJS
rm -f $OBJS
