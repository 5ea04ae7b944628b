#!/bin/bash
# Builds side-by-side variants of the library for A/B timing on the GPU box (dev/time_step.py with XPSI_B200_LIB):
#   bash dev/build_variants.sh tag "-DXB_S2_UNROLL=0" [tag2 "flags2" ...]
set -e
cd "$(dirname "$0")/../xpsi_b200/csrc"
make -s
mkdir -p ../../dev/variants
while [ $# -ge 2 ]; do
  TAG=$1; FLAGS=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $FLAGS -c integrate_azinv.cu -o /tmp/azinv_$TAG.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../dev/variants/lib_$TAG.so /tmp/azinv_$TAG.o integrate_general.o integrate_tinv.o embed.o tools.o marginal.o api.o -lcudart
  echo built dev/variants/lib_$TAG.so
done
