#!/bin/bash
# Build libuivr.so for sm_100a.  -fmad=false is part of the arithmetic contract:
# nvcc must not fuse a*b+c on its own; every FMA in the source is an explicit fmaf().
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false \
    -Xcompiler -fPIC -shared ${UIVR_NVCC_EXTRA} -o ${UIVR_OUT:-libuivr.so} uivr_api.cu
