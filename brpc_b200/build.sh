#!/bin/sh
# Builds brpc_b200/libb2rpc.so (the C-ABI product library) for sm_100a, in-tree.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -Xcompiler -fPIC,-Wall -Xptxas -v -shared $B2_NVCC_DEFS \
    -o "${B2_OUT:-$HERE/libb2rpc.so}" "$HERE/csrc/b2_api.cu" 2>&1
