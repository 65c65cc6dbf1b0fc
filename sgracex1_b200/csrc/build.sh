#!/bin/bash
# Builds sgracex1_b200/libsgrace_b200.so for sm_100a (cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libsgrace_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
HOSTCXX=/usr/bin/g++
[ -x "$HOSTCXX" ] || HOSTCXX=g++
"$NVCC" -ccbin "$HOSTCXX" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -Xcompiler -fPIC -Xcompiler -Wall -shared ${SGRACE_NVCC_EXTRA} \
    -o "$OUT" "$HERE/sgrace_abi.cu" -lcuda
echo "built $OUT"
