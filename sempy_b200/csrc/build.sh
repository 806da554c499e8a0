#!/bin/bash
# Build libsemb200.so (the C-ABI library) in-tree for sm_100a.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libsemb200.so"
NVCC="${NVCC:-nvcc}"
"$NVCC" -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
  -Xcompiler -fPIC -shared -o "$OUT" "$HERE/semb200.cu" -ldl "$@"
echo "built $OUT"
