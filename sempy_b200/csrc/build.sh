#!/bin/bash
# Build libsemb200.so (the C-ABI library) in-tree for sm_100a: one object per translation unit,
# compiled in parallel, linked into one shared library.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libsemb200.so"
OBJ="$HERE/_obj"
NVCC="${NVCC:-nvcc}"
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC"
mkdir -p "$OBJ"
pids=()
for src in semb200 k1_ax8 k1_axl_lo k1_axl_hi k1_axl_small k1_generic; do
  "$NVCC" $FLAGS "$@" -c "$HERE/$src.cu" -o "$OBJ/$src.o" &
  pids+=($!)
done
for pid in "${pids[@]}"; do wait "$pid"; done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" "$OBJ"/semb200.o "$OBJ"/k1_ax8.o "$OBJ"/k1_axl_lo.o \
  "$OBJ"/k1_axl_hi.o "$OBJ"/k1_axl_small.o "$OBJ"/k1_generic.o -ldl -lpthread
echo "built $OUT"
