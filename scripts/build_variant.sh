#!/bin/bash
# Dev tool: build sempy_b200/libsemb200_<tag>.so with extra -D flags for the K1 translation units
# (the other objects are taken from the last regular build).  usage: build_variant.sh <tag> [-DKNOB=value ...]
set -e
HERE="$(cd "$(dirname "$0")/../sempy_b200/csrc" && pwd)"
TAG="$1"; shift
OBJ="$HERE/_obj/var_$TAG"
mkdir -p "$OBJ"
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC"
pids=()
for src in k1_ax8 k1_axl_lo k1_axl_hi k1_axl_small; do
  nvcc $FLAGS "$@" -c "$HERE/$src.cu" -o "$OBJ/$src.o" &
  pids+=($!)
done
for pid in "${pids[@]}"; do wait "$pid"; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$HERE/../libsemb200_$TAG.so" "$HERE/_obj/semb200.o" "$OBJ/k1_ax8.o" \
  "$OBJ/k1_axl_lo.o" "$OBJ/k1_axl_hi.o" "$OBJ/k1_axl_small.o" "$HERE/_obj/k1_generic.o" -ldl -lpthread
echo "built libsemb200_$TAG.so"
