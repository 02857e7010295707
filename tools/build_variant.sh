#!/bin/bash
# usage: tools/build_variant.sh NAME [-DFLAG=...]...  -> variants/libgradus_b200_NAME.so (tuning aid; GB200_LIB selects it)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
T=$(mktemp -d)
cd "$ROOT/gradus.jl_b200/csrc"
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c gb200_trace.cu -o $T/trace.o &
nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c gb200_api.cu -o $T/api.o &
nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c gb200_dual.cu -o $T/dual.o &
wait
mkdir -p "$ROOT/variants"
nvcc $ARCH -shared -o "$ROOT/variants/libgradus_b200_$NAME.so" $T/trace.o $T/api.o $T/dual.o -cudart static
rm -rf $T
echo built variants/libgradus_b200_$NAME.so
