#!/bin/bash
# usage: tools/build_variant.sh NAME 'sed-expr' ['sed-expr' ...]   -> tools/variants/lib_NAME.so (kmeans_tc.cu patched)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p tools/variants
args=(); for e in "$@"; do args+=(-e "$e"); done
sed "${args[@]}" leida-python_b200/csrc/kmeans_tc.cu > /tmp/kmeans_tc_$name.cu
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v -I include -I leida-python_b200/csrc \
  -c /tmp/kmeans_tc_$name.cu -o /tmp/kmeans_tc_$name.o 2>&1 | grep -A2 "k_tc_assign" | tail -1
objs=$(ls leida-python_b200/csrc/build/*.o | grep -v kmeans_tc.o)
nvcc -shared -o tools/variants/lib_$name.so $objs /tmp/kmeans_tc_$name.o
