#!/bin/bash
# Build a second library with other compile-time parameters, next to the default one:
#   tools/build_variant.sh bn256 -DLEIDA_TC_BN=256     -> leida-python_b200/pyleida/_lib/libleida_b200_bn256.so
# select it at run time with LEIDA_B200_LIB=<path>.
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/leida-python_b200/csrc
OUT=/tmp/leida_variant_$NAME
mkdir -p $OUT
for f in capi signal clean eigdense kmeans kmeans_tc scores dynamics stats; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I $ROOT/include -I $SRC "$@" -c $SRC/$f.cu -o $OUT/$f.o &
done
wait
nvcc -shared -o $ROOT/leida-python_b200/pyleida/_lib/libleida_b200_$NAME.so $OUT/*.o
echo built $ROOT/leida-python_b200/pyleida/_lib/libleida_b200_$NAME.so
