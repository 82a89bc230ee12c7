#!/bin/bash
# Kernel-experiment build of the library: tools/build_variant.sh <name> <file.cu> <extra nvcc flags...>
# recompiles ONE source with the extra flags and links it with the other objects of the regular build into
# variants/<name>/libglomar_b200.so (git-ignored); select it with GLOMAR_B200_LIB=variants/<name>/libglomar_b200.so.
set -e
NAME=$1; SRC=$2; shift 2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CS=$ROOT/glomargridding_b200/csrc
OUT=$ROOT/variants/$NAME
mkdir -p $OUT
BASE=$(basename $SRC .cu)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -O2 \
     --expt-relaxed-constexpr "$@" -c $CS/$SRC -o $OUT/$BASE.o
OBJS=$(ls $CS/build/*.o | grep -v "/$BASE.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libglomar_b200.so $OBJS $OUT/$BASE.o -lcudart -ldl
rm -f $OUT/$BASE.o
echo built $OUT/libglomar_b200.so
