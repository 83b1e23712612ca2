#!/bin/bash
# Experiment build: tools/build_variant.sh <suffix> <file.cu> <nvcc flags...>  ->  libhpose<suffix>.so with ONE translation unit
# recompiled with the extra flags (the other objects are taken from the default build).  Select it with HPOSE_LIB_SUFFIX=<suffix>.
set -e
SUF=$1; SRC=$2; shift 2
PKG=$(dirname "$0")/../head-pose-estimation-model_b200
mkdir -p $PKG/csrc/build$SUF
cp -p $PKG/csrc/build/*.o $PKG/csrc/build$SUF/
(cd $PKG/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c $SRC -o build$SUF/${SRC%.cu}.o &&
 nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../libhpose$SUF.so build$SUF/*.o -ldl)
echo built libhpose$SUF.so
