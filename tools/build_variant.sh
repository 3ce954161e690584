#!/bin/sh
# Builds one tile-geometry variant of the library for on-GPU sweeps:
#   tools/build_variant.sh TAG "-DXA_DEC_TBQ=256 -DXA_DEC_NT=128 ..."
# -> build/variants/TAG/libbjxa_b200.so   (select with BJXA_LIB=...)
set -e
tag=$1; defs=$2
d=build/variants/$tag
mkdir -p $d
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $defs \
    -c -o $d/xa_kernels.o bjxa_b200/csrc/xa_kernels.cu
gcc -std=c99 -O2 -fPIC -c -o $d/bjxa_host.o bjxa_b200/csrc/bjxa_host.c
gcc -std=c99 -O2 -fPIC -c -o $d/bjxa_corpus.o bjxa_b200/csrc/bjxa_corpus.c
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $d/libbjxa_b200.so $d/xa_kernels.o $d/bjxa_host.o $d/bjxa_corpus.o \
    -Xlinker --version-script=bjxa_b200/csrc/libbjxa.map -cudart static -lpthread -ldl -lrt
echo built $d
