#!/bin/bash
# Development aid: build liblfcuda.so with extra -D flags into ab/<name>.so for A/B runs on the GPU box
# (select one with LF_LFCUDA_SO=ab/<name>.so).  Usage: tools/build_variant.sh <name> [-DLF_...=...]...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
defs=""; for a in "$@"; do case $a in -D*) defs="$defs $a";; esac; done   # only the -D flags go to g++
mkdir -p ab build/ab
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=${LF_FMAD:-false} -Xcompiler -fPIC -std=c++17 \
     -Iinclude -Ilavaframe_b200/csrc "$@" -c lavaframe_b200/csrc/lf_kernels.cu -o build/ab/$name.kernels.o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -Xcompiler -fPIC -std=c++17 \
     -Iinclude -Ilavaframe_b200/csrc -c lavaframe_b200/csrc/lf_tlas.cu -o build/ab/$name.tlas.o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -Xcompiler -fPIC -std=c++17 \
     -Iinclude -Ilavaframe_b200/csrc -c lavaframe_b200/csrc/lf_blas.cu -o build/ab/$name.blas.o
g++ -O2 -fPIC -std=c++17 -ffp-contract=off -Iinclude -Ilavaframe_b200/csrc -I/usr/local/cuda/include $defs -c lavaframe_b200/csrc/lfcuda.cpp -o build/ab/$name.lfcuda.o
g++ -O2 -fPIC -std=c++17 -ffp-contract=off -Iinclude -Ilavaframe_b200/csrc -I/usr/local/cuda/include $defs -c lavaframe_b200/csrc/lf_repack.cpp -o build/ab/$name.repack.o
g++ -O2 -fPIC -std=c++17 -ffp-contract=off -Iinclude -Ilavaframe_b200/csrc -I/usr/local/cuda/include $defs -c lavaframe_b200/csrc/lfcuda_group.cpp -o build/ab/$name.group.o
g++ -shared -o ab/$name.so build/ab/$name.kernels.o build/ab/$name.lfcuda.o build/ab/$name.repack.o build/ab/$name.group.o build/ab/$name.tlas.o build/ab/$name.blas.o -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
echo "built ab/$name.so"
