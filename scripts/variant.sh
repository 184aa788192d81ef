#!/bin/bash
# Builds a VARIANT of the product library for kernel experiments: recompiles one translation unit of
# fnft_b200/csrc/cuda with extra -D flags and links it with the other objects of the main build.
#   scripts/variant.sh <name> <tu> [nvcc flags...]     e.g.  scripts/variant.sh upA k_tree_up -DFNFTB_UP_PREFETCH=1
# -> fnft_b200/lib/var/libfnft_b200_<name>.so (git-ignored; select it with FNFT_B200_LIB=<path>)
set -e
cd "$(dirname "$0")/.."
name=$1; tu=$2; shift 2
mkdir -p build/var fnft_b200/lib/var
make -j8 > /dev/null
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" \
     -c fnft_b200/csrc/cuda/$tu.cu -o build/var/${name}_$tu.o
objs=$(ls build/host_*.o build/cuda_*.o | grep -v "cuda_$tu.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o fnft_b200/lib/var/libfnft_b200_$name.so $objs build/var/${name}_$tu.o -lm
echo fnft_b200/lib/var/libfnft_b200_$name.so
