#!/bin/bash
# Kernel tuning: build an alternative libterrarium_b200.so with extra -D flags for the fast-math translation unit.
#   profiles/build_variant.sh NAME "-DTRM_EULER2_BLOCKS=5 ..."   ->  terrarium.jl_b200/csrc/variants/libtrm_NAME.so
# Select it at run time with TRM_LIB=<path> (terrarium.jl_b200/_lib.py). The variants directory is git-ignored.
set -e
cd "$(dirname "$0")/../terrarium.jl_b200/csrc"
mkdir -p variants
name=$1; flags=$2
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DTRM_MAX_BLOCK=128 -DTRM_MIN_BLOCKS=4 -Xcompiler -fPIC,-ffp-contract=off -fmad=true $flags -c kernels_fast.cu -o variants/kernels_fast_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libtrm_$name.so kernels_faithful.o variants/kernels_fast_$name.o terrarium_b200.o -cudart shared
echo variants/libtrm_$name.so
