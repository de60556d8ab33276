#!/bin/bash
# Builds tuning variants of libkofft_cuda.so side by side (kofft_b200/lib/libkofft_cuda_<name>.so);
# select one at run time with KOFFT_CUDA_LIB=<path>.   usage: build_variants.sh name "-DFOO=1 ..." [name flags ...]
set -e
cd "$(dirname "$0")/../kofft_b200/csrc"
while [ $# -ge 2 ]; do
    name=$1; flags=$2; shift 2
    make -j16 EXTRA="$flags" OBJDIR=../../build/obj_$name LIBNAME=libkofft_cuda_$name.so > /dev/null
    echo "built kofft_b200/lib/libkofft_cuda_$name.so  ($flags)"
done
