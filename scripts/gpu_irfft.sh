#!/bin/bash
TAG=${1:-r04l}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in default "$@"; do
    if [ $v != default ]; then export KOFFT_CUDA_LIB=$PWD/kofft_b200/lib/libkofft_cuda_$v.so; fi
    timeout 300 python scripts/bench_irfft_split.py 2> $OUT/irfft_$v.err | tee -a $OUT/summary.txt
done
unset KOFFT_CUDA_LIB
[ -n "$NO_NCU" ] || timeout 600 ncu --set full --clock-control none --import-source on -k regex:split32 -s 3 -c 1 -o $OUT/prof_isplit \
    python scripts/one_kernel.py isplit > $OUT/ncu_isplit.log 2>&1; echo "ncu exit $?" | tee -a $OUT/summary.txt
