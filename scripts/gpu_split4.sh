#!/bin/bash
# split kernel: default + build variants, rfft 2^16 / c2c 2^15 lines only, plus the split parity tests per variant
TAG=${1:-r04i}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in default "$@"; do
    echo "== variant $v" | tee -a $OUT/summary.txt
    if [ $v != default ]; then export KOFFT_CUDA_LIB=$PWD/kofft_b200/lib/libkofft_cuda_$v.so; fi
    timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "split_kernel or config3" 2>&1 | tail -1 | tee -a $OUT/summary.txt
    timeout 600 python scripts/bench_split.py exact 2> $OUT/split_$v.err | grep -E "rfft_65536|c2c_32768|c2c_16384" | grep split32 | tee -a $OUT/summary.txt
done
