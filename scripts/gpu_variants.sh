#!/bin/bash
# GPU session: parity tests + kernel timings for the default library and tuning variants.
# usage (under gpurun): bash scripts/gpu_variants.sh <tag> "<which...>" [variant ...]
TAG=${1:-var}; WHICH=${2:-stft}; shift 2
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -5 $OUT/pytest_gpu.log
timeout 600 python scripts/bench_kernels.py $WHICH > $OUT/kernels_default.jsonl 2> $OUT/kernels_default.err
cat $OUT/kernels_default.jsonl; tail -3 $OUT/kernels_default.err
for v in "$@"; do
    KOFFT_CUDA_LIB=$PWD/kofft_b200/lib/libkofft_cuda_$v.so timeout 600 python scripts/bench_kernels.py $WHICH > $OUT/kernels_$v.jsonl 2> $OUT/kernels_$v.err
    cat $OUT/kernels_$v.jsonl; tail -3 $OUT/kernels_$v.err
done
