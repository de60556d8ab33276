#!/bin/bash
# ad-hoc probes: FP32 scalar/packed issue rates (scripts/microbench/fp32x2.cu)
TAG=${1:-probe}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp32x2 scripts/microbench/fp32x2.cu && /tmp/fp32x2 > $OUT/fp32x2.txt 2>&1
cat $OUT/fp32x2.txt
