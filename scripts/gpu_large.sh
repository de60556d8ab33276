#!/bin/bash
# GPU-box session for the N > 16384 paths: parity and timing of the three implementations, ncu capture
# of the pipelined kernel.
# Usage (under gpurun): bash scripts/gpu_large.sh [tag]
TAG=${1:-large}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
echo "== pytest large" | tee -a $OUT/summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "large or config3" > $OUT/pytest_large.log 2>&1
echo "pytest exit $?" | tee -a $OUT/summary.txt
tail -15 $OUT/pytest_large.log | tee -a $OUT/summary.txt
echo "== sweep" | tee -a $OUT/summary.txt
KOFFT_CUDA_VERBOSE=1 timeout 600 python scripts/bench_kernels.py rfft large > $OUT/large_kernels.jsonl 2> $OUT/sweep.err
echo "sweep exit $?" | tee -a $OUT/summary.txt
cat $OUT/large_kernels.jsonl | tee -a $OUT/summary.txt
tail -5 $OUT/sweep.err | tee -a $OUT/summary.txt
echo "== ncu full (pipelined rfft)" | tee -a $OUT/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:large_pipe -s 2 -c 1 -o $OUT/prof_rfft_pipe \
    python scripts/one_kernel.py rfft > $OUT/ncu_rfft_pipe.log 2>&1; echo "ncu exit $?" | tee -a $OUT/summary.txt
tail -3 $OUT/ncu_rfft_pipe.log | tee -a $OUT/summary.txt
ls -la $OUT | tee -a $OUT/summary.txt
