#!/bin/bash
# GPU-box session for the f64 twin: parity tests, per-size timing, ncu capture of the N = 4096 kernel.
TAG=${1:-f64}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest f64" | tee -a $OUT/summary.txt
timeout 900 python -m pytest tests/test_gpu_f64.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest_f64.log 2>&1
echo "pytest exit $?" | tee -a $OUT/summary.txt
tail -15 $OUT/pytest_f64.log | tee -a $OUT/summary.txt
echo "== timing" | tee -a $OUT/summary.txt
timeout 600 python scripts/bench_kernels.py f64 > $OUT/f64_kernels.jsonl 2> $OUT/f64.err
cat $OUT/f64_kernels.jsonl | tee -a $OUT/summary.txt
tail -3 $OUT/f64.err | tee -a $OUT/summary.txt
echo "== ncu full" | tee -a $OUT/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_f64_kernel -s 2 -c 1 -o $OUT/prof_f64 \
    python scripts/one_kernel.py f64 > $OUT/ncu_f64.log 2>&1; echo "ncu exit $?" | tee -a $OUT/summary.txt
tail -3 $OUT/ncu_f64.log | tee -a $OUT/summary.txt
