#!/bin/bash
# GPU session for the split kernel: parity subset, timings against the older paths, one ncu capture.
TAG=${1:-r03a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
echo "== pytest split / large" | tee -a $OUT/summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "split_kernel or large or config3" > $OUT/pytest_split.log 2>&1
echo "pytest exit $?" | tee -a $OUT/summary.txt
tail -15 $OUT/pytest_split.log | tee -a $OUT/summary.txt
echo "== timings" | tee -a $OUT/summary.txt
timeout 600 python scripts/bench_split.py both > $OUT/split.jsonl 2> $OUT/split.err; echo "bench_split exit $?" | tee -a $OUT/summary.txt
cat $OUT/split.jsonl | tee -a $OUT/summary.txt
tail -5 $OUT/split.err | tee -a $OUT/summary.txt
echo "== ncu split" | tee -a $OUT/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:split32 -s 3 -c 1 -o $OUT/prof_split \
    python scripts/one_kernel.py split > $OUT/ncu_split.log 2>&1; echo "ncu exit $?" | tee -a $OUT/summary.txt
python scripts/summarize_ncu.py $TAG prof_split rfft_split >> $OUT/summary.txt 2>&1
mkdir -p $OUT/profiles; cp profiles/${TAG}_* $OUT/profiles/ 2>/dev/null
ls -la $OUT | tee -a $OUT/summary.txt
