#!/bin/bash
# quick A/B for the split kernel: parity subset, rfft/irfft/c2c timings, persisting-L2 variant, one ncu capture
TAG=${1:-r03g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "split_kernel or large or config3" > $OUT/pytest_split.log 2>&1
echo "pytest exit $?" | tee -a $OUT/summary.txt
tail -5 $OUT/pytest_split.log | tee -a $OUT/summary.txt
timeout 600 python scripts/bench_split.py exit > $OUT/split.jsonl 2> $OUT/split.err; echo "bench_split exit $?" | tee -a $OUT/summary.txt
cat $OUT/split.jsonl | tee -a $OUT/summary.txt
echo "== persisting L2" | tee -a $OUT/summary.txt
KOFFT_L2_PERSIST=1 KOFFT_CUDA_VERBOSE=1 timeout 600 python scripts/bench_split.py exact > $OUT/split_persist.jsonl 2> $OUT/split_persist.err; echo "bench_split exit $?" | tee -a $OUT/summary.txt
head -3 $OUT/split_persist.jsonl | tee -a $OUT/summary.txt; grep persisting $OUT/split_persist.err | head -2 | tee -a $OUT/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:split32 -s 3 -c 1 -o $OUT/prof_split \
    python scripts/one_kernel.py split > $OUT/ncu_split.log 2>&1; echo "ncu exit $?" | tee -a $OUT/summary.txt
KOFFT_L2_PERSIST=1 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:split32 -s 3 -c 1 \
    python scripts/one_kernel.py split 2>&1 | grep -E "dram__|gpu__time" | tee -a $OUT/summary.txt
python scripts/summarize_ncu.py $TAG prof_split rfft_split >> $OUT/summary.txt 2>&1
mkdir -p $OUT/profiles; cp profiles/${TAG}_* $OUT/profiles/ 2>/dev/null
