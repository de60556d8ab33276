#!/bin/bash
# split kernel: default + build variants (rfft / c2c 2^15 only), exact mode
TAG=${1:-r03h}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "split_kernel or large or config3" > $OUT/pytest_split.log 2>&1
echo "pytest exit $?" | tee -a $OUT/summary.txt
tail -3 $OUT/pytest_split.log | tee -a $OUT/summary.txt
timeout 600 python scripts/bench_split.py exact > $OUT/split.jsonl 2> $OUT/split.err; echo "bench_split exit $?" | tee -a $OUT/summary.txt
cat $OUT/split.jsonl | tee -a $OUT/summary.txt
for v in "$@"; do
    echo "== variant $v" | tee -a $OUT/summary.txt
    KOFFT_CUDA_LIB=$PWD/kofft_b200/lib/libkofft_cuda_$v.so timeout 600 python scripts/bench_split.py exact 2> $OUT/split_$v.err | grep -E "rfft_65536|c2c_32768" | grep split32 | tee -a $OUT/summary.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:split32 -s 3 -c 1 -o $OUT/prof_split \
    python scripts/one_kernel.py split > $OUT/ncu_split.log 2>&1; echo "ncu exit $?" | tee -a $OUT/summary.txt
python scripts/summarize_ncu.py $TAG prof_split rfft_split >> $OUT/summary.txt 2>&1
mkdir -p $OUT/profiles; cp profiles/${TAG}_* $OUT/profiles/ 2>/dev/null
