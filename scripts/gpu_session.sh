#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, ncu launch list + full capture of the headline kernel.
# Usage (from the repo root, under gpurun): bash scripts/gpu_session.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
echo "== smoke" | tee -a $OUT/summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/summary.txt
tail -3 $OUT/smoke.log | tee -a $OUT/summary.txt
echo "== pytest -m gpu" | tee -a $OUT/summary.txt
timeout 2400 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/summary.txt
tail -40 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
echo "== bench" | tee -a $OUT/summary.txt
KOFFT_CUDA_VERBOSE=1 timeout 900 python bench.py --steps 50 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?" | tee -a $OUT/summary.txt
cat $OUT/bench.json | tee -a $OUT/summary.txt
tail -5 $OUT/bench.err | tee -a $OUT/summary.txt
echo "== bench reference arm" | tee -a $OUT/summary.txt
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > $OUT/bench_ref.json 2>> $OUT/bench.err
cat $OUT/bench_ref.json | tee -a $OUT/summary.txt
echo "== ncu launch list" | tee -a $OUT/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 > $OUT/ncu_launches.log 2>&1; echo "ncu list exit $?" | tee -a $OUT/summary.txt
echo "== ncu full" | tee -a $OUT/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_cta_kernel -s 3 -c 2 -o $OUT/prof_c2c \
    python scripts/one_kernel.py c2c > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?" | tee -a $OUT/summary.txt
ls -la $OUT | tee -a $OUT/summary.txt
echo "== ncu stft / rfft" | tee -a $OUT/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_cta_kernel -s 3 -c 1 -o $OUT/prof_stft \
    python scripts/one_kernel.py stft > $OUT/ncu_stft.log 2>&1; echo "ncu stft exit $?" | tee -a $OUT/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:split32 -s 3 -c 1 -o $OUT/prof_rfft_split \
    python scripts/one_kernel.py split > $OUT/ncu_rfft_split.log 2>&1; echo "ncu rfft (split kernel, default) exit $?" | tee -a $OUT/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:split32 -s 3 -c 1 -o $OUT/prof_irfft_split \
    python scripts/one_kernel.py isplit > $OUT/ncu_irfft_split.log 2>&1; echo "ncu irfft (split kernel) exit $?" | tee -a $OUT/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_wide -s 3 -c 1 -o $OUT/prof_wide \
    python scripts/one_kernel.py wide > $OUT/ncu_wide.log 2>&1; echo "ncu c2c 8192 (wide kernel) exit $?" | tee -a $OUT/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:istft_fused -s 2 -c 1 -o $OUT/prof_istft \
    python scripts/one_kernel.py istft > $OUT/ncu_istft.log 2>&1; echo "ncu istft exit $?" | tee -a $OUT/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_f64_kernel -s 2 -c 1 -o $OUT/prof_f64 \
    python scripts/one_kernel.py f64 > $OUT/ncu_f64.log 2>&1; echo "ncu f64 exit $?" | tee -a $OUT/summary.txt
echo "== per-shape kernel timings" | tee -a $OUT/summary.txt
timeout 900 python scripts/bench_kernels.py > $OUT/kernels.jsonl 2> $OUT/kernels.err; echo "kernels exit $?" | tee -a $OUT/summary.txt
cat $OUT/kernels.jsonl | tee -a $OUT/summary.txt
# summaries are produced on the box (the raw .ncu-rep files together exceed what gpurun copies back);
# only the headline capture travels home as a report
echo "== summaries" | tee -a $OUT/summary.txt
mkdir -p $OUT/profiles
python scripts/summarize_ncu.py $TAG prof_c2c c2c --headline >> $OUT/summary.txt 2>&1
for pair in "prof_stft stft" "prof_rfft_split rfft_split" "prof_irfft_split irfft_split" "prof_wide c2c8192_wide" "prof_istft istft" "prof_f64 f64"; do
    set -- $pair
    [ -f $OUT/$1.ncu-rep ] && python scripts/summarize_ncu.py $TAG $1 $2 >> $OUT/summary.txt 2>&1
done
cp profiles/${TAG}_* profiles/headline_kernel_traffic.json $OUT/profiles/ 2>/dev/null
for r in prof_stft prof_rfft_split prof_irfft_split prof_wide prof_istft prof_f64; do rm -f $OUT/$r.ncu-rep; done
du -sh $OUT | tee -a $OUT/summary.txt
