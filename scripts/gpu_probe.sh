#!/bin/bash
# ad-hoc probes: STFT kernel ncu capture, write-only bandwidth ceiling
TAG=${1:-probe}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python - > $OUT/bw.txt 2>&1 <<'PY'
import torch, time
x = torch.empty(1<<31, dtype=torch.float32, device='cuda')   # 8 GiB
y = torch.empty_like(x)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.zero_()); print('fill 8GiB  ms', ms, 'GB/s', x.numel()*4/ms/1e6)
ms = t(lambda: y.copy_(x)); print('copy 8GiB  ms', ms, 'GB/s (r+w)', 2*x.numel()*4/ms/1e6)
ms = t(lambda: x.sum()); print('read 8GiB  ms', ms, 'GB/s', x.numel()*4/ms/1e6)
PY
cat $OUT/bw.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_cta_kernel -s 3 -c 1 -o $OUT/prof_stft \
    python scripts/one_kernel.py stft > $OUT/ncu_stft.log 2>&1; echo "ncu stft exit $?"
ls -la $OUT
