"""Fused STFT magnitudes (config 4 shape on 16 channels), EXACT.   KOFFT_CUDA_LIB=<variant> python scripts/bench_stft_mag.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kofft_b200  # noqa: E402
from kofft_b200 import spectrogram as SP, window as W  # noqa: E402
from scripts.bench_kernels import timeit  # noqa: E402

ch, length, hop, win = 16, 28_800_000, 512, 2048
nframes = -(-length // hop)
g = torch.Generator(device="cuda").manual_seed(4)
sig = (torch.rand((ch, length), generator=g, device="cuda") * 2 - 1).contiguous()
w = torch.from_numpy(W.hann(win)).cuda()
fft = kofft_b200.CudaFftImpl(device=0, exact=True)
ms, best = timeit(lambda: SP.stft_magnitudes_batch(fft, sig, w, hop, nframes), 6, 2)
m, mx = SP.stft_magnitudes_batch(fft, sig, w, hop, nframes)
print(json.dumps({"lib": os.path.basename(os.environ.get("KOFFT_CUDA_LIB", "default")), "ms_median": round(ms, 4), "ms_best": round(best, 4),
                  "checksum": float(m[:, ::997].double().sum().item()), "max": float(mx.max().item())}), flush=True)
