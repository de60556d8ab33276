"""Fused STFT (config 4 shape on 16 channels), EXACT and FAST, timed.   KOFFT_CUDA_LIB=<variant> python scripts/bench_stft.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kofft_b200  # noqa: E402
from kofft_b200 import stft as S, window as W  # noqa: E402
from scripts.bench_kernels import PEAK, timeit  # noqa: E402

ch, length, hop, win = 16, 28_800_000, 512, 2048
nframes = -(-length // hop)
g = torch.Generator(device="cuda").manual_seed(4)
sig = (torch.rand((ch, length), generator=g, device="cuda") * 2 - 1).contiguous()
w = torch.from_numpy(W.hann(win)).cuda()
frames = torch.empty((ch, nframes, win), dtype=torch.complex64, device="cuda")
for exact in (True, False):
    fft = kofft_b200.CudaFftImpl(device=0, exact=exact)
    ms, best = timeit(lambda: S.stft_batch(fft, sig, w, hop, nframes, out=frames), 6, 2)
    algo = 4 * ch * length + 8 * ch * nframes * win
    chk = float(torch.view_as_real(frames[:, ::997]).double().sum().item())
    print(json.dumps({"lib": os.path.basename(os.environ.get("KOFFT_CUDA_LIB", "default")), "mode": "exact" if exact else "fast",
                      "ms_median": round(ms, 4), "ms_best": round(best, 4), "frac_of_measured_peak": round(algo / ms / 1e6 / PEAK, 4),
                      "checksum": chk}), flush=True)
