"""Fused ISTFT (config 4 shape on 16 channels), EXACT and FAST, timed; round trip against the signal.
   KOFFT_CUDA_LIB=<variant> python scripts/bench_istft.py"""
import json
import os
import sys

import numpy as np
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
ref = None
for exact in (True, False):
    fft = kofft_b200.CudaFftImpl(device=0, exact=exact)
    frames = S.stft_batch(fft, sig, w, hop, nframes)
    out = torch.zeros((ch, length), device="cuda")
    ms, best = timeit(lambda: S.istft_batch(fft, frames, w, hop, out), 6, 2)
    out.zero_()
    S.istft_batch(fft, frames, w, hop, out)
    inner = slice(win, length - win)
    rt = float(((out[:, inner] - sig[:, inner]).double().norm() / sig[:, inner].double().norm()).item())
    algo = 4 * ch * length + 8 * ch * nframes * win
    h = float(out.double().sum().item())
    print(json.dumps({"lib": os.path.basename(os.environ.get("KOFFT_CUDA_LIB", "default")), "mode": "exact" if exact else "fast",
                      "ms_median": round(ms, 4), "ms_best": round(best, 4), "frac_of_measured_peak": round(algo / ms / 1e6 / PEAK, 4),
                      "roundtrip_rel_l2": rt, "checksum": h}), flush=True)
    del frames, out
