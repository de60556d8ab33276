"""Dense C2C rows of 8192 / 16384 points: the wide single-CTA kernel (fft_wide.cuh) against the other paths.
   KOFFT_WIDE_MASK=0x6000 (both) | 0x2000 (8192 only, default) | 0 (off: split kernel / one-CTA kernel)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kofft_b200  # noqa: E402
from scripts.bench_kernels import PEAK, timeit  # noqa: E402

fft = kofft_b200.CudaFftImpl(device=0, exact=(os.environ.get("MODE", "exact") == "exact"))
g = torch.Generator(device="cuda").manual_seed(0)
for n in (8192, 16384):
    rows = 2 ** 28 // n
    x = torch.view_as_complex(torch.rand((rows, n, 2), generator=g, device="cuda") * 2 - 1).contiguous()
    y = torch.empty_like(x)
    for inverse in (False, True):
        ms, best = timeit(lambda: fft.fft_batch(x, out=y, inverse=inverse), 8, 2)
        print(json.dumps({"what": f"c2c_{n}x{rows}", "inverse": inverse, "wide_mask": os.environ.get("KOFFT_WIDE_MASK", "default"),
                          "mode": os.environ.get("MODE", "exact"), "ms_median": round(ms, 4), "ms_best": round(best, 4),
                          "frac_of_measured_peak": round(2 * x.numel() * 8 / ms / 1e6 / PEAK, 4)}), flush=True)
    del x, y
