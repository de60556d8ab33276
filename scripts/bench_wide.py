"""Complex cores of 8192 / 16384 points: the wide single-CTA kernel (fft_wide.cuh) against the other paths, per kind.
   python scripts/bench_wide.py        one JSON line per (shape, kind, path)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kofft_b200  # noqa: E402
from scripts.bench_kernels import PEAK, timeit  # noqa: E402

mode = os.environ.get("MODE", "exact")
fft = kofft_b200.CudaFftImpl(device=0, exact=(mode == "exact"))
g = torch.Generator(device="cuda").manual_seed(0)


def rep(what, path, ms, best, nbytes):
    print(json.dumps({"what": what, "path": path, "mode": mode, "ms_median": round(ms, 4), "ms_best": round(best, 4),
                      "frac_of_measured_peak": round(nbytes / ms / 1e6 / PEAK, 4)}), flush=True)


for n in (8192, 16384):
    rows = 2 ** 28 // n
    x = torch.view_as_complex(torch.rand((rows, n, 2), generator=g, device="cuda") * 2 - 1).contiguous()
    y = torch.empty_like(x)
    xr = (torch.rand((rows, 2 * n), generator=g, device="cuda") * 2 - 1).contiguous()
    yr = torch.empty((rows, n + 1), dtype=torch.complex64, device="cuda")
    zr = torch.empty_like(xr)
    for path, mask in (("wide", 0xFF), ("other", 0)):
        fft.ctx.set_wide_mask(mask)
        ms, best = timeit(lambda: fft.fft_batch(x, out=y), 8, 2)
        rep(f"c2c_{n}x{rows}", path, ms, best, 2 * x.numel() * 8)
        ms, best = timeit(lambda: fft.rfft_batch(xr, out=yr), 8, 2)
        rep(f"rfft_{2 * n}x{rows}", path, ms, best, xr.numel() * 4 + yr.numel() * 8)
        ms, best = timeit(lambda: fft.irfft_batch(yr, 2 * n, out=zr), 8, 2)
        rep(f"irfft_{2 * n}x{rows}", path, ms, best, xr.numel() * 4 + yr.numel() * 8)
    fft.ctx.set_wide_mask(None)
    del x, y, xr, yr, zr
