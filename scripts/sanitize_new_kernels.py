"""Small runs of the kernels added in round 2 (wide single-CTA kernel, split kernel incl. the irfft untwist role) for
compute-sanitizer:   compute-sanitizer --tool memcheck|racecheck|synccheck python scripts/sanitize_new_kernels.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kofft_b200  # noqa: E402

fft = kofft_b200.CudaFftImpl(device=0, exact=True)
g = torch.Generator(device="cuda").manual_seed(0)
fft.ctx.set_max_ctas(8)  # few CTAs: the tools serialise everything
for n in (8192, 16384):
    x = torch.view_as_complex(torch.rand((20, n, 2), generator=g, device="cuda") * 2 - 1).contiguous()
    y = torch.empty_like(x)
    fft.fft_batch(x, out=y)
    fft.fft_batch(x, out=y, inverse=True)
    xr = (torch.rand((20, 2 * n), generator=g, device="cuda") * 2 - 1).contiguous()
    fft.ctx.set_wide_mask(0xFF)
    yr = fft.rfft_batch(xr)
    fft.irfft_batch(yr, 2 * n)
    fft.ctx.set_wide_mask(None)
xr = (torch.rand((10, 65536), generator=g, device="cuda") * 2 - 1).contiguous()
yr = fft.rfft_batch(xr)           # split kernel, TMA tiles, shuffle twist
zr = fft.irfft_batch(yr, 65536)   # split kernel, untwist role
xc = torch.view_as_complex(torch.rand((10, 32768, 2), generator=g, device="cuda") * 2 - 1).contiguous()
fft.fft_batch(xc, out=torch.empty_like(xc))
torch.cuda.synchronize()
print("done", float(zr.abs().sum()))
