"""Headline kernel only: C2C N = 4096 x 65536 (and 1024 / 2048 / 256 for regressions), EXACT.  KOFFT_CUDA_LIB=<variant> python scripts/bench_headline.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kofft_b200  # noqa: E402
from scripts.bench_kernels import PEAK, timeit  # noqa: E402

fft = kofft_b200.CudaFftImpl(device=0, exact=True)
g = torch.Generator(device="cuda").manual_seed(0)
for n in (4096, 2048, 1024, 256):
    rows = 2 ** 28 // n
    x = torch.view_as_complex(torch.rand((rows, n, 2), generator=g, device="cuda") * 2 - 1).contiguous()
    y = torch.empty_like(x)
    ms, best = timeit(lambda: fft.fft_batch(x, out=y), 20, 5)
    print(json.dumps({"lib": os.path.basename(os.environ.get("KOFFT_CUDA_LIB", "default")), "n": n, "rows": rows, "ms_median": round(ms, 4),
                      "ms_best": round(best, 4), "frac_of_measured_peak": round(2 * x.numel() * 8 / ms / 1e6 / PEAK, 4)}), flush=True)
    del x, y
