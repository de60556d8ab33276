"""Timing experiment: the split kernel's two roles on their own, and one team on its own.
   KOFFT_CUDA_LIB=<variant> python scripts/bench_split_roles.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kofft_b200  # noqa: E402
from scripts.bench_kernels import timeit  # noqa: E402

fft = kofft_b200.CudaFftImpl(device=0, exact=True)
g = torch.Generator(device="cuda").manual_seed(0)
for rows, max_ctas in ((16384, 0), (443, 4), (886, 8), (16384 // 2, 74)):
    x = (torch.rand((rows, 65536), generator=g, device="cuda") * 2 - 1).contiguous()
    out = torch.empty((rows, 32769), dtype=torch.complex64, device="cuda")
    fft.ctx.set_max_ctas(max_ctas)
    ms, best = timeit(lambda: fft.rfft_batch(x, out=out), 8, 2)
    teams = (max_ctas or 148) // 4
    print(json.dumps({"lib": os.path.basename(os.environ.get("KOFFT_CUDA_LIB", "default")), "rows": rows, "ctas": max_ctas or 148,
                      "ms": round(ms, 4), "us_per_tile_per_team": round(ms * 1e3 / (rows / teams), 3)}), flush=True)
    del x, out
fft.ctx.set_max_ctas(0)
