"""irfft 2^16 x 16384 through the split kernel (untwist in the B warps), timed, and compared bit for bit with the
older pipelined path.   KOFFT_CUDA_LIB=<variant> python scripts/bench_irfft_split.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kofft_b200  # noqa: E402
from scripts.bench_kernels import PEAK, timeit  # noqa: E402

fft = kofft_b200.CudaFftImpl(device=0, exact=True)
g = torch.Generator(device="cuda").manual_seed(0)
rows = 16384
x = torch.view_as_complex(torch.rand((rows, 32769, 2), generator=g, device="cuda") * 2 - 1).contiguous()
y = torch.empty((rows, 65536), device="cuda")
nbytes = x.numel() * 8 + y.numel() * 4
ms, best = timeit(lambda: fft.irfft_batch(x, 65536, out=y), 8, 2)
fft.ctx.set_split_min_log2n(16)
ref = torch.empty((1024, 65536), device="cuda")
fft.irfft_batch(x[:1024], 65536, out=ref)
same = bool(torch.equal(ref, y[:1024]))
print(json.dumps({"lib": os.path.basename(os.environ.get("KOFFT_CUDA_LIB", "default")), "what": "irfft_65536x16384", "ms_median": round(ms, 4),
                  "ms_best": round(best, 4), "frac_of_measured_peak": round(nbytes / ms / 1e6 / PEAK, 4),
                  "bit_identical_to_pipelined_path": same}), flush=True)
