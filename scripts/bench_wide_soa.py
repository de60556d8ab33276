import json, os, sys
import torch
sys.path.insert(0, "/root/repo")
sys.path.insert(0, os.getcwd())
import kofft_b200
from scripts.bench_kernels import PEAK, timeit
fft = kofft_b200.CudaFftImpl(device=0, exact=True)
g = torch.Generator(device="cuda").manual_seed(0)
for n in (8192, 16384):
    rows = 2 ** 27 // n
    re = (torch.rand((rows, n), generator=g, device="cuda") * 2 - 1).contiguous()
    im = (torch.rand((rows, n), generator=g, device="cuda") * 2 - 1).contiguous()
    for path, mask in (("wide", 0xFF), ("other", 0x1F)):
        fft.ctx.set_wide_mask(mask)
        ms, best = timeit(lambda: fft.fft_split_batch(re, im), 8, 2)
        print(json.dumps({"what": f"split_rows_{n}x{rows}", "path": path, "ms_median": round(ms, 4), "frac_of_measured_peak": round(2 * rows * n * 8 / ms / 1e6 / PEAK, 4)}), flush=True)
    fft.ctx.set_wide_mask(None)
