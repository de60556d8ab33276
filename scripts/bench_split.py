"""Times the warp-specialised split kernel (fft_split32.cuh) against the older paths for the same shapes.
   python scripts/bench_split.py [exact|fast|both]       one JSON line per measurement"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kofft_b200  # noqa: E402
from scripts.bench_kernels import PEAK, timeit  # noqa: E402


def rep(name, mode, path, ms, best, nbytes):
    print(json.dumps({"what": name, "mode": mode, "path": path, "ms_median": round(ms, 4), "ms_best": round(best, 4),
                      "hbm_gbs": round(nbytes / ms / 1e6, 1), "frac_of_measured_peak": round(nbytes / ms / 1e6 / PEAK, 4)}), flush=True)


def main():
    sel = sys.argv[1] if len(sys.argv) > 1 else "both"
    if sel == "exit":
        sel = "exact"
    g = torch.Generator(device="cuda").manual_seed(0)
    for exact in ([True, False] if sel == "both" else [sel == "exact"]):
        mode = "exact" if exact else "fast"
        fft = kofft_b200.CudaFftImpl(device=0, exact=exact)
        C = fft.ctx
        x = (torch.rand((16384, 65536), generator=g, device="cuda") * 2 - 1).contiguous()
        out = torch.empty((16384, 32769), dtype=torch.complex64, device="cuda")
        nbytes = x.numel() * 4 + out.numel() * 8
        for path, min_l in (("split32", 15), ("pipelined", 16)):
            C.set_split_min_log2n(min_l)
            ms, best = timeit(lambda: fft.rfft_batch(x, out=out), 8, 2)
            rep("rfft_65536x16384", mode, path, ms, best, nbytes)
        back = torch.empty_like(x)
        for path, min_l in (("split32", 15), ("pipelined", 16)):
            C.set_split_min_log2n(min_l)
            ms, best = timeit(lambda: fft.irfft_batch(out, 65536, out=back), 6, 2)
            rep("irfft_65536x16384", mode, path, ms, best, nbytes)
        del x, out, back
        for n in (8192, 16384, 32768):
            rows = 2 ** 28 // n
            xc = torch.view_as_complex(torch.rand((rows, n, 2), generator=g, device="cuda") * 2 - 1).contiguous()
            yc = torch.empty_like(xc)
            for path, min_l in (("split32", 13), ("older", 16)):
                C.set_split_min_log2n(min_l)
                ms, best = timeit(lambda: fft.fft_batch(xc, out=yc), 8, 2)
                rep(f"c2c_{n}x{rows}", mode, path, ms, best, 2 * xc.numel() * 8)
            del xc, yc
        for n in (16384, 32768):
            rows = 2 ** 28 // n
            xr = (torch.rand((rows, n), generator=g, device="cuda") * 2 - 1).contiguous()
            yr = torch.empty((rows, n // 2 + 1), dtype=torch.complex64, device="cuda")
            for path, min_l in (("split32", 13), ("older", 16)):
                C.set_split_min_log2n(min_l)
                ms, best = timeit(lambda: fft.rfft_batch(xr, out=yr), 8, 2)
                rep(f"rfft_{n}x{rows}", mode, path, ms, best, xr.numel() * 4 + yr.numel() * 8)
            del xr, yr
        C.set_split_min_log2n(14)


if __name__ == "__main__":
    main()
