"""N = 2^15 / 2^16 through the two-kernel (mode 0) and pipelined (mode 2) paths for the kinds the auto mode
does not cover by measurement yet: C2C inverse and split (SoA) rows.  python scripts/bench_large_kinds.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kofft_b200  # noqa: E402

fft = kofft_b200.CudaFftImpl(device=0, exact=True)
g = torch.Generator(device="cuda").manual_seed(0)


def timeit(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    ts.sort()
    return ts[4]


for n in (32768, 65536):
    b = 2 ** 28 // n
    x = torch.view_as_complex((torch.rand((b, n, 2), generator=g, device="cuda") * 2 - 1).contiguous())
    y = torch.empty_like(x)
    re = (torch.rand((b, n), generator=g, device="cuda") * 2 - 1).contiguous()
    im = (torch.rand((b, n), generator=g, device="cuda") * 2 - 1).contiguous()
    for mode in (0, 2):
        fft.ctx.set_large_mode(mode)
        print(f"n={n} mode={mode} c2c_inv {timeit(lambda: fft.fft_batch(x, inverse=True, out=y)):.3f} ms  "
              f"split_fwd {timeit(lambda: fft.fft_split_batch(re, im)):.3f} ms", flush=True)
