"""Stress of the pipelined large-N kernel's dependency flags: many launches, several grid sizes and
batch sizes, every result compared bit for bit with the two-kernel path.  python scripts/stress_large_pipe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kofft_b200  # noqa: E402

fft = kofft_b200.CudaFftImpl(device=0, exact=True)
C = fft.ctx
g = torch.Generator(device="cuda").manual_seed(1)
bad = 0
runs = 0
for n, rows in ((65536, 1111), (131072, 517), (65536, 37), (131072, 300)):
    xr = (torch.rand((rows, n), generator=g, device="cuda") * 2 - 1).contiguous()
    xc = torch.view_as_complex((torch.rand((rows, n // 2, 2), generator=g, device="cuda") * 2 - 1).contiguous())
    C.set_large_mode(0)
    ref_r = fft.rfft_batch(xr).clone()
    ref_c = fft.fft_batch(xc, out=torch.empty_like(xc)).clone()
    torch.cuda.synchronize()
    C.set_large_mode(2)
    for max_ctas in (0, 296, 160, 152, 136, 64, 16):
        C.set_max_ctas(max_ctas)
        for rep in range(12):
            y = fft.rfft_batch(xr)
            z = fft.fft_batch(xc, out=torch.empty_like(xc))
            torch.cuda.synchronize()
            runs += 2
            if not torch.equal(torch.view_as_real(y), torch.view_as_real(ref_r)):
                bad += 1
                print(f"MISMATCH rfft n={n} rows={rows} max_ctas={max_ctas} rep={rep}", flush=True)
            if not torch.equal(torch.view_as_real(z), torch.view_as_real(ref_c)):
                bad += 1
                print(f"MISMATCH c2c n={n // 2} rows={rows} max_ctas={max_ctas} rep={rep}", flush=True)
    C.set_max_ctas(0)
C.set_large_mode(3)
print(f"stress: {runs} pipelined launches compared with the two-kernel path, {bad} mismatches")
sys.exit(1 if bad else 0)
