import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import kofft_b200
fft = kofft_b200.CudaFftImpl(device=0, exact=True)
C = fft.ctx
g = torch.Generator(device="cuda").manual_seed(0)
for n, rows in ((131072, 150), (65536, 150), (131072, 40)):
    xr = (torch.rand((rows, n), generator=g, device="cuda") * 2 - 1).contiguous()
    C.set_large_mode(0)
    ref = fft.rfft_batch(xr).clone()
    torch.cuda.synchronize()
    for max_ctas in (0, 16, 32, 64, 160):
        for staged in (True, False):
            C.set_large_mode(2)
            C.set_max_ctas(max_ctas)
            C.set_tma_staging(staged)
            for rep in range(2):
                y = fft.rfft_batch(xr)
                torch.cuda.synchronize()
                bad = (torch.view_as_real(y) != torch.view_as_real(ref)).any(dim=2)
                nb = int(bad.sum())
                msg = ""
                if nb:
                    r, c = torch.nonzero(bad, as_tuple=True)
                    rr = sorted(set(r.tolist()))
                    cc = c[r == rr[0]].tolist()
                    msg = f" rows {rr[:12]}{'...' if len(rr) > 12 else ''} cols(row {rr[0]}): n={len(cc)} first {cc[:8]} last {cc[-3:]} k%256 set {sorted(set(v % 256 for v in cc))[:20]}"
                print(f"n={n} rows={rows} max_ctas={max_ctas} staged={staged} rep={rep}: bad={nb}{msg}", flush=True)
    C.set_max_ctas(0); C.set_tma_staging(True)
