import os, sys
sys.path.insert(0, os.getcwd())
import torch, kofft_b200
for exact in (True, False):
    fft = kofft_b200.CudaFftImpl(device=0, exact=exact)
    g = torch.Generator(device="cuda").manual_seed(0)
    n, b = 65536, 16384
    x = torch.view_as_complex((torch.rand((b, n // 2 + 1, 2), generator=g, device="cuda") * 2 - 1).contiguous())
    out = torch.empty((b, n), dtype=torch.float32, device="cuda")
    for mode in (0, 2):
        fft.ctx.set_large_mode(mode)
        for _ in range(2): fft.irfft_batch(x, n, out=out)
        torch.cuda.synchronize()
        ts = []
        for _ in range(8):
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fft.irfft_batch(x, n, out=out); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
        ts.sort()
        print("irfft 65536x16384", "exact" if exact else "fast", "mode", mode, "median %.3f ms" % ts[4], flush=True)
