import os, sys
sys.path.insert(0, os.getcwd())
import torch, kofft_b200
fft = kofft_b200.CudaFftImpl(device=0, exact=True)
g = torch.Generator(device="cuda").manual_seed(0)
x = (torch.rand((16384, 65536), generator=g, device="cuda") * 2 - 1).contiguous()
out = torch.empty((16384, 32769), dtype=torch.complex64, device="cuda")
fft.ctx.set_large_mode(2)
for mc in (0, 296, 288, 272, 256, 224, 192, 148):
    fft.ctx.set_max_ctas(mc)
    for _ in range(2): fft.rfft_batch(x, out=out)
    torch.cuda.synchronize()
    ts=[]
    for _ in range(6):
        a,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fft.rfft_batch(x, out=out); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
    ts.sort(); print("max_ctas", mc, "median %.3f ms" % ts[3], flush=True)
