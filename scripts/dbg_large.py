import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import kofft_b200
fft = kofft_b200.CudaFftImpl(device=0, exact=True)
staged = int(sys.argv[1]) if len(sys.argv) > 1 else 1
fft.ctx.set_tma_staging(bool(staged))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
rows = 40
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.view_as_complex((torch.rand((rows, n, 2), generator=g, device="cuda") * 2 - 1).contiguous())
y = torch.empty_like(x)
fft.fft_batch(x, out=y)
torch.cuda.synchronize()
print("ok staged", staged, n, float(y.abs().max()))
