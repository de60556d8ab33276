import os, sys
sys.path.insert(0, os.getcwd())
import torch, kofft_b200
fft = kofft_b200.CudaFftImpl(device=0, exact=True)
g = torch.Generator(device="cuda").manual_seed(0)
n, b = 65536, 16384
x = torch.view_as_complex((torch.rand((b, n // 2 + 1, 2), generator=g, device="cuda") * 2 - 1).contiguous())
out = torch.empty((b, n), dtype=torch.float32, device="cuda")
for mode in (0, 2, 3):
    fft.ctx.set_large_mode(mode)
    fft.irfft_batch(x, n, out=out); torch.cuda.synchronize()
    l0 = fft.ctx.launch_count
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fft.irfft_batch(x, n, out=out); e.record(); torch.cuda.synchronize()
    print("mode", mode, "launches per call", fft.ctx.launch_count - l0, "ms %.3f" % a.elapsed_time(e), flush=True)
