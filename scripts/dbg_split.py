import sys, os, traceback, atexit
sys.path.insert(0, ".")
import numpy as np, torch
import kofft_b200
atexit.register(lambda: print("atexit handler ran", flush=True))
fft = kofft_b200.CudaFftImpl(device=0, exact=True)
print("ctx ok", flush=True)
rng = np.random.default_rng(0)
staged = sys.argv[1] == "1"
fft.ctx.set_tma_staging(staged)
x = (rng.uniform(-1, 1, (3, 32768)) + 1j * rng.uniform(-1, 1, (3, 32768))).astype(np.complex64)
d = torch.from_numpy(x).cuda()
y = torch.empty_like(d)
print("launch staged=", staged, flush=True)
try:
    fft.fft_batch(d, out=y)
    print("returned", flush=True)
    torch.cuda.synchronize()
    print("synced", fft.ctx.fallback_count, flush=True)
except BaseException as e:
    traceback.print_exc()
    print("exception", repr(e), flush=True)
