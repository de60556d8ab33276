import os, sys, json
sys.path.insert(0, os.getcwd())
import torch
import kofft_b200
from kofft_b200 import stft as S, window as W, spectrogram as SP
fft = kofft_b200.CudaFftImpl(device=0, exact=(os.environ.get("MODE","exact")=="exact"))
g = torch.Generator(device="cuda").manual_seed(0)
ch, length, hop, win = 64, 28_800_000, 512, 2048
nframes = -(-length // hop)
sig = (torch.rand((ch, length), generator=g, device="cuda") * 2 - 1).contiguous()
w = torch.from_numpy(W.hann(win)).cuda()
frames = torch.empty((ch, nframes, win), dtype=torch.complex64, device="cuda")
for _ in range(3): S.stft_batch(fft, sig, w, hop, nframes, out=frames)
torch.cuda.synchronize()
ts=[]
for _ in range(24):
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); S.stft_batch(fft, sig, w, hop, nframes, out=frames); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ts.sort()
print(os.environ.get("KOFFT_CUDA_LIB","default").split("_")[-1], os.environ.get("MODE","exact"), "min %.2f p25 %.2f med %.2f p75 %.2f max %.2f" % (ts[0], ts[6], ts[12], ts[18], ts[-1]))
