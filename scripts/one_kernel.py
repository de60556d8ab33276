"""Launches one headline kernel a few times (target command for `ncu --set full`)."""
import sys

import torch

sys.path.insert(0, ".")
import kofft_b200  # noqa: E402
from kofft_b200 import stft as S  # noqa: E402
from kofft_b200 import window as W  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c2c"
exact = not (len(sys.argv) > 2 and sys.argv[2] == "fast")
fft = kofft_b200.CudaFftImpl(device=0, exact=exact)
g = torch.Generator(device="cuda").manual_seed(0)
if which == "c2c":
    x = torch.view_as_complex(torch.rand((65536, 4096, 2), generator=g, device="cuda") * 2 - 1).contiguous()
    y = torch.empty_like(x)
    for _ in range(6):
        fft.fft_batch(x, out=y)
elif which == "f64":
    x = torch.view_as_complex((torch.rand((16384, 4096, 2), generator=g, device="cuda", dtype=torch.float64) * 2 - 1).contiguous())
    y = torch.empty_like(x)
    f64 = kofft_b200.CudaFftImpl64(ctx=fft.ctx)
    for _ in range(5):
        f64.fft_batch(x, out=y)
elif which == "stft":
    ch, length, hop, win = 16, 28_800_000, 512, 2048
    nframes = -(-length // hop)
    sig = (torch.rand((ch, length), generator=g, device="cuda") * 2 - 1).contiguous()
    w = torch.from_numpy(W.hann(win)).cuda()
    frames = torch.empty((ch, nframes, win), dtype=torch.complex64, device="cuda")
    for _ in range(6):
        S.stft_batch(fft, sig, w, hop, nframes, out=frames)
elif which == "istft":
    ch, length, hop, win = 16, 28_800_000, 512, 2048
    nframes = -(-length // hop)
    w = torch.from_numpy(W.hann(win)).cuda()
    frames = torch.view_as_complex(torch.rand((ch, nframes, win, 2), generator=g, device="cuda") * 2 - 1).contiguous()
    out = torch.zeros((ch, length), device="cuda")
    for _ in range(4):
        S.istft_batch(fft, frames, w, hop, out)
elif which == "rfft2":  # two-kernel large-N path (column pass + row pass per L2-sized chunk)
    fft.ctx.set_large_mode(0)
    x = (torch.rand((4096, 65536), generator=g, device="cuda") * 2 - 1).contiguous()
    for _ in range(3):
        fft.rfft_batch(x)
elif which == "split":  # rfft 2^16 through the warp-specialised split kernel (default for 2^15 cores)
    x = (torch.rand((4096, 65536), generator=g, device="cuda") * 2 - 1).contiguous()
    y = torch.empty((4096, 32769), dtype=torch.complex64, device="cuda")
    for _ in range(6):
        fft.rfft_batch(x, out=y)
elif which == "wide":  # C2C 8192 through the wide single-CTA kernel
    x = torch.view_as_complex(torch.rand((16384, 8192, 2), generator=g, device="cuda") * 2 - 1).contiguous()
    y = torch.empty_like(x)
    for _ in range(6):
        fft.fft_batch(x, out=y)
elif which == "isplit":  # irfft 2^16 through the split kernel (the B warps untwist ahead of pass A)
    x = torch.view_as_complex(torch.rand((4096, 32769, 2), generator=g, device="cuda") * 2 - 1).contiguous()
    y = torch.empty((4096, 65536), device="cuda")
    for _ in range(6):
        fft.irfft_batch(x, 65536, out=y)
elif which == "rfft":  # the persistent pipelined kernel (16 elements per thread)
    fft.ctx.set_split_min_log2n(16)
    x = (torch.rand((4096, 65536), generator=g, device="cuda") * 2 - 1).contiguous()
    for _ in range(6):
        fft.rfft_batch(x)
torch.cuda.synchronize()
