"""Times the device-pointer kernels of every BASELINE config shape (CUDA events, data resident in HBM).
   python scripts/bench_kernels.py [which ...]    which: c2c c2c2048 stft stftmag istft rfft large f64   (default: all)
Prints one JSON line per measurement.  Used to compare tuning variants (KOFFT_CUDA_LIB=...)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kofft_b200  # noqa: E402
from kofft_b200 import stft as S  # noqa: E402
from kofft_b200 import window as W  # noqa: E402

PEAK = 6449.4
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def report(name, mode, ms, best, nbytes, **kw):
    print(json.dumps({"what": name, "mode": mode, "ms_median": round(ms, 4), "ms_best": round(best, 4),
                      "hbm_gbs": round(nbytes / ms / 1e6, 1), "frac_of_measured_peak": round(nbytes / ms / 1e6 / PEAK, 4),
                      "lib": os.environ.get("KOFFT_CUDA_LIB", "default"), **kw}), flush=True)


def main():
    which = sys.argv[1:] or ["c2c", "c2c2048", "stft", "stftmag", "istft", "rfft", "large", "f64"]
    g = torch.Generator(device="cuda").manual_seed(0)
    for exact in (True, False):
        mode = "exact" if exact else "fast"
        fft = kofft_b200.CudaFftImpl(device=0, exact=exact)
        if "c2c" in which:
            x = torch.view_as_complex(torch.rand((65536, 4096, 2), generator=g, device="cuda") * 2 - 1).contiguous()
            y = torch.empty_like(x)
            ms, best = timeit(lambda: fft.fft_batch(x, out=y), 20)
            report("c2c_4096x65536", mode, ms, best, 2 * x.numel() * 8)
            del x, y
        if "c2c2048" in which:
            for n in (256, 1024, 2048, 8192, 16384):
                x = torch.view_as_complex(torch.rand((2 ** 28 // n, n, 2), generator=g, device="cuda") * 2 - 1).contiguous()
                y = torch.empty_like(x)
                ms, best = timeit(lambda: fft.fft_batch(x, out=y), 10)
                report(f"c2c_{n}x{2 ** 28 // n}", mode, ms, best, 2 * x.numel() * 8)
                del x, y
        if "stft" in which or "istft" in which or "stftmag" in which:
            ch, length, hop, win = 64, 28_800_000, 512, 2048
            nframes = -(-length // hop)
            sig = (torch.rand((ch, length), generator=g, device="cuda") * 2 - 1).contiguous()
            w = torch.from_numpy(W.hann(win)).cuda()
            frames = torch.empty((ch, nframes, win), dtype=torch.complex64, device="cuda")
            nbytes = 4 * ch * length + 8 * ch * nframes * win
            if "stft" in which:
                ms, best = timeit(lambda: S.stft_batch(fft, sig, w, hop, nframes, out=frames), 6, 2)
                report("stft_2048_512_64ch", mode, ms, best, nbytes, frames_per_s=round(ch * nframes / ms * 1e3))
            if "stftmag" in which:
                from kofft_b200 import spectrogram as SP
                mags = torch.empty((ch, nframes, win // 2), dtype=torch.float32, device="cuda")
                mbytes = 4 * ch * length + 4 * ch * nframes * (win // 2)
                ms, best = timeit(lambda: SP.stft_magnitudes_batch(fft, sig, w, hop, nframes, out=mags), 6, 2)
                report("stft_magnitudes_2048_512_64ch", mode, ms, best, mbytes, frames_per_s=round(ch * nframes / ms * 1e3))
                del mags
            if "istft" in which:
                S.stft_batch(fft, sig, w, hop, nframes, out=frames)
                out = torch.zeros((ch, length), device="cuda")
                for fused in (True, False):
                    fft.ctx.set_istft_fusion(fused)
                    ms, best = timeit(lambda: S.istft_batch(fft, frames, w, hop, out), 4, 1)
                    report("istft_2048_512_64ch_" + ("fused" if fused else "two_kernel"), mode, ms, best, nbytes,
                           frames_per_s=round(ch * nframes / ms * 1e3))
                fft.ctx.set_istft_fusion(True)
                del out
            del sig, frames
        if "f64" in which and exact:
            fft64 = kofft_b200.CudaFftImpl64(ctx=fft.ctx)
            for n in (256, 1024, 4096, 8192):
                rows = 2 ** 26 // n
                x = torch.view_as_complex((torch.rand((rows, n, 2), generator=g, device="cuda", dtype=torch.float64) * 2 - 1).contiguous())
                y = torch.empty_like(x)
                ms, best = timeit(lambda: fft64.fft_batch(x, out=y), 10)
                report(f"c2c_f64_{n}x{rows}", "f64", ms, best, 2 * x.numel() * 16)
                del x, y
            for n in (4096, 16384):
                rows = 2 ** 27 // n
                x = (torch.rand((rows, n), generator=g, device="cuda", dtype=torch.float64) * 2 - 1).contiguous()
                y = torch.empty((rows, n // 2 + 1), dtype=torch.complex128, device="cuda")
                ms, best = timeit(lambda: fft64.rfft_batch(x, out=y), 10)
                report(f"rfft_f64_{n}x{rows}", "f64", ms, best, x.numel() * 8 + y.numel() * 16)
                del x, y
        modes = [("pipelined", 2), ("two_kernel", 0), ("cluster", 1)]
        if "rfft" in which:
            x = (torch.rand((16384, 65536), generator=g, device="cuda") * 2 - 1).contiguous()
            out = torch.empty((16384, 32769), dtype=torch.complex64, device="cuda")
            nbytes = x.numel() * 4 + out.numel() * 8
            for name, m in modes:
                fft.ctx.set_large_mode(m)
                ms, best = timeit(lambda: fft.rfft_batch(x, out=out), 6, 2)
                report("rfft_65536x16384_" + name, mode, ms, best, nbytes)
            fft.ctx.set_large_mode(3)
            del x, out
        if "large" in which:
            for n in (32768, 65536):
                x = torch.view_as_complex(torch.rand((2 ** 28 // n, n, 2), generator=g, device="cuda") * 2 - 1).contiguous()
                y = torch.empty_like(x)
                for name, m in modes:
                    fft.ctx.set_large_mode(m)
                    ms, best = timeit(lambda: fft.fft_batch(x, out=y), 6, 2)
                    report(f"c2c_{n}x{2 ** 28 // n}_" + name, mode, ms, best, 2 * x.numel() * 8)
                fft.ctx.set_large_mode(3)
                del x, y
        fft.close() if hasattr(fft, "close") else None


if __name__ == "__main__":
    main()
