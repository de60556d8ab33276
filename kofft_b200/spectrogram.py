"""Spectrogram magnitudes (reference: src/visual/spectrogram.rs:52-76, `stft_magnitudes`).

The reference computes a Hann STFT, keeps |X[k]| for k < win_len/2 of every frame and tracks the
largest magnitude.  Here the magnitude and the running maximum are fused behind the last FFT
stage, so the complex frames never reach HBM."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .errors import check
from .fft import CudaFftImpl, _check_tensor, _f32, _is_tensor, _stream_of

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def stft_magnitudes(fft: CudaFftImpl, samples, win_len: int, hop: int):
    """samples: 1-D float32 numpy array -> (mags [ceil(len/hop)][win_len/2] float32, max_mag)."""
    s = _f32(np.ascontiguousarray(samples, dtype=np.float32), "samples", writable=False)
    if hop == 0:
        from .errors import InvalidHopSize

        raise InvalidHopSize()
    nframes = -(-s.size // hop)
    mags = np.zeros((nframes, win_len // 2), dtype=np.float32)
    mx = C.c_float(0.0)
    check(_lib.lib().kofft_cuda_stft_magnitudes_host_f32(fft.ctx.handle, s.ctypes.data, s.size, win_len, hop,
                                                         mags.ctypes.data, nframes, C.byref(mx)))
    return mags, np.float32(mx.value)


def stft_magnitudes_batch(fft: CudaFftImpl, signal, window, hop: int, nframes: int, out=None):
    """Device tensors: signal [channels, len], window [win_len] -> (mags [channels, nframes, win_len/2], max [channels])."""
    dev = fft.ctx.device
    _check_tensor(signal, torch.float32, 2, dev, "signal")
    _check_tensor(window, torch.float32, 1, dev, "window")
    ch, ln = signal.shape
    win_len = window.shape[0]
    if out is None:
        out = torch.empty((ch, nframes, win_len // 2), dtype=torch.float32, device=signal.device)
    else:
        _check_tensor(out, torch.float32, 3, dev, "out", (ch, nframes, win_len // 2))
    mx = torch.empty((ch,), dtype=torch.float32, device=signal.device)
    check(_lib.lib().kofft_cuda_stft_magnitudes_f32(fft.ctx.handle, signal.data_ptr(), ln, ch, window.data_ptr(), win_len,
                                                    hop, out.data_ptr(), nframes, mx.data_ptr(), _stream_of(signal)))
    return out, mx
