"""STFT / ISTFT interface (reference: src/stft.rs).

  * `stft`             src/stft.rs:76-105      * `istft`             src/stft.rs:117-156
  * `parallel`         src/stft.rs:232-263     * `inverse_parallel`  src/stft.rs:289-343
  * `frame`            src/stft.rs:355-372     * `inverse_frame`     src/stft.rs:384-399
  * `StftStream`       src/stft.rs:160-206     * `IstftStream`       src/stft.rs:407-520

Framing + windowing is fused into the first load of the FFT kernel and window + overlap-add
+ normalisation behind the inverse FFT, so a whole signal (or many channels) is one call.
`stft_batch` / `istft_batch` are the additional inherent entry points for dense
[channels, ...] data on the host (numpy) or on the device (CUDA torch tensors).
"""
from __future__ import annotations

from typing import List

import numpy as np

from . import _lib
from .errors import InvalidHopSize, MismatchedLengths, check
from .fft import CudaFftImpl, _check_tensor, _f32, _is_tensor, _stream_of

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def _div_ceil(a: int, b: int) -> int:
    return -(-a // b)


# -- dense batched entry points -------------------------------------------------------------------
def stft_batch(fft: CudaFftImpl, signal, window, hop_size: int, nframes: int, out=None):
    """signal [channels, len] -> frames [channels, nframes, win_len] complex64."""
    lib = _lib.lib()
    if _is_tensor(signal):
        dev = fft.ctx.device
        _check_tensor(signal, torch.float32, 2, dev, "signal")
        _check_tensor(window, torch.float32, 1, dev, "window")
        ch, ln = signal.shape
        win_len = window.shape[0]
        if out is None:
            out = torch.empty((ch, nframes, win_len), dtype=torch.complex64, device=signal.device)
        else:
            _check_tensor(out, torch.complex64, 3, dev, "out", (ch, nframes, win_len))
        check(lib.kofft_cuda_stft_f32(fft.ctx.handle, signal.data_ptr(), ln, ch, window.data_ptr(), win_len,
                                      hop_size, out.data_ptr(), nframes, _stream_of(signal)))
        return out
    s = _f32(np.ascontiguousarray(signal, dtype=np.float32), "signal", writable=False)
    w = _f32(np.ascontiguousarray(window, dtype=np.float32), "window", writable=False)
    ch, ln = s.shape
    if out is None:
        out = np.zeros((ch, nframes, w.size), dtype=np.complex64)
    check(lib.kofft_cuda_stft_host_f32(fft.ctx.handle, s.ctypes.data, ln, ch, w.ctypes.data, w.size, hop_size,
                                       out.ctypes.data, nframes))
    return out


def istft_batch(fft: CudaFftImpl, frames, window, hop_size: int, output, norm=None, zero_uncovered: bool = False):
    """frames [channels, nframes, win_len] -> accumulates into output [channels, out_len]."""
    lib = _lib.lib()
    if _is_tensor(frames):
        dev = fft.ctx.device
        _check_tensor(frames, torch.complex64, 3, dev, "frames")
        ch, nframes, win_len = frames.shape
        _check_tensor(window, torch.float32, 1, dev, "window", (win_len,))
        _check_tensor(output, torch.float32, 2, dev, "output")
        if output.shape[0] != ch:
            raise MismatchedLengths()
        if norm is not None:
            _check_tensor(norm, torch.float32, 2, dev, "norm", output.shape)
        out_len = output.shape[1]
        check(lib.kofft_cuda_istft_f32(fft.ctx.handle, frames.data_ptr(), nframes, ch, window.data_ptr(), win_len,
                                       hop_size, output.data_ptr(), out_len,
                                       norm.data_ptr() if norm is not None else None, int(zero_uncovered),
                                       _stream_of(frames)))
        return output
    f = np.ascontiguousarray(frames, dtype=np.complex64)
    w = np.ascontiguousarray(window, dtype=np.float32)
    ch, nframes, win_len = f.shape
    out = _f32(output, "output")
    if norm is None and not zero_uncovered:
        norm = np.empty_like(out)
    check(lib.kofft_cuda_istft_host_f32(fft.ctx.handle, f.ctypes.data, nframes, ch, w.ctypes.data, win_len,
                                        hop_size, out.ctypes.data, out.shape[1],
                                        norm.ctypes.data if norm is not None else None,
                                        norm.size if norm is not None else 0, int(zero_uncovered)))
    return out


# -- the reference's functions ------------------------------------------------------------------------
def stft(signal, window, hop_size: int, output: List, fft: CudaFftImpl) -> None:
    """src/stft.rs:76-105.  `output` is a list with one slot per frame (the reference's
    `&mut [Vec<Complex32>]`); every slot is replaced by a complex64 array of len(window)."""
    if hop_size == 0:
        raise InvalidHopSize()
    sig = np.ascontiguousarray(signal, dtype=np.float32)
    win = np.ascontiguousarray(window, dtype=np.float32)
    if len(output) < _div_ceil(sig.size, hop_size):
        raise MismatchedLengths()
    frames = stft_batch(fft, sig.reshape(1, -1), win, hop_size, len(output))
    for i in range(len(output)):
        output[i] = frames[0, i]


def parallel(signal, window, hop_size: int, output: List, fft: CudaFftImpl) -> None:
    """src/stft.rs:232-263 -- same frames as `stft`, but the reference's rayon variant does not
    check the frame count: it fills exactly len(output) frames."""
    if hop_size == 0:
        raise InvalidHopSize()
    sig = np.ascontiguousarray(signal, dtype=np.float32)
    win = np.ascontiguousarray(window, dtype=np.float32)
    n = len(output)
    if n == 0:
        return
    # only the samples frames 0..n-1 can see; any extra frames this implies are discarded
    seen = sig[: (n - 1) * hop_size + win.size]
    nframes = max(n, _div_ceil(seen.size, hop_size))
    frames = stft_batch(fft, seen.reshape(1, -1), win, hop_size, nframes)[0]
    for i in range(n):
        output[i] = frames[i]


def istft(frames: List, window, hop_size: int, output: np.ndarray, scratch: np.ndarray, fft: CudaFftImpl) -> None:
    """src/stft.rs:117-156.  `output` is accumulated into and `scratch` receives the summed
    window power, as in the reference.  (The reference also leaves the time-domain frames in
    `frames`; the GPU path does not write them back.)"""
    if hop_size == 0:
        raise InvalidHopSize()
    if len(scratch) != len(output):
        raise MismatchedLengths()
    win = np.ascontiguousarray(window, dtype=np.float32)
    for fr in frames:
        if len(fr) != win.size:
            raise MismatchedLengths()
    out = _f32(output, "output")
    sc = _f32(scratch, "scratch")
    if len(frames) == 0:
        sc[...] = 0.0
        return
    dense = np.ascontiguousarray(np.stack([np.asarray(f, dtype=np.complex64) for f in frames]))[None]
    istft_batch(fft, dense, win, hop_size, out.reshape(1, -1), sc.reshape(1, -1))


def inverse_parallel(frames: List, window, hop_size: int, output: np.ndarray, fft: CudaFftImpl) -> None:
    """src/stft.rs:289-343: like `istft` but samples no frame covers are set to 0."""
    if hop_size == 0:
        raise InvalidHopSize()
    win = np.ascontiguousarray(window, dtype=np.float32)
    out = _f32(output, "output")
    if len(frames) == 0:
        out[...] = 0.0
        return
    dense = np.ascontiguousarray(np.stack([np.asarray(f, dtype=np.complex64) for f in frames]))[None]
    istft_batch(fft, dense, win, hop_size, out.reshape(1, -1), None, zero_uncovered=True)


def frame(signal, window, start: int, frame_out: np.ndarray, fft: CudaFftImpl) -> None:
    """src/stft.rs:355-372: one frame starting at `start`."""
    sig = np.ascontiguousarray(signal, dtype=np.float32)
    win = np.ascontiguousarray(window, dtype=np.float32)
    seg = np.zeros(win.size, dtype=np.float32)
    if start < sig.size:
        part = sig[start:start + win.size]
        seg[: part.size] = part
    res = stft_batch(fft, seg.reshape(1, -1), win, max(win.size, 1), 1)
    frame_out[...] = res[0, 0]


def inverse_frame(frame_in: np.ndarray, window, start: int, output: np.ndarray, fft: CudaFftImpl) -> None:
    """src/stft.rs:384-399: ifft in place, then windowed overlap-add without normalisation."""
    win = np.ascontiguousarray(window, dtype=np.float32)
    fft.ifft(frame_in)
    for i in range(win.size):
        if start + i < len(output):
            output[start + i] = np.float32(output[start + i]) + np.float32(frame_in[i].real) * win[i]


class StftStream:
    """src/stft.rs:160-206"""

    def __init__(self, signal, window, hop_size: int, fft: CudaFftImpl):
        if hop_size == 0:
            raise InvalidHopSize()
        self.signal = np.ascontiguousarray(signal, dtype=np.float32)
        self.window = np.ascontiguousarray(window, dtype=np.float32)
        self.hop_size = hop_size
        self.pos = 0
        self.fft = fft

    def next_frame(self, out: np.ndarray) -> bool:
        if len(out) != self.window.size:
            raise MismatchedLengths()
        if self.pos >= self.signal.size:
            return False
        frame(self.signal, self.window, self.pos, out, self.fft)
        self.pos += self.hop_size
        return True


class IstftStream:
    """src/stft.rs:407-520: overlap-add with normalisation, `hop` samples out per frame in."""

    def __init__(self, win_len: int, hop: int, window, fft: CudaFftImpl):
        if hop == 0:
            raise InvalidHopSize()
        self.win_len, self.hop = win_len, hop
        self.window = np.ascontiguousarray(window, dtype=np.float32)
        self.fft = fft
        self.buffer = np.zeros(win_len + hop * 2, dtype=np.float32)
        self.norm_buf = np.zeros(win_len + hop * 2, dtype=np.float32)
        self.buf_pos = self.out_pos = self.frame_count = 0

    def push_frame(self, frame_in) -> np.ndarray:
        if len(frame_in) != self.win_len:
            raise MismatchedLengths()
        tb = np.ascontiguousarray(np.array(frame_in, dtype=np.complex64))
        self.fft.ifft(tb)
        w = self.window
        seg = slice(self.buf_pos, self.buf_pos + self.win_len)
        self.buffer[seg] += tb.real.astype(np.float32) * w
        self.norm_buf[seg] += w * w
        self.frame_count += 1
        o0, o1 = self.out_pos, self.out_pos + self.hop
        nz = self.norm_buf[o0:o1] > np.float32(1e-8)
        self.buffer[o0:o1][nz] /= self.norm_buf[o0:o1][nz]
        self.norm_buf[o0:o1] = 0.0
        self.out_pos += self.hop
        self.buf_pos += self.hop
        if self.buf_pos + self.win_len > self.buffer.size:
            new_len = self.buf_pos + self.win_len
            self.buffer = np.concatenate([self.buffer, np.zeros(new_len - self.buffer.size, np.float32)])
            self.norm_buf = np.concatenate([self.norm_buf, np.zeros(new_len - self.norm_buf.size, np.float32)])
        tail = slice(self.buf_pos + self.win_len - self.hop, self.buf_pos + self.win_len)
        self.buffer[tail] = 0.0
        self.norm_buf[tail] = 0.0
        return self.buffer[o0:o1]

    def flush(self) -> np.ndarray:
        if self.frame_count == 0:
            return np.zeros(0, dtype=np.float32)
        o0, o1 = self.out_pos, self.buf_pos + self.win_len - self.hop
        if o0 >= o1:
            return np.zeros(0, dtype=np.float32)
        nz = self.norm_buf[o0:o1] > np.float32(1e-8)
        self.buffer[o0:o1][nz] /= self.norm_buf[o0:o1][nz]
        self.norm_buf[o0:o1] = 0.0
        self.out_pos = o1
        return self.buffer[o0:o1]


# -- device-resident streams (kofft_cuda_{stft,istft}_stream_*) -------------------------------------
class DeviceStftStream:
    """`StftStream` (src/stft.rs:160-206) with its state on the GPU: push any number of new samples for
    every channel, get the frames that became complete; `flush()` emits the remaining zero-padded ones.
    The concatenated frames are bit-identical to `stft_batch` of the whole signal."""

    def __init__(self, fft: CudaFftImpl, channels: int, window, hop_size: int):
        import ctypes as C

        self.fft, self.channels, self.hop = fft, channels, hop_size
        w = np.ascontiguousarray(window, dtype=np.float32)
        self.win_len = w.size
        self._lib = _lib.lib()
        self._h = C.c_void_p()
        check(self._lib.kofft_cuda_stft_stream_create(fft.ctx.handle, channels, w.ctypes.data, w.size, hop_size,
                                                      C.byref(self._h)))

    def _push(self, samples, flush: bool):
        import ctypes as C

        n = 0 if samples is None else samples.shape[1]
        if samples is not None:
            if (not samples.is_cuda or samples.dtype != torch.float32 or samples.dim() != 2
                    or samples.shape[0] != self.channels or samples.stride(1) != 1):
                raise TypeError("expected a CUDA float32 tensor [channels, n] with unit inner stride")
        k = int(self._lib.kofft_cuda_stft_stream_frames(self._h, n, int(flush)))
        dev = samples.device if samples is not None else torch.device("cuda", self.fft.ctx.device)
        frames = torch.empty((self.channels, k, self.win_len), dtype=torch.complex64, device=dev)
        got = C.c_size_t(0)
        stream = _stream_of(frames)
        check(self._lib.kofft_cuda_stft_stream_push(self._h, samples.data_ptr() if n else None, n,
                                                    samples.stride(0) if n else 0, frames.data_ptr(), k, C.byref(got),
                                                    int(flush), stream))
        assert got.value == k
        return frames

    def push(self, samples):
        """samples: CUDA float32 [channels, n] -> complex64 [channels, k, win_len] (k may be 0)."""
        return self._push(samples, False)

    def flush(self):
        return self._push(None, True)

    def close(self) -> None:
        if self._h:
            self._lib.kofft_cuda_stft_stream_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class DeviceIstftStream:
    """`IstftStream` (src/stft.rs:407-520) with its state on the GPU: push k frames per channel, get the
    next k * hop normalised samples; `flush()` returns the win_len - hop samples after the last frame.
    Bit-identical to `istft_batch` of all frames (the reference asserts stream == offline)."""

    def __init__(self, fft: CudaFftImpl, channels: int, window, hop: int):
        import ctypes as C

        self.fft, self.channels, self.hop = fft, channels, hop
        w = np.ascontiguousarray(window, dtype=np.float32)
        self.win_len = w.size
        self._lib = _lib.lib()
        self._h = C.c_void_p()
        check(self._lib.kofft_cuda_istft_stream_create(fft.ctx.handle, channels, w.ctypes.data, w.size, hop,
                                                       C.byref(self._h)))
        self._pushed = 0
        self._flushed = False

    def push_frames(self, frames):
        """frames: CUDA complex64 [channels, k, win_len] contiguous -> float32 [channels, k * hop]."""
        import ctypes as C

        if (not frames.is_cuda or frames.dtype != torch.complex64 or frames.dim() != 3 or not frames.is_contiguous()
                or frames.shape[0] != self.channels):
            raise TypeError("expected a contiguous CUDA complex64 tensor [channels, k, win_len]")
        if frames.shape[2] != self.win_len:
            raise MismatchedLengths()  # push_frame, src/stft.rs:450-452
        k = frames.shape[1]
        out = torch.empty((self.channels, k * self.hop), dtype=torch.float32, device=frames.device)
        got = C.c_size_t(0)
        check(self._lib.kofft_cuda_istft_stream_push(self._h, frames.data_ptr(), k, out.data_ptr(), max(k * self.hop, 1),
                                                     C.byref(got), 0, _stream_of(frames)))
        self._pushed += k
        return out

    def flush(self):
        import ctypes as C

        n = max(self.win_len - self.hop, 0) if (self._pushed and not self._flushed) else 0
        out = torch.empty((self.channels, n), dtype=torch.float32, device=torch.device("cuda", self.fft.ctx.device))
        got = C.c_size_t(0)
        check(self._lib.kofft_cuda_istft_stream_push(self._h, None, 0, out.data_ptr() if n else None, max(n, 1),
                                                     C.byref(got), 1, _stream_of(out)))
        assert got.value == n
        self._flushed = True
        return out

    def close(self) -> None:
        if self._h:
            self._lib.kofft_cuda_istft_stream_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass
