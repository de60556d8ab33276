"""f64 twin of the plugin surface: `ScalarFftImpl<f64>` / `FftPlanner<f64>` (src/fft.rs:914-1051 behind
the generic dispatch :1054-1082 and `ifft` :1134-1174), on the GPU through the C ABI's *_f64 entry points.

Every length the reference takes: powers of two (single-CTA kernel to 8192, multi-pass kernels to 2^26) and, through
kofft's Bluestein path, the others.  Same conventions as `CudaFftImpl`: numpy complex128 arrays go
through the host-pointer calls (in place, synchronous), CUDA torch.complex128 tensors through the
device-pointer call (stream-ordered).  No CPU fallback."""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _lib
from .errors import MismatchedLengths, check
from .fft import Context, _is_tensor, _stream_of

try:  # torch is only needed for device tensors
    import torch
except Exception:  # pragma: no cover
    torch = None


def _c128(a, name="input") -> np.ndarray:
    if not isinstance(a, np.ndarray) or a.dtype != np.complex128:
        raise TypeError(f"{name}: expected a numpy complex128 array")
    if not a.flags.c_contiguous or not a.flags.writeable:
        raise TypeError(f"{name}: expected a writable C-contiguous array")
    return a


class FftPlanner64:
    """`FftPlanner<f64>`: the twiddle table of the f64 recurrence (src/fft.rs:391-405 with T = f64)."""

    def __init__(self):
        self._cache: dict[int, np.ndarray] = {}

    def get_twiddles(self, n: int) -> np.ndarray:
        if n not in self._cache:
            out = np.empty(n // 2, dtype=np.complex128)
            check(_lib.lib().kofft_cuda_twiddles_host_f64(n, out.ctypes.data))
            out.flags.writeable = False
            self._cache[n] = out
        return self._cache[n]


class RfftPlanner64:
    """`RfftPlanner<f64>`: T'[k] = exp(-i pi k / m) from the f64 recurrence (src/rfft.rs:172-183)."""

    def __init__(self):
        self._cache: dict[int, np.ndarray] = {}

    def get_twiddles(self, m: int) -> np.ndarray:
        if m not in self._cache:
            out = np.empty(m, dtype=np.complex128)
            check(_lib.lib().kofft_cuda_rfft_twiddles_host_f64(m, out.ctypes.data))
            out.flags.writeable = False
            self._cache[m] = out
        return self._cache[m]


class CudaFftImpl64:
    """`impl FftImpl<f64>`: fft / ifft in place, plus the batched inherent method."""

    def __init__(self, device: Optional[int] = None, ctx: Optional[Context] = None):
        self.ctx = ctx if ctx is not None else Context(device)
        self._lib = _lib.lib()

    def fft(self, input: np.ndarray) -> None:
        a = _c128(input)
        check(self._lib.kofft_cuda_fft_host_f64(self.ctx.handle, a.ctypes.data, a.size, 0))

    def ifft(self, input: np.ndarray) -> None:
        a = _c128(input)
        check(self._lib.kofft_cuda_fft_host_f64(self.ctx.handle, a.ctypes.data, a.size, 1))

    @staticmethod
    def _f64(a, name="input") -> np.ndarray:
        if not isinstance(a, np.ndarray) or a.dtype != np.float64 or not a.flags.c_contiguous or not a.flags.writeable:
            raise TypeError(f"{name}: expected a writable C-contiguous numpy float64 array")
        return a

    def fft_split(self, re: np.ndarray, im: np.ndarray) -> None:
        """`fft_split` for f64 (src/fft.rs:1365-1391; tests/split64.rs)."""
        r, i = self._f64(re, "re"), self._f64(im, "im")
        check(self._lib.kofft_cuda_fft_split_host_f64(self.ctx.handle, r.ctypes.data, r.size, i.ctypes.data, i.size, 0))

    def ifft_split(self, re: np.ndarray, im: np.ndarray) -> None:
        r, i = self._f64(re, "re"), self._f64(im, "im")
        check(self._lib.kofft_cuda_fft_split_host_f64(self.ctx.handle, r.ctypes.data, r.size, i.ctypes.data, i.size, 1))

    def fft_strided(self, input: np.ndarray, stride: int, scratch: np.ndarray) -> None:
        """`fft_strided` (src/fft.rs:1175-1199): n = len(scratch) elements, `stride` apart, in place."""
        a = _c128(input)
        check(self._lib.kofft_cuda_fft_strided_host_f64(self.ctx.handle, a.ctypes.data, a.size, stride, len(scratch), 0))

    def ifft_strided(self, input: np.ndarray, stride: int, scratch: np.ndarray) -> None:
        a = _c128(input)
        check(self._lib.kofft_cuda_fft_strided_host_f64(self.ctx.handle, a.ctypes.data, a.size, stride, len(scratch), 1))

    def fft_out_of_place_strided(self, input: np.ndarray, in_stride: int, output: np.ndarray, out_stride: int) -> None:
        a, o = _c128(np.ascontiguousarray(input), "input"), _c128(output, "output")
        check(self._lib.kofft_cuda_fft_out_of_place_strided_host_f64(self.ctx.handle, a.ctypes.data, a.size, in_stride,
                                                                    o.ctypes.data, o.size, out_stride, 0))

    def ifft_out_of_place_strided(self, input: np.ndarray, in_stride: int, output: np.ndarray, out_stride: int) -> None:
        a, o = _c128(np.ascontiguousarray(input), "input"), _c128(output, "output")
        check(self._lib.kofft_cuda_fft_out_of_place_strided_host_f64(self.ctx.handle, a.ctypes.data, a.size, in_stride,
                                                                    o.ctypes.data, o.size, out_stride, 1))

    # -- RealFftImpl<f64> (src/rfft.rs:775-837) ------------------------------------------------------
    def rfft(self, input: np.ndarray, output: np.ndarray) -> None:
        """`rfft(input, output)`: n reals -> n/2 + 1 bins; errors as the reference (src/rfft.rs:433-443)."""
        a = self._f64(input)
        o = _c128(output, "output")
        if a.size and a.size % 2 == 0 and o.size != a.size // 2 + 1:
            raise MismatchedLengths()
        check(self._lib.kofft_cuda_rfft_batch_host_f64(self.ctx.handle, a.ctypes.data, a.size, 1, o.ctypes.data))

    def irfft(self, input: np.ndarray, output: np.ndarray) -> None:
        a = _c128(np.ascontiguousarray(input), "input")
        o = self._f64(output, "output")
        if o.size and o.size % 2 == 0 and a.size != o.size // 2 + 1:
            raise MismatchedLengths()
        check(self._lib.kofft_cuda_irfft_batch_host_f64(self.ctx.handle, a.ctypes.data, o.size, 1, o.ctypes.data))

    def rfft_batch(self, x, out=None):
        """[batch][n] float64 -> [batch][n/2+1] complex128 (numpy: host path; CUDA tensors: stream-ordered)."""
        if _is_tensor(x):
            if x.dtype != torch.float64 or x.dim() != 2 or not x.is_contiguous() or not x.is_cuda:
                raise TypeError("expected a contiguous CUDA float64 tensor [batch, n]")
            b, n = x.shape
            if out is None:
                out = torch.empty((b, n // 2 + 1), dtype=torch.complex128, device=x.device)
            check(self._lib.kofft_cuda_rfft_f64(self.ctx.handle, x.data_ptr(), out.data_ptr(), n, b, _stream_of(x)))
            return out
        a = self._f64(np.ascontiguousarray(x))
        b, n = a.shape
        o = np.empty((b, n // 2 + 1), dtype=np.complex128) if out is None else out
        check(self._lib.kofft_cuda_rfft_batch_host_f64(self.ctx.handle, a.ctypes.data, n, b, o.ctypes.data))
        return o

    def irfft_batch(self, x, n: int, out=None):
        if _is_tensor(x):
            if x.dtype != torch.complex128 or x.dim() != 2 or not x.is_contiguous() or not x.is_cuda:
                raise TypeError("expected a contiguous CUDA complex128 tensor [batch, n/2+1]")
            if x.shape[1] != n // 2 + 1:
                raise MismatchedLengths()
            if out is None:
                out = torch.empty((x.shape[0], n), dtype=torch.float64, device=x.device)
            check(self._lib.kofft_cuda_irfft_f64(self.ctx.handle, x.data_ptr(), out.data_ptr(), n, x.shape[0], _stream_of(x)))
            return out
        a = _c128(np.ascontiguousarray(x))
        if a.shape[1] != n // 2 + 1:
            raise MismatchedLengths()
        o = np.empty((a.shape[0], n), dtype=np.float64) if out is None else out
        check(self._lib.kofft_cuda_irfft_batch_host_f64(self.ctx.handle, a.ctypes.data, n, a.shape[0], o.ctypes.data))
        return o

    def fft_batch(self, x, inverse: bool = False, out=None):
        """Dense rows [batch][n]: numpy complex128 (in place, host path) or CUDA torch.complex128."""
        if _is_tensor(x):
            if x.dtype != torch.complex128 or x.dim() != 2 or not x.is_contiguous() or not x.is_cuda:
                raise TypeError("expected a contiguous CUDA complex128 tensor [batch, n]")
            out = x if out is None else out
            if out.shape != x.shape or out.dtype != x.dtype or not out.is_contiguous():
                raise MismatchedLengths()
            check(self._lib.kofft_cuda_fft_c2c_f64(self.ctx.handle, x.data_ptr(), out.data_ptr(), x.shape[1],
                                                  x.shape[0], int(inverse), _stream_of(x)))
            return out
        a = _c128(x)
        if a.ndim != 2:
            raise TypeError("expected a 2-D array [batch, n]")
        check(self._lib.kofft_cuda_fft_batch_host_f64(self.ctx.handle, a.ctypes.data, a.shape[1], a.shape[0],
                                                     int(inverse)))
        return a
