"""f64 twin of the plugin surface: `ScalarFftImpl<f64>` / `FftPlanner<f64>` (src/fft.rs:914-1051 behind
the generic dispatch :1054-1082 and `ifft` :1134-1174), on the GPU through the C ABI's *_f64 entry points.

Power-of-two lengths 1 .. 8192.  Same conventions as `CudaFftImpl`: numpy complex128 arrays go
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
