"""Host-side mirror of kofft's FFT plugin interface over the C ABI of libkofft_cuda.so.

Reference interface (paths relative to the kofft repository):
  * `FftPlanner<T>`            src/fft.rs:332-445
  * `trait FftImpl<T>`         src/fft.rs:466-587   (7 required + 6 provided methods)
  * `FftStrategy`              src/fft.rs:456-463
  * free functions             src/fft.rs:2119-2191 (`fft_parallel`, `batch`, ...)
  * `new_fft_impl()`           src/fft.rs:1954-1985

`CudaFftImpl` keeps the reference's method names, argument meaning and error behaviour:
host buffers are numpy arrays (complex64 / float32) that are transformed *in place* like the
reference's `&mut [Complex<f32>]`, and errors are raised as the `FftError` variants of
`kofft_b200.errors`.  The additional batched methods (`fft_batch`, `rfft_batch`, ...) also
accept CUDA `torch.Tensor`s, in which case nothing leaves the device.

The Rust toolchain is absent from the build image, so this Python mirror is what the parity
tests drive; the `kofft-cuda` crate sources that bind the same C ABI are in `kofft-cuda/`
(see INTEGRATION.md).
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import Optional

import numpy as np

from . import _lib
from .errors import InvalidStride, MismatchedLengths, check

try:  # torch is only needed for the device-pointer paths
    import torch
except Exception:  # pragma: no cover
    torch = None


class FftStrategy(enum.Enum):
    """src/fft.rs:456-463"""

    Radix2 = 0
    Radix4 = 1
    SplitRadix = 2
    Auto = 3


def _is_tensor(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


def _check_tensor(t, dtype, ndim=None, device=None, name="input", shape=None):
    """Every tensor whose data_ptr() reaches the C ABI goes through here: the kernels trust the pointer."""
    if not _is_tensor(t):
        raise TypeError(f"{name} must be a torch tensor")
    if not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor")
    if t.dtype != dtype:
        raise TypeError(f"{name} must have dtype {dtype}, not {t.dtype}")
    if not t.is_contiguous():
        raise TypeError(f"{name} must be contiguous")
    if ndim is not None and t.dim() != ndim:
        raise TypeError(f"{name} must have {ndim} dimensions")
    if device is not None and t.device.index != device:
        raise TypeError(f"{name} lives on cuda:{t.device.index}, the context on cuda:{device}")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise MismatchedLengths()
    return t


def _c64(a, name="input") -> np.ndarray:
    if not isinstance(a, np.ndarray) or a.dtype != np.complex64:
        raise TypeError(f"{name} must be a numpy complex64 array (it is transformed in place)")
    if not a.flags.c_contiguous or not a.flags.writeable:
        raise TypeError(f"{name} must be C-contiguous and writable")
    return a


def _f32(a, name="input", writable=True) -> np.ndarray:
    if not isinstance(a, np.ndarray) or a.dtype != np.float32:
        raise TypeError(f"{name} must be a numpy float32 array")
    if not a.flags.c_contiguous or (writable and not a.flags.writeable):
        raise TypeError(f"{name} must be C-contiguous" + (" and writable" if writable else ""))
    return a


class Context:
    """Owns one `kofft_cuda_ctx` (one device, one stream, the device-resident tables)."""

    def __init__(self, device: Optional[int] = None, exact: bool = True):
        lib = _lib.lib()
        if device is None:
            device = torch.cuda.current_device() if (torch is not None and torch.cuda.is_available()) else 0
        handle = C.c_void_p()
        check(lib.kofft_cuda_create(C.byref(handle), int(device)))
        self._h = handle
        self.device = int(device)
        self.set_exact(exact)

    @property
    def handle(self) -> C.c_void_p:
        if self._h is None:
            raise RuntimeError("context already destroyed")
        return self._h

    def set_exact(self, exact: bool) -> None:
        check(_lib.lib().kofft_cuda_set_exact(self.handle, int(bool(exact))))

    @property
    def exact(self) -> bool:
        return bool(_lib.lib().kofft_cuda_get_exact(self.handle))

    def set_max_ctas(self, n: int) -> None:
        check(_lib.lib().kofft_cuda_set_max_ctas(self.handle, int(n)))

    def set_tma_staging(self, enable: bool) -> None:
        check(_lib.lib().kofft_cuda_set_tma_staging(self.handle, int(bool(enable))))

    def set_host_pipeline(self, chunk_bytes: int) -> None:
        """chunk size of the H2D / kernel / D2H pipeline behind the host-pointer batch calls (0 = off)"""
        check(_lib.lib().kofft_cuda_set_host_pipeline(self.handle, int(chunk_bytes)))

    def set_istft_fusion(self, enable: bool, run_frames: int = 0) -> None:
        check(_lib.lib().kofft_cuda_set_istft_fusion(self.handle, int(bool(enable)), int(run_frames)))

    def set_cluster_fusion(self, enable: bool) -> None:
        check(_lib.lib().kofft_cuda_set_cluster_fusion(self.handle, int(bool(enable))))

    LARGE_TWO_KERNEL, LARGE_CLUSTER, LARGE_PIPELINED, LARGE_AUTO = 0, 1, 2, 3

    def set_large_mode(self, mode: int) -> None:
        """N > 16384: 3 = auto (default: pipelined for rfft / irfft, two kernels otherwise), 2 = persistent pipelined
        kernel, 0 = two kernels per chunk, 1 = cluster kernel."""
        check(_lib.lib().kofft_cuda_set_large_mode(self.handle, int(mode)))

    def set_split_min_log2n(self, min_log2n: int) -> None:
        """complex cores of 2^min_log2n .. 2^15 points run the warp-specialised split kernel (default 14; 16 = off)"""
        check(_lib.lib().kofft_cuda_set_split_min_log2n(self.handle, int(min_log2n)))

    def set_wide_mask(self, mask: int | None) -> None:
        """bit (L - 13) + 2 g: cores of 2^L points (L = 13, 14) of group g (0 dense C2C, 1 rfft, 2 irfft, 3 SoA / strided)
        run the wide single-CTA kernel; None = the default, 0 = off"""
        check(_lib.lib().kofft_cuda_set_wide_mask(self.handle, 0x80000000 if mask is None else int(mask)))

    def set_split_all_kinds(self, all_kinds: bool) -> None:
        """also route irfft and strided / SoA rows through the split kernel (default: C2C and rfft only)"""
        check(_lib.lib().kofft_cuda_set_split_all_kinds(self.handle, int(bool(all_kinds))))

    @property
    def fallback_count(self) -> int:
        """times a cooperative (persistent) launch was not possible and a slower path computed the result"""
        return int(_lib.lib().kofft_cuda_fallback_count(self.handle))

    def set_rfft_table_fma(self, fma: bool) -> None:
        check(_lib.lib().kofft_cuda_set_rfft_table_fma(self.handle, int(bool(fma))))
        self._rfft_fma = bool(fma)

    @property
    def rfft_table_fma(self) -> bool:
        return getattr(self, "_rfft_fma", False)

    @property
    def launch_count(self) -> int:
        return int(_lib.lib().kofft_cuda_launch_count(self.handle))

    @property
    def stream(self) -> int:
        return int(_lib.lib().kofft_cuda_stream(self.handle) or 0)

    def synchronize(self) -> None:
        check(_lib.lib().kofft_cuda_synchronize(self.handle))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            _lib.lib().kofft_cuda_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class FftPlanner:
    """`FftPlanner<f32>` (src/fft.rs:332-445): twiddle cache, now with device-resident twins."""

    def __init__(self, ctx: Optional[Context] = None):
        self._ctx = ctx
        self._cache: dict[int, np.ndarray] = {}

    def get_twiddles(self, n: int) -> np.ndarray:
        """n/2 entries exp(-2 pi i k / n) from kofft's f32 recurrence (src/fft.rs:391-405).
        Repeated calls return the same array object (the reference returns the same Arc)."""
        if n not in self._cache:
            out = np.empty(n // 2, dtype=np.complex64)
            check(_lib.lib().kofft_cuda_twiddles_host_f32(n, out.ctypes.data))
            out.flags.writeable = False
            self._cache[n] = out
        return self._cache[n]

    def device_twiddles(self, n: int) -> int:
        """Device pointer of the uploaded table; stable for the context's life."""
        if self._ctx is None:
            raise RuntimeError("planner has no device context")
        p = C.c_void_p()
        check(_lib.lib().kofft_cuda_get_twiddles(self._ctx.handle, n, C.byref(p)))
        return int(p.value)

    def plan_strategy(self, n: int) -> FftStrategy:
        """src/fft.rs:438-444"""
        return FftStrategy.SplitRadix if (n > 1 and (n & (n - 1)) == 0) else FftStrategy.Auto


def _stream_of(t) -> int:
    return int(torch.cuda.current_stream(t.device).cuda_stream)


class CudaFftImpl:
    """`impl FftImpl<f32>` on the B200 (mirrors `ScalarFftImpl<f32>`, src/fft.rs:1053-1440)."""

    def __init__(self, device: Optional[int] = None, exact: bool = True, ctx: Optional[Context] = None):
        self.ctx = ctx if ctx is not None else Context(device, exact)
        self.planner = FftPlanner(self.ctx)
        self._lib = _lib.lib()

    # -- required trait methods -------------------------------------------------------------
    def fft(self, input: np.ndarray) -> None:
        """FftImpl::fft (src/fft.rs:467, 1054-1133), in place."""
        a = _c64(input)
        check(self._lib.kofft_cuda_fft_host_f32(self.ctx.handle, a.ctypes.data, a.size, 0))

    def ifft(self, input: np.ndarray) -> None:
        """FftImpl::ifft (src/fft.rs:468, 1134-1174), in place."""
        a = _c64(input)
        check(self._lib.kofft_cuda_fft_host_f32(self.ctx.handle, a.ctypes.data, a.size, 1))

    def fft_strided(self, input: np.ndarray, stride: int, scratch: np.ndarray) -> None:
        """src/fft.rs:494-499, 1175-1199: n = len(scratch) elements at `stride`."""
        a = _c64(input)
        check(self._lib.kofft_cuda_fft_strided_host_f32(self.ctx.handle, a.ctypes.data, a.size, stride,
                                                       len(scratch), 0))

    def ifft_strided(self, input: np.ndarray, stride: int, scratch: np.ndarray) -> None:
        """src/fft.rs:500-506, 1235-1259"""
        a = _c64(input)
        check(self._lib.kofft_cuda_fft_strided_host_f32(self.ctx.handle, a.ctypes.data, a.size, stride,
                                                       len(scratch), 1))

    def fft_out_of_place_strided(self, input: np.ndarray, in_stride: int, output: np.ndarray,
                                 out_stride: int) -> None:
        """src/fft.rs:508-514, 1260-1298"""
        a, o = _c64(input), _c64(output, "output")
        check(self._lib.kofft_cuda_fft_out_of_place_strided_host_f32(
            self.ctx.handle, a.ctypes.data, a.size, in_stride, o.ctypes.data, o.size, out_stride, 0))

    def ifft_out_of_place_strided(self, input: np.ndarray, in_stride: int, output: np.ndarray,
                                  out_stride: int) -> None:
        """src/fft.rs:516-522, 1299-1336"""
        a, o = _c64(input), _c64(output, "output")
        check(self._lib.kofft_cuda_fft_out_of_place_strided_host_f32(
            self.ctx.handle, a.ctypes.data, a.size, in_stride, o.ctypes.data, o.size, out_stride, 1))

    def fft_with_strategy(self, input: np.ndarray, strategy: FftStrategy) -> None:
        """src/fft.rs:524-528, 1337-1363.  Every strategy runs the same faithful Stockham
        kernels (in the reference Radix2/SplitRadix/Auto already do; its separate in-place
        radix-4 routine for `Radix4` is outside this backend's scope)."""
        if not isinstance(strategy, FftStrategy):
            raise TypeError("strategy must be an FftStrategy")
        self.fft(input)

    # -- provided trait methods -------------------------------------------------------------
    def fft_out_of_place(self, input: np.ndarray, output: np.ndarray) -> None:
        """src/fft.rs:469-479"""
        if len(input) != len(output):
            raise MismatchedLengths()
        o = _c64(output, "output")
        o[...] = input
        self.fft(o)

    def ifft_out_of_place(self, input: np.ndarray, output: np.ndarray) -> None:
        """src/fft.rs:480-490"""
        if len(input) != len(output):
            raise MismatchedLengths()
        o = _c64(output, "output")
        o[...] = input
        self.ifft(o)

    def fft_strided_alloc(self, input: np.ndarray, stride: int) -> None:
        """src/fft.rs:530-541, 1201-1216"""
        n = 0 if stride == 0 else len(input) // stride
        self.fft_strided(input, stride, np.empty(n, dtype=np.complex64))

    def ifft_strided_alloc(self, input: np.ndarray, stride: int) -> None:
        """src/fft.rs:543-554, 1218-1233"""
        n = 0 if stride == 0 else len(input) // stride
        self.ifft_strided(input, stride, np.empty(n, dtype=np.complex64))

    def fft_split(self, re: np.ndarray, im: np.ndarray) -> None:
        """src/fft.rs:556-570, 1365-1391"""
        r, i = _f32(re, "re"), _f32(im, "im")
        check(self._lib.kofft_cuda_fft_split_host_f32(self.ctx.handle, r.ctypes.data, r.size, i.ctypes.data,
                                                     i.size, 0))

    def ifft_split(self, re: np.ndarray, im: np.ndarray) -> None:
        """src/fft.rs:572-586, 1393-1439"""
        r, i = _f32(re, "re"), _f32(im, "im")
        check(self._lib.kofft_cuda_fft_split_host_f32(self.ctx.handle, r.ctypes.data, r.size, i.ctypes.data,
                                                     i.size, 1))

    def fft_vec(self, input) -> np.ndarray:
        """ScalarFftImpl::fft_vec (src/fft.rs:1443-1449): allocate, copy, transform."""
        out = np.ascontiguousarray(np.array(input, dtype=np.complex64))
        self.fft(out)
        return out

    # -- RealFftImpl blanket methods (src/rfft.rs:775-837) -------------------------------------
    def rfft_with_scratch(self, input: np.ndarray, output: np.ndarray, scratch: np.ndarray) -> None:
        a, o = _f32(input), _c64(output, "output")
        check(self._lib.kofft_cuda_rfft_host_f32(self.ctx.handle, a.ctypes.data, a.size, o.ctypes.data, o.size,
                                                len(scratch)))

    def rfft(self, input: np.ndarray, output: np.ndarray) -> None:
        self.rfft_with_scratch(input, output, np.empty(len(input) // 2, dtype=np.complex64))

    def irfft_with_scratch(self, input: np.ndarray, output: np.ndarray, scratch: np.ndarray) -> None:
        a, o = _c64(input), _f32(output, "output")
        check(self._lib.kofft_cuda_irfft_host_f32(self.ctx.handle, a.ctypes.data, a.size, o.ctypes.data, o.size,
                                                 len(scratch)))

    def irfft(self, input: np.ndarray, output: np.ndarray) -> None:
        self.irfft_with_scratch(input, output, np.empty(len(output) // 2, dtype=np.complex64))

    # -- batched entry points (inherent methods of CudaFftImpl) ------------------
    def fft_batch(self, x, inverse: bool = False, out=None):
        """`batch()` / `batch_inverse()` (src/fft.rs:2156-2175) over dense rows [batch][n].

        numpy complex64 2-D array: transformed in place through the host-pointer ABI.
        CUDA torch.complex64 2-D tensor: stream-ordered on the current torch stream; `out`
        defaults to in place.  Returns the array/tensor holding the result."""
        if _is_tensor(x):
            _check_tensor(x, torch.complex64, 2, self.ctx.device)
            out = x if out is None else _check_tensor(out, torch.complex64, 2, self.ctx.device, "out", x.shape)
            check(self._lib.kofft_cuda_fft_c2c_f32(self.ctx.handle, x.data_ptr(), out.data_ptr(), x.shape[1],
                                                  x.shape[0], int(inverse), _stream_of(x)))
            return out
        a = _c64(x)
        if a.ndim != 2:
            raise TypeError("expected [batch, n]")
        check(self._lib.kofft_cuda_fft_batch_host_f32(self.ctx.handle, a.ctypes.data, a.shape[1], a.shape[0],
                                                     int(inverse)))
        return a

    def fft_split_batch(self, re, im, inverse: bool = False):
        """SoA rows on the device: re, im CUDA float32 [batch, n], in place."""
        _check_tensor(re, torch.float32, 2, self.ctx.device, "re")
        _check_tensor(im, torch.float32, 2, self.ctx.device, "im", re.shape)
        check(self._lib.kofft_cuda_fft_split_f32(self.ctx.handle, re.data_ptr(), im.data_ptr(), re.data_ptr(),
                                                im.data_ptr(), re.shape[1], re.shape[0], int(inverse),
                                                _stream_of(re)))
        return re, im

    def fft_strided_batch(self, x, n: int, batch: int, in_stride: int, in_dist: int, out=None,
                          out_stride: Optional[int] = None, out_dist: Optional[int] = None, inverse: bool = False):
        """Batched `fft_out_of_place_strided` on a flat CUDA complex64 tensor."""
        _check_tensor(x, torch.complex64, None, self.ctx.device)
        out = x if out is None else _check_tensor(out, torch.complex64, None, self.ctx.device, "out")
        out_stride = in_stride if out_stride is None else out_stride
        out_dist = in_dist if out_dist is None else out_dist
        if in_stride == 0 or out_stride == 0:
            raise InvalidStride()
        if n and batch:  # the last element either side touches must exist
            if (batch - 1) * in_dist + (n - 1) * in_stride >= x.numel() or \
                    (batch - 1) * out_dist + (n - 1) * out_stride >= out.numel():
                raise MismatchedLengths()
        check(self._lib.kofft_cuda_fft_strided_f32(self.ctx.handle, x.data_ptr(), in_stride, in_dist,
                                                  out.data_ptr(), out_stride, out_dist, n, batch, int(inverse),
                                                  _stream_of(x)))
        return out

    def rfft_batch(self, x, out=None):
        """Fused pack + FFT + twist per row: [batch, n] f32 -> [batch, n/2+1] complex64."""
        if _is_tensor(x):
            _check_tensor(x, torch.float32, 2, self.ctx.device)
            b, n = x.shape
            if out is None:
                out = torch.empty((b, n // 2 + 1), dtype=torch.complex64, device=x.device)
            else:
                _check_tensor(out, torch.complex64, 2, self.ctx.device, "out", (b, n // 2 + 1))
            check(self._lib.kofft_cuda_rfft_f32(self.ctx.handle, x.data_ptr(), out.data_ptr(), n, b, _stream_of(x)))
            return out
        a = _f32(x, writable=False)
        b, n = a.shape
        if out is None:
            out = np.empty((b, n // 2 + 1), dtype=np.complex64)
        check(self._lib.kofft_cuda_rfft_batch_host_f32(self.ctx.handle, a.ctypes.data, n, b, out.ctypes.data))
        return out

    def irfft_batch(self, x, n: int, out=None):
        """Fused untwist + IFFT + unpack per row: [batch, n/2+1] complex64 -> [batch, n] f32."""
        if _is_tensor(x):
            _check_tensor(x, torch.complex64, 2, self.ctx.device)
            b = x.shape[0]
            if x.shape[1] != n // 2 + 1:
                raise MismatchedLengths()
            if out is None:
                out = torch.empty((b, n), dtype=torch.float32, device=x.device)
            else:
                _check_tensor(out, torch.float32, 2, self.ctx.device, "out", (b, n))
            check(self._lib.kofft_cuda_irfft_f32(self.ctx.handle, x.data_ptr(), out.data_ptr(), n, b, _stream_of(x)))
            return out
        a = _c64(x)
        b = a.shape[0]
        if a.shape[1] != n // 2 + 1:
            raise MismatchedLengths()
        if out is None:
            out = np.empty((b, n), dtype=np.float32)
        check(self._lib.kofft_cuda_irfft_batch_host_f32(self.ctx.handle, a.ctypes.data, n, b, out.ctypes.data))
        return out


def new_fft_impl(device: Optional[int] = None, exact: bool = True) -> CudaFftImpl:
    """`new_fft_impl()` (src/fft.rs:1954-1985) — here there is exactly one backend."""
    return CudaFftImpl(device, exact)


# -- free functions (src/fft.rs:2119-2191) -------------------------------------------------------
_default_impl: Optional[CudaFftImpl] = None


def _default() -> CudaFftImpl:
    global _default_impl
    if _default_impl is None:
        _default_impl = CudaFftImpl()
    return _default_impl


def fft_parallel(input: np.ndarray) -> None:
    """src/fft.rs:2119-2121"""
    _default().fft(input)


def ifft_parallel(input: np.ndarray) -> None:
    """src/fft.rs:2125-2127"""
    _default().ifft(input)


def fft_split(re: np.ndarray, im: np.ndarray) -> None:
    """src/fft.rs:2129-2131"""
    _default().fft_split(re, im)


def ifft_split(re: np.ndarray, im: np.ndarray) -> None:
    """src/fft.rs:2133-2135"""
    _default().ifft_split(re, im)


def batch(fft: CudaFftImpl, batches) -> None:
    """src/fft.rs:2156-2164: `batches` is a list of 1-D complex64 arrays (ragged allowed, like
    the reference's `&mut [Vec<Complex<T>>]`) or one dense 2-D array.  Rows of equal length
    are transformed by one batched launch."""
    _batch(fft, batches, False)


def batch_inverse(fft: CudaFftImpl, batches) -> None:
    """src/fft.rs:2166-2175"""
    _batch(fft, batches, True)


def multi_channel(fft: CudaFftImpl, channels) -> None:
    """src/fft.rs:2177-2183"""
    _batch(fft, channels, False)


def multi_channel_inverse(fft: CudaFftImpl, channels) -> None:
    """src/fft.rs:2185-2191"""
    _batch(fft, channels, True)


def _batch(fft: CudaFftImpl, batches, inverse: bool) -> None:
    if isinstance(batches, np.ndarray) and batches.ndim == 2:
        fft.fft_batch(batches, inverse)
        return
    rows = list(batches)
    # the reference stops at the first failing row (`?`): validate in order before launching
    i = 0
    while i < len(rows):
        n = len(rows[i])
        j = i
        while j < len(rows) and len(rows[j]) == n:
            j += 1
        if j - i == 1:
            (fft.ifft if inverse else fft.fft)(rows[i])
        else:
            dense = np.ascontiguousarray(np.stack(rows[i:j]).astype(np.complex64, copy=False))
            fft.fft_batch(dense, inverse)
            for r, row in zip(rows[i:j], dense):
                r[...] = row
        i = j
