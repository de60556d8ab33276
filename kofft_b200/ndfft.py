"""N-dimensional transforms (reference: src/ndfft.rs:74-153): `fft2d_inplace`, `fft3d_inplace`.

Same argument meaning and error behaviour as the reference (flat row-major `data`, scratch
buffers whose LENGTHS are validated); rows are one batched launch, the column / depth passes are
the strided kernel batched over every column."""
from __future__ import annotations

from . import _lib
from .errors import check
from .fft import CudaFftImpl, _c64, _is_tensor, _stream_of


def fft2d_inplace(data, rows: int, cols: int, fft: CudaFftImpl, scratch_col=None):
    """src/ndfft.rs:74-105.  numpy complex64 (host, in place) or CUDA complex64 tensor (device, in place)."""
    lib = _lib.lib()
    if _is_tensor(data):
        from .errors import MismatchedLengths

        if rows * cols != data.numel():
            raise MismatchedLengths()
        check(lib.kofft_cuda_fft2d_f32(fft.ctx.handle, data.data_ptr(), rows, cols, _stream_of(data)))
        return data
    a = _c64(data, "data")
    n_scratch = rows if scratch_col is None else len(scratch_col)
    check(lib.kofft_cuda_fft2d_host_f32(fft.ctx.handle, a.ctypes.data, a.size, rows, cols, n_scratch))
    return a


def fft3d_inplace(data, depth: int, rows: int, cols: int, fft: CudaFftImpl, scratch=None):
    """src/ndfft.rs:114-156.  scratch: optional (tube, row, col) whose lengths are validated."""
    lib = _lib.lib()
    if _is_tensor(data):
        from .errors import MismatchedLengths

        if depth * rows * cols != data.numel():
            raise MismatchedLengths()
        check(lib.kofft_cuda_fft3d_f32(fft.ctx.handle, data.data_ptr(), depth, rows, cols, _stream_of(data)))
        return data
    a = _c64(data, "data")
    lens = (depth, rows, cols) if scratch is None else tuple(len(s) for s in scratch)
    check(lib.kofft_cuda_fft3d_host_f32(fft.ctx.handle, a.ctypes.data, a.size, depth, rows, cols, *lens))
    return a
