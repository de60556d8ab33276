"""ctypes binding of the CPU oracle (oracle/kofft_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under kofft_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkofft_oracle.so")

ERRORS = {
    1: "EmptyInput",
    2: "NonPowerOfTwoNoStd",
    3: "MismatchedLengths",
    4: "InvalidStride",
    5: "InvalidHopSize",
    6: "InvalidValue",
}


class OracleError(Exception):
    def __init__(self, code: int):
        self.code = code
        self.variant = ERRORS.get(code, f"internal({code})")
        super().__init__(self.variant)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("kofft_oracle.c", "kofft_oracle_f64.c")]
    fast = os.path.join(_HERE, "libkofft_oracle_fast.so")
    newest = max(os.path.getmtime(f) for f in srcs + [os.path.join(_HERE, "Makefile")])
    if force or not os.path.exists(_SO) or not os.path.exists(fast) or min(os.path.getmtime(_SO), os.path.getmtime(fast)) < newest:
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        sz, fp, ip = C.c_size_t, C.c_void_p, C.c_int
        sigs = {
            "kofft_oracle_twiddles_f32": (None, [sz, fp]),
            "kofft_oracle_rfft_twiddles_f32": (None, [sz, fp, ip]),
            "kofft_oracle_fft_f32": (ip, [fp, sz]),
            "kofft_oracle_ifft_f32": (ip, [fp, sz]),
            "kofft_oracle_fft_split_f32": (ip, [fp, fp, sz, sz]),
            "kofft_oracle_ifft_split_f32": (ip, [fp, fp, sz, sz]),
            "kofft_oracle_fft_strided_f32": (ip, [fp, sz, sz, sz, ip]),
            "kofft_oracle_fft_out_of_place_strided_f32": (ip, [fp, sz, sz, fp, sz, sz, ip]),
            "kofft_oracle_rfft_f32": (ip, [fp, sz, fp, sz, ip]),
            "kofft_oracle_irfft_f32": (ip, [fp, sz, fp, sz, ip]),
            "kofft_oracle_hann_f32": (None, [sz, fp]),
            "kofft_oracle_hamming_f32": (None, [sz, fp]),
            "kofft_oracle_blackman_f32": (None, [sz, fp]),
            "kofft_oracle_kaiser_f32": (None, [sz, C.c_float, fp]),
            "kofft_oracle_stft_f32": (ip, [fp, sz, fp, sz, sz, fp, sz]),
            "kofft_oracle_istft_f32": (ip, [fp, sz, fp, sz, sz, fp, sz, fp, sz]),
            "kofft_oracle_istft_parallel_f32": (ip, [fp, sz, fp, sz, sz, fp, sz]),
            "kofft_oracle_istft_stream_f32": (ip, [fp, sz, fp, sz, sz, fp, C.POINTER(sz)]),
            "kofft_oracle_dft_bins_f64": (None, [fp, sz, fp, sz, fp]),
            "kofft_oracle_fft_batch_f32": (ip, [fp, sz, sz, ip, ip]),
            "kofft_oracle_rfft_batch_f32": (ip, [fp, sz, sz, fp, ip, ip]),
            "kofft_oracle_irfft_batch_f32": (ip, [fp, sz, sz, fp, ip, ip]),
            "kofft_oracle_stft_batch_f32": (ip, [fp, sz, sz, fp, sz, sz, fp, sz, ip, ip]),
            "kofft_oracle_twiddles_f64": (None, [sz, fp]),
            "kofft_oracle_fft_f64": (ip, [fp, sz, ip]),
            "kofft_oracle_fft_batch_f64": (ip, [fp, sz, sz, ip, ip]),
            "kofft_oracle_rfft_twiddles_f64": (None, [sz, fp]),
            "kofft_oracle_rfft_batch_f64": (ip, [fp, sz, sz, fp, ip]),
        }
        for name, (res, args) in sigs.items():
            f = getattr(_lib, name)
            f.restype = res
            f.argtypes = args
    return _lib


def _p(a: np.ndarray) -> int:
    return a.ctypes.data


def _chk(rc: int) -> None:
    if rc != 0:
        raise OracleError(rc)


def _c64(a) -> np.ndarray:
    return np.ascontiguousarray(np.array(a, dtype=np.complex64, copy=True))


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.array(a, dtype=np.float32, copy=True))


# ---- tables -----------------------------------------------------------------------------
def twiddles(n: int) -> np.ndarray:
    out = np.empty(n // 2, dtype=np.complex64)
    lib().kofft_oracle_twiddles_f32(n, _p(out))
    return out


def rfft_twiddles(m: int, fma_mul: bool = False) -> np.ndarray:
    out = np.empty(m, dtype=np.complex64)
    lib().kofft_oracle_rfft_twiddles_f32(m, _p(out), int(fma_mul))
    return out


# ---- C2C --------------------------------------------------------------------------------
def fft(x) -> np.ndarray:
    a = _c64(x)
    _chk(lib().kofft_oracle_fft_f32(_p(a), a.size))
    return a


def ifft(x) -> np.ndarray:
    a = _c64(x)
    _chk(lib().kofft_oracle_ifft_f32(_p(a), a.size))
    return a


def fft_split(re, im, inverse: bool = False):
    r, i = _f32(re), _f32(im)
    f = lib().kofft_oracle_ifft_split_f32 if inverse else lib().kofft_oracle_fft_split_f32
    _chk(f(_p(r), _p(i), r.size, i.size))
    return r, i


def fft_strided(x, stride: int, n: int, inverse: bool = False) -> np.ndarray:
    a = _c64(x)
    _chk(lib().kofft_oracle_fft_strided_f32(_p(a), a.size, stride, n, int(inverse)))
    return a


def fft_out_of_place_strided(x, in_stride: int, out, out_stride: int, inverse: bool = False) -> np.ndarray:
    a, o = _c64(x), _c64(out)
    _chk(lib().kofft_oracle_fft_out_of_place_strided_f32(_p(a), a.size, in_stride, _p(o), o.size, out_stride,
                                                         int(inverse)))
    return o


def fft_batch(x, inverse: bool = False, nthreads: int = 1) -> np.ndarray:
    a = _c64(x)
    assert a.ndim == 2
    _chk(lib().kofft_oracle_fft_batch_f32(_p(a), a.shape[1], a.shape[0], int(inverse), nthreads))
    return a


_fast = None


def lib_fast() -> C.CDLL:
    """The timing-only -O3 / 128-bit-vector build of the same sources (oracle/Makefile); only the batch entry
    points used by bench.py are bound.  tests/test_oracle_golden.py asserts it is bit-identical to lib()."""
    global _fast
    if _fast is None:
        build()
        _fast = C.CDLL(os.path.join(_HERE, "libkofft_oracle_fast.so"))
        sz, fp, ip = C.c_size_t, C.c_void_p, C.c_int
        _fast.kofft_oracle_fft_batch_f32.restype = ip
        _fast.kofft_oracle_fft_batch_f32.argtypes = [fp, sz, sz, ip, ip]
        _fast.kofft_oracle_stft_batch_f32.restype = ip
        _fast.kofft_oracle_stft_batch_f32.argtypes = [fp, sz, sz, fp, sz, sz, fp, sz, ip, ip]
        _fast.kofft_oracle_rfft_batch_f32.restype = ip
        _fast.kofft_oracle_rfft_batch_f32.argtypes = lib().kofft_oracle_rfft_batch_f32.argtypes
    return _fast


def fft_batch_inplace(a: np.ndarray, inverse: bool = False, nthreads: int = 1, fast: bool = False) -> None:
    """Timed-baseline entry: no copies.  fast: the -O3 vectorised build (same bits, see lib_fast)."""
    assert a.dtype == np.complex64 and a.flags.c_contiguous and a.ndim == 2
    _chk((lib_fast() if fast else lib()).kofft_oracle_fft_batch_f32(_p(a), a.shape[1], a.shape[0], int(inverse), nthreads))


# ---- f64 twin (kofft_oracle_f64.c) ----------------------------------------------------------
def _c128(a) -> np.ndarray:
    return np.ascontiguousarray(np.array(a, dtype=np.complex128, copy=True))


def twiddles_f64(n: int) -> np.ndarray:
    out = np.empty(max(n // 2, 1), dtype=np.complex128)
    lib().kofft_oracle_twiddles_f64(n, _p(out))
    return out[: n // 2]


def fft_f64(x, inverse: bool = False) -> np.ndarray:
    a = _c128(x)
    _chk(lib().kofft_oracle_fft_f64(_p(a), a.size, int(inverse)))
    return a


def fft_batch_f64(x, inverse: bool = False, nthreads: int = 1) -> np.ndarray:
    a = _c128(x)
    assert a.ndim == 2
    _chk(lib().kofft_oracle_fft_batch_f64(_p(a), a.shape[1], a.shape[0], int(inverse), nthreads))
    return a


def fft_batch_f64_inplace(a: np.ndarray, inverse: bool = False, nthreads: int = 1) -> None:
    """Timed-baseline entry: no copies."""
    assert a.dtype == np.complex128 and a.flags.c_contiguous and a.ndim == 2
    _chk(lib().kofft_oracle_fft_batch_f64(_p(a), a.shape[1], a.shape[0], int(inverse), nthreads))


def rfft_twiddles_f64(m: int) -> np.ndarray:
    out = np.empty(max(m, 1), dtype=np.complex128)
    lib().kofft_oracle_rfft_twiddles_f64(m, _p(out))
    return out[:m]


def rfft_batch_f64(x) -> np.ndarray:
    a = np.ascontiguousarray(np.array(x, dtype=np.float64, copy=True))
    assert a.ndim == 2
    out = np.zeros((a.shape[0], a.shape[1] // 2 + 1), dtype=np.complex128)
    _chk(lib().kofft_oracle_rfft_batch_f64(_p(a), a.shape[1], a.shape[0], _p(out), 0))
    return out


def irfft_batch_f64(x, n: int) -> np.ndarray:
    a = _c128(x)
    assert a.ndim == 2 and a.shape[1] == n // 2 + 1
    out = np.zeros((a.shape[0], n), dtype=np.float64)
    _chk(lib().kofft_oracle_rfft_batch_f64(_p(a), n, a.shape[0], _p(out), 1))
    return out


# ---- real -------------------------------------------------------------------------------
def rfft(x, out_len: int | None = None, fma_mul: bool = False) -> np.ndarray:
    a = _f32(x)
    n = a.size
    out = np.zeros(n // 2 + 1 if out_len is None else out_len, dtype=np.complex64)
    _chk(lib().kofft_oracle_rfft_f32(_p(a), n, _p(out), out.size, int(fma_mul)))
    return out


def irfft(x, n: int, fma_mul: bool = False) -> np.ndarray:
    a = _c64(x)
    out = np.zeros(n, dtype=np.float32)
    _chk(lib().kofft_oracle_irfft_f32(_p(a), a.size, _p(out), n, int(fma_mul)))
    return out


def rfft_batch(x, fma_mul: bool = False, nthreads: int = 1) -> np.ndarray:
    a = _f32(x)
    assert a.ndim == 2
    b, n = a.shape
    out = np.zeros((b, n // 2 + 1), dtype=np.complex64)
    _chk(lib().kofft_oracle_rfft_batch_f32(_p(a), n, b, _p(out), int(fma_mul), nthreads))
    return out


def irfft_batch(x, n: int, fma_mul: bool = False, nthreads: int = 1) -> np.ndarray:
    a = _c64(x)
    assert a.ndim == 2 and a.shape[1] == n // 2 + 1
    out = np.zeros((a.shape[0], n), dtype=np.float32)
    _chk(lib().kofft_oracle_irfft_batch_f32(_p(a), n, a.shape[0], _p(out), int(fma_mul), nthreads))
    return out


# ---- windows ----------------------------------------------------------------------------
def _win(name: str, n: int, *extra) -> np.ndarray:
    out = np.empty(n, dtype=np.float32)
    getattr(lib(), f"kofft_oracle_{name}_f32")(n, *extra, _p(out))
    return out


def hann(n: int) -> np.ndarray:
    return _win("hann", n)


def hamming(n: int) -> np.ndarray:
    return _win("hamming", n)


def blackman(n: int) -> np.ndarray:
    return _win("blackman", n)


def kaiser(n: int, beta: float) -> np.ndarray:
    return _win("kaiser", n, C.c_float(beta))


# ---- STFT -------------------------------------------------------------------------------
def stft(signal, window, hop: int, nframes: int) -> np.ndarray:
    s, w = _f32(signal), _f32(window)
    frames = np.zeros((nframes, w.size), dtype=np.complex64)
    _chk(lib().kofft_oracle_stft_f32(_p(s), s.size, _p(w), w.size, hop, _p(frames), nframes))
    return frames


def istft(frames, window, hop: int, output, scratch_len: int | None = None) -> np.ndarray:
    """Returns the accumulated+normalised output (the reference adds into `output`)."""
    f, w, out = _c64(frames), _f32(window), _f32(output)
    scratch = np.ones(out.size if scratch_len is None else scratch_len, dtype=np.float32)
    _chk(lib().kofft_oracle_istft_f32(_p(f), f.shape[0], _p(w), w.size, hop, _p(out), out.size, _p(scratch),
                                      scratch.size))
    return out


def istft_parallel(frames, window, hop: int, output) -> np.ndarray:
    f, w, out = _c64(frames), _f32(window), _f32(output)
    _chk(lib().kofft_oracle_istft_parallel_f32(_p(f), f.shape[0], _p(w), w.size, hop, _p(out), out.size))
    return out


def istft_stream(frames, window, hop: int) -> np.ndarray:
    f, w = _c64(frames), _f32(window)
    out = np.zeros(f.shape[0] * hop + w.size, dtype=np.float32)
    n = C.c_size_t(0)
    _chk(lib().kofft_oracle_istft_stream_f32(_p(f), f.shape[0], _p(w), w.size, hop, _p(out), C.byref(n)))
    return out[: n.value].copy()


def stft_batch(signal, window, hop: int, nframes: int, fresh_planner: bool = False, nthreads: int = 1,
               fast: bool = False) -> np.ndarray:
    s, w = _f32(signal), _f32(window)
    assert s.ndim == 2
    ch, ln = s.shape
    frames = np.zeros((ch, nframes, w.size), dtype=np.complex64)
    _chk((lib_fast() if fast else lib()).kofft_oracle_stft_batch_f32(_p(s), ln, ch, _p(w), w.size, hop, _p(frames), nframes,
                                           int(fresh_planner), nthreads))
    return frames


# ---- f64 DFT (error yardstick) ----------------------------------------------------------
def dft_bins_f64(x, bins) -> np.ndarray:
    a = _c64(x)
    b = np.ascontiguousarray(np.array(bins, dtype=np.uint64))
    out = np.empty(b.size, dtype=np.complex128)
    lib().kofft_oracle_dft_bins_f64(_p(a), a.size, _p(b), b.size, _p(out))
    return out


def stft_magnitudes(samples, win_len: int, hop: int):
    """src/visual/spectrogram.rs:52-76: Hann STFT, |X[k]| = sqrt(re*re + im*im) for k < win_len/2
    (each operation rounded to f32, no fusion) and the largest magnitude (NaN never wins)."""
    s = _f32(samples)
    nframes = -(-s.size // hop)
    fr = stft(s, hann(win_len), hop, nframes)[:, : win_len // 2]
    re, im = fr.real.astype(np.float32), fr.imag.astype(np.float32)
    mags = np.sqrt((re * re).astype(np.float32) + (im * im).astype(np.float32), dtype=np.float32)
    mx = np.float32(0.0)
    if mags.size:
        finite_max = np.nanmax(mags) if not np.all(np.isnan(mags)) else np.float32(0.0)
        mx = np.float32(max(mx, finite_max))
    return mags, mx
