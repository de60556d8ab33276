"""Window generators with kofft's exact f32 arithmetic (reference: src/window.rs:24-61).

These are host-side table generators inside libkofft_cuda.so (the multiply by the window is
fused into the first load of the STFT kernel and the last store of the ISTFT kernel).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .errors import check

HANN, HAMMING, BLACKMAN, KAISER = 0, 1, 2, 3


def _gen(kind: int, length: int, beta: float = 0.0) -> np.ndarray:
    out = np.empty(length, dtype=np.float32)
    check(_lib.lib().kofft_cuda_window_host_f32(kind, length, C.c_float(beta), out.ctypes.data))
    return out


def hann(length: int) -> np.ndarray:
    """src/window.rs:24-28 (periodic: divides by `len`)"""
    return _gen(HANN, length)


def hamming(length: int) -> np.ndarray:
    """src/window.rs:31-35"""
    return _gen(HAMMING, length)


def blackman(length: int) -> np.ndarray:
    """src/window.rs:38-48"""
    return _gen(BLACKMAN, length)


def kaiser(length: int, beta: float) -> np.ndarray:
    """src/window.rs:52-61 (19-term I0 series, symmetric)"""
    return _gen(KAISER, length, beta)
