"""Builds and binds tests/emu/emu_engine.cpp (CPU emulation of the CUDA CTA; tests only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkofft_emu.so")
_SRC = os.path.join(_HERE, "emu_engine.cpp")
_CSRC = os.path.join(_HERE, "..", "..", "kofft_b200", "csrc")

KIND = {"c2c_fwd": 0, "c2c_inv": 1, "gen_fwd": 2, "gen_inv": 3, "stft": 4, "istft": 5, "rfft": 6, "irfft": 7, "stft_mag": 8}


def _stale() -> bool:
    if not os.path.exists(_SO):
        return True
    t = os.path.getmtime(_SO)
    deps = [_SRC] + [os.path.join(_CSRC, f) for f in ("hostdev.h", "fft_engine.cuh", "fft_kernels.cuh",
                                                      "small_kernels.cuh")]
    return any(os.path.getmtime(d) > t for d in deps)


class Emu:
    def __init__(self, lib):
        self.lib = lib
        lib.kofft_emu_run.restype = C.c_int
        lib.kofft_emu_run.argtypes = ([C.c_int, C.c_int, C.c_long, C.c_long] + [C.c_void_p] * 5 + [C.c_long] * 4
                                      + [C.c_float, C.c_void_p])
        lib.kofft_emu_bank_audit.restype = C.c_int
        lib.kofft_emu_bank_audit.argtypes = [C.c_int, C.POINTER(C.c_int)]

    def run(self, kind, exact, n, rows, table, inp=None, in2=None, out=None, out2=None, aux=None,
            p=(0, 0, 0, 0), scale=1.0, staged=False):
        self.lib.kofft_emu_set_staged(int(staged))
        def ptr(a):
            return a.ctypes.data if a is not None else None

        rc = self.lib.kofft_emu_run(KIND[kind], int(exact), n, rows, ptr(inp), ptr(in2), ptr(out), ptr(out2),
                                    ptr(aux), *[int(v) for v in p], C.c_float(scale), ptr(table))
        assert rc == 0, rc

    def bank_audit(self, L):
        out = (C.c_int * 6)()
        assert self.lib.kofft_emu_bank_audit(L, out) == 0
        return list(out)


def load() -> Emu:
    if _stale():
        subprocess.run(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-shared", "-fPIC",
                        "-fvisibility=hidden", "-o", _SO, _SRC], check=True)
    return Emu(C.CDLL(_SO))


# ---- whole-kernel emulation (cuda_emu.h coroutines running the real kernel bodies) ---------------
_SOK = os.path.join(_HERE, "libkofft_emuk.so")
_SRCK = os.path.join(_HERE, "emu_kernels.cpp")


def _stale_k() -> bool:
    if not os.path.exists(_SOK):
        return True
    t = os.path.getmtime(_SOK)
    deps = [_SRCK, os.path.join(_HERE, "cuda_emu.h")] + [
        os.path.join(_CSRC, f) for f in ("hostdev.h", "async_copy.cuh", "fft_engine.cuh", "fft_kernels.cuh",
                                         "fft_large.cuh", "fft_split32.cuh", "fft_wide.cuh", "fft_f64.cuh", "istft_fused.cuh", "small_kernels.cuh")]
    return any(os.path.getmtime(d) > t for d in deps)


class EmuKernels:
    def __init__(self, lib):
        self.lib = lib
        for f in (lib.kofft_emuk_cta, lib.kofft_emuk_large, lib.kofft_emuk_split32):
            f.restype = C.c_int
            f.argtypes = ([C.c_int, C.c_int, C.c_int, C.c_long] + [C.c_void_p] * 5 + [C.c_long] * 4
                          + [C.c_float, C.c_void_p, C.c_int, C.c_int])

    @staticmethod
    def _ptr(a):
        return a.ctypes.data if a is not None else None

    def cta(self, kind, exact, L, rows, table, inp=None, in2=None, out=None, out2=None, aux=None,
            p=(0, 0, 0, 0), scale=1.0, staged=False, grid=2):
        """CtaFft::run<STAGED> on `grid` persistent CTAs."""
        rc = self.lib.kofft_emuk_cta(KIND[kind], int(exact), L, rows, self._ptr(inp), self._ptr(in2), self._ptr(out),
                                     self._ptr(out2), self._ptr(aux), *[int(v) for v in p], C.c_float(scale),
                                     self._ptr(table), int(staged), grid)
        assert rc == 0, rc

    def large(self, kind, exact, L, rows, table, inp=None, in2=None, out=None, out2=None, aux=None,
              p=(0, 0, 0, 0), scale=1.0, grid_col=5, grid_row=16, fused=False, staged=False, pipe=False, skew=None):
        """ColPass::run + RowPass::run (two chunks) for a complex core of length 2^L, or (fused)
        LargeFused::run on `grid_col` thread-block clusters, or (pipe) LargePipe::run on `grid_col` CTAs
        (teams of 8 / 16 CTAs run concurrently, so the dependency flags are exercised)."""
        self.lib.kofft_emuk_set_fused(int(fused))
        self.lib.kofft_emuk_set_pipe(int(bool(pipe)))
        self.lib.kofft_emuk_set_skew(*(skew or (0, 0)))  # (late_from, ratio): CTAs of a team that run late
        self.lib.kofft_emuk_set_large_staged(int(staged))
        rc = self.lib.kofft_emuk_large(KIND[kind], int(exact), L, rows, self._ptr(inp), self._ptr(in2),
                                       self._ptr(out), self._ptr(out2), self._ptr(aux), *[int(v) for v in p],
                                       C.c_float(scale), self._ptr(table), grid_col, grid_row)
        assert rc == 0, rc


def _split32(self, kind, exact, L, rows, table, inp=None, in2=None, out=None, out2=None, aux=None,
             p=(0, 0, 0, 0), scale=1.0, grid=8, skew=None, staged=False):
    """Split32::run (fft_split32.cuh) on `grid` CTAs of 512 threads (teams of 4 / 2 / 1 CTAs run concurrently)."""
    self.lib.kofft_emuk_set_skew(*(skew or (0, 0)))
    rc = self.lib.kofft_emuk_split32(KIND[kind], int(exact), L, rows, self._ptr(inp), self._ptr(in2), self._ptr(out),
                                     self._ptr(out2), self._ptr(aux), *[int(v) for v in p], C.c_float(scale),
                                     self._ptr(table), grid, int(staged))
    self.lib.kofft_emuk_set_skew(0, 0)
    assert rc == 0, rc


EmuKernels.split32 = _split32


def _wide(self, kind, exact, L, rows, table, inp=None, in2=None, out=None, out2=None, aux=None,
          p=(0, 0, 0, 0), scale=1.0, grid=2, staged=False):
    """WideCta::run (fft_wide.cuh): a complex core of 2^L points in one CTA of 2^L / 32 threads, 32 elements per thread."""
    f = self.lib.kofft_emuk_wide
    f.restype = C.c_int
    f.argtypes = ([C.c_int, C.c_int, C.c_int, C.c_long] + [C.c_void_p] * 5 + [C.c_long] * 4
                  + [C.c_float, C.c_void_p, C.c_int, C.c_int])
    rc = f(KIND[kind], int(exact), L, rows, self._ptr(inp), self._ptr(in2), self._ptr(out), self._ptr(out2), self._ptr(aux),
           *[int(v) for v in p], C.c_float(scale), self._ptr(table), grid, int(staged))
    assert rc == 0, rc


EmuKernels.wide = _wide


def _wide_bank_audit(self, L, staged):
    """worst number of extra lanes per 8-byte bank pair over every half-warp exchange access of WideCta<L>"""
    f = self.lib.kofft_emuk_wide_bank_audit
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_int]
    return f(L, int(staged))


EmuKernels.wide_bank_audit = _wide_bank_audit


def _f64(self, n, rows, inp, out, table, inverse=False, grid=2, staged=False):
    """CtaFftD::run<STAGED> (n >= 32) / the literal kernels (n <= 16) for complex128 rows."""
    self.lib.kofft_emuk_set_f64_staged(int(staged))
    f = self.lib.kofft_emuk_f64
    f.restype = C.c_int
    f.argtypes = [C.c_long, C.c_long, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int]
    scale = 1.0 / float(np.float32(n))
    rc = f(n, rows, inp.ctypes.data, out.ctypes.data, int(inverse), scale, table.ctypes.data if table is not None else None, grid)
    assert rc == 0, rc


EmuKernels.f64 = _f64


def _f64_split(self, n, rows, re, im, ore, oim, table, inverse=False, grid=2):
    """CtaFftD::run over IoGenericD: split (SoA) float64 rows."""
    f = self.lib.kofft_emuk_f64_split
    f.restype = C.c_int
    f.argtypes = [C.c_long, C.c_long] + [C.c_void_p] * 4 + [C.c_int, C.c_double, C.c_void_p, C.c_int]
    rc = f(n, rows, re.ctypes.data, im.ctypes.data, ore.ctypes.data, oim.ctypes.data, int(inverse), 1.0 / float(np.float32(n)),
           table.ctypes.data if table is not None else None, grid)
    assert rc == 0, rc


EmuKernels.f64_split = _f64_split


def _f64_real(self, m, rows, inp, out, table, rtw, which, grid=2, staged=False):
    """CtaFftD::run over IoRfftD (which = 1) / IoIrfftD (which = 2); m = n/2 is the engine length."""
    self.lib.kofft_emuk_set_f64_staged(int(staged))
    f = self.lib.kofft_emuk_f64_real
    f.restype = C.c_int
    f.argtypes = [C.c_long, C.c_long, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_int]
    rc = f(m, rows, inp.ctypes.data, out.ctypes.data, which, 1.0 / float(np.float32(m)),
           table.ctypes.data if table is not None else None, rtw.ctypes.data, grid)
    assert rc == 0, rc


EmuKernels.f64_real = _f64_real


def _bind_istft(lib):
    f = lib.kofft_emuk_istft_fused
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_long,
                  C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_int]


def _istft_fused(self, exact, L, frames, window, output, norm, hop, run_frames, zero_uncovered, table, grid=3):
    """IstftFused::run: frames [ch, nframes, N] -> accumulates into output [ch, out_len] in place."""
    ch, nframes, _ = frames.shape
    rc = self.lib.kofft_emuk_istft_fused(int(exact), L, frames.ctypes.data, window.ctypes.data, output.ctypes.data,
                                         norm.ctypes.data if norm is not None else None, ch, nframes, hop,
                                         output.shape[1], run_frames, int(zero_uncovered), table.ctypes.data, grid)
    assert rc == 0, rc


EmuKernels.istft_fused = _istft_fused


def load_kernels() -> EmuKernels:
    if _stale_k():
        # KOFFT_EMU_DEFS: build-time knobs of the kernels under test (e.g. "-DKOFFT_SPLIT_LAZY_BAR2=1")
        subprocess.run(["g++", "-O1", "-DKOFFT_EMU", *os.environ.get("KOFFT_EMU_DEFS", "").split(), "-ffp-contract=off",
                        "-fno-fast-math", "-std=c++17", "-shared",
                        "-fPIC", "-fvisibility=hidden", "-I", _HERE, "-o", _SOK, _SRCK], check=True)
    lib = C.CDLL(_SOK)
    _bind_istft(lib)
    return EmuKernels(lib)
