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

KIND = {"c2c_fwd": 0, "c2c_inv": 1, "gen_fwd": 2, "gen_inv": 3, "stft": 4, "istft": 5, "rfft": 6, "irfft": 7}


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
