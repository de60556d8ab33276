"""Generates tests/golden/oracle_pin.npz from the oracle (oracle/kofft_oracle.c).

The reference is a Rust crate and cannot be run in this image (no rustc/cargo), so these are
NOT reference outputs: they pin the oracle's own tables and a few outputs, produced in the
build container (glibc 2.39), so that tests on another box detect any libm/compiler drift.
The reference's own golden vectors are closed-form and restated in test_oracle_golden.py.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import kofft_oracle as ko  # noqa: E402

ko.build()
rng = np.random.default_rng(2026)
d = {}
for n in (32, 1024, 2048, 4096, 32768):
    d[f"tw_{n}"] = ko.twiddles(n)
d["rtw_32768"] = ko.rfft_twiddles(32768)
d["rtw_fma_32768"] = ko.rfft_twiddles(32768, True)
x = (rng.uniform(-1, 1, 4096) + 1j * rng.uniform(-1, 1, 4096)).astype(np.complex64)
d["x_4096"], d["y_4096"], d["yi_4096"] = x, ko.fft(x), ko.ifft(x)
xr = rng.uniform(-1, 1, 8192).astype(np.float32)
d["xr_8192"], d["yr_8192"] = xr, ko.rfft(xr)
d["hann_2048"] = ko.hann(2048)
sig = rng.uniform(-1, 1, 4096).astype(np.float32)
d["sig"] = sig
d["stft_frames"] = ko.stft(sig, d["hann_2048"], 512, 8)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_pin.npz"), **d)
print("wrote oracle_pin.npz", {k: v.shape for k, v in d.items()})
