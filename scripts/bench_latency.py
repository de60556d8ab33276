"""BASELINE configs[0]: one 1024-point C2C FFT then IFFT through the host-pointer entry point, latency per call.
   KOFFT_SMALL_ZERO_COPY=0|1 python scripts/bench_latency.py"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kofft_b200  # noqa: E402

fft = kofft_b200.CudaFftImpl(device=0, exact=True)
for n in (64, 1024, 4096, 8000):
    x = (np.random.default_rng(n).uniform(-1, 1, n) + 1j * np.random.default_rng(n + 1).uniform(-1, 1, n)).astype(np.complex64)
    y = x.copy()
    for _ in range(50):
        fft.fft(y)
        fft.ifft(y)
    ts = []
    for _ in range(500):
        t0 = time.perf_counter()
        fft.fft(y)
        fft.ifft(y)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    print(json.dumps({"n": n, "zero_copy": os.environ.get("KOFFT_SMALL_ZERO_COPY", "1"), "fft_ifft_pair_us_median": round(ts[250] * 1e6, 2),
                      "p90": round(ts[450] * 1e6, 2), "min": round(ts[0] * 1e6, 2), "max_roundtrip_err": float(np.abs(y - x).max())}), flush=True)
