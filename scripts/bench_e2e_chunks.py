"""End-to-end host-pointer batch call (kofft_cuda_fft_batch_host_f32, pinned rows) for several pipeline
chunk sizes.  python scripts/bench_e2e_chunks.py"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kofft_b200  # noqa: E402

N, ROWS = 4096, 65536
fft = kofft_b200.CudaFftImpl(device=0, exact=True)
host = torch.empty((ROWS, N), dtype=torch.complex64, pin_memory=True)
host.uniform_(-1, 1) if hasattr(host, "uniform_") else None
h = host.numpy()
for mib in (0, 4, 8, 16, 32, 64, 128, 256):
    fft.ctx.set_host_pipeline(mib << 20)
    fft.fft_batch(h)
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        fft.fft_batch(h)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    ms = ts[len(ts) // 2] * 1e3
    print(json.dumps({"chunk_mib": mib, "ms": round(ms, 2), "gb_per_s_each_way": round(ROWS * N * 8 / ms / 1e6, 1),
                      "gflops": round(5 * N * 12 * ROWS / ms / 1e6, 1)}), flush=True)
