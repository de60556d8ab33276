"""Pure-write, pure-read and copy bandwidth of the device with library kernels (context for write-dominated kernels:
the STFT writes 7x what it reads)."""
import json

import torch

n = 1 << 31  # 8 GiB of f32
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")


def t(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = t(lambda: a.fill_(1.0))
print(json.dumps({"op": "fill 8 GiB (write only)", "ms": ms, "gbs": n * 4 / ms / 1e6}))
ms = t(lambda: a.zero_())
print(json.dumps({"op": "memset 8 GiB (write only)", "ms": ms, "gbs": n * 4 / ms / 1e6}))
ms = t(lambda: b.copy_(a))
print(json.dumps({"op": "copy 8 GiB (read + write)", "ms": ms, "gbs": 2 * n * 4 / ms / 1e6}))
ms = t(lambda: a.sum())
print(json.dumps({"op": "sum 8 GiB (read only)", "ms": ms, "gbs": n * 4 / ms / 1e6}))
