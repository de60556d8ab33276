"""Times the sharded single transform (BASELINE configs[4]) -- one process per GPU under
torch.distributed.run, or a single process (world 1).  2^27 points (1 GiB) per GPU by default.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 scripts/bench_dist.py [log2n]
Prints one JSON line on rank 0: wall time per transform (max over ranks), per-phase device times."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kofft_b200  # noqa: E402
from kofft_b200 import dist as D  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 27 + int(np.log2(world))
natural = not (len(sys.argv) > 2 and sys.argv[2] == "transposed")
n = 1 << log2n
shard = n // world
ctx = kofft_b200.Context(device=local)
d = D.DistFft(ctx, rank, world, log2n)
if world > 1:
    d.connect()
else:
    D.run_local([d], [torch.zeros(shard, dtype=torch.complex64, device="cuda")],
                [torch.zeros(shard, dtype=torch.complex64, device="cuda")])
g = torch.Generator(device="cuda").manual_seed(5 + rank)
x = torch.view_as_complex(torch.rand((shard, 2), generator=g, device="cuda") * 2 - 1).contiguous()
out = torch.empty_like(x)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


phase_ms = [[] for _ in range(4)]
walls = []
nph = 3  # natural order: the result stays in buffer A (zero-copy view), no final device copy
for it in range(2 + 5):
    barrier()
    t0 = time.perf_counter()
    for p in range(nph):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        d.phase(p, x, None if natural else out, False, natural)
        b.record()
        barrier()
        if it >= 2:
            phase_ms[p].append(a.elapsed_time(b))
    if it >= 2:
        walls.append((time.perf_counter() - t0) * 1e3)
wall = torch.tensor([float(np.median(walls))] + [float(np.median(v)) if v else 0.0 for v in phase_ms], device="cuda")
if world > 1:
    dist.all_reduce(wall, op=dist.ReduceOp.MAX)
# sampled-bin check against f64 direct sums is in tests/; here: Parseval on this rank's slices
ex = float((x.abs().double() ** 2).sum())
res = d.result_view() if natural else out
eo = float((res.abs().double() ** 2).sum())
tot = torch.tensor([ex, eo], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(tot)
if rank == 0:
    w = wall.tolist()
    flops = 5.0 * n * log2n
    print(json.dumps({"what": "dist_c2c_f32", "log2n": log2n, "world": world, "natural_order": natural,
                      "ms_per_transform_wall_max_over_ranks": round(w[0], 3),
                      "phase_ms_device_max_over_ranks": [round(v, 3) for v in w[1:1 + nph]],
                      "gflops": round(flops / w[0] / 1e6, 1),
                      "bytes_per_gpu_per_exchange": shard * 8 * (world - 1) // world,
                      "parseval_ratio": tot[1].item() / (tot[0].item() * n)}), flush=True)
d.close()
if world > 1:
    dist.destroy_process_group()
