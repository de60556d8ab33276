"""Sweep of the sharded transform's overlap knobs (pieces per phase x SMs left to the scatter kernel), one process
per GPU under torch.distributed.run.  One JSON line per configuration on rank 0."""
import json
import os
import statistics
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
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
log2n = 27 + int(np.log2(world))
shard = (1 << log2n) // world
ctx = kofft_b200.Context(device=local)
g = torch.Generator(device="cuda").manual_seed(5 + rank)
x = torch.view_as_complex(torch.rand((shard, 2), generator=g, device="cuda") * 2 - 1).contiguous()
tok = torch.zeros(1, device="cuda")


def mx(v):
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for pieces in [int(v) for v in os.environ.get("SWEEP_PIECES", "1,2,4,8").split(",")]:
    for reserve in [int(v) for v in os.environ.get("SWEEP_RESERVE", "16,32,48").split(",")]:
        os.environ["KOFFT_DIST_PIECES"] = str(pieces)
        os.environ["KOFFT_DIST_RESERVE_SMS"] = str(reserve)
        d = D.DistFft(ctx, rank, world, log2n)
        d.connect()
        walls, phases = [], [[], [], []]
        for it in range(6):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            d.transform(x, natural_order=True)
            torch.cuda.synchronize()
            if it >= 2:
                walls.append(time.perf_counter() - t0)
        for it in range(3):  # per-phase device times (host barrier between phases so that they do not overlap)
            for p in range(3):
                dist.barrier()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                d.phase(p, x, None, False, True)
                b.record()
                torch.cuda.synchronize()
                phases[p].append(a.elapsed_time(b))
        w = mx(statistics.median(walls))
        ph = [mx(statistics.median(v)) for v in phases]
        if rank == 0:
            print(json.dumps({"log2n": log2n, "world": world, "pieces": pieces, "reserve_sms": reserve,
                              "ms_natural_order": round(w * 1e3, 3), "phase_ms": [round(v, 3) for v in ph]}), flush=True)
        d.close()
dist.barrier()
dist.destroy_process_group()
