"""Worker for test_dist_one_process_per_gpu_ipc (launched with torch.distributed.run)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kofft_b200  # noqa: E402
from kofft_b200 import dist as D  # noqa: E402

log2n = int(sys.argv[1])
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
n = 1 << log2n
shard = n // world
rng = np.random.default_rng(11)  # every rank generates the same full input
x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
ctx = kofft_b200.Context(device=local)
d = D.DistFft(ctx, rank, world, log2n)
d.connect()
xs = torch.from_numpy(x[rank * shard:(rank + 1) * shard]).cuda()
out = d.transform(xs)
want = np.fft.fft(x.astype(np.complex128))[rank * shard:(rank + 1) * shard]
err = float(np.linalg.norm(out.cpu().numpy() - want) / np.linalg.norm(want))
back = d.transform(out, inverse=True)
err2 = float(np.linalg.norm(back.cpu().numpy() - x[rank * shard:(rank + 1) * shard]) / np.linalg.norm(x[:shard]))
# the library-collective arm (NCCL all-to-all instead of the kernels' peer stores) must give the same bits
col = d.transform_collective(xs).clone()
p2p = d.transform(xs)
same = float(torch.equal(torch.view_as_real(col), torch.view_as_real(p2p)))
# closed-form / direct-sum validation on the GPUs, the routine bench.py runs at 2^30
from kofft_b200 import dist_validate  # noqa: E402

v = dist_validate.validate(d, chunk=1 << 20, nbins=16)
vc = dist_validate.validate(d, transform=lambda t: d.transform_collective(t), chunk=1 << 20, nbins=8)
errs = torch.tensor([err, err2, 1.0 - same, v["rel_err_tones"], v["rel_err_impulse"], v["max_bin_err"],
                     vc["rel_err_tones"], vc["rel_err_impulse"], vc["max_bin_err"]], device="cuda")
dist.all_reduce(errs, op=dist.ReduceOp.MAX)
d.close()
dist.barrier()
if rank == 0:
    assert errs[0].item() < 2e-6 and errs[1].item() < 2e-6, errs
    assert errs[2].item() == 0.0, "collective arm differs from the peer-store arm"
    assert errs[3:].max().item() < 5e-6, errs
    print(f"dist_worker ok world={world} log2n={log2n} rel_l2={errs[0].item():.2e} roundtrip={errs[1].item():.2e}")
dist.destroy_process_group()
