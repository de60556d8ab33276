"""Multi-rank path on CPU: world_size 2 over gloo.  Each rank transforms its own block of rows
(no data-path collective -- rows are independent, src/fft.rs:2160-2162); the only communication
is the test's own all_gather of the results, which rank 0 compares with the oracle on the whole
batch.  The per-rank compute runs the real kernel body on the CPU emulator (tests/emu), because
the build container has no GPU; on the GPU box the same partition feeds CudaFftImpl (bench.py)."""
import os
import socket

import numpy as np
import pytest

from kofft_b200.shard import owner_of, shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
            for u in range(0, total, max(1, total // 50)):
                r = owner_of(u, total, world)
                assert blocks[r][0] <= u < blocks[r][1]
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, rows, path):
    import torch.distributed as dist
    import torch

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import kofft_oracle as ko
        from tests.emu import emu_binding

        emuk = emu_binding.load_kernels()
        rng = np.random.default_rng(0)  # same bytes on every rank
        x = (rng.uniform(-1, 1, (rows, n)) + 1j * rng.uniform(-1, 1, (rows, n))).astype(np.complex64)
        lo, hi = shard_range(rows, rank, world)
        mine = np.ascontiguousarray(x[lo:hi])
        out = np.zeros_like(mine)
        L = int(np.log2(n))
        emuk.cta("c2c_fwd", True, L, hi - lo, ko.twiddles(n), inp=mine, out=out, staged=True, grid=2)
        # gather the shards (different sizes: pad to the largest)
        width = max(shard_range(rows, r, world)[1] - shard_range(rows, r, world)[0] for r in range(world))
        pad = np.zeros((width, n), np.complex64)
        pad[: hi - lo] = out
        t = torch.from_numpy(pad.view(np.float32))
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        if rank == 0:
            parts = []
            for r in range(world):
                a, b = shard_range(rows, r, world)
                parts.append(gathered[r].numpy().view(np.complex64)[: b - a])
            full = np.concatenate(parts)
            ok = np.array_equal(full, ko.fft_batch(x))
            with open(path, "w") as f:
                f.write("ok" if ok else "mismatch")
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_row_sharding_matches_oracle(tmp_path, oracle, emuk):
    import torch.multiprocessing as mp

    path = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), 1024, 9, path), nprocs=2, join=True)
    assert open(path).read() == "ok"
