"""One C2C transform sharded over the GPUs of a box (BASELINE configs[4]).

No reference equivalent: kofft is single-process CPU code and its planner table degenerates at
2^30.  Host side of `kofft_cuda_dist_*` (include/kofft_cuda.h): a four-step split
N = N1 * N2 whose all-to-all exchanges are peer-to-peer stores issued by the kernels themselves
over NVLink.  Two ways to drive it:

  * one process per GPU (`DistFft` + `torch.distributed`): the process group is used only to
    exchange the 128-byte CUDA IPC handles and as the barrier between phases;
  * one process, several devices (`run_local`): also usable with every rank on ONE device, which
    is how the single-GPU tests exercise the multi-rank index math.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

from . import _lib
from .errors import check
from .fft import Context

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


class DistFft:
    """Rank `rank` of `world` for one N = 2**log2n transform; owns two shard-sized device buffers."""

    def __init__(self, ctx: Context, rank: int, world: int, log2n: int):
        self.ctx, self.rank, self.world, self.log2n = ctx, int(rank), int(world), int(log2n)
        h = C.c_void_p()
        check(_lib.lib().kofft_cuda_dist_create(ctx.handle, self.rank, self.world, self.log2n, C.byref(h)))
        self._h = h
        self.shard_len = int(_lib.lib().kofft_cuda_dist_shard_len(h))
        self.n1 = 1 << (self.log2n // 2)
        self.n2 = 1 << (self.log2n - self.log2n // 2)

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("DistFft already destroyed")
        return self._h

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            _lib.lib().kofft_cuda_dist_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # -- one process per GPU ------------------------------------------------------------------
    def connect(self, group=None) -> None:
        """Exchange IPC handles over a torch.distributed process group and map the peers' buffers."""
        import torch.distributed as dist

        mine = C.create_string_buffer(128)
        check(_lib.lib().kofft_cuda_dist_ipc_handles(self.handle, mine))
        gathered = [None] * self.world
        dist.all_gather_object(gathered, bytes(mine.raw), group=group)
        blob = C.create_string_buffer(b"".join(gathered), 128 * self.world)
        check(_lib.lib().kofft_cuda_dist_connect_ipc(self.handle, blob))
        self._group = group

    def set_pieces(self, pieces: int) -> None:
        check(_lib.lib().kofft_cuda_dist_set_pieces(self.handle, int(pieces)))

    def result_view(self):
        """Buffer A as a CUDA complex64 tensor (zero-copy): where a natural-order transform leaves this
        rank's slice of the spectrum when `out` is None.  Valid until ANY rank starts the next
        transform (`transform()` begins with a barrier for exactly this reason)."""
        ptr = int(_lib.lib().kofft_cuda_dist_buffer(self.handle, 0))

        class _Raw:
            __cuda_array_interface__ = {"shape": (self.shard_len, 2), "typestr": "<f4", "data": (ptr, False), "version": 3}

        return torch.view_as_complex(torch.as_tensor(_Raw(), device=f"cuda:{self.ctx.device}"))

    def phase(self, p: int, x, out, inverse: bool = False, natural_order: bool = True, stream: Optional[int] = None):
        s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        optr = C.c_void_p(out.data_ptr()) if out is not None else C.c_void_p(0)
        check(_lib.lib().kofft_cuda_dist_phase(self.handle, int(p), C.c_void_p(x.data_ptr()), optr,
                                               int(inverse), int(natural_order), C.c_void_p(s)))

    def transform(self, x, out=None, inverse: bool = False, natural_order: bool = True, barrier=None):
        """x: this rank's contiguous slice (CUDA complex64, N/world elements).  Returns this rank's
        slice of the spectrum (natural_order) or rows [rank*N1/world ...) of X[k1 + N1*k2] as [r][k2]."""
        import torch.distributed as dist

        if x.data_ptr() == int(_lib.lib().kofft_cuda_dist_buffer(self.handle, 0)):
            x = x.clone()  # phase 0 scatters INTO buffer A on every rank: the input must not live there
        zero_copy = out is None and natural_order
        if out is None and not natural_order:
            out = torch.empty_like(x)
        if barrier is None:
            # A stream-ordered barrier: a one-element all-reduce on the current stream completes on a rank only after
            # every rank's earlier work on that stream (its phase, peer stores included) has been issued and its own
            # contribution has arrived -- no host synchronisation between the phases (the host-side barrier this
            # replaces, stream synchronise + dist.barrier, cost ~0.1 ms per phase).
            if not hasattr(self, "_token"):
                self._token = torch.zeros(1, device=x.device)

            def barrier():
                if self.world > 1:
                    dist.all_reduce(self._token, group=getattr(self, "_group", None))
        # entry barrier: phase 0 stores into EVERY rank's buffer A, which may still hold the previous
        # (zero-copy) result a slower rank is reading; all ranks have let go of it once they are here
        barrier()
        for p in range(3 if (zero_copy or not natural_order) else 4):
            self.phase(p, x, out, inverse, natural_order)
            barrier()
        return self.result_view() if zero_copy else out


def _buffer_view(d: DistFft, which: int):
    ptr = int(_lib.lib().kofft_cuda_dist_buffer(d.handle, which))

    class _Raw:
        __cuda_array_interface__ = {"shape": (d.shard_len, 2), "typestr": "<f4", "data": (ptr, False), "version": 3}

    return torch.view_as_complex(torch.as_tensor(_Raw(), device=f"cuda:{d.ctx.device}"))


def _transform_collective(self, x, inverse: bool = False, group=None):
    """The same four-step transform with `torch.distributed.all_to_all_single` (ncclAlltoAll) for the three
    exchanges instead of the kernels' peer-to-peer stores: pack -> all-to-all -> unpack per exchange (two extra
    passes over the shard each).  Natural-order result, returned as a view of buffer A.  This is the baseline arm
    bench.py times beside `transform`; both give bit-identical results."""
    import torch.distributed as dist

    lib = _lib.lib()
    s = torch.cuda.current_stream().cuda_stream
    world, r1, c2 = self.world, self.n1 // self.world, self.n2 // self.world
    if not hasattr(self, "_send"):
        self._send = torch.empty(self.shard_len, dtype=torch.complex64, device=x.device)
        self._recv = torch.empty(self.shard_len, dtype=torch.complex64, device=x.device)
    send, recv = self._send, self._recv
    bufs = [_buffer_view(self, 0), _buffer_view(self, 1)]

    def exchange(step, src, dst):
        rows, cb = (c2, r1) if step == 1 else (r1, c2)
        check(lib.kofft_cuda_dist_pack(self.handle, step, C.c_void_p(src.data_ptr()), C.c_void_p(send.data_ptr()),
                                       int(inverse), C.c_void_p(s)))
        if world > 1:
            dist.all_to_all_single(torch.view_as_real(recv), torch.view_as_real(send), group=group)
            got = recv
        else:
            got = send
        # block s = [cb][rows] from rank s -> columns [s*rows, (s+1)*rows) of the destination [cb][world*rows]
        dst.view(cb, world, rows).copy_(got.view(world, cb, rows).permute(1, 0, 2))

    exchange(0, x, bufs[0])
    check(lib.kofft_cuda_dist_local_fft(self.handle, 0, int(inverse), C.c_void_p(s)))
    exchange(1, bufs[0], bufs[1])
    check(lib.kofft_cuda_dist_local_fft(self.handle, 1, int(inverse), C.c_void_p(s)))
    exchange(2, bufs[1], bufs[0])
    return bufs[0]


DistFft.transform_collective = _transform_collective


def run_local(dists: Sequence[DistFft], xs, outs, inverse: bool = False, natural_order: bool = True) -> None:
    """All ranks driven from this process (devices may repeat): phases with device-synchronising barriers."""
    world = len(dists)
    hs = (C.c_void_p * world)(*[d.handle for d in dists])
    check(_lib.lib().kofft_cuda_dist_connect_local(hs, world))
    ins = (C.c_void_p * world)(*[C.c_void_p(x.data_ptr()) for x in xs])
    os_ = (C.c_void_p * world)(*[C.c_void_p(o.data_ptr()) for o in outs])
    check(_lib.lib().kofft_cuda_dist_run_local(hs, world, ins, os_, int(inverse), int(natural_order)))
