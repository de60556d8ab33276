"""Full-size validation of the sharded transform against closed forms and f64 direct sums, computed on the GPUs.

kofft has no usable output at these sizes (its f32 twiddle recurrence degenerates), so the sharded
2^30 transform is checked the way SURVEY.md 8d prescribes for config 5, at the REAL size:
  (i)   K tones  sum_m a_m exp(2 pi i f_m n / N)  -> exact spikes N a_m at f_m, zero elsewhere
  (ii)  an impulse at n0                          -> the exact phase ramp exp(-2 pi i k n0 / N) at EVERY bin
  (iii) uniform noise (seed 5) with 64 random bins compared with f64 direct sums over all N points
Every rank generates / checks its own slice; phases use exact integer reduction ((f * n) mod N in int64) and
f64 sin / cos, so the yardstick is accurate to ~1e-15.  Used by bench.py (figures land in SCALE_r*.json) and by
tests/test_dist_fft.py at smaller sizes."""
from __future__ import annotations

import math

import torch


def _phase(idx: torch.Tensor, mult: int, n: int, sign: float) -> torch.Tensor:
    """exp(sign * 2 pi i * ((idx * mult) mod n) / n) as complex128, idx int64 (products stay below 2^62)."""
    p = torch.remainder(idx * int(mult), n).to(torch.float64) * (sign * 2.0 * math.pi / n)
    return torch.complex(torch.cos(p), torch.sin(p))


def _allreduce(t: torch.Tensor, world: int, op=None):
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(t, op=op if op is not None else dist.ReduceOp.SUM)
    return t


def validate(dfft, transform=None, chunk: int = 1 << 24, nbins: int = 64) -> dict:
    """dfft: a connected DistFft; transform(x) -> this rank's natural-order slice (default dfft.transform).
    Returns {"rel_err_tones", "rel_err_impulse", "max_bin_err", ...} (identical on every rank)."""
    import torch.distributed as dist

    world, rank, log2n = dfft.world, dfft.rank, dfft.log2n
    n = 1 << log2n
    shard = n // world
    dev = torch.device("cuda", dfft.ctx.device)
    lo = rank * shard
    run = transform if transform is not None else (lambda x: dfft.transform(x))
    out = {}

    # ---- (i) tones -------------------------------------------------------------------------------------------
    tones = [(3, 1.0), (4097, 0.5), (n // 2 + 17, 0.25), (n - 5, 2.0), (n // 3, 0.75)]
    x = torch.zeros(shard, dtype=torch.complex64, device=dev)
    for c0 in range(0, shard, chunk):
        idx = torch.arange(lo + c0, lo + min(c0 + chunk, shard), dtype=torch.int64, device=dev)
        acc = torch.zeros(idx.numel(), dtype=torch.complex128, device=dev)
        for f, a in tones:
            acc += a * _phase(idx, f, n, +1.0)
        x[c0:c0 + idx.numel()] = acc.to(torch.complex64)
    y = run(x)
    # ||Y - exact||^2 = sum |Y|^2 over the non-tone bins + sum |Y[f] - N a|^2 over the tone bins
    e2 = (y.real.double() ** 2 + y.imag.double() ** 2).sum()
    worst = torch.zeros(1, dtype=torch.float64, device=dev)
    for f, a in tones:
        if lo <= f < lo + shard:
            v = y[f - lo].to(torch.complex128)
            e2 = e2 - (v.real ** 2 + v.imag ** 2) + ((v.real - n * a) ** 2 + v.imag ** 2)
            worst = torch.maximum(worst, (((v.real - n * a) ** 2 + v.imag ** 2).sqrt() / (n * a)).reshape(1))
    ref2 = sum((n * a) ** 2 for _, a in tones)
    e2 = _allreduce(e2.reshape(1).clone(), world)
    worst = _allreduce(worst, world, dist.ReduceOp.MAX if world > 1 else None)
    out["rel_err_tones"] = float(torch.sqrt(torch.clamp(e2, min=0.0) / ref2).item())
    out["worst_tone_rel_err"] = float(worst.item())
    del x, y

    # ---- (ii) impulse: every bin of the spectrum against the exact phase ramp ---------------------------------
    n0 = (n // 7) | 1
    x = torch.zeros(shard, dtype=torch.complex64, device=dev)
    if lo <= n0 < lo + shard:
        x[n0 - lo] = 1.0
    y = run(x)
    e2 = torch.zeros(1, dtype=torch.float64, device=dev)
    emax = torch.zeros(1, dtype=torch.float64, device=dev)
    for c0 in range(0, shard, chunk):
        k = torch.arange(lo + c0, lo + min(c0 + chunk, shard), dtype=torch.int64, device=dev)
        d = y[c0:c0 + k.numel()].to(torch.complex128) - _phase(k, n0, n, -1.0)
        m2 = d.real ** 2 + d.imag ** 2
        e2 += m2.sum()
        emax = torch.maximum(emax, m2.max().sqrt().reshape(1))
    e2 = _allreduce(e2, world)
    emax = _allreduce(emax, world, dist.ReduceOp.MAX if world > 1 else None)
    out["rel_err_impulse"] = float(torch.sqrt(e2 / n).item())  # ||exact||^2 = N
    out["max_abs_err_impulse"] = float(emax.item())
    del x, y

    # ---- (iii) uniform noise, seed 5: random bins against f64 direct sums over all N points --------------------
    g = torch.Generator(device=dev).manual_seed(5 + rank)
    x = torch.view_as_complex(torch.rand((shard, 2), generator=g, device=dev) * 2 - 1).contiguous()
    keep = x.clone()  # transform() may reuse its input buffers
    y = run(x)
    gb = torch.Generator().manual_seed(5)
    bins = torch.randint(0, n, (nbins,), generator=gb).tolist()
    sums = torch.zeros(nbins, dtype=torch.complex128, device=dev)
    for c0 in range(0, shard, chunk):
        idx = torch.arange(lo + c0, lo + min(c0 + chunk, shard), dtype=torch.int64, device=dev)
        xc = keep[c0:c0 + idx.numel()].to(torch.complex128)
        for j, k in enumerate(bins):
            sums[j] += (xc * _phase(idx, k, n, -1.0)).sum()
    sr = _allreduce(torch.view_as_real(sums).clone(), world)
    exact = torch.view_as_complex(sr)
    mine = torch.zeros(nbins, dtype=torch.complex128, device=dev)
    for j, k in enumerate(bins):
        if lo <= k < lo + shard:
            mine[j] = y[k - lo].to(torch.complex128)
    mr = _allreduce(torch.view_as_real(mine).clone(), world)
    got = torch.view_as_complex(mr)
    rms = float(torch.sqrt((exact.real ** 2 + exact.imag ** 2).mean()).item())
    err = (got - exact).abs()
    out["max_bin_err"] = float(err.max().item() / rms)  # relative to the rms bin magnitude (~ sqrt(2N/3))
    out["rel_l2_bins"] = float(torch.sqrt((err ** 2).sum() / (exact.abs() ** 2).sum()).item())
    out["bins_checked"] = nbins
    out["log2n"] = log2n
    return out
