"""Sharded single transform (BASELINE configs[4]).  There is no kofft output to compare with at
these sizes: the checks are against f64 (numpy) on closed-form and random inputs.

CPU part: a numpy model of the phase/index algebra that `kofft_cuda_dist_phase` implements
(transpose_scatter semantics + four-step split), for world = 1, 2, 4, 8 -- no GPU code runs.
GPU part (-m gpu): the real kernels through the C ABI with every rank on one device
(`connect_local`), and across devices / processes when the box has more than one GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- numpy model of the kernels' contract ---------------------------------------------------------
def model_scatter(src, world, cb, dsts, rank, twiddle, log2n):
    """D_d[c, off + r] = S[r, d*cb + c] * W_N^{(row0 + r) * (d*cb + c)}   (dist_kernels.cu)"""
    rows = src.shape[0]
    off = row0 = rows * rank
    n = 1 << log2n
    for d in range(world):
        blk = src[:, d * cb:(d + 1) * cb]
        if twiddle:
            r = (row0 + np.arange(rows))[:, None]
            c = (d * cb + np.arange(cb))[None, :]
            w = np.exp(-2j * np.pi * ((r * c) % n) / n)
            blk = blk * (np.conj(w) if twiddle == 2 else w)
        dsts[d][:, off:off + rows] = blk.T


def model_transform(x, world, log2n, inverse=False, natural=True):
    l1 = log2n // 2
    n1, n2 = 1 << l1, 1 << (log2n - l1)
    r1, c2 = n1 // world, n2 // world
    f = (lambda a: np.fft.ifft(a, axis=1)) if inverse else (lambda a: np.fft.fft(a, axis=1))
    shards = [x[g * r1 * n2:(g + 1) * r1 * n2].reshape(r1, n2) for g in range(world)]
    A = [np.zeros((c2, n1), complex) for _ in range(world)]
    B = [np.zeros((r1, n2), complex) for _ in range(world)]
    for g in range(world):  # phase 0
        model_scatter(shards[g], world, c2, A, g, 0, log2n)
    for g in range(world):  # phase 1
        A[g] = f(A[g])
    for g in range(world):
        model_scatter(A[g], world, r1, B, g, 2 if inverse else 1, log2n)
    for g in range(world):  # phase 2
        B[g] = f(B[g])
    if not natural:
        return B
    for g in range(world):
        model_scatter(B[g], world, c2, A, g, 0, log2n)
    return [a.reshape(-1) for a in A]


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("log2n", [10, 13])
def test_model_matches_numpy_fft(world, log2n):
    if (log2n // 2) - int(np.log2(world)) < 2:
        pytest.skip("too few rows per rank for this toy size")
    rng = np.random.default_rng(log2n * 10 + world)
    n = 1 << log2n
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    want = np.fft.fft(x)
    got = np.concatenate(model_transform(x, world, log2n))
    assert np.allclose(got, want, rtol=0, atol=1e-9 * np.abs(want).max())
    # transposed order: rank g, [r][k2] = X[(g r1 + r) + n1 k2]
    n1 = 1 << (log2n // 2)
    B = model_transform(x, world, log2n, natural=False)
    r1 = n1 // world
    for g in range(world):
        k = (g * r1 + np.arange(r1))[:, None] + n1 * np.arange(n >> (log2n // 2))[None, :]
        assert np.allclose(B[g], want[k], rtol=0, atol=1e-9 * np.abs(want).max())
    back = np.concatenate(model_transform(want, world, log2n, inverse=True))
    assert np.allclose(back, x, rtol=0, atol=1e-9)


# ---- the real thing ----------------------------------------------------------------------------
def _rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _run_local(log2n, world, devices, x, inverse=False, natural=True):
    import torch

    import kofft_b200
    from kofft_b200 import dist as D

    ctxs = {d: kofft_b200.Context(device=d) for d in set(devices)}
    dists = [D.DistFft(ctxs[devices[g]], g, world, log2n) for g in range(world)]
    shard = (1 << log2n) // world
    xs = [torch.from_numpy(x[g * shard:(g + 1) * shard]).to(f"cuda:{devices[g]}") for g in range(world)]
    outs = [torch.empty_like(t) for t in xs]
    D.run_local(dists, xs, outs, inverse=inverse, natural_order=natural)
    res = [o.cpu().numpy() for o in outs]
    for d in dists:
        d.close()
    return res


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2, 4])
@pytest.mark.parametrize("log2n", [16, 20, 23])
def test_dist_single_device_virtual_ranks(world, log2n):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    rng = np.random.default_rng(log2n + world)
    n = 1 << log2n
    x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
    want = np.fft.fft(x.astype(np.complex128))
    got = np.concatenate(_run_local(log2n, world, [0] * world, x))
    assert _rel(got, want) < 2e-6, _rel(got, want)
    # transposed-order output
    n1 = 1 << (log2n // 2)
    r1 = n1 // world
    parts = _run_local(log2n, world, [0] * world, x, natural=False)
    for g in range(world):
        k = (g * r1 + np.arange(r1))[:, None] + n1 * np.arange(n // n1)[None, :]
        assert _rel(parts[g].reshape(r1, -1), want[k]) < 2e-6
    # inverse round trip
    back = np.concatenate(_run_local(log2n, world, [0] * world, got.astype(np.complex64), inverse=True))
    assert _rel(back, x.astype(np.complex128)) < 2e-6


@pytest.mark.gpu
def test_dist_closed_form_tones_and_impulse():
    """K tones give exact spikes, an impulse gives an exact phase ramp (SURVEY.md 8d config 5 inputs)."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    log2n, world = 24, 4
    n = 1 << log2n
    idx = np.arange(n)
    bins, amps = [3, 4097, n // 2 + 17, n - 5], [1.0, 0.5, 0.25, 2.0]
    x = np.zeros(n, np.complex128)
    for b, a in zip(bins, amps):
        x += a * np.exp(2j * np.pi * ((b * idx) % n) / n)
    got = np.concatenate(_run_local(log2n, world, [0] * world, x.astype(np.complex64)))
    for b, a in zip(bins, amps):
        assert abs(got[b] / n - a) < 1e-5
    mask = np.ones(n, bool)
    mask[bins] = False
    assert np.abs(got[mask]).max() / n < 1e-5
    n0 = 123457
    imp = np.zeros(n, np.complex64)
    imp[n0] = 1
    got = np.concatenate(_run_local(log2n, world, [0] * world, imp))
    k = np.array([0, 1, 2, 77, n // 3, n - 1])
    assert np.allclose(got[k], np.exp(-2j * np.pi * ((k * n0) % n) / n), atol=2e-6)


@pytest.mark.gpu
def test_dist_across_devices_one_process():
    import torch

    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2+ GPUs")
    world = 2
    while world * 2 <= min(torch.cuda.device_count(), 8):
        world *= 2
    log2n = 24
    rng = np.random.default_rng(7)
    n = 1 << log2n
    x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
    want = np.fft.fft(x.astype(np.complex128))
    got = np.concatenate(_run_local(log2n, world, list(range(world)), x))
    assert _rel(got, want) < 2e-6


@pytest.mark.gpu
def test_dist_one_process_per_gpu_ipc():
    """torchrun, one rank per GPU: IPC handle exchange + barriers over the process group."""
    import torch

    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2+ GPUs")
    world = 2
    while world * 2 <= min(torch.cuda.device_count(), 8):
        world *= 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "dist_worker.py"), "24"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "dist_worker ok" in r.stdout


@pytest.mark.gpu
def test_dist_collective_arm_and_validation_single_rank():
    """world = 1 runs the pack / (no) all-to-all / unpack arm and the closed-form + direct-sum validation that
    bench.py reports at 2^30, here at 2^22 on one GPU; the two arms must agree bit for bit."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import kofft_b200
    from kofft_b200 import dist as D
    from kofft_b200 import dist_validate

    log2n = 22
    ctx = kofft_b200.Context(device=0)
    d = D.DistFft(ctx, 0, 1, log2n)
    z = torch.zeros(1 << log2n, dtype=torch.complex64, device="cuda")
    D.run_local([d], [z], [torch.zeros_like(z)])  # connects the single rank
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.view_as_complex(torch.rand((1 << log2n, 2), generator=g, device="cuda") * 2 - 1).contiguous()
    a = d.transform(x).clone()
    b = d.transform_collective(x)
    assert torch.equal(torch.view_as_real(a), torch.view_as_real(b))
    v = dist_validate.validate(d, chunk=1 << 20, nbins=16)
    assert v["rel_err_tones"] < 2e-6 and v["rel_err_impulse"] < 2e-6 and v["max_bin_err"] < 5e-6, v
    d.close()
