"""The real kernel bodies (CtaFft::run incl. the TMA/mbarrier staging protocol, and the large-N
ColPass::run / RowPass::run) executed on the CPU by the cooperative block emulator
(tests/emu/cuda_emu.h: CUDA threads are coroutines, __syncthreads a counting barrier) and
compared bit-for-bit with the oracle.  Persistent grids smaller than the work exercise the
row-group loops, buffer alternation and barrier phases exactly as on the GPU."""
import numpy as np
import pytest

from tests.conftest import rel_l2, uniform_c64

TOL = 1e-5


@pytest.mark.parametrize("L", [5, 7, 8, 9, 11, 12, 13, 14])
@pytest.mark.parametrize("staged", [False, True])
def test_cta_kernel_body(emuk, oracle, L, staged):
    n = 1 << L
    rng = np.random.default_rng(L)
    rows = 11 if L < 13 else 3
    x = uniform_c64(rng, (rows, n))
    tab = oracle.twiddles(n)
    for inverse in (False, True):
        ref = oracle.fft_batch(x, inverse=inverse)
        for grid in (1, 3):
            y = np.zeros_like(x)
            emuk.cta("c2c_inv" if inverse else "c2c_fwd", True, L, rows, tab, inp=x, out=y,
                     scale=float(np.float32(1) / np.float32(n)), staged=staged, grid=grid)
            assert np.array_equal(y, ref), (L, staged, inverse, grid)
    y = np.zeros_like(x)
    emuk.cta("c2c_fwd", False, L, rows, tab, inp=x, out=y, staged=staged, grid=2)
    assert rel_l2(y, oracle.fft_batch(x)) <= TOL


@pytest.mark.parametrize("staged", [False, True])
def test_cta_kernel_body_stft_rfft(emuk, oracle, staged):
    rng = np.random.default_rng(3)
    win_len, hop, length, ch = 2048, 512, 8192, 2
    sig = rng.uniform(-1, 1, (ch, length)).astype(np.float32)
    w = oracle.hann(win_len)
    nframes = length // hop + 2  # even (TPC = 2), includes frames that run past the end
    ref = oracle.stft_batch(sig, w, hop, nframes)
    frames = np.zeros_like(ref)
    emuk.cta("stft", True, 11, ch * nframes, oracle.twiddles(win_len), inp=sig, out=frames, aux=w,
             p=(length, nframes, hop, 0), staged=staged, grid=3)
    assert np.array_equal(frames, ref)
    time = np.zeros((ch, nframes, win_len), np.float32)
    emuk.cta("istft", True, 11, ch * nframes, oracle.twiddles(win_len), inp=ref, out=time, aux=w,
             scale=float(np.float32(1) / np.float32(win_len)), staged=staged, grid=3)
    want = np.stack([oracle.fft_batch(ref[c], inverse=True).real * w for c in range(ch)]).astype(np.float32)
    assert np.array_equal(time, want)
    m = 4096
    x = rng.uniform(-1, 1, (5, 2 * m)).astype(np.float32)
    y = np.zeros((5, m + 1), np.complex64)
    emuk.cta("rfft", True, 12, 5, oracle.twiddles(m), inp=x, out=y, aux=oracle.rfft_twiddles(m), staged=staged, grid=2)
    assert np.array_equal(y, oracle.rfft_batch(x))


@pytest.mark.parametrize("L", [15, 16])
def test_large_two_pass_c2c(emuk, oracle, L):
    n = 1 << L
    rng = np.random.default_rng(L)
    rows = 3
    x = uniform_c64(rng, (rows, n))
    tab = oracle.twiddles(n)
    ref = oracle.fft_batch(x)
    y = np.zeros_like(x)
    emuk.large("c2c_fwd", True, L, rows, tab, inp=x, out=y, grid_col=5, grid_row=16)
    assert np.array_equal(y, ref)
    y[...] = 0
    emuk.large("c2c_inv", True, L, rows, tab, inp=x, out=y, scale=float(np.float32(1) / np.float32(n)),
               grid_col=7, grid_row=32)
    assert np.array_equal(y, oracle.fft_batch(x, inverse=True))
    y[...] = 0
    emuk.large("c2c_fwd", False, L, rows, tab, inp=x, out=y)
    assert rel_l2(y, ref) <= TOL
    # SoA through the generic policy
    re, im = np.ascontiguousarray(x.real), np.ascontiguousarray(x.imag)
    ore, oim = np.zeros_like(re), np.zeros_like(im)
    emuk.large("gen_fwd", True, L, rows, tab, inp=re, in2=im, out=ore, out2=oim, p=(1, n, 1, n))
    assert np.array_equal(ore, ref.real) and np.array_equal(oim, ref.imag)


@pytest.mark.parametrize("L", [15, 16])
def test_large_two_pass_rfft_irfft(emuk, oracle, L):
    """BASELINE configs[2] is rfft N = 2^16, i.e. a half-length core m = 2^15 with the Hermitian
    twist fused behind the row pass (mirrored k-blocks)."""
    m = 1 << L
    rng = np.random.default_rng(100 + L)
    rows = 2
    x = rng.uniform(-1, 1, (rows, 2 * m)).astype(np.float32)
    tab, rtw = oracle.twiddles(m), oracle.rfft_twiddles(m)
    ref = oracle.rfft_batch(x)
    y = np.zeros((rows, m + 1), np.complex64)
    emuk.large("rfft", True, L, rows, tab, inp=x, out=y, aux=rtw)
    assert np.array_equal(y, ref)
    z = np.zeros((rows, 2 * m), np.float32)
    emuk.large("irfft", True, L, rows, tab, inp=ref, out=z, aux=rtw, scale=float(np.float32(1) / np.float32(m)))
    assert np.array_equal(z, oracle.irfft_batch(ref, 2 * m))
    y[...] = 0
    emuk.large("rfft", False, L, rows, tab, inp=x, out=y, aux=rtw)
    assert rel_l2(y, ref) <= TOL


@pytest.mark.parametrize("L,hop,length,extra_out,run_frames,grid", [
    (9, 128, 3000, 7, 4, 3),       # several runs per channel, halo frames, tail beyond the last frame
    (11, 512, 9000, 0, 3, 2),      # BASELINE window/hop, out_len == signal length (frames clipped)
    (10, 1024, 5000, 3000, 2, 2),  # hop == N (no overlap), output longer than the covered range
    (9, 200, 2600, 100, 64, 1),    # hop does not divide N; one run per channel
    (12, 1024, 9000, 0, 2, 2),     # one frame per CTA group (TPC = 1)
    (11, 512, 30000, 0, 24, 2),    # long runs: the steady-state fast path (register window power, fixed offsets)
])
def test_fused_istft_kernel_body(emuk, oracle, L, hop, length, extra_out, run_frames, grid):
    """IstftFused::run (ifft + window + ordered overlap-add + normalisation in one kernel) is
    bit-identical to the reference's sequential istft / inverse_parallel, including the
    accumulate-into-output semantics and samples no frame reaches."""
    N = 1 << L
    rng = np.random.default_rng(L * hop)
    ch = 2
    sig = rng.uniform(-1, 1, (ch, length)).astype(np.float32)
    w = oracle.hann(N)
    nframes = -(-length // hop) + 1
    frames = oracle.stft_batch(sig, w, hop, nframes)
    out_len = length + extra_out
    base = rng.uniform(-1, 1, (ch, out_len)).astype(np.float32)
    tab = oracle.twiddles(N)
    for zero_uncovered in (0, 1):
        got = base.copy()
        norm = np.full_like(got, -1.0)
        emuk.istft_fused(True, L, frames, w, got, norm, hop, run_frames, zero_uncovered, tab, grid=grid)
        if zero_uncovered:
            want = np.stack([oracle.istft_parallel(frames[c], w, hop, base[c]) for c in range(ch)])
        else:
            want = np.stack([oracle.istft(frames[c], w, hop, base[c]) for c in range(ch)])
        assert np.array_equal(got, want), (zero_uncovered, np.flatnonzero(got != want)[:10])
        # the reference's `scratch`: summed window power per sample, in frame order
        nrm = np.zeros(out_len, np.float32)
        for f in range(nframes):
            for i in range(N):
                if f * hop + i < out_len:
                    nrm[f * hop + i] = np.float32(nrm[f * hop + i] + np.float32(w[i] * w[i]))
        assert np.array_equal(norm[0], nrm) and np.array_equal(norm[1], nrm)
    got = base.copy()
    emuk.istft_fused(False, L, frames, w, got, None, hop, run_frames, 0, tab, grid=grid)
    want = np.stack([oracle.istft(frames[c], w, hop, base[c]) for c in range(ch)])
    assert rel_l2(got, want) <= TOL


@pytest.mark.parametrize("L", [15, 16])
def test_large_fused_cluster_kernel_body(emuk, oracle, L):
    """LargeFused::run: thread-block clusters, cluster barrier, double-buffered scratch."""
    n = 1 << L
    rng = np.random.default_rng(7 + L)
    rows = 5
    x = uniform_c64(rng, (rows, n))
    tab = oracle.twiddles(n)
    y = np.zeros_like(x)
    emuk.large("c2c_fwd", True, L, rows, tab, inp=x, out=y, grid_col=2, fused=True)
    assert np.array_equal(y, oracle.fft_batch(x))
    xr = rng.uniform(-1, 1, (rows, 2 * n)).astype(np.float32)
    rtw = oracle.rfft_twiddles(n)
    yr = np.zeros((rows, n + 1), np.complex64)
    emuk.large("rfft", True, L, rows, tab, inp=xr, out=yr, aux=rtw, grid_col=2, fused=True)
    assert np.array_equal(yr, oracle.rfft_batch(xr))
    zr = np.zeros((rows, 2 * n), np.float32)
    emuk.large("irfft", True, L, rows, tab, inp=yr, out=zr, aux=rtw, scale=float(np.float32(1) / np.float32(n)),
               grid_col=3, fused=True)
    assert np.array_equal(zr, oracle.irfft_batch(yr, 2 * n))


@pytest.mark.parametrize("L,grid,rows,staged,skew", [(15, 16, 7, True, None), (16, 16, 4, False, (5, 7)), (15, 8, 5, True, (4, 25))])
def test_large_pipelined_kernel_body(emuk, oracle, L, grid, rows, staged, skew):
    """LargePipe::run: teams of 8 / 16 CTAs (run concurrently by the emulator), pass A of a team's next
    transform ahead of pass B of the current one, dependency flags, three rotating intermediate
    slots per team, pass A's tile staged in the idle exchange buffer with asynchronous copies.
    skew = (late_from, ratio): the team's CTAs from rank late_from on run `ratio` times slower, so the
    others run ahead as far as the flags allow (a single running counter per team fails this)."""
    n = 1 << L
    rng = np.random.default_rng(900 + L + grid)
    x = uniform_c64(rng, (rows, n))
    tab = oracle.twiddles(n)
    y = np.zeros_like(x)
    emuk.large("c2c_fwd", True, L, rows, tab, inp=x, out=y, grid_col=grid, pipe=True, staged=staged, skew=skew)
    assert np.array_equal(y, oracle.fft_batch(x))
    y[...] = 0
    emuk.large("c2c_inv", True, L, rows, tab, inp=x, out=y, scale=float(np.float32(1) / np.float32(n)),
               grid_col=grid, pipe=True, staged=staged, skew=skew)
    assert np.array_equal(y, oracle.fft_batch(x, inverse=True))
    xr = rng.uniform(-1, 1, (rows, 2 * n)).astype(np.float32)
    rtw = oracle.rfft_twiddles(n)
    yr = np.zeros((rows, n + 1), np.complex64)
    emuk.large("rfft", True, L, rows, tab, inp=xr, out=yr, aux=rtw, grid_col=grid, pipe=True, staged=staged, skew=skew)
    ref = oracle.rfft_batch(xr)
    assert np.array_equal(yr, ref)
    zr = np.zeros((rows, 2 * n), np.float32)
    emuk.large("irfft", True, L, rows, tab, inp=ref, out=zr, aux=rtw, scale=float(np.float32(1) / np.float32(n)),
               grid_col=grid, pipe=True, staged=staged, skew=skew)
    assert np.array_equal(zr, oracle.irfft_batch(ref, 2 * n))
    re, im = np.ascontiguousarray(x.real), np.ascontiguousarray(x.imag)
    ore, oim = np.zeros_like(re), np.zeros_like(im)
    emuk.large("gen_fwd", True, L, rows, tab, inp=re, in2=im, out=ore, out2=oim, p=(1, n, 1, n), grid_col=grid,
               pipe=True, staged=staged, skew=skew)
    want = oracle.fft_batch(x)
    assert np.array_equal(ore, want.real) and np.array_equal(oim, want.imag)


@pytest.mark.parametrize("L", [15, 16])
def test_large_two_pass_prefetching_variants(emuk, oracle, L):
    """ColPass::run<true> (cp.async column tiles) + RowPass::run<true> (TMA bulk row tiles, mirrored
    half staged ascending and read back reversed for the rfft twist) are bit-identical too."""
    n = 1 << L
    rng = np.random.default_rng(300 + L)
    rows = 3
    x = uniform_c64(rng, (rows, n))
    tab = oracle.twiddles(n)
    y = np.zeros_like(x)
    emuk.large("c2c_fwd", True, L, rows, tab, inp=x, out=y, grid_col=5, grid_row=16, staged=True)
    assert np.array_equal(y, oracle.fft_batch(x))
    y[...] = 0
    emuk.large("c2c_inv", True, L, rows, tab, inp=x, out=y, scale=float(np.float32(1) / np.float32(n)),
               grid_col=3, grid_row=32, staged=True)
    assert np.array_equal(y, oracle.fft_batch(x, inverse=True))
    xr = rng.uniform(-1, 1, (rows, 2 * n)).astype(np.float32)
    rtw = oracle.rfft_twiddles(n)
    yr = np.zeros((rows, n + 1), np.complex64)
    emuk.large("rfft", True, L, rows, tab, inp=xr, out=yr, aux=rtw, staged=True)
    ref = oracle.rfft_batch(xr)
    assert np.array_equal(yr, ref)
    zr = np.zeros((rows, 2 * n), np.float32)  # irfft: column pass has no raw rows -> plain loads, row pass staged
    emuk.large("irfft", True, L, rows, tab, inp=ref, out=zr, aux=rtw, scale=float(np.float32(1) / np.float32(n)), staged=True)
    assert np.array_equal(zr, oracle.irfft_batch(ref, 2 * n))


@pytest.mark.parametrize("staged", [False, True])
def test_cta_kernel_body_stft_magnitudes(emuk, oracle, staged):
    """IoStftMag: |X[k]|, k < win_len/2, and the running maximum fused behind the last FFT stage
    (src/visual/spectrogram.rs:52-76) -- bit-identical magnitudes, exact maximum."""
    rng = np.random.default_rng(8)
    win_len, hop, length = 2048, 512, 9216
    sig = rng.uniform(-1, 1, length).astype(np.float32)
    want, want_max = oracle.stft_magnitudes(sig, win_len, hop)
    nframes = want.shape[0]
    mags = np.zeros_like(want)
    mx = np.zeros(1, np.int32)
    emuk.cta("stft_mag", True, 11, nframes, oracle.twiddles(win_len), inp=sig, out=mags, out2=mx, aux=oracle.hann(win_len),
             p=(length, nframes, hop, 0), staged=staged, grid=3)
    assert np.array_equal(mags, want)
    assert mx.view(np.float32)[0] == want_max


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 256, 1024, 4096, 8192])
def test_f64_kernel_body(emuk, oracle, n):
    """CtaFftD::run (fft_f64.cuh; N <= 16: the literal kernels instantiated for double2) against the
    f64 oracle, bit for bit, forward and inverse, persistent grids smaller than the work; staged = the
    next row group prefetched into the idle exchange buffer by a TMA bulk copy (ragged last group)."""
    rng = np.random.default_rng(64 + n)
    rows = 7 if n <= 1024 else 3
    x = (rng.uniform(-1, 1, (rows, n)) + 1j * rng.uniform(-1, 1, (rows, n))).astype(np.complex128)
    tab = oracle.twiddles_f64(n) if n >= 32 else None
    for inverse in (False, True):
        ref = oracle.fft_batch_f64(x, inverse=inverse)
        for grid, staged in ((1, False), (2, True), (1, True)):
            y = np.zeros_like(x)
            emuk.f64(n, rows, x, y, tab, inverse=inverse, grid=grid, staged=staged)
            assert np.array_equal(y, ref), (n, inverse, grid, staged)


@pytest.mark.parametrize("n", [8, 64, 2048])
def test_f64_split_rows_kernel_body(emuk, oracle, n):
    """IoGenericD (split / strided addressing of the f64 engine) against the f64 oracle."""
    rng = np.random.default_rng(128 + n)
    rows = 5
    x = (rng.uniform(-1, 1, (rows, n)) + 1j * rng.uniform(-1, 1, (rows, n))).astype(np.complex128)
    re, im = np.ascontiguousarray(x.real), np.ascontiguousarray(x.imag)
    tab = oracle.twiddles_f64(n) if n >= 32 else None
    for inverse in (False, True):
        ref = oracle.fft_batch_f64(x, inverse=inverse)
        ore, oim = np.zeros_like(re), np.zeros_like(im)
        emuk.f64_split(n, rows, re, im, ore, oim, tab, inverse=inverse)
        assert np.array_equal(ore, ref.real) and np.array_equal(oim, ref.imag), (n, inverse)


@pytest.mark.parametrize("n", [2, 8, 32, 64, 512, 4096, 16384])
def test_f64_rfft_irfft_kernel_body(emuk, oracle, n):
    """IoRfftD (twist behind one more exchange; staged variant) and IoIrfftD (untwist at the load)
    against the f64 oracle, bit for bit."""
    m = n // 2
    rng = np.random.default_rng(256 + n)
    rows = 5 if n <= 4096 else 2
    x = rng.uniform(-1, 1, (rows, n))
    tab = oracle.twiddles_f64(m) if m >= 32 else None
    rtw = oracle.rfft_twiddles_f64(m)
    ref = oracle.rfft_batch_f64(x)
    for staged in ((False, True) if m >= 32 else (False,)):
        y = np.zeros((rows, m + 1), np.complex128)
        emuk.f64_real(m, rows, x, y, tab, rtw, 1, grid=2, staged=staged)
        assert np.array_equal(y, ref), (n, staged)
    z = np.zeros((rows, n), np.float64)
    emuk.f64_real(m, rows, ref, z, tab, rtw, 2, grid=1)
    assert np.array_equal(z, oracle.irfft_batch_f64(ref, n)), n


@pytest.mark.parametrize("L,staged", [(13, False), (13, True), (14, False), (14, True)])
def test_wide_cta_shared_memory_bank_conflicts(emuk, L, staged):
    """every exchange access of a half-warp of the wide kernel -- padded or in-place (swapped) first layout, XOR-swizzled
    second layout, the rfft side buffer -- hits 16 distinct 8-byte bank pairs (enumerated with the kernel's own index
    functions)"""
    assert emuk.wide_bank_audit(L, staged) == 0


@pytest.mark.parametrize("L,staged", [(13, False), (13, True), (14, False), (14, True)])
def test_wide_cta_kernel_body(emuk, oracle, L, staged):
    """WideCta::run (fft_wide.cuh): N = 8192 / 16384 in one CTA with 32 elements per thread (three register passes, two
    exchanges in one buffer with two paddings); more rows than CTAs so the buffer is reused.  staged: the row lands by
    TMA bulk copies in the exchange buffer, pass 0 in place (warp-level swap at 8192).  Bit-exact both ways."""
    import functools

    emuk_wide = functools.partial(emuk.wide, staged=staged)
    n = 1 << L
    rng = np.random.default_rng(1700 + L)
    rows = 5
    x = uniform_c64(rng, (rows, n))
    tab = oracle.twiddles(n)
    inv_n = float(np.float32(1) / np.float32(n))
    y = np.zeros_like(x)
    emuk_wide("c2c_fwd", True, L, rows, tab, inp=x, out=y, grid=2)
    assert np.array_equal(y, oracle.fft_batch(x))
    y[...] = 0
    emuk_wide("c2c_inv", True, L, rows, tab, inp=x, out=y, scale=inv_n, grid=2)
    assert np.array_equal(y, oracle.fft_batch(x, inverse=True))
    y[...] = 0
    emuk_wide("c2c_fwd", False, L, rows, tab, inp=x, out=y, grid=3)
    assert rel_l2(y, oracle.fft_batch(x)) <= TOL
    # rfft (twist behind one more exchange), irfft (untwist at the load), SoA rows
    xr = rng.uniform(-1, 1, (rows, 2 * n)).astype(np.float32)
    rtw = oracle.rfft_twiddles(n)
    yr = np.zeros((rows, n + 1), np.complex64)
    emuk_wide("rfft", True, L, rows, tab, inp=xr, out=yr, aux=rtw, grid=2)
    ref = oracle.rfft_batch(xr)
    assert np.array_equal(yr, ref)
    if not staged:
        zr = np.zeros((rows, 2 * n), np.float32)
        emuk_wide("irfft", True, L, rows, tab, inp=ref, out=zr, aux=rtw, scale=inv_n, grid=2)
        assert np.array_equal(zr, oracle.irfft_batch(ref, 2 * n))
        re, im = np.ascontiguousarray(x.real), np.ascontiguousarray(x.imag)
        ore, oim = np.zeros_like(re), np.zeros_like(im)
        emuk_wide("gen_fwd", True, L, rows, tab, inp=re, in2=im, out=ore, out2=oim, p=(1, n, 1, n), grid=2)
        want = oracle.fft_batch(x)
        assert np.array_equal(ore, want.real) and np.array_equal(oim, want.imag)


@pytest.mark.parametrize("L,grid,rows,skew,staged", [(15, 8, 11, None, True), (15, 4, 6, (2, 9), True), (15, 8, 9, (1, 4), False),
                                                     (14, 4, 7, None, False), (14, 2, 6, (1, 5), False), (13, 2, 5, None, False)])
def test_split32_kernel_body(emuk, oracle, L, grid, rows, skew, staged):
    """Split32::run (fft_split32.cuh): warp-specialised CTAs (A warps: 2^(L-5)-point column transforms of an
    8192-element tile with one exchange; B warps: 32-point rows in one thread's registers, shuffle twist), teams
    of 4 / 2 / 1 CTAs, per-slot dependency counters between the roles; more transforms than 2 * SLOTS per team so
    the intermediate ring wraps.  skew: part of the team runs late.  staged: pass A's tiles arrive by TMA tensor-map
    loads in the exchange buffer (C2C and rfft; irfft / SoA rows keep the plain loads)."""
    import functools

    emuk_split32 = functools.partial(emuk.split32, staged=staged)
    n = 1 << L
    rng = np.random.default_rng(1300 + L + grid)
    x = uniform_c64(rng, (rows, n))
    tab = oracle.twiddles(n)
    y = np.zeros_like(x)
    emuk_split32("c2c_fwd", True, L, rows, tab, inp=x, out=y, grid=grid, skew=skew)
    assert np.array_equal(y, oracle.fft_batch(x))
    y[...] = 0
    emuk_split32("c2c_inv", True, L, rows, tab, inp=x, out=y, scale=float(np.float32(1) / np.float32(n)), grid=grid, skew=skew)
    assert np.array_equal(y, oracle.fft_batch(x, inverse=True))
    xr = rng.uniform(-1, 1, (rows, 2 * n)).astype(np.float32)
    rtw = oracle.rfft_twiddles(n)
    yr = np.zeros((rows, n + 1), np.complex64)
    emuk_split32("rfft", True, L, rows, tab, inp=xr, out=yr, aux=rtw, grid=grid, skew=skew)
    ref = oracle.rfft_batch(xr)
    assert np.array_equal(yr, ref)
    zr = np.zeros((rows, 2 * n), np.float32)
    emuk_split32("irfft", True, L, rows, tab, inp=ref, out=zr, aux=rtw, scale=float(np.float32(1) / np.float32(n)), grid=grid, skew=skew)
    assert np.array_equal(zr, oracle.irfft_batch(ref, 2 * n))
    yr[...] = 0
    emuk_split32("rfft", False, L, rows, tab, inp=xr, out=yr, aux=rtw, grid=grid)
    assert rel_l2(yr, ref) <= TOL
    if L == 15 and skew is None:
        re, im = np.ascontiguousarray(x.real), np.ascontiguousarray(x.imag)
        ore, oim = np.zeros_like(re), np.zeros_like(im)
        emuk_split32("gen_fwd", True, L, rows, tab, inp=re, in2=im, out=ore, out2=oim, p=(1, n, 1, n), grid=grid)
        want = oracle.fft_batch(x)
        assert np.array_equal(ore, want.real) and np.array_equal(oim, want.imag)
