"""CPU emulation of the CUDA CTA (tests/emu) against the oracle: index math, twiddle
selection, padding and every fused I/O policy, bit-for-bit in EXACT mode and within the
north-star tolerance (rel-L2 <= 1e-5) in FAST mode.  This is the GPU-less stand-in for the
-m gpu parity tests; it compiles the very same headers the kernels are built from."""
import numpy as np
import pytest

from tests.conftest import rel_l2, uniform_c64

SIZES = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384]
TOL = 1e-5  # BASELINE.json north_star: relative L2 vs kofft's f32 path


def table_for(oracle, n):
    return oracle.twiddles(n) if n > 16 else np.zeros(1, np.complex64)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("inverse", [False, True])
def test_c2c(emu, oracle, n, inverse):
    rng = np.random.default_rng(n)
    rows = 5 if n <= 4096 else 2  # ragged w.r.t. transforms-per-CTA on purpose
    x = uniform_c64(rng, (rows, n))
    ref = oracle.fft_batch(x, inverse=inverse) if n > 1 else x.copy()
    scale = float(np.float32(1.0) / np.float32(n))
    for exact in (True, False):
        for staged in (False, True):  # staged = the TMA bulk-copy input path
            y = np.zeros_like(x)
            emu.run("c2c_inv" if inverse else "c2c_fwd", exact, n, rows, table_for(oracle, n), inp=x, out=y,
                    scale=scale, staged=staged)
            if exact:
                assert np.array_equal(y, ref)
            else:
                assert rel_l2(y, ref) <= TOL


@pytest.mark.parametrize("n", [4, 32, 256, 2048])
def test_reference_input_patterns(emu, oracle, n):
    # large-dynamic-range inputs the reference's tests/benches use: (i, 0), (i, 2i), (i, -i/2)
    i = np.arange(n, dtype=np.float32)
    for x in [(i + 0j), (i + 2j * i), (i - 0.5j * i), (np.sin(i) + 1j * np.cos(i))]:
        x = x.astype(np.complex64).reshape(1, n)
        y = np.zeros_like(x)
        emu.run("c2c_fwd", True, n, 1, table_for(oracle, n), inp=x, out=y)
        assert np.array_equal(y, oracle.fft_batch(x))


@pytest.mark.parametrize("n", [8, 64, 1024])
@pytest.mark.parametrize("inverse", [False, True])
def test_split_and_strided(emu, oracle, n, inverse):
    rng = np.random.default_rng(7 * n)
    rows = 3
    x = uniform_c64(rng, (rows, n))
    ref = oracle.fft_batch(x, inverse=inverse)
    scale = float(np.float32(1.0) / np.float32(n))
    kind = "gen_inv" if inverse else "gen_fwd"
    # SoA
    re, im = np.ascontiguousarray(x.real), np.ascontiguousarray(x.imag)
    ore, oim = np.zeros_like(re), np.zeros_like(im)
    emu.run(kind, True, n, rows, table_for(oracle, n), inp=re, in2=im, out=ore, out2=oim, p=(1, n, 1, n), scale=scale)
    assert np.array_equal(ore, ref.real) and np.array_equal(oim, ref.imag)
    # interleaved with stride 3 in, stride 2 out
    buf = np.zeros((rows, n * 3), np.complex64)
    buf[:, ::3] = x
    outb = np.full((rows, n * 2), 7 + 7j, np.complex64)
    fin, fout = buf.view(np.float32), outb.view(np.float32)
    emu.run(kind, True, n, rows, table_for(oracle, n), inp=fin, in2=fin.ravel()[1:], out=fout, out2=fout.ravel()[1:],
            p=(6, 6 * n, 4, 4 * n), scale=scale)
    assert np.array_equal(outb[:, ::2], ref)
    assert np.all(outb[:, 1::2] == 7 + 7j)  # elements between the strides are untouched


@pytest.mark.parametrize("m", [1, 2, 8, 16, 32, 128, 4096, 16384])
def test_rfft_irfft(emu, oracle, m):
    rng = np.random.default_rng(m)
    n, rows = 2 * m, (3 if m <= 4096 else 1)
    x = rng.uniform(-1, 1, (rows, n)).astype(np.float32)
    ref = oracle.rfft_batch(x)
    rtw = oracle.rfft_twiddles(m)
    for exact in (True, False):
        for staged in (False, True):
            y = np.zeros((rows, m + 1), np.complex64)
            emu.run("rfft", exact, m, rows, table_for(oracle, m), inp=x, out=y, aux=rtw, staged=staged)
            assert np.array_equal(y, ref) if exact else rel_l2(y, ref) <= TOL
    back_ref = oracle.irfft_batch(ref, n)
    for exact in (True, False):
        z = np.zeros((rows, n), np.float32)
        emu.run("irfft", exact, m, rows, table_for(oracle, m), inp=ref, out=z, aux=rtw,
                scale=float(np.float32(1.0) / np.float32(m)))
        assert np.array_equal(z, back_ref) if exact else rel_l2(z, back_ref) <= TOL


@pytest.mark.parametrize("win_len,hop,length,staged", [
    (4, 2, 8, False), (16, 4, 50, False), (64, 16, 1000, False), (2048, 512, 5000, False), (256, 300, 700, False),
    (2048, 512, 5120, True), (1024, 256, 6144, True), (64, 16, 1024, True), (256, 64, 4096, True), (4096, 1024, 20480, True)])
def test_stft_and_istft_stage1(emu, oracle, win_len, hop, length, staged):
    rng = np.random.default_rng(win_len + hop)
    ch = 2
    sig = rng.uniform(-1, 1, (ch, length)).astype(np.float32)
    w = oracle.hann(win_len)
    nframes = -(-length // hop) + 1  # one frame more than required: the reference fills it too
    if staged:  # what the host requires before it picks the TMA-staged kernel (kofft_cuda_stft_f32)
        tpc = max(1, 4096 // win_len)
        nframes = -(-nframes // tpc) * tpc + tpc  # multiple of TPC, incl. frames wholly past the end
        assert length % 4 == 0 and hop % 4 == 0 and sig.ctypes.data % 16 == 0
    ref = oracle.stft_batch(sig, w, hop, nframes)
    frames = np.zeros((ch, nframes, win_len), np.complex64)
    emu.run("stft", True, win_len, ch * nframes, table_for(oracle, win_len), inp=sig, out=frames, aux=w,
            p=(length, nframes, hop, 0), staged=staged)
    assert np.array_equal(frames, ref)
    # istft stage 1: (ifft(frame).re) * window
    time = np.zeros((ch, nframes, win_len), np.float32)
    emu.run("istft", True, win_len, ch * nframes, table_for(oracle, win_len), inp=ref, out=time, aux=w,
            scale=float(np.float32(1.0) / np.float32(win_len)), staged=staged)
    exp = np.stack([oracle.fft_batch(ref[c], inverse=True).real * w for c in range(ch)])
    assert np.array_equal(time, exp.astype(np.float32))
    # ordered overlap-add of those frames == the oracle's istft (pure numpy restatement of ola.cu)
    for c in range(ch):
        out_len = length + 3
        acc = rng.uniform(-1, 1, out_len).astype(np.float32)  # istft accumulates into `output`
        want = oracle.istft(ref[c], w, hop, acc)
        got = acc.copy()
        nrm = np.zeros(out_len, np.float32)
        for f in range(nframes):
            for i in range(win_len):
                p = f * hop + i
                if p < out_len:
                    got[p] = np.float32(got[p] + time[c, f, i])
                    nrm[p] = np.float32(nrm[p] + np.float32(w[i] * w[i]))
        ok = nrm > np.float32(1e-8)
        got[ok] = got[ok] / nrm[ok]
        assert np.array_equal(got, want)


@pytest.mark.parametrize("L", range(5, 15))
def test_shared_memory_bank_conflicts(emu, L):
    """every exchange store/load of a half-warp hits 16 distinct 8-byte bank pairs"""
    assert all(v in (0, 1) for v in emu.bank_audit(L)), emu.bank_audit(L)
