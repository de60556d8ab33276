"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI of
libkofft_cuda.so via the host-side mirror of the reference interface (kofft_b200), and is
compared with the CPU oracle on the same seeded inputs:

  * EXACT mode (default): bit-identical to the oracle.
  * FAST mode: relative L2 <= 1e-5 (the north-star tolerance, BASELINE.json).

The first block restates the reference's own tests against `CudaFftImpl` so they read like
the originals (file:line cited per test)."""
import numpy as np
import pytest

from tests.conftest import rel_l2, uniform_c64

pytestmark = pytest.mark.gpu
TOL = 1e-5  # BASELINE.json north_star tolerance: relative L2 vs kofft's f32 output


def c32(*vals):
    return np.array(vals, dtype=np.complex64)


# ------------------------------------------------------------------------------------------
# 1. the reference's tests, restated against the GPU backend
# ------------------------------------------------------------------------------------------
def test_fft_ifft_f32(cuda_fft):  # src/lib.rs:178-199
    data = c32(1, 0, 0, 0)
    cuda_fft.fft(data)
    assert np.all(np.abs(data.real - 1) < 1e-6) and np.all(np.abs(data.imag) < 1e-6)
    cuda_fft.ifft(data)
    assert abs(data[0].real - 1) < 1e-6 and np.all(np.abs(data[1:]) < 1e-6)


def test_fft_all_zeros_all_ones(cuda_fft):  # src/lib.rs:242-264
    data = np.zeros(8, np.complex64)
    cuda_fft.fft(data)
    assert np.all(np.abs(data) < 1e-6)
    data = np.ones(8, np.complex64)
    cuda_fft.fft(data)
    assert abs(data[0].real - 8) < 1e-6 and np.all(np.abs(data[1:]) < 1e-6)


def test_fft_symmetries(cuda_fft):  # src/lib.rs:360-388
    d = c32(1, 2, 3, 4)
    cuda_fft.fft(d)
    assert abs(d[1].real - d[3].real) < 1e-6 and abs(d[1].imag + d[3].imag) < 1e-6
    d = c32(1j, 2j, 3j, 4j)
    cuda_fft.fft(d)
    assert abs(d[1].real + d[3].real) < 1e-6 and abs(d[1].imag - d[3].imag) < 1e-6


def test_fft_empty_single_nonpow2(cuda_fft):  # src/lib.rs:313-318, 342-349
    import kofft_b200 as k

    with pytest.raises(k.EmptyInput):
        cuda_fft.fft(np.zeros(0, np.complex64))
    with pytest.raises(k.EmptyInput):
        cuda_fft.ifft(np.zeros(0, np.complex64))
    d = c32(1)
    cuda_fft.fft(d)
    assert d[0] == 1
    d = c32(1, 2, 3)  # non-power-of-two: Bluestein, like the reference's std build (src/fft.rs:1083-1132)
    cuda_fft.fft(d)
    assert np.allclose(d, np.fft.fft([1, 2, 3]), atol=1e-5)
    # the rfft / stft cores reach Bluestein too (they call fft.fft(): src/rfft.rs:447, src/stft.rs:102), the
    # device-resident streams included; only the fused magnitude kernel is power-of-two only
    assert cuda_fft.rfft_batch(np.ones((1, 24), np.float32)).shape == (1, 13)
    from kofft_b200 import stft as S

    S.DeviceStftStream(cuda_fft, 1, np.ones(12, np.float32), 4)


def test_fft_out_of_place(cuda_fft, oracle):  # src/lib.rs:281-311, 320-329
    import kofft_b200 as k

    inp = c32(1, 2, 3, 4)
    out = np.zeros(4, np.complex64)
    cuda_fft.fft_out_of_place(inp, out)
    assert np.array_equal(inp, c32(1, 2, 3, 4)) and np.array_equal(out, oracle.fft(inp))
    assert np.array_equal(cuda_fft.fft_vec(inp), oracle.fft(inp))
    with pytest.raises(k.MismatchedLengths):
        cuda_fft.fft_out_of_place(c32(1, 2), np.zeros(3, np.complex64))
    cuda_fft.ifft_out_of_place(out.copy(), out)
    assert np.all(np.abs(out - inp) < 1e-6)


def test_roundtrip_repeated_and_large_values(cuda_fft):  # src/lib.rs:390-429
    d = c32(1000, 2000, 3000, 4000)
    orig = d.copy()
    cuda_fft.fft(d)
    cuda_fft.ifft(d)
    assert np.all(np.abs(d - orig) < 1e-3)
    d = c32(1, 2, 3, 4)
    for _ in range(10):
        cuda_fft.fft(d)
        cuda_fft.ifft(d)
    assert np.all(np.abs(d - c32(1, 2, 3, 4)) < 1e-4)


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32])
def test_fft_matches_dft_for_powers_of_two(cuda_fft, n):  # tests/pow2.rs:18-31
    i = np.arange(n, dtype=np.float32)
    data = (i - 0.5j * i).astype(np.complex64)
    expected = np.fft.fft(data.astype(np.complex128))
    cuda_fft.fft(data)
    assert np.all(np.abs(data.real - expected.real) < 1e-2) and np.all(np.abs(data.imag - expected.imag) < 1e-2)


@pytest.mark.parametrize("n", [32, 64, 128, 256, 512, 1024])
def test_stockham_matches_fft(cuda_fft, oracle, n):  # tests/stockham_parity.rs, stockham_large.rs
    import kofft_b200 as k

    i = np.arange(n, dtype=np.float32)
    src = ((i - 0.25j * i) if n <= 256 else (np.sin(i) + 1j * np.cos(i))).astype(np.complex64)
    data, expected = src.copy(), src.copy()
    cuda_fft.fft(expected)
    cuda_fft.fft_with_strategy(data, k.FftStrategy.SplitRadix)  # == stockham_fft (src/fft.rs:1360)
    assert np.array_equal(data, expected) and np.array_equal(data, oracle.fft(src))


def test_parallel_stockham_4096(cuda_fft, oracle):  # tests/parallel_stockham.rs:6-27
    import kofft_b200 as k

    i = np.arange(4096, dtype=np.float32)
    a = (i + 2j * i).astype(np.complex64)
    b = a.copy()
    k.fft_parallel(a)
    cuda_fft.fft(b)
    assert np.all(np.abs(a - b) < 1e-4) and np.array_equal(b, oracle.fft((i + 2j * i).astype(np.complex64)))


def test_split(cuda_fft):  # tests/split.rs:11-27, 48-78
    import kofft_b200 as k

    data = np.arange(16, dtype=np.float32).astype(np.complex64)
    re, im = data.real.copy(), data.imag.copy()
    aos = data.copy()
    cuda_fft.fft(aos)
    k.fft_split(re, im)
    assert np.all(np.abs(aos.real - re) < 1e-6) and np.all(np.abs(aos.imag - im) < 1e-6)
    i = np.arange(64, dtype=np.float32)
    re, im = i.copy(), -i
    cuda_fft.fft_split(re, im)
    cuda_fft.ifft_split(re, im)
    assert np.all(np.abs(re - i) < 1e-4) and np.all(np.abs(im + i) < 1e-4)
    with pytest.raises(k.MismatchedLengths):
        cuda_fft.fft_split(np.zeros(4, np.float32), np.zeros(3, np.float32))


def test_planner_device_table(cuda_fft, oracle):  # tests/twiddle.rs:7-13 + "device-resident twiddle tables"
    import torch

    p = cuda_fft.planner
    t = p.get_twiddles(8)
    e = np.exp(-2j * np.pi / 8)
    assert abs(t[1].real - e.real) < 1e-6 and abs(t[1].imag - e.imag) < 1e-6
    ptr1, ptr2 = p.device_twiddles(4096), p.device_twiddles(4096)
    assert ptr1 == ptr2 and ptr1 != 0  # pointer-stable like the reference's Arc
    # read the device copy back and compare with the oracle's table
    from cuda import cudart

    host = np.empty(2048, np.complex64)
    (err,) = cudart.cudaMemcpy(host.ctypes.data, ptr1, host.nbytes, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
    assert int(err) == 0
    assert np.array_equal(host, oracle.twiddles(4096))


def test_rfft_roundtrip_and_errors(cuda_fft, oracle):  # src/lib.rs:431-478, src/rfft.rs:892-906
    import kofft_b200 as k

    for vals in ([1, 2, 3, 4], [1, 2, 3, 4, 5, 6, 7, 8]):
        inp = np.array(vals, np.float32)
        freq = np.zeros(len(inp) // 2 + 1, np.complex64)
        scratch = np.zeros(len(inp) // 2, np.complex64)
        cuda_fft.rfft_with_scratch(inp, freq, scratch)
        assert abs(freq[0].imag) < 1e-6 and abs(freq[-1].imag) < 1e-6
        assert np.array_equal(freq, oracle.rfft(inp))
        out = np.zeros(len(inp), np.float32)
        cuda_fft.irfft_with_scratch(freq, out, scratch)
        assert np.all(np.abs(out - inp) < 1e-5)
    with pytest.raises(k.MismatchedLengths):
        cuda_fft.rfft(np.array([1, 2, 3, 4], np.float32), np.zeros(4, np.complex64))
    with pytest.raises(k.InvalidValue):
        cuda_fft.rfft(np.array([1, 2, 3], np.float32), np.zeros(2, np.complex64))
    with pytest.raises(k.EmptyInput):
        cuda_fft.rfft(np.zeros(0, np.float32), np.zeros(1, np.complex64))
    planner = k.RfftPlanner()  # tests/rfft_dispatch.rs:5-24
    inp = np.array([1, 2, 3, 4], np.float32)
    freq, scratch, out = np.zeros(3, np.complex64), np.zeros(2, np.complex64), np.zeros(4, np.float32)
    planner.rfft_with_scratch(cuda_fft, inp, freq, scratch)
    planner.irfft_with_scratch(cuda_fft, freq, out, scratch)
    assert np.all(np.abs(out - inp) < 1e-5)


def test_rfft_sin32(cuda_fft, oracle):  # tests/rfft_arch_parity.rs:11-46
    x = np.sin(np.arange(32, dtype=np.float32))
    f = np.zeros(17, np.complex64)
    cuda_fft.rfft(x, f)
    assert np.array_equal(f, oracle.rfft(x))


def test_stft_errors(cuda_fft):  # tests/stft.rs:6-14, src/stft.rs:83-89
    import kofft_b200 as k
    from kofft_b200 import stft as S
    from kofft_b200.window import hann

    with pytest.raises(k.MismatchedLengths):
        S.stft(np.zeros(10, np.float32), hann(4), 4, [None] * 2, cuda_fft)
    with pytest.raises(k.InvalidHopSize):
        S.stft(np.zeros(10, np.float32), hann(4), 0, [None] * 3, cuda_fft)
    with pytest.raises(k.InvalidHopSize):
        S.istft([], hann(4), 0, np.zeros(4, np.float32), np.zeros(4, np.float32), cuda_fft)
    with pytest.raises(k.MismatchedLengths):
        S.istft([np.zeros(4, np.complex64)], hann(4), 2, np.zeros(4, np.float32), np.zeros(5, np.float32), cuda_fft)
    with pytest.raises(k.MismatchedLengths):
        S.istft([np.zeros(3, np.complex64)], hann(4), 2, np.zeros(4, np.float32), np.zeros(4, np.float32), cuda_fft)


def test_stft_istft_roundtrip(cuda_fft):  # src/stft.rs:527-630, 800-813
    from kofft_b200 import stft as S

    signal = np.arange(1, 9, dtype=np.float32)
    window = np.ones(4, np.float32)
    frames = [None] * 4
    S.stft(signal, window, 2, frames, cuda_fft)
    out, scratch = np.zeros(8, np.float32), np.zeros(8, np.float32)
    S.istft(frames, window, 2, out, scratch, cuda_fft)
    assert np.all(np.abs(out - signal) < 1e-4)


def test_istft_stream_reconstructs_and_flushes(cuda_fft):  # tests/istft_stream.rs:5-49 (assert_eq!)
    from kofft_b200 import stft as S

    signal = np.arange(1, 9, dtype=np.float32)
    win_len, hop = 4, 2
    window = np.ones(win_len, np.float32)
    st = S.StftStream(signal, window, hop, cuda_fft)
    ist = S.IstftStream(win_len, hop, window.copy(), cuda_fft)
    frame = np.zeros(win_len, np.complex64)
    frames, stream_out = [], []
    while st.next_frame(frame):
        frames.append(frame.copy())
        stream_out.extend(ist.push_frame(frame).tolist())
    tail = ist.flush().copy()
    stream_out.extend(tail.tolist())
    offline = np.zeros(len(signal) + win_len - hop, np.float32)
    S.istft([f.copy() for f in frames], window, hop, offline, np.zeros_like(offline), cuda_fft)
    assert np.array_equal(np.array(stream_out[: len(signal)], np.float32), offline[: len(signal)])
    assert len(tail) == win_len - hop and np.array_equal(tail, offline[len(signal):])


def test_zero_window(cuda_fft):  # src/stft.rs:699-720
    from kofft_b200 import stft as S

    signal = np.arange(1, 9, dtype=np.float32)
    window = np.zeros(4, np.float32)
    frames = [None] * 4
    S.stft(signal, window, 2, frames, cuda_fft)
    assert not np.stack(frames).any()
    out = np.zeros(8, np.float32)
    S.istft(frames, window, 2, out, np.zeros(8, np.float32), cuda_fft)
    assert not out.any()


def test_parallel_equals_serial(cuda_fft):  # src/visual/spectrogram.rs:281-297 (bit-exact)
    from kofft_b200 import stft as S
    from kofft_b200.window import hann

    signal = np.arange(16, dtype=np.float32)
    a, b = [None] * 8, [None] * 8
    S.stft(signal, hann(4), 2, a, cuda_fft)
    S.parallel(signal, hann(4), 2, b, cuda_fft)
    assert np.array_equal(np.stack(a), np.stack(b))


# ------------------------------------------------------------------------------------------
# 2. parity with the oracle on seeded inputs, every size
# ------------------------------------------------------------------------------------------
SIZES = [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384]


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("inverse", [False, True])
def test_c2c_batch_bit_exact(cuda_fft, cuda_fft_fast, oracle, n, inverse):
    rng = np.random.default_rng(n + 100 * inverse)
    rows = 37 if n <= 4096 else 5  # ragged vs transforms-per-CTA and vs the grid
    x = uniform_c64(rng, (rows, n))
    ref = oracle.fft_batch(x, inverse=inverse, nthreads=4)
    y = x.copy()
    cuda_fft.fft_batch(y, inverse=inverse)
    assert np.array_equal(y, ref), f"exact mode differs: rel {rel_l2(y, ref):.3e}"
    z = x.copy()
    cuda_fft_fast.fft_batch(z, inverse=inverse)
    assert rel_l2(z, ref) <= TOL
    worst = max(rel_l2(z[r], ref[r]) for r in range(rows))
    assert worst <= TOL


def test_config1_basic_usage(cuda_fft, oracle):
    """BASELINE configs[0]: 1024-point FFT + IFFT of (sin(0.1 i), 0) (examples/basic_usage.rs:232-241)."""
    i = np.arange(1024, dtype=np.float32)
    x = np.sin(np.float32(0.1) * i).astype(np.float32).astype(np.complex64)
    y = x.copy()
    cuda_fft.fft(y)
    assert np.array_equal(y, oracle.fft(x))
    cuda_fft.ifft(y)
    assert np.array_equal(y, oracle.ifft(oracle.fft(x)))
    assert np.max(np.abs(y - x)) < 1e-4


def test_reference_dynamic_range_inputs(cuda_fft, oracle):
    # kofft-bench uses (i, 0) (bench_fft.rs:109); tests use (i, 2i): large dynamic range
    for n in (1024, 4096):
        i = np.arange(n, dtype=np.float32)
        for x in ((i + 0j), (i + 2j * i)):
            x = x.astype(np.complex64)
            y = x.copy()
            cuda_fft.fft(y)
            assert np.array_equal(y, oracle.fft(x))


def test_device_tensor_path(cuda_fft, oracle):
    import torch

    rng = np.random.default_rng(5)
    x = uniform_c64(rng, (300, 2048))
    d = torch.from_numpy(x).cuda()
    out = torch.empty_like(d)
    cuda_fft.fft_batch(d, out=out)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), oracle.fft_batch(x, nthreads=4))
    assert np.array_equal(d.cpu().numpy(), x)  # out-of-place leaves the input alone
    cuda_fft.fft_batch(d, inverse=True)  # in place
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy(), oracle.fft_batch(x, inverse=True, nthreads=4))


@pytest.mark.parametrize("n", [1, 2, 16, 64, 1024])
def test_strided_and_out_of_place_strided(cuda_fft, oracle, n):
    import kofft_b200 as k

    rng = np.random.default_rng(n)
    stride = 3
    buf = uniform_c64(rng, ((n - 1) * stride + 1 + 2,))
    for inverse in (False, True):
        a = buf.copy()
        scratch = np.zeros(n, np.complex64)
        (cuda_fft.ifft_strided if inverse else cuda_fft.fft_strided)(a, stride, scratch)
        assert np.array_equal(a, oracle.fft_strided(buf, stride, n, inverse))
    a = buf.copy()
    cuda_fft.fft_strided_alloc(a[: n * stride], stride)
    inp = uniform_c64(rng, (n * 2,))
    out = np.full(n * 4, 9 - 9j, np.complex64)
    cuda_fft.fft_out_of_place_strided(inp, 2, out, 4)
    assert np.array_equal(out, oracle.fft_out_of_place_strided(inp, 2, np.full(n * 4, 9 - 9j, np.complex64), 4))
    with pytest.raises(k.InvalidStride):
        cuda_fft.fft_strided(buf.copy(), 0, np.zeros(n, np.complex64))
    with pytest.raises(k.InvalidStride):
        cuda_fft.fft_out_of_place_strided(inp, 0, out, 4)
    if n > 2:
        with pytest.raises(k.MismatchedLengths):
            cuda_fft.fft_strided(buf[: n].copy(), stride, np.zeros(n, np.complex64))


@pytest.mark.parametrize("m", [1, 2, 4, 16, 32, 256, 2048, 16384])
def test_rfft_irfft_batch(cuda_fft, cuda_fft_fast, oracle, m):
    rng = np.random.default_rng(m)
    n, rows = 2 * m, (9 if m <= 2048 else 3)
    x = rng.uniform(-1, 1, (rows, n)).astype(np.float32)
    ref = oracle.rfft_batch(x, nthreads=4)
    y = cuda_fft.rfft_batch(x)
    assert np.array_equal(y, ref), rel_l2(y, ref)
    assert rel_l2(cuda_fft_fast.rfft_batch(x), ref) <= TOL
    back = oracle.irfft_batch(ref, n, nthreads=4)
    assert np.array_equal(cuda_fft.irfft_batch(ref, n), back)
    assert rel_l2(cuda_fft_fast.irfft_batch(ref, n), back) <= TOL


def test_rfft_fma_table_flavour(oracle):
    """xtask's `-C target-feature=+fma` build fuses Complex::mul in the rfft table recurrence."""
    import kofft_b200 as k

    fft = k.CudaFftImpl(device=0, exact=True)
    planner = k.RfftPlanner(fma_mul=True)
    x = np.random.default_rng(3).uniform(-1, 1, 4096).astype(np.float32)
    f = np.zeros(2049, np.complex64)
    planner.rfft(fft, x, f)
    assert np.array_equal(f, oracle.rfft(x, fma_mul=True))
    assert not np.array_equal(f, oracle.rfft(x, fma_mul=False))


@pytest.mark.parametrize("win_len,hop,length,kind", [(4, 2, 8, "hann"), (16, 4, 50, "hamming"), (64, 16, 1000, "blackman"),
                                                     (2048, 512, 40000, "hann"), (256, 300, 700, "kaiser"),
                                                     (1024, 256, 9999, "hann")])
def test_stft_istft_batch(cuda_fft, cuda_fft_fast, oracle, win_len, hop, length, kind):
    from kofft_b200 import stft as S
    from kofft_b200 import window as W

    rng = np.random.default_rng(win_len + hop)
    ch = 3
    sig = rng.uniform(-1, 1, (ch, length)).astype(np.float32)
    w = {"hann": W.hann, "hamming": W.hamming, "blackman": W.blackman,
         "kaiser": lambda n: W.kaiser(n, 8.6)}[kind](win_len)
    nframes = -(-length // hop) + 2  # more frames than required: the reference fills them all
    ref = oracle.stft_batch(sig, w, hop, nframes, nthreads=4)
    frames = S.stft_batch(cuda_fft, sig, w, hop, nframes)
    assert np.array_equal(frames, ref)
    assert rel_l2(S.stft_batch(cuda_fft_fast, sig, w, hop, nframes), ref) <= TOL
    out_len = length + 5
    base = rng.uniform(-1, 1, (ch, out_len)).astype(np.float32)  # istft ACCUMULATES into output
    want = np.stack([oracle.istft(ref[c], w, hop, base[c]) for c in range(ch)])
    got = base.copy()
    norm = np.zeros_like(got)
    S.istft_batch(cuda_fft, ref, w, hop, got, norm)
    assert np.array_equal(got, want)
    want0 = np.stack([oracle.istft_parallel(ref[c], w, hop, base[c]) for c in range(ch)])
    got0 = base.copy()
    S.istft_batch(cuda_fft, ref, w, hop, got0, None, zero_uncovered=True)
    assert np.array_equal(got0, want0)
    fast = base.copy()
    S.istft_batch(cuda_fft_fast, ref, w, hop, fast, np.zeros_like(fast))
    assert rel_l2(fast, want) <= TOL


@pytest.mark.parametrize("win_len,hop,length,run_frames", [(2048, 512, 60000, 8), (512, 128, 5000, 64), (4096, 1024, 50000, 3),
                                                         (1024, 1024, 9000, 5), (512, 200, 4000, 4)])
def test_istft_fused_and_two_kernel_paths_agree(cuda_fft, oracle, win_len, hop, length, run_frames):
    """istft runs as ONE fused kernel for windows 512..4096; the two-kernel path is the fallback.
    Both bit-identical to the oracle, incl. accumulate-into-output, norm, uncovered samples."""
    from kofft_b200 import stft as S
    from kofft_b200 import window as W

    rng = np.random.default_rng(win_len + hop)
    ch = 3
    sig = rng.uniform(-1, 1, (ch, length)).astype(np.float32)
    w = W.hann(win_len)
    nframes = -(-length // hop) + 1
    frames = oracle.stft_batch(sig, w, hop, nframes, nthreads=4)
    out_len = length + 700
    base = rng.uniform(-1, 1, (ch, out_len)).astype(np.float32)
    want = np.stack([oracle.istft(frames[c], w, hop, base[c]) for c in range(ch)])
    want0 = np.stack([oracle.istft_parallel(frames[c], w, hop, base[c]) for c in range(ch)])
    try:
        for fused in (True, False):
            cuda_fft.ctx.set_istft_fusion(fused, run_frames)
            got, norm = base.copy(), np.zeros_like(base)
            S.istft_batch(cuda_fft, frames, w, hop, got, norm)
            assert np.array_equal(got, want), f"fused={fused}"
            got0 = base.copy()
            S.istft_batch(cuda_fft, frames, w, hop, got0, None, zero_uncovered=True)
            assert np.array_equal(got0, want0), f"fused={fused}"
            if fused:
                norm_fused = norm
            else:
                assert np.array_equal(norm, norm_fused)
    finally:
        cuda_fft.ctx.set_istft_fusion(True, 64)


def test_stft_device_tensors(cuda_fft, oracle):
    import torch
    from kofft_b200 import stft as S
    from kofft_b200 import window as W

    rng = np.random.default_rng(11)
    sig = rng.uniform(-1, 1, (4, 48000)).astype(np.float32)
    w = W.hann(2048)
    nframes = -(-48000 // 512)
    d_sig, d_w = torch.from_numpy(sig).cuda(), torch.from_numpy(w).cuda()
    frames = S.stft_batch(cuda_fft, d_sig, d_w, 512, nframes)
    torch.cuda.synchronize()
    ref = oracle.stft_batch(sig, w, 512, nframes, nthreads=4)
    assert np.array_equal(frames.cpu().numpy(), ref)
    out = torch.zeros((4, 48000), device="cuda")
    norm = torch.zeros_like(out)
    S.istft_batch(cuda_fft, frames, d_w, 512, out, norm)
    torch.cuda.synchronize()
    want = np.stack([oracle.istft(ref[c], w, 512, np.zeros(48000, np.float32)) for c in range(4)])
    assert np.array_equal(out.cpu().numpy(), want)
    # reconstruction away from the edges (sample 0 has w = 0; full overlap from win_len on)
    assert np.max(np.abs(want[:, 2048:-2048] - sig[:, 2048:-2048])) < 2e-3


def test_batch_free_functions_ragged(cuda_fft, oracle):  # src/fft.rs:2156-2191
    import kofft_b200 as k

    rng = np.random.default_rng(9)
    rows = [uniform_c64(rng, (n,)) for n in (64, 64, 64, 8, 1024, 1024, 1)]
    want = [oracle.fft(r) if len(r) > 1 else r.copy() for r in rows]
    k.batch(cuda_fft, rows)
    for a, b in zip(rows, want):
        assert np.array_equal(a, b)
    k.batch_inverse(cuda_fft, rows)
    k.multi_channel(cuda_fft, rows)
    for a, b in zip(rows, want):  # fft(ifft(fft(x))): kofft's own round-trip error is ~3e-5 at N=1024
        assert rel_l2(a, b) < 1e-3


# ------------------------------------------------------------------------------------------
# 3. BASELINE.json full sizes: sampled rows against the oracle + size-independent properties
# ------------------------------------------------------------------------------------------
def test_config2_full_size_c2c_4096x65536(cuda_fft, cuda_fft_fast, oracle):
    import torch

    n, batch = 4096, 65536
    g = torch.Generator(device="cuda").manual_seed(0)
    x = (torch.rand((batch, n, 2), generator=g, device="cuda") * 2 - 1)
    x = torch.view_as_complex(x).contiguous()
    y = torch.empty_like(x)
    cuda_fft.fft_batch(x, out=y)
    torch.cuda.synchronize()
    rows = np.random.default_rng(0).choice(batch, 96, replace=False)
    rows = np.concatenate([rows, [0, 1, batch - 1]])
    xs = x[torch.from_numpy(rows).cuda()].cpu().numpy()
    ys = y[torch.from_numpy(rows).cuda()].cpu().numpy()
    ref = oracle.fft_batch(xs, nthreads=8)
    assert np.array_equal(ys, ref)  # bit-exact on every sampled row
    # no worse than the reference's own error against an f64 DFT (north star, second criterion)
    f64 = np.fft.fft(xs[:16].astype(np.complex128), axis=1)
    assert rel_l2(ys[:16], f64) <= rel_l2(ref[:16], f64) * (1 + 1e-6)
    # FAST mode on the whole batch: rel-L2 vs EXACT over everything, and worst row
    z = torch.empty_like(x)
    cuda_fft_fast.fft_batch(x, out=z)
    torch.cuda.synchronize()
    num = torch.linalg.vector_norm((z - y).view(batch, -1), dim=1)
    den = torch.linalg.vector_norm(y.view(batch, -1), dim=1)
    assert float((num / den).max()) <= TOL
    del z
    # Parseval per row (size-independent): sum|Y|^2 = N sum|x|^2 up to kofft's table error
    ex = torch.linalg.vector_norm(x, dim=1) ** 2
    ey = torch.linalg.vector_norm(y, dim=1) ** 2
    assert float(((ey / (n * ex)) - 1).abs().max()) < 2e-3
    # linearity: F(a x1 + x2) = a F(x1) + F(x2) on the first 4096 rows (f32 round-off bound)
    a = 0.5
    lin_in = (a * x[:4096] + x[4096:8192]).contiguous()
    lin = torch.empty_like(lin_in)
    cuda_fft.fft_batch(lin_in, out=lin)
    torch.cuda.synchronize()
    comb = a * y[:4096] + y[4096:8192]
    assert float(torch.linalg.vector_norm(lin - comb) / torch.linalg.vector_norm(comb)) < 1e-6
    # round trip through the inverse, in place, whole batch
    cuda_fft.fft_batch(y, inverse=True)
    torch.cuda.synchronize()
    err = torch.linalg.vector_norm((y - x).view(batch, -1), dim=1) / torch.linalg.vector_norm(x.view(batch, -1), dim=1)
    assert float(err.max()) < 5e-4  # kofft's own fft->ifft round trip at N=4096 is ~1.7e-4
    back = oracle.fft_batch(ref, inverse=True, nthreads=8)
    assert np.array_equal(y[torch.from_numpy(rows).cuda()].cpu().numpy(), back)


def test_config4_stft_full_shape_sampled(cuda_fft, oracle):
    """BASELINE configs[3] shape (Hann 2048, hop 512, 48 kHz) on as many channels as fit
    comfortably: sampled frames bit-exact vs the oracle, ISTFT round trip."""
    import torch
    from kofft_b200 import stft as S
    from kofft_b200 import window as W

    free, _ = torch.cuda.mem_get_info()
    ch = 64 if free > 100e9 else 8
    length = 28_800_000 if free > 100e9 else 2_880_000
    hop, win_len = 512, 2048
    nframes = -(-length // hop)
    g = torch.Generator(device="cuda").manual_seed(2)
    t = torch.arange(length, device="cuda", dtype=torch.float32) / 48000.0
    sig = (0.5 * torch.sin(2 * np.pi * 440.0 * t) + 0.3 * torch.sin(2 * np.pi * 1000.0 * t)
           + 0.1 * torch.sin(2 * np.pi * 5000.0 * t))
    sig = (sig.unsqueeze(0) + 0.1 * (torch.rand((ch, length), generator=g, device="cuda") * 2 - 1)).contiguous()
    w = W.hann(win_len)
    d_w = torch.from_numpy(w).cuda()
    frames = S.stft_batch(cuda_fft, sig, d_w, hop, nframes)
    torch.cuda.synchronize()
    rng = np.random.default_rng(4)
    for c in rng.choice(ch, 3, replace=False):
        fs = np.concatenate([rng.choice(nframes, 24, replace=False), [0, 1, nframes - 4, nframes - 1]])
        host_sig = sig[c].cpu().numpy()
        for f in fs:
            seg = np.zeros(win_len, np.float32)
            part = host_sig[f * hop: f * hop + win_len]
            seg[: len(part)] = part
            want = oracle.fft((seg * w).astype(np.complex64))
            assert np.array_equal(frames[c, f].cpu().numpy(), want), (c, f)
    out = torch.zeros((ch, length), device="cuda")
    S.istft_batch(cuda_fft, frames, d_w, hop, out)
    torch.cuda.synchronize()
    # fully overlapped interior reconstructs the input (Hann^2 / sum Hann^2 at 75 % overlap)
    err = (out[:, win_len:-win_len] - sig[:, win_len:-win_len]).abs().max()
    assert float(err) < 5e-3
    c = int(rng.integers(ch))
    n_chk = 20 * hop + win_len
    want = oracle.istft(frames[c, : 20 + 4].cpu().numpy(), w, hop, np.zeros(n_chk, np.float32))
    assert np.array_equal(out[c, : 20 * hop].cpu().numpy(), want[: 20 * hop])


# ------------------------------------------------------------------------------------------
# 4. kernel variants: TMA-staged input prefetch vs plain loads, unaligned fallbacks
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [32, 256, 2048, 4096, 8192])
def test_tma_staged_and_plain_variants_agree(cuda_fft, oracle, n):
    import torch

    rng = np.random.default_rng(n)
    rows = 1000 if n <= 4096 else 77
    x = uniform_c64(rng, (rows + 1, n))
    ref = oracle.fft_batch(x[:rows], nthreads=8)
    d = torch.from_numpy(x).cuda()
    try:
        for tma in (True, False):
            cuda_fft.ctx.set_tma_staging(tma)
            out = torch.zeros((rows, n), dtype=torch.complex64, device="cuda")
            cuda_fft.fft_batch(d[:rows], out=out)
            torch.cuda.synchronize()
            assert np.array_equal(out.cpu().numpy(), ref), f"tma={tma}"
        # an input that is only 8-byte aligned must silently take the plain-load kernel
        cuda_fft.ctx.set_tma_staging(True)
        flat = d.view(-1)
        shifted = flat[1: 1 + rows * n].view(rows, n)
        assert shifted.data_ptr() % 16 == 8
        out = torch.zeros((rows, n), dtype=torch.complex64, device="cuda")
        cuda_fft.fft_batch(shifted, out=out)
        torch.cuda.synchronize()
        want = oracle.fft_batch(x.reshape(-1)[1: 1 + rows * n].reshape(rows, n), nthreads=8)
        assert np.array_equal(out.cpu().numpy(), want)
    finally:
        cuda_fft.ctx.set_tma_staging(True)


def test_stream_ordering_on_a_side_stream(cuda_fft, oracle):
    """Device entry points are ordered on the caller's stream (here a non-default torch stream)."""
    import torch

    rng = np.random.default_rng(1)
    x = uniform_c64(rng, (512, 1024))
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        d = torch.from_numpy(x).cuda(non_blocking=True)
        out = torch.empty_like(d)
        for _ in range(3):
            cuda_fft.fft_batch(d, out=out)
            cuda_fft.fft_batch(out, inverse=True, out=d)
        cuda_fft.fft_batch(d, out=out)
    side.synchronize()
    want = x
    for _ in range(3):
        want = oracle.fft_batch(oracle.fft_batch(want, nthreads=8), inverse=True, nthreads=8)
    assert np.array_equal(out.cpu().numpy(), oracle.fft_batch(want, nthreads=8))


# ------------------------------------------------------------------------------------------
# 5. large-N two-pass path (N = 2^15, 2^16; rfft N = 2^16, 2^17)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [32768, 65536])
def test_large_c2c(cuda_fft, cuda_fft_fast, oracle, n):
    rng = np.random.default_rng(n)
    rows = 7
    x = uniform_c64(rng, (rows, n))
    for inverse in (False, True):
        ref = oracle.fft_batch(x, inverse=inverse, nthreads=8)
        y = x.copy()
        cuda_fft.fft_batch(y, inverse=inverse)
        assert np.array_equal(y, ref), rel_l2(y, ref)
        z = x.copy()
        cuda_fft_fast.fft_batch(z, inverse=inverse)
        assert rel_l2(z, ref) <= TOL
    one = x[0].copy()
    cuda_fft.fft(one)  # trait-level single transform
    assert np.array_equal(one, oracle.fft(x[0]))
    re, im = np.ascontiguousarray(x[1].real), np.ascontiguousarray(x[1].imag)
    cuda_fft.fft_split(re, im)
    ref1 = oracle.fft(x[1])
    assert np.array_equal(re, ref1.real) and np.array_equal(im, ref1.imag)


@pytest.mark.parametrize("n", [32768, 65536])
def test_large_paths_agree(cuda_fft, oracle, n):
    """N > 16384 has three implementations of the two-pass split: the persistent pipelined kernel
    (teams of CTAs, dependency flags; the default for rfft), two kernels per batch chunk (the default
    otherwise) and the thread-block-cluster kernel.  All must be bit-identical to the oracle (enough
    rows that every team / cluster iterates several times and reuses its intermediate slots)."""
    rng = np.random.default_rng(n + 1)
    rows = 150
    x = uniform_c64(rng, (rows, n))
    ref = oracle.fft_batch(x, nthreads=8)
    iref = oracle.fft_batch(x, inverse=True, nthreads=8)
    xr = rng.uniform(-1, 1, (rows, 2 * n)).astype(np.float32)
    rref = oracle.rfft_batch(xr, nthreads=8)
    bref = oracle.irfft_batch(rref, 2 * n, nthreads=8)
    C = cuda_fft.ctx
    try:
        C.set_split_min_log2n(16)  # 2^15 would otherwise take the split kernel whatever the mode (tested below)
        for mode in (C.LARGE_PIPELINED, C.LARGE_CLUSTER, C.LARGE_TWO_KERNEL):
            C.set_cluster_fusion(False)
            C.set_large_mode(mode)
            tag = f"mode={mode}"
            y = x.copy()
            cuda_fft.fft_batch(y)
            assert np.array_equal(y, ref), tag
            y = x.copy()
            cuda_fft.fft_batch(y, inverse=True)
            assert np.array_equal(y, iref), tag
            assert np.array_equal(cuda_fft.rfft_batch(xr), rref), tag
            assert np.array_equal(cuda_fft.irfft_batch(rref, 2 * n), bref), tag
    finally:
        C.set_large_mode(C.LARGE_AUTO)
        C.set_split_min_log2n(14)


@pytest.mark.parametrize("n", [8192, 16384, 32768])
def test_split_kernel_on_device(cuda_fft, cuda_fft_fast, oracle, n):
    """The warp-specialised split kernel (fft_split32.cuh; default for 2^15, selectable for 2^13 / 2^14): C2C
    both directions, rfft / irfft, split (SoA) rows; ~20 transforms per team so every intermediate slot is
    reused several times; grids of one team up to the full device; repeated launches (counters re-zeroed);
    no cooperative-launch fallback happened."""
    import torch

    rows = 800 if n == 32768 else 3000
    g = torch.Generator(device="cuda").manual_seed(n)
    x = torch.view_as_complex((torch.rand((rows, n, 2), generator=g, device="cuda") * 2 - 1).contiguous())
    xr = (torch.rand((rows, 2 * n), generator=g, device="cuda") * 2 - 1).contiguous()
    C = cuda_fft.ctx
    fb0 = C.fallback_count
    outs, routs = [], []
    try:
        C.set_split_min_log2n(13)
        C.set_split_all_kinds(True)
        C.set_wide_mask(0)  # dense C2C rows of 8192 / 16384 points would otherwise take the wide single-CTA kernel
        cuda_fft_fast.ctx.set_split_min_log2n(13)
        cuda_fft_fast.ctx.set_wide_mask(0)
        for max_ctas in (0, 4, 24, 140):
            C.set_max_ctas(max_ctas)
            y = torch.empty_like(x)
            for _ in range(3):
                cuda_fft.fft_batch(x, out=y)
            torch.cuda.synchronize()
            outs.append(y)
            routs.append(cuda_fft.rfft_batch(xr))
            torch.cuda.synchronize()
        C.set_max_ctas(0)
        pick = [0, 1, 2, 36, 37, 73, 74, 148, 149, rows // 2, rows - 2, rows - 1]
        xs, xrs = x[pick].cpu().numpy(), xr[pick].cpu().numpy()
        ref = oracle.fft_batch(xs, nthreads=8)
        rref = oracle.rfft_batch(xrs, nthreads=8)
        assert np.array_equal(outs[0][pick].cpu().numpy(), ref)
        assert np.array_equal(routs[0][pick].cpu().numpy(), rref)
        for y, yr in zip(outs[1:], routs[1:]):
            assert torch.equal(torch.view_as_real(y), torch.view_as_real(outs[0]))
            assert torch.equal(torch.view_as_real(yr), torch.view_as_real(routs[0]))
        # inverse, irfft, SoA rows, FAST mode on the sampled rows (host-pointer path)
        y = xs.copy()
        cuda_fft.fft_batch(y, inverse=True)
        assert np.array_equal(y, oracle.fft_batch(xs, inverse=True, nthreads=8))
        assert np.array_equal(cuda_fft.irfft_batch(rref, 2 * n), oracle.irfft_batch(rref, 2 * n, nthreads=8))
        re, im = np.ascontiguousarray(xs[3].real), np.ascontiguousarray(xs[3].imag)
        cuda_fft.fft_split(re, im)
        assert np.array_equal(re, ref[3].real) and np.array_equal(im, ref[3].imag)
        z = xs.copy()
        cuda_fft_fast.fft_batch(z)
        assert rel_l2(z, ref) <= TOL
        assert rel_l2(cuda_fft_fast.rfft_batch(xrs), rref) <= TOL
        assert C.fallback_count == fb0
    finally:
        C.set_max_ctas(0)
        C.set_split_min_log2n(14)
        C.set_split_all_kinds(False)
        C.set_wide_mask(None)
        cuda_fft_fast.ctx.set_split_min_log2n(14)
        cuda_fft_fast.ctx.set_wide_mask(None)


@pytest.mark.parametrize("rows", [1, 2, 3, 5, 38, 149])
def test_irfft_65536_few_rows_per_team(cuda_fft, oracle, rows):
    """irfft 2^16 through the split kernel (the B warps untwist three transforms ahead of pass A): fewer transforms per
    team than the look-ahead, teams without work, one more row than teams; bit-identical to the oracle."""
    import torch

    rng = np.random.default_rng(6500 + rows)
    spec = (rng.uniform(-1, 1, (rows, 32769)) + 1j * rng.uniform(-1, 1, (rows, 32769))).astype(np.complex64)
    back = cuda_fft.irfft_batch(torch.from_numpy(spec).cuda(), 65536)
    torch.cuda.synchronize()
    assert np.array_equal(back.cpu().numpy(), oracle.irfft_batch(spec, 65536, nthreads=8))
    assert cuda_fft.ctx.fallback_count == 0


@pytest.mark.parametrize("n", [8192, 16384])
def test_wide_kernel_on_device(cuda_fft, cuda_fft_fast, oracle, n):
    """The wide single-CTA kernel (fft_wide.cuh; default for dense C2C rows of 8192 / 16384 points): more rows than
    resident CTAs so every CTA reuses its buffer, grids from one CTA to the full device, both directions, the TMA-staged
    variant (16-byte aligned rows) and the plain-load variant (rows at an odd complex offset); identical to the
    split / single-CTA kernels on the whole batch and to the oracle on sampled rows; FAST mode inside the tolerance."""
    import torch

    rows = 700
    g = torch.Generator(device="cuda").manual_seed(n + 1)
    flat = (torch.rand((rows * n + 1, 2), generator=g, device="cuda") * 2 - 1).contiguous()
    x = torch.view_as_complex(flat[: rows * n]).view(rows, n)
    xo = torch.view_as_complex(flat[1:]).view(rows, n)  # 8-byte aligned only: plain loads
    C = cuda_fft.ctx
    try:
        outs = []
        for max_ctas in (0, 1, 37):
            C.set_max_ctas(max_ctas)
            y = torch.empty_like(x)
            for _ in range(2):
                cuda_fft.fft_batch(x, out=y)
            torch.cuda.synchronize()
            outs.append(y)
        C.set_max_ctas(0)
        yo = torch.empty((rows, n), dtype=torch.complex64, device="cuda")
        cuda_fft.fft_batch(xo, out=yo)
        yi = torch.empty_like(x)
        cuda_fft.fft_batch(x, out=yi, inverse=True)
        C.set_wide_mask(0)
        y_other = torch.empty_like(x)
        cuda_fft.fft_batch(x, out=y_other)
        yo_other = torch.empty_like(yo)
        cuda_fft.fft_batch(xo, out=yo_other)
        yi_other = torch.empty_like(x)
        cuda_fft.fft_batch(x, out=yi_other, inverse=True)
        torch.cuda.synchronize()
        for y in outs:
            assert torch.equal(torch.view_as_real(y), torch.view_as_real(y_other))
        assert torch.equal(torch.view_as_real(yo), torch.view_as_real(yo_other))
        assert torch.equal(torch.view_as_real(yi), torch.view_as_real(yi_other))
        pick = [0, 1, 147, 148, 295, 296, 297, rows - 1]
        xs = x[pick].cpu().numpy()
        ref = oracle.fft_batch(xs, nthreads=8)
        assert np.array_equal(outs[0][pick].cpu().numpy(), ref)
        assert np.array_equal(yi[pick].cpu().numpy(), oracle.fft_batch(xs, inverse=True, nthreads=8))
        assert np.array_equal(yo[pick].cpu().numpy(), oracle.fft_batch(xo[pick].cpu().numpy(), nthreads=8))
        z = xs.copy()
        cuda_fft_fast.fft_batch(z)
        assert rel_l2(z, ref) <= TOL
        # the other kinds through the same kernel (mask 0xff): rfft with the twist behind one more exchange, irfft with
        # the untwist at the load, SoA rows; identical to the default paths and to the oracle
        xr = (torch.rand((rows, 2 * n), generator=g, device="cuda") * 2 - 1).contiguous()
        C.set_wide_mask(0xFF)
        yr = cuda_fft.rfft_batch(xr)
        zr = cuda_fft.irfft_batch(yr, 2 * n)
        re, im = np.ascontiguousarray(xs[3].real), np.ascontiguousarray(xs[3].imag)
        cuda_fft.fft_split(re, im)
        C.set_wide_mask(0)
        yr0 = cuda_fft.rfft_batch(xr)
        zr0 = cuda_fft.irfft_batch(yr, 2 * n)
        torch.cuda.synchronize()
        assert torch.equal(torch.view_as_real(yr), torch.view_as_real(yr0)) and torch.equal(zr, zr0)
        assert np.array_equal(yr[pick].cpu().numpy(), oracle.rfft_batch(xr[pick].cpu().numpy(), nthreads=8))
        assert np.array_equal(re, ref[3].real) and np.array_equal(im, ref[3].imag)
    finally:
        C.set_max_ctas(0)
        C.set_wide_mask(None)


def test_large_pipelined_many_transforms_per_team_on_device(cuda_fft, oracle):
    """The pipelined kernel on device-resident rows: ~40 transforms per team (slot reuse, both
    dependency counters), grids of one team up to the full device, repeated launches (the counters
    are re-zeroed); sampled rows bit-identical to the oracle, whole batch identical across grids."""
    import torch

    n, rows = 32768, 1500
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.view_as_complex((torch.rand((rows, n, 2), generator=g, device="cuda") * 2 - 1).contiguous())
    xr = (torch.rand((rows, 2 * n), generator=g, device="cuda") * 2 - 1).contiguous()
    C = cuda_fft.ctx
    outs, routs = [], []
    try:
        C.set_split_min_log2n(16)
        C.set_large_mode(C.LARGE_PIPELINED)
        for max_ctas in (0, 8, 24, 160):
            C.set_max_ctas(max_ctas)
            y = torch.empty_like(x)
            for _ in range(3):
                cuda_fft.fft_batch(x, out=y)
            torch.cuda.synchronize()
            outs.append(y)
            routs.append(cuda_fft.rfft_batch(xr))
            torch.cuda.synchronize()
    finally:
        C.set_max_ctas(0)
        C.set_large_mode(C.LARGE_AUTO)
        C.set_split_min_log2n(14)
    pick = [0, 1, 36, 37, 73, 74, 700, rows - 2, rows - 1]
    assert np.array_equal(outs[0][pick].cpu().numpy(), oracle.fft_batch(x[pick].cpu().numpy(), nthreads=8))
    assert np.array_equal(routs[0][pick].cpu().numpy(), oracle.rfft_batch(xr[pick].cpu().numpy(), nthreads=8))
    for y, yr in zip(outs[1:], routs[1:]):
        assert torch.equal(torch.view_as_real(y), torch.view_as_real(outs[0]))
        assert torch.equal(torch.view_as_real(yr), torch.view_as_real(routs[0]))


@pytest.mark.parametrize("n", [65536, 131072])
def test_large_rfft_irfft(cuda_fft, cuda_fft_fast, oracle, n):
    rng = np.random.default_rng(n)
    rows = 5
    x = rng.uniform(-1, 1, (rows, n)).astype(np.float32)
    ref = oracle.rfft_batch(x, nthreads=8)
    y = cuda_fft.rfft_batch(x)
    assert np.array_equal(y, ref), rel_l2(y, ref)
    assert rel_l2(cuda_fft_fast.rfft_batch(x), ref) <= TOL
    back = oracle.irfft_batch(ref, n, nthreads=8)
    assert np.array_equal(cuda_fft.irfft_batch(ref, n), back)
    # the reference's bench input x_i = i (kofft-bench/benches/bench_fft.rs:300)
    ramp = np.arange(n, dtype=np.float32).reshape(1, n)
    assert np.array_equal(cuda_fft.rfft_batch(ramp), oracle.rfft_batch(ramp))


def test_config3_full_size_rfft_65536x16384(cuda_fft, oracle):
    """BASELINE configs[2]: rfft N = 2^16 x batch 16384 (4 GiB in, 4 GiB out)."""
    import torch

    n, batch = 65536, 16384
    m = n // 2
    g = torch.Generator(device="cuda").manual_seed(1)
    x = (torch.rand((batch, n), generator=g, device="cuda") * 2 - 1).contiguous()
    y = cuda_fft.rfft_batch(x)
    torch.cuda.synchronize()
    rows = np.concatenate([np.random.default_rng(1).choice(batch, 24, replace=False), [0, batch - 1]])
    idx = torch.from_numpy(rows).cuda()
    xs, ys = x[idx].cpu().numpy(), y[idx].cpu().numpy()
    ref = oracle.rfft_batch(xs, nthreads=8)
    assert np.array_equal(ys, ref)
    # no worse than the reference's own error against an f64 transform
    f64 = np.fft.rfft(xs[:8].astype(np.float64), axis=1)
    assert rel_l2(ys[:8], f64) <= rel_l2(ref[:8], f64) * (1 + 1e-6)
    # DC / Nyquist bins are purely real for every row (src/rfft.rs:450-452)
    assert float(y[:, 0].imag.abs().max()) == 0.0 and float(y[:, m].imag.abs().max()) == 0.0
    # round trip through irfft, whole batch (kofft's own rfft->irfft error at this size is ~5e-3)
    z = cuda_fft.irfft_batch(y, n)
    torch.cuda.synchronize()
    err = torch.linalg.vector_norm(z - x, dim=1) / torch.linalg.vector_norm(x, dim=1)
    assert float(err.max()) < 1e-2
    assert np.array_equal(z[idx].cpu().numpy(), oracle.irfft_batch(ref, n, nthreads=8))


@pytest.mark.parametrize("chunk_bytes", [1 << 16, 1 << 20, 3 << 20])
def test_host_pipeline_matches_single_shot(cuda_fft, oracle, chunk_bytes):
    """Host-pointer batch calls cut into chunks over three streams (H2D / kernels / D2H) give the
    same bits as one copy + one launch, for ragged chunk counts, pinned and pageable buffers."""
    import torch

    ctx = cuda_fft.ctx
    rng = np.random.default_rng(chunk_bytes)
    try:
        # C2C in place, 301 rows of 1024 (2.4 MB): several chunks with a ragged tail
        x = uniform_c64(rng, (301, 1024))
        ref = oracle.fft_batch(x, nthreads=4)
        ctx.set_host_pipeline(chunk_bytes)
        y = x.copy()
        cuda_fft.fft_batch(y)
        assert np.array_equal(y, ref)
        pinned = torch.from_numpy(x.copy()).pin_memory()
        cuda_fft.fft_batch(pinned.numpy(), inverse=True)
        assert np.array_equal(pinned.numpy(), oracle.fft_batch(x, inverse=True, nthreads=4))
        # rfft / irfft out of place (row strides differ between input and output)
        r = rng.uniform(-1, 1, (77, 8192)).astype(np.float32)
        spec = cuda_fft.rfft_batch(r)
        assert np.array_equal(spec, oracle.rfft_batch(r))
        back = cuda_fft.irfft_batch(spec, 8192)
        ctx.set_host_pipeline(0)
        assert np.array_equal(back, cuda_fft.irfft_batch(spec, 8192))
        # large-N path shares one scratch between chunks: kernels must stay on one stream
        ctx.set_host_pipeline(chunk_bytes)
        big = uniform_c64(rng, (9, 32768))
        z = big.copy()
        cuda_fft.fft_batch(z)
        assert np.array_equal(z, oracle.fft_batch(big, nthreads=4))
    finally:
        ctx.set_host_pipeline(32 << 20)


@pytest.mark.parametrize("win_len,hop,length", [(2048, 512, 100_000), (1024, 256, 33_333), (256, 64, 5000), (4096, 4096, 40_000),
                                                 (1000, 250, 20_000), (16, 4, 999), (15, 5, 300)])
def test_stft_magnitudes_fused(cuda_fft, oracle, win_len, hop, length):
    """stft_magnitudes (src/visual/spectrogram.rs:52-76) with |X| and max fused behind the FFT; windows the fused
    kernel does not cover (not a power of two, below 32 points) take the general stft path and an |.| / max kernel."""
    import torch

    from kofft_b200 import spectrogram as SP

    rng = np.random.default_rng(win_len + hop)
    sig = rng.uniform(-1, 1, length).astype(np.float32)
    want, want_max = oracle.stft_magnitudes(sig, win_len, hop)
    mags, mx = SP.stft_magnitudes(cuda_fft, sig, win_len, hop)
    assert mags.shape == want.shape and np.array_equal(mags, want)
    assert mx == want_max
    # device tensors, several channels: one maximum per channel
    sig2 = rng.uniform(-1, 1, (3, length)).astype(np.float32)
    sig2[1] *= 0.25
    nframes = -(-length // hop)
    dm, dmx = SP.stft_magnitudes_batch(cuda_fft, torch.from_numpy(sig2).cuda(), torch.from_numpy(oracle.hann(win_len)).cuda(),
                                       hop, nframes)
    torch.cuda.synchronize()
    for c in range(3):
        w, wm = oracle.stft_magnitudes(sig2[c], win_len, hop)
        assert np.array_equal(dm[c].cpu().numpy(), w) and dmx[c].item() == wm
    with pytest.raises(Exception):
        SP.stft_magnitudes(cuda_fft, sig, win_len, 0)


@pytest.mark.parametrize("n", [3, 5, 6, 7, 12, 15, 100, 1000, 3000, 4097, 32767])
def test_bluestein_non_power_of_two(cuda_fft, cuda_fft_fast, oracle, n):
    """Non-power-of-two lengths: chirp tables generated like FftPlanner::get_bluestein
    (src/fft.rs:411-433), two power-of-two transforms of m = next_pow2(2n-1), Complex::mul unfused
    (src/fft.rs:1083-1132).  EXACT is bit-identical to the oracle; KAT tests/bluestein.rs:47-65."""
    rng = np.random.default_rng(n)
    rows = 3 if n > 1000 else 7
    x = uniform_c64(rng, (rows, n))
    ref = oracle.fft_batch(x)
    y = x.copy()
    cuda_fft.fft_batch(y)
    assert np.array_equal(y, ref), rel_l2(y, ref)
    z = x.copy()
    cuda_fft.fft_batch(z, inverse=True)
    assert np.array_equal(z, oracle.fft_batch(x, inverse=True))
    f = x.copy()
    cuda_fft_fast.fft_batch(f)
    assert rel_l2(f, ref) <= 20 * TOL  # FAST differs by ~1e-7 per transform, amplified by the chirp products
    if n == 15:  # tests/bluestein.rs:47-65: (i, -i) against a naive DFT, 1e-3 absolute
        d = (np.arange(n) - 1j * np.arange(n)).astype(np.complex64)
        want = np.fft.fft(d.astype(np.complex128))
        cuda_fft.fft(d)
        assert np.max(np.abs(d - want)) < 1e-3


def test_bluestein_beyond_32768(cuda_fft, oracle):
    """Non-power-of-two lengths whose padded length exceeds 2^16 run through the huge-N passes (m = 2^17, 2^18)."""
    rng = np.random.default_rng(40000)
    for n in (40000, 100003):
        x = uniform_c64(rng, (2, n))
        y = x.copy()
        cuda_fft.fft_batch(y)
        assert np.array_equal(y, oracle.fft_batch(x, nthreads=4)), n
    y = x.copy()
    cuda_fft.fft_batch(y, inverse=True)
    assert np.array_equal(y, oracle.fft_batch(x, inverse=True, nthreads=4))


def test_lengths_beyond_the_single_gpu_range_are_an_error(cuda_fft):
    import torch

    import kofft_b200 as k

    x = torch.zeros((1, 1 << 28), dtype=torch.complex64, device="cuda")
    with pytest.raises(k.CudaBackendError):
        cuda_fft.fft_batch(x)


@pytest.mark.parametrize("log2n", [17, 18, 19, 20, 21, 22])
def test_huge_c2c(cuda_fft, cuda_fft_fast, oracle, log2n):
    """Lengths above 2^16 (the reference has no upper bound, src/fft.rs:1054-1082, and benchmarks 2^20,
    benchmarks/README.md:5-7): 256-point column pass + register passes through global memory; every radix of the
    last pass (2^17: 1 stage, 2^18: 2, 2^19: 3, 2^20: 4) and several passes (2^21, 2^22).  Bit-exact."""
    n = 1 << log2n
    rng = np.random.default_rng(log2n)
    rows = 3 if log2n <= 20 else 2
    x = uniform_c64(rng, (rows, n))
    ref = oracle.fft_batch(x, nthreads=8)
    y = x.copy()
    cuda_fft.fft_batch(y)
    assert np.array_equal(y, ref), rel_l2(y, ref)
    if log2n in (17, 20, 22):
        y = x.copy()
        cuda_fft.fft_batch(y, inverse=True)
        assert np.array_equal(y, oracle.fft_batch(x, inverse=True, nthreads=8))
        z = x.copy()
        cuda_fft_fast.fft_batch(z)
        assert rel_l2(z, ref) <= TOL
        re, im = np.ascontiguousarray(x[1].real), np.ascontiguousarray(x[1].imag)
        cuda_fft.fft_split(re, im)
        assert np.array_equal(re, ref[1].real) and np.array_equal(im, ref[1].imag)
    if log2n == 20:  # the reference's bench input (i, 0) (kofft-bench/benches/bench_fft.rs:109), trait-level call
        ramp = (np.arange(n, dtype=np.float32) + 0j).astype(np.complex64)
        want = oracle.fft(ramp)
        cuda_fft.fft(ramp)
        assert np.array_equal(ramp, want)


@pytest.mark.parametrize("log2n", [18, 20, 21])
def test_huge_rfft_irfft(cuda_fft, cuda_fft_fast, oracle, log2n):
    """rfft / irfft above 2^17 (complex core above 2^16): column pass, register passes, twist kernel."""
    n = 1 << log2n
    rng = np.random.default_rng(100 + log2n)
    x = rng.uniform(-1, 1, (2, n)).astype(np.float32)
    ref = oracle.rfft_batch(x, nthreads=8)
    y = cuda_fft.rfft_batch(x)
    assert np.array_equal(y, ref), rel_l2(y, ref)
    assert np.array_equal(cuda_fft.irfft_batch(ref, n), oracle.irfft_batch(ref, n, nthreads=8))
    assert rel_l2(cuda_fft_fast.rfft_batch(x), ref) <= TOL
    import torch

    d = torch.from_numpy(x).cuda()  # device-pointer path
    assert np.array_equal(cuda_fft.rfft_batch(d).cpu().numpy(), ref)


def test_ndfft_2d_3d(cuda_fft, oracle):
    """fft2d_inplace / fft3d_inplace (src/ndfft.rs:74-153): rows then strided columns (then depth),
    bit-identical to composing the oracle's 1-D transform the same way; error cases :85-93, :125-133."""
    import torch

    import kofft_b200 as k
    from kofft_b200 import ndfft

    rng = np.random.default_rng(77)
    rows, cols = 64, 256
    x = uniform_c64(rng, (rows, cols))
    want = oracle.fft_batch(x)                                   # rows
    want = np.ascontiguousarray(oracle.fft_batch(np.ascontiguousarray(want.T)).T)  # columns (gather, fft, scatter)
    got = x.copy().reshape(-1)
    ndfft.fft2d_inplace(got, rows, cols, cuda_fft, np.zeros(rows, np.complex64))
    assert np.array_equal(got.reshape(rows, cols), want)
    assert np.allclose(want, np.fft.fft2(x.astype(np.complex128)), atol=1e-2)
    d = torch.from_numpy(x.copy()).cuda()
    ndfft.fft2d_inplace(d, rows, cols, cuda_fft)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy(), want)
    with pytest.raises(k.MismatchedLengths):
        ndfft.fft2d_inplace(x.copy().reshape(-1), rows, cols + 1, cuda_fft, np.zeros(rows, np.complex64))
    with pytest.raises(k.MismatchedLengths):
        ndfft.fft2d_inplace(x.copy().reshape(-1), rows, cols, cuda_fft, np.zeros(rows + 1, np.complex64))
    depth, rows, cols = 8, 16, 32
    v = uniform_c64(rng, (depth, rows, cols))
    w3 = np.moveaxis(oracle.fft_batch(np.ascontiguousarray(np.moveaxis(v, 0, 2).reshape(-1, depth))).reshape(rows, cols, depth), 2, 0)
    w3 = np.moveaxis(oracle.fft_batch(np.ascontiguousarray(np.moveaxis(w3, 1, 2).reshape(-1, rows))).reshape(depth, cols, rows), 2, 1)
    w3 = oracle.fft_batch(np.ascontiguousarray(w3.reshape(-1, cols))).reshape(depth, rows, cols)
    g3 = v.copy().reshape(-1)
    ndfft.fft3d_inplace(g3, depth, rows, cols, cuda_fft)
    assert np.array_equal(g3.reshape(depth, rows, cols), w3)
    assert np.allclose(w3, np.fft.fftn(v.astype(np.complex128)), atol=1e-2)
    with pytest.raises(k.MismatchedLengths):
        ndfft.fft3d_inplace(v.copy().reshape(-1), depth, rows, cols, cuda_fft, (np.zeros(depth), np.zeros(rows), np.zeros(cols + 1)))


@pytest.mark.parametrize("win_len,hop,chunks", [
    (2048, 512, [100, 5000, 1948, 512, 511, 20000, 3]),   # BASELINE window / hop, ragged pushes
    (1024, 300, [4096, 1, 1023, 7000]),                   # hop does not divide the window
    (512, 512, [512, 1000, 24]),                          # no overlap
    (256, 700, [300, 1000, 5, 2000, 1, 699, 702]),        # hop > window: samples between frames are skipped across pushes
    (1000, 250, [999, 1, 4000, 137, 2500]),               # non-power-of-two window: frames through Bluestein
])
def test_device_streams_match_offline(cuda_fft, oracle, win_len, hop, chunks):
    """Device twins of StftStream / IstftStream (src/stft.rs:160-206, 407-520; tests/istft_stream.rs):
    whatever the push sizes, the frames equal the offline stft and the samples equal the offline istft
    (and the oracle's restatement of IstftStream), bit for bit, for several channels at once."""
    import torch

    from kofft_b200 import stft as S

    ch = 3
    total = sum(chunks)
    rng = np.random.default_rng(win_len + hop)
    sig = rng.uniform(-1, 1, (ch, total)).astype(np.float32)
    w = oracle.hann(win_len)
    nframes = -(-total // hop)
    ref_frames = oracle.stft_batch(sig, w, hop, nframes)
    d_sig = torch.from_numpy(sig).cuda()
    st = S.DeviceStftStream(cuda_fft, ch, w, hop)
    got, pos = [], 0
    for n in chunks:
        got.append(st.push(d_sig[:, pos:pos + n]))  # a strided view: rows are `total` apart
        pos += n
    got.append(st.flush())
    assert st.flush().shape[1] == 0
    frames = torch.cat(got, dim=1)
    torch.cuda.synchronize()
    assert frames.shape[1] == nframes
    assert np.array_equal(frames.cpu().numpy(), ref_frames)
    if hop > win_len:
        return  # the inverse stream is defined for overlapping frames

    ist = S.DeviceIstftStream(cuda_fft, ch, w, hop)
    assert ist.flush().shape[1] == 0  # nothing before the first frame (src/stft.rs:498-500)
    ist = S.DeviceIstftStream(cuda_fft, ch, w, hop)
    outs, f0 = [], 0
    for k in (1, 2, 7, 1, 30, 5, 10 ** 9):
        k = min(k, nframes - f0)
        if k <= 0:
            break
        outs.append(ist.push_frames(frames[:, f0:f0 + k].contiguous()))
        f0 += k
    outs.append(ist.flush())
    assert ist.flush().shape[1] == 0
    rec = torch.cat(outs, dim=1).cpu().numpy()
    out_len = nframes * hop + max(win_len - hop, 0)
    assert rec.shape[1] == out_len
    for c in range(ch):
        want = oracle.istft(ref_frames[c], w, hop, np.zeros(out_len, np.float32))
        assert np.array_equal(rec[c], want), c
        assert np.array_equal(rec[c], oracle.istft_stream(ref_frames[c], w, hop)[:out_len])


def test_device_stream_errors(cuda_fft, oracle):
    from kofft_b200 import stft as S
    from kofft_b200.errors import InvalidHopSize

    with pytest.raises(InvalidHopSize):
        S.DeviceStftStream(cuda_fft, 1, oracle.hann(256), 0)
    with pytest.raises(InvalidHopSize):
        S.DeviceIstftStream(cuda_fft, 1, oracle.hann(256), 0)


def test_two_streams_share_one_context(cuda_fft, oracle):
    """The device-pointer calls are stream-ordered on the CALLER's stream, but the large-N intermediate, the
    dependency flags and the Bluestein / unfused-ISTFT workspaces are one per context: calls issued on two
    streams must not overlap on them (the library orders them with events).  Same results as one stream."""
    import torch

    n, rows = 32768, 300
    g = torch.Generator(device="cuda").manual_seed(5)
    xs = [torch.view_as_complex((torch.rand((rows, n, 2), generator=g, device="cuda") * 2 - 1).contiguous()) for _ in range(2)]
    xr = [(torch.rand((rows, 2 * n), generator=g, device="cuda") * 2 - 1).contiguous() for _ in range(2)]
    xb = [torch.view_as_complex((torch.rand((64, 3000, 2), generator=g, device="cuda") * 2 - 1).contiguous()) for _ in range(2)]
    want = [cuda_fft.fft_batch(x, out=torch.empty_like(x)) for x in xs]
    wantr = [cuda_fft.rfft_batch(x) for x in xr]
    wantb = [cuda_fft.fft_batch(x, out=torch.empty_like(x)) for x in xb]  # non-power-of-two: Bluestein workspace
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    got = [[None] * 2 for _ in range(3)]
    for rep in range(6):
        for si in (0, 1):
            with torch.cuda.stream(streams[si]):
                got[0][si] = cuda_fft.fft_batch(xs[si], out=torch.empty_like(xs[si]))
                got[1][si] = cuda_fft.rfft_batch(xr[si])
                got[2][si] = cuda_fft.fft_batch(xb[si], out=torch.empty_like(xb[si]))
    torch.cuda.synchronize()
    for si in (0, 1):
        assert torch.equal(torch.view_as_real(got[0][si]), torch.view_as_real(want[si]))
        assert torch.equal(torch.view_as_real(got[1][si]), torch.view_as_real(wantr[si]))
        assert torch.equal(torch.view_as_real(got[2][si]), torch.view_as_real(wantb[si]))
    assert np.array_equal(want[0][:3].cpu().numpy(), oracle.fft_batch(xs[0][:3].cpu().numpy(), nthreads=4))


def test_tensor_arguments_are_validated(cuda_fft):
    """A wrong dtype, a CPU tensor, a non-contiguous view or an undersized `out` raises instead of handing a bad
    pointer to the kernels."""
    import torch

    from kofft_b200 import spectrogram as SP
    from kofft_b200 import stft as S
    from kofft_b200.errors import MismatchedLengths

    x = torch.zeros((4, 256), dtype=torch.complex64, device="cuda")
    with pytest.raises(TypeError):
        cuda_fft.fft_batch(x.to(torch.complex128))
    with pytest.raises(TypeError):
        cuda_fft.fft_batch(x[:, ::2])
    with pytest.raises(MismatchedLengths):
        cuda_fft.fft_batch(x, out=torch.zeros((4, 128), dtype=torch.complex64, device="cuda"))
    r = torch.zeros((4, 256), dtype=torch.float32, device="cuda")
    with pytest.raises(TypeError):
        cuda_fft.rfft_batch(r.double())
    with pytest.raises(MismatchedLengths):
        cuda_fft.rfft_batch(r, out=torch.zeros((4, 128), dtype=torch.complex64, device="cuda"))
    with pytest.raises(TypeError):
        cuda_fft.irfft_batch(torch.zeros((4, 129), dtype=torch.complex64), 256)  # CPU tensor
    with pytest.raises(TypeError):
        cuda_fft.fft_split_batch(r, r.double())
    w = torch.ones(64, device="cuda")
    sig = torch.zeros((2, 1000), device="cuda")
    with pytest.raises(TypeError):
        S.stft_batch(cuda_fft, sig, w.cpu(), 16, 63)
    with pytest.raises(MismatchedLengths):
        S.stft_batch(cuda_fft, sig, w, 16, 63, out=torch.zeros((2, 63, 32), dtype=torch.complex64, device="cuda"))
    frames = torch.zeros((2, 63, 64), dtype=torch.complex64, device="cuda")
    with pytest.raises(TypeError):
        S.istft_batch(cuda_fft, frames, w.double(), 16, torch.zeros((2, 1000), device="cuda"))
    with pytest.raises(MismatchedLengths):
        S.istft_batch(cuda_fft, frames, w, 16, torch.zeros((3, 1000), device="cuda"))
    with pytest.raises(TypeError):
        SP.stft_magnitudes_batch(cuda_fft, sig.double(), w, 16, 63)
    with pytest.raises(MismatchedLengths):
        cuda_fft.fft_strided_batch(torch.zeros(100, dtype=torch.complex64, device="cuda"), 64, 2, 1, 64)


def test_rfft_planner_flavour_does_not_stick(cuda_fft, oracle):
    """RfftPlanner(fma_mul=True) uses the fused-multiply table for ITS calls only (src/rfft.rs:172-183 under +fma);
    the shared context goes back to the default table afterwards."""
    from kofft_b200.rfft import RfftPlanner

    rng = np.random.default_rng(17)
    x = rng.uniform(-1, 1, 4096).astype(np.float32)
    ref = oracle.rfft_batch(x[None])[0]
    out = np.zeros(2049, np.complex64)
    RfftPlanner(fma_mul=True).rfft(cuda_fft, x, out)
    again = np.zeros(2049, np.complex64)
    cuda_fft.rfft(x, again)
    assert np.array_equal(again, ref)
    assert np.array_equal(cuda_fft.rfft_batch(x[None])[0], ref)


@pytest.mark.parametrize("n", [6, 30, 100, 1000, 6000])
def test_non_power_of_two_through_every_core(cuda_fft, cuda_fft_fast, oracle, n):
    """The reference's rfft / irfft / stft / istft / split / strided all end in fft.fft(), which takes Bluestein for
    non-power-of-two lengths in the std build (src/rfft.rs:447, 502, src/stft.rs:102, 141, src/fft.rs:797-809,
    1191-1197).  Same here: bit-identical to the oracle, which restates those call chains."""
    from kofft_b200 import stft as S

    rng = np.random.default_rng(n)
    # rfft / irfft with a non-power-of-two half length m = n
    x = rng.uniform(-1, 1, (3, 2 * n)).astype(np.float32)
    ref = oracle.rfft_batch(x, nthreads=2)
    got = cuda_fft.rfft_batch(x)
    assert np.array_equal(got, ref), rel_l2(got, ref)
    assert np.array_equal(cuda_fft.irfft_batch(ref, 2 * n), oracle.irfft_batch(ref, 2 * n, nthreads=2))
    assert rel_l2(cuda_fft_fast.rfft_batch(x), ref) <= TOL
    one = np.zeros(n + 1, np.complex64)
    cuda_fft.rfft(x[0].copy(), one)  # trait-level call
    assert np.array_equal(one, ref[0])
    # split (SoA) and strided rows
    c = uniform_c64(rng, (n,))
    want = oracle.fft(c)
    re, im = np.ascontiguousarray(c.real), np.ascontiguousarray(c.imag)
    cuda_fft.fft_split(re, im)
    assert np.array_equal(re, want.real) and np.array_equal(im, want.imag)
    re, im = np.ascontiguousarray(c.real), np.ascontiguousarray(c.imag)
    cuda_fft.ifft_split(re, im)
    iwant = oracle.fft_batch(c[None], inverse=True)[0]
    assert np.array_equal(re, iwant.real) and np.array_equal(im, iwant.imag)
    buf = np.zeros(3 * n, np.complex64)
    buf[::3] = c
    cuda_fft.fft_strided(buf, 3, np.zeros(n, np.complex64))
    assert np.array_equal(buf[::3], want) and not buf[1::3].any()
    # stft / istft with a non-power-of-two window
    hop = max(1, n // 3)
    sig = rng.uniform(-1, 1, (2, 7 * n + 5)).astype(np.float32)
    w = oracle.hann(n)
    nframes = -(-sig.shape[1] // hop)
    fref = oracle.stft_batch(sig, w, hop, nframes)
    frames = S.stft_batch(cuda_fft, sig, w, hop, nframes)
    assert np.array_equal(frames, fref)
    rec = np.zeros_like(sig)
    S.istft_batch(cuda_fft, fref, w, hop, rec)
    wantr = np.stack([oracle.istft(fref[ch], w, hop, np.zeros(sig.shape[1], np.float32)) for ch in range(2)])
    assert np.array_equal(rec, wantr)


def test_persistent_kernels_flag_stress(cuda_fft):
    """The dependency flags of the two persistent cooperative kernels (split kernel: A warps / B warps over four
    intermediate slots; pipelined kernel: teams of 8 / 16 CTAs): many back-to-back launches, odd batch sizes, grids
    from one team up to the whole device, interleaved with other kernels on a second stream -- every result
    compared bit for bit with the two-kernel path, and no cooperative-launch fallback may have happened."""
    import torch

    C = cuda_fft.ctx
    g = torch.Generator(device="cuda").manual_seed(1)
    side = torch.cuda.Stream()
    noise = torch.rand((1 << 24,), device="cuda")
    fb0 = C.fallback_count
    runs = 0
    try:
        for n, rows in ((65536, 1111), (131072, 317), (65536, 37)):
            xr = (torch.rand((rows, n), generator=g, device="cuda") * 2 - 1).contiguous()
            xc = torch.view_as_complex((torch.rand((rows, n // 2, 2), generator=g, device="cuda") * 2 - 1).contiguous())
            C.set_split_min_log2n(16)
            C.set_large_mode(C.LARGE_TWO_KERNEL)
            ref_r = cuda_fft.rfft_batch(xr).clone()
            ref_c = cuda_fft.fft_batch(xc, out=torch.empty_like(xc)).clone()
            ref_i = cuda_fft.irfft_batch(ref_r, n).clone()  # (split kernel: rows untwisted by the B warps, their own flags)
            torch.cuda.synchronize()
            for split_min, mode in ((14, C.LARGE_AUTO), (16, C.LARGE_PIPELINED)):
                C.set_split_min_log2n(split_min)
                C.set_large_mode(mode)
                for max_ctas in (0, 148, 136, 64, 32, 16):  # at least one team of the widest kernel (16 CTAs)
                    C.set_max_ctas(max_ctas)
                    for rep in range(4):
                        with torch.cuda.stream(side):  # unrelated work competing for the SMs between the launches
                            noise.mul_(1.0001)
                        y = cuda_fft.rfft_batch(xr)
                        z = cuda_fft.fft_batch(xc, out=torch.empty_like(xc))
                        runs += 2
                        assert torch.equal(torch.view_as_real(y), torch.view_as_real(ref_r)), (n, rows, split_min, max_ctas, rep)
                        assert torch.equal(torch.view_as_real(z), torch.view_as_real(ref_c)), (n, rows, split_min, max_ctas, rep)
                        if rep < 2:
                            assert torch.equal(cuda_fft.irfft_batch(ref_r, n), ref_i), (n, rows, split_min, max_ctas, rep)
                C.set_max_ctas(0)
        torch.cuda.synchronize()
        assert C.fallback_count == fb0
        assert runs == 3 * 2 * 6 * 4 * 2
    finally:
        C.set_max_ctas(0)
        C.set_large_mode(C.LARGE_AUTO)
        C.set_split_min_log2n(14)


def test_fast_mode_full_size_rfft_and_stft(cuda_fft, cuda_fft_fast, oracle):
    """FAST mode (fused multiply-add butterflies) is the mode that meets the 70 % STFT target, so it gets the same
    full-size treatment as EXACT: BASELINE configs[2] (rfft 2^16 x 16384) and configs[3] (Hann 2048 / hop 512, as many
    channels as fit) -- whole batch against EXACT (which is bit-identical to the oracle) and sampled rows / frames
    against the oracle itself: rel-L2 and the worst row / frame within the north star's 1e-5."""
    import torch

    from kofft_b200 import stft as S
    from kofft_b200 import window as W

    def rel_rows(a, b):  # per-row rel-L2 of complex tensors [rows, n]
        d = (torch.view_as_real(a) - torch.view_as_real(b)).double().pow(2).sum(dim=(-2, -1))
        r = torch.view_as_real(b).double().pow(2).sum(dim=(-2, -1)).clamp_min(1e-300)
        return torch.sqrt(d.sum() / r.sum()).item(), torch.sqrt((d / r).max()).item()

    n, batch = 65536, 16384
    g = torch.Generator(device="cuda").manual_seed(1)
    x = (torch.rand((batch, n), generator=g, device="cuda") * 2 - 1).contiguous()
    ye = cuda_fft.rfft_batch(x)
    yf = cuda_fft_fast.rfft_batch(x)
    torch.cuda.synchronize()
    whole, worst = rel_rows(yf, ye)
    assert whole <= TOL and worst <= TOL, (whole, worst)
    rows = np.random.default_rng(2).choice(batch, 12, replace=False)
    ref = oracle.rfft_batch(x[torch.from_numpy(rows).cuda()].cpu().numpy(), nthreads=8)
    assert rel_l2(yf[torch.from_numpy(rows).cuda()].cpu().numpy(), ref) <= TOL
    zf = cuda_fft_fast.irfft_batch(yf, n)
    ze = cuda_fft.irfft_batch(ye, n)
    assert float((zf - ze).double().norm() / ze.double().norm()) <= TOL
    del x, ye, yf, zf, ze
    torch.cuda.empty_cache()

    free, _ = torch.cuda.mem_get_info()
    ch = 64 if free > 150e9 else (16 if free > 50e9 else 4)
    length, hop, win_len = 28_800_000, 512, 2048
    nframes = -(-length // hop)
    sig = (torch.rand((ch, length), generator=g, device="cuda") * 2 - 1).contiguous()
    w = W.hann(win_len)
    d_w = torch.from_numpy(w).cuda()
    fe = S.stft_batch(cuda_fft, sig, d_w, hop, nframes)
    ff = S.stft_batch(cuda_fft_fast, sig, d_w, hop, nframes)
    torch.cuda.synchronize()
    worst_all, whole_num, whole_den = 0.0, 0.0, 0.0
    for c in range(ch):  # channel by channel: the f64 temporaries stay small
        d = (torch.view_as_real(ff[c]) - torch.view_as_real(fe[c])).double().pow(2).sum(dim=(1, 2))
        r = torch.view_as_real(fe[c]).double().pow(2).sum(dim=(1, 2)).clamp_min(1e-300)
        worst_all = max(worst_all, float(torch.sqrt((d / r).max())))
        whole_num += float(d.sum())
        whole_den += float(r.sum())
    assert (whole_num / whole_den) ** 0.5 <= TOL and worst_all <= TOL, ((whole_num / whole_den) ** 0.5, worst_all)
    rng = np.random.default_rng(4)
    c = int(rng.integers(ch))
    host_sig = sig[c].cpu().numpy()
    for f in np.concatenate([rng.choice(nframes, 16, replace=False), [0, nframes - 1]]):
        seg = np.zeros(win_len, np.float32)
        part = host_sig[f * hop: f * hop + win_len]
        seg[: len(part)] = part
        want = oracle.fft((seg * w).astype(np.complex64))
        assert rel_l2(ff[c, f].cpu().numpy(), want) <= TOL, (c, f)
    del ff
    oe = torch.zeros((ch, length), device="cuda")
    of = torch.zeros((ch, length), device="cuda")
    S.istft_batch(cuda_fft, fe, d_w, hop, oe)
    S.istft_batch(cuda_fft_fast, fe, d_w, hop, of)
    torch.cuda.synchronize()
    per = (of - oe).double().norm(dim=1) / oe.double().norm(dim=1)
    assert float(per.max()) <= TOL, float(per.max())


def test_randomised_differential_against_the_oracle(cuda_fft, oracle):
    """Seeded random walk over the entry points, lengths (every size class: literal kernels, single-CTA engine, wide
    kernel, split kernel, multi-pass, Bluestein), batch sizes and directions; each call compared bit for bit with the
    oracle.  The sizes are bounded so the oracle finishes in seconds."""
    rng = np.random.default_rng(20261018)
    pow2 = [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072]
    other = [3, 5, 6, 7, 12, 24, 100, 255, 1000, 1023, 4097, 10000]
    for it in range(60):
        kind = rng.choice(["c2c", "ic2c", "rfft", "irfft", "split", "strided"])
        n = int(rng.choice(pow2 if rng.random() < 0.7 else other))
        rows = int(rng.integers(1, 6 if n >= 16384 else 40))
        tag = (it, kind, n, rows)
        if kind in ("c2c", "ic2c"):
            x = uniform_c64(rng, (rows, n))
            y = x.copy()
            cuda_fft.fft_batch(y, inverse=(kind == "ic2c"))
            assert np.array_equal(y, oracle.fft_batch(x, inverse=(kind == "ic2c"), nthreads=4)), tag
        elif kind == "rfft":
            xr = rng.uniform(-1, 1, (rows, 2 * n)).astype(np.float32)
            assert np.array_equal(cuda_fft.rfft_batch(xr), oracle.rfft_batch(xr, nthreads=4)), tag
        elif kind == "irfft":
            spec = uniform_c64(rng, (rows, n + 1))
            assert np.array_equal(cuda_fft.irfft_batch(spec, 2 * n), oracle.irfft_batch(spec, 2 * n, nthreads=4)), tag
        elif kind == "split":
            x = uniform_c64(rng, (1, n))[0]
            re, im = np.ascontiguousarray(x.real), np.ascontiguousarray(x.imag)
            cuda_fft.fft_split(re, im)
            ref = oracle.fft_batch(x[None, :])[0]
            assert np.array_equal(re, ref.real) and np.array_equal(im, ref.imag), tag
        else:
            stride = int(rng.integers(2, 5))
            buf = uniform_c64(rng, (1, n * stride))[0]
            want = buf.copy()
            want[::stride] = oracle.fft_batch(buf[::stride][None, :])[0]
            cuda_fft.fft_strided(buf, stride, np.zeros(n, np.complex64))
            assert np.array_equal(buf, want), tag
