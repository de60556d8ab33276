"""The oracle against every value-pinning test the reference holds for the hot path
(SURVEY.md section 4 / 8c).  Each test cites the reference test it restates; inputs are the
reference's closed-form inputs, tolerances are the reference's own.  CPU only."""
import numpy as np
import pytest

from tests.conftest import rel_l2


def naive_dft_f32(x):
    """The reference tests' own f32 DFT helper (tests/pow2.rs:3-16)."""
    n = len(x)
    out = np.zeros(n, dtype=np.complex64)
    for k in range(n):
        acc = np.complex64(0)
        for j in range(n):
            ang = np.float32(-2.0) * np.float32(np.pi) * np.float32(k * j) / np.float32(n)
            tw = np.complex64(complex(np.cos(ang, dtype=np.float32), np.sin(ang, dtype=np.float32)))
            acc = np.complex64(acc + np.complex64(x[j]) * tw)
        out[k] = acc
    return out


def test_impulse_fft_ifft(oracle):  # src/lib.rs:178-199
    y = oracle.fft(np.array([1, 0, 0, 0], np.complex64))
    assert np.all(np.abs(y.real - 1) < 1e-6) and np.all(np.abs(y.imag) < 1e-6)
    z = oracle.ifft(y)
    assert abs(z[0].real - 1) < 1e-6 and np.all(np.abs(z[1:]) < 1e-6)


def test_all_zeros_and_ones(oracle):  # src/lib.rs:242-264
    assert np.all(np.abs(oracle.fft(np.zeros(8, np.complex64))) < 1e-6)
    y = oracle.fft(np.ones(8, np.complex64))
    assert abs(y[0].real - 8) < 1e-6 and np.all(np.abs(y[1:]) < 1e-6)


def test_hermitian_symmetry(oracle):  # src/lib.rs:360-388
    y = oracle.fft(np.array([1, 2, 3, 4], np.complex64))
    assert abs(y[1].real - y[3].real) < 1e-6 and abs(y[1].imag + y[3].imag) < 1e-6
    y = oracle.fft(np.array([1j, 2j, 3j, 4j], np.complex64))
    assert abs(y[1].real + y[3].real) < 1e-6 and abs(y[1].imag - y[3].imag) < 1e-6


def test_cosine_peak(oracle):  # src/lib.rs:218-240
    n = 8
    x = np.cos(2 * np.pi * np.arange(n, dtype=np.float32) / n).astype(np.complex64)
    mags = np.abs(oracle.fft(x))
    assert int(np.argmax(mags)) in (1, n - 1)


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32])
def test_pow2_vs_naive_dft(oracle, n):  # tests/pow2.rs:18-31, tests/small_kernels.rs:18-32
    i = np.arange(n, dtype=np.float32)
    x = (i - 0.5j * i).astype(np.complex64)
    y, ref = oracle.fft(x), naive_dft_f32(x)
    assert np.all(np.abs(y.real - ref.real) < 1e-2) and np.all(np.abs(y.imag - ref.imag) < 1e-2)


@pytest.mark.parametrize("n", [8, 16])
def test_direct_fft8_16(oracle, n):  # tests/small_kernels.rs:35-56
    i = np.arange(n, dtype=np.float32)
    x = (np.sin(i) + 1j * np.cos(i)).astype(np.complex64)
    y, ref = oracle.fft(x), naive_dft_f32(x)
    assert np.all(np.abs(y - ref) < 1e-2)


@pytest.mark.parametrize("n", [32, 64, 128, 256, 512, 1024])
def test_stockham_sizes_vs_f64(oracle, n):
    # tests/stockham_parity.rs, stockham_large.rs compare stockham_fft with fft (same code
    # path, tol 1e-3); here the same inputs are also held against an independent f64 DFT.
    i = np.arange(n, dtype=np.float32)
    x = ((i - 0.25j * i) if n <= 256 else (np.sin(i) + 1j * np.cos(i))).astype(np.complex64)
    y = oracle.fft(x)
    ref = np.fft.fft(x.astype(np.complex128))
    assert rel_l2(y, ref) < 5e-5


def test_parallel_stockham_input(oracle):  # tests/parallel_stockham.rs:6-27 (N=4096, x=(i,2i))
    i = np.arange(4096, dtype=np.float32)
    x = (i + 2j * i).astype(np.complex64)
    y = oracle.fft(x)
    assert np.array_equal(y, oracle.fft(x))  # "parallel" == serial: one code path, deterministic
    assert rel_l2(y, np.fft.fft(x.astype(np.complex128))) < 5e-4


def test_split_matches_aos(oracle):  # tests/split.rs:11-27 (N=16), :48-67 (round trip N=64)
    x = np.arange(16, dtype=np.float32).astype(np.complex64)
    re, im = oracle.fft_split(x.real.copy(), x.imag.copy())
    aos = oracle.fft(x)
    assert np.all(np.abs(aos.real - re) < 1e-6) and np.all(np.abs(aos.imag - im) < 1e-6)
    i = np.arange(64, dtype=np.float32)
    re, im = oracle.fft_split(i, -i)
    re, im = oracle.fft_split(re, im, inverse=True)
    assert np.all(np.abs(re - i) < 1e-4) and np.all(np.abs(im + i) < 1e-4)


def test_split_errors(oracle):  # tests/split.rs:69-78
    with pytest.raises(oracle.OracleError) as e:
        oracle.fft_split(np.zeros(4, np.float32), np.zeros(3, np.float32))
    assert e.value.variant == "MismatchedLengths"


def test_empty_and_single(oracle):  # src/lib.rs:313-318, 342-349
    with pytest.raises(oracle.OracleError) as e:
        oracle.fft(np.zeros(0, np.complex64))
    assert e.value.variant == "EmptyInput"
    y = oracle.fft(np.array([1 + 0j], np.complex64))
    assert y[0] == 1


def test_planner_twiddles(oracle):  # tests/twiddle.rs:7-13 (n=8, entry 1, 1e-6)
    t = oracle.twiddles(8)
    e = np.exp(-2j * np.pi / 8)
    assert len(t) == 4 and abs(t[1].real - e.real) < 1e-6 and abs(t[1].imag - e.imag) < 1e-6


def test_rfft_twiddles(oracle):  # tests/rfft_twiddles.rs:5-11 (m=8, entry 1, 1e-6)
    t = oracle.rfft_twiddles(8)
    e = np.exp(-1j * np.pi / 8)
    assert len(t) == 8 and abs(t[1].real - e.real) < 1e-6 and abs(t[1].imag - e.imag) < 1e-6


@pytest.mark.parametrize("x", [[1, 2, 3, 4], [1, 2, 3, 4, 5, 6, 7, 8]])
def test_rfft_roundtrip(oracle, x):  # src/lib.rs:431-450, tests/rfft_dispatch.rs, rfft.rs:892-906
    x = np.array(x, np.float32)
    f = oracle.rfft(x)
    assert abs(f[0].imag) < 1e-6 and abs(f[-1].imag) < 1e-6  # src/lib.rs:452-467
    assert np.all(np.abs(oracle.irfft(f, len(x)) - x) < 1e-5)
    assert rel_l2(f, np.fft.rfft(x.astype(np.float64))) < 1e-6


def test_rfft_errors(oracle):  # src/lib.rs:469-478; src/rfft.rs:433-443
    with pytest.raises(oracle.OracleError) as e:
        oracle.rfft(np.arange(4, dtype=np.float32), out_len=4)
    assert e.value.variant == "MismatchedLengths"
    with pytest.raises(oracle.OracleError) as e:
        oracle.rfft(np.arange(3, dtype=np.float32))
    assert e.value.variant == "InvalidValue"
    with pytest.raises(oracle.OracleError) as e:
        oracle.rfft(np.zeros(0, np.float32))
    assert e.value.variant == "EmptyInput"


def test_rfft_sin32(oracle):  # tests/rfft_arch_parity.rs:11-46 (size 32, sin(i)); SIMD == scalar
    x = np.sin(np.arange(32, dtype=np.float32))
    f = oracle.rfft(x)
    assert rel_l2(f, np.fft.rfft(x.astype(np.float64))) < 1e-5


def test_stft_insufficient_frames(oracle):  # tests/stft.rs:6-14
    with pytest.raises(oracle.OracleError) as e:
        oracle.stft(np.zeros(10, np.float32), oracle.hann(4), 4, 2)
    assert e.value.variant == "MismatchedLengths"
    with pytest.raises(oracle.OracleError) as e:
        oracle.stft(np.zeros(10, np.float32), oracle.hann(4), 0, 4)
    assert e.value.variant == "InvalidHopSize"


def test_stft_istft_roundtrip_rect(oracle):  # src/stft.rs:527-630, 800-813 (1..8, rect window)
    sig = np.arange(1, 9, dtype=np.float32)
    w = np.ones(4, np.float32)
    frames = oracle.stft(sig, w, 2, 4)
    out = oracle.istft(frames, w, 2, np.zeros(8, np.float32))
    assert np.all(np.abs(out - sig) < 1e-4)


def test_istft_stream_equals_offline(oracle):  # tests/istft_stream.rs:5-49 (bit-exact)
    sig = np.arange(1, 9, dtype=np.float32)
    w = np.ones(4, np.float32)
    frames = oracle.stft(sig, w, 2, 4)  # StftStream yields exactly these 4 frames
    stream = oracle.istft_stream(frames, w, 2)
    offline = oracle.istft(frames, w, 2, np.zeros(len(sig) + 4 - 2, np.float32))
    assert np.array_equal(stream[: len(sig)], offline[: len(sig)])
    assert len(stream) - len(sig) == 2 and np.array_equal(stream[len(sig):], offline[len(sig):])


def test_zero_window_guard(oracle):  # src/stft.rs:699-720: all-zero window -> output stays 0
    sig = np.arange(1, 9, dtype=np.float32)
    w = np.zeros(4, np.float32)
    frames = oracle.stft(sig, w, 2, 4)
    assert not frames.any()
    assert not oracle.istft(frames, w, 2, np.zeros(8, np.float32)).any()


def test_hann_stft_spectrogram_input(oracle):  # src/visual/spectrogram.rs:281-297 (0..15, hann 4, hop 2)
    sig = np.arange(16, dtype=np.float32)
    w = oracle.hann(4)
    a = oracle.stft(sig, w, 2, 8)
    b = oracle.stft_batch(sig.reshape(1, -1), w, 2, 8, fresh_planner=True, nthreads=2)[0]
    assert np.array_equal(a, b)  # parallel (fresh planner per frame) == serial, bit-exact


def test_windows(oracle):  # src/window.rs tests (:75-149): endpoints and symmetry
    h = oracle.hann(8)
    assert h[0] == 0 and abs(h[4] - 1) < 1e-6
    hm = oracle.hamming(8)
    assert abs(hm[0] - 0.08) < 1e-6
    b = oracle.blackman(8)
    assert abs(b[0]) < 1e-6
    k = oracle.kaiser(9, 5.0)
    assert abs(k[4] - 1) < 1e-6 and np.allclose(k, k[::-1], atol=1e-6)


def test_error_magnitudes_match_survey(oracle):
    """SURVEY.md 0.4 / BASELINE.md 4: kofft's f32 recurrence tables put its FFT this far from
    an f64 DFT; the oracle must reproduce those magnitudes (they are why exact tables are
    not good enough for parity)."""
    rng = np.random.default_rng(0)
    for n, lo, hi in [(1024, 5e-6, 4e-5), (2048, 2e-6, 3e-5), (4096, 3e-5, 3e-4), (32768, 3e-4, 3e-3)]:
        x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
        err = rel_l2(oracle.fft(x), np.fft.fft(x.astype(np.complex128)))
        assert lo < err < hi, (n, err)


def test_oracle_pinned_fixture(oracle):
    """tests/golden/oracle_pin.npz (made by tests/golden/make_golden.py in the build
    container) pins the oracle's own tables and outputs so a different libm / compiler on
    another box cannot silently move the yardstick."""
    import os

    path = os.path.join(os.path.dirname(__file__), "golden", "oracle_pin.npz")
    g = np.load(path)
    for n in (32, 1024, 2048, 4096, 32768):
        assert np.array_equal(oracle.twiddles(n), g[f"tw_{n}"]), n
    assert np.array_equal(oracle.rfft_twiddles(32768), g["rtw_32768"])
    assert np.array_equal(oracle.rfft_twiddles(32768, True), g["rtw_fma_32768"])
    x = g["x_4096"]
    assert np.array_equal(oracle.fft(x), g["y_4096"])
    assert np.array_equal(oracle.ifft(x), g["yi_4096"])
    assert np.array_equal(oracle.rfft(g["xr_8192"]), g["yr_8192"])
    assert np.array_equal(oracle.hann(2048), g["hann_2048"])
    assert np.array_equal(oracle.stft(g["sig"], g["hann_2048"], 512, 8), g["stft_frames"])


def test_bluestein_matches_dft(oracle):  # tests/bluestein.rs:47-65 (n = 15, (i, -i), 1e-3 absolute) + other lengths
    n = 15
    x = (np.arange(n) - 1j * np.arange(n)).astype(np.complex64)
    want = np.fft.fft(x.astype(np.complex128))
    got = oracle.fft(x)
    assert np.max(np.abs(got.real - want.real)) < 1e-3 and np.max(np.abs(got.imag - want.imag)) < 1e-3
    rng = np.random.default_rng(15)
    for n in (3, 5, 6, 7, 12, 100, 1000):
        x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
        want = np.fft.fft(x.astype(np.complex128))
        assert np.linalg.norm(oracle.fft(x) - want) / np.linalg.norm(want) < 2e-4
        assert np.linalg.norm(oracle.ifft(oracle.fft(x)) - x) / np.linalg.norm(x) < 4e-4


def test_stft_magnitudes_restatement(oracle):  # src/visual/spectrogram.rs:52-76
    rng = np.random.default_rng(52)
    sig = rng.uniform(-1, 1, 3000).astype(np.float32)
    mags, mx = oracle.stft_magnitudes(sig, 256, 64)
    assert mags.shape == (47, 128) and mags.dtype == np.float32
    fr = oracle.stft(sig, oracle.hann(256), 64, 47)[:, :128]
    assert np.allclose(mags, np.abs(fr), rtol=2e-7, atol=0) or np.max(np.abs(mags - np.abs(fr))) < 1e-5
    assert mx == mags.max() and mx > 0


# ---- f64 twin (oracle/kofft_oracle_f64.c) -------------------------------------------------------
def test_f64_split_matches_aos_and_roundtrip(oracle):
    """tests/split64.rs:4-17 (fft of (i, 0), n = 16, the split and AoS entries agree -- one code path
    here) and :36-56 (ifft(fft(x)) == x within 1e-9 for x = (i, -i), n = 16)."""
    n = 16
    x = np.arange(n, dtype=np.float64) + 0j
    y = oracle.fft_f64(x)
    ref = np.fft.fft(x)
    assert np.allclose(y, ref, atol=1e-5)  # the n = 16 kernel's constants are f32 literals widened to f64
    x2 = np.arange(n, dtype=np.float64) * (1 - 1j)
    back = oracle.fft_f64(oracle.fft_f64(x2), inverse=True)
    assert np.abs(back - x2).max() < 1e-5
    # power-of-two sizes on the Stockham path: f64 accuracy
    for n in (32, 64, 1024, 8192):
        rng = np.random.default_rng(n)
        x = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
        y = oracle.fft_f64(x)
        assert np.linalg.norm(y - np.fft.fft(x)) / np.linalg.norm(y) < 1e-12
        assert np.linalg.norm(oracle.fft_f64(y, inverse=True) - x) / np.linalg.norm(x) < 1e-12


def test_f64_small_kernels_use_f32_literals(oracle):
    """src/fft_kernels.rs builds its constants with T::from_f32(0.70710677) etc., so the f64 results for
    n = 8, 16 carry an f32-sized error (~1e-8) while n = 2, 4 and n >= 32 are f64-accurate: the oracle
    must reproduce that, not 'fix' it."""
    rng = np.random.default_rng(5)
    errs = {}
    for n in (2, 4, 8, 16, 32):
        x = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
        y = oracle.fft_f64(x)
        errs[n] = np.linalg.norm(y - np.fft.fft(x)) / np.linalg.norm(y)
    assert errs[2] < 1e-15 and errs[4] < 1e-15 and errs[32] < 1e-14
    assert 1e-10 < errs[8] < 1e-6 and 1e-10 < errs[16] < 1e-6


def test_f64_twiddle_table_and_errors(oracle):
    """static / planner table checks of tests/twiddle.rs restated for f64: entry 1 of n = 8 is
    exp(-2 pi i / 8); the recurrence stays within 1e-12 of the exact roots up to n = 8192."""
    t = oracle.twiddles_f64(8)
    assert abs(t[1] - np.exp(-2j * np.pi / 8)) < 1e-15
    t = oracle.twiddles_f64(8192)
    assert np.abs(t - np.exp(-2j * np.pi * np.arange(4096) / 8192)).max() < 1e-12
    with pytest.raises(oracle.OracleError) as e:
        oracle.fft_f64(np.zeros(0, np.complex128))
    assert e.value.variant == "EmptyInput"


def test_f64_rfft_roundtrip_dispatch(oracle):
    """tests/rfft_dispatch.rs:24-41: rfft then irfft of [1, 2, 3, 4] in f64 returns the input within 1e-10;
    plus agreement with numpy and the table's first entries."""
    x = np.array([[1.0, 2.0, 3.0, 4.0]])
    y = oracle.rfft_batch_f64(x)
    assert np.allclose(y, np.fft.rfft(x, axis=1), atol=1e-12)
    assert np.abs(oracle.irfft_batch_f64(y, 4) - x).max() < 1e-10
    for n in (64, 1024, 16384):
        rng = np.random.default_rng(n)
        x = rng.uniform(-1, 1, (2, n))
        y = oracle.rfft_batch_f64(x)
        assert np.linalg.norm(y - np.fft.rfft(x, axis=1)) / np.linalg.norm(y) < 1e-11
        assert np.abs(oracle.irfft_batch_f64(y, n) - x).max() < 1e-11
    t = oracle.rfft_twiddles_f64(8)
    assert t[0] == 1.0 and abs(t[1] - np.exp(-1j * np.pi / 8)) < 1e-15  # tests/rfft_twiddles.rs for f64


def test_timing_build_of_the_port_is_bit_identical(oracle):
    """bench.py's CPU legs time the -O3 / 128-bit-vector build of the same sources (oracle/Makefile; the reference's
    hot loop is explicit 4-wide SSE, src/fft.rs:845-862).  Vectorising must not change a bit."""
    rng = np.random.default_rng(99)
    for n in (64, 1024, 4096, 32768):
        x = (rng.uniform(-1, 1, (6, n)) + 1j * rng.uniform(-1, 1, (6, n))).astype(np.complex64)
        a, b = x.copy(), x.copy()
        oracle.fft_batch_inplace(a, nthreads=2)
        oracle.fft_batch_inplace(b, nthreads=2, fast=True)
        assert np.array_equal(a, b), n
    sig = rng.uniform(-1, 1, (2, 30000)).astype(np.float32)
    w = oracle.hann(2048)
    nf = -(-30000 // 512)
    assert np.array_equal(oracle.stft_batch(sig, w, 512, nf), oracle.stft_batch(sig, w, 512, nf, fast=True))


@pytest.mark.parametrize("n", [3, 6, 12, 100, 1000])
def test_f64_bluestein_restatement(oracle, n):
    """ScalarFftImpl<f64>::fft for non-power-of-two n (src/fft.rs:411-433, 1083-1132, T = f64): a DFT to the accuracy the
    reference's own constants allow (m <= 16: f32 literals in the small kernels), ifft round trip, linearity."""
    rng = np.random.default_rng(640 + n)
    x = (rng.uniform(-1, 1, (2, n)) + 1j * rng.uniform(-1, 1, (2, n))).astype(np.complex128)
    y = oracle.fft_batch_f64(x)
    want = np.fft.fft(x, axis=1)
    tol = 1e-6 if n <= 8 else 1e-11
    assert np.linalg.norm(y - want) / np.linalg.norm(want) < tol
    back = oracle.fft_batch_f64(y, inverse=True)
    assert np.abs(back - x).max() < (1e-6 if n <= 8 else 1e-10)
    s = oracle.fft_batch_f64((x[0] + x[1])[None, :])[0]
    assert np.abs(s - (y[0] + y[1])).max() < (1e-6 if n <= 8 else 1e-10)
