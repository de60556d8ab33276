"""GPU parity of the f64 twin (`impl FftImpl<f64> for CudaFftImpl64`) through the C ABI's *_f64 entry
points, against the f64 oracle (oracle/kofft_oracle_f64.c): bit-exact for every power-of-two length
1 .. 8192, both directions, host-pointer and device-pointer paths, plus the reference's own f64
checks (tests/split64.rs) and the error behaviour."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fft64(cuda_fft):
    import kofft_b200

    return kofft_b200.CudaFftImpl64(ctx=cuda_fft.ctx)


def uniform_c128(rng, shape):
    return (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(np.complex128)


@pytest.mark.parametrize("n", [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_f64_fft_ifft_bit_exact(fft64, oracle, n):
    rng = np.random.default_rng(640 + n)
    rows = 37 if n <= 1024 else 9
    x = uniform_c128(rng, (rows, n))
    for inverse in (False, True):
        ref = oracle.fft_batch_f64(x, inverse=inverse, nthreads=4)
        y = x.copy()
        fft64.fft_batch(y, inverse=inverse)
        assert np.array_equal(y, ref), (n, inverse)
    one = x[0].copy()
    fft64.fft(one)  # trait-level single transform, in place
    assert np.array_equal(one, oracle.fft_f64(x[0]))
    fft64.ifft(one)
    assert np.array_equal(one, oracle.fft_f64(oracle.fft_f64(x[0]), inverse=True))


def test_f64_device_pointer_path_and_large_batch(fft64, oracle):
    """Stream-ordered device-pointer call, out of place and in place, a batch that makes the persistent
    CTAs loop many times; sampled rows bit-exact, Parseval and round trip over the whole batch."""
    import torch

    n, rows = 4096, 8192
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.view_as_complex((torch.rand((rows, n, 2), generator=g, device="cuda", dtype=torch.float64) * 2 - 1).contiguous())
    y = torch.empty_like(x)
    fft64.fft_batch(x, out=y)
    torch.cuda.synchronize()
    pick = [0, 1, 2, 777, rows - 1]
    assert np.array_equal(y[pick].cpu().numpy(), oracle.fft_batch_f64(x[pick].cpu().numpy()))
    ex, ey = float((x.abs() ** 2).sum()), float((y.abs() ** 2).sum())
    assert abs(ey / n - ex) / ex < 1e-12  # Parseval
    fft64.fft_batch(y, inverse=True)  # in place
    torch.cuda.synchronize()
    assert float((y - x).abs().max()) < 1e-11  # 12 stages of the f64 recurrence table, there and back


def test_f64_reference_checks(fft64, oracle):
    """tests/split64.rs: fft of (i, 0) for n = 16 and the ifft round trip of (i, -i); the f32-literal
    constants of the n = 8 / 16 kernels show up as a 1e-8 error against an exact DFT, as in the reference."""
    n = 16
    x = (np.arange(n) + 0j).astype(np.complex128)
    y = x.copy()
    fft64.fft(y)
    assert np.array_equal(y, oracle.fft_f64(x))
    assert 1e-10 < np.linalg.norm(y - np.fft.fft(x)) / np.linalg.norm(y) < 1e-6
    z = (np.arange(n) * (1 - 1j)).astype(np.complex128)
    w = z.copy()
    fft64.fft(w)
    fft64.ifft(w)
    assert np.abs(w - z).max() < 1e-5
    big = (np.arange(1024) * (1 - 1j)).astype(np.complex128)
    w = big.copy()
    fft64.fft(w)
    fft64.ifft(w)
    assert np.abs(w - big).max() < 1e-9


def test_f64_tables_and_errors(fft64, oracle):
    import kofft_b200
    from kofft_b200.errors import EmptyInput

    for n in (8, 32, 4096, 8192):
        assert np.array_equal(kofft_b200.FftPlanner64().get_twiddles(n), oracle.twiddles_f64(n))
    with pytest.raises(EmptyInput):
        fft64.fft(np.zeros(0, np.complex128))
    one = np.array([3 - 2j], np.complex128)
    fft64.fft(one)
    fft64.ifft(one)
    assert one[0] == 3 - 2j  # n == 1 is a no-op (src/fft.rs:1059-1061, 1139-1141)


def test_f64_split_strided_surface(fft64, oracle):
    """tests/split64.rs: fft_split == AoS fft (n = 32, x = (i, 0)) and the ifft_split round trip (n = 64,
    x = (i, -i)); plus fft_strided / fft_out_of_place_strided for f64 with the reference's argument
    meaning and errors (src/fft.rs:1175-1336), all bit-identical to the f64 oracle."""
    from kofft_b200.errors import InvalidStride, MismatchedLengths

    n = 32
    data = (np.arange(n) + 0j).astype(np.complex128)
    re, im = np.ascontiguousarray(data.real), np.ascontiguousarray(data.imag)
    fft64.fft_split(re, im)
    ref = oracle.fft_f64(data)
    assert np.array_equal(re, ref.real) and np.array_equal(im, ref.imag)
    n = 64
    data = (np.arange(n) * (1 - 1j)).astype(np.complex128)
    re, im = np.ascontiguousarray(data.real), np.ascontiguousarray(data.imag)
    fft64.fft_split(re, im)
    fft64.ifft_split(re, im)
    back = oracle.fft_f64(oracle.fft_f64(data), inverse=True)
    assert np.array_equal(re, back.real) and np.array_equal(im, back.imag)
    assert np.abs(re - data.real).max() < 1e-8 and np.abs(im - data.imag).max() < 1e-8
    with pytest.raises(MismatchedLengths):
        fft64.fft_split(np.zeros(8), np.zeros(4))
    # strided, in place: every 3rd element of a longer buffer
    rng = np.random.default_rng(3)
    n, stride = 256, 3
    buf = (rng.uniform(-1, 1, n * stride) + 1j * rng.uniform(-1, 1, n * stride)).astype(np.complex128)
    want = buf.copy()
    want[::stride] = oracle.fft_f64(buf[::stride])
    fft64.fft_strided(buf, stride, np.zeros(n, np.complex128))
    assert np.array_equal(buf, want)
    want[::stride] = oracle.fft_f64(want[::stride], inverse=True)
    fft64.ifft_strided(buf, stride, np.zeros(n, np.complex128))
    assert np.array_equal(buf, want)
    with pytest.raises(InvalidStride):
        fft64.fft_strided(buf, 0, np.zeros(n, np.complex128))
    with pytest.raises(MismatchedLengths):
        fft64.fft_strided(buf[:10], stride, np.zeros(n, np.complex128))
    # out of place, different strides; untouched output elements survive
    src = (rng.uniform(-1, 1, 128 * 2) + 1j * rng.uniform(-1, 1, 128 * 2)).astype(np.complex128)
    dst = np.full(128 * 5, 7 + 7j, np.complex128)
    fft64.fft_out_of_place_strided(src, 2, dst, 5)
    want = np.full(128 * 5, 7 + 7j, np.complex128)
    want[::5] = oracle.fft_f64(src[::2])
    assert np.array_equal(dst, want)
    with pytest.raises(InvalidStride):
        fft64.fft_out_of_place_strided(src, 0, dst, 5)


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 256, 2048, 4096, 8192, 16384])
def test_f64_rfft_irfft_bit_exact(fft64, oracle, n):
    """RealFftImpl<f64>: rfft / irfft for every supported length, host and device paths, against the
    f64 oracle bit for bit; tests/rfft_dispatch.rs round trip."""
    import torch

    rng = np.random.default_rng(700 + n)
    rows = 9
    x = rng.uniform(-1, 1, (rows, n))
    ref = oracle.rfft_batch_f64(x)
    y = fft64.rfft_batch(x)
    assert np.array_equal(y, ref), n
    back = oracle.irfft_batch_f64(ref, n)
    assert np.array_equal(fft64.irfft_batch(ref, n), back), n
    dx = torch.from_numpy(x).cuda()
    dy = fft64.rfft_batch(dx)
    dz = fft64.irfft_batch(dy, n)
    torch.cuda.synchronize()
    assert np.array_equal(dy.cpu().numpy(), ref) and np.array_equal(dz.cpu().numpy(), back)
    one_out = np.zeros(n // 2 + 1, np.complex128)
    fft64.rfft(x[0].copy(), one_out)
    assert np.array_equal(one_out, ref[0])
    one_back = np.zeros(n, np.float64)
    fft64.irfft(one_out, one_back)
    assert np.array_equal(one_back, back[0])


def test_f64_rfft_reference_checks_and_errors(fft64, oracle):
    import kofft_b200
    from kofft_b200.errors import EmptyInput, InvalidValue, MismatchedLengths

    x = np.array([1.0, 2.0, 3.0, 4.0])
    freq = np.zeros(3, np.complex128)
    fft64.rfft(x.copy(), freq)
    out = np.zeros(4)
    fft64.irfft(freq, out)
    assert np.abs(out - x).max() < 1e-10  # tests/rfft_dispatch.rs:24-41
    assert np.array_equal(kofft_b200.RfftPlanner64().get_twiddles(4096), oracle.rfft_twiddles_f64(4096))
    with pytest.raises(EmptyInput):
        fft64.rfft(np.zeros(0), np.zeros(1, np.complex128))
    with pytest.raises(InvalidValue):
        fft64.rfft(np.zeros(7), np.zeros(4, np.complex128))
    with pytest.raises(MismatchedLengths):
        fft64.rfft(np.zeros(8), np.zeros(4, np.complex128))


@pytest.mark.parametrize("log2n", [14, 15, 17, 18, 20])
def test_f64_above_8192(fft64, oracle, log2n):
    """ScalarFftImpl<f64> has no upper bound either (src/fft.rs:914-1051): above the single-CTA kernel's 8192 points
    the dense C2C rows make several register passes through global memory (fft_huge.cu); every last-pass radix
    (2^14: 2 stages, 2^15: 3, 2^17: 1, 2^20: 4).  Bit-exact, both directions, host and device paths."""
    import torch

    n = 1 << log2n
    rng = np.random.default_rng(6400 + log2n)
    x = uniform_c128(rng, (3, n))
    for inverse in (False, True):
        ref = oracle.fft_batch_f64(x, inverse=inverse, nthreads=4)
        y = x.copy()
        fft64.fft_batch(y, inverse=inverse)
        assert np.array_equal(y, ref), (log2n, inverse)
    d = torch.from_numpy(x).cuda()
    out = torch.empty_like(d)
    fft64.fft_batch(d, out=out)
    assert np.array_equal(out.cpu().numpy(), oracle.fft_batch_f64(x, nthreads=4))


@pytest.mark.parametrize("n", [3, 5, 6, 12, 100, 1000, 4097, 10007, 70001])
def test_f64_bluestein_bit_exact(fft64, oracle, n):
    """ScalarFftImpl<f64>::fft for non-power-of-two n (src/fft.rs:411-433, 1083-1132 with T = f64: the chirp angle's i*i
    passes through f32; transforms of m = next_pow2(2n - 1) points -- 4097 and up take the multi-pass f64 kernels):
    both directions, host and device paths, bit-identical to the f64 oracle; close to numpy's DFT."""
    import torch

    rng = np.random.default_rng(6464 + n)
    rows = 3 if n < 20000 else 1
    x = uniform_c128(rng, (rows, n))
    for inverse in (False, True):
        ref = oracle.fft_batch_f64(x, inverse=inverse, nthreads=4)
        y = x.copy()
        fft64.fft_batch(y, inverse=inverse)
        assert np.array_equal(y, ref), (n, inverse)
    d = torch.from_numpy(x).cuda()
    out = torch.empty_like(d)
    fft64.fft_batch(d, out=out)
    got = out.cpu().numpy()
    assert np.array_equal(got, oracle.fft_batch_f64(x, nthreads=4))
    want = np.fft.fft(x, axis=1)
    # as in the reference: m <= 16 runs the literal kernels with their f32 constants; above n = 4096 the f32-rounded
    # square in the chirp angle limits the accuracy
    tol = 1e-6 if n <= 8 else (1e-9 if n <= 4097 else 1e-2)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < tol


@pytest.mark.parametrize("n", [12, 1000, 16384, 40000])
def test_f64_split_strided_real_beyond_the_single_cta_range(fft64, oracle, n):
    """fft_split / ifft_split, fft_strided and rfft / irfft for f64 lengths the single-CTA kernel does not cover (not a
    power of two, or above 8192 points): the reference's own gather / fft / scatter and pack / fft / twist around the
    dense core (src/fft.rs:921-933, 1191-1197, 1414-1425; src/rfft.rs:425-508), bit-identical to the f64 oracle."""
    rng = np.random.default_rng(6500 + n)
    x = uniform_c128(rng, (1, n))[0]
    ref = oracle.fft_f64(x)
    re, im = np.ascontiguousarray(x.real), np.ascontiguousarray(x.imag)
    fft64.fft_split(re, im)
    assert np.array_equal(re, ref.real) and np.array_equal(im, ref.imag)
    # ifft_split: negate, fft_split, negate, * (1.0 / n as f64)
    re, im = np.ascontiguousarray(x.real), np.ascontiguousarray(x.imag)
    fft64.ifft_split(re, im)
    f = oracle.fft_f64(np.conj(x))
    scale = 1.0 / float(n)
    assert np.array_equal(re, f.real * scale) and np.array_equal(im, (-f.imag) * scale)
    # strided, in place: every 2nd element
    buf = uniform_c128(rng, (1, 2 * n))[0]
    want = buf.copy()
    want[::2] = oracle.fft_f64(buf[::2])
    fft64.fft_strided(buf, 2, np.zeros(n, np.complex128))
    assert np.array_equal(buf, want)
    # real transforms of 2 n samples (complex core of n points)
    xr = rng.uniform(-1, 1, (2, 2 * n))
    spec = oracle.rfft_batch_f64(xr)
    assert np.array_equal(fft64.rfft_batch(xr), spec)
    assert np.array_equal(fft64.irfft_batch(spec, 2 * n), oracle.irfft_batch_f64(spec, 2 * n))
