"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares,
its host table generators are bit-identical to the oracle's, argument validation mirrors the
reference's error order, and it refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "kofft_cuda.h")


@pytest.fixture(scope="module")
def lib():
    from kofft_b200 import _lib

    return _lib.lib()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kofft_cuda_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    from kofft_b200 import _lib

    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/kofft_cuda.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_header_cites_reference_for_each_entry():
    text = open(HEADER).read()
    assert text.count("src/") >= 25


@pytest.mark.parametrize("n", [2, 8, 32, 1024, 2048, 4096, 16384, 32768, 65536])
def test_fft_twiddle_table_bit_identical_to_oracle(lib, oracle, n):
    out = np.empty(n // 2, np.complex64)
    assert lib.kofft_cuda_twiddles_host_f32(n, out.ctypes.data) == 0
    assert np.array_equal(out, oracle.twiddles(n))


@pytest.mark.parametrize("m", [1, 2, 8, 256, 4096, 32768])
@pytest.mark.parametrize("fma", [0, 1])
def test_rfft_twiddle_table_bit_identical_to_oracle(lib, oracle, m, fma):
    out = np.empty(m, np.complex64)
    assert lib.kofft_cuda_rfft_twiddles_host_f32(m, out.ctypes.data, fma) == 0
    assert np.array_equal(out, oracle.rfft_twiddles(m, bool(fma)))


@pytest.mark.parametrize("length", [1, 2, 4, 7, 64, 2048])
def test_windows_bit_identical_to_oracle(oracle, length):
    from kofft_b200 import window

    assert np.array_equal(window.hann(length), oracle.hann(length))
    assert np.array_equal(window.hamming(length), oracle.hamming(length))
    assert np.array_equal(window.blackman(length), oracle.blackman(length))
    if length > 1:
        assert np.array_equal(window.kaiser(length, 8.6), oracle.kaiser(length, 8.6))


def test_planner_returns_same_table_object():
    # tests/rfft_twiddles.rs:13-15 / tests/twiddle.rs:12-13: repeated lookups give the same pointer
    import kofft_b200

    p = kofft_b200.FftPlanner()
    assert p.get_twiddles(8) is p.get_twiddles(8)
    r = kofft_b200.RfftPlanner()
    t = r.get_twiddles(8)
    assert t is r.get_twiddles(8) and len(t) == 8
    e = np.exp(-1j * np.pi / 8)
    assert abs(t[1].real - e.real) < 1e-6 and abs(t[1].imag - e.imag) < 1e-6
    from kofft_b200 import FftStrategy

    assert p.plan_strategy(8) == FftStrategy.SplitRadix and p.plan_strategy(12) == FftStrategy.Auto


def test_argument_validation_precedes_device_work(lib):
    """Length / stride / hop checks return the reference's FftError code before the context
    is touched (ctx = NULL here), in the reference's order."""
    z = None
    assert lib.kofft_cuda_fft_c2c_f32(z, z, z, 0, 1, 0, z) == 1        # EmptyInput
    assert lib.kofft_cuda_rfft_f32(z, z, z, 24, 1, z) == 6             # m = 12 is fine (Bluestein, as the std build): fails on the NULL context only
    assert lib.kofft_cuda_stft_stream_create(z, 1, z, 12, 4, z) == 6   # null arguments
    assert lib.kofft_cuda_fft_strided_f32(z, z, 0, 8, z, 1, 8, 8, 1, 0, z) == 4  # InvalidStride
    assert lib.kofft_cuda_rfft_f32(z, z, z, 0, 1, z) == 1
    assert lib.kofft_cuda_rfft_f32(z, z, z, 7, 1, z) == 6              # InvalidValue (odd)
    assert lib.kofft_cuda_irfft_f32(z, z, z, 0, 1, z) == 1
    assert lib.kofft_cuda_stft_f32(z, z, 10, 1, z, 4, 0, z, 3, z) == 5  # InvalidHopSize
    assert lib.kofft_cuda_stft_f32(z, z, 10, 1, z, 4, 4, z, 2, z) == 3  # tests/stft.rs:6-14
    assert lib.kofft_cuda_istft_f32(z, z, 3, 1, z, 4, 0, z, 8, z, 0, z) == 5
    assert lib.kofft_cuda_fft_split_host_f32(z, z, 4, z, 3, 0) == 3     # tests/split.rs:69-78
    assert lib.kofft_cuda_fft_strided_host_f32(z, z, 8, 0, 4, 0) == 4
    assert lib.kofft_cuda_fft_strided_host_f32(z, z, 8, 3, 4, 0) == 3   # input shorter than (n-1)*stride+1
    assert lib.kofft_cuda_fft_out_of_place_strided_host_f32(z, z, 8, 3, z, 8, 2, 0) == 4
    assert lib.kofft_cuda_fft_out_of_place_strided_host_f32(z, z, 8, 2, z, 8, 4, 0) == 3
    assert lib.kofft_cuda_rfft_host_f32(z, z, 4, z, 4, 2) == 3          # src/lib.rs:469-478
    assert lib.kofft_cuda_rfft_host_f32(z, z, 4, z, 3, 1) == 3          # scratch too short
    assert lib.kofft_cuda_irfft_host_f32(z, z, 4, z, 4, 2) == 3


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, not compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import kofft_b200

    with pytest.raises(kofft_b200.CudaBackendError):
        kofft_b200.CudaFftImpl()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "kofft_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(base, f)).read()
                assert "kofft_oracle" not in text and "oracle/" not in text, os.path.join(base, f)
