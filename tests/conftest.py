import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/kofft_oracle.c) -- the checker, never the thing under test."""
    from oracle import kofft_oracle as ko

    ko.build()
    ko.lib()
    return ko


@pytest.fixture(scope="session")
def emu():
    """CPU emulation of the CUDA CTA (tests/emu) -- test scaffolding for GPU-less CI."""
    from tests.emu import emu_binding

    return emu_binding.load()


@pytest.fixture(scope="session")
def emuk():
    """The real kernel bodies run by the cooperative block emulator (tests/emu/cuda_emu.h)."""
    from tests.emu import emu_binding

    return emu_binding.load_kernels()


@pytest.fixture(scope="session")
def cuda_fft():
    """The product: CudaFftImpl over libkofft_cuda.so on cuda:0 (exact mode)."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import kofft_b200

    return kofft_b200.CudaFftImpl(device=0, exact=True)


@pytest.fixture(scope="session")
def cuda_fft_fast():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import kofft_b200

    return kofft_b200.CudaFftImpl(device=0, exact=False)


def rel_l2(a, b):
    a = np.asarray(a).astype(np.complex128).ravel()
    b = np.asarray(b).astype(np.complex128).ravel()
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def uniform_c64(rng, shape):
    return (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(np.complex64)
