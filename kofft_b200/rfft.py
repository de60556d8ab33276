"""Real FFT interface (reference: src/rfft.rs).

  * `RfftPlanner<f32>`     src/rfft.rs:194-338 (twiddle cache exp(-i pi k / m), `:172-183`)
  * blanket `RealFftImpl`  src/rfft.rs:775-837 (methods live on `CudaFftImpl`)

On the GPU the pack step is a pointer reinterpretation and the Hermitian twist is fused
behind the last FFT stage, so `scratch` is only length-checked.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .errors import check
from .fft import CudaFftImpl


class RfftPlanner:
    """Caches rfft twiddle tables by half-length m (src/rfft.rs:194-262)."""

    PRECOMPUTED = (2, 4, 8, 16, 32, 64, 128, 256)  # src/rfft.rs:215

    def __init__(self, fma_mul: bool = False):
        self.fma_mul = fma_mul
        self._cache: dict[int, np.ndarray] = {}
        for m in self.PRECOMPUTED:
            self.get_twiddles(m)

    def get_twiddles(self, m: int) -> np.ndarray:
        """src/rfft.rs:246-252; repeated calls return the same array (tests/rfft_twiddles.rs:13-15)."""
        if m not in self._cache:
            out = np.empty(m, dtype=np.complex64)
            check(_lib.lib().kofft_cuda_rfft_twiddles_host_f32(m, out.ctypes.data, int(self.fma_mul)))
            out.flags.writeable = False
            self._cache[m] = out
        return self._cache[m]

    def get_pack_twiddles(self, m: int) -> np.ndarray:
        """src/rfft.rs:255-261: the reference keeps an identical second table."""
        return self.get_twiddles(m)

    def rfft_with_scratch(self, fft: CudaFftImpl, input, output, scratch) -> None:
        """src/rfft.rs:264-282"""
        saved = fft.ctx.rfft_table_fma  # the table flavour is this planner's, not the shared context's
        fft.ctx.set_rfft_table_fma(self.fma_mul)
        try:
            fft.rfft_with_scratch(input, output, scratch)
        finally:
            fft.ctx.set_rfft_table_fma(saved)

    def rfft(self, fft: CudaFftImpl, input, output) -> None:
        """src/rfft.rs:285-299"""
        self.rfft_with_scratch(fft, input, output, np.empty(len(input) // 2, dtype=np.complex64))

    def irfft_with_scratch(self, fft: CudaFftImpl, input, output, scratch) -> None:
        """src/rfft.rs:302-320"""
        saved = fft.ctx.rfft_table_fma
        fft.ctx.set_rfft_table_fma(self.fma_mul)
        try:
            fft.irfft_with_scratch(input, output, scratch)
        finally:
            fft.ctx.set_rfft_table_fma(saved)

    def irfft(self, fft: CudaFftImpl, input, output) -> None:
        """src/rfft.rs:323-337"""
        self.irfft_with_scratch(fft, input, output, np.empty(len(output) // 2, dtype=np.complex64))
