"""kofft_b200 — B200 (sm_100a) backend for kofft's batched FFT / rfft / STFT hot path.

Package layout (only what the path needs):
  csrc/      hand-written CUDA kernels + the C ABI (include/kofft_cuda.h) -> lib/libkofft_cuda.so
  fft.py     mirror of `FftImpl` / `FftPlanner` / the batch free functions   (src/fft.rs)
  fft64.py   the f64 twin: `ScalarFftImpl<f64>` / `FftPlanner<f64>`            (src/fft.rs:914-1051)
  rfft.py    mirror of `RfftPlanner` / `RealFftImpl`                          (src/rfft.rs)
  stft.py    mirror of `stft` / `istft` / streams                              (src/stft.rs)
  spectrogram.py  `stft_magnitudes` with the magnitude + maximum fused behind the FFT (src/visual/spectrogram.rs)
  ndfft.py   `fft2d_inplace` / `fft3d_inplace`                               (src/ndfft.rs)
  dist.py    one transform sharded over several GPUs (BASELINE configs[4])
  window.py  `hann` / `hamming` / `blackman` / `kaiser`                        (src/window.rs)

There is no CPU fallback: the compute entry points raise if the CUDA library is missing.
"""
from . import errors, window  # noqa: F401
from .errors import (CudaBackendError, EmptyInput, FftError, InvalidHopSize, InvalidStride,  # noqa: F401
                     InvalidValue, MismatchedLengths, NonPowerOfTwoNoStd)
from .fft import (Context, CudaFftImpl, FftPlanner, FftStrategy, batch, batch_inverse,  # noqa: F401
                  fft_parallel, fft_split, ifft_parallel, ifft_split, multi_channel,
                  multi_channel_inverse, new_fft_impl)
from .fft64 import CudaFftImpl64, FftPlanner64, RfftPlanner64  # noqa: F401
from .rfft import RfftPlanner  # noqa: F401
from . import stft  # noqa: F401
from . import spectrogram  # noqa: F401
from . import ndfft  # noqa: F401

__all__ = [
    "Context", "CudaFftImpl", "CudaFftImpl64", "FftPlanner", "FftPlanner64", "RfftPlanner64", "FftStrategy", "RfftPlanner", "new_fft_impl",
    "batch", "batch_inverse", "multi_channel", "multi_channel_inverse",
    "fft_parallel", "ifft_parallel", "fft_split", "ifft_split",
    "FftError", "EmptyInput", "NonPowerOfTwoNoStd", "MismatchedLengths", "InvalidStride",
    "InvalidHopSize", "InvalidValue", "CudaBackendError", "stft", "spectrogram", "ndfft", "window", "errors",
]
