"""kofft's FftError (reference: src/fft.rs:446-454) as Python exceptions.

The C ABI returns 1..6 for the six variants in declaration order; negative codes are CUDA
runtime errors, surfaced as CudaBackendError (the reference has no variant for a backend
failure, so this one is new).
"""
from __future__ import annotations


class FftError(Exception):
    """Base class; `variant` holds the reference's enum variant name."""

    variant = "FftError"
    code = 0

    def __init__(self, message: str = ""):
        super().__init__(message or self.variant)

    def __eq__(self, other):  # FftError derives PartialEq in the reference
        return isinstance(other, FftError) and other.code == self.code

    def __hash__(self):
        return hash(self.code)


class EmptyInput(FftError):
    variant, code = "EmptyInput", 1


class NonPowerOfTwoNoStd(FftError):
    variant, code = "NonPowerOfTwoNoStd", 2


class MismatchedLengths(FftError):
    variant, code = "MismatchedLengths", 3


class InvalidStride(FftError):
    variant, code = "InvalidStride", 4


class InvalidHopSize(FftError):
    variant, code = "InvalidHopSize", 5


class InvalidValue(FftError):
    variant, code = "InvalidValue", 6


class CudaBackendError(RuntimeError):
    def __init__(self, code: int, message: str):
        self.code = code
        super().__init__(f"CUDA backend error {code}: {message}")


_BY_CODE = {c.code: c for c in (EmptyInput, NonPowerOfTwoNoStd, MismatchedLengths, InvalidStride,
                                InvalidHopSize, InvalidValue)}


def check(rc: int) -> None:
    """Raise the exception matching a C-ABI return code (0 = Ok)."""
    if rc == 0:
        return
    if rc in _BY_CODE:
        raise _BY_CODE[rc]()
    from ._lib import last_error

    raise CudaBackendError(rc, last_error())
