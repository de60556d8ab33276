"""ctypes loader for libkofft_cuda.so (the C ABI declared in include/kofft_cuda.h).

There is no fallback of any kind: if the shared library has not been built
(`python -c "import __graft_entry__ as g; g.build()"` or `make -C kofft_b200/csrc -j8`)
importing a compute entry point raises, and without a CUDA device `Context()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# KOFFT_CUDA_LIB: load another build of the same ABI (tuning variants, scripts/build_variants.sh)
LIB_PATH = os.environ.get("KOFFT_CUDA_LIB") or os.path.join(_HERE, "lib", "libkofft_cuda.so")

_sz, _vp, _i, _f = C.c_size_t, C.c_void_p, C.c_int, C.c_float

# name -> (restype, argtypes); kept in one place so tests can check that the library
# exports every symbol include/kofft_cuda.h declares.
SIGNATURES = {
    "kofft_cuda_create": (_i, [C.POINTER(_vp), _i]),
    "kofft_cuda_destroy": (None, [_vp]),
    "kofft_cuda_last_error": (C.c_char_p, []),
    "kofft_cuda_device": (_i, [_vp]),
    "kofft_cuda_stream": (_vp, [_vp]),
    "kofft_cuda_synchronize": (_i, [_vp]),
    "kofft_cuda_set_exact": (_i, [_vp, _i]),
    "kofft_cuda_get_exact": (_i, [_vp]),
    "kofft_cuda_launch_count": (C.c_ulonglong, [_vp]),
    "kofft_cuda_set_max_ctas": (_i, [_vp, _i]),
    "kofft_cuda_set_tma_staging": (_i, [_vp, _i]),
    "kofft_cuda_set_host_pipeline": (_i, [_vp, _sz]),
    "kofft_cuda_set_cluster_fusion": (_i, [_vp, _i]),
    "kofft_cuda_set_large_mode": (_i, [_vp, _i]),
    "kofft_cuda_set_split_min_log2n": (_i, [_vp, _i]),
    "kofft_cuda_set_split_all_kinds": (_i, [_vp, _i]),
    "kofft_cuda_set_wide_mask": (_i, [_vp, C.c_uint]),
    "kofft_cuda_fallback_count": (C.c_ulonglong, [_vp]),
    "kofft_cuda_set_istft_fusion": (_i, [_vp, _i, _i]),
    "kofft_cuda_twiddles_host_f32": (_i, [_sz, _vp]),
    "kofft_cuda_rfft_twiddles_host_f32": (_i, [_sz, _vp, _i]),
    "kofft_cuda_get_twiddles": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "kofft_cuda_get_rfft_twiddles": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "kofft_cuda_set_rfft_table_fma": (_i, [_vp, _i]),
    "kofft_cuda_window_host_f32": (_i, [_i, _sz, _f, _vp]),
    "kofft_cuda_fft_c2c_f32": (_i, [_vp, _vp, _vp, _sz, _sz, _i, _vp]),
    "kofft_cuda_stft_stream_create": (_i, [_vp, _sz, _vp, _sz, _sz, C.POINTER(_vp)]),
    "kofft_cuda_stft_stream_destroy": (None, [_vp]),
    "kofft_cuda_stft_stream_frames": (_sz, [_vp, _sz, _i]),
    "kofft_cuda_stft_stream_push": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, C.POINTER(_sz), _i, _vp]),
    "kofft_cuda_istft_stream_create": (_i, [_vp, _sz, _vp, _sz, _sz, C.POINTER(_vp)]),
    "kofft_cuda_istft_stream_destroy": (None, [_vp]),
    "kofft_cuda_istft_stream_push": (_i, [_vp, _vp, _sz, _vp, _sz, C.POINTER(_sz), _i, _vp]),
    "kofft_cuda_twiddles_host_f64": (_i, [_sz, _vp]),
    "kofft_cuda_fft_c2c_f64": (_i, [_vp, _vp, _vp, _sz, _sz, _i, _vp]),
    "kofft_cuda_fft_strided_f64": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _sz, _sz, _i, _vp]),
    "kofft_cuda_fft_split_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _sz, _i, _vp]),
    "kofft_cuda_fft_split_host_f64": (_i, [_vp, _vp, _sz, _vp, _sz, _i]),
    "kofft_cuda_fft_strided_host_f64": (_i, [_vp, _vp, _sz, _sz, _sz, _i]),
    "kofft_cuda_fft_out_of_place_strided_host_f64": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _i]),
    "kofft_cuda_rfft_twiddles_host_f64": (_i, [_sz, _vp]),
    "kofft_cuda_rfft_f64": (_i, [_vp, _vp, _vp, _sz, _sz, _vp]),
    "kofft_cuda_irfft_f64": (_i, [_vp, _vp, _vp, _sz, _sz, _vp]),
    "kofft_cuda_rfft_batch_host_f64": (_i, [_vp, _vp, _sz, _sz, _vp]),
    "kofft_cuda_irfft_batch_host_f64": (_i, [_vp, _vp, _sz, _sz, _vp]),
    "kofft_cuda_fft_host_f64": (_i, [_vp, _vp, _sz, _i]),
    "kofft_cuda_fft_batch_host_f64": (_i, [_vp, _vp, _sz, _sz, _i]),
    "kofft_cuda_fft_strided_f32": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _sz, _sz, _i, _vp]),
    "kofft_cuda_fft2d_f32": (_i, [_vp, _vp, _sz, _sz, _vp]),
    "kofft_cuda_fft3d_f32": (_i, [_vp, _vp, _sz, _sz, _sz, _vp]),
    "kofft_cuda_fft2d_host_f32": (_i, [_vp, _vp, _sz, _sz, _sz, _sz]),
    "kofft_cuda_fft3d_host_f32": (_i, [_vp, _vp, _sz, _sz, _sz, _sz, _sz, _sz, _sz]),
    "kofft_cuda_fft_split_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _sz, _i, _vp]),
    "kofft_cuda_rfft_f32": (_i, [_vp, _vp, _vp, _sz, _sz, _vp]),
    "kofft_cuda_irfft_f32": (_i, [_vp, _vp, _vp, _sz, _sz, _vp]),
    "kofft_cuda_stft_f32": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _vp, _sz, _vp]),
    "kofft_cuda_stft_magnitudes_f32": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _vp, _sz, _vp, _vp]),
    "kofft_cuda_stft_magnitudes_host_f32": (_i, [_vp, _vp, _sz, _sz, _sz, _vp, _sz, _vp]),
    "kofft_cuda_istft_f32": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _vp, _sz, _vp, _i, _vp]),
    "kofft_cuda_fft_host_f32": (_i, [_vp, _vp, _sz, _i]),
    "kofft_cuda_fft_batch_host_f32": (_i, [_vp, _vp, _sz, _sz, _i]),
    "kofft_cuda_fft_split_host_f32": (_i, [_vp, _vp, _sz, _vp, _sz, _i]),
    "kofft_cuda_fft_strided_host_f32": (_i, [_vp, _vp, _sz, _sz, _sz, _i]),
    "kofft_cuda_fft_out_of_place_strided_host_f32": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _i]),
    "kofft_cuda_rfft_host_f32": (_i, [_vp, _vp, _sz, _vp, _sz, _sz]),
    "kofft_cuda_rfft_batch_host_f32": (_i, [_vp, _vp, _sz, _sz, _vp]),
    "kofft_cuda_irfft_host_f32": (_i, [_vp, _vp, _sz, _vp, _sz, _sz]),
    "kofft_cuda_irfft_batch_host_f32": (_i, [_vp, _vp, _sz, _sz, _vp]),
    "kofft_cuda_stft_host_f32": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _vp, _sz]),
    "kofft_cuda_dist_create": (_i, [_vp, _i, _i, _i, C.POINTER(_vp)]),
    "kofft_cuda_dist_destroy": (None, [_vp]),
    "kofft_cuda_dist_shard_len": (_sz, [_vp]),
    "kofft_cuda_dist_set_pieces": (_i, [_vp, _i]),
    "kofft_cuda_dist_buffer": (_vp, [_vp, _i]),
    "kofft_cuda_dist_ipc_handles": (_i, [_vp, _vp]),
    "kofft_cuda_dist_connect_ipc": (_i, [_vp, _vp]),
    "kofft_cuda_dist_connect_local": (_i, [C.POINTER(_vp), _i]),
    "kofft_cuda_dist_phase": (_i, [_vp, _i, _vp, _vp, _i, _i, _vp]),
    "kofft_cuda_dist_pack": (_i, [_vp, _i, _vp, _vp, _i, _vp]),
    "kofft_cuda_dist_local_fft": (_i, [_vp, _i, _i, _vp]),
    "kofft_cuda_dist_run_local": (_i, [C.POINTER(_vp), _i, C.POINTER(_vp), C.POINTER(_vp), _i, _i]),
    "kofft_cuda_istft_host_f32": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _vp, _sz, _vp, _sz, _i]),
}

_lib = None


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C kofft_b200/csrc -j8` "
                "(or __graft_entry__.build()). kofft_b200 has no CPU fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the ABI and the header diverge
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    msg = lib().kofft_cuda_last_error()
    return msg.decode() if msg else ""
