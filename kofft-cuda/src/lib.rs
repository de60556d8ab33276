//! `CudaFftImpl`: kofft's `FftImpl<f32>` on an NVIDIA B200 (sm_100a).
//!
//! Drop-in for the reference's `ScalarFftImpl<f32>` on the batched FFT / rfft / STFT path:
//! the trait methods keep their argument meaning and error behaviour and produce results
//! bit-identical to kofft's scalar/SSE f32 path (the kernels reuse kofft's own twiddle tables
//! and stage structure).  `rfft`/`irfft` arrive through kofft's blanket `RealFftImpl`; the
//! fused batched variants (`fft_batch`, `rfft_batch`, `stft`, `istft`) are inherent methods
//! because the blanket impl cannot be specialised.
//!
//! FFI: `include/kofft_cuda.h` (C ABI of libkofft_cuda).  Return codes 1..=6 are the
//! `FftError` variants in declaration order.  A negative code is a CUDA failure, for which kofft
//! has no variant: the shim panics with the backend's message -- except for "not supported"
//! (lengths beyond the single-GPU range of 2^27, f64 lengths the f64 kernels do not cover), which is
//! an ordinary property of the input and comes back as `FftError::InvalidValue`.  There is no CPU
//! fallback of any kind.
#![allow(clippy::missing_safety_doc)]

use core::ffi::{c_char, c_int, c_void};
use kofft::fft::{Complex32, Complex64, FftError, FftImpl, FftStrategy};

#[repr(C)]
pub struct RawCtx {
    _private: [u8; 0],
}

extern "C" {
    fn kofft_cuda_create(out: *mut *mut RawCtx, device: c_int) -> c_int;
    fn kofft_cuda_destroy(ctx: *mut RawCtx);
    fn kofft_cuda_last_error() -> *const c_char;
    fn kofft_cuda_set_exact(ctx: *mut RawCtx, exact: c_int) -> c_int;
    fn kofft_cuda_fft_host_f32(ctx: *mut RawCtx, data: *mut f32, n: usize, inverse: c_int) -> c_int;
    fn kofft_cuda_fft_batch_host_f32(ctx: *mut RawCtx, data: *mut f32, n: usize, batch: usize, inverse: c_int) -> c_int;
    fn kofft_cuda_fft_host_f64(ctx: *mut RawCtx, data: *mut f64, n: usize, inverse: c_int) -> c_int;
    fn kofft_cuda_fft_batch_host_f64(ctx: *mut RawCtx, data: *mut f64, n: usize, batch: usize, inverse: c_int) -> c_int;
    fn kofft_cuda_rfft_batch_host_f64(ctx: *mut RawCtx, input: *const f64, n: usize, batch: usize, output: *mut f64) -> c_int;
    fn kofft_cuda_irfft_batch_host_f64(ctx: *mut RawCtx, input: *const f64, n: usize, batch: usize, output: *mut f64) -> c_int;
    fn kofft_cuda_fft_split_host_f64(ctx: *mut RawCtx, re: *mut f64, re_len: usize, im: *mut f64, im_len: usize,
                                     inverse: c_int) -> c_int;
    fn kofft_cuda_fft_strided_host_f64(ctx: *mut RawCtx, input: *mut f64, input_len: usize, stride: usize, n: usize,
                                       inverse: c_int) -> c_int;
    fn kofft_cuda_fft_out_of_place_strided_host_f64(ctx: *mut RawCtx, input: *const f64, input_len: usize,
                                                    in_stride: usize, output: *mut f64, output_len: usize,
                                                    out_stride: usize, inverse: c_int) -> c_int;
    fn kofft_cuda_fft_split_host_f32(ctx: *mut RawCtx, re: *mut f32, re_len: usize, im: *mut f32, im_len: usize,
                                     inverse: c_int) -> c_int;
    fn kofft_cuda_fft_strided_host_f32(ctx: *mut RawCtx, input: *mut f32, input_len: usize, stride: usize, n: usize,
                                       inverse: c_int) -> c_int;
    fn kofft_cuda_fft_out_of_place_strided_host_f32(ctx: *mut RawCtx, input: *const f32, input_len: usize,
                                                    in_stride: usize, output: *mut f32, output_len: usize,
                                                    out_stride: usize, inverse: c_int) -> c_int;
    fn kofft_cuda_rfft_batch_host_f32(ctx: *mut RawCtx, input: *const f32, n: usize, batch: usize, output: *mut f32) -> c_int;
    fn kofft_cuda_irfft_batch_host_f32(ctx: *mut RawCtx, input: *const f32, n: usize, batch: usize, output: *mut f32) -> c_int;
    fn kofft_cuda_stft_host_f32(ctx: *mut RawCtx, signal: *const f32, len: usize, channels: usize, window: *const f32,
                                win_len: usize, hop: usize, frames: *mut f32, nframes: usize) -> c_int;
    fn kofft_cuda_istft_host_f32(ctx: *mut RawCtx, frames: *const f32, nframes: usize, channels: usize,
                                 window: *const f32, win_len: usize, hop: usize, output: *mut f32, out_len: usize,
                                 scratch: *mut f32, scratch_len: usize, zero_uncovered: c_int) -> c_int;
    fn kofft_cuda_stft_magnitudes_host_f32(ctx: *mut RawCtx, samples: *const f32, len: usize, win_len: usize, hop: usize,
                                           mags: *mut f32, nframes: usize, max_mag: *mut f32) -> c_int;
    fn kofft_cuda_fft2d_host_f32(ctx: *mut RawCtx, data: *mut f32, data_len: usize, rows: usize, cols: usize,
                                 scratch_col_len: usize) -> c_int;
    fn kofft_cuda_fft3d_host_f32(ctx: *mut RawCtx, data: *mut f32, data_len: usize, depth: usize, rows: usize,
                                 cols: usize, tube_len: usize, row_len: usize, col_len: usize) -> c_int;
    fn kofft_cuda_set_host_pipeline(ctx: *mut RawCtx, chunk_bytes: usize) -> c_int;
    // one transform sharded over several GPUs (include/kofft_cuda.h, "kofft_cuda_dist_*")
    pub fn kofft_cuda_dist_create(ctx: *mut RawCtx, rank: c_int, world: c_int, log2n: c_int, out: *mut *mut c_void) -> c_int;
    pub fn kofft_cuda_dist_destroy(d: *mut c_void);
    pub fn kofft_cuda_dist_ipc_handles(d: *mut c_void, out128: *mut c_void) -> c_int;
    pub fn kofft_cuda_dist_connect_ipc(d: *mut c_void, all_handles: *const c_void) -> c_int;
    pub fn kofft_cuda_dist_connect_local(dists: *const *mut c_void, world: c_int) -> c_int;
    pub fn kofft_cuda_dist_phase(d: *mut c_void, phase: c_int, d_in: *const c_void, d_out: *mut c_void, inverse: c_int,
                                 natural_order: c_int, stream: *mut c_void) -> c_int;
    pub fn kofft_cuda_dist_run_local(dists: *const *mut c_void, world: c_int, d_in: *const *const c_void,
                                     d_out: *const *mut c_void, inverse: c_int, natural_order: c_int) -> c_int;
    // device-pointer entry points (stream-ordered) for callers that keep data on the GPU
    pub fn kofft_cuda_fft_c2c_f32(ctx: *mut RawCtx, d_in: *const c_void, d_out: *mut c_void, n: usize, batch: usize,
                                  inverse: c_int, stream: *mut c_void) -> c_int;
    pub fn kofft_cuda_rfft_f32(ctx: *mut RawCtx, d_in: *const f32, d_out: *mut c_void, n: usize, batch: usize,
                               stream: *mut c_void) -> c_int;
    pub fn kofft_cuda_stft_f32(ctx: *mut RawCtx, d_signal: *const f32, len: usize, channels: usize,
                               d_window: *const f32, win_len: usize, hop: usize, d_frames: *mut c_void,
                               nframes: usize, stream: *mut c_void) -> c_int;
}

/// -(cudaErrorNotSupported): the backend has no kernel for this length (see include/kofft_cuda.h, "Lengths")
const NOT_SUPPORTED: c_int = -801;

fn check(rc: c_int) -> Result<(), FftError> {
    match rc {
        0 => Ok(()),
        1 => Err(FftError::EmptyInput),
        2 => Err(FftError::NonPowerOfTwoNoStd),
        3 => Err(FftError::MismatchedLengths),
        4 => Err(FftError::InvalidStride),
        5 => Err(FftError::InvalidHopSize),
        6 => Err(FftError::InvalidValue),
        NOT_SUPPORTED => Err(FftError::InvalidValue),
        _ => {
            // kofft's FftError has no variant for a backend failure
            let msg = unsafe { std::ffi::CStr::from_ptr(kofft_cuda_last_error()) };
            panic!("kofft-cuda: CUDA backend error {rc}: {}", msg.to_string_lossy());
        }
    }
}

/// Owns one device context (stream, device-resident twiddle tables, workspaces).
/// `!Sync` like the reference's `ScalarFftImpl` (src/fft.rs:589-605); `Send`.
pub struct CudaFftImpl {
    ctx: *mut RawCtx,
}

unsafe impl Send for CudaFftImpl {}

impl CudaFftImpl {
    /// Fails (no CPU fallback) if `device` is not a CUDA device of compute capability 10.x.
    pub fn new(device: i32) -> Result<Self, String> {
        let mut ctx = core::ptr::null_mut();
        let rc = unsafe { kofft_cuda_create(&mut ctx, device) };
        if rc != 0 {
            let msg = unsafe { std::ffi::CStr::from_ptr(kofft_cuda_last_error()) };
            return Err(msg.to_string_lossy().into_owned());
        }
        Ok(Self { ctx })
    }
    /// `false` selects the FMA-contracted butterflies (~1e-7 relative difference).
    pub fn set_exact(&self, exact: bool) {
        unsafe { kofft_cuda_set_exact(self.ctx, exact as c_int) };
    }
    pub fn raw(&self) -> *mut RawCtx {
        self.ctx
    }

    /// `batch()` (src/fft.rs:2156-2164) over dense rows `[batch][n]` in one launch.
    pub fn fft_batch(&self, rows: &mut [Complex32], n: usize, inverse: bool) -> Result<(), FftError> {
        if n == 0 {
            return Err(FftError::EmptyInput);
        }
        if rows.len() % n != 0 {
            return Err(FftError::MismatchedLengths);
        }
        check(unsafe {
            kofft_cuda_fft_batch_host_f32(self.ctx, rows.as_mut_ptr() as *mut f32, n, rows.len() / n, inverse as c_int)
        })
    }
    /// Fused pack + FFT + twist per row: `[batch][n]` reals -> `[batch][n/2+1]` bins.
    pub fn rfft_batch(&self, input: &[f32], n: usize, output: &mut [Complex32]) -> Result<(), FftError> {
        if n == 0 {
            return Err(FftError::EmptyInput);
        }
        let batch = input.len() / n;
        if input.len() % n != 0 || output.len() != batch * (n / 2 + 1) {
            return Err(FftError::MismatchedLengths);
        }
        check(unsafe { kofft_cuda_rfft_batch_host_f32(self.ctx, input.as_ptr(), n, batch, output.as_mut_ptr() as *mut f32) })
    }
    pub fn irfft_batch(&self, input: &[Complex32], n: usize, output: &mut [f32]) -> Result<(), FftError> {
        if n == 0 {
            return Err(FftError::EmptyInput);
        }
        let batch = output.len() / n;
        if output.len() % n != 0 || input.len() != batch * (n / 2 + 1) {
            return Err(FftError::MismatchedLengths);
        }
        check(unsafe { kofft_cuda_irfft_batch_host_f32(self.ctx, input.as_ptr() as *const f32, n, batch, output.as_mut_ptr()) })
    }
    /// `stft()` (src/stft.rs:76-105) for `channels` signals of `len` samples; `frames` is dense
    /// `[channels][nframes][window.len()]`.
    pub fn stft(&self, signal: &[f32], channels: usize, window: &[f32], hop: usize, frames: &mut [Complex32],
                nframes: usize) -> Result<(), FftError> {
        if channels == 0 || signal.len() % channels != 0 || frames.len() != channels * nframes * window.len() {
            return Err(FftError::MismatchedLengths);
        }
        check(unsafe {
            kofft_cuda_stft_host_f32(self.ctx, signal.as_ptr(), signal.len() / channels, channels, window.as_ptr(),
                                     window.len(), hop, frames.as_mut_ptr() as *mut f32, nframes)
        })
    }
    /// `istft()` (src/stft.rs:117-156): accumulates into `output`, fills `scratch` with the
    /// window-power sums, normalises where the sum exceeds 1e-8.
    pub fn istft(&self, frames: &[Complex32], channels: usize, window: &[f32], hop: usize, output: &mut [f32],
                 scratch: &mut [f32]) -> Result<(), FftError> {
        if scratch.len() != output.len() || channels == 0 || window.is_empty() {
            return Err(FftError::MismatchedLengths);
        }
        let nframes = frames.len() / (channels * window.len());
        check(unsafe {
            kofft_cuda_istft_host_f32(self.ctx, frames.as_ptr() as *const f32, nframes, channels, window.as_ptr(),
                                      window.len(), hop, output.as_mut_ptr(), output.len() / channels,
                                      scratch.as_mut_ptr(), scratch.len(), 0)
        })
    }
}

impl CudaFftImpl {
    /// `stft_magnitudes(samples, win_len, hop)` (src/visual/spectrogram.rs:52-76) with the magnitude
    /// and the running maximum fused behind the last FFT stage.  Returns (flat mags, frames, max).
    pub fn stft_magnitudes(&self, samples: &[f32], win_len: usize, hop: usize) -> Result<(Vec<f32>, usize, f32), FftError> {
        if hop == 0 {
            return Err(FftError::InvalidHopSize);
        }
        let nframes = samples.len().div_ceil(hop);
        let mut mags = vec![0.0f32; nframes * (win_len / 2)];
        let mut max_mag = 0.0f32;
        check(unsafe {
            kofft_cuda_stft_magnitudes_host_f32(self.ctx, samples.as_ptr(), samples.len(), win_len, hop, mags.as_mut_ptr(),
                                                nframes, &mut max_mag)
        })?;
        Ok((mags, nframes, max_mag))
    }
    /// `fft2d_inplace` (src/ndfft.rs:74-105)
    pub fn fft2d_inplace(&self, data: &mut [Complex32], rows: usize, cols: usize, scratch_col: &mut [Complex32]) -> Result<(), FftError> {
        check(unsafe { kofft_cuda_fft2d_host_f32(self.ctx, data.as_mut_ptr() as *mut f32, data.len(), rows, cols, scratch_col.len()) })
    }
    /// `fft3d_inplace` (src/ndfft.rs:114-156); scratch lengths = (depth, rows, cols)
    pub fn fft3d_inplace(&self, data: &mut [Complex32], depth: usize, rows: usize, cols: usize,
                         scratch_lens: (usize, usize, usize)) -> Result<(), FftError> {
        check(unsafe {
            kofft_cuda_fft3d_host_f32(self.ctx, data.as_mut_ptr() as *mut f32, data.len(), depth, rows, cols, scratch_lens.0,
                                      scratch_lens.1, scratch_lens.2)
        })
    }
    /// chunk size of the H2D / kernel / D2H pipeline behind the batch calls (0 = off, default 32 MiB)
    pub fn set_host_pipeline(&self, chunk_bytes: usize) {
        unsafe { kofft_cuda_set_host_pipeline(self.ctx, chunk_bytes) };
    }
}

impl Drop for CudaFftImpl {
    fn drop(&mut self) {
        unsafe { kofft_cuda_destroy(self.ctx) }
    }
}

impl FftImpl<f32> for CudaFftImpl {
    fn fft(&self, input: &mut [Complex32]) -> Result<(), FftError> {
        check(unsafe { kofft_cuda_fft_host_f32(self.ctx, input.as_mut_ptr() as *mut f32, input.len(), 0) })
    }
    fn ifft(&self, input: &mut [Complex32]) -> Result<(), FftError> {
        check(unsafe { kofft_cuda_fft_host_f32(self.ctx, input.as_mut_ptr() as *mut f32, input.len(), 1) })
    }
    fn fft_strided(&self, input: &mut [Complex32], stride: usize, scratch: &mut [Complex32]) -> Result<(), FftError> {
        check(unsafe {
            kofft_cuda_fft_strided_host_f32(self.ctx, input.as_mut_ptr() as *mut f32, input.len(), stride, scratch.len(), 0)
        })
    }
    fn ifft_strided(&self, input: &mut [Complex32], stride: usize, scratch: &mut [Complex32]) -> Result<(), FftError> {
        check(unsafe {
            kofft_cuda_fft_strided_host_f32(self.ctx, input.as_mut_ptr() as *mut f32, input.len(), stride, scratch.len(), 1)
        })
    }
    fn fft_out_of_place_strided(&self, input: &[Complex32], in_stride: usize, output: &mut [Complex32],
                                out_stride: usize) -> Result<(), FftError> {
        check(unsafe {
            kofft_cuda_fft_out_of_place_strided_host_f32(self.ctx, input.as_ptr() as *const f32, input.len(), in_stride,
                                                         output.as_mut_ptr() as *mut f32, output.len(), out_stride, 0)
        })
    }
    fn ifft_out_of_place_strided(&self, input: &[Complex32], in_stride: usize, output: &mut [Complex32],
                                 out_stride: usize) -> Result<(), FftError> {
        check(unsafe {
            kofft_cuda_fft_out_of_place_strided_host_f32(self.ctx, input.as_ptr() as *const f32, input.len(), in_stride,
                                                         output.as_mut_ptr() as *mut f32, output.len(), out_stride, 1)
        })
    }
    fn fft_with_strategy(&self, input: &mut [Complex32], _strategy: FftStrategy) -> Result<(), FftError> {
        // every strategy runs the faithful Stockham kernels (Radix2/SplitRadix/Auto already do in the reference)
        self.fft(input)
    }
    fn fft_split(&self, re: &mut [f32], im: &mut [f32]) -> Result<(), FftError> {
        check(unsafe { kofft_cuda_fft_split_host_f32(self.ctx, re.as_mut_ptr(), re.len(), im.as_mut_ptr(), im.len(), 0) })
    }
    fn ifft_split(&self, re: &mut [f32], im: &mut [f32]) -> Result<(), FftError> {
        check(unsafe { kofft_cuda_fft_split_host_f32(self.ctx, re.as_mut_ptr(), re.len(), im.as_mut_ptr(), im.len(), 1) })
    }
}

/// The f64 twin: kofft's `ScalarFftImpl<f64>` (src/fft.rs:914-1051) on the GPU, power-of-two lengths
/// 1..=8192, bit-identical, the whole trait surface (strided, out of place, split) on the device.
pub struct CudaFftImpl64 {
    ctx: *mut RawCtx,
}

unsafe impl Send for CudaFftImpl64 {}

impl CudaFftImpl64 {
    pub fn new(device: i32) -> Result<Self, String> {
        let mut ctx = core::ptr::null_mut();
        let rc = unsafe { kofft_cuda_create(&mut ctx, device) };
        if rc != 0 {
            let msg = unsafe { std::ffi::CStr::from_ptr(kofft_cuda_last_error()) };
            return Err(msg.to_string_lossy().into_owned());
        }
        Ok(Self { ctx })
    }
    /// `batch()` over dense rows `[batch][n]` in one launch.
    pub fn fft_batch(&self, rows: &mut [Complex64], n: usize, inverse: bool) -> Result<(), FftError> {
        if n == 0 {
            return Err(FftError::EmptyInput);
        }
        if rows.len() % n != 0 {
            return Err(FftError::MismatchedLengths);
        }
        check(unsafe {
            kofft_cuda_fft_batch_host_f64(self.ctx, rows.as_mut_ptr() as *mut f64, n, rows.len() / n, inverse as c_int)
        })
    }
}

impl CudaFftImpl64 {
    /// Fused pack + FFT + twist per row: `[batch][n]` reals -> `[batch][n/2+1]` bins (`rfft` / `irfft` one at a
    /// time arrive through kofft's blanket `RealFftImpl<f64>`, src/rfft.rs:837).
    pub fn rfft_batch(&self, input: &[f64], n: usize, output: &mut [Complex64]) -> Result<(), FftError> {
        if n == 0 {
            return Err(FftError::EmptyInput);
        }
        if input.len() % n != 0 || output.len() != input.len() / n * (n / 2 + 1) {
            return Err(FftError::MismatchedLengths);
        }
        check(unsafe {
            kofft_cuda_rfft_batch_host_f64(self.ctx, input.as_ptr(), n, input.len() / n, output.as_mut_ptr() as *mut f64)
        })
    }
    pub fn irfft_batch(&self, input: &[Complex64], n: usize, output: &mut [f64]) -> Result<(), FftError> {
        if n == 0 {
            return Err(FftError::EmptyInput);
        }
        if output.len() % n != 0 || input.len() != output.len() / n * (n / 2 + 1) {
            return Err(FftError::MismatchedLengths);
        }
        check(unsafe {
            kofft_cuda_irfft_batch_host_f64(self.ctx, input.as_ptr() as *const f64, n, output.len() / n, output.as_mut_ptr())
        })
    }
}

impl Drop for CudaFftImpl64 {
    fn drop(&mut self) {
        unsafe { kofft_cuda_destroy(self.ctx) }
    }
}

impl FftImpl<f64> for CudaFftImpl64 {
    fn fft(&self, input: &mut [Complex64]) -> Result<(), FftError> {
        check(unsafe { kofft_cuda_fft_host_f64(self.ctx, input.as_mut_ptr() as *mut f64, input.len(), 0) })
    }
    fn ifft(&self, input: &mut [Complex64]) -> Result<(), FftError> {
        check(unsafe { kofft_cuda_fft_host_f64(self.ctx, input.as_mut_ptr() as *mut f64, input.len(), 1) })
    }
    fn fft_strided(&self, input: &mut [Complex64], stride: usize, scratch: &mut [Complex64]) -> Result<(), FftError> {
        check(unsafe {
            kofft_cuda_fft_strided_host_f64(self.ctx, input.as_mut_ptr() as *mut f64, input.len(), stride, scratch.len(), 0)
        })
    }
    fn ifft_strided(&self, input: &mut [Complex64], stride: usize, scratch: &mut [Complex64]) -> Result<(), FftError> {
        check(unsafe {
            kofft_cuda_fft_strided_host_f64(self.ctx, input.as_mut_ptr() as *mut f64, input.len(), stride, scratch.len(), 1)
        })
    }
    fn fft_out_of_place_strided(&self, input: &[Complex64], in_stride: usize, output: &mut [Complex64],
                                out_stride: usize) -> Result<(), FftError> {
        check(unsafe {
            kofft_cuda_fft_out_of_place_strided_host_f64(self.ctx, input.as_ptr() as *const f64, input.len(), in_stride,
                                                         output.as_mut_ptr() as *mut f64, output.len(), out_stride, 0)
        })
    }
    fn ifft_out_of_place_strided(&self, input: &[Complex64], in_stride: usize, output: &mut [Complex64],
                                 out_stride: usize) -> Result<(), FftError> {
        check(unsafe {
            kofft_cuda_fft_out_of_place_strided_host_f64(self.ctx, input.as_ptr() as *const f64, input.len(), in_stride,
                                                         output.as_mut_ptr() as *mut f64, output.len(), out_stride, 1)
        })
    }
    fn fft_split(&self, re: &mut [f64], im: &mut [f64]) -> Result<(), FftError> {
        check(unsafe { kofft_cuda_fft_split_host_f64(self.ctx, re.as_mut_ptr(), re.len(), im.as_mut_ptr(), im.len(), 0) })
    }
    fn ifft_split(&self, re: &mut [f64], im: &mut [f64]) -> Result<(), FftError> {
        check(unsafe { kofft_cuda_fft_split_host_f64(self.ctx, re.as_mut_ptr(), re.len(), im.as_mut_ptr(), im.len(), 1) })
    }
    fn fft_with_strategy(&self, input: &mut [Complex64], _strategy: FftStrategy) -> Result<(), FftError> {
        self.fft(input)
    }
}

/// What `kofft::fft::new_fft_impl()` (src/fft.rs:1954-1985) would return with the `cuda` feature.
pub fn new_cuda_fft_impl(device: i32) -> Result<Box<dyn FftImpl<f32>>, String> {
    Ok(Box::new(CudaFftImpl::new(device)?))
}
