/*
 * kofft_cuda.h -- C ABI of libkofft_cuda.so: the B200 (sm_100a) backend for the batched
 * FFT / rfft / STFT hot path of okian/kofft.
 *
 * This is the boundary a `kofft-cuda` crate binds (`extern "C"`); its `CudaFftImpl`
 * implements kofft's `FftImpl<f32>` trait on top of the host-pointer entry points and
 * exposes the device-pointer batched entry points as inherent methods.  Every function
 * cites the reference interface it replaces (paths relative to the kofft repository).
 *
 * Conventions
 *  - Complex data is interleaved {re, im} f32 == `#[repr(C)] Complex<f32>` (src/num.rs:105-110).
 *  - Return value: 0 = Ok; 1..6 = kofft's `FftError` variants in declaration order
 *    (src/fft.rs:446-454); negative = -(cudaError_t) with text in kofft_cuda_last_error().
 *  - Lengths: any power of two up to 2^27 (rfft / irfft: 2^28) on one GPU; above 2^16 the transform makes
 *    several trips through global memory (fft_huge.cu).  Non-power-of-two lengths n <= 2^26 take the
 *    reference's Bluestein path (src/fft.rs:411-433, 1083-1132: chirp, two transforms of length
 *    next_pow2(2n-1)) with the same tables and arithmetic -- from fft / ifft and, as in the reference's
 *    std build, from the rfft / irfft / stft / istft / split / strided / magnitude entry points and the
 *    device-resident streams, which all end in fft.fft().  KOFFT_ERR_NON_POWER_OF_TWO_NO_STD (the no_std build's
 *    answer) is therefore never returned.
 *  - Host-pointer functions are synchronous and never retain the caller's pointers.
 *    Device-pointer functions are stream-ordered on `stream`, a cudaStream_t passed as
 *    void* with CUDA's own meaning (NULL = the legacy default stream; kofft_cuda_stream()
 *    returns the context's private stream), and return as soon as the work is enqueued.
 *  - A context belongs to one device and is not thread-safe (the reference's
 *    ScalarFftImpl is !Sync for the same reason, src/fft.rs:589-605).
 *  - There is no CPU fallback: without a CUDA device kofft_cuda_create fails.
 */
#ifndef KOFFT_CUDA_H
#define KOFFT_CUDA_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* FftError, src/fft.rs:446-454 */
#define KOFFT_OK 0
#define KOFFT_ERR_EMPTY_INPUT 1
#define KOFFT_ERR_NON_POWER_OF_TWO_NO_STD 2
#define KOFFT_ERR_MISMATCHED_LENGTHS 3
#define KOFFT_ERR_INVALID_STRIDE 4
#define KOFFT_ERR_INVALID_HOP_SIZE 5
#define KOFFT_ERR_INVALID_VALUE 6

/* window kinds for kofft_cuda_window_host_f32 (src/window.rs:24-61) */
#define KOFFT_WINDOW_HANN 0
#define KOFFT_WINDOW_HAMMING 1
#define KOFFT_WINDOW_BLACKMAN 2
#define KOFFT_WINDOW_KAISER 3

typedef struct kofft_cuda_ctx kofft_cuda_ctx;

/* ---- context: the device twin of ScalarFftImpl + FftPlanner (src/fft.rs:332-445, 600-632) */
int kofft_cuda_create(kofft_cuda_ctx **out, int device);
void kofft_cuda_destroy(kofft_cuda_ctx *ctx);
const char *kofft_cuda_last_error(void);
int kofft_cuda_device(const kofft_cuda_ctx *ctx);
void *kofft_cuda_stream(const kofft_cuda_ctx *ctx); /* cudaStream_t */
int kofft_cuda_synchronize(kofft_cuda_ctx *ctx);
/* exact != 0 (default): every product and sum rounded as in the reference -> bit-identical
 * results.  exact == 0: packed FMUL2/FFMA2 butterflies (about 1e-7 relative difference). */
int kofft_cuda_set_exact(kofft_cuda_ctx *ctx, int exact);
int kofft_cuda_get_exact(const kofft_cuda_ctx *ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
unsigned long long kofft_cuda_launch_count(const kofft_cuda_ctx *ctx);
/* enable (default) / disable the TMA-staged input prefetch (cp.async.bulk + mbarrier); the
 * library falls back to plain loads by itself when a pointer is not 16-byte aligned */
int kofft_cuda_set_tma_staging(kofft_cuda_ctx *ctx, int enable);
/* istft, window 512..4096: enable (default) / disable the single fused kernel (inverse FFT +
 * window + ordered overlap-add + normalisation); run_frames > 0 sets the frames a CTA owns per
 * run (default 128).  Disabled or out of range -> two kernels with an f32 intermediate. */
int kofft_cuda_set_istft_fusion(kofft_cuda_ctx *ctx, int enable, int run_frames);
/* N > 16384 (rfft above 32768): which of the three implementations of the two-pass split runs.
 *   mode 3 (default): mode 2 for rfft and irfft (where it measured faster: rfft 2^16 4.06 vs 4.29 ms, irfft 4.40 vs
 *          6.11 ms, profiles/r02a, r02p), mode 0 otherwise.
 *   mode 2: one persistent cooperative kernel; teams of 8 / 16 CTAs run pass A two transforms ahead
 *          of pass B behind dependency flags, and the intermediate (4 transforms per team) is pinned
 *          in L2, so HBM sees the rows once in and once out.
 *   mode 0: two kernels (column pass, row pass) per 256 MB batch chunk.
 *   mode 1: one persistent thread-block-cluster kernel (cluster barrier between the passes).
 * All are bit-identical. */
int kofft_cuda_set_large_mode(kofft_cuda_ctx *ctx, int mode);
/* Lengths 2^min_log2n .. 2^15 of the complex core (rfft / irfft: twice that) run the warp-specialised split
 * kernel (fft_split32.cuh): ONE persistent cooperative launch, 512-thread CTAs whose "A warps" run the
 * 2^(L-5)-point column transforms of an 8192-element tile while their "B warps" finish 32-point rows in
 * registers (rfft twist by warp shuffles), intermediate pinned in L2.  Default 14 (2^14 and 2^15: where it measured
 * faster than the single-CTA / two-pass kernels); 13 adds 2^13, 16 switches it off (kofft_cuda_set_large_mode then
 * picks the implementation for 2^15).  Bit-identical. */
int kofft_cuda_set_split_min_log2n(kofft_cuda_ctx *ctx, int min_log2n);
/* By default the split kernel serves C2C and rfft (2^15 cores: tiles arrive by TMA tensor-map loads).  all_kinds != 0
 * also routes irfft and the strided / split (SoA) entry points through it (plain loads; measured slower than the
 * older paths, kept for parity testing). */
int kofft_cuda_set_split_all_kinds(kofft_cuda_ctx *ctx, int all_kinds);
/* Complex cores of 8192 / 16384 points can run the wide single-CTA kernel (fft_wide.cuh: 32 elements per thread, the row
 * lands by TMA bulk copies in the one exchange buffer, two CTAs per SM at 8192; default: dense C2C, rfft and SoA /
 * strided rows at both lengths, irfft at 2^13).  Bit (L - 13) + 2 g of `mask` enables it
 * for 2^L-point cores of group g: 0 dense C2C rows, 1 rfft, 2 irfft, 3 SoA / strided rows; a mask with bit 31 set
 * restores the default (the kinds where it measured faster).  0 hands everything back to the split / single-CTA kernels.
 * Bit-identical. */
int kofft_cuda_set_wide_mask(kofft_cuda_ctx *ctx, unsigned mask);
/* The persistent large-N kernels need a cooperative launch (every CTA resident).  When the device cannot grant
 * it (shared or partitioned GPU) the library computes the same bits with the slower multi-kernel path, bumps
 * this counter and leaves a note in kofft_cuda_last_error(). */
unsigned long long kofft_cuda_fallback_count(const kofft_cuda_ctx *ctx);
/* enable != 0: mode 1 above; 0: back to the current non-cluster mode */
int kofft_cuda_set_cluster_fusion(kofft_cuda_ctx *ctx, int enable);
/* Host-pointer batch entry points (fft_batch_host, rfft_batch_host, irfft_batch_host): batches
 * larger than chunk_bytes are cut into chunks of about chunk_bytes that flow through three
 * streams (H2D copy engine, kernels, D2H copy engine), so both PCIe directions and the SMs work
 * at the same time.  Default 32 MiB; 0 = one copy in, one launch, one copy out.  Pin the host
 * buffers (cudaHostRegister) for full PCIe rate; pageable buffers work but copy synchronously. */
int kofft_cuda_set_host_pipeline(kofft_cuda_ctx *ctx, size_t chunk_bytes);
/* 0 = let the library size the grid (occupancy x SM count); otherwise cap the CTA count */
int kofft_cuda_set_max_ctas(kofft_cuda_ctx *ctx, int max_ctas);

/* ---- planner tables -------------------------------------------------------------------- */
/* FftPlanner::get_twiddles (src/fft.rs:370-408): n/2 complex, f32 recurrence, bit-exact. */
int kofft_cuda_twiddles_host_f32(size_t n, float *out);
/* build_twiddle_table (src/rfft.rs:172-183): m complex; fma_mul selects the
 * `target-feature=+fma` flavour of Complex::mul (src/num.rs:173-178). */
int kofft_cuda_rfft_twiddles_host_f32(size_t m, float *out, int fma_mul);
/* device-resident copies, created on first use, pointer-stable for the context's life
 * (the reference's Arc<[Complex<T>]> cache, tests/rfft_twiddles.rs:13-15). */
int kofft_cuda_get_twiddles(kofft_cuda_ctx *ctx, size_t n, const void **dev_ptr);
int kofft_cuda_get_rfft_twiddles(kofft_cuda_ctx *ctx, size_t m, const void **dev_ptr);
/* which Complex::mul flavour builds the rfft table (default 0 = default cargo build) */
int kofft_cuda_set_rfft_table_fma(kofft_cuda_ctx *ctx, int fma_mul);
/* hann / hamming / blackman / kaiser (src/window.rs:24-61); beta only for kaiser */
int kofft_cuda_window_host_f32(int kind, size_t len, float beta, float *out);

/* ---- device-pointer batched entry points ------------------------------------------------ */
/* batch() / batch_inverse() over dense rows [batch][n] (src/fft.rs:2156-2175), each row one
 * FftImpl::fft / ifft (src/fft.rs:1054-1174).  d_in == d_out is allowed. */
int kofft_cuda_fft_c2c_f32(kofft_cuda_ctx *ctx, const void *d_in, void *d_out, size_t n, size_t batch,
                           int inverse, void *stream);
/* fft_out_of_place_strided generalised to a batch (src/fft.rs:1260-1336): element e of row r
 * is at d_in[r*in_dist + e*in_stride] (units: complex elements). */
int kofft_cuda_fft_strided_f32(kofft_cuda_ctx *ctx, const void *d_in, size_t in_stride, size_t in_dist,
                               void *d_out, size_t out_stride, size_t out_dist, size_t n, size_t batch,
                               int inverse, void *stream);
/* fft2d_inplace / fft3d_inplace (src/ndfft.rs:74-153): row-major data, in place; the column /
 * depth passes are the strided entry point batched over all columns. */
int kofft_cuda_fft2d_f32(kofft_cuda_ctx *ctx, void *d_data, size_t rows, size_t cols, void *stream);
int kofft_cuda_fft3d_f32(kofft_cuda_ctx *ctx, void *d_data, size_t depth, size_t rows, size_t cols, void *stream);
/* fft_split / ifft_split (src/fft.rs:1365-1439): SoA rows [batch][n]. */
int kofft_cuda_fft_split_f32(kofft_cuda_ctx *ctx, const float *d_in_re, const float *d_in_im, float *d_out_re,
                             float *d_out_im, size_t n, size_t batch, int inverse, void *stream);
/* rfft_with_scratch per row (src/rfft.rs:425-465): in [batch][n] f32, out [batch][n/2+1] complex. */
int kofft_cuda_rfft_f32(kofft_cuda_ctx *ctx, const float *d_in, void *d_out, size_t n, size_t batch,
                        void *stream);
/* irfft_with_scratch per row (src/rfft.rs:468-508): in [batch][n/2+1] complex, out [batch][n] f32. */
int kofft_cuda_irfft_f32(kofft_cuda_ctx *ctx, const void *d_in, float *d_out, size_t n, size_t batch,
                         void *stream);
/* stft() per channel (src/stft.rs:76-105): signal [channels][len], window [win_len],
 * frames [channels][nframes][win_len] complex; nframes >= ceil(len/hop). */
int kofft_cuda_stft_f32(kofft_cuda_ctx *ctx, const float *d_signal, size_t len, size_t channels,
                        const float *d_window, size_t win_len, size_t hop, void *d_frames, size_t nframes,
                        void *stream);
/* stft_magnitudes() per channel (src/visual/spectrogram.rs:52-76): the only in-tree consumer of
 * large STFTs keeps |X[k]| for k < win_len/2 and the largest magnitude.  Fused behind the last
 * FFT stage: mags [channels][nframes][win_len/2] f32 (8x fewer output bytes than the complex
 * frames), d_max [channels] f32.  win_len >= 32. */
int kofft_cuda_stft_magnitudes_f32(kofft_cuda_ctx *ctx, const float *d_signal, size_t len, size_t channels,
                                   const float *d_window, size_t win_len, size_t hop, float *d_mags,
                                   size_t nframes, float *d_max, void *stream);
/* istft() per channel (src/stft.rs:117-156): frames [channels][nframes][win_len] (left
 * untouched), output [channels][out_len] is ACCUMULATED into as in the reference, d_norm
 * (optional, [channels][out_len]) receives the reference's `scratch` (sum of window^2).
 * zero_uncovered != 0 selects inverse_parallel's variant (src/stft.rs:335-341). */
int kofft_cuda_istft_f32(kofft_cuda_ctx *ctx, const void *d_frames, size_t nframes, size_t channels,
                         const float *d_window, size_t win_len, size_t hop, float *d_output, size_t out_len,
                         float *d_norm, int zero_uncovered, void *stream);

/* ---- host-pointer drop-ins: what `impl FftImpl<f32> for CudaFftImpl` calls ---------------- */
/* FftImpl::fft / ifft (src/fft.rs:467-468): in place on n complex. */
int kofft_cuda_fft_host_f32(kofft_cuda_ctx *ctx, float *data, size_t n, int inverse);
/* batch() / batch_inverse() on dense rows (src/fft.rs:2156-2175). */
int kofft_cuda_fft_batch_host_f32(kofft_cuda_ctx *ctx, float *data, size_t n, size_t batch, int inverse);
/* ---- f64 twin: what `impl FftImpl<f64> for CudaFftImpl64` calls (ScalarFftImpl<f64>, src/fft.rs:914-1051
 * behind the same dispatch :1054-1082 and ifft :1134-1174).  Elements are interleaved doubles = #[repr(C)]
 * Complex<f64>.  Every entry point takes every power of two up to 2^26 complex points and, through the reference's
 * Bluestein path with T = f64 (:411-433, 1083-1132), every other length up to 2^25: powers of two up to 8192 in one
 * fused kernel, the rest through the dense C2C core with the reference's own gather / scatter / twist / untwist loops
 * around it (src/fft.rs:921-933, 1191-1197; src/rfft.rs:425-508).  Beyond that: negative (not supported). */
/* FftPlanner::<f64>::get_twiddles(n) (src/fft.rs:391-405 with T = f64): n/2 complex doubles */
int kofft_cuda_twiddles_host_f64(size_t n, double *out);
/* batched, device pointers, stream-ordered, in place allowed */
int kofft_cuda_fft_c2c_f64(kofft_cuda_ctx *ctx, const void *d_in, void *d_out, size_t n, size_t batch, int inverse,
                           void *stream);
/* strided rows of interleaved complex doubles / split (SoA) rows, device pointers, as the f32 twins */
int kofft_cuda_fft_strided_f64(kofft_cuda_ctx *ctx, const void *d_in, size_t in_stride, size_t in_dist, void *d_out,
                               size_t out_stride, size_t out_dist, size_t n, size_t batch, int inverse, void *stream);
int kofft_cuda_fft_split_f64(kofft_cuda_ctx *ctx, const double *d_in_re, const double *d_in_im, double *d_out_re,
                             double *d_out_im, size_t n, size_t batch, int inverse, void *stream);
/* RealFftImpl<f64> (src/rfft.rs:775-837 -> rfft_direct / irfft_direct :425-508 with T = f64), batched; the
 * table is build_twiddle_table::<f64> (:172-183).  n even, n/2 a power of two <= 8192; errors as the f32 twins. */
int kofft_cuda_rfft_twiddles_host_f64(size_t m, double *out);
int kofft_cuda_rfft_f64(kofft_cuda_ctx *ctx, const double *d_in, void *d_out, size_t n, size_t batch, void *stream);
int kofft_cuda_irfft_f64(kofft_cuda_ctx *ctx, const void *d_in, double *d_out, size_t n, size_t batch, void *stream);
int kofft_cuda_rfft_batch_host_f64(kofft_cuda_ctx *ctx, const double *input, size_t n, size_t batch, double *output);
int kofft_cuda_irfft_batch_host_f64(kofft_cuda_ctx *ctx, const double *input, size_t n, size_t batch, double *output);
/* FftImpl::<f64>::fft / ifft in place on n complex doubles; batch variant on dense rows */
int kofft_cuda_fft_host_f64(kofft_cuda_ctx *ctx, double *data, size_t n, int inverse);
int kofft_cuda_fft_batch_host_f64(kofft_cuda_ctx *ctx, double *data, size_t n, size_t batch, int inverse);
/* FftImpl::<f64>::fft_split / ifft_split (src/fft.rs:1365-1439; KAT tests/split64.rs),
 * fft_strided / ifft_strided (:1175-1259), fft_out_of_place_strided / ifft_... (:1260-1336) */
int kofft_cuda_fft_split_host_f64(kofft_cuda_ctx *ctx, double *re, size_t re_len, double *im, size_t im_len, int inverse);
int kofft_cuda_fft_strided_host_f64(kofft_cuda_ctx *ctx, double *input, size_t input_len, size_t stride, size_t n,
                                    int inverse);
int kofft_cuda_fft_out_of_place_strided_host_f64(kofft_cuda_ctx *ctx, const double *input, size_t input_len,
                                                 size_t in_stride, double *output, size_t output_len, size_t out_stride,
                                                 int inverse);
/* FftImpl::fft_split / ifft_split (src/fft.rs:556-586, 1365-1439). */
int kofft_cuda_fft_split_host_f32(kofft_cuda_ctx *ctx, float *re, size_t re_len, float *im, size_t im_len,
                                  int inverse);
/* FftImpl::fft_strided / ifft_strided (src/fft.rs:494-506, 1175-1199): input_len complex
 * elements, n = scratch.len(). */
int kofft_cuda_fft_strided_host_f32(kofft_cuda_ctx *ctx, float *input, size_t input_len, size_t stride,
                                    size_t n, int inverse);
/* FftImpl::fft_out_of_place_strided / ifft_... (src/fft.rs:508-522, 1260-1336). */
int kofft_cuda_fft_out_of_place_strided_host_f32(kofft_cuda_ctx *ctx, const float *input, size_t input_len,
                                                 size_t in_stride, float *output, size_t output_len,
                                                 size_t out_stride, int inverse);
/* fft2d_inplace(data, rows, cols, fft, scratch_col) / fft3d_inplace (src/ndfft.rs:74-153); the
 * scratch lengths are checked like the reference's (the scratch itself is not needed). */
int kofft_cuda_fft2d_host_f32(kofft_cuda_ctx *ctx, float *data, size_t data_len, size_t rows, size_t cols,
                              size_t scratch_col_len);
int kofft_cuda_fft3d_host_f32(kofft_cuda_ctx *ctx, float *data, size_t data_len, size_t depth, size_t rows,
                              size_t cols, size_t tube_len, size_t row_len, size_t col_len);
/* RealFftImpl::rfft_with_scratch (src/rfft.rs:780-788): output_len must be n/2+1 and
 * scratch_len >= n/2 (the scratch itself is not needed on the GPU; its length is checked
 * so error behaviour matches). */
int kofft_cuda_rfft_host_f32(kofft_cuda_ctx *ctx, const float *input, size_t n, float *output,
                             size_t output_len, size_t scratch_len);
int kofft_cuda_rfft_batch_host_f32(kofft_cuda_ctx *ctx, const float *input, size_t n, size_t batch,
                                   float *output);
/* RealFftImpl::irfft_with_scratch (src/rfft.rs:809-817). */
int kofft_cuda_irfft_host_f32(kofft_cuda_ctx *ctx, const float *input, size_t input_len, float *output,
                              size_t n, size_t scratch_len);
int kofft_cuda_irfft_batch_host_f32(kofft_cuda_ctx *ctx, const float *input, size_t n, size_t batch,
                                    float *output);
/* stft() (src/stft.rs:76-105) with `output` flattened to [nframes][win_len] complex;
 * channels > 1 processes [channels][len] -> [channels][nframes][win_len]. */
int kofft_cuda_stft_host_f32(kofft_cuda_ctx *ctx, const float *signal, size_t len, size_t channels,
                             const float *window, size_t win_len, size_t hop, float *frames, size_t nframes);
/* stft_magnitudes(samples, win_len, hop) (src/visual/spectrogram.rs:52-76): Hann window,
 * nframes = ceil(len / hop) rows of win_len/2 magnitudes, *max_mag = the largest one. */
int kofft_cuda_stft_magnitudes_host_f32(kofft_cuda_ctx *ctx, const float *samples, size_t len, size_t win_len,
                                        size_t hop, float *mags, size_t nframes, float *max_mag);
/* istft() (src/stft.rs:117-156); scratch receives the window-power sums like the reference.
 * zero_uncovered != 0: inverse_parallel semantics (scratch may then be NULL). */
int kofft_cuda_istft_host_f32(kofft_cuda_ctx *ctx, const float *frames, size_t nframes, size_t channels,
                              const float *window, size_t win_len, size_t hop, float *output, size_t out_len,
                              float *scratch, size_t scratch_len, int zero_uncovered);

/* ---- streaming STFT / ISTFT with device-resident state -----------------------------------------
 * Device twins of StftStream (src/stft.rs:160-206) and IstftStream (src/stft.rs:407-520): the
 * reference's streams are host structs around one fft per frame; here the carried samples / the open
 * frame tails live on the device and one push handles any number of samples or frames for many
 * channels with the batched kernels.  What a stream delivers is bit-identical to the offline
 * stft / istft of the whole signal (the property the reference tests, tests/istft_stream.rs). */
typedef struct kofft_cuda_stft_stream kofft_cuda_stft_stream;
typedef struct kofft_cuda_istft_stream kofft_cuda_istft_stream;
/* window: host pointer, copied.  hop == 0 -> InvalidHopSize (StftStream::new :178-180). */
int kofft_cuda_stft_stream_create(kofft_cuda_ctx *ctx, size_t channels, const float *window, size_t win_len, size_t hop,
                                  kofft_cuda_stft_stream **out);
void kofft_cuda_stft_stream_destroy(kofft_cuda_stft_stream *s);
/* frames the next push of n samples per channel emits (flush != 0: the remaining, zero-padded ones) */
size_t kofft_cuda_stft_stream_frames(const kofft_cuda_stft_stream *s, size_t n, int flush);
/* d_samples [channels][n] (row stride ld floats) -> d_frames [channels][*nframes_out][win_len] complex, dense;
 * frames_cap = frames per channel d_frames can hold.  flush != 0 after the last samples (next_frame() keeps
 * returning frames while pos < len, zero-padded past the end, :193-199). */
int kofft_cuda_stft_stream_push(kofft_cuda_stft_stream *s, const float *d_samples, size_t n, size_t ld, void *d_frames,
                                size_t frames_cap, size_t *nframes_out, int flush, void *stream);
/* hop == 0 -> InvalidHopSize (IstftStream::new :434-436). */
int kofft_cuda_istft_stream_create(kofft_cuda_ctx *ctx, size_t channels, const float *window, size_t win_len, size_t hop,
                                   kofft_cuda_istft_stream **out);
void kofft_cuda_istft_stream_destroy(kofft_cuda_istft_stream *s);
/* push_frame() for k frames per channel at once (:449-495): d_frames [channels][k][win_len] complex, dense ->
 * the next k * hop normalised samples of every channel in d_out (row stride ld floats).  flush != 0 (:497-519):
 * the win_len - hop samples after the last frame (0 before the first frame or the second time). */
int kofft_cuda_istft_stream_push(kofft_cuda_istft_stream *s, const void *d_frames, size_t k, float *d_out, size_t ld,
                                 size_t *nsamples_out, int flush, void *stream);

/* ---- one C2C transform sharded over the GPUs of a box (BASELINE configs[4]) -------------------
 * No reference equivalent: kofft is single-process CPU code and `FftPlanner::get_twiddles`
 * (src/fft.rs:391-405) degenerates at these sizes (cos(2 pi / 2^30) rounds to 1.0f), so this
 * path uses correctly rounded twiddles and is validated against f64 instead.
 *
 * N = 2^log2n = N1 * N2 (N1 = 2^(log2n/2)); rank g of `world` (a power of two <= 16) owns the
 * contiguous slice [g N/world, (g+1) N/world) of the input, i.e. rows [g N1/world, ...) of x
 * viewed as [N1][N2].  The transform is four phases per rank; the all-to-all exchanges between
 * the phases are performed by the kernels themselves as peer-to-peer stores over NVLink into the
 * other ranks' buffers (no NCCL call, no staging).  The CALLER must place a barrier between
 * phases: every rank finishes phase p (stream synchronize) before any rank starts phase p + 1.
 *   natural_order == 0: d_out [N1/world][N2] holds X[(g N1/world + r) + N1 k2] at [r][k2]
 *                       ("transposed" spectrum, two exchanges)
 *   natural_order != 0: d_out holds the contiguous slice X[g N/world ...] (three exchanges); pass
 *                       d_out == NULL to leave it in buffer A (kofft_cuda_dist_buffer(d, 0)) and
 *                       save the final device copy
 * One process per GPU: create, exchange the 128-byte IPC handles of all ranks (any transport),
 * connect_ipc.  One process driving several GPUs: create one per device, connect_local, then
 * kofft_cuda_dist_run_local runs all phases with device-synchronising barriers. */
typedef struct kofft_cuda_dist kofft_cuda_dist;
int kofft_cuda_dist_create(kofft_cuda_ctx *ctx, int rank, int world, int log2n, kofft_cuda_dist **out);
void kofft_cuda_dist_destroy(kofft_cuda_dist *d);
size_t kofft_cuda_dist_shard_len(const kofft_cuda_dist *d); /* complex elements per rank */
/* the local transforms of a phase run in `pieces` (default 4, 1..8) pieces, each scattered to the
 * peers on a second stream while the next one is transformed (NVLink traffic overlaps compute) */
int kofft_cuda_dist_set_pieces(kofft_cuda_dist *d, int pieces);
void *kofft_cuda_dist_buffer(const kofft_cuda_dist *d, int which); /* 0: A, 1: B (device) */
int kofft_cuda_dist_ipc_handles(kofft_cuda_dist *d, void *out128);
int kofft_cuda_dist_connect_ipc(kofft_cuda_dist *d, const void *all_handles /* world * 128 bytes, rank order */);
int kofft_cuda_dist_connect_local(kofft_cuda_dist *const *dists, int world);
int kofft_cuda_dist_phase(kofft_cuda_dist *d, int phase /* 0..3 */, const void *d_in, void *d_out, int inverse,
                          int natural_order, void *stream);
/* The same transform with a LIBRARY collective for the exchanges (NCCL all-to-all): the baseline the P2P-store
 * exchange above is measured against (bench.py reports both).  kofft_cuda_dist_pack transposes (step 1: and
 * twiddles) the local matrix into a send buffer laid out by destination rank; the caller runs the all-to-all and
 * copies block s of the receive buffer, [cb][rows], to columns [s*rows, (s+1)*rows) of the destination buffer
 * [cb][world*rows] (buffer A after steps 0 and 2, buffer B after step 1); kofft_cuda_dist_local_fft runs the local
 * transforms of buffer A (which = 0) / B (which = 1) in place.  kofft_b200/dist.py: DistFft.transform_collective. */
int kofft_cuda_dist_pack(kofft_cuda_dist *d, int step /* 0..2 */, const void *d_src, void *d_send, int inverse, void *stream);
int kofft_cuda_dist_local_fft(kofft_cuda_dist *d, int which /* 0: A, 1: B */, int inverse, void *stream);
int kofft_cuda_dist_run_local(kofft_cuda_dist *const *dists, int world, const void *const *d_in,
                              void *const *d_out, int inverse, int natural_order);

#ifdef __cplusplus
}
#endif
#endif /* KOFFT_CUDA_H */
