// launch.h -- type-erased launch interface between the C ABI (kofft_cuda.cu) and the
// per-size kernel translation units (fft_inst.cu is compiled once per L with -DKOFFT_L=<L>
// so the ten sizes build in parallel).
#pragma once
#include <cuda_runtime.h>

#include "fft_engine.cuh"

namespace kofft {

enum Kind : int {
    KIND_C2C_FWD = 0,
    KIND_C2C_INV = 1,
    KIND_GEN_FWD = 2,
    KIND_GEN_INV = 3,
    KIND_STFT = 4,
    KIND_ISTFT = 5,
    KIND_RFFT = 6,
    KIND_IRFFT = 7,
    KIND_STFT_MAG = 8,
    KIND_COUNT = 9
};

// Untyped operands; their meaning per kind is documented at the IO policy they feed
// (fft_kernels.cuh).
struct IoArgs {
    const void *in = nullptr;   // c2c/rfft/irfft/istft: input rows; gen: re; stft: signal
    const void *in2 = nullptr;  // gen: im
    void *out = nullptr;        // output rows; gen: re
    void *out2 = nullptr;       // gen: im
    const void *aux = nullptr;  // stft/istft: window; rfft/irfft: T' table
    long n = 0;                 // transform length of the complex core
    long p0 = 0, p1 = 0, p2 = 0, p3 = 0; // gen: in_es,in_rs,out_es,out_rs; stft: len,nframes,hop
    float scale = 1.0f;         // 1/n for inverse kinds
};

struct LaunchArgs {
    int kind = 0;
    bool exact = true;
    IoArgs io;
    Tw0 tw0;
    const float2 *table = nullptr; // device-resident FftPlanner table for n
    long rows = 0;
    int num_sms = 148;
    int max_ctas = 0; // 0 = occupancy * num_sms
    bool staged = false; // input rows may be fetched with TMA bulk copies (16-byte aligned, contiguous)
    cudaStream_t stream = nullptr;
};

// Function attributes (dynamic shared memory opt-in) and occupancy are per device: every kernel
// instantiation keeps one cached value per device so one process can drive several GPUs.
struct PerDevice {
    int v[64] = {};
    int &get()
    {
        int d = 0;
        cudaGetDevice(&d);
        return v[d & 63];
    }
};

// N = 32 .. 16384 (L = 5 .. 14)
cudaError_t launch_cta_fft(int L, const LaunchArgs &a);
// N = 1 .. 16
cudaError_t launch_small_fft(int n, const LaunchArgs &a);
// N = 2^15, 2^16: two kernels (column pass, row pass) over one chunk of the batch
struct LargeArgs {
    int lsub = 0;             // L - 8
    long row0 = 0;            // first transform of the chunk
    long chunk_rows = 0;      // transforms in the chunk
    float2 *scratch = nullptr; // two-kernel path: chunk_rows * 2^L complex, reused by every chunk
                               // fused path: clusters * 2 * 2^L complex (L2-resident either way)
    bool fused = true;         // one persistent cluster kernel instead of two kernels per chunk
    int max_clusters = 0;      // fused path: 0 = as many clusters as fit on the device
    bool stage_rows = true;    // two-kernel path: row pass prefetches its next tile with TMA bulk copies
    // pipelined persistent kernel (LargePipe): chunk_rows = every transform of the batch, scratch =
    // pipe_max_teams * kLargePipeSlots * 2^L complex, flags = pipe_max_teams * kPipeFlagStride counters
    bool pipe = false;
    int pipe_max_teams = 0;
    unsigned *flags = nullptr;
    int launches = 0;          // out: kernels launched
};
// upper bound on the CTAs of the pipelined kernel per SM (sizes its scratch); counters per team
constexpr int kMaxPipeCtasPerSm = 2;
constexpr int kPipeFlagStride = 32;
#ifndef KOFFT_PIPE_AHEAD
#define KOFFT_PIPE_AHEAD 2
#endif
#ifndef KOFFT_PIPE_SLOTS
#define KOFFT_PIPE_SLOTS (KOFFT_PIPE_AHEAD + 2)
#endif
constexpr int kLargePipeSlots = KOFFT_PIPE_SLOTS; // intermediate transforms per team (LargePipe::SLOTS)
// upper bound on the clusters the fused kernel runs with (sizes its scratch)
constexpr int kMaxFusedClusters = 148;
cudaError_t launch_large_fft(int L, const LaunchArgs &a, LargeArgs &g);

// N = 2^13 .. 2^15: the warp-specialised split kernel (fft_split32.cuh), one persistent cooperative launch
#ifndef KOFFT_SPLIT_SLOTS
#define KOFFT_SPLIT_SLOTS 4
#endif
constexpr int kSplitSlots = KOFFT_SPLIT_SLOTS; // intermediate transforms per team (Split32::SLOTS)
#ifndef KOFFT_SPLIT_ZSLOTS
#define KOFFT_SPLIT_ZSLOTS 4
#endif
constexpr int kSplitZSlots = KOFFT_SPLIT_ZSLOTS; // irfft at 2^15: untwisted rows per team (Split32::ZSLOTS)
struct SplitArgs {
    float2 v0[32] = {};        // pass-0 twiddles: v[(2^t - 1) + c] = T[c << (L-1-t)]  (Tw0W)
    float2 *scratch = nullptr; // max_teams * kSplitSlots * 2^L complex (stays in L2)
    unsigned *flags = nullptr; // max_teams * kPipeFlagStride counters
    int max_teams = 0;
    bool persist_l2 = false;   // mark the intermediate as a persisting-L2 access window for this launch
    bool pre_rows = false;     // scratch has room for max_teams * kSplitZSlots * 2^L more (irfft at 2^15: untwisted rows)
};
cudaError_t launch_split32_fft(int L, const LaunchArgs &a, SplitArgs &g);
// N = 8192 / 16384, dense C2C rows: one CTA per transform with 32 elements per thread (fft_wide.cuh);
// v0[(2^t - 1) + c] = T[c << (L-1-t)], t < 5
cudaError_t launch_wide_fft(int L, const LaunchArgs &a, const float2 *v0);

// power-of-two lengths above 2^16 (fft_huge.cu): 256-point column pass + register passes through two scratch buffers
constexpr int kHugeMaxLog2 = 27;
struct HugeArgs {
    float2 *scratch[2] = {nullptr, nullptr}; // chunk_rows * 2^L complex each
    long chunk_rows = 1;                      // transforms per trip through the scratch buffers
    int launches = 0;                         // out: kernels launched
};
cudaError_t launch_huge_fft(int L, const LaunchArgs &a, HugeArgs &g);

// per-L entry points (one per fft_inst.cu build)
#define KOFFT_DECL_L(L) cudaError_t launch_cta_fft_L##L(const LaunchArgs &a);
KOFFT_DECL_L(5) KOFFT_DECL_L(6) KOFFT_DECL_L(7) KOFFT_DECL_L(8) KOFFT_DECL_L(9)
KOFFT_DECL_L(10) KOFFT_DECL_L(11) KOFFT_DECL_L(12) KOFFT_DECL_L(13) KOFFT_DECL_L(14)
#undef KOFFT_DECL_L

// istft stage 2: ordered overlap-add gather + normalisation (src/stft.rs:142-154)
struct OlaArgs {
    const float *time;  // [channels][nframes][win_len] windowed real frames
    const float *window;
    float *output;      // [channels][out_len], accumulated into (reference semantics)
    float *norm;        // optional [channels][out_len] (the reference's `scratch`), may be null
    long channels, nframes, win_len, hop, out_len;
    int zero_uncovered; // 0: istft (leave sample untouched), 1: inverse_parallel (write 0)
};
cudaError_t launch_ola(const OlaArgs &a, cudaStream_t stream);

// Bluestein's element-wise steps (bluestein.cu): 0 = pre, 1 = mid, 2 = post
struct BluesteinArgs {
    const float2 *x = nullptr;     // [rows][n] input
    float2 *out = nullptr;         // [rows][n] output
    float2 *a = nullptr;           // [rows][m] workspace
    const float2 *chirp = nullptr; // n entries
    const float2 *bfft = nullptr;  // m entries: fft(b)
    long n = 0, m = 0, rows = 0;
    int inverse = 0;
    float scale_m = 1.0f, scale_n = 1.0f;
};
cudaError_t launch_bluestein_step(int step, const BluesteinArgs &b, bool exact, int num_sms, cudaStream_t s);
// the same for ScalarFftImpl<f64> (no FAST variant)
struct BluesteinArgsD {
    const double2 *x = nullptr;
    double2 *out = nullptr;
    double2 *a = nullptr;
    const double2 *chirp = nullptr;
    const double2 *bfft = nullptr;
    long n = 0, m = 0, rows = 0;
    int inverse = 0;
    double scale_m = 1.0, scale_n = 1.0;
};
cudaError_t launch_bluestein_step_f64(int step, const BluesteinArgsD &b, int num_sms, cudaStream_t s);
// f64 element-wise steps around a dense C2C core that the single-CTA f64 kernel does not cover (n > 8192 or not a power
// of two): gather / scatter of strided or split rows, rfft twist, irfft untwist (bluestein.cu)
struct ElementwiseArgsD {
    int op = 0;                    // EW_GATHER, EW_SCATTER, EW_UNTWIST, EW_TWIST
    long n = 0, rows = 0;
    const double *re = nullptr, *im = nullptr;
    double *out_re = nullptr, *out_im = nullptr;
    long es = 0, rs = 0;           // element and row strides in doubles
    double2 *a = nullptr;          // dense rows (gather / untwist / twist: destination; scatter: source)
    const double2 *x = nullptr;    // untwist: X [rows][m+1]; twist: Y [rows][m]
    const double2 *rtw = nullptr;
    int neg_im = 0;                // gather: negate the imaginary parts; scatter: negate them and ...
    double scale = 1.0;            // ... multiply both parts by this (ifft_split, src/fft.rs:1417-1425)
};
cudaError_t launch_elementwise_f64(const ElementwiseArgsD &e, int num_sms, cudaStream_t s);

// element-wise steps around a non-power-of-two core (bluestein.cu): the reference's gather / scatter, framing,
// untwist / twist and real * window loops, one kernel each
enum ElementwiseOp : int { EW_GATHER = 0, EW_SCATTER = 1, EW_FRAME = 2, EW_TIME = 3, EW_UNTWIST = 4, EW_TWIST = 5, EW_MAG = 6 };
struct ElementwiseArgs {
    int op = 0;
    long n = 0, rows = 0;          // core length (twist / untwist: m) and rows
    const float *re = nullptr, *im = nullptr; // gather: source planes; frame: re = signal
    float *out_re = nullptr, *out_im = nullptr; // scatter: destination planes; time: out_re = time frames
    long es = 0, rs = 0;           // gather / scatter: element and row strides in floats
    float2 *a = nullptr;           // the dense complex rows (gather / frame / untwist / twist: destination)
    const float2 *x = nullptr;     // untwist: X [rows][m+1]; twist: Y [rows][m]
    const float2 *rtw = nullptr;   // T' (src/rfft.rs:172-183)
    const float *aux_f = nullptr;  // frame / time: window
    int *max_bits = nullptr;       // mag: running maximum of the rows as float bits (x = frames, out_re = magnitudes)
    long len = 0, nframes = 0, hop = 0;
};
cudaError_t launch_elementwise(const ElementwiseArgs &e, bool exact, int num_sms, cudaStream_t s);

// f64 twin (fft_f64.cuh): FftImpl<f64>::fft / ifft, N = 1 .. 8192 (power of two), contiguous rows
struct LaunchF64Args {
    const double2 *in = nullptr;
    double2 *out = nullptr;
    long n = 0, rows = 0;
    bool inverse = false;
    double scale = 1.0;               // 1/n for the inverse
    const double2 *table = nullptr;   // device-resident FftPlanner<f64> table for n (n >= 32)
    Tw0D tw0 = {};                    // pass-0 twiddles: v[(2^t - 1) + c] = T[c << (L-1-t)]
    int num_sms = 148, max_ctas = 0;
    bool staged = true;               // TMA prefetch of the next row group into the idle exchange buffer
    // generic (strided / split) addressing instead of dense interleaved rows: element e of row r at
    // re[r*rs + e*es], strides in doubles
    bool generic = false;
    const double *in_re = nullptr, *in_im = nullptr;
    double *out_re = nullptr, *out_im = nullptr;
    long in_es = 0, in_rs = 0, out_es = 0, out_rs = 0;
    // real transforms: 1 = rfft (in: [rows][m] complex view of the reals, out: [rows][m+1]), 2 = irfft; n = m
    int real = 0;
    const double2 *rtw = nullptr;     // T'[k] = exp(-i pi k / m), m entries
    cudaStream_t stream = nullptr;
};
cudaError_t launch_fft_f64(const LaunchF64Args &a);
// f64 above 8192 points: register passes through two scratch buffers (fft_huge.cu), dense rows
constexpr int kHugeMaxLog2F64 = 26;
cudaError_t launch_huge_fft_f64(int L, const LaunchF64Args &a, double2 *scratch0, double2 *scratch1, int *launches);

// fused single-kernel istft (istft_fused.cuh), N = 512 .. 4096
struct IstftFusedArgs;
cudaError_t launch_istft_fused(int L, const LaunchArgs &a, const IstftFusedArgs &f);

} // namespace kofft
