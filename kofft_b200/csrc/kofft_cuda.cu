// kofft_cuda.cu -- the C ABI (include/kofft_cuda.h): context, device-resident planner tables,
// argument checks that mirror the reference's error behaviour, and kernel dispatch.
// the library is built with -fvisibility=hidden; only the C ABI is exported
#pragma GCC visibility push(default)
#include "../../include/kofft_cuda.h"
#pragma GCC visibility pop

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "dist_kernels.h"
#include "host_tables.h"
#include "istft_fused.cuh"
#include "launch.h"

using namespace kofft;

namespace {

thread_local std::string g_last_error;

int fail_cuda(cudaError_t e, const char *where)
{
    (void)cudaGetLastError(); // do not let a reported error leak into the next call's launch check
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
    g_last_error = buf;
    int code = -static_cast<int>(e);
    return code == 0 ? -1 : code;
}

int fail_msg(int code, const char *msg)
{
    g_last_error = msg;
    return code;
}

#define CU(call)                                          \
    do {                                                  \
        cudaError_t e__ = (call);                         \
        if (e__ != cudaSuccess) return fail_cuda(e__, #call); \
    } while (0)

bool is_pow2(size_t n) { return n != 0 && (n & (n - 1)) == 0; }
int log2_of(size_t n)
{
    int l = 0;
    while ((size_t(1) << l) < n) l++;
    return l;
}

struct Table {
    std::vector<float> host; // interleaved; dropped after the upload for n > 2^20 (pre0 keeps what the host still needs)
    float2 *dev = nullptr;
    float2 pre0[16] = {};    // pass-0 twiddles of a radix-16 first pass: pre0[(2^t - 1) + c] = T[c << (L-1-t)], t < 4
};

} // namespace

// default of kofft_cuda_set_wide_mask: bit (L - 13) + 2 g, g = 0 dense C2C, 1 rfft, 2 irfft, 3 SoA / strided rows
constexpr size_t kSmallHostBytes = 64 * 1024; // host-pointer C2C calls up to this size run in place on mapped host memory
constexpr unsigned kWideDefault = 0xDFu; // dense C2C, rfft and SoA / strided rows at both lengths, irfft at 2^13 (measured: profiles/r04r, r05a, r05e)

struct kofft_cuda_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int num_sms = 148;
    bool exact = true;
    bool rfft_fma = false;
    int max_ctas = 0;
    unsigned long long launches = 0;
    std::map<std::pair<size_t, int>, Table> fft_tables;  // key (n, accurate)
    struct Blue {
        float2 *chirp = nullptr, *bfft = nullptr;
        size_t m = 0;
    };
    std::map<size_t, Blue> blue_tables; // key n (non-power-of-two): the planner's bluestein_cache
    struct BlueD {
        double2 *chirp = nullptr, *bfft = nullptr;
        size_t m = 0;
    };
    std::map<size_t, BlueD> blue_tables_f64; // FftPlanner<f64>::bluestein_cache
    bool accurate_tables = false; // true: correctly rounded roots of unity instead of the reference's recurrence
    std::map<std::pair<size_t, int>, Table> rfft_tables; // key (m, fma)
    struct TableD {
        std::vector<double> host; // interleaved
        double2 *dev = nullptr;
    };
    std::map<size_t, TableD> fft_tables_f64;  // FftPlanner<f64>
    std::map<size_t, TableD> rfft_tables_f64; // RfftPlanner<f64>, key m
    // grow-only device workspaces: [0] host-API staging in, [1] staging out, [2] istft time frames,
    // [3] small staging (windows), [4] two-pass (N > 16384) intermediate
    // [5] dense complex rows of a non-power-of-two core (rfft / istft / strided / split), [6] its istft time frames
    static constexpr int kNumWs = 8;
    void *ws[kNumWs] = {};
    size_t ws_bytes[kNumWs] = {};
    // The device-pointer entry points are stream-ordered on the CALLER's stream, but [2] and [4] (and the
    // dependency flags of the persistent kernels) are one per context: a call on another stream first waits for
    // the event the previous user recorded (ws_acquire / ws_release), so two streams never share a scratch.
    cudaEvent_t ws_event[kNumWs] = {};
    cudaStream_t ws_stream[kNumWs] = {};
    bool ws_used[kNumWs] = {};
    size_t large_scratch_bytes = size_t(256) << 20; // two-pass intermediate per batch chunk (measured: kernel
                                                   // length matters more than L2 residency, profiles/r01n)
    size_t istft_ws_limit = size_t(1) << 30;
    size_t huge_scratch_bytes = size_t(1) << 30; // N > 2^16: both scratch buffers together (at least one transform each)
    bool use_tma = true; // TMA-staged input prefetch where alignment allows
    bool large_fused = false; // N > 16384: one persistent thread-block-cluster kernel instead of two kernels per chunk
    // N > 16384 default: one persistent cooperative kernel, per-team dependency flags, intermediate
    // pinned in L2 (fft_large.cuh LargePipe)
    bool large_pipe = true;
    bool large_auto = true; // pipelined kernel only where it measured faster (rfft, irfft), two kernels otherwise
    unsigned *pipe_flags = nullptr;
    // N = 2^13 .. 2^15: the warp-specialised split kernel (fft_split32.cuh) serves lengths 2^split_min_l .. 2^15
    // (16 = off).  Cooperative launch; when the device cannot make every CTA resident the older paths compute
    // the same bits and coop_fallbacks counts it.
    int split_min_l = 14;
    void *small_host = nullptr, *small_dev = nullptr; // pinned, device-mapped staging of the small host-pointer calls
    bool small_zero_copy = true;                       // KOFFT_SMALL_ZERO_COPY=0: copy-engine round trip instead
    unsigned wide_mask = kWideDefault; // which kinds at 2^13 / 2^14 run the wide single-CTA kernel (kofft_cuda_set_wide_mask)
    bool split_irfft = true; // irfft at 2^15 through the split kernel with the untwist in its B warps (KOFFT_SPLIT_IRFFT=0: older path)
    bool split_all_kinds = false; // default: C2C and rfft, where it measured faster; irfft / SoA rows keep the older paths
    unsigned long long coop_fallbacks = 0;
    int l2_persist_mode = 0; // 0 off, 1 requested (KOFFT_L2_PERSIST=1), 2 active
    bool istft_fused = true; // N = 512..4096: overlap-add fused behind the inverse FFT (one kernel)
    int istft_run_frames = 128;
    // host-pointer batch entry points: the batch is cut into chunks that flow through three
    // streams (H2D copy engine, SMs, D2H copy engine) so both PCIe directions and the kernels overlap
    size_t host_chunk_bytes = size_t(32) << 20; // 0 = one copy in, one launch, one copy out
    static constexpr int kPipeSlots = 4;
    cudaStream_t pipe_stream[3] = {nullptr, nullptr, nullptr}; // [0] H2D, [1] compute, [2] D2H
    cudaEvent_t pipe_event[3][kPipeSlots] = {};                // [0] H2D done, [1] kernel done, [2] D2H done
    bool pipe_ready = false;
};

namespace {

int ensure_ws(kofft_cuda_ctx *ctx, int which, size_t bytes, void **out)
{
    if (bytes == 0) bytes = 16;
    if (ctx->ws_bytes[which] < bytes) {
        if (ctx->ws[which]) {
            CU(cudaDeviceSynchronize()); // the old block may be in use on a caller's stream
            CU(cudaFree(ctx->ws[which]));
            ctx->ws[which] = nullptr;
            ctx->ws_bytes[which] = 0;
        }
        CU(cudaMalloc(&ctx->ws[which], bytes));
        ctx->ws_bytes[which] = bytes;
    }
    *out = ctx->ws[which];
    return 0;
}

// before enqueuing work that uses workspace `which` on stream s / after enqueuing it
int ws_acquire(kofft_cuda_ctx *ctx, int which, cudaStream_t s)
{
    if (ctx->ws_used[which] && ctx->ws_stream[which] != s) CU(cudaStreamWaitEvent(s, ctx->ws_event[which], 0));
    return 0;
}
int ws_release(kofft_cuda_ctx *ctx, int which, cudaStream_t s)
{
    if (!ctx->ws_event[which]) CU(cudaEventCreateWithFlags(&ctx->ws_event[which], cudaEventDisableTiming));
    CU(cudaEventRecord(ctx->ws_event[which], s));
    ctx->ws_stream[which] = s;
    ctx->ws_used[which] = true;
    return 0;
}

int get_fft_table(kofft_cuda_ctx *ctx, size_t n, const Table **out)
{
    const auto key = std::make_pair(n, ctx->accurate_tables ? 1 : 0);
    auto it = ctx->fft_tables.find(key);
    if (it == ctx->fft_tables.end()) {
        Table t;
        size_t half = n / 2;
        t.host.resize(2 * (half ? half : 1));
        if (ctx->accurate_tables)
            host_accurate_twiddles(n, 1, half, t.host.data());
        else
            host_fft_twiddles(n, t.host.data());
        CU(cudaMalloc(&t.dev, sizeof(float2) * (half ? half : 1)));
        CU(cudaMemcpyAsync(t.dev, t.host.data(), sizeof(float2) * half, cudaMemcpyHostToDevice, ctx->stream));
        if (n >= 16) {
            const int L = log2_of(n);
            for (int tl = 0; tl < 4; tl++)
                for (int c = 0; c < (1 << tl); c++) {
                    const size_t idx = static_cast<size_t>(c) << (L - 1 - tl);
                    t.pre0[(1 << tl) - 1 + c] = make_float2(t.host[2 * idx], t.host[2 * idx + 1]);
                }
        }
        // the host vector must outlive the async copy: it is owned by the map entry below
        it = ctx->fft_tables.emplace(key, std::move(t)).first;
        CU(cudaStreamSynchronize(ctx->stream));
        if (n > (size_t(1) << 20)) std::vector<float>().swap(it->second.host); // 4 n bytes: keep big tables on the device only
    }
    *out = &it->second;
    return 0;
}

int get_rfft_table(kofft_cuda_ctx *ctx, size_t m, const Table **out)
{
    auto key = std::make_pair(m, ctx->rfft_fma ? 1 : 0);
    auto it = ctx->rfft_tables.find(key);
    if (it == ctx->rfft_tables.end()) {
        Table t;
        t.host.resize(2 * (m ? m : 1));
        host_rfft_twiddles(m, t.host.data(), ctx->rfft_fma);
        CU(cudaMalloc(&t.dev, sizeof(float2) * (m ? m : 1)));
        CU(cudaMemcpyAsync(t.dev, t.host.data(), sizeof(float2) * m, cudaMemcpyHostToDevice, ctx->stream));
        it = ctx->rfft_tables.emplace(key, std::move(t)).first;
        CU(cudaStreamSynchronize(ctx->stream));
        if (m > (size_t(1) << 20)) std::vector<float>().swap(it->second.host);
    }
    *out = &it->second;
    return 0;
}

// pass-0 twiddles (group k = 0) of the single-CTA engine: v[(2^t - 1) + c] = T[c << (L-1-t)]
void fill_tw0(const Table *t, int L, Tw0 *tw0)
{
    memset(tw0, 0, sizeof *tw0);
    const int NP = L <= 8 ? 2 : (L <= 12 ? 3 : 4);
    const int R0 = L - 4 * (NP - 1);
    for (int tl = 0; tl < R0; tl++)
        for (int c = 0; c < (1 << tl); c++) {
            size_t idx = static_cast<size_t>(c) << (L - 1 - tl);
            tw0->v[(1 << tl) - 1 + c] = make_float2(t->host[2 * idx], t->host[2 * idx + 1]);
        }
}

int dispatch_impl(kofft_cuda_ctx *ctx, int kind, const IoArgs &io, size_t n, size_t rows, cudaStream_t stream, bool staged);

// Common dispatch: the complex core has length n (power of two, >= 1), `rows` transforms.
int dispatch(kofft_cuda_ctx *ctx, int kind, const IoArgs &io, size_t n, size_t rows, cudaStream_t stream,
             bool staged = false)
{
    if (rows == 0) return KOFFT_OK;
    if (n < 8192) return dispatch_impl(ctx, kind, io, n, rows, stream, staged);
    // lengths that may run through the per-context intermediate / dependency flags (workspace 4)
    int rc = ws_acquire(ctx, 4, stream);
    if (rc) return rc;
    rc = dispatch_impl(ctx, kind, io, n, rows, stream, staged);
    if (rc) return rc;
    return ws_release(ctx, 4, stream);
}

int dispatch_impl(kofft_cuda_ctx *ctx, int kind, const IoArgs &io, size_t n, size_t rows, cudaStream_t stream, bool staged)
{
    if (rows == 0) return KOFFT_OK;
    LaunchArgs a;
    a.staged = staged && ctx->use_tma;
    a.kind = kind;
    a.exact = ctx->exact;
    a.io = io;
    a.io.n = static_cast<long>(n);
    a.rows = static_cast<long>(rows);
    a.num_sms = ctx->num_sms;
    a.max_ctas = ctx->max_ctas;
    a.stream = stream;
    memset(&a.tw0, 0, sizeof a.tw0);
    (void)cudaGetLastError();
    cudaError_t e;
    if (n <= 16) {
        e = launch_small_fft(static_cast<int>(n), a);
    } else {
        const int L = log2_of(n);
        if (L > kHugeMaxLog2)
            return fail_msg(-static_cast<int>(cudaErrorNotSupported),
                            "transform lengths above 2^27 (rfft above 2^28) need the multi-GPU path (kofft_cuda_dist_*)");
        const Table *t = nullptr;
        int rc = get_fft_table(ctx, n, &t);
        if (rc) return rc;
        a.table = t->dev;
        if (L > 16) {
            // 256-point column pass + register passes through two scratch buffers (fft_huge.cu)
            if (kind == KIND_STFT || kind == KIND_ISTFT || kind == KIND_STFT_MAG)
                return fail_msg(-static_cast<int>(cudaErrorNotSupported), "STFT windows above 16384 are not supported");
            for (int i = 0; i < 15; i++) a.tw0.v[i] = t->pre0[i];
            const size_t row_bytes = n * sizeof(float2);
            size_t chunk = ctx->huge_scratch_bytes / 2 / row_bytes;
            if (chunk < 1) chunk = 1;
            if (chunk > rows) chunk = rows;
            void *scratch = nullptr;
            rc = ensure_ws(ctx, 4, 2 * chunk * row_bytes, &scratch);
            if (rc) return rc;
            HugeArgs g;
            g.scratch[0] = static_cast<float2 *>(scratch);
            g.scratch[1] = g.scratch[0] + chunk * n;
            g.chunk_rows = static_cast<long>(chunk);
            e = launch_huge_fft(L, a, g);
            if (e != cudaSuccess) return fail_cuda(e, "huge-N kernel launch");
            ctx->launches += g.launches;
            return KOFFT_OK;
        }
        // wide_mask bit (L - 13) + 2 g, g = 0 dense C2C rows, 1 rfft, 2 irfft, 3 SoA / strided rows
        const int wide_group = (kind == KIND_C2C_FWD || kind == KIND_C2C_INV) ? 0
                               : kind == KIND_RFFT                             ? 1
                               : kind == KIND_IRFFT                            ? 2
                               : (kind == KIND_GEN_FWD || kind == KIND_GEN_INV) ? 3
                                                                                : -1;
        const int wide_bit = wide_group < 0 ? -1 : (L - 13) + 2 * wide_group;
        if ((L == 13 || L == 14) && wide_bit >= 0 && ((ctx->wide_mask >> wide_bit) & 1)) {
            // one CTA per transform, 32 elements per thread, two CTAs per SM at 8192 (fft_wide.cuh)
            float2 v0[32] = {};
            for (int tl = 0; tl < 5; tl++)
                for (int c = 0; c < (1 << tl); c++) {
                    size_t idx = static_cast<size_t>(c) << (L - 1 - tl);
                    v0[(1 << tl) - 1 + c] = make_float2(t->host[2 * idx], t->host[2 * idx + 1]);
                }
            e = launch_wide_fft(L, a, v0);
            if (e != cudaSuccess) return fail_cuda(e, "wide kernel launch");
            ctx->launches += 1;
            return KOFFT_OK;
        }
        // irfft at 2^15 complex points: the split kernel's B warps untwist the rows ahead of pass A (2^13, 2^14 and the
        // SoA rows keep the older paths, where they measured faster)
        const bool split_kind = kind == KIND_C2C_FWD || kind == KIND_C2C_INV || kind == KIND_RFFT ||
                                (kind == KIND_IRFFT && L == 15 && ctx->split_irfft) ||
                                (ctx->split_all_kinds && (kind == KIND_GEN_FWD || kind == KIND_GEN_INV || kind == KIND_IRFFT));
        if (L >= ctx->split_min_l && L >= 13 && L <= 15 && split_kind) {
            SplitArgs g;
            const int ra0 = L - 10; // pass A's first register pass: stages 0 .. L-11 of the big transform
            for (int tl = 0; tl < ra0; tl++)
                for (int c = 0; c < (1 << tl); c++) {
                    size_t idx = static_cast<size_t>(c) << (L - 1 - tl);
                    g.v0[(1 << tl) - 1 + c] = make_float2(t->host[2 * idx], t->host[2 * idx + 1]);
                }
            const int nt = 1 << (L - 13); // CTAs per team
            g.max_teams = ctx->num_sms / nt;
            void *scratch = nullptr;
            g.pre_rows = kind == KIND_IRFFT && L == 15 && ctx->split_irfft;
            rc = ensure_ws(ctx, 4, size_t(kSplitSlots + (g.pre_rows ? kSplitZSlots : 0)) * g.max_teams * n * sizeof(float2), &scratch);
            if (rc) return rc;
            if (!ctx->pipe_flags) // sized for the smallest team of any persistent large-N kernel
                CU(cudaMalloc(&ctx->pipe_flags, sizeof(unsigned) * kPipeFlagStride * (kMaxPipeCtasPerSm * ctx->num_sms)));
            g.scratch = static_cast<float2 *>(scratch);
            g.flags = ctx->pipe_flags;
            if (ctx->l2_persist_mode == 1) { // tuning aid (KOFFT_L2_PERSIST=1): reserve persisting L2 for the intermediate
                int max_persist = 0;
                CU(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device));
                size_t want = size_t(kSplitSlots) * g.max_teams * n * sizeof(float2);
                if (want > size_t(max_persist)) want = size_t(max_persist);
                CU(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
                if (getenv("KOFFT_CUDA_VERBOSE")) fprintf(stderr, "[kofft_cuda] persisting L2: %zu of max %d bytes\n", want, max_persist);
                ctx->l2_persist_mode = 2;
            }
            g.persist_l2 = ctx->l2_persist_mode == 2;
            e = launch_split32_fft(L, a, g);
            if (e == cudaSuccess) {
                ctx->launches += 1;
                return KOFFT_OK;
            }
            if (e != cudaErrorCooperativeLaunchTooLarge && e != cudaErrorLaunchOutOfResources && e != cudaErrorNotSupported)
                return fail_cuda(e, "split kernel launch");
            (void)cudaGetLastError();
            ctx->coop_fallbacks++;
            g_last_error = "cooperative launch not possible (device shared or partitioned): fell back to a slower path";
        }
        if (L > 14) {
            // two-pass path: pass A's pass-0 twiddles are those of stage 0..3 of the big transform
            if (kind == KIND_STFT || kind == KIND_ISTFT)
                return fail_msg(-static_cast<int>(cudaErrorNotSupported), "STFT windows above 16384 are not supported");
            for (int tl = 0; tl < 4; tl++)
                for (int c = 0; c < (1 << tl); c++) {
                    size_t idx = static_cast<size_t>(c) << (L - 1 - tl);
                    a.tw0.v[(1 << tl) - 1 + c] = make_float2(t->host[2 * idx], t->host[2 * idx + 1]);
                }
            const size_t row_bytes = n * sizeof(float2);
            void *scratch = nullptr;
            if (ctx->large_pipe && !ctx->large_fused && (!ctx->large_auto || kind == KIND_RFFT || kind == KIND_IRFFT)) {
                const int nkb = L == 15 ? 8 : 16;
                const int max_teams = kMaxPipeCtasPerSm * ctx->num_sms / nkb;
                rc = ensure_ws(ctx, 4, size_t(kLargePipeSlots) * max_teams * row_bytes, &scratch);
                if (rc) return rc;
                if (!ctx->pipe_flags) // sized for the smallest team of any persistent large-N kernel
                    CU(cudaMalloc(&ctx->pipe_flags, sizeof(unsigned) * kPipeFlagStride * (kMaxPipeCtasPerSm * ctx->num_sms)));
                LargeArgs g;
                g.lsub = L - 8;
                g.row0 = 0;
                g.chunk_rows = static_cast<long>(rows);
                g.scratch = static_cast<float2 *>(scratch);
                g.fused = false;
                g.pipe = true;
                g.pipe_max_teams = max_teams;
                g.flags = ctx->pipe_flags;
                e = launch_large_fft(L, a, g);
                if (e == cudaSuccess) {
                    ctx->launches += g.launches;
                    return KOFFT_OK;
                }
                // the cooperative launch needs every CTA resident at once; where the device cannot grant
                // that (e.g. a partitioned GPU) the two-kernel path below computes the same bits
                if (e != cudaErrorCooperativeLaunchTooLarge && e != cudaErrorLaunchOutOfResources && e != cudaErrorNotSupported)
                    return fail_cuda(e, "large-N pipelined kernel launch");
                (void)cudaGetLastError();
                ctx->coop_fallbacks++;
                g_last_error = "cooperative launch not possible (device shared or partitioned): fell back to two kernels per chunk";
            }
            if (ctx->large_fused) {
                // one persistent launch; each cluster double-buffers one transform in scratch
                rc = ensure_ws(ctx, 4, size_t(kMaxFusedClusters) * 2 * row_bytes, &scratch);
                if (rc) return rc;
                LargeArgs g;
                g.lsub = L - 8;
                g.row0 = 0;
                g.chunk_rows = static_cast<long>(rows);
                g.scratch = static_cast<float2 *>(scratch);
                g.fused = true;
                g.max_clusters = kMaxFusedClusters;
                e = launch_large_fft(L, a, g);
                if (e == cudaSuccess) {
                    ctx->launches += 1;
                    return KOFFT_OK;
                }
                (void)cudaGetLastError(); // cluster launch not possible: fall through to two kernels
            }
            // chunk the batch so the intermediate stays L2-resident
            size_t chunk = ctx->large_scratch_bytes / row_bytes;
            if (chunk < 1) chunk = 1;
            if (chunk > rows) chunk = rows;
            rc = ensure_ws(ctx, 4, chunk * row_bytes, &scratch);
            if (rc) return rc;
            for (size_t r0 = 0; r0 < rows; r0 += chunk) {
                LargeArgs g;
                g.lsub = L - 8;
                g.row0 = static_cast<long>(r0);
                g.chunk_rows = static_cast<long>(rows - r0 < chunk ? rows - r0 : chunk);
                g.scratch = static_cast<float2 *>(scratch);
                g.fused = false;
                g.stage_rows = ctx->use_tma;
                e = launch_large_fft(L, a, g);
                if (e != cudaSuccess) return fail_cuda(e, "large-N kernel launch");
                ctx->launches += 2;
            }
            return KOFFT_OK;
        }
        // pass-0 twiddles (group k = 0): v[(2^t - 1) + c] = T[c << (L-1-t)]
        const int NP = L <= 8 ? 2 : (L <= 12 ? 3 : 4);
        const int R0 = L - 4 * (NP - 1);
        for (int tl = 0; tl < R0; tl++)
            for (int c = 0; c < (1 << tl); c++) {
                size_t idx = static_cast<size_t>(c) << (L - 1 - tl);
                a.tw0.v[(1 << tl) - 1 + c] = make_float2(t->host[2 * idx], t->host[2 * idx + 1]);
            }
        e = launch_cta_fft(L, a);
    }
    if (e != cudaSuccess) return fail_cuda(e, "kernel launch");
    ctx->launches++;
    return KOFFT_OK;
}

// length checks shared by every entry that ends in FftImpl::fft(n) (src/fft.rs:1054-1082): the std build takes
// Bluestein for non-power-of-two lengths (:1083-1132), so only the empty input is an error
int check_fft_len(size_t n)
{
    if (n == 0) return KOFFT_ERR_EMPTY_INPUT;
    return KOFFT_OK;
}
int elementwise(kofft_cuda_ctx *ctx, const ElementwiseArgs &e, cudaStream_t s)
{
    (void)cudaGetLastError();
    cudaError_t err = launch_elementwise(e, ctx->exact, ctx->num_sms, s);
    if (err != cudaSuccess) return fail_cuda(err, "element-wise kernel launch");
    ctx->launches++;
    return KOFFT_OK;
}
// rows of the non-power-of-two core per trip through workspace `which`
size_t rows_per_trip(const kofft_cuda_ctx *ctx, size_t row_bytes, size_t rows)
{
    size_t chunk = row_bytes ? ctx->istft_ws_limit / row_bytes : rows;
    if (chunk < 1) chunk = 1;
    return chunk > rows ? rows : chunk;
}

// `stream` is a cudaStream_t exactly as the caller passed it (NULL = CUDA's legacy default
// stream, which is what frameworks hand out as "the current stream" most of the time)
cudaStream_t pick_stream(kofft_cuda_ctx *, void *stream)
{
    return static_cast<cudaStream_t>(stream);
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// transforms per CTA of the single-CTA engine (Plan<L, min_cta>::TPC; 16 elements per thread)
// bytes of landing zone per point of a staged STFT group: the complex-sized stage holds real samples; the half-size
// stage build knob (KOFFT_STFT_TW1_SMEM) leaves a group only half of it
constexpr long kStftStageBytesPerPoint = IoTraits<IoStft>::kStageHalf ? 4 : 8;

long tpc_of(size_t n, int min_cta = 256)
{
    const size_t per_cta = static_cast<size_t>(min_cta) * 16;
    return n >= per_cta ? 1 : static_cast<long>(per_cta / n);
}

} // namespace

extern "C" {

int kofft_cuda_create(kofft_cuda_ctx **out, int device)
{
    if (!out) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null out pointer");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDeviceCount (no CUDA device: there is no CPU fallback)");
    if (device < 0 || device >= count) return fail_msg(-static_cast<int>(cudaErrorInvalidDevice), "invalid device index");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail_msg(-static_cast<int>(cudaErrorNoKernelImageForDevice),
                        "libkofft_cuda is built for sm_100a (B200) only");
    kofft_cuda_ctx *ctx = new kofft_cuda_ctx();
    if (const char *mb = getenv("KOFFT_LARGE_SCRATCH_MB"))
        if (atoi(mb) > 0) ctx->large_scratch_bytes = size_t(atoi(mb)) << 20;
    // N > 16384 path selection and tuning knobs (see kofft_cuda_set_large_mode)
    if (const char *m = getenv("KOFFT_LARGE_MODE")) {
        ctx->large_auto = strcmp(m, "auto") == 0;
        ctx->large_pipe = ctx->large_auto || strcmp(m, "pipe") == 0;
        ctx->large_fused = strcmp(m, "cluster") == 0;
    }

    if (const char *m = getenv("KOFFT_L2_PERSIST"))
        if (atoi(m) == 1) ctx->l2_persist_mode = 1;
    if (const char *m = getenv("KOFFT_SPLIT_MIN_L"))
        if (atoi(m) >= 13 && atoi(m) <= 16) ctx->split_min_l = atoi(m);
    if (const char *m = getenv("KOFFT_SPLIT_IRFFT")) ctx->split_irfft = atoi(m) != 0;
    if (const char *m = getenv("KOFFT_SMALL_ZERO_COPY")) ctx->small_zero_copy = atoi(m) != 0;
    if (const char *m = getenv("KOFFT_WIDE_MASK")) ctx->wide_mask = static_cast<unsigned>(strtoul(m, nullptr, 0));

    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ctx;
        return fail_cuda(e, "cudaStreamCreate");
    }
    *out = ctx;
    return KOFFT_OK;
}

void kofft_cuda_destroy(kofft_cuda_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize(); // tables and workspaces may still be in use on the callers' streams
    for (int i = 0; i < kofft_cuda_ctx::kNumWs; i++)
        if (ctx->ws_event[i]) cudaEventDestroy(ctx->ws_event[i]);
    for (auto &kv : ctx->fft_tables) cudaFree(kv.second.dev);
    for (auto &kv : ctx->rfft_tables) cudaFree(kv.second.dev);
    for (auto &kv : ctx->blue_tables) {
        cudaFree(kv.second.chirp);
        cudaFree(kv.second.bfft);
    }
    for (int i = 0; i < kofft_cuda_ctx::kNumWs; i++)
        if (ctx->ws[i]) cudaFree(ctx->ws[i]);
    if (ctx->pipe_flags) cudaFree(ctx->pipe_flags);
    if (ctx->small_host) cudaFreeHost(ctx->small_host);
    for (auto &kv : ctx->blue_tables_f64) {
        cudaFree(kv.second.chirp);
        cudaFree(kv.second.bfft);
    }
    for (auto &kv : ctx->fft_tables_f64)
        if (kv.second.dev) cudaFree(kv.second.dev);
    for (auto &kv : ctx->rfft_tables_f64)
        if (kv.second.dev) cudaFree(kv.second.dev);
    if (ctx->pipe_ready) {
        for (int i = 0; i < 3; i++) {
            cudaStreamSynchronize(ctx->pipe_stream[i]);
            cudaStreamDestroy(ctx->pipe_stream[i]);
            for (int j = 0; j < kofft_cuda_ctx::kPipeSlots; j++) cudaEventDestroy(ctx->pipe_event[i][j]);
        }
    }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *kofft_cuda_last_error(void) { return g_last_error.c_str(); }
int kofft_cuda_device(const kofft_cuda_ctx *ctx) { return ctx ? ctx->device : -1; }
void *kofft_cuda_stream(const kofft_cuda_ctx *ctx) { return ctx ? ctx->stream : nullptr; }

int kofft_cuda_synchronize(kofft_cuda_ctx *ctx)
{
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return KOFFT_OK;
}

int kofft_cuda_set_exact(kofft_cuda_ctx *ctx, int exact)
{
    ctx->exact = exact != 0;
    return KOFFT_OK;
}
int kofft_cuda_get_exact(const kofft_cuda_ctx *ctx) { return ctx->exact ? 1 : 0; }
unsigned long long kofft_cuda_launch_count(const kofft_cuda_ctx *ctx) { return ctx->launches; }
int kofft_cuda_set_max_ctas(kofft_cuda_ctx *ctx, int max_ctas)
{
    ctx->max_ctas = max_ctas < 0 ? 0 : max_ctas;
    return KOFFT_OK;
}
int kofft_cuda_set_tma_staging(kofft_cuda_ctx *ctx, int enable)
{
    ctx->use_tma = enable != 0;
    return KOFFT_OK;
}
int kofft_cuda_set_istft_fusion(kofft_cuda_ctx *ctx, int enable, int run_frames)
{
    ctx->istft_fused = enable != 0;
    if (run_frames > 0) ctx->istft_run_frames = run_frames;
    return KOFFT_OK;
}
int kofft_cuda_set_large_mode(kofft_cuda_ctx *ctx, int mode)
{
    if (mode < 0 || mode > 3) return fail_msg(KOFFT_ERR_INVALID_VALUE, "set_large_mode: mode 0..3");
    ctx->large_auto = mode == 3;
    ctx->large_pipe = mode >= 2;
    ctx->large_fused = mode == 1;
    return KOFFT_OK;
}
int kofft_cuda_set_split_min_log2n(kofft_cuda_ctx *ctx, int min_log2n)
{
    if (min_log2n < 13 || min_log2n > 16) return fail_msg(KOFFT_ERR_INVALID_VALUE, "set_split_min_log2n: 13..16 (16 = off)");
    ctx->split_min_l = min_log2n;
    return KOFFT_OK;
}
int kofft_cuda_set_split_all_kinds(kofft_cuda_ctx *ctx, int all_kinds)
{
    ctx->split_all_kinds = all_kinds != 0;
    return KOFFT_OK;
}
int kofft_cuda_set_wide_mask(kofft_cuda_ctx *ctx, unsigned mask)
{
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    ctx->wide_mask = (mask & 0x80000000u) ? kWideDefault : mask;
    return KOFFT_OK;
}
unsigned long long kofft_cuda_fallback_count(const kofft_cuda_ctx *ctx) { return ctx->coop_fallbacks; }
int kofft_cuda_set_cluster_fusion(kofft_cuda_ctx *ctx, int enable)
{
    ctx->large_fused = enable != 0;
    return KOFFT_OK;
}
int kofft_cuda_set_host_pipeline(kofft_cuda_ctx *ctx, size_t chunk_bytes)
{
    ctx->host_chunk_bytes = chunk_bytes;
    return KOFFT_OK;
}
int kofft_cuda_set_rfft_table_fma(kofft_cuda_ctx *ctx, int fma_mul)
{
    ctx->rfft_fma = fma_mul != 0;
    return KOFFT_OK;
}

// ---- planner tables ---------------------------------------------------------------------
int kofft_cuda_twiddles_host_f32(size_t n, float *out)
{
    host_fft_twiddles(n, out);
    return KOFFT_OK;
}
int kofft_cuda_rfft_twiddles_host_f32(size_t m, float *out, int fma_mul)
{
    host_rfft_twiddles(m, out, fma_mul != 0);
    return KOFFT_OK;
}
int kofft_cuda_get_twiddles(kofft_cuda_ctx *ctx, size_t n, const void **dev_ptr)
{
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    const Table *t = nullptr;
    int rc = get_fft_table(ctx, n, &t);
    if (rc) return rc;
    *dev_ptr = t->dev;
    return KOFFT_OK;
}
int kofft_cuda_get_rfft_twiddles(kofft_cuda_ctx *ctx, size_t m, const void **dev_ptr)
{
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    const Table *t = nullptr;
    int rc = get_rfft_table(ctx, m, &t);
    if (rc) return rc;
    *dev_ptr = t->dev;
    return KOFFT_OK;
}
int kofft_cuda_window_host_f32(int kind, size_t len, float beta, float *out)
{
    if (host_window(kind, len, beta, out) != 0) return fail_msg(KOFFT_ERR_INVALID_VALUE, "unknown window kind");
    return KOFFT_OK;
}

// ---- device-pointer entry points ----------------------------------------------------------
static int bluestein_c2c(kofft_cuda_ctx *ctx, const void *d_in, void *d_out, size_t n, size_t batch, int inverse,
                         cudaStream_t s);

int kofft_cuda_fft_c2c_f32(kofft_cuda_ctx *ctx, const void *d_in, void *d_out, size_t n, size_t batch,
                           int inverse, void *stream)
{
    if (n != 0 && !is_pow2(n)) { // the reference's std build takes Bluestein here (src/fft.rs:1083-1132)
        if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
        if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
        return bluestein_c2c(ctx, d_in, d_out, n, batch, inverse, pick_stream(ctx, stream));
    }
    int rc = check_fft_len(n);
    if (rc) return rc;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = pick_stream(ctx, stream);
    if (n == 1) { // fft of one point is the identity (src/fft.rs:1059-1061)
        if (d_in != d_out && batch)
            CU(cudaMemcpyAsync(d_out, d_in, batch * sizeof(float2), cudaMemcpyDeviceToDevice, s));
        return KOFFT_OK;
    }
    IoArgs io;
    io.in = d_in;
    io.out = d_out;
    io.scale = 1.0f / static_cast<float>(n); // src/fft.rs:1163
    return dispatch(ctx, inverse ? KIND_C2C_INV : KIND_C2C_FWD, io, n, batch, s, aligned16(d_in) && n >= 2);
}

// Non-power-of-two C2C: chirp multiply, two length-m power-of-two transforms, chirp multiply.
static int bluestein_c2c(kofft_cuda_ctx *ctx, const void *d_in, void *d_out, size_t n, size_t batch, int inverse,
                         cudaStream_t s)
{
    size_t m = 1;
    while (m < 2 * n - 1) m <<= 1; // (2n - 1).next_power_of_two()
    if (m > (size_t(1) << kHugeMaxLog2))
        return fail_msg(-static_cast<int>(cudaErrorNotSupported), "non-power-of-two lengths above 2^26 are not supported");
    if (batch == 0) return KOFFT_OK;
    auto it = ctx->blue_tables.find(n);
    if (it == ctx->blue_tables.end()) {
        kofft_cuda_ctx::Blue bt;
        bt.m = m;
        std::vector<float> chirp(2 * n), b(2 * m);
        host_bluestein_chirp(n, m, chirp.data(), b.data());
        CU(cudaMalloc(&bt.chirp, n * sizeof(float2)));
        CU(cudaMalloc(&bt.bfft, m * sizeof(float2)));
        CU(cudaMemcpy(bt.chirp, chirp.data(), n * sizeof(float2), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(bt.bfft, b.data(), m * sizeof(float2), cudaMemcpyHostToDevice));
        // fft(b) with the reference's arithmetic whatever mode the context is in (the planner table is mode-free)
        const bool saved_exact = ctx->exact, saved_acc = ctx->accurate_tables;
        ctx->exact = true;
        ctx->accurate_tables = false;
        int rc = kofft_cuda_fft_c2c_f32(ctx, bt.bfft, bt.bfft, m, 1, 0, ctx->stream);
        ctx->exact = saved_exact;
        ctx->accurate_tables = saved_acc;
        if (rc) return rc;
        CU(cudaStreamSynchronize(ctx->stream));
        it = ctx->blue_tables.emplace(n, bt).first;
    }
    const kofft_cuda_ctx::Blue &bt = it->second;
    // workspace rows of m complex, a bounded number of rows at a time
    size_t chunk = ctx->istft_ws_limit / (m * sizeof(float2));
    if (chunk < 1) chunk = 1;
    if (chunk > batch) chunk = batch;
    void *ws = nullptr;
    int rc = ensure_ws(ctx, 2, chunk * m * sizeof(float2), &ws);
    if (rc) return rc;
    rc = ws_acquire(ctx, 2, s);
    if (rc) return rc;
    for (size_t r0 = 0; r0 < batch; r0 += chunk) {
        BluesteinArgs b;
        b.rows = static_cast<long>(batch - r0 < chunk ? batch - r0 : chunk);
        b.x = static_cast<const float2 *>(d_in) + r0 * n;
        b.out = static_cast<float2 *>(d_out) + r0 * n;
        b.a = static_cast<float2 *>(ws);
        b.chirp = bt.chirp;
        b.bfft = bt.bfft;
        b.n = static_cast<long>(n);
        b.m = static_cast<long>(m);
        b.inverse = inverse ? 1 : 0;
        b.scale_m = 1.0f / static_cast<float>(m); // src/fft.rs:1116
        b.scale_n = 1.0f / static_cast<float>(n); // src/fft.rs:1163
        (void)cudaGetLastError();
        for (int step = 0; step < 3; step++) {
            cudaError_t e = launch_bluestein_step(step, b, ctx->exact, ctx->num_sms, s);
            if (e != cudaSuccess) return fail_cuda(e, "bluestein step launch");
            ctx->launches++;
            if (step < 2) {
                rc = kofft_cuda_fft_c2c_f32(ctx, b.a, b.a, m, static_cast<size_t>(b.rows), 0, s);
                if (rc) return rc;
            }
        }
    }
    return ws_release(ctx, 2, s);
}

int kofft_cuda_fft_strided_f32(kofft_cuda_ctx *ctx, const void *d_in, size_t in_stride, size_t in_dist,
                               void *d_out, size_t out_stride, size_t out_dist, size_t n, size_t batch,
                               int inverse, void *stream)
{
    if (in_stride == 0 || out_stride == 0) return KOFFT_ERR_INVALID_STRIDE;
    int rc = check_fft_len(n);
    if (rc) return rc;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    if (!is_pow2(n)) { // gather, fft (Bluestein), scatter: src/fft.rs:1191-1197
        cudaStream_t s = pick_stream(ctx, stream);
        const size_t chunk = rows_per_trip(ctx, n * sizeof(float2), batch);
        void *buf = nullptr;
        rc = ensure_ws(ctx, 5, chunk * n * sizeof(float2), &buf);
        if (rc) return rc;
        rc = ws_acquire(ctx, 5, s);
        if (rc) return rc;
        for (size_t r0 = 0; r0 < batch; r0 += chunk) {
            ElementwiseArgs e;
            e.op = EW_GATHER;
            e.n = static_cast<long>(n);
            e.rows = static_cast<long>(batch - r0 < chunk ? batch - r0 : chunk);
            e.re = static_cast<const float *>(d_in) + 2 * r0 * in_dist;
            e.im = e.re + 1;
            e.es = 2 * static_cast<long>(in_stride);
            e.rs = 2 * static_cast<long>(in_dist);
            e.a = static_cast<float2 *>(buf);
            rc = elementwise(ctx, e, s);
            if (rc) return rc;
            rc = bluestein_c2c(ctx, buf, buf, n, static_cast<size_t>(e.rows), inverse, s);
            if (rc) return rc;
            e.op = EW_SCATTER;
            e.out_re = static_cast<float *>(d_out) + 2 * r0 * out_dist;
            e.out_im = e.out_re + 1;
            e.es = 2 * static_cast<long>(out_stride);
            e.rs = 2 * static_cast<long>(out_dist);
            rc = elementwise(ctx, e, s);
            if (rc) return rc;
        }
        return ws_release(ctx, 5, s);
    }
    IoArgs io;
    io.in = d_in;
    io.in2 = static_cast<const float *>(d_in) + 1;
    io.out = d_out;
    io.out2 = static_cast<float *>(d_out) + 1;
    io.p0 = 2 * static_cast<long>(in_stride);
    io.p1 = 2 * static_cast<long>(in_dist);
    io.p2 = 2 * static_cast<long>(out_stride);
    io.p3 = 2 * static_cast<long>(out_dist);
    io.scale = 1.0f / static_cast<float>(n);
    return dispatch(ctx, inverse ? KIND_GEN_INV : KIND_GEN_FWD, io, n, batch, pick_stream(ctx, stream));
}

// fft2d_inplace (src/ndfft.rs:74-101): rows, then columns through the strided entry point.
int kofft_cuda_fft2d_f32(kofft_cuda_ctx *ctx, void *d_data, size_t rows, size_t cols, void *stream)
{
    if (rows == 0 || cols == 0) return KOFFT_OK; // :88-90
    int rc = kofft_cuda_fft_c2c_f32(ctx, d_data, d_data, cols, rows, 0, stream); // :95-98
    if (rc) return rc;
    // column c: elements d_data[c + r*cols], r < rows (:100-102); all columns in one batched launch
    return kofft_cuda_fft_strided_f32(ctx, d_data, cols, 1, d_data, cols, 1, rows, cols, 0, stream);
}

// fft3d_inplace (src/ndfft.rs:114-153): depth axis, row axis, column axis.
int kofft_cuda_fft3d_f32(kofft_cuda_ctx *ctx, void *d_data, size_t depth, size_t rows, size_t cols, void *stream)
{
    if (depth == 0 || rows == 0 || cols == 0) return KOFFT_OK; // :128-130
    const size_t plane = rows * cols;
    int rc = kofft_cuda_fft_strided_f32(ctx, d_data, plane, 1, d_data, plane, 1, depth, plane, 0, stream); // z :135-140
    if (rc) return rc;
    for (size_t d = 0; d < depth; d++) { // y :142-147
        float2 *p = static_cast<float2 *>(d_data) + d * plane;
        rc = kofft_cuda_fft_strided_f32(ctx, p, cols, 1, p, cols, 1, rows, cols, 0, stream);
        if (rc) return rc;
    }
    return kofft_cuda_fft_c2c_f32(ctx, d_data, d_data, cols, depth * rows, 0, stream); // x :149-154
}

static int host_roundtrip_begin(kofft_cuda_ctx *ctx, const void *src, size_t bytes, int which, void **dev);

int kofft_cuda_fft2d_host_f32(kofft_cuda_ctx *ctx, float *data, size_t data_len, size_t rows, size_t cols,
                              size_t scratch_col_len)
{
    if (rows * cols != data_len) return KOFFT_ERR_MISMATCHED_LENGTHS; // src/ndfft.rs:85-87
    if (rows == 0 || cols == 0) return KOFFT_OK;
    if (scratch_col_len != rows) return KOFFT_ERR_MISMATCHED_LENGTHS;  // :91-93
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    void *d = nullptr;
    int rc = host_roundtrip_begin(ctx, data, data_len * sizeof(float2), 0, &d);
    if (rc) return rc;
    rc = kofft_cuda_fft2d_f32(ctx, d, rows, cols, ctx->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(data, d, data_len * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KOFFT_OK;
}

int kofft_cuda_fft3d_host_f32(kofft_cuda_ctx *ctx, float *data, size_t data_len, size_t depth, size_t rows,
                              size_t cols, size_t tube_len, size_t row_len, size_t col_len)
{
    if (depth * rows * cols != data_len) return KOFFT_ERR_MISMATCHED_LENGTHS; // src/ndfft.rs:125-127
    if (depth == 0 || rows == 0 || cols == 0) return KOFFT_OK;
    if (tube_len != depth || row_len != rows || col_len != cols) return KOFFT_ERR_MISMATCHED_LENGTHS; // :131-133
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    void *d = nullptr;
    int rc = host_roundtrip_begin(ctx, data, data_len * sizeof(float2), 0, &d);
    if (rc) return rc;
    rc = kofft_cuda_fft3d_f32(ctx, d, depth, rows, cols, ctx->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(data, d, data_len * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KOFFT_OK;
}

int kofft_cuda_fft_split_f32(kofft_cuda_ctx *ctx, const float *d_in_re, const float *d_in_im, float *d_out_re,
                             float *d_out_im, size_t n, size_t batch, int inverse, void *stream)
{
    int rc = check_fft_len(n);
    if (rc) return rc;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    if (!is_pow2(n)) { // AoS copy, fft (Bluestein), copy back: src/fft.rs:797-809
        cudaStream_t s = pick_stream(ctx, stream);
        const size_t chunk = rows_per_trip(ctx, n * sizeof(float2), batch);
        void *buf = nullptr;
        rc = ensure_ws(ctx, 5, chunk * n * sizeof(float2), &buf);
        if (rc) return rc;
        rc = ws_acquire(ctx, 5, s);
        if (rc) return rc;
        for (size_t r0 = 0; r0 < batch; r0 += chunk) {
            ElementwiseArgs e;
            e.op = EW_GATHER;
            e.n = static_cast<long>(n);
            e.rows = static_cast<long>(batch - r0 < chunk ? batch - r0 : chunk);
            e.re = d_in_re + r0 * n;
            e.im = d_in_im + r0 * n;
            e.es = 1;
            e.rs = static_cast<long>(n);
            e.a = static_cast<float2 *>(buf);
            rc = elementwise(ctx, e, s);
            if (rc) return rc;
            rc = bluestein_c2c(ctx, buf, buf, n, static_cast<size_t>(e.rows), inverse, s);
            if (rc) return rc;
            e.op = EW_SCATTER;
            e.out_re = d_out_re + r0 * n;
            e.out_im = d_out_im + r0 * n;
            rc = elementwise(ctx, e, s);
            if (rc) return rc;
        }
        return ws_release(ctx, 5, s);
    }
    IoArgs io;
    io.in = d_in_re;
    io.in2 = d_in_im;
    io.out = d_out_re;
    io.out2 = d_out_im;
    io.p0 = 1;
    io.p1 = static_cast<long>(n);
    io.p2 = 1;
    io.p3 = static_cast<long>(n);
    io.scale = 1.0f / static_cast<float>(n); // src/fft.rs:1413
    return dispatch(ctx, inverse ? KIND_GEN_INV : KIND_GEN_FWD, io, n, batch, pick_stream(ctx, stream));
}

int kofft_cuda_rfft_f32(kofft_cuda_ctx *ctx, const float *d_in, void *d_out, size_t n, size_t batch, void *stream)
{
    if (n == 0) return KOFFT_ERR_EMPTY_INPUT;       // src/rfft.rs:434-436
    if (n % 2 != 0) return KOFFT_ERR_INVALID_VALUE; // :437-439
    const size_t m = n / 2;
    int rc = check_fft_len(m);
    if (rc) return rc;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    const Table *t = nullptr;
    rc = get_rfft_table(ctx, m, &t);
    if (rc) return rc;
    if (!is_pow2(m)) { // pack (a reinterpretation), fft(m) through Bluestein, twist: src/rfft.rs:444-463
        cudaStream_t s = pick_stream(ctx, stream);
        const size_t chunk = rows_per_trip(ctx, m * sizeof(float2), batch);
        void *y = nullptr;
        rc = ensure_ws(ctx, 5, chunk * m * sizeof(float2), &y);
        if (rc) return rc;
        rc = ws_acquire(ctx, 5, s);
        if (rc) return rc;
        for (size_t r0 = 0; r0 < batch; r0 += chunk) {
            const size_t nr = batch - r0 < chunk ? batch - r0 : chunk;
            rc = bluestein_c2c(ctx, d_in + r0 * n, y, m, nr, 0, s);
            if (rc) return rc;
            ElementwiseArgs e;
            e.op = EW_TWIST;
            e.n = static_cast<long>(m);
            e.rows = static_cast<long>(nr);
            e.x = static_cast<const float2 *>(y);
            e.rtw = t->dev;
            e.a = static_cast<float2 *>(d_out) + r0 * (m + 1);
            rc = elementwise(ctx, e, s);
            if (rc) return rc;
        }
        return ws_release(ctx, 5, s);
    }
    IoArgs io;
    io.in = d_in;
    io.out = d_out;
    io.aux = t->dev;
    return dispatch(ctx, KIND_RFFT, io, m, batch, pick_stream(ctx, stream), aligned16(d_in) && m >= 2);
}

int kofft_cuda_irfft_f32(kofft_cuda_ctx *ctx, const void *d_in, float *d_out, size_t n, size_t batch, void *stream)
{
    if (n == 0) return KOFFT_ERR_EMPTY_INPUT;       // src/rfft.rs:477-479
    if (n % 2 != 0) return KOFFT_ERR_INVALID_VALUE; // :480-482
    const size_t m = n / 2;
    int rc = check_fft_len(m);
    if (rc) return rc;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    const Table *t = nullptr;
    rc = get_rfft_table(ctx, m, &t);
    if (rc) return rc;
    if (!is_pow2(m)) { // untwist, ifft(m) through Bluestein, unpack (a reinterpretation): src/rfft.rs:485-507
        cudaStream_t s = pick_stream(ctx, stream);
        ElementwiseArgs e;
        e.op = EW_UNTWIST;
        e.n = static_cast<long>(m);
        e.rows = static_cast<long>(batch);
        e.x = static_cast<const float2 *>(d_in);
        e.rtw = t->dev;
        e.a = reinterpret_cast<float2 *>(d_out);
        rc = elementwise(ctx, e, s);
        if (rc) return rc;
        return bluestein_c2c(ctx, d_out, d_out, m, batch, 1, s);
    }
    IoArgs io;
    io.in = d_in;
    io.out = d_out;
    io.aux = t->dev;
    io.scale = 1.0f / static_cast<float>(m);
    return dispatch(ctx, KIND_IRFFT, io, m, batch, pick_stream(ctx, stream));
}

int kofft_cuda_stft_f32(kofft_cuda_ctx *ctx, const float *d_signal, size_t len, size_t channels,
                        const float *d_window, size_t win_len, size_t hop, void *d_frames, size_t nframes,
                        void *stream)
{
    if (hop == 0) return KOFFT_ERR_INVALID_HOP_SIZE;          // src/stft.rs:83-85
    const size_t required = (len + hop - 1) / hop;            // :86
    if (nframes < required) return KOFFT_ERR_MISMATCHED_LENGTHS; // :87-89
    if (nframes == 0 || channels == 0) return KOFFT_OK;
    int rc = check_fft_len(win_len);                          // first fft.fft(frame) :102
    if (rc) return rc;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    if (!is_pow2(win_len) || win_len > 16384) { // framing kernel, then the C2C core (Bluestein / large-N) in place on the frames
        cudaStream_t s = pick_stream(ctx, stream);
        ElementwiseArgs e;
        e.op = EW_FRAME;
        e.n = static_cast<long>(win_len);
        e.rows = static_cast<long>(channels * nframes);
        e.re = d_signal;
        e.aux_f = d_window;
        e.a = static_cast<float2 *>(d_frames);
        e.len = static_cast<long>(len);
        e.nframes = static_cast<long>(nframes);
        e.hop = static_cast<long>(hop);
        rc = elementwise(ctx, e, s);
        if (rc) return rc;
        return kofft_cuda_fft_c2c_f32(ctx, d_frames, d_frames, win_len, channels * nframes, 0, stream);
    }
    IoArgs io;
    io.in = d_signal;
    io.aux = d_window;
    io.out = d_frames;
    io.p0 = static_cast<long>(len);
    io.p1 = static_cast<long>(nframes);
    io.p2 = static_cast<long>(hop);
    // TMA staging needs 16-byte aligned, whole-float4 segments that never straddle a channel
    const long tpc = tpc_of(win_len, IoTraits<IoStft>::kMinCta);
    const bool staged = aligned16(d_signal) && len % 4 == 0 && hop % 4 == 0 && nframes % tpc == 0 &&
                        ((tpc - 1) * static_cast<long>(hop) + static_cast<long>(win_len)) * 4 <= tpc * static_cast<long>(win_len) * kStftStageBytesPerPoint;
    return dispatch(ctx, KIND_STFT, io, win_len, channels * nframes, pick_stream(ctx, stream), staged);
}

int kofft_cuda_stft_magnitudes_f32(kofft_cuda_ctx *ctx, const float *d_signal, size_t len, size_t channels,
                                   const float *d_window, size_t win_len, size_t hop, float *d_mags, size_t nframes,
                                   float *d_max, void *stream)
{
    if (hop == 0) return KOFFT_ERR_INVALID_HOP_SIZE;             // src/stft.rs:83-85 via compute_stft
    if (nframes < (len + hop - 1) / hop) return KOFFT_ERR_MISMATCHED_LENGTHS;
    if (channels == 0) return KOFFT_OK;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = pick_stream(ctx, stream);
    CU(cudaMemsetAsync(d_max, 0, channels * sizeof(float), s)); // max_mag starts at 0.0 (spectrogram.rs:64)
    if (nframes == 0) return KOFFT_OK;
    int rc = check_fft_len(win_len);
    if (rc) return rc;
    if (!is_pow2(win_len) || win_len < 32 || win_len > 16384) {
        // frames through the general stft path (Bluestein / literal kernels / large-N core), one channel at a time,
        // then |.| and the running maximum (src/visual/spectrogram.rs:52-76)
        void *fr = nullptr;
        rc = ensure_ws(ctx, 7, nframes * win_len * sizeof(float2), &fr);
        if (rc) return rc;
        rc = ws_acquire(ctx, 7, s);
        if (rc) return rc;
        for (size_t c = 0; c < channels; c++) {
            rc = kofft_cuda_stft_f32(ctx, d_signal + c * len, len, 1, d_window, win_len, hop, fr, nframes, s);
            if (rc) return rc;
            ElementwiseArgs e;
            e.op = EW_MAG;
            e.n = static_cast<long>(win_len);
            e.rows = static_cast<long>(nframes);
            e.x = static_cast<const float2 *>(fr);
            e.out_re = d_mags + c * nframes * (win_len / 2);
            e.max_bits = reinterpret_cast<int *>(d_max + c);
            rc = elementwise(ctx, e, s);
            if (rc) return rc;
        }
        return ws_release(ctx, 7, s);
    }
    const long tpc = tpc_of(win_len, IoTraits<IoStftMag>::kMinCta);
    const bool staged = aligned16(d_signal) && len % 4 == 0 && hop % 4 == 0 && nframes % tpc == 0 &&
                        ((tpc - 1) * static_cast<long>(hop) + static_cast<long>(win_len)) * 4 <= tpc * static_cast<long>(win_len) * kStftStageBytesPerPoint;
    // one launch per channel: each has its own running maximum
    for (size_t c = 0; c < channels; c++) {
        IoArgs io;
        io.in = d_signal + c * len;
        io.aux = d_window;
        io.out = d_mags + c * nframes * (win_len / 2);
        io.out2 = d_max + c;
        io.p0 = static_cast<long>(len);
        io.p1 = static_cast<long>(nframes);
        io.p2 = static_cast<long>(hop);
        rc = dispatch(ctx, KIND_STFT_MAG, io, win_len, nframes, s, staged);
        if (rc) return rc;
    }
    return KOFFT_OK;
}

static int host_roundtrip_begin(kofft_cuda_ctx *ctx, const void *src, size_t bytes, int which, void **dev);

int kofft_cuda_stft_magnitudes_host_f32(kofft_cuda_ctx *ctx, const float *samples, size_t len, size_t win_len,
                                        size_t hop, float *mags, size_t nframes, float *max_mag)
{
    if (hop == 0) return KOFFT_ERR_INVALID_HOP_SIZE;
    if (nframes < (len + hop - 1) / hop) return KOFFT_ERR_MISMATCHED_LENGTHS;
    int rc = nframes ? check_fft_len(win_len) : KOFFT_OK;
    if (rc) return rc;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    std::vector<float> w(win_len ? win_len : 1);
    host_window(KOFFT_WINDOW_HANN, win_len, 0.0f, w.data()); // stft_magnitudes always uses hann (spectrogram.rs:57)
    void *dsig = nullptr, *dwin = nullptr, *dout = nullptr;
    rc = host_roundtrip_begin(ctx, samples, len * sizeof(float), 0, &dsig);
    if (rc) return rc;
    rc = host_roundtrip_begin(ctx, w.data(), win_len * sizeof(float), 3, &dwin);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream)); // `w` is a local: its upload must finish before it goes away
    const size_t mbytes = nframes * (win_len / 2) * sizeof(float);
    rc = ensure_ws(ctx, 1, mbytes + 256, &dout);
    if (rc) return rc;
    float *d_mags = static_cast<float *>(dout);
    float *d_max = reinterpret_cast<float *>(static_cast<char *>(dout) + ((mbytes + 15) & ~size_t(15)));
    rc = kofft_cuda_stft_magnitudes_f32(ctx, static_cast<const float *>(dsig), len, 1, static_cast<const float *>(dwin),
                                        win_len, hop, d_mags, nframes, d_max, ctx->stream);
    if (rc) return rc;
    if (mbytes) CU(cudaMemcpyAsync(mags, d_mags, mbytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(max_mag, d_max, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KOFFT_OK;
}

int kofft_cuda_istft_f32(kofft_cuda_ctx *ctx, const void *d_frames, size_t nframes, size_t channels,
                         const float *d_window, size_t win_len, size_t hop, float *d_output, size_t out_len,
                         float *d_norm, int zero_uncovered, void *stream)
{
    if (hop == 0) return KOFFT_ERR_INVALID_HOP_SIZE; // src/stft.rs:125-127
    if (channels == 0) return KOFFT_OK;
    if (nframes > 0) {
        int rc = check_fft_len(win_len); // fft.ifft(frame) :141
        if (rc) return rc;
    }
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = pick_stream(ctx, stream);
    const int Lw = log2_of(win_len);
    if (ctx->istft_fused && nframes > 0 && out_len > 0 && is_pow2(win_len) && Lw >= 9 && Lw <= 12 && hop <= win_len &&
        aligned16(d_frames)) {
        // one kernel: ifft + window + ordered overlap-add + normalisation (istft_fused.cuh)
        const Table *t = nullptr;
        int rc = get_fft_table(ctx, win_len, &t);
        if (rc) return rc;
        LaunchArgs a;
        a.exact = ctx->exact;
        a.table = t->dev;
        a.num_sms = ctx->num_sms;
        a.max_ctas = ctx->max_ctas;
        a.stream = s;
        fill_tw0(t, Lw, &a.tw0);
        IstftFusedArgs f;
        f.frames = static_cast<const float2 *>(d_frames);
        f.window = d_window;
        f.output = d_output;
        f.norm = d_norm;
        f.channels = static_cast<long>(channels);
        f.nframes = static_cast<long>(nframes);
        f.hop = static_cast<long>(hop);
        f.out_len = static_cast<long>(out_len);
        f.run_frames = ctx->istft_run_frames;
        f.zero_uncovered = zero_uncovered;
        f.scale = 1.0f / static_cast<float>(win_len);
        (void)cudaGetLastError();
        cudaError_t e = launch_istft_fused(Lw, a, f);
        if (e != cudaSuccess) return fail_cuda(e, "fused istft launch");
        ctx->launches++;
        return KOFFT_OK;
    }
    if (nframes > 0 && (!is_pow2(win_len) || win_len > 16384)) {
        // ifft of every frame through the C2C core (Bluestein / large-N), real * window, ordered overlap-add
        const size_t per_ch = nframes * win_len * sizeof(float2);
        size_t chunk = per_ch ? ctx->istft_ws_limit / per_ch : channels;
        if (chunk < 1) chunk = 1;
        if (chunk > channels) chunk = channels;
        void *z = nullptr, *tm = nullptr;
        int rc = ensure_ws(ctx, 5, chunk * per_ch, &z);
        if (rc) return rc;
        rc = ensure_ws(ctx, 6, chunk * per_ch / 2, &tm);
        if (rc) return rc;
        rc = ws_acquire(ctx, 5, s);
        if (rc) return rc;
        rc = ws_acquire(ctx, 6, s);
        if (rc) return rc;
        for (size_t c0 = 0; c0 < channels; c0 += chunk) {
            const size_t nc = (channels - c0 < chunk) ? channels - c0 : chunk;
            rc = kofft_cuda_fft_c2c_f32(ctx, static_cast<const float2 *>(d_frames) + c0 * nframes * win_len, z, win_len,
                                        nc * nframes, 1, stream);
            if (rc) return rc;
            ElementwiseArgs e;
            e.op = EW_TIME;
            e.n = static_cast<long>(win_len);
            e.rows = static_cast<long>(nc * nframes);
            e.a = static_cast<float2 *>(z);
            e.aux_f = d_window;
            e.out_re = static_cast<float *>(tm);
            rc = elementwise(ctx, e, s);
            if (rc) return rc;
            OlaArgs o;
            o.time = static_cast<const float *>(tm);
            o.window = d_window;
            o.output = d_output + c0 * out_len;
            o.norm = d_norm ? d_norm + c0 * out_len : nullptr;
            o.channels = static_cast<long>(nc);
            o.nframes = static_cast<long>(nframes);
            o.win_len = static_cast<long>(win_len);
            o.hop = static_cast<long>(hop);
            o.out_len = static_cast<long>(out_len);
            o.zero_uncovered = zero_uncovered;
            cudaError_t e2 = launch_ola(o, s);
            if (e2 != cudaSuccess) return fail_cuda(e2, "ola launch");
            ctx->launches++;
        }
        rc = ws_release(ctx, 5, s);
        if (rc) return rc;
        return ws_release(ctx, 6, s);
    }
    // stage 1 writes windowed real frames into a bounded workspace, a few channels at a time
    const size_t per_channel = nframes * win_len * sizeof(float);
    size_t chunk = per_channel ? ctx->istft_ws_limit / per_channel : channels;
    if (chunk < 1) chunk = 1;
    if (chunk > channels) chunk = channels;
    void *time = nullptr;
    int rc = ensure_ws(ctx, 2, chunk * per_channel, &time);
    if (rc) return rc;
    rc = ws_acquire(ctx, 2, s);
    if (rc) return rc;
    for (size_t c0 = 0; c0 < channels; c0 += chunk) {
        const size_t nc = (channels - c0 < chunk) ? channels - c0 : chunk;
        if (nframes > 0) {
            IoArgs io;
            io.in = static_cast<const float2 *>(d_frames) + c0 * nframes * win_len;
            io.aux = d_window;
            io.out = time;
            io.scale = 1.0f / static_cast<float>(win_len);
            rc = dispatch(ctx, KIND_ISTFT, io, win_len, nc * nframes, s, aligned16(io.in) && win_len >= 2);
            if (rc) return rc;
        }
        OlaArgs o;
        o.time = static_cast<const float *>(time);
        o.window = d_window;
        o.output = d_output + c0 * out_len;
        o.norm = d_norm ? d_norm + c0 * out_len : nullptr;
        o.channels = static_cast<long>(nc);
        o.nframes = static_cast<long>(nframes);
        o.win_len = static_cast<long>(win_len);
        o.hop = static_cast<long>(hop);
        o.out_len = static_cast<long>(out_len);
        o.zero_uncovered = zero_uncovered;
        cudaError_t e = launch_ola(o, s);
        if (e != cudaSuccess) return fail_cuda(e, "ola launch");
        ctx->launches++;
    }
    return ws_release(ctx, 2, s);
}

} // extern "C"

namespace {

// Chunked, three-stream pipeline behind the host-pointer batch entry points.  `rows` independent
// rows of in_row_bytes (host, read) -> out_row_bytes (host, written); launch(d_in, d_out, nrows,
// stream) enqueues the kernels for one chunk.  Chunk i uses staging slot i % kPipeSlots; the copy
// engines and the SMs are decoupled by events, so H2D of chunk i+1, the kernels of chunk i and
// D2H of chunk i-1 run at the same time.  All kernels go to ONE stream (they may share context
// workspaces).  in_place: the kernel overwrites its input slot, which is then copied back.
template <class Launch>
int host_pipeline(kofft_cuda_ctx *ctx, const void *h_in, size_t in_row_bytes, void *h_out, size_t out_row_bytes,
                  size_t rows, bool in_place, Launch launch)
{
    constexpr int NS = kofft_cuda_ctx::kPipeSlots;
    if (!ctx->pipe_ready) {
        for (int i = 0; i < 3; i++) {
            CU(cudaStreamCreateWithFlags(&ctx->pipe_stream[i], cudaStreamNonBlocking));
            for (int j = 0; j < NS; j++) CU(cudaEventCreateWithFlags(&ctx->pipe_event[i][j], cudaEventDisableTiming));
        }
        ctx->pipe_ready = true;
    }
    const size_t big_row = in_row_bytes > out_row_bytes ? in_row_bytes : out_row_bytes;
    size_t chunk = ctx->host_chunk_bytes / big_row;
    if (chunk < 1) chunk = 1;
    if (chunk > rows) chunk = rows;
    const size_t nchunks = (rows + chunk - 1) / chunk;
    const int slots = nchunks < size_t(NS) ? static_cast<int>(nchunks) : NS;
    // slot strides rounded up to 256 bytes so every slot keeps the alignment TMA staging wants
    const size_t in_slot = (chunk * in_row_bytes + 255) & ~size_t(255);
    const size_t out_slot = (chunk * out_row_bytes + 255) & ~size_t(255);
    void *d_in = nullptr, *d_out = nullptr;
    int rc = ensure_ws(ctx, 0, in_slot * slots, &d_in);
    if (rc) return rc;
    if (!in_place) {
        rc = ensure_ws(ctx, 1, out_slot * slots, &d_out);
        if (rc) return rc;
    }
    cudaStream_t s_in = ctx->pipe_stream[0], s_k = ctx->pipe_stream[1], s_out = ctx->pipe_stream[2];
    for (size_t i = 0; i < nchunks; i++) {
        const int slot = static_cast<int>(i % slots);
        const size_t r0 = i * chunk;
        const size_t nr = rows - r0 < chunk ? rows - r0 : chunk;
        char *di = static_cast<char *>(d_in) + slot * in_slot;
        char *dout = in_place ? di : static_cast<char *>(d_out) + slot * out_slot;
        if (i >= size_t(slots)) CU(cudaStreamWaitEvent(s_in, ctx->pipe_event[2][slot], 0)); // slot drained
        CU(cudaMemcpyAsync(di, static_cast<const char *>(h_in) + r0 * in_row_bytes, nr * in_row_bytes,
                           cudaMemcpyHostToDevice, s_in));
        CU(cudaEventRecord(ctx->pipe_event[0][slot], s_in));
        CU(cudaStreamWaitEvent(s_k, ctx->pipe_event[0][slot], 0));
        rc = launch(di, dout, nr, s_k);
        if (rc) {
            cudaDeviceSynchronize();
            return rc;
        }
        CU(cudaEventRecord(ctx->pipe_event[1][slot], s_k));
        CU(cudaStreamWaitEvent(s_out, ctx->pipe_event[1][slot], 0));
        CU(cudaMemcpyAsync(static_cast<char *>(h_out) + r0 * out_row_bytes, dout, nr * out_row_bytes,
                           cudaMemcpyDeviceToHost, s_out));
        CU(cudaEventRecord(ctx->pipe_event[2][slot], s_out));
    }
    CU(cudaStreamSynchronize(s_out));
    return KOFFT_OK;
}

bool use_host_pipeline(const kofft_cuda_ctx *ctx, size_t total_bytes)
{
    return ctx->host_chunk_bytes != 0 && total_bytes > ctx->host_chunk_bytes;
}

} // namespace

extern "C" {

// ---- host-pointer drop-ins ------------------------------------------------------------------
static int host_roundtrip_begin(kofft_cuda_ctx *ctx, const void *src, size_t bytes, int which, void **dev)
{
    int rc = ensure_ws(ctx, which, bytes, dev);
    if (rc) return rc;
    if (src && bytes) CU(cudaMemcpyAsync(*dev, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

int kofft_cuda_fft_batch_host_f32(kofft_cuda_ctx *ctx, float *data, size_t n, size_t batch, int inverse)
{
    int rc = n == 0 ? KOFFT_ERR_EMPTY_INPUT : KOFFT_OK; // non-power-of-two lengths take Bluestein in fft_c2c
    if (rc) return rc;
    if (n == 1 || batch == 0) return KOFFT_OK;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    const size_t bytes = n * batch * sizeof(float2);
    if (use_host_pipeline(ctx, bytes))
        return host_pipeline(ctx, data, n * sizeof(float2), data, n * sizeof(float2), batch, true,
                             [&](void *di, void *dout, size_t nr, cudaStream_t s) {
                                 return kofft_cuda_fft_c2c_f32(ctx, di, dout, n, nr, inverse, s);
                             });
    if (bytes <= kSmallHostBytes && ctx->small_zero_copy) {
        // latency path (BASELINE configs[0]: one 1024-point transform): the kernel works in place on a pinned,
        // device-mapped staging buffer -- one launch and one synchronisation, no copy engine round trips
        if (!ctx->small_host) {
            CU(cudaHostAlloc(&ctx->small_host, kSmallHostBytes, cudaHostAllocMapped));
            CU(cudaHostGetDevicePointer(&ctx->small_dev, ctx->small_host, 0));
        }
        memcpy(ctx->small_host, data, bytes);
        rc = kofft_cuda_fft_c2c_f32(ctx, ctx->small_dev, ctx->small_dev, n, batch, inverse, ctx->stream);
        if (rc) return rc;
        CU(cudaStreamSynchronize(ctx->stream));
        memcpy(data, ctx->small_host, bytes);
        return KOFFT_OK;
    }
    void *d = nullptr;
    rc = host_roundtrip_begin(ctx, data, bytes, 0, &d);
    if (rc) return rc;
    rc = kofft_cuda_fft_c2c_f32(ctx, d, d, n, batch, inverse, ctx->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(data, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KOFFT_OK;
}

int kofft_cuda_fft_host_f32(kofft_cuda_ctx *ctx, float *data, size_t n, int inverse)
{
    return kofft_cuda_fft_batch_host_f32(ctx, data, n, 1, inverse);
}

// ---- streaming STFT / ISTFT with device-resident state (src/stft.rs:160-206, 407-520) ----------------
// The reference's StftStream / IstftStream are host structs around one fft per frame.  Here the state
// lives on the device and a push handles any number of samples / frames for many channels with the
// batched kernels; the frames / samples produced are bit-identical to the offline stft / istft (and
// therefore to the reference's streams, whose own test asserts stream == offline, tests/istft_stream.rs).
struct kofft_cuda_stft_stream {
    kofft_cuda_ctx *ctx = nullptr;
    size_t channels = 0, win_len = 0, hop = 0;
    float *d_window = nullptr;
    float *d_carry = nullptr; // [channels][win_len]: samples received but not yet behind an emitted frame
    size_t carry_len = 0;     // < win_len between calls
    size_t skip = 0;          // hop > win_len: samples still to drop before the next frame starts (carry_len == 0 then)
    float *d_work = nullptr;  // [channels][carry_len + n], grow-only
    size_t work_floats = 0;
};

namespace {
int stft_launch_frames(kofft_cuda_ctx *ctx, const float *d_signal, size_t len, size_t channels, const float *d_window,
                       size_t win_len, size_t hop, void *d_frames, size_t nframes, cudaStream_t s)
{
    // exactly kofft_cuda_stft_f32 without the "enough frames for the whole signal" check: a push emits
    // only the frames that are complete
    if (!is_pow2(win_len) || win_len > 16384) { // framing kernel, then the C2C core (Bluestein / large-N) in place
        ElementwiseArgs e;
        e.op = EW_FRAME;
        e.n = static_cast<long>(win_len);
        e.rows = static_cast<long>(channels * nframes);
        e.re = d_signal;
        e.aux_f = d_window;
        e.a = static_cast<float2 *>(d_frames);
        e.len = static_cast<long>(len);
        e.nframes = static_cast<long>(nframes);
        e.hop = static_cast<long>(hop);
        int rc = elementwise(ctx, e, s);
        if (rc) return rc;
        return kofft_cuda_fft_c2c_f32(ctx, d_frames, d_frames, win_len, channels * nframes, 0, s);
    }
    IoArgs io;
    io.in = d_signal;
    io.aux = d_window;
    io.out = d_frames;
    io.p0 = static_cast<long>(len);
    io.p1 = static_cast<long>(nframes);
    io.p2 = static_cast<long>(hop);
    const long tpc = tpc_of(win_len, IoTraits<IoStft>::kMinCta);
    const bool staged = aligned16(d_signal) && len % 4 == 0 && hop % 4 == 0 && nframes % tpc == 0 &&
                        ((tpc - 1) * static_cast<long>(hop) + static_cast<long>(win_len)) * 4 <= tpc * static_cast<long>(win_len) * kStftStageBytesPerPoint;
    return dispatch(ctx, KIND_STFT, io, win_len, channels * nframes, s, staged);
}
int stream_grow(float **buf, size_t *have, size_t want)
{
    if (*have >= want) return 0;
    if (*buf) {
        CU(cudaDeviceSynchronize());
        CU(cudaFree(*buf));
        *buf = nullptr;
        *have = 0;
    }
    CU(cudaMalloc(buf, want * sizeof(float)));
    *have = want;
    return 0;
}
} // namespace

int kofft_cuda_stft_stream_create(kofft_cuda_ctx *ctx, size_t channels, const float *window, size_t win_len, size_t hop,
                                  kofft_cuda_stft_stream **out)
{
    if (!ctx || !out || !window) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null argument");
    if (hop == 0) return KOFFT_ERR_INVALID_HOP_SIZE; // StftStream::new, src/stft.rs:178-180
    int rc = check_fft_len(win_len); // any length: the frames go through fft.fft(), Bluestein included (src/stft.rs:201)
    if (rc) return rc;
    if (channels == 0) return fail_msg(KOFFT_ERR_INVALID_VALUE, "channels == 0");
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    auto *s = new kofft_cuda_stft_stream;
    s->ctx = ctx;
    s->channels = channels;
    s->win_len = win_len;
    s->hop = hop;
    CU(cudaMalloc(&s->d_window, win_len * sizeof(float)));
    CU(cudaMemcpy(s->d_window, window, win_len * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&s->d_carry, channels * win_len * sizeof(float)));
    *out = s;
    return KOFFT_OK;
}

void kofft_cuda_stft_stream_destroy(kofft_cuda_stft_stream *s)
{
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    cudaFree(s->d_window);
    cudaFree(s->d_carry);
    cudaFree(s->d_work);
    delete s;
}

// frames a push of n more samples per channel will emit (flush: n = 0, flush != 0)
size_t kofft_cuda_stft_stream_frames(const kofft_cuda_stft_stream *s, size_t n, int flush)
{
    const size_t have = s->carry_len + (n > s->skip ? n - s->skip : 0);
    if (flush) return (have + s->hop - 1) / s->hop; // every start position < total length (src/stft.rs:193)
    return have >= s->win_len ? (have - s->win_len) / s->hop + 1 : 0;
}

int kofft_cuda_stft_stream_push(kofft_cuda_stft_stream *s, const float *d_samples, size_t n, size_t ld, void *d_frames,
                                size_t frames_cap, size_t *nframes_out, int flush, void *stream)
{
    if (!s) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null stream");
    kofft_cuda_ctx *ctx = s->ctx;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream);
    if (n && ld < n) return KOFFT_ERR_MISMATCHED_LENGTHS;
    const size_t k = kofft_cuda_stft_stream_frames(s, n, flush);
    if (nframes_out) *nframes_out = k;
    if (k > frames_cap) return KOFFT_ERR_MISMATCHED_LENGTHS;
    // hop > win_len: the samples between the end of one frame's hop and the next frame's start are dropped
    const size_t drop = s->skip < n ? s->skip : n;
    s->skip -= drop;
    d_samples += drop;
    n -= drop;
    const size_t have = s->carry_len + n;
    if (have == 0) return KOFFT_OK;
    int rc = stream_grow(&s->d_work, &s->work_floats, s->channels * have);
    if (rc) return rc;
    if (s->carry_len)
        CU(cudaMemcpy2DAsync(s->d_work, have * 4, s->d_carry, s->win_len * 4, s->carry_len * 4, s->channels,
                             cudaMemcpyDeviceToDevice, st));
    if (n)
        CU(cudaMemcpy2DAsync(s->d_work + s->carry_len, have * 4, d_samples, ld * 4, n * 4, s->channels,
                             cudaMemcpyDeviceToDevice, st));
    if (k) {
        rc = stft_launch_frames(ctx, s->d_work, have, s->channels, s->d_window, s->win_len, s->hop, d_frames, k, st);
        if (rc) return rc;
    }
    const size_t used = k * s->hop;
    if (!flush && used > have) s->skip += used - have; // the next frame starts beyond what has arrived
    const size_t rest = (flush || used >= have) ? 0 : have - used;
    if (rest)
        CU(cudaMemcpy2DAsync(s->d_carry, s->win_len * 4, s->d_work + used, have * 4, rest * 4, s->channels,
                             cudaMemcpyDeviceToDevice, st));
    s->carry_len = rest;
    return KOFFT_OK;
}

struct kofft_cuda_istft_stream {
    kofft_cuda_ctx *ctx = nullptr;
    size_t channels = 0, win_len = 0, hop = 0, halo = 0;
    float *d_window = nullptr;
    float *d_hist = nullptr;  // [channels][halo][win_len] complex: the last frames, whose tails are still open
    size_t hist = 0;          // frames held (<= halo)
    size_t pushed = 0;
    bool flushed = false;
    float *d_workf = nullptr, *d_worko = nullptr; // [channels][hist + k][win_len] complex, [channels][out_len]
    size_t workf_floats = 0, worko_floats = 0;
};

int kofft_cuda_istft_stream_create(kofft_cuda_ctx *ctx, size_t channels, const float *window, size_t win_len, size_t hop,
                                   kofft_cuda_istft_stream **out)
{
    if (!ctx || !out || !window) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null argument");
    if (hop == 0) return KOFFT_ERR_INVALID_HOP_SIZE; // IstftStream::new, src/stft.rs:434-436
    int rc = check_fft_len(win_len);
    if (rc) return rc;
    if (channels == 0) return fail_msg(KOFFT_ERR_INVALID_VALUE, "channels == 0");
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    auto *s = new kofft_cuda_istft_stream;
    s->ctx = ctx;
    s->channels = channels;
    s->win_len = win_len;
    s->hop = hop;
    s->halo = (win_len + hop - 1) / hop - 1; // earlier frames that still reach a frame's first hop samples
    CU(cudaMalloc(&s->d_window, win_len * sizeof(float)));
    CU(cudaMemcpy(s->d_window, window, win_len * sizeof(float), cudaMemcpyHostToDevice));
    if (s->halo) CU(cudaMalloc(&s->d_hist, channels * s->halo * win_len * 2 * sizeof(float)));
    *out = s;
    return KOFFT_OK;
}

void kofft_cuda_istft_stream_destroy(kofft_cuda_istft_stream *s)
{
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    cudaFree(s->d_window);
    cudaFree(s->d_hist);
    cudaFree(s->d_workf);
    cudaFree(s->d_worko);
    delete s;
}

// push k frames per channel ([channels][k][win_len] complex, dense): writes the next k * hop samples of every
// channel to d_out (row stride ld floats).  flush != 0 (k = 0): the win_len - hop samples after the last frame.
int kofft_cuda_istft_stream_push(kofft_cuda_istft_stream *s, const void *d_frames, size_t k, float *d_out, size_t ld,
                                 size_t *nsamples_out, int flush, void *stream)
{
    if (!s) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null stream");
    kofft_cuda_ctx *ctx = s->ctx;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream);
    const size_t w2 = s->win_len * 2; // floats per frame
    if (flush) k = 0;
    const size_t tail = s->win_len > s->hop ? s->win_len - s->hop : 0;
    // IstftStream::flush (src/stft.rs:497-519): nothing before the first frame, nothing the second time
    const size_t produce = flush ? ((s->pushed == 0 || s->flushed) ? 0 : tail) : k * s->hop;
    if (nsamples_out) *nsamples_out = produce;
    if (produce == 0) return KOFFT_OK;
    if (ld < produce) return KOFFT_ERR_MISMATCHED_LENGTHS;
    const size_t h = s->hist, nf = h + k;
    const size_t out_len = flush ? h * s->hop + tail : nf * s->hop;
    int rc = stream_grow(&s->d_workf, &s->workf_floats, s->channels * nf * w2);
    if (rc) return rc;
    rc = stream_grow(&s->d_worko, &s->worko_floats, s->channels * out_len);
    if (rc) return rc;
    if (h)
        CU(cudaMemcpy2DAsync(s->d_workf, nf * w2 * 4, s->d_hist, s->halo * w2 * 4, h * w2 * 4, s->channels,
                             cudaMemcpyDeviceToDevice, st));
    if (k)
        CU(cudaMemcpy2DAsync(s->d_workf + h * w2, nf * w2 * 4, d_frames, k * w2 * 4, k * w2 * 4, s->channels,
                             cudaMemcpyDeviceToDevice, st));
    CU(cudaMemsetAsync(s->d_worko, 0, s->channels * out_len * sizeof(float), st)); // istft accumulates into its output
    rc = kofft_cuda_istft_f32(ctx, s->d_workf, nf, s->channels, s->d_window, s->win_len, s->hop, s->d_worko, out_len,
                              nullptr, 0, st);
    if (rc) return rc;
    // the regions of the frames held from earlier pushes were final (and delivered) before; the new ones follow
    CU(cudaMemcpy2DAsync(d_out, ld * 4, s->d_worko + h * s->hop, out_len * 4, produce * 4, s->channels,
                         cudaMemcpyDeviceToDevice, st));
    if (flush) {
        s->flushed = true;
        return KOFFT_OK;
    }
    const size_t keep = nf < s->halo ? nf : s->halo;
    if (keep) // (src and dst never overlap: the source is the work buffer)
        CU(cudaMemcpy2DAsync(s->d_hist, s->halo * w2 * 4, s->d_workf + (nf - keep) * w2, nf * w2 * 4, keep * w2 * 4,
                             s->channels, cudaMemcpyDeviceToDevice, st));
    s->hist = keep;
    s->pushed += k;
    return KOFFT_OK;
}

// ---- f64 twin: FftImpl<f64>::fft / ifft (src/fft.rs:914-1051, 1054-1082, 1134-1174) -----------------
int kofft_cuda_twiddles_host_f64(size_t n, double *out)
{
    host_fft_twiddles_f64(n, out);
    return KOFFT_OK;
}

namespace {
// shared by the dense, strided and split f64 entry points: length checks, the device-resident
// FftPlanner<f64> table, pass-0 twiddles, launch.  a: addressing filled in by the caller.
int f64_check_len(kofft_cuda_ctx *ctx, size_t n, bool bluestein_ok = false)
{
    if (n == 0) return KOFFT_ERR_EMPTY_INPUT; // src/fft.rs:1055-1058
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    if (!is_pow2(n) && !bluestein_ok)
        return fail_msg(-static_cast<int>(cudaErrorNotSupported), "f64: non-power-of-two length in a power-of-two-only path");
    return KOFFT_OK;
}
int f64_dispatch(kofft_cuda_ctx *ctx, LaunchF64Args &a, size_t n, size_t batch, int inverse, cudaStream_t s)
{
    if (batch == 0) return KOFFT_OK;
    a.n = static_cast<long>(n);
    a.rows = static_cast<long>(batch);
    a.inverse = inverse != 0;
    // T::one() / T::from_f32(n as f32) (ifft, src/fft.rs:1167); 1.0 / n as f64 (ifft_split, :1421): equal for these n
    a.scale = 1.0 / static_cast<double>(static_cast<float>(n));
    a.num_sms = ctx->num_sms;
    a.max_ctas = ctx->max_ctas;
    a.stream = s;
    if (n >= 32) {
        auto it = ctx->fft_tables_f64.find(n);
        if (it == ctx->fft_tables_f64.end()) {
            kofft_cuda_ctx::TableD t;
            t.host.resize(n);
            host_fft_twiddles_f64(n, t.host.data());
            CU(cudaMalloc(&t.dev, (n / 2) * sizeof(double2)));
            CU(cudaMemcpy(t.dev, t.host.data(), (n / 2) * sizeof(double2), cudaMemcpyHostToDevice));
            it = ctx->fft_tables_f64.emplace(n, std::move(t)).first;
        }
        a.table = it->second.dev;
        const int L = log2_of(n);
        const int NP = L <= 8 ? 2 : (L <= 12 ? 3 : 4);
        const int R0 = L - 4 * (NP - 1);
        const std::vector<double> &h = it->second.host;
        for (int tl = 0; tl < R0; tl++)
            for (int c = 0; c < (1 << tl); c++) {
                const size_t idx = static_cast<size_t>(c) << (L - 1 - tl);
                a.tw0.v[(1 << tl) - 1 + c] = make_double2(h[2 * idx], h[2 * idx + 1]);
            }
    }
    (void)cudaGetLastError();
    cudaError_t e = launch_fft_f64(a);
    if (e != cudaSuccess) return fail_cuda(e, "f64 kernel launch");
    ctx->launches += 1;
    return KOFFT_OK;
}
} // namespace

// Non-power-of-two C2C for T = f64 (src/fft.rs:411-433, 1083-1132): as bluestein_c2c, on double2
static int bluestein_c2c_f64(kofft_cuda_ctx *ctx, const void *d_in, void *d_out, size_t n, size_t batch, int inverse,
                             cudaStream_t s)
{
    size_t m = 1;
    while (m < 2 * n - 1) m <<= 1; // (2n - 1).next_power_of_two()
    if (m > (size_t(1) << kHugeMaxLog2F64))
        return fail_msg(-static_cast<int>(cudaErrorNotSupported), "f64: non-power-of-two lengths above 2^25 are not supported");
    if (batch == 0) return KOFFT_OK;
    auto it = ctx->blue_tables_f64.find(n);
    if (it == ctx->blue_tables_f64.end()) {
        kofft_cuda_ctx::BlueD bt;
        bt.m = m;
        std::vector<double> chirp(2 * n), b(2 * m);
        host_bluestein_chirp_f64(n, m, chirp.data(), b.data());
        CU(cudaMalloc(&bt.chirp, n * sizeof(double2)));
        CU(cudaMalloc(&bt.bfft, m * sizeof(double2)));
        CU(cudaMemcpy(bt.chirp, chirp.data(), n * sizeof(double2), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(bt.bfft, b.data(), m * sizeof(double2), cudaMemcpyHostToDevice));
        int rc = kofft_cuda_fft_c2c_f64(ctx, bt.bfft, bt.bfft, m, 1, 0, ctx->stream); // fft(b), the planner's own transform
        if (rc) return rc;
        CU(cudaStreamSynchronize(ctx->stream));
        it = ctx->blue_tables_f64.emplace(n, bt).first;
    }
    const kofft_cuda_ctx::BlueD &bt = it->second;
    size_t chunk = ctx->istft_ws_limit / (m * sizeof(double2));
    if (chunk < 1) chunk = 1;
    if (chunk > batch) chunk = batch;
    void *ws = nullptr;
    int rc = ensure_ws(ctx, 2, chunk * m * sizeof(double2), &ws);
    if (rc) return rc;
    rc = ws_acquire(ctx, 2, s);
    if (rc) return rc;
    for (size_t r0 = 0; r0 < batch; r0 += chunk) {
        BluesteinArgsD b;
        b.rows = static_cast<long>(batch - r0 < chunk ? batch - r0 : chunk);
        b.x = static_cast<const double2 *>(d_in) + r0 * n;
        b.out = static_cast<double2 *>(d_out) + r0 * n;
        b.a = static_cast<double2 *>(ws);
        b.chirp = bt.chirp;
        b.bfft = bt.bfft;
        b.n = static_cast<long>(n);
        b.m = static_cast<long>(m);
        b.inverse = inverse ? 1 : 0;
        b.scale_m = 1.0 / static_cast<double>(static_cast<float>(m)); // T::one() / T::from_f32(m as f32), src/fft.rs:1116
        b.scale_n = 1.0 / static_cast<double>(static_cast<float>(n)); // src/fft.rs:1167
        (void)cudaGetLastError();
        for (int step = 0; step < 3; step++) {
            cudaError_t e = launch_bluestein_step_f64(step, b, ctx->num_sms, s);
            if (e != cudaSuccess) return fail_cuda(e, "f64 bluestein step launch");
            ctx->launches++;
            if (step < 2) {
                rc = kofft_cuda_fft_c2c_f64(ctx, b.a, b.a, m, static_cast<size_t>(b.rows), 0, s);
                if (rc) return rc;
            }
        }
    }
    return ws_release(ctx, 2, s);
}

int kofft_cuda_fft_c2c_f64(kofft_cuda_ctx *ctx, const void *d_in, void *d_out, size_t n, size_t batch, int inverse,
                           void *stream)
{
    int rc = f64_check_len(ctx, n, true);
    if (rc) return rc;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = pick_stream(ctx, stream);
    if (!is_pow2(n)) return bluestein_c2c_f64(ctx, d_in, d_out, n, batch, inverse, s);
    if (n == 1) { // identity for fft; ifft: conj, conj, * (1/1) (src/fft.rs:1139-1141 returns early)
        if (d_in != d_out && batch)
            CU(cudaMemcpyAsync(d_out, d_in, batch * sizeof(double2), cudaMemcpyDeviceToDevice, s));
        return KOFFT_OK;
    }
    if (n > 8192) {
        const int L = log2_of(n);
        if (L > kHugeMaxLog2F64)
            return fail_msg(-static_cast<int>(cudaErrorNotSupported), "f64: transform lengths above 2^26 are not supported");
        auto it = ctx->fft_tables_f64.find(n);
        if (it == ctx->fft_tables_f64.end()) {
            kofft_cuda_ctx::TableD t;
            t.host.resize(n);
            host_fft_twiddles_f64(n, t.host.data());
            CU(cudaMalloc(&t.dev, (n / 2) * sizeof(double2)));
            CU(cudaMemcpy(t.dev, t.host.data(), (n / 2) * sizeof(double2), cudaMemcpyHostToDevice));
            std::vector<double>().swap(t.host);
            it = ctx->fft_tables_f64.emplace(n, std::move(t)).first;
        }
        const size_t row_bytes = n * sizeof(double2);
        size_t chunk = ctx->huge_scratch_bytes / 2 / row_bytes;
        if (chunk < 1) chunk = 1;
        if (chunk > batch) chunk = batch;
        void *scratch = nullptr;
        rc = ensure_ws(ctx, 4, 2 * chunk * row_bytes, &scratch);
        if (rc) return rc;
        rc = ws_acquire(ctx, 4, s);
        if (rc) return rc;
        for (size_t r0 = 0; r0 < batch; r0 += chunk) {
            LaunchF64Args a;
            a.in = static_cast<const double2 *>(d_in) + r0 * n;
            a.out = static_cast<double2 *>(d_out) + r0 * n;
            a.n = static_cast<long>(n);
            a.rows = static_cast<long>(batch - r0 < chunk ? batch - r0 : chunk);
            a.inverse = inverse != 0;
            a.scale = 1.0 / static_cast<double>(static_cast<float>(n)); // src/fft.rs:1167
            a.table = it->second.dev;
            a.num_sms = ctx->num_sms;
            a.stream = s;
            int launches = 0;
            (void)cudaGetLastError();
            cudaError_t e = launch_huge_fft_f64(L, a, static_cast<double2 *>(scratch), static_cast<double2 *>(scratch) + chunk * n, &launches);
            if (e != cudaSuccess) return fail_cuda(e, "f64 huge-N kernel launch");
            ctx->launches += launches;
        }
        return ws_release(ctx, 4, s);
    }
    LaunchF64Args a;
    a.in = static_cast<const double2 *>(d_in);
    a.out = static_cast<double2 *>(d_out);
    // measured (profiles/r02b, r02c): the prefetch pays from N = 1024 up (4096: 65 -> 80 % of the HBM peak);
    // below, its extra barrier per row group costs more than the latency it hides (256: 93 -> 84 %)
    a.staged = ctx->use_tma && aligned16(d_in) && n >= 1024;
    return f64_dispatch(ctx, a, n, batch, inverse, s);
}

namespace {
// the single-CTA f64 kernel covers powers of two up to 8192; everything else goes through the dense C2C core
// (kofft_cuda_fft_c2c_f64: multi-pass kernels / Bluestein) with element-wise kernels around it, as the reference's
// own gather / fft / scatter (src/fft.rs:1191-1197, 921-933) and rfft_direct / irfft_direct (src/rfft.rs:425-508)
bool f64_needs_dense_core(size_t n) { return !is_pow2(n) || n > 8192; }

// split_inverse: ifft_split's own conj / fft / conj * (1.0 / n as f64) (src/fft.rs:1414-1425)
int f64_generic_dense(kofft_cuda_ctx *ctx, const double *in_re, const double *in_im, long in_es, long in_rs, double *out_re,
                      double *out_im, long out_es, long out_rs, size_t n, size_t batch, int inverse, bool split_inverse,
                      cudaStream_t s)
{
    if (batch == 0) return KOFFT_OK;
    const size_t chunk = rows_per_trip(ctx, n * sizeof(double2), batch);
    void *buf = nullptr;
    int rc = ensure_ws(ctx, 5, chunk * n * sizeof(double2), &buf);
    if (rc) return rc;
    rc = ws_acquire(ctx, 5, s);
    if (rc) return rc;
    for (size_t r0 = 0; r0 < batch; r0 += chunk) {
        const size_t nr = batch - r0 < chunk ? batch - r0 : chunk;
        ElementwiseArgsD e;
        e.op = EW_GATHER;
        e.n = static_cast<long>(n);
        e.rows = static_cast<long>(nr);
        e.re = in_re + r0 * in_rs;
        e.im = in_im + r0 * in_rs;
        e.es = in_es;
        e.rs = in_rs;
        e.a = static_cast<double2 *>(buf);
        e.neg_im = split_inverse ? 1 : 0;
        (void)cudaGetLastError();
        cudaError_t ce = launch_elementwise_f64(e, ctx->num_sms, s);
        if (ce != cudaSuccess) return fail_cuda(ce, "f64 gather launch");
        ctx->launches++;
        rc = kofft_cuda_fft_c2c_f64(ctx, buf, buf, n, nr, split_inverse ? 0 : inverse, s);
        if (rc) return rc;
        e.op = EW_SCATTER;
        e.out_re = out_re + r0 * out_rs;
        e.out_im = out_im + r0 * out_rs;
        e.es = out_es;
        e.rs = out_rs;
        e.scale = 1.0 / static_cast<double>(n);
        ce = launch_elementwise_f64(e, ctx->num_sms, s);
        if (ce != cudaSuccess) return fail_cuda(ce, "f64 scatter launch");
        ctx->launches++;
    }
    return ws_release(ctx, 5, s);
}
} // namespace

// strided rows of interleaved complex doubles (strides / distances in complex elements), as
// kofft_cuda_fft_strided_f32: FftImpl<f64>::fft_strided / fft_out_of_place_strided (src/fft.rs:1175-1336)
int kofft_cuda_fft_strided_f64(kofft_cuda_ctx *ctx, const void *d_in, size_t in_stride, size_t in_dist, void *d_out,
                               size_t out_stride, size_t out_dist, size_t n, size_t batch, int inverse, void *stream)
{
    if (in_stride == 0 || out_stride == 0) return KOFFT_ERR_INVALID_STRIDE;
    int rc = f64_check_len(ctx, n, true);
    if (rc) return rc;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    if (f64_needs_dense_core(n))
        return f64_generic_dense(ctx, static_cast<const double *>(d_in), static_cast<const double *>(d_in) + 1,
                                 2 * static_cast<long>(in_stride), 2 * static_cast<long>(in_dist), static_cast<double *>(d_out),
                                 static_cast<double *>(d_out) + 1, 2 * static_cast<long>(out_stride),
                                 2 * static_cast<long>(out_dist), n, batch, inverse, false, pick_stream(ctx, stream));
    LaunchF64Args a;
    a.generic = true;
    a.in_re = static_cast<const double *>(d_in);
    a.in_im = a.in_re + 1;
    a.out_re = static_cast<double *>(d_out);
    a.out_im = a.out_re + 1;
    a.in_es = 2 * static_cast<long>(in_stride);
    a.in_rs = 2 * static_cast<long>(in_dist);
    a.out_es = 2 * static_cast<long>(out_stride);
    a.out_rs = 2 * static_cast<long>(out_dist);
    return f64_dispatch(ctx, a, n, batch, inverse, pick_stream(ctx, stream));
}

// split (SoA) rows: FftImpl<f64>::fft_split / ifft_split (src/fft.rs:556-586 -> 1365-1439), batched
int kofft_cuda_fft_split_f64(kofft_cuda_ctx *ctx, const double *d_in_re, const double *d_in_im, double *d_out_re,
                             double *d_out_im, size_t n, size_t batch, int inverse, void *stream)
{
    int rc = f64_check_len(ctx, n, true);
    if (rc) return rc;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    if (f64_needs_dense_core(n))
        return f64_generic_dense(ctx, d_in_re, d_in_im, 1, static_cast<long>(n), d_out_re, d_out_im, 1, static_cast<long>(n), n,
                                 batch, inverse, inverse != 0, pick_stream(ctx, stream));
    LaunchF64Args a;
    a.generic = true;
    a.in_re = d_in_re;
    a.in_im = d_in_im;
    a.out_re = d_out_re;
    a.out_im = d_out_im;
    a.in_es = a.out_es = 1;
    a.in_rs = a.out_rs = static_cast<long>(n);
    return f64_dispatch(ctx, a, n, batch, inverse, pick_stream(ctx, stream));
}

// ---- f64 real transforms: RealFftImpl<f64> (src/rfft.rs:775-837 -> rfft_direct / irfft_direct :425-508) ----
int kofft_cuda_rfft_twiddles_host_f64(size_t m, double *out)
{
    host_rfft_twiddles_f64(m, out);
    return KOFFT_OK;
}

namespace {
int get_rfft_table_f64(kofft_cuda_ctx *ctx, size_t m, const double2 **out)
{
    auto it = ctx->rfft_tables_f64.find(m);
    if (it == ctx->rfft_tables_f64.end()) {
        kofft_cuda_ctx::TableD t;
        t.host.resize(2 * m);
        host_rfft_twiddles_f64(m, t.host.data());
        CU(cudaMalloc(&t.dev, m * sizeof(double2)));
        CU(cudaMemcpy(t.dev, t.host.data(), m * sizeof(double2), cudaMemcpyHostToDevice));
        it = ctx->rfft_tables_f64.emplace(m, std::move(t)).first;
    }
    *out = it->second.dev;
    return 0;
}
int real_f64(kofft_cuda_ctx *ctx, const void *d_in, void *d_out, size_t n, size_t batch, int which, void *stream)
{
    if (n == 0) return KOFFT_ERR_EMPTY_INPUT;        // src/rfft.rs:434-436 / 477-479
    if (n % 2 != 0) return KOFFT_ERR_INVALID_VALUE;  // :437-439 / 480-482
    const size_t m = n / 2;
    int rc = f64_check_len(ctx, m, true);
    if (rc) return rc;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    if (f64_needs_dense_core(m)) {
        if (batch == 0) return KOFFT_OK;
        cudaStream_t s = pick_stream(ctx, stream);
        const double2 *rtw = nullptr;
        rc = get_rfft_table_f64(ctx, m, &rtw);
        if (rc) return rc;
        ElementwiseArgsD e;
        e.n = static_cast<long>(m);
        e.rtw = rtw;
        (void)cudaGetLastError();
        if (which == 2) { // untwist into the output rows, ifft(m) in place, unpack = reinterpretation (src/rfft.rs:485-507)
            e.op = EW_UNTWIST;
            e.rows = static_cast<long>(batch);
            e.x = static_cast<const double2 *>(d_in);
            e.a = static_cast<double2 *>(d_out);
            cudaError_t ce = launch_elementwise_f64(e, ctx->num_sms, s);
            if (ce != cudaSuccess) return fail_cuda(ce, "f64 untwist launch");
            ctx->launches++;
            return kofft_cuda_fft_c2c_f64(ctx, d_out, d_out, m, batch, 1, s);
        }
        // pack = reinterpretation, fft(m) into a workspace, twist (src/rfft.rs:444-463)
        const size_t chunk = rows_per_trip(ctx, m * sizeof(double2), batch);
        void *y = nullptr;
        rc = ensure_ws(ctx, 5, chunk * m * sizeof(double2), &y);
        if (rc) return rc;
        rc = ws_acquire(ctx, 5, s);
        if (rc) return rc;
        for (size_t r0 = 0; r0 < batch; r0 += chunk) {
            const size_t nr = batch - r0 < chunk ? batch - r0 : chunk;
            rc = kofft_cuda_fft_c2c_f64(ctx, static_cast<const double *>(d_in) + r0 * n, y, m, nr, 0, s);
            if (rc) return rc;
            e.op = EW_TWIST;
            e.rows = static_cast<long>(nr);
            e.x = static_cast<const double2 *>(y);
            e.a = static_cast<double2 *>(d_out) + r0 * (m + 1);
            cudaError_t ce = launch_elementwise_f64(e, ctx->num_sms, s);
            if (ce != cudaSuccess) return fail_cuda(ce, "f64 twist launch");
            ctx->launches++;
        }
        return ws_release(ctx, 5, s);
    }
    LaunchF64Args a;
    a.real = which;
    a.in = static_cast<const double2 *>(d_in);
    a.out = static_cast<double2 *>(d_out);
    rc = get_rfft_table_f64(ctx, m, &a.rtw);
    if (rc) return rc;
    a.staged = which == 1 && ctx->use_tma && aligned16(d_in) && m >= 1024;
    return f64_dispatch(ctx, a, m, batch, which == 2, pick_stream(ctx, stream));
}
} // namespace

// d_in [batch][n] doubles -> d_out [batch][n/2+1] complex doubles
int kofft_cuda_rfft_f64(kofft_cuda_ctx *ctx, const double *d_in, void *d_out, size_t n, size_t batch, void *stream)
{
    return real_f64(ctx, d_in, d_out, n, batch, 1, stream);
}
// d_in [batch][n/2+1] complex doubles -> d_out [batch][n] doubles
int kofft_cuda_irfft_f64(kofft_cuda_ctx *ctx, const void *d_in, double *d_out, size_t n, size_t batch, void *stream)
{
    return real_f64(ctx, d_in, d_out, n, batch, 2, stream);
}

int kofft_cuda_rfft_batch_host_f64(kofft_cuda_ctx *ctx, const double *input, size_t n, size_t batch, double *output)
{
    if (n == 0) return KOFFT_ERR_EMPTY_INPUT;
    if (n % 2 != 0) return KOFFT_ERR_INVALID_VALUE;
    int rc = f64_check_len(ctx, n / 2, true);
    if (rc) return rc;
    if (batch == 0) return KOFFT_OK;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    const size_t ibytes = batch * n * sizeof(double), obytes = batch * (n / 2 + 1) * sizeof(double2);
    void *din = nullptr, *dout = nullptr;
    rc = host_roundtrip_begin(ctx, input, ibytes, 0, &din);
    if (rc) return rc;
    rc = ensure_ws(ctx, 1, obytes, &dout);
    if (rc) return rc;
    rc = kofft_cuda_rfft_f64(ctx, static_cast<const double *>(din), dout, n, batch, ctx->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(output, dout, obytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KOFFT_OK;
}

int kofft_cuda_irfft_batch_host_f64(kofft_cuda_ctx *ctx, const double *input, size_t n, size_t batch, double *output)
{
    if (n == 0) return KOFFT_ERR_EMPTY_INPUT;
    if (n % 2 != 0) return KOFFT_ERR_INVALID_VALUE;
    int rc = f64_check_len(ctx, n / 2, true);
    if (rc) return rc;
    if (batch == 0) return KOFFT_OK;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    const size_t ibytes = batch * (n / 2 + 1) * sizeof(double2), obytes = batch * n * sizeof(double);
    void *din = nullptr, *dout = nullptr;
    rc = host_roundtrip_begin(ctx, input, ibytes, 0, &din);
    if (rc) return rc;
    rc = ensure_ws(ctx, 1, obytes, &dout);
    if (rc) return rc;
    rc = kofft_cuda_irfft_f64(ctx, din, static_cast<double *>(dout), n, batch, ctx->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(output, dout, obytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KOFFT_OK;
}

int kofft_cuda_fft_batch_host_f64(kofft_cuda_ctx *ctx, double *data, size_t n, size_t batch, int inverse)
{
    if (n == 0) return KOFFT_ERR_EMPTY_INPUT;
    if (n == 1 || batch == 0) return KOFFT_OK;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    const size_t bytes = n * batch * sizeof(double2);
    void *d = nullptr;
    int rc = host_roundtrip_begin(ctx, data, bytes, 0, &d);
    if (rc) return rc;
    rc = kofft_cuda_fft_c2c_f64(ctx, d, d, n, batch, inverse, ctx->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(data, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KOFFT_OK;
}

int kofft_cuda_fft_host_f64(kofft_cuda_ctx *ctx, double *data, size_t n, int inverse)
{
    return kofft_cuda_fft_batch_host_f64(ctx, data, n, 1, inverse);
}

int kofft_cuda_fft_split_host_f64(kofft_cuda_ctx *ctx, double *re, size_t re_len, double *im, size_t im_len, int inverse)
{
    if (re_len != im_len) return KOFFT_ERR_MISMATCHED_LENGTHS; // src/fft.rs:1366-1368
    const size_t n = re_len;
    int rc = f64_check_len(ctx, n, true);
    if (rc) return rc;
    if (n == 1) return KOFFT_OK; // identity; ifft_split: (negate, negate, * 1/1)
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    void *d = nullptr;
    rc = ensure_ws(ctx, 0, 2 * n * sizeof(double), &d);
    if (rc) return rc;
    double *dre = static_cast<double *>(d), *dim = dre + n;
    CU(cudaMemcpyAsync(dre, re, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dim, im, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    rc = kofft_cuda_fft_split_f64(ctx, dre, dim, dre, dim, n, 1, inverse, ctx->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(re, dre, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(im, dim, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KOFFT_OK;
}

int kofft_cuda_fft_strided_host_f64(kofft_cuda_ctx *ctx, double *input, size_t input_len, size_t stride, size_t n,
                                    int inverse)
{
    if (stride == 0) return KOFFT_ERR_INVALID_STRIDE;                           // src/fft.rs:1181-1183
    if (n == 0) return KOFFT_OK;                                                // :1185-1187
    if (input_len < (n - 1) * stride + 1) return KOFFT_ERR_MISMATCHED_LENGTHS; // :1188-1190
    int rc = f64_check_len(ctx, n, true);
    if (rc) return rc;
    if (n == 1) return KOFFT_OK;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    const size_t span = (n - 1) * stride + 1;
    void *d = nullptr;
    rc = host_roundtrip_begin(ctx, input, span * sizeof(double2), 0, &d);
    if (rc) return rc;
    rc = kofft_cuda_fft_strided_f64(ctx, d, stride, span, d, stride, span, n, 1, inverse, ctx->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(input, d, span * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KOFFT_OK;
}

int kofft_cuda_fft_out_of_place_strided_host_f64(kofft_cuda_ctx *ctx, const double *input, size_t input_len,
                                                 size_t in_stride, double *output, size_t output_len, size_t out_stride,
                                                 int inverse)
{
    if (in_stride == 0 || out_stride == 0) return KOFFT_ERR_INVALID_STRIDE;   // src/fft.rs:1267-1269
    if (input_len % in_stride != 0 || output_len % out_stride != 0) return KOFFT_ERR_INVALID_STRIDE; // :1270-1272
    const size_t n = input_len / in_stride;
    if (output_len / out_stride != n) return KOFFT_ERR_MISMATCHED_LENGTHS;    // :1274-1276
    int rc = f64_check_len(ctx, n, true);
    if (rc) return rc;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    void *din = nullptr, *dout = nullptr;
    rc = host_roundtrip_begin(ctx, input, input_len * sizeof(double2), 0, &din);
    if (rc) return rc;
    rc = host_roundtrip_begin(ctx, output, output_len * sizeof(double2), 1, &dout); // untouched elements survive
    if (rc) return rc;
    if (n == 1) { // a one-point transform copies the element (ifft: conj, conj, * 1)
        CU(cudaMemcpyAsync(dout, din, sizeof(double2), cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        rc = kofft_cuda_fft_strided_f64(ctx, din, in_stride, input_len, dout, out_stride, output_len, n, 1, inverse,
                                        ctx->stream);
        if (rc) return rc;
    }
    CU(cudaMemcpyAsync(output, dout, output_len * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KOFFT_OK;
}

int kofft_cuda_fft_split_host_f32(kofft_cuda_ctx *ctx, float *re, size_t re_len, float *im, size_t im_len,
                                  int inverse)
{
    if (re_len != im_len) return KOFFT_ERR_MISMATCHED_LENGTHS; // src/fft.rs:1366-1368
    const size_t n = re_len;
    int rc = check_fft_len(n);
    if (rc) return rc;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    void *d = nullptr;
    rc = ensure_ws(ctx, 0, 2 * n * sizeof(float), &d);
    if (rc) return rc;
    float *dre = static_cast<float *>(d), *dim = dre + n;
    CU(cudaMemcpyAsync(dre, re, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dim, im, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if (n > 1) { // n == 1 is the identity for fft_split and (negate, negate, *1/1) for ifft_split
        rc = kofft_cuda_fft_split_f32(ctx, dre, dim, dre, dim, n, 1, inverse, ctx->stream);
        if (rc) return rc;
    }
    CU(cudaMemcpyAsync(re, dre, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(im, dim, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KOFFT_OK;
}

int kofft_cuda_fft_strided_host_f32(kofft_cuda_ctx *ctx, float *input, size_t input_len, size_t stride, size_t n,
                                    int inverse)
{
    if (stride == 0) return KOFFT_ERR_INVALID_STRIDE;                           // src/fft.rs:1181-1183
    if (n == 0) return KOFFT_OK;                                                // :1185-1187
    if (input_len < (n - 1) * stride + 1) return KOFFT_ERR_MISMATCHED_LENGTHS; // :1188-1190
    int rc = check_fft_len(n);
    if (rc) return rc;
    if (n == 1) return KOFFT_OK;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    const size_t span = (n - 1) * stride + 1;
    void *d = nullptr;
    rc = host_roundtrip_begin(ctx, input, span * sizeof(float2), 0, &d);
    if (rc) return rc;
    rc = kofft_cuda_fft_strided_f32(ctx, d, stride, span, d, stride, span, n, 1, inverse, ctx->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(input, d, span * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KOFFT_OK;
}

int kofft_cuda_fft_out_of_place_strided_host_f32(kofft_cuda_ctx *ctx, const float *input, size_t input_len,
                                                 size_t in_stride, float *output, size_t output_len,
                                                 size_t out_stride, int inverse)
{
    if (in_stride == 0 || out_stride == 0) return KOFFT_ERR_INVALID_STRIDE;   // src/fft.rs:1267-1269
    if (input_len % in_stride != 0 || output_len % out_stride != 0) return KOFFT_ERR_INVALID_STRIDE; // :1270-1272
    const size_t n = input_len / in_stride;
    if (output_len / out_stride != n) return KOFFT_ERR_MISMATCHED_LENGTHS;    // :1274-1276
    int rc = check_fft_len(n);
    if (rc) return rc;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    void *din = nullptr, *dout = nullptr;
    rc = host_roundtrip_begin(ctx, input, input_len * sizeof(float2), 0, &din);
    if (rc) return rc;
    // the scatter only touches every out_stride-th element: stage the caller's buffer so the
    // untouched elements survive the copy back
    rc = host_roundtrip_begin(ctx, output, output_len * sizeof(float2), 1, &dout);
    if (rc) return rc;
    rc = kofft_cuda_fft_strided_f32(ctx, din, in_stride, input_len, dout, out_stride, output_len, n, 1, inverse,
                                    ctx->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(output, dout, output_len * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KOFFT_OK;
}

int kofft_cuda_rfft_batch_host_f32(kofft_cuda_ctx *ctx, const float *input, size_t n, size_t batch, float *output)
{
    if (n == 0) return KOFFT_ERR_EMPTY_INPUT;
    if (n % 2 != 0) return KOFFT_ERR_INVALID_VALUE;
    const size_t m = n / 2;
    int rc = check_fft_len(m);
    if (rc) return rc;
    if (batch == 0) return KOFFT_OK;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    if (use_host_pipeline(ctx, (m + 1) * batch * sizeof(float2)))
        return host_pipeline(ctx, input, n * sizeof(float), output, (m + 1) * sizeof(float2), batch, false,
                             [&](void *di, void *dout, size_t nr, cudaStream_t s) {
                                 return kofft_cuda_rfft_f32(ctx, static_cast<const float *>(di), dout, n, nr, s);
                             });
    void *din = nullptr, *dout = nullptr;
    rc = host_roundtrip_begin(ctx, input, n * batch * sizeof(float), 0, &din);
    if (rc) return rc;
    rc = ensure_ws(ctx, 1, (m + 1) * batch * sizeof(float2), &dout);
    if (rc) return rc;
    rc = kofft_cuda_rfft_f32(ctx, static_cast<const float *>(din), dout, n, batch, ctx->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(output, dout, (m + 1) * batch * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KOFFT_OK;
}

int kofft_cuda_rfft_host_f32(kofft_cuda_ctx *ctx, const float *input, size_t n, float *output, size_t output_len,
                             size_t scratch_len)
{
    if (n == 0) return KOFFT_ERR_EMPTY_INPUT;
    if (n % 2 != 0) return KOFFT_ERR_INVALID_VALUE;
    if (output_len != n / 2 + 1 || scratch_len < n / 2) return KOFFT_ERR_MISMATCHED_LENGTHS; // src/rfft.rs:441-443
    return kofft_cuda_rfft_batch_host_f32(ctx, input, n, 1, output);
}

int kofft_cuda_irfft_batch_host_f32(kofft_cuda_ctx *ctx, const float *input, size_t n, size_t batch, float *output)
{
    if (n == 0) return KOFFT_ERR_EMPTY_INPUT;
    if (n % 2 != 0) return KOFFT_ERR_INVALID_VALUE;
    const size_t m = n / 2;
    int rc = check_fft_len(m);
    if (rc) return rc;
    if (batch == 0) return KOFFT_OK;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    if (use_host_pipeline(ctx, (m + 1) * batch * sizeof(float2)))
        return host_pipeline(ctx, input, (m + 1) * sizeof(float2), output, n * sizeof(float), batch, false,
                             [&](void *di, void *dout, size_t nr, cudaStream_t s) {
                                 return kofft_cuda_irfft_f32(ctx, di, static_cast<float *>(dout), n, nr, s);
                             });
    void *din = nullptr, *dout = nullptr;
    rc = host_roundtrip_begin(ctx, input, (m + 1) * batch * sizeof(float2), 0, &din);
    if (rc) return rc;
    rc = ensure_ws(ctx, 1, n * batch * sizeof(float), &dout);
    if (rc) return rc;
    rc = kofft_cuda_irfft_f32(ctx, din, static_cast<float *>(dout), n, batch, ctx->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(output, dout, n * batch * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KOFFT_OK;
}

int kofft_cuda_irfft_host_f32(kofft_cuda_ctx *ctx, const float *input, size_t input_len, float *output, size_t n,
                              size_t scratch_len)
{
    if (n == 0) return KOFFT_ERR_EMPTY_INPUT;
    if (n % 2 != 0) return KOFFT_ERR_INVALID_VALUE;
    if (input_len != n / 2 + 1 || scratch_len < n / 2) return KOFFT_ERR_MISMATCHED_LENGTHS; // src/rfft.rs:484-486
    return kofft_cuda_irfft_batch_host_f32(ctx, input, n, 1, output);
}

int kofft_cuda_stft_host_f32(kofft_cuda_ctx *ctx, const float *signal, size_t len, size_t channels,
                             const float *window, size_t win_len, size_t hop, float *frames, size_t nframes)
{
    if (hop == 0) return KOFFT_ERR_INVALID_HOP_SIZE;
    if (nframes < (len + hop - 1) / hop) return KOFFT_ERR_MISMATCHED_LENGTHS;
    if (nframes == 0 || channels == 0) return KOFFT_OK;
    int rc = check_fft_len(win_len);
    if (rc) return rc;
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    void *dsig = nullptr, *dfr = nullptr, *dwin = nullptr;
    rc = host_roundtrip_begin(ctx, signal, len * channels * sizeof(float), 0, &dsig);
    if (rc) return rc;
    rc = host_roundtrip_begin(ctx, window, win_len * sizeof(float), 3, &dwin);
    if (rc) return rc;
    const size_t fbytes = channels * nframes * win_len * sizeof(float2);
    rc = ensure_ws(ctx, 1, fbytes, &dfr);
    if (rc) return rc;
    rc = kofft_cuda_stft_f32(ctx, static_cast<const float *>(dsig), len, channels, static_cast<const float *>(dwin),
                             win_len, hop, dfr, nframes, ctx->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(frames, dfr, fbytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KOFFT_OK;
}

int kofft_cuda_istft_host_f32(kofft_cuda_ctx *ctx, const float *frames, size_t nframes, size_t channels,
                              const float *window, size_t win_len, size_t hop, float *output, size_t out_len,
                              float *scratch, size_t scratch_len, int zero_uncovered)
{
    if (hop == 0) return KOFFT_ERR_INVALID_HOP_SIZE;
    if (!zero_uncovered && scratch_len != out_len * channels)
        return KOFFT_ERR_MISMATCHED_LENGTHS; // src/stft.rs:128-130 (per channel)
    if (channels == 0) return KOFFT_OK;
    if (nframes > 0) {
        int rc = check_fft_len(win_len);
        if (rc) return rc;
    }
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    void *dfr = nullptr, *dwin = nullptr, *dout = nullptr;
    int rc = host_roundtrip_begin(ctx, frames, channels * nframes * win_len * sizeof(float2), 0, &dfr);
    if (rc) return rc;
    rc = host_roundtrip_begin(ctx, window, win_len * sizeof(float), 3, &dwin);
    if (rc) return rc;
    // output (accumulated into) followed by norm in one staging buffer
    const size_t obytes = channels * out_len * sizeof(float);
    rc = ensure_ws(ctx, 1, 2 * obytes, &dout);
    if (rc) return rc;
    float *d_out = static_cast<float *>(dout);
    float *d_norm = d_out + channels * out_len;
    if (obytes) CU(cudaMemcpyAsync(d_out, output, obytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = kofft_cuda_istft_f32(ctx, dfr, nframes, channels, static_cast<const float *>(dwin), win_len, hop, d_out,
                              out_len, d_norm, zero_uncovered, ctx->stream);
    if (rc) return rc;
    if (obytes) CU(cudaMemcpyAsync(output, d_out, obytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (scratch && obytes) CU(cudaMemcpyAsync(scratch, d_norm, obytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KOFFT_OK;
}

} // extern "C"

// ---- one transform sharded over several GPUs (BASELINE configs[4]) -----------------
// Four-step split N = N1 * N2 (n = N2 n1 + n2, k = k1 + N1 k2):
//   X[k1 + N1 k2] = sum_n2 W_N2^{n2 k2} * W_N^{n2 k1} * sum_n1 W_N1^{n1 k1} x[N2 n1 + n2]
// Rank g owns rows n1 in [g R1, (g+1) R1) of x viewed as [N1][N2] (a contiguous slice of the
// input), R1 = N1 / world, C2 = N2 / world.
//   phase 0  transpose-scatter x_g [R1][N2] -> A_h [C2][N1] on every rank h      (P2P stores)
//   phase 1  N1-point FFTs of the C2 rows of A (in place), then transpose-scatter with the
//            inter-step twiddle W_N^{n2 k1} -> B_h [R1][N2] on every rank h       (P2P stores)
//   phase 2  N2-point FFTs of the R1 rows of B: B[r][k2] = X[(g R1 + r) + N1 k2] (transposed
//            order, written to d_out) -- or, natural order: in place, then transpose-scatter
//            -> A_h [C2][N1] (= X[h N/world ...], the contiguous slice of the spectrum)
//   phase 3  natural order only: copy A -> d_out
// Ranks must not start phase p+1 before every rank has finished phase p (the caller puts a
// barrier between phases: stream synchronize + process barrier, or kofft_cuda_dist_run_local).
// Twiddles are correctly rounded (f64-evaluated) -- the reference's recurrence table degenerates
// at this size, so this path is validated against f64, not against kofft.
struct kofft_cuda_dist {
    kofft_cuda_ctx *ctx = nullptr;
    int rank = 0, world = 1, log2n = 0, l1 = 0, l2 = 0, llo = 0;
    size_t n1 = 0, n2 = 0, r1 = 0, c2 = 0;
    float2 *bufA = nullptr, *bufB = nullptr;
    float2 *peerA[kMaxDistWorld] = {}, *peerB[kMaxDistWorld] = {};
    bool ipc_open[kMaxDistWorld] = {};
    bool connected = false;
    float2 *tlo = nullptr, *thi = nullptr;
    // the local transforms of a phase are cut into pieces; piece i is scattered to the peers on a
    // second stream while piece i+1 is transformed, so NVLink traffic overlaps the butterflies
    // 8 pieces / 64 SMs: 8-GPU sweeps profiles/r04e (2^30: 5.03 ms against 5.57 ms with 4 / 32)
    int pieces = 8;
    int reserve_sms = 64; // SMs the piece transforms leave free so the concurrent scatter kernel gets on the machine
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fft[16] = {}, ev_done = nullptr;
};

namespace {

// rows [r_begin, r_begin + r_count) of the local matrix src [rows][world * cb]
int dist_scatter(kofft_cuda_dist *d, const float2 *src, float2 *const *peers, size_t rows, size_t cb, int twiddle,
                 cudaStream_t s, size_t r_begin = 0, size_t r_count = 0)
{
    if (r_count == 0) r_count = rows - r_begin;
    ScatterArgs a;
    a.src = src + r_begin * cb * d->world;
    for (int g = 0; g < d->world; g++) a.dst[g] = peers[g];
    a.rows = static_cast<long>(r_count);
    a.cb = static_cast<long>(cb);
    a.world = d->world;
    a.rank = d->rank;
    a.dst_pitch = static_cast<long>(rows) * d->world;
    a.dst_off = static_cast<long>(rows) * d->rank + static_cast<long>(r_begin);
    a.twiddle = twiddle;
    a.row0 = static_cast<long>(rows) * d->rank + static_cast<long>(r_begin);
    a.log2n = d->log2n;
    a.llo = d->llo;
    a.tlo = d->tlo;
    a.thi = d->thi;
    (void)cudaGetLastError();
    cudaError_t e = launch_transpose_scatter(a, d->ctx->num_sms, s);
    if (e != cudaSuccess) return fail_cuda(e, "transpose_scatter launch");
    d->ctx->launches++;
    return KOFFT_OK;
}

// batched C2C with correctly rounded tables, restoring the context's table mode afterwards
int dist_local_fft(kofft_cuda_dist *d, const void *in, void *out, size_t n, size_t batch, int inverse, cudaStream_t s)
{
    const bool saved = d->ctx->accurate_tables;
    d->ctx->accurate_tables = true;
    int rc = kofft_cuda_fft_c2c_f32(d->ctx, in, out, n, batch, inverse, s);
    d->ctx->accurate_tables = saved;
    return rc;
}

// `batch` transforms of length n in place on buf, each piece followed by its transpose-scatter on the
// side stream; on return `s` waits for the last scatter, so the phase completes in stream order.
int dist_fft_then_scatter(kofft_cuda_dist *d, float2 *buf, size_t n, size_t batch, int inverse, float2 *const *peers,
                          int twiddle, cudaStream_t s)
{
    int pieces = d->pieces;
    while (pieces > 1 && (batch % pieces != 0 || (batch / pieces) % 32 != 0)) pieces >>= 1;
    if (pieces <= 1 || !d->side) {
        int rc = dist_local_fft(d, buf, buf, n, batch, inverse, s);
        if (rc) return rc;
        return dist_scatter(d, buf, peers, batch, n / d->world, twiddle, s);
    }
    const size_t per = batch / pieces;
    // the transform kernels are persistent and would fill every SM: shrink their grid a little so
    // the scatter of the previous piece (side stream) runs at the same time
    const int all_sms = d->ctx->num_sms;
    const int fft_sms = all_sms - d->reserve_sms > all_sms / 2 ? all_sms - d->reserve_sms : all_sms;
    for (int i = 0; i < pieces; i++) {
        float2 *p = buf + i * per * n;
        d->ctx->num_sms = i == 0 ? all_sms : fft_sms; // nothing to overlap with during the first piece
        int rc = dist_local_fft(d, p, p, n, per, inverse, s);
        d->ctx->num_sms = all_sms;
        if (rc) return rc;
        CU(cudaEventRecord(d->ev_fft[i], s));
        CU(cudaStreamWaitEvent(d->side, d->ev_fft[i], 0));
        rc = dist_scatter(d, buf, peers, batch, n / d->world, twiddle, d->side, i * per, per);
        if (rc) return rc;
    }
    CU(cudaEventRecord(d->ev_done, d->side));
    CU(cudaStreamWaitEvent(s, d->ev_done, 0));
    return KOFFT_OK;
}

} // namespace

extern "C" {

int kofft_cuda_dist_create(kofft_cuda_ctx *ctx, int rank, int world, int log2n, kofft_cuda_dist **out)
{
    if (!out) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null out pointer");
    *out = nullptr;
    if (world < 1 || world > kMaxDistWorld || (world & (world - 1)) != 0 || rank < 0 || rank >= world)
        return fail_msg(KOFFT_ERR_INVALID_VALUE, "world must be a power of two <= 16 and 0 <= rank < world");
    const int l1 = log2n / 2, l2 = log2n - l1;
    const int lw = log2_of(static_cast<size_t>(world));
    if (l2 > 16 || l1 - lw < 5)
        return fail_msg(KOFFT_ERR_INVALID_VALUE, "log2n must satisfy 10 + 2 log2(world) <= log2n <= 32");
    if (!ctx) return fail_msg(KOFFT_ERR_INVALID_VALUE, "null context");
    CU(cudaSetDevice(ctx->device));
    kofft_cuda_dist *d = new kofft_cuda_dist();
    d->ctx = ctx;
    d->rank = rank;
    d->world = world;
    d->log2n = log2n;
    d->l1 = l1;
    d->l2 = l2;
    d->n1 = size_t(1) << l1;
    d->n2 = size_t(1) << l2;
    d->r1 = d->n1 / world;
    d->c2 = d->n2 / world;
    d->llo = (log2n + 1) / 2;
    const size_t shard = (size_t(1) << log2n) / world * sizeof(float2);
    const size_t nlo = size_t(1) << d->llo, nhi = size_t(1) << (log2n - d->llo);
    std::vector<float> h(2 * (nlo > nhi ? nlo : nhi));
    cudaError_t e = cudaMalloc(&d->bufA, shard);
    if (e == cudaSuccess) e = cudaMalloc(&d->bufB, shard);
    if (e == cudaSuccess) e = cudaMalloc(&d->tlo, nlo * sizeof(float2));
    if (e == cudaSuccess) e = cudaMalloc(&d->thi, nhi * sizeof(float2));
    if (e == cudaSuccess) {
        host_accurate_twiddles(size_t(1) << log2n, 1, nlo, h.data());
        e = cudaMemcpy(d->tlo, h.data(), nlo * sizeof(float2), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) {
        host_accurate_twiddles(size_t(1) << log2n, nlo, nhi, h.data());
        e = cudaMemcpy(d->thi, h.data(), nhi * sizeof(float2), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        cudaFree(d->bufA);
        cudaFree(d->bufB);
        cudaFree(d->tlo);
        cudaFree(d->thi);
        delete d;
        return fail_cuda(e, "kofft_cuda_dist_create");
    }
    d->peerA[rank] = d->bufA;
    d->peerB[rank] = d->bufB;
    if (const char *e = getenv("KOFFT_DIST_PIECES")) // tuning aids
        if (atoi(e) >= 1 && atoi(e) <= 16) d->pieces = atoi(e);
    if (const char *e = getenv("KOFFT_DIST_RESERVE_SMS"))
        if (atoi(e) >= 0 && atoi(e) < ctx->num_sms) d->reserve_sms = atoi(e);
    if (cudaStreamCreateWithFlags(&d->side, cudaStreamNonBlocking) == cudaSuccess) {
        for (int i = 0; i < 16; i++) cudaEventCreateWithFlags(&d->ev_fft[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&d->ev_done, cudaEventDisableTiming);
    } else {
        d->side = nullptr;
        (void)cudaGetLastError();
    }
    d->connected = world == 1;
    *out = d;
    return KOFFT_OK;
}

void kofft_cuda_dist_destroy(kofft_cuda_dist *d)
{
    if (!d) return;
    cudaSetDevice(d->ctx->device);
    cudaDeviceSynchronize();
    for (int g = 0; g < d->world; g++)
        if (d->ipc_open[g]) {
            cudaIpcCloseMemHandle(d->peerA[g]);
            cudaIpcCloseMemHandle(d->peerB[g]);
        }
    if (d->side) {
        cudaStreamDestroy(d->side);
        for (int i = 0; i < 16; i++) cudaEventDestroy(d->ev_fft[i]);
        cudaEventDestroy(d->ev_done);
    }
    cudaFree(d->bufA);
    cudaFree(d->bufB);
    cudaFree(d->tlo);
    cudaFree(d->thi);
    delete d;
}

int kofft_cuda_dist_set_pieces(kofft_cuda_dist *d, int pieces)
{
    d->pieces = pieces < 1 ? 1 : (pieces > 16 ? 16 : pieces);
    return KOFFT_OK;
}
size_t kofft_cuda_dist_shard_len(const kofft_cuda_dist *d) { return (size_t(1) << d->log2n) / d->world; }
void *kofft_cuda_dist_buffer(const kofft_cuda_dist *d, int which) { return which == 0 ? d->bufA : d->bufB; }

int kofft_cuda_dist_ipc_handles(kofft_cuda_dist *d, void *out128)
{
    CU(cudaSetDevice(d->ctx->device));
    cudaIpcMemHandle_t h[2];
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "two 64-byte handles per rank");
    CU(cudaIpcGetMemHandle(&h[0], d->bufA));
    CU(cudaIpcGetMemHandle(&h[1], d->bufB));
    memcpy(out128, h, sizeof h);
    return KOFFT_OK;
}

int kofft_cuda_dist_connect_ipc(kofft_cuda_dist *d, const void *all_handles)
{
    CU(cudaSetDevice(d->ctx->device));
    const cudaIpcMemHandle_t *h = static_cast<const cudaIpcMemHandle_t *>(all_handles);
    for (int g = 0; g < d->world; g++) {
        if (g == d->rank || d->ipc_open[g]) continue;
        void *pa = nullptr, *pb = nullptr;
        CU(cudaIpcOpenMemHandle(&pa, h[2 * g], cudaIpcMemLazyEnablePeerAccess));
        CU(cudaIpcOpenMemHandle(&pb, h[2 * g + 1], cudaIpcMemLazyEnablePeerAccess));
        d->peerA[g] = static_cast<float2 *>(pa);
        d->peerB[g] = static_cast<float2 *>(pb);
        d->ipc_open[g] = true;
    }
    d->connected = true;
    return KOFFT_OK;
}

int kofft_cuda_dist_connect_local(kofft_cuda_dist *const *dists, int world)
{
    for (int g = 0; g < world; g++)
        if (!dists[g] || dists[g]->world != world || dists[g]->rank != g || dists[g]->log2n != dists[0]->log2n)
            return fail_msg(KOFFT_ERR_INVALID_VALUE, "dists[g] must be rank g of the same world and size");
    for (int g = 0; g < world; g++) {
        CU(cudaSetDevice(dists[g]->ctx->device));
        for (int h = 0; h < world; h++) {
            dists[g]->peerA[h] = dists[h]->bufA;
            dists[g]->peerB[h] = dists[h]->bufB;
            const int dg = dists[g]->ctx->device, dh = dists[h]->ctx->device;
            if (dg != dh) {
                int can = 0;
                CU(cudaDeviceCanAccessPeer(&can, dg, dh));
                if (!can) return fail_msg(-static_cast<int>(cudaErrorPeerAccessUnsupported), "no peer access between the devices");
                cudaError_t e = cudaDeviceEnablePeerAccess(dh, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail_cuda(e, "cudaDeviceEnablePeerAccess");
                (void)cudaGetLastError();
            }
        }
        dists[g]->connected = true;
    }
    return KOFFT_OK;
}

int kofft_cuda_dist_phase(kofft_cuda_dist *d, int phase, const void *d_in, void *d_out, int inverse,
                          int natural_order, void *stream)
{
    if (!d->connected) return fail_msg(KOFFT_ERR_INVALID_VALUE, "connect the ranks first (connect_ipc / connect_local)");
    CU(cudaSetDevice(d->ctx->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int tw = inverse ? 2 : 1;
    switch (phase) {
    case 0:
        return dist_scatter(d, static_cast<const float2 *>(d_in), d->peerA, d->r1, d->c2, 0, s);
    case 1:
        return dist_fft_then_scatter(d, d->bufA, d->n1, d->c2, inverse, d->peerB, tw, s);
    case 2:
        if (!natural_order) return dist_local_fft(d, d->bufB, d_out, d->n2, d->r1, inverse, s);
        return dist_fft_then_scatter(d, d->bufB, d->n2, d->r1, inverse, d->peerA, 0, s);
    case 3: // natural order: the result is in buffer A; d_out == NULL or == buffer A leaves it there
        if (natural_order && d_out && d_out != d->bufA)
            CU(cudaMemcpyAsync(d_out, d->bufA, kofft_cuda_dist_shard_len(d) * sizeof(float2), cudaMemcpyDeviceToDevice, s));
        return KOFFT_OK;
    default:
        return fail_msg(KOFFT_ERR_INVALID_VALUE, "phase must be 0..3");
    }
}

// ---- the library-collective arm of the same transform (the baseline the P2P-store exchange is measured against):
// every exchange becomes  pack (local transpose [+ twiddle] into a send buffer laid out by destination)
// -> all-to-all by the CALLER (ncclAlltoAll / torch.distributed.all_to_all_single) -> unpack by the caller:
// what rank s sent to rank d arrives as a dense block [cb][rows_s]; the destination buffer is [cb][world][rows].
// step 0: d_src = this rank's input slice x [R1][N2]            (no twiddle)   blocks [C2][R1]
// step 1: d_src = buffer A [C2][N1] after its N1-point transforms (W_N^{n2 k1}) blocks [R1][C2]
// step 2: d_src = buffer B [R1][N2] after its N2-point transforms (no twiddle)  blocks [C2][R1]
int kofft_cuda_dist_pack(kofft_cuda_dist *d, int step, const void *d_src, void *d_send, int inverse, void *stream)
{
    CU(cudaSetDevice(d->ctx->device));
    if (step < 0 || step > 2) return fail_msg(KOFFT_ERR_INVALID_VALUE, "dist_pack: step 0..2");
    const size_t rows = step == 1 ? d->c2 : d->r1, cb = step == 1 ? d->r1 : d->c2;
    ScatterArgs a;
    a.src = static_cast<const float2 *>(d_src);
    for (int g = 0; g < d->world; g++) a.dst[g] = static_cast<float2 *>(d_send) + static_cast<size_t>(g) * rows * cb;
    a.rows = static_cast<long>(rows);
    a.cb = static_cast<long>(cb);
    a.world = d->world;
    a.rank = d->rank;
    a.dst_pitch = static_cast<long>(rows);
    a.dst_off = 0;
    a.twiddle = step == 1 ? (inverse ? 2 : 1) : 0;
    a.row0 = static_cast<long>(rows) * d->rank;
    a.log2n = d->log2n;
    a.llo = d->llo;
    a.tlo = d->tlo;
    a.thi = d->thi;
    (void)cudaGetLastError();
    cudaError_t e = launch_transpose_scatter(a, d->ctx->num_sms, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda(e, "dist_pack launch");
    d->ctx->launches++;
    return KOFFT_OK;
}

// the local transforms of a phase on their own: which = 0: the C2 rows of buffer A (N1 points each), 1: the R1 rows
// of buffer B (N2 points each); in place, correctly rounded tables (as kofft_cuda_dist_phase uses)
int kofft_cuda_dist_local_fft(kofft_cuda_dist *d, int which, int inverse, void *stream)
{
    CU(cudaSetDevice(d->ctx->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (which == 0) return dist_local_fft(d, d->bufA, d->bufA, d->n1, d->c2, inverse, s);
    return dist_local_fft(d, d->bufB, d->bufB, d->n2, d->r1, inverse, s);
}

int kofft_cuda_dist_run_local(kofft_cuda_dist *const *dists, int world, const void *const *d_in, void *const *d_out,
                              int inverse, int natural_order)
{
    for (int phase = 0; phase < 4; phase++) {
        for (int g = 0; g < world; g++) {
            int rc = kofft_cuda_dist_phase(dists[g], phase, d_in[g], d_out[g], inverse, natural_order, dists[g]->ctx->stream);
            if (rc) return rc;
        }
        for (int g = 0; g < world; g++) { // the barrier between phases
            CU(cudaSetDevice(dists[g]->ctx->device));
            CU(cudaStreamSynchronize(dists[g]->ctx->stream));
        }
    }
    return KOFFT_OK;
}

} // extern "C"
