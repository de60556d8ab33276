// emu_kernels.cpp -- runs the *actual* kernel bodies (CtaFft::run, ColPass::run, RowPass::run)
// on the CPU with the cooperative block emulator in cuda_emu.h (CUDA threads = coroutines,
// __syncthreads = counting barrier, mbarrier / TMA bulk copy = host stand-ins).  Test
// scaffolding for the GPU-less CI: never shipped, never imported by kofft_b200/.
//
// Build: g++ -O1 -DKOFFT_EMU -ffp-contract=off -std=c++17 -shared -fPIC -I tests/emu
#include <cstring>
#include <vector>

#include "../../kofft_b200/csrc/fft_f64.cuh"
#include "../../kofft_b200/csrc/fft_large.cuh"
#include "../../kofft_b200/csrc/fft_split32.cuh"
#include "../../kofft_b200/csrc/fft_wide.cuh"
#include "../../kofft_b200/csrc/small_kernels.cuh"
#include "../../kofft_b200/csrc/istft_fused.cuh"

using namespace kofft;

namespace {

struct Args {
    const void *in, *in2;
    void *out, *out2;
    const void *aux;
    long n, p0, p1, p2, p3;
    float scale;
};

Tw0 make_tw0(int L, int R0, const float *table)
{
    Tw0 tw0;
    memset(&tw0, 0, sizeof tw0);
    for (int tl = 0; tl < R0; tl++)
        for (int c = 0; c < (1 << tl); c++) {
            long idx = (long)c << (L - 1 - tl);
            tw0.v[(1 << tl) - 1 + c] = make_float2(table[2 * idx], table[2 * idx + 1]);
        }
    return tw0;
}

// ---- single-CTA engine: the real CtaFft::run<STAGED> -------------------------------------------
template <int L, bool EXACT, class IO>
int run_cta(const IO &io, const float *table, long rows, bool staged, int grid)
{
    using P = Plan<L, IoTraits<IO>::kMinCta>; // the plan the shipped kernel uses for this IO
    Tw0 tw0 = make_tw0(L, P::R0, table);
    const long groups = (rows + P::TPC - 1) / P::TPC;
    if (grid > groups) grid = (int)groups;
    if (grid < 1) grid = 1;
    std::vector<float2> smem((P::SMEM_BYTES_STAGED + 256) / 8);
    float2 *sm = reinterpret_cast<float2 *>((reinterpret_cast<uintptr_t>(smem.data()) + 127) & ~uintptr_t(127));
    const float2 *tab = reinterpret_cast<const float2 *>(table);
    if constexpr (IO::kStageable && P::CAN_STAGE) {
        if (staged) {
            cuda_emu::launch(grid, P::CTA, [&] { CtaFft<P, EXACT, IO>::template run<true>(io, tw0, tab, rows, sm); });
            return 0;
        }
    }
    cuda_emu::launch(grid, P::CTA, [&] { CtaFft<P, EXACT, IO>::template run<false>(io, tw0, tab, rows, sm); });
    return 0;
}

template <bool EXACT, class IO>
int run_cta_sized(int L, const IO &io, const float *table, long rows, bool staged, int grid)
{
    switch (L) {
#define CASE_L(L) case L: return run_cta<L, EXACT>(io, table, rows, staged, grid);
    CASE_L(5) CASE_L(6) CASE_L(7) CASE_L(8) CASE_L(9) CASE_L(10) CASE_L(11) CASE_L(12) CASE_L(13) CASE_L(14)
#undef CASE_L
    default: return -1;
    }
}

// ---- two-pass large-N path: the real ColPass::run + RowPass::run --------------------------------
static bool g_large_staged = false; // the prefetching variants (cp.async column tiles, TMA row tiles)

template <int LB, bool EXACT, class IO, int EPI>
int run_large(const IO &io, const float *table, long rows, int grid_col, int grid_row)
{
    const int L = LB + LARGE_S1;
    const long n = 1L << L;
    Tw0 tw0 = make_tw0(L, 4, table);
    const float2 *tab = reinterpret_cast<const float2 *>(table);
    std::vector<float2> scratch((size_t)rows * n);
    using C = ColPass<EXACT, IO>;
    using R = RowPass<LB, EXACT, IO, EPI>;
    std::vector<float2> smem((std::max(C::SMEM_BYTES_STAGED, R::SMEM_BYTES_STAGED) + 256) / 8);
    float2 *sm = reinterpret_cast<float2 *>((reinterpret_cast<uintptr_t>(smem.data()) + 127) & ~uintptr_t(127));
    // process the batch in two chunks to exercise row0
    const long half = rows > 1 ? rows / 2 : rows;
    for (long r0 = 0; r0 < rows; r0 += half) {
        const long nr = std::min(half, rows - r0);
        const long tiles_c = nr << (LB - 4);
        int gc = (int)std::min<long>(grid_col, tiles_c);
        bool col_staged = false;
        if constexpr (IoTraits<IO>::kRowPtr) {
            if (g_large_staged) {
                col_staged = true;
                cuda_emu::launch(gc, 256, [&] { C::template run<true>(io, tw0, tab, LB, tiles_c, r0, scratch.data(), sm); });
            }
        }
        if (!col_staged) cuda_emu::launch(gc, 256, [&] { C::template run<false>(io, tw0, tab, LB, tiles_c, r0, scratch.data(), sm); });
        const long tiles_r = nr * R::NKB;
        long gr = std::min<long>(grid_row, tiles_r) / R::NKB * R::NKB;
        if (gr < R::NKB) gr = R::NKB;
        if (g_large_staged)
            cuda_emu::launch((int)gr, 256, [&] { R::template run<true>(io, tab, tiles_r, r0, scratch.data(), sm); });
        else
            cuda_emu::launch((int)gr, 256, [&] { R::template run<false>(io, tab, tiles_r, r0, scratch.data(), sm); });
    }
    return 0;
}

// the fused cluster kernel body: grid_col is reused as the number of clusters
template <int LB, bool EXACT, class IO, int EPI>
int run_fused(const IO &io, const float *table, long rows, int nclusters)
{
    using F = LargeFused<LB, EXACT, IO, EPI>;
    const int L = LB + LARGE_S1;
    const long n = 1L << L;
    Tw0 tw0 = make_tw0(L, 4, table);
    const float2 *tab = reinterpret_cast<const float2 *>(table);
    if (nclusters > rows) nclusters = (int)rows;
    if (nclusters < 1) nclusters = 1;
    std::vector<float2> scratch((size_t)nclusters * 2 * n);
    // one shared-memory block per CTA of the cluster
    const size_t per = (F::SMEM_BYTES + 255) / 8;
    std::vector<float2> smem(per * F::CLUSTER + 32);
    float2 *base = reinterpret_cast<float2 *>((reinterpret_cast<uintptr_t>(smem.data()) + 127) & ~uintptr_t(127));
    const unsigned grid = (unsigned)nclusters * F::CLUSTER;
    cuda_emu::launch(grid, 256, [&] {
        const unsigned rank = cuda_emu::cluster_rank();
        float2 *sm = base + (size_t)rank * ((per + 15) / 16 * 16);
        F::run(io, tw0, tab, rows, scratch.data(), sm, (int)rank, blockIdx.x / F::CLUSTER, gridDim.x / F::CLUSTER);
    }, F::CLUSTER);
    return 0;
}

// the pipelined persistent kernel body (LargePipe::run): `grid` CTAs = grid / NKB teams; the CTAs of a
// team run as interleaved coroutines (the emulator's cluster mode), so the dependency flags are live
static bool g_pipe = false;

template <int LB, bool EXACT, class IO, int EPI, bool STAGED>
int run_pipe_v(const IO &io, const float *table, long rows, int grid)
{
    using F = LargePipe<LB, EXACT, IO, EPI, STAGED>;
    const int L = LB + LARGE_S1;
    const long n = 1L << L;
    Tw0 tw0 = make_tw0(L, 4, table);
    const float2 *tab = reinterpret_cast<const float2 *>(table);
    long teams = grid / F::NKB;
    if (teams < 1) teams = 1;
    if (teams > rows) teams = rows;
    std::vector<float2> scratch((size_t)teams * F::SLOTS * n);
    std::vector<unsigned> flags((size_t)teams * F::FLAG_STRIDE, 0u);
    // one shared-memory block per CTA of the team
    const size_t per = ((F::SMEM_BYTES + 255) / 8 + 15) / 16 * 16;
    std::vector<float2> smem(per * F::NKB + 32);
    float2 *base = reinterpret_cast<float2 *>((reinterpret_cast<uintptr_t>(smem.data()) + 127) & ~uintptr_t(127));
    cuda_emu::launch((unsigned)(teams * F::NKB), 256, [&] {
        float2 *sm = base + (size_t)cuda_emu::cluster_rank() * per;
        F::run(io, tw0, tab, rows, scratch.data(), sm, flags.data());
    }, F::NKB);
    return 0;
}

template <int LB, bool EXACT, class IO, int EPI>
int run_pipe(const IO &io, const float *table, long rows, int grid)
{
    if constexpr (IoTraits<IO>::kRowPtr) {
        if (g_large_staged) return run_pipe_v<LB, EXACT, IO, EPI, true>(io, table, rows, grid);
    }
    return run_pipe_v<LB, EXACT, IO, EPI, false>(io, table, rows, grid);
}

static bool g_fused = false;

template <int LB, bool EXACT, class IO, int EPI>
int run_large_any(const IO &io, const float *table, long rows, int gc, int gr)
{
    if (g_pipe) return run_pipe<LB, EXACT, IO, EPI>(io, table, rows, gc);
    return g_fused ? run_fused<LB, EXACT, IO, EPI>(io, table, rows, gc) : run_large<LB, EXACT, IO, EPI>(io, table, rows, gc, gr);
}

template <int LB, bool EXACT>
int run_large_kind(int kind, const Args &q, const float *table, long rows, int gc, int gr)
{
    switch (kind) {
    case 0: { IoC2C<false> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale}; return run_large_any<LB, EXACT, IoC2C<false>, ROW_STORE>(io, table, rows, gc, gr); }
    case 1: { IoC2C<true> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale}; return run_large_any<LB, EXACT, IoC2C<true>, ROW_STORE>(io, table, rows, gc, gr); }
    case 2: { IoGeneric<false> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2, q.p0, q.p1, q.p2, q.p3, q.scale}; return run_large_any<LB, EXACT, IoGeneric<false>, ROW_STORE>(io, table, rows, gc, gr); }
    case 3: { IoGeneric<true> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2, q.p0, q.p1, q.p2, q.p3, q.scale}; return run_large_any<LB, EXACT, IoGeneric<true>, ROW_STORE>(io, table, rows, gc, gr); }
    case 6: { IoRfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n}; return run_large_any<LB, EXACT, IoRfft<EXACT>, ROW_TWIST>(io, table, rows, gc, gr); }
    case 7: { IoIrfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n, q.scale}; return run_large_any<LB, EXACT, IoIrfft<EXACT>, ROW_STORE>(io, table, rows, gc, gr); }
    default: return -2;
    }
}

template <bool EXACT>
int run_cta_kind(int kind, int L, const Args &q, const float *table, long rows, bool staged, int grid)
{
    switch (kind) {
    case 0: { IoC2C<false> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale}; return run_cta_sized<EXACT>(L, io, table, rows, staged, grid); }
    case 1: { IoC2C<true> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale}; return run_cta_sized<EXACT>(L, io, table, rows, staged, grid); }
    case 2: { IoGeneric<false> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2, q.p0, q.p1, q.p2, q.p3, q.scale}; return run_cta_sized<EXACT>(L, io, table, rows, staged, grid); }
    case 3: { IoGeneric<true> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2, q.p0, q.p1, q.p2, q.p3, q.scale}; return run_cta_sized<EXACT>(L, io, table, rows, staged, grid); }
    case 4: { IoStft io{(const float *)q.in, (const float *)q.aux, (float2 *)q.out, q.p0, q.p1, q.p2, q.n}; return run_cta_sized<EXACT>(L, io, table, rows, staged, grid); }
    case 5: { IoIstft io{(const float2 *)q.in, (const float *)q.aux, (float *)q.out, q.n, q.scale}; return run_cta_sized<EXACT>(L, io, table, rows, staged, grid); }
    case 8: { IoStftMag io{{(const float *)q.in, (const float *)q.aux, nullptr, q.p0, q.p1, q.p2, q.n}, (float *)q.out, (int *)q.out2, 0.0f}; return run_cta_sized<EXACT>(L, io, table, rows, staged, grid); }
    case 6: { IoRfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n}; return run_cta_sized<EXACT>(L, io, table, rows, staged, grid); }
    case 7: { IoIrfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n, q.scale}; return run_cta_sized<EXACT>(L, io, table, rows, staged, grid); }
    default: return -2;
    }
}

} // namespace

#define API extern "C" __attribute__((visibility("default")))

// the real single-CTA kernel body; grid = number of (persistent) CTAs to emulate
API int kofft_emuk_cta(int kind, int exact, int L, long rows, const void *in, const void *in2, void *out, void *out2,
                       const void *aux, long p0, long p1, long p2, long p3, float scale, const float *table,
                       int staged, int grid)
{
    Args q{in, in2, out, out2, aux, 1L << L, p0, p1, p2, p3, scale};
    return exact ? run_cta_kind<true>(kind, L, q, table, rows, staged != 0, grid)
                 : run_cta_kind<false>(kind, L, q, table, rows, staged != 0, grid);
}

API void kofft_emuk_set_fused(int fused) { g_fused = fused != 0; }
API void kofft_emuk_set_pipe(int pipe) { g_pipe = pipe != 0; }
// CTAs with team rank >= late_from run `ratio` times slower than the others (0 / 1 = off)
API void kofft_emuk_set_skew(int late_from, int ratio)
{
    cuda_emu::g_late_from = ratio > 1 ? (unsigned)late_from : ~0u;
    cuda_emu::g_late_ratio = ratio > 1 ? (unsigned long long)ratio : 1;
}
API void kofft_emuk_set_large_staged(int staged) { g_large_staged = staged != 0; }

// the real two-pass kernel bodies; L = 15 or 16 is the length of the complex core
API int kofft_emuk_large(int kind, int exact, int L, long rows, const void *in, const void *in2, void *out,
                         void *out2, const void *aux, long p0, long p1, long p2, long p3, float scale,
                         const float *table, int grid_col, int grid_row)
{
    Args q{in, in2, out, out2, aux, 1L << L, p0, p1, p2, p3, scale};
    if (L == 15)
        return exact ? run_large_kind<7, true>(kind, q, table, rows, grid_col, grid_row)
                     : run_large_kind<7, false>(kind, q, table, rows, grid_col, grid_row);
    if (L == 16)
        return exact ? run_large_kind<8, true>(kind, q, table, rows, grid_col, grid_row)
                     : run_large_kind<8, false>(kind, q, table, rows, grid_col, grid_row);
    return -1;
}


// ---- the warp-specialised split kernel (fft_split32.cuh): Split32::run on `grid` CTAs of 512 threads; the CTAs
// of a team run as interleaved coroutines so the A-role / B-role dependency flags are live
static bool g_split_staged = false;

template <int LA, bool EXACT, class IO, int EPI, bool STAGED, bool PRE = false>
static int run_split32_v(const IO &io, const float *table, long rows, int grid)
{
    using F = Split32<LA, EXACT, IO, EPI, STAGED, PRE>;
    const long n = 1L << F::L;
    Tw0W tw0;
    memset(&tw0, 0, sizeof tw0);
    for (int tl = 0; tl < F::RA0; tl++)
        for (int c = 0; c < (1 << tl); c++) {
            long idx = (long)c << (F::L - 1 - tl);
            tw0.v[(1 << tl) - 1 + c] = make_float2(table[2 * idx], table[2 * idx + 1]);
        }
    const float2 *tab = reinterpret_cast<const float2 *>(table);
    long teams = grid / F::NT;
    if (teams < 1) teams = 1;
    if (teams > rows) teams = rows;
    std::vector<float2> scratch((size_t)teams * (F::SLOTS + (PRE ? F::ZSLOTS : 0)) * n);
    std::vector<unsigned> flags((size_t)teams * F::FLAG_STRIDE, 0u);
    const size_t per = ((F::SMEM_BYTES + 255) / 8 + 15) / 16 * 16;
    std::vector<float2> smem(per * F::NT + 32);
    float2 *base = reinterpret_cast<float2 *>((reinterpret_cast<uintptr_t>(smem.data()) + 127) & ~uintptr_t(127));
    const size_t keep = cuda_emu::g_stack_bytes;
    cuda_emu::g_stack_bytes = 64 * 1024;
    TmaMap map;
    memset(&map, 0, sizeof map);
    if constexpr (STAGED) { // the [rows * 2^LA][32 complex] view of the input, boxes of 256 rows x 8 complex
        if constexpr (PRE) { // the teams' untwisted rows behind the intermediate
            map.base = scratch.data() + (size_t)teams * F::SLOTS * n;
            map.dim1 = (unsigned long long)(teams * F::ZSLOTS) << LA;
        } else {
            map.base = io.row_ptr(0);
            map.dim1 = (unsigned long long)rows << LA;
        }
        map.dim0 = 64;
        map.stride1 = 256;
        map.box0 = 2 * F::COLS;
        map.box1 = 256;
    }
    cuda_emu::launch((unsigned)(teams * F::NT), F::CTA, [&] {
        float2 *sm = base + (size_t)cuda_emu::cluster_rank() * per;
        F::run(io, tw0, tab, rows, scratch.data(), sm, flags.data(), &map);
    }, F::NT);
    cuda_emu::g_stack_bytes = keep;
    return 0;
}

template <int LA, bool EXACT, class IO, int EPI>
static int run_split32(const IO &io, const float *table, long rows, int grid)
{
    if constexpr (LA == 10 && IoTraits<IO>::kRowPtr) {
        if (g_split_staged) return run_split32_v<LA, EXACT, IO, EPI, true>(io, table, rows, grid);
    }
    return run_split32_v<LA, EXACT, IO, EPI, false>(io, table, rows, grid);
}

template <int LA, bool EXACT>
static int run_split32_kind(int kind, const Args &q, const float *table, long rows, int grid)
{
    switch (kind) {
    case 0: { IoC2C<false> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale}; return run_split32<LA, EXACT, IoC2C<false>, SPLIT_STORE>(io, table, rows, grid); }
    case 1: { IoC2C<true> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale}; return run_split32<LA, EXACT, IoC2C<true>, SPLIT_STORE>(io, table, rows, grid); }
    case 2: { IoGeneric<false> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2, q.p0, q.p1, q.p2, q.p3, q.scale}; return run_split32<LA, EXACT, IoGeneric<false>, SPLIT_STORE>(io, table, rows, grid); }
    case 3: { IoGeneric<true> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2, q.p0, q.p1, q.p2, q.p3, q.scale}; return run_split32<LA, EXACT, IoGeneric<true>, SPLIT_STORE>(io, table, rows, grid); }
    case 6: { IoRfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n}; return run_split32<LA, EXACT, IoRfft<EXACT>, SPLIT_TWIST>(io, table, rows, grid); }
    case 7: {
        IoIrfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n, q.scale};
        if constexpr (LA == 10) {
            if (g_split_staged) return run_split32_v<LA, EXACT, IoIrfft<EXACT>, SPLIT_STORE, true, true>(io, table, rows, grid);
        }
        return run_split32<LA, EXACT, IoIrfft<EXACT>, SPLIT_STORE>(io, table, rows, grid);
    }
    default: return -2;
    }
}

// L = 13 .. 15 is the length of the complex core
API int kofft_emuk_split32(int kind, int exact, int L, long rows, const void *in, const void *in2, void *out, void *out2,
                           const void *aux, long p0, long p1, long p2, long p3, float scale, const float *table,
                           int grid, int staged)
{
    g_split_staged = staged != 0;
    Args q{in, in2, out, out2, aux, 1L << L, p0, p1, p2, p3, scale};
    switch (L) {
    case 13: return exact ? run_split32_kind<8, true>(kind, q, table, rows, grid) : run_split32_kind<8, false>(kind, q, table, rows, grid);
    case 14: return exact ? run_split32_kind<9, true>(kind, q, table, rows, grid) : run_split32_kind<9, false>(kind, q, table, rows, grid);
    case 15: return exact ? run_split32_kind<10, true>(kind, q, table, rows, grid) : run_split32_kind<10, false>(kind, q, table, rows, grid);
    default: return -1;
    }
}

// ---- the wide single-CTA kernel (fft_wide.cuh): WideCta::run on `grid` CTAs of N / 32 threads
static bool g_wide_staged = false;
template <int L, bool EXACT, class IO, bool STAGED>
static int run_wide_v(const IO &io, const float *table, long rows, int grid)
{
    using F = WideCta<L, EXACT, IO, STAGED>;
    Tw0W tw0;
    memset(&tw0, 0, sizeof tw0);
    for (int tl = 0; tl < 5; tl++)
        for (int c = 0; c < (1 << tl); c++) {
            long idx = (long)c << (L - 1 - tl);
            tw0.v[(1 << tl) - 1 + c] = make_float2(table[2 * idx], table[2 * idx + 1]);
        }
    std::vector<float2> smem(F::SMEM_BYTES / 8 + 32);
    float2 *sm = reinterpret_cast<float2 *>((reinterpret_cast<uintptr_t>(smem.data()) + 127) & ~uintptr_t(127));
    const float2 *tab = reinterpret_cast<const float2 *>(table);
    if (grid > rows) grid = (int)rows;
    if (grid < 1) grid = 1;
    const size_t keep = cuda_emu::g_stack_bytes;
    cuda_emu::g_stack_bytes = 64 * 1024;
    cuda_emu::launch(grid, F::CTA, [&] { F::run(io, tw0, tab, rows, sm); });
    cuda_emu::g_stack_bytes = keep;
    return 0;
}

template <int L, bool EXACT, class IO>
static int run_wide(const IO &io, const float *table, long rows, int grid)
{
    return g_wide_staged ? run_wide_v<L, EXACT, IO, true>(io, table, rows, grid) : run_wide_v<L, EXACT, IO, false>(io, table, rows, grid);
}

template <int L, bool EXACT>
static int run_wide_kind(int kind, const Args &q, const float *table, long rows, int grid)
{
    switch (kind) {
    case 0: { IoC2C<false> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale}; return run_wide<L, EXACT>(io, table, rows, grid); }
    case 1: { IoC2C<true> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale}; return run_wide<L, EXACT>(io, table, rows, grid); }
    case 2: { IoGeneric<false> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2, q.p0, q.p1, q.p2, q.p3, q.scale}; return run_wide_v<L, EXACT, IoGeneric<false>, false>(io, table, rows, grid); }
    case 3: { IoGeneric<true> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2, q.p0, q.p1, q.p2, q.p3, q.scale}; return run_wide_v<L, EXACT, IoGeneric<true>, false>(io, table, rows, grid); }
    case 6: { IoRfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n}; return run_wide<L, EXACT>(io, table, rows, grid); }
    case 7: { IoIrfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n, q.scale}; return run_wide_v<L, EXACT, IoIrfft<EXACT>, false>(io, table, rows, grid); }
    default: return -2;
    }
}

// same argument list as kofft_emuk_split32; L = 13, 14 is the length of the complex core
API int kofft_emuk_wide(int kind, int exact, int L, long rows, const void *in, const void *in2, void *out, void *out2,
                        const void *aux, long p0, long p1, long p2, long p3, float scale, const float *table,
                        int grid, int staged)
{
    g_wide_staged = staged != 0;
    Args q{in, in2, out, out2, aux, 1L << L, p0, p1, p2, p3, scale};
    switch (L) {
    case 13: return exact ? run_wide_kind<13, true>(kind, q, table, rows, grid) : run_wide_kind<13, false>(kind, q, table, rows, grid);
    case 14: return exact ? run_wide_kind<14, true>(kind, q, table, rows, grid) : run_wide_kind<14, false>(kind, q, table, rows, grid);
    default: return -1;
    }
}

// ---- bank audit of the wide kernel's shared-memory exchanges: for every access of a half-warp (16 lanes x 8 bytes) the
// number of lanes that share an 8-byte bank pair with another lane, using the kernel's own index functions
template <int L>
static int wide_bank_audit(int staged)
{
    using F = WideCta<L, true, IoC2C<false>, false>;
    using P0 = typename F::P0;
    using P1 = typename F::P1;
    using P2 = typename F::P2;
    int worst = 0;
    auto audit = [&](auto addr_of_lane) { // addr_of_lane(lane in 0..15) -> float2 index
        int cnt[16] = {0};
        for (int l = 0; l < 16; l++) cnt[addr_of_lane(l) & 15]++;
        for (int b = 0; b < 16; b++)
            if (cnt[b] - 1 > worst) worst = cnt[b] - 1;
    };
    constexpr int CTA = F::CTA;
    for (int t0 = 0; t0 < CTA; t0 += 16)
        for (int q = 0; q < 32; q++) {
            if (staged) {
                const int c = bitrev(q, 5);
                audit([&](int l) { const int t = t0 + l; return t + q * CTA; });                                      // pass-0 load
                audit([&](int l) { const int t = t0 + l; return ((F::SWAP && (c & 1)) ? (t ^ 8) : t) + c * CTA; });  // pass-0 store
                audit([&](int l) {                                                                                   // pass-1 load
                    const int t = t0 + l, k1 = t >> P1::LJ, sb = F::SWAP ? (k1 & 1) : 0;
                    const int base = (k1 << (L - 5)) + (t & (P1::J - 1));
                    return base + ((q & 1) ? -(sb << P1::LJ) : (sb << P1::LJ)) + (q << P1::LJ);
                });
            } else {
                audit([&](int l) { return F::pad_a(P0::dst_index(t0 + l, 0, q)); });
                audit([&](int l) { return F::pad_a(P1::src_index(t0 + l, 0, q)); });
            }
            audit([&](int l) { return F::pad_b(P1::dst_index(t0 + l, 0, q)); }); // pass-1 store
        }
    for (int t0 = 0; t0 < CTA; t0 += 16)
        for (int u = 0; u < P2::U; u++)
            for (int q = 0; q < P2::R; q++) {
                audit([&](int l) { return F::pad_b(P2::src_index(t0 + l, u, q)); }); // pass-2 load
                const int c = bitrev(q, F::R2), HR = P2::R / 2;                      // rfft side buffer
                if (c >= HR)
                    audit([&](int l) { return (c - HR) * 1024 + P2::bfly(t0 + l, u); });
                else
                    audit([&](int l) {
                        const int b = P2::bfly(t0 + l, u);
                        return b == 0 ? (HR - c) * 1024 : (HR - 1 - c) * 1024 + (1024 - b);
                    });
            }
    return worst;
}

API int kofft_emuk_wide_bank_audit(int L, int staged) { return L == 13 ? wide_bank_audit<13>(staged) : (L == 14 ? wide_bank_audit<14>(staged) : -1); }

// the real fused istft kernel body (istft_fused.cuh)
template <int L, bool EXACT>
static int run_istft(const IstftFusedArgs &a, const float *table, int grid)
{
    using K = IstftFused<L, EXACT>;
    Tw0 tw0 = make_tw0(L, Plan<L>::R0, table);
    std::vector<float2> smem((K::smem_bytes(a.hop) + 256) / 8);
    float2 *sm = reinterpret_cast<float2 *>((reinterpret_cast<uintptr_t>(smem.data()) + 127) & ~uintptr_t(127));
    const float2 *tab = reinterpret_cast<const float2 *>(table);
    cuda_emu::launch(grid, K::CTA, [&] { K::run(a, tw0, tab, sm); });
    return 0;
}

API int kofft_emuk_istft_fused(int exact, int L, const void *frames, const float *window, float *output, float *norm,
                               long channels, long nframes, long hop, long out_len, int run_frames,
                               int zero_uncovered, const float *table, int grid)
{
    IstftFusedArgs a;
    a.frames = (const float2 *)frames;
    a.window = window;
    a.output = output;
    a.norm = norm;
    a.channels = channels;
    a.nframes = nframes;
    a.hop = hop;
    a.out_len = out_len;
    a.run_frames = run_frames;
    a.zero_uncovered = zero_uncovered;
    a.scale = 1.0f / (float)(1L << L);
    switch (L) {
#define CASE_L(L) case L: return exact ? run_istft<L, true>(a, table, grid) : run_istft<L, false>(a, table, grid);
    CASE_L(9) CASE_L(10) CASE_L(11) CASE_L(12)
#undef CASE_L
    default: return -1;
    }
}

// ---- f64 twin: the real CtaFftD::run (fft_f64.cuh) and the N <= 16 literal kernels for double2 ----
static bool g_f64_staged = false;
API void kofft_emuk_set_f64_staged(int staged) { g_f64_staged = staged != 0; }
template <int L, class IO>
static int run_f64_L(const IO &io, const double *table, long rows, int grid)
{
    using P = PlanD<L>;
    Tw0D tw0;
    memset(&tw0, 0, sizeof tw0);
    for (int tl = 0; tl < P::R0; tl++)
        for (int c = 0; c < (1 << tl); c++) {
            long idx = (long)c << (L - 1 - tl);
            tw0.v[(1 << tl) - 1 + c] = make_double2(table[2 * idx], table[2 * idx + 1]);
        }
    const long groups = (rows + P::TPC - 1) / P::TPC;
    if (grid > groups) grid = (int)groups;
    if (grid < 1) grid = 1;
    std::vector<double2> smem(P::SMEM_BYTES / 16 + 32);
    const double2 *tab = reinterpret_cast<const double2 *>(table);
    double2 *sm = reinterpret_cast<double2 *>((reinterpret_cast<uintptr_t>(smem.data()) + 127) & ~uintptr_t(127));
    if constexpr (IO::kStageable) {
        if (g_f64_staged) {
            cuda_emu::launch(grid, P::CTA, [&] { CtaFftD<L, IO>::template run<true>(io, tw0, tab, rows, sm); });
            return 0;
        }
    }
    cuda_emu::launch(grid, P::CTA, [&] { CtaFftD<L, IO>::template run<false>(io, tw0, tab, rows, sm); });
    return 0;
}

template <int N, class IO>
static int run_f64_small(const IO &io, long rows)
{
    for (long r = 0; r < rows; r++) small_transform<N, true, IO>(io, r);
    return 0;
}

template <class IO>
static int run_f64_io(const IO &io, long n, const double *table, long rows, int grid)
{
    switch (n) {
    case 1: return run_f64_small<1>(io, rows);
    case 2: return run_f64_small<2>(io, rows);
    case 4: return run_f64_small<4>(io, rows);
    case 8: return run_f64_small<8>(io, rows);
    case 16: return run_f64_small<16>(io, rows);
#define CASE_L(L) case (1 << L): return run_f64_L<L>(io, table, rows, grid);
    CASE_L(5) CASE_L(6) CASE_L(7) CASE_L(8) CASE_L(9) CASE_L(10) CASE_L(11) CASE_L(12) CASE_L(13)
#undef CASE_L
    default: return -1;
    }
}

// split (SoA) rows through IoGenericD: re / im [rows][n]
API int kofft_emuk_f64_split(long n, long rows, const double *in_re, const double *in_im, double *out_re, double *out_im,
                             int inverse, double scale, const double *table, int grid)
{
    const bool keep = g_f64_staged;
    g_f64_staged = false; // generic rows are never staged
    int rc;
    if (inverse) {
        IoGenericD<true> io{in_re, in_im, out_re, out_im, 1, n, 1, n, scale};
        rc = run_f64_io(io, n, table, rows, grid);
    } else {
        IoGenericD<false> io{in_re, in_im, out_re, out_im, 1, n, 1, n, scale};
        rc = run_f64_io(io, n, table, rows, grid);
    }
    g_f64_staged = keep;
    return rc;
}

// f64 real transforms: m = n/2 is the engine length; which = 1 rfft (in [rows][n] doubles, out [rows][m+1] complex),
// 2 irfft (the reverse); rtw: T' table of m entries
API int kofft_emuk_f64_real(long m, long rows, const void *in, void *out, int which, double scale, const double *table,
                            const double *rtw, int grid)
{
    if (which == 1) {
        IoRfftD io{(const double2 *)in, (double2 *)out, (const double2 *)rtw, m};
        return run_f64_io(io, m, table, rows, grid);
    }
    const bool keep = g_f64_staged;
    g_f64_staged = false;
    IoIrfftD io{(const double2 *)in, (double2 *)out, (const double2 *)rtw, m, scale};
    int rc = run_f64_io(io, m, table, rows, grid);
    g_f64_staged = keep;
    return rc;
}

// in / out: [rows][n] complex doubles; table: FftPlanner<f64> table of n (n >= 32)
API int kofft_emuk_f64(long n, long rows, const void *in, void *out, int inverse, double scale, const double *table, int grid)
{
    if (inverse) {
        IoC2CD<true> io{(const double2 *)in, (double2 *)out, n, scale};
        return run_f64_io(io, n, table, rows, grid);
    }
    IoC2CD<false> io{(const double2 *)in, (double2 *)out, n, scale};
    return run_f64_io(io, n, table, rows, grid);
}
