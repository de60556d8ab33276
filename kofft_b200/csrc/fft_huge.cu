// fft_huge.cu -- power-of-two lengths above 2^16 on one GPU (kofft has no upper bound: its benchmark table goes to
// N = 2^20, benchmarks/README.md:5-7; dispatch src/fft.rs:1054-1082, Stockham :642-706, :789-912).
//
// The faithful radix-2 Stockham again, cut into passes that each make one trip through global memory:
//   pass A     stages 0 .. 7: the 256-point column pass of fft_large.cuh (any L: its column stride is a run-time
//              value), straight from the caller's rows through the IO policy (ifft conjugation, irfft untwist,
//              strided / SoA rows) into scratch buffer 0.
//   pass s     stages s .. s+r-1 (r <= 4) in registers: thread (k, j') owns the 2^r elements
//              i = k 2^(L-s) + q 2^(L-s-r) + j', runs r layers of the reference's butterflies with the reference's
//              table entries T[(k + c_low 2^s) 2^(L-1-s-t)] and writes i' = (k + c 2^s) 2^(L-s-r) + j'
//              (fft_engine.cuh derives this).  j' is the fastest thread index, so loads and stores are coalesced
//              runs in every pass (in the last pass j' is empty and consecutive threads are consecutive k: each
//              store instruction of a warp writes 256 contiguous bytes).  The passes ping-pong between two
//              scratch buffers; the last one stores through the IO policy (or, for rfft, is followed by the
//              Hermitian twist kernel, src/rfft.rs:450-463).
// HBM traffic is (1 + ceil((L-8)/4)) round trips instead of one, so these lengths run at a fraction of the
// roofline (measured numbers in DESIGN.md); the arithmetic is bit-identical to the reference like every other path.
#include "fft_large.cuh"
#include "launch.h"

namespace kofft {

namespace {

constexpr int kPassThreads = 256;

// R = log2 radix of this pass.  LAST: results leave through io.store(row, K, v); else dst[b * n + i'].
template <int R, bool EXACT, class IO, bool LAST>
__global__ void __launch_bounds__(kPassThreads)
    huge_pass_kernel(const __grid_constant__ IO io, const float2 *__restrict__ table, const float2 *__restrict__ src,
                     float2 *__restrict__ dst, int L, int s, long rows, long row0)
{
    constexpr int RAD = 1 << R;
    const int lj = L - s - R;              // log2 of the j' range
    const long per = 1L << (L - R);        // threads per transform
    const long total = rows * per;
    const long n = 1L << L;
    for (long g = blockIdx.x * (long)kPassThreads + threadIdx.x; g < total; g += (long)gridDim.x * kPassThreads) {
        const long b = g >> (L - R);
        const long bf = g & (per - 1);
        const long k = bf >> lj, j = bf & ((1L << lj) - 1);
        const float2 *in = src + b * n + (k << (L - s)) + j;
        float2 x[RAD];
#pragma unroll
        for (int q = 0; q < RAD; q++) x[q] = KOFFT_LDCG(in + ((long)q << lj));
#pragma unroll
        for (int tl = 0; tl < R; tl++) {
            const int bit = 1 << (R - 1 - tl);
#pragma unroll
            for (int w0 = 0; w0 < RAD; w0++) {
                if (w0 & bit) continue;
                const int c_low = bitrev(w0 >> (R - tl), tl);
                const float2 tw = KOFFT_LDG(table + ((k + ((long)c_low << s)) << (L - 1 - s - tl)));
                butterfly<EXACT>(x[w0], x[w0 | bit], tw);
            }
        }
        const long obase = (k << lj) + j;
#pragma unroll
        for (int w = 0; w < RAD; w++) {
            const long idx = obase + ((long)bitrev(w, R) << (s + lj));
            if constexpr (LAST)
                io.store(row0 + b, (int)idx, x[w]);
            else
                dst[b * n + idx] = x[w];
        }
    }
}

// rfft epilogue over Y = fft(m) of the packed rows: out[k] = twist(Y[k], Y[m-k]) (src/rfft.rs:450-463)
template <bool EXACT>
__global__ void __launch_bounds__(kPassThreads)
    huge_twist_kernel(const __grid_constant__ IoRfft<EXACT> io, const float2 *__restrict__ y, long rows, long row0)
{
    const long m = io.m;
    const long total = rows * (m + 1);
    for (long g = blockIdx.x * (long)kPassThreads + threadIdx.x; g < total; g += (long)gridDim.x * kPassThreads) {
        const long b = g / (m + 1), k = g - b * (m + 1);
        const float2 *Y = y + b * m;
        float2 *o = io.out + (row0 + b) * (m + 1);
        if (k == 0) {
            const float2 a = KOFFT_LDCG(Y);
            o[0] = make_float2(add_rn(a.x, a.y), 0.0f);
        } else if (k == m) {
            const float2 a = KOFFT_LDCG(Y);
            o[m] = make_float2(sub_rn(a.x, a.y), 0.0f);
        } else {
            o[k] = io.twist(KOFFT_LDCG(Y + k), KOFFT_LDCG(Y + (m - k)), KOFFT_LDG(io.rtw + k));
        }
    }
}

template <int R, bool EXACT, class IO, bool LAST>
cudaError_t launch_pass(const IO &io, const LaunchArgs &a, const float2 *src, float2 *dst, int L, int s, long rows, long row0)
{
    const long total = rows << (L - R);
    long blocks = (total + kPassThreads - 1) / kPassThreads;
    const long cap = (long)a.num_sms * 16;
    if (blocks > cap) blocks = cap;
    huge_pass_kernel<R, EXACT, IO, LAST><<<(unsigned)blocks, kPassThreads, 0, a.stream>>>(io, a.table, src, dst, L, s, rows, row0);
    return cudaGetLastError();
}

template <bool EXACT, class IO, bool LAST>
cudaError_t launch_pass_r(int r, const IO &io, const LaunchArgs &a, const float2 *src, float2 *dst, int L, int s, long rows, long row0)
{
    switch (r) {
    case 1: return launch_pass<1, EXACT, IO, LAST>(io, a, src, dst, L, s, rows, row0);
    case 2: return launch_pass<2, EXACT, IO, LAST>(io, a, src, dst, L, s, rows, row0);
    case 3: return launch_pass<3, EXACT, IO, LAST>(io, a, src, dst, L, s, rows, row0);
    default: return launch_pass<4, EXACT, IO, LAST>(io, a, src, dst, L, s, rows, row0);
    }
}

// TWIST: the last pass writes Y to scratch and the twist kernel produces the output
template <bool EXACT, class IO, bool TWIST>
cudaError_t run_chunk(const IO &io, const LaunchArgs &a, HugeArgs &g, int L, long rows, long row0)
{
    // pass A: the 256-point column pass (stages 0 .. 7)
    {
        using C = ColPass<EXACT, IO>;
        auto kern = colpass_kernel<EXACT, IO, false>;
        static PerDevice occ_pd;
        int &occ = occ_pd.get();
        if (occ == 0) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
            if (e != cudaSuccess) return e;
            int o = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, 256, C::SMEM_BYTES);
            if (e != cudaSuccess) return e;
            occ = o > 0 ? o : 1;
        }
        const long tiles = rows << (L - 8 - 4);
        const long cap = (long)occ * a.num_sms;
        kern<<<(unsigned)(tiles < cap ? tiles : cap), 256, C::SMEM_BYTES, a.stream>>>(io, a.tw0, a.table, L - 8, tiles, row0, g.scratch[0]);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        g.launches++;
    }
    int cur = 0;
    for (int s = 8; s < L;) {
        const int left = L - s;
        const int r = left >= 4 ? 4 : left;
        const bool last = s + r == L;
        cudaError_t e;
        if (last && !TWIST)
            e = launch_pass_r<EXACT, IO, true>(r, io, a, g.scratch[cur], nullptr, L, s, rows, row0);
        else
            e = launch_pass_r<EXACT, IO, false>(r, io, a, g.scratch[cur], g.scratch[cur ^ 1], L, s, rows, row0);
        if (e != cudaSuccess) return e;
        g.launches++;
        cur ^= 1;
        s += r;
    }
    if constexpr (TWIST) {
        const long total = rows * ((1L << L) + 1);
        long blocks = (total + kPassThreads - 1) / kPassThreads;
        const long cap = (long)a.num_sms * 16;
        if (blocks > cap) blocks = cap;
        huge_twist_kernel<EXACT><<<(unsigned)blocks, kPassThreads, 0, a.stream>>>(io, g.scratch[cur], rows, row0);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        g.launches++;
    }
    return cudaSuccess;
}

template <bool EXACT, class IO, bool TWIST>
cudaError_t run_all(const IO &io, const LaunchArgs &a, HugeArgs &g, int L)
{
    for (long r0 = 0; r0 < a.rows; r0 += g.chunk_rows) {
        const long nr = a.rows - r0 < g.chunk_rows ? a.rows - r0 : g.chunk_rows;
        cudaError_t e = run_chunk<EXACT, IO, TWIST>(io, a, g, L, nr, r0);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

template <bool EXACT>
cudaError_t launch_kind(const LaunchArgs &a, HugeArgs &g, int L)
{
    const IoArgs &q = a.io;
    switch (a.kind) {
    case KIND_C2C_FWD: {
        IoC2C<false> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale};
        return run_all<EXACT, IoC2C<false>, false>(io, a, g, L);
    }
    case KIND_C2C_INV: {
        IoC2C<true> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale};
        return run_all<EXACT, IoC2C<true>, false>(io, a, g, L);
    }
    case KIND_GEN_FWD: {
        IoGeneric<false> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2,
                            q.p0, q.p1, q.p2, q.p3, q.scale};
        return run_all<EXACT, IoGeneric<false>, false>(io, a, g, L);
    }
    case KIND_GEN_INV: {
        IoGeneric<true> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2,
                           q.p0, q.p1, q.p2, q.p3, q.scale};
        return run_all<EXACT, IoGeneric<true>, false>(io, a, g, L);
    }
    case KIND_RFFT: {
        IoRfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n};
        return run_all<EXACT, IoRfft<EXACT>, true>(io, a, g, L);
    }
    case KIND_IRFFT: {
        IoIrfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n, q.scale};
        return run_all<EXACT, IoIrfft<EXACT>, false>(io, a, g, L);
    }
    default:
        return cudaErrorNotSupported;
    }
}

// ---- f64 twin above 8192 points (ScalarFftImpl<f64>, src/fft.rs:914-1051): every stage range is a register pass
// through global memory (there is no f64 column pass); dense rows only.  First pass: conj on load for the inverse;
// last pass: conj * (1/n) on store (src/fft.rs:1163-1172).
template <int R, bool FIRST, bool LAST>
__global__ void __launch_bounds__(kPassThreads)
    huge_pass_f64_kernel(const double2 *__restrict__ table, const double2 *__restrict__ src, double2 *__restrict__ dst, int L,
                         int s, long rows, int inverse, double scale)
{
    constexpr int RAD = 1 << R;
    const int lj = L - s - R;
    const long per = 1L << (L - R);
    const long total = rows * per;
    const long n = 1L << L;
    for (long g = blockIdx.x * (long)kPassThreads + threadIdx.x; g < total; g += (long)gridDim.x * kPassThreads) {
        const long b = g >> (L - R);
        const long bf = g & (per - 1);
        const long k = bf >> lj, j = bf & ((1L << lj) - 1);
        const double2 *in = src + b * n + (k << (L - s)) + j;
        double2 x[RAD];
#pragma unroll
        for (int q = 0; q < RAD; q++) {
            x[q] = in[(long)q << lj];
            if (FIRST && inverse) x[q].y = -x[q].y;
        }
#pragma unroll
        for (int tl = 0; tl < R; tl++) {
            const int bit = 1 << (R - 1 - tl);
#pragma unroll
            for (int w0 = 0; w0 < RAD; w0++) {
                if (w0 & bit) continue;
                const int c_low = bitrev(w0 >> (R - tl), tl);
                butterfly_f64(x[w0], x[w0 | bit], __ldg(table + ((k + ((long)c_low << s)) << (L - 1 - s - tl))));
            }
        }
        double2 *o = dst + b * n + (k << lj) + j;
#pragma unroll
        for (int w = 0; w < RAD; w++) {
            double2 v = x[w];
            if (LAST && inverse) {
                v.y = -v.y;
                v.x = dmul(v.x, scale);
                v.y = dmul(v.y, scale);
            }
            o[(long)bitrev(w, R) << (s + lj)] = v;
        }
    }
}

template <int R>
cudaError_t launch_pass_f64(bool first, bool last, const LaunchF64Args &a, const double2 *src, double2 *dst, int L, int s)
{
    const long total = a.rows << (L - R);
    long blocks = (total + kPassThreads - 1) / kPassThreads;
    const long cap = (long)a.num_sms * 16;
    if (blocks > cap) blocks = cap;
    const unsigned gb = (unsigned)blocks;
    const int inv = a.inverse ? 1 : 0;
    if (first && last)
        return cudaErrorNotSupported; // L >= 14 always takes several passes
    if (first)
        huge_pass_f64_kernel<R, true, false><<<gb, kPassThreads, 0, a.stream>>>(a.table, src, dst, L, s, a.rows, inv, a.scale);
    else if (last)
        huge_pass_f64_kernel<R, false, true><<<gb, kPassThreads, 0, a.stream>>>(a.table, src, dst, L, s, a.rows, inv, a.scale);
    else
        huge_pass_f64_kernel<R, false, false><<<gb, kPassThreads, 0, a.stream>>>(a.table, src, dst, L, s, a.rows, inv, a.scale);
    return cudaGetLastError();
}

} // namespace

// a.in -> a.out through scratch[0] / scratch[1] (rows * n double2 each); launches: out
cudaError_t launch_huge_fft_f64(int L, const LaunchF64Args &a, double2 *scratch0, double2 *scratch1, int *launches)
{
    if (L < 5 || L > kHugeMaxLog2F64) return cudaErrorNotSupported;
    double2 *bufs[2] = {scratch0, scratch1};
    const double2 *src = a.in;
    int cur = 0;
    *launches = 0;
    for (int s = 0; s < L;) {
        const int left = L - s;
        const int r = left >= 4 ? 4 : left;
        const bool first = s == 0, last = s + r == L;
        double2 *dst = last ? a.out : bufs[cur];
        cudaError_t e;
        switch (r) {
        case 1: e = launch_pass_f64<1>(first, last, a, src, dst, L, s); break;
        case 2: e = launch_pass_f64<2>(first, last, a, src, dst, L, s); break;
        case 3: e = launch_pass_f64<3>(first, last, a, src, dst, L, s); break;
        default: e = launch_pass_f64<4>(first, last, a, src, dst, L, s); break;
        }
        if (e != cudaSuccess) return e;
        (*launches)++;
        src = dst;
        cur ^= 1;
        s += r;
    }
    return cudaSuccess;
}

cudaError_t launch_huge_fft(int L, const LaunchArgs &a, HugeArgs &g)
{
    if (L < 12 || L > kHugeMaxLog2) return cudaErrorNotSupported;
    g.launches = 0;
    return a.exact ? launch_kind<true>(a, g, L) : launch_kind<false>(a, g, L);
}

} // namespace kofft
