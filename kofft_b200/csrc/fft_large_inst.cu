// fft_large_inst.cu -- instantiates the two-pass kernels (fft_large.cuh) for N = 2^15, 2^16.
#include "fft_large.cuh"
#include "launch.h"

namespace kofft {

namespace {

template <class K>
cudaError_t prep(K kern, int smem, int threads, int *occ)
{
    if (*occ == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        int o = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, threads, smem);
        if (e != cudaSuccess) return e;
        *occ = o > 0 ? o : 1;
    }
    return cudaSuccess;
}

template <bool EXACT, class IO>
cudaError_t launch_col(const IO &io, const LaunchArgs &a, const LargeArgs &g)
{
    using C = ColPass<EXACT, IO>;
    auto kern = colpass_kernel<EXACT, IO>;
    static int occ = 0;
    cudaError_t e = prep(kern, C::SMEM_BYTES, 256, &occ);
    if (e != cudaSuccess) return e;
    const long tiles = g.chunk_rows << (g.lsub - 4);
    long cap = a.max_ctas > 0 ? a.max_ctas : (long)occ * a.num_sms;
    int grid = (int)(tiles < cap ? tiles : cap);
    kern<<<grid, 256, C::SMEM_BYTES, a.stream>>>(io, a.tw0, a.table, g.lsub, tiles, g.row0, g.scratch);
    return cudaGetLastError();
}

template <int LB, bool EXACT, class IO, int EPI>
cudaError_t launch_row(const IO &io, const LaunchArgs &a, const LargeArgs &g)
{
    using R = RowPass<LB, EXACT, IO, EPI>;
    auto kern = rowpass_kernel<LB, EXACT, IO, EPI>;
    static int occ = 0;
    cudaError_t e = prep(kern, R::SMEM_BYTES, 256, &occ);
    if (e != cudaSuccess) return e;
    const long tiles = g.chunk_rows * R::NKB;
    long cap = a.max_ctas > 0 ? a.max_ctas : (long)occ * a.num_sms;
    long grid = tiles < cap ? tiles : cap;
    grid = grid / R::NKB * R::NKB; // a CTA keeps its k-block: the grid is a multiple of NKB
    if (grid < R::NKB) grid = R::NKB;
    kern<<<(int)grid, 256, R::SMEM_BYTES, a.stream>>>(io, a.table, tiles, g.row0, g.scratch);
    return cudaGetLastError();
}

template <int LB, bool EXACT>
cudaError_t launch_pair(const LaunchArgs &a, const LargeArgs &g)
{
    const IoArgs &q = a.io;
    cudaError_t e;
    switch (a.kind) {
    case KIND_C2C_FWD: {
        IoC2C<false> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale};
        if ((e = launch_col<EXACT>(io, a, g)) != cudaSuccess) return e;
        return launch_row<LB, EXACT, IoC2C<false>, ROW_STORE>(io, a, g);
    }
    case KIND_C2C_INV: {
        IoC2C<true> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale};
        if ((e = launch_col<EXACT>(io, a, g)) != cudaSuccess) return e;
        return launch_row<LB, EXACT, IoC2C<true>, ROW_STORE>(io, a, g);
    }
    case KIND_GEN_FWD: {
        IoGeneric<false> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2,
                            q.p0, q.p1, q.p2, q.p3, q.scale};
        if ((e = launch_col<EXACT>(io, a, g)) != cudaSuccess) return e;
        return launch_row<LB, EXACT, IoGeneric<false>, ROW_STORE>(io, a, g);
    }
    case KIND_GEN_INV: {
        IoGeneric<true> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2,
                           q.p0, q.p1, q.p2, q.p3, q.scale};
        if ((e = launch_col<EXACT>(io, a, g)) != cudaSuccess) return e;
        return launch_row<LB, EXACT, IoGeneric<true>, ROW_STORE>(io, a, g);
    }
    case KIND_RFFT: {
        IoRfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n};
        if ((e = launch_col<EXACT>(io, a, g)) != cudaSuccess) return e;
        return launch_row<LB, EXACT, IoRfft<EXACT>, ROW_TWIST>(io, a, g);
    }
    case KIND_IRFFT: {
        IoIrfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n, q.scale};
        if ((e = launch_col<EXACT>(io, a, g)) != cudaSuccess) return e;
        return launch_row<LB, EXACT, IoIrfft<EXACT>, ROW_STORE>(io, a, g);
    }
    default:
        return cudaErrorNotSupported;
    }
}

} // namespace

cudaError_t launch_large_fft(int L, const LaunchArgs &a, const LargeArgs &g)
{
    switch (L) {
    case 15: return a.exact ? launch_pair<7, true>(a, g) : launch_pair<7, false>(a, g);
    case 16: return a.exact ? launch_pair<8, true>(a, g) : launch_pair<8, false>(a, g);
    default: return cudaErrorNotSupported;
    }
}

} // namespace kofft
