// fft_large_inst.cu -- instantiates the two-pass kernels (fft_large.cuh) for N = 2^15, 2^16.
#include <cstdio>
#include <cstdlib>

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

template <bool EXACT, class IO, bool STAGED>
cudaError_t launch_col_v(const IO &io, const LaunchArgs &a, const LargeArgs &g)
{
    using C = ColPass<EXACT, IO>;
    constexpr int smem = STAGED ? C::SMEM_BYTES_STAGED : C::SMEM_BYTES;
    auto kern = colpass_kernel<EXACT, IO, STAGED>;
    static PerDevice occ_pd;
    int &occ = occ_pd.get();
    cudaError_t e = prep(kern, smem, 256, &occ);
    if (e != cudaSuccess) return e;
    const long tiles = g.chunk_rows << (g.lsub - 4);
    long cap = a.max_ctas > 0 ? a.max_ctas : (long)occ * a.num_sms;
    int grid = (int)(tiles < cap ? tiles : cap);
    kern<<<grid, 256, smem, a.stream>>>(io, a.tw0, a.table, g.lsub, tiles, g.row0, g.scratch);
    return cudaGetLastError();
}

// a.staged: the host verified 16-byte alignment of the rows (asynchronous 16-byte copies)
template <bool EXACT, class IO>
cudaError_t launch_col(const IO &io, const LaunchArgs &a, const LargeArgs &g)
{
    if constexpr (IoTraits<IO>::kRowPtr) {
        if (a.staged) return launch_col_v<EXACT, IO, true>(io, a, g);
    }
    return launch_col_v<EXACT, IO, false>(io, a, g);
}

template <int LB, bool EXACT, class IO, int EPI, bool STAGED>
cudaError_t launch_row_v(const IO &io, const LaunchArgs &a, const LargeArgs &g);

template <int LB, bool EXACT, class IO, int EPI>
cudaError_t launch_row(const IO &io, const LaunchArgs &a, const LargeArgs &g)
{
    // the intermediate is the library's own 16-byte aligned scratch: staging is always possible.  The
    // twist variant spends its shared memory on the CTA's rfft twiddles instead (two CTAs per SM).
    if constexpr (EPI != ROW_TWIST) {
        if (g.stage_rows) return launch_row_v<LB, EXACT, IO, EPI, true>(io, a, g);
    }
    return launch_row_v<LB, EXACT, IO, EPI, false>(io, a, g);
}

template <int LB, bool EXACT, class IO, int EPI, bool STAGED>
cudaError_t launch_row_v(const IO &io, const LaunchArgs &a, const LargeArgs &g)
{
    using R = RowPass<LB, EXACT, IO, EPI>;
    constexpr int smem_bytes = STAGED ? R::SMEM_BYTES_STAGED : R::SMEM_BYTES;
    auto kern = rowpass_kernel<LB, EXACT, IO, EPI, STAGED>;
    static PerDevice occ_pd;
    int &occ = occ_pd.get();
    cudaError_t e = prep(kern, smem_bytes, 256, &occ);
    if (e != cudaSuccess) return e;
    const long tiles = g.chunk_rows * R::NKB;
    long cap = a.max_ctas > 0 ? a.max_ctas : (long)occ * a.num_sms;
    long grid = tiles < cap ? tiles : cap;
    grid = grid / R::NKB * R::NKB; // a CTA keeps its k-block: the grid is a multiple of NKB
    if (grid < R::NKB) grid = R::NKB;
    kern<<<(int)grid, 256, smem_bytes, a.stream>>>(io, a.table, tiles, g.row0, g.scratch);
    return cudaGetLastError();
}

// one persistent launch: clusters of NKB CTAs, one transform per cluster at a time
template <int LB, bool EXACT, class IO, int EPI>
cudaError_t launch_fused(const IO &io, const LaunchArgs &a, const LargeArgs &g)
{
    using F = LargeFused<LB, EXACT, IO, EPI>;
    auto kern = large_fused_kernel<LB, EXACT, IO, EPI>;
    static PerDevice mc_pd;
    int &max_clusters = mc_pd.get();
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = F::CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = F::SMEM_BYTES;
    cfg.stream = a.stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (max_clusters == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, F::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        if (F::CLUSTER > 8) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (e != cudaSuccess) return e;
        }
        cfg.gridDim = dim3(F::CLUSTER * a.num_sms, 1, 1);
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
        if (e != cudaSuccess) return e;
        if (n < 1) return cudaErrorLaunchOutOfResources;
        max_clusters = n;
        if (getenv("KOFFT_CUDA_VERBOSE"))
            fprintf(stderr, "[kofft_cuda] large_fused LB=%d cluster=%d smem=%d: max active clusters %d\n", LB,
                    F::CLUSTER, F::SMEM_BYTES, n);
    }
    long nclusters = g.chunk_rows < max_clusters ? g.chunk_rows : max_clusters;
    if (g.max_clusters > 0 && nclusters > g.max_clusters) nclusters = g.max_clusters;
    if (nclusters < 1) return cudaSuccess;
    cfg.gridDim = dim3((unsigned)(nclusters * F::CLUSTER), 1, 1);
    return cudaLaunchKernelEx(&cfg, kern, io, a.tw0, a.table, g.chunk_rows, g.scratch);
}

// one persistent cooperative launch: teams of NKB CTAs, pass A of a team's next transform overlapped
// with pass B of its current one, dependency flags inside the team
template <int LB, bool EXACT, class IO, int EPI, bool STAGED>
cudaError_t launch_pipe_v(const IO &io, const LaunchArgs &a, LargeArgs &g)
{
    using F = LargePipe<LB, EXACT, IO, EPI, STAGED>;
    static_assert(F::SLOTS == kLargePipeSlots && F::FLAG_STRIDE == kPipeFlagStride, "host-side sizes");
    auto kern = large_pipe_kernel<LB, EXACT, IO, EPI, STAGED>;
    static PerDevice occ_pd;
    int &occ = occ_pd.get();
    cudaError_t e = prep(kern, F::SMEM_BYTES, 256, &occ);
    if (e != cudaSuccess) return e;
    int per_sm = occ < kMaxPipeCtasPerSm ? occ : kMaxPipeCtasPerSm;
    long cap = a.max_ctas > 0 ? a.max_ctas : (long)per_sm * a.num_sms;
    if (cap > (long)per_sm * a.num_sms) cap = (long)per_sm * a.num_sms; // every CTA must be resident
    long teams = cap / F::NKB;
    if (teams < 1) return cudaErrorLaunchOutOfResources;
    long rows = g.chunk_rows;
    if (teams > rows) teams = rows;
    if (teams > g.pipe_max_teams) teams = g.pipe_max_teams; // what the caller sized scratch and flags for
    float2 *scratch = g.scratch;
    unsigned *flags = g.flags;
    const float2 *table = a.table;
    void *args[] = {(void *)&io, (void *)&a.tw0, (void *)&table, (void *)&rows, (void *)&scratch, (void *)&flags};
    e = cudaMemsetAsync(flags, 0, sizeof(unsigned) * F::FLAG_STRIDE * teams, a.stream);
    if (e != cudaSuccess) return e;
    g.launches = 1;
    return cudaLaunchCooperativeKernel((const void *)kern, dim3((unsigned)(teams * F::NKB)), dim3(256), args, F::SMEM_BYTES,
                                       a.stream);
}

// a.staged: the host verified 16-byte alignment of the rows (asynchronous 16-byte copies)
template <int LB, bool EXACT, class IO, int EPI>
cudaError_t launch_pipe(const IO &io, const LaunchArgs &a, LargeArgs &g)
{
    if constexpr (IoTraits<IO>::kRowPtr) {
        if (a.staged) return launch_pipe_v<LB, EXACT, IO, EPI, true>(io, a, g);
    }
    return launch_pipe_v<LB, EXACT, IO, EPI, false>(io, a, g);
}

template <int LB, bool EXACT, class IO, int EPI>
cudaError_t launch_both(const IO &io, const LaunchArgs &a, LargeArgs &g)
{
    if (g.pipe) return launch_pipe<LB, EXACT, IO, EPI>(io, a, g);
    if (g.fused) {
        g.launches = 1;
        return launch_fused<LB, EXACT, IO, EPI>(io, a, g);
    }
    g.launches = 2;
    cudaError_t e = launch_col<EXACT>(io, a, g);
    if (e != cudaSuccess) return e;
    return launch_row<LB, EXACT, IO, EPI>(io, a, g);
}

template <int LB, bool EXACT>
cudaError_t launch_pair(const LaunchArgs &a, LargeArgs &g)
{
    const IoArgs &q = a.io;
    switch (a.kind) {
    case KIND_C2C_FWD: {
        IoC2C<false> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale};
        return launch_both<LB, EXACT, IoC2C<false>, ROW_STORE>(io, a, g);
    }
    case KIND_C2C_INV: {
        IoC2C<true> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale};
        return launch_both<LB, EXACT, IoC2C<true>, ROW_STORE>(io, a, g);
    }
    case KIND_GEN_FWD: {
        IoGeneric<false> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2,
                            q.p0, q.p1, q.p2, q.p3, q.scale};
        return launch_both<LB, EXACT, IoGeneric<false>, ROW_STORE>(io, a, g);
    }
    case KIND_GEN_INV: {
        IoGeneric<true> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2,
                           q.p0, q.p1, q.p2, q.p3, q.scale};
        return launch_both<LB, EXACT, IoGeneric<true>, ROW_STORE>(io, a, g);
    }
    case KIND_RFFT: {
        IoRfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n};
        return launch_both<LB, EXACT, IoRfft<EXACT>, ROW_TWIST>(io, a, g);
    }
    case KIND_IRFFT: {
        IoIrfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n, q.scale};
        return launch_both<LB, EXACT, IoIrfft<EXACT>, ROW_STORE>(io, a, g);
    }
    default:
        return cudaErrorNotSupported;
    }
}

} // namespace

cudaError_t launch_large_fft(int L, const LaunchArgs &a, LargeArgs &g)
{
    switch (L) {
    case 15: return a.exact ? launch_pair<7, true>(a, g) : launch_pair<7, false>(a, g);
    case 16: return a.exact ? launch_pair<8, true>(a, g) : launch_pair<8, false>(a, g);
    default: return cudaErrorNotSupported;
    }
}

} // namespace kofft
