// fft_split32_inst.cu -- instantiates the warp-specialised split kernel (fft_split32.cuh), N = 2^13 .. 2^15.
#include <cstdio>
#include <cstdlib>

#include "fft_split32.cuh"
#include "launch.h"

namespace kofft {

namespace {

// one persistent cooperative launch, one 512-thread CTA per SM, teams of NT CTAs
template <int LA, bool EXACT, class IO, int EPI>
cudaError_t launch_split(const IO &io, const LaunchArgs &a, SplitArgs &g)
{
    using F = Split32<LA, EXACT, IO, EPI>;
    static_assert(F::SLOTS == kSplitSlots && F::FLAG_STRIDE == kPipeFlagStride, "host-side sizes");
    auto kern = split32_kernel<LA, EXACT, IO, EPI>;
    static PerDevice occ_pd;
    int &occ = occ_pd.get();
    if (occ == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, F::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        int o = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, F::CTA, F::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        if (o < 1) return cudaErrorLaunchOutOfResources;
        occ = o;
    }
    long cap = a.num_sms; // one CTA per SM: the roles split the SM's registers between them
    if (a.max_ctas > 0 && a.max_ctas < cap) cap = a.max_ctas;
    long teams = cap / F::NT;
    if (teams < 1) return cudaErrorLaunchOutOfResources;
    long rows = a.rows;
    if (teams > rows) teams = rows;
    if (teams > g.max_teams) teams = g.max_teams; // what the caller sized scratch and flags for
    Tw0W tw0;
    for (int i = 0; i < 32; i++) tw0.v[i] = g.v0[i];
    float2 *scratch = g.scratch;
    unsigned *flags = g.flags;
    const float2 *table = a.table;
    void *args[] = {(void *)&io, (void *)&tw0, (void *)&table, (void *)&rows, (void *)&scratch, (void *)&flags};
    cudaError_t e = cudaMemsetAsync(flags, 0, sizeof(unsigned) * F::FLAG_STRIDE * teams, a.stream);
    if (e != cudaSuccess) return e;
    return cudaLaunchCooperativeKernel((const void *)kern, dim3((unsigned)(teams * F::NT)), dim3(F::CTA), args,
                                       F::SMEM_BYTES, a.stream);
}

template <int LA, bool EXACT>
cudaError_t launch_kind(const LaunchArgs &a, SplitArgs &g)
{
    const IoArgs &q = a.io;
    switch (a.kind) {
    case KIND_C2C_FWD: {
        IoC2C<false> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale};
        return launch_split<LA, EXACT, IoC2C<false>, SPLIT_STORE>(io, a, g);
    }
    case KIND_C2C_INV: {
        IoC2C<true> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale};
        return launch_split<LA, EXACT, IoC2C<true>, SPLIT_STORE>(io, a, g);
    }
    case KIND_GEN_FWD: {
        IoGeneric<false> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2,
                            q.p0, q.p1, q.p2, q.p3, q.scale};
        return launch_split<LA, EXACT, IoGeneric<false>, SPLIT_STORE>(io, a, g);
    }
    case KIND_GEN_INV: {
        IoGeneric<true> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2,
                           q.p0, q.p1, q.p2, q.p3, q.scale};
        return launch_split<LA, EXACT, IoGeneric<true>, SPLIT_STORE>(io, a, g);
    }
    case KIND_RFFT: {
        IoRfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n};
        return launch_split<LA, EXACT, IoRfft<EXACT>, SPLIT_TWIST>(io, a, g);
    }
    case KIND_IRFFT: {
        IoIrfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n, q.scale};
        return launch_split<LA, EXACT, IoIrfft<EXACT>, SPLIT_STORE>(io, a, g);
    }
    default:
        return cudaErrorNotSupported;
    }
}

} // namespace

cudaError_t launch_split32_fft(int L, const LaunchArgs &a, SplitArgs &g)
{
    switch (L) {
#ifndef KOFFT_SPLIT_ONLY_15
    case 13: return a.exact ? launch_kind<8, true>(a, g) : launch_kind<8, false>(a, g);
    case 14: return a.exact ? launch_kind<9, true>(a, g) : launch_kind<9, false>(a, g);
#endif
    case 15: return a.exact ? launch_kind<10, true>(a, g) : launch_kind<10, false>(a, g);
    default: return cudaErrorNotSupported;
    }
}

} // namespace kofft
