// fft_inst.cu -- instantiates the single-CTA kernels for one size; compile with -DKOFFT_L=<L>.
#include "fft_kernels.cuh"
#include "launch.h"

#ifndef KOFFT_L
#error "compile with -DKOFFT_L=<log2 N>"
#endif

namespace kofft {

namespace {

template <int L, bool EXACT, class IO, bool STAGED>
cudaError_t launch_variant(const IO &io, const LaunchArgs &a)
{
    using P = Plan<L, IoTraits<IO>::kMinCta>;
    constexpr int smem = STAGED ? CtaFft<P, EXACT, IO>::SMEM_STAGED : P::SMEM_BYTES;
    auto kern = fft_cta_kernel<L, EXACT, IO, STAGED>;
    static PerDevice occ_pd; // per instantiation and device
    int &occ = occ_pd.get();
    if (occ == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        int o = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, P::CTA, smem);
        if (e != cudaSuccess) return e;
        occ = o > 0 ? o : 1;
    }
    long groups = (a.rows + P::TPC - 1) / P::TPC;
    long cap = a.max_ctas > 0 ? a.max_ctas : (long)occ * a.num_sms;
    int grid = (int)(groups < cap ? groups : cap);
    if (grid <= 0) return cudaSuccess;
    kern<<<grid, P::CTA, smem, a.stream>>>(io, a.tw0, a.table, a.rows);
    return cudaGetLastError();
}

// a.staged: the host verified alignment / contiguity for the TMA-staged variant
template <int L, bool EXACT, class IO>
cudaError_t launch_one(const IO &io, const LaunchArgs &a)
{
    if constexpr (IO::kStageable && Plan<L, IoTraits<IO>::kMinCta>::CAN_STAGE) {
        if (a.staged) return launch_variant<L, EXACT, IO, true>(io, a);
    }
    return launch_variant<L, EXACT, IO, false>(io, a);
}

template <int L, bool EXACT>
cudaError_t launch_kind(const LaunchArgs &a)
{
    const IoArgs &q = a.io;
    switch (a.kind) {
    case KIND_C2C_FWD: {
        IoC2C<false> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale};
        return launch_one<L, EXACT>(io, a);
    }
    case KIND_C2C_INV: {
        IoC2C<true> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale};
        return launch_one<L, EXACT>(io, a);
    }
    case KIND_GEN_FWD: {
        IoGeneric<false> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2,
                            q.p0, q.p1, q.p2, q.p3, q.scale};
        return launch_one<L, EXACT>(io, a);
    }
    case KIND_GEN_INV: {
        IoGeneric<true> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2,
                           q.p0, q.p1, q.p2, q.p3, q.scale};
        return launch_one<L, EXACT>(io, a);
    }
    case KIND_STFT: {
        IoStft io{(const float *)q.in, (const float *)q.aux, (float2 *)q.out, q.p0, q.p1, q.p2, q.n};
        return launch_one<L, EXACT>(io, a);
    }
    case KIND_STFT_MAG: {
        IoStftMag io{{(const float *)q.in, (const float *)q.aux, nullptr, q.p0, q.p1, q.p2, q.n}, (float *)q.out,
                     (int *)q.out2, 0.0f};
        return launch_one<L, EXACT>(io, a);
    }
    case KIND_ISTFT: {
        IoIstft io{(const float2 *)q.in, (const float *)q.aux, (float *)q.out, q.n, q.scale};
        return launch_one<L, EXACT>(io, a);
    }
    case KIND_RFFT: {
        IoRfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n};
        return launch_one<L, EXACT>(io, a);
    }
    case KIND_IRFFT: {
        IoIrfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n, q.scale};
        return launch_one<L, EXACT>(io, a);
    }
    default:
        return cudaErrorInvalidValue;
    }
}

} // namespace

#define KOFFT_CAT2(a, b) a##b
#define KOFFT_CAT(a, b) KOFFT_CAT2(a, b)

cudaError_t KOFFT_CAT(launch_cta_fft_L, KOFFT_L)(const LaunchArgs &a)
{
    return a.exact ? launch_kind<KOFFT_L, true>(a) : launch_kind<KOFFT_L, false>(a);
}

} // namespace kofft
