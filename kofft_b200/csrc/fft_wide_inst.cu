// fft_wide_inst.cu -- instantiates the wide single-CTA kernel (fft_wide.cuh) for complex cores of N = 8192 / 16384.
#include "fft_wide.cuh"
#include "launch.h"

namespace kofft {

namespace {

template <int L, bool EXACT, class IO, bool STAGED>
cudaError_t launch_wide_v(const IO &io, const LaunchArgs &a, const float2 *v0)
{
    using F = WideCta<L, EXACT, IO, STAGED>;
    auto kern = fft_wide_kernel<L, EXACT, IO, STAGED>;
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
    long cap = a.max_ctas > 0 ? a.max_ctas : (long)occ * a.num_sms;
    const int grid = (int)(a.rows < cap ? a.rows : cap);
    if (grid <= 0) return cudaSuccess;
    Tw0W tw0;
    for (int i = 0; i < 32; i++) tw0.v[i] = v0[i];
    kern<<<grid, F::CTA, F::SMEM_BYTES, a.stream>>>(io, tw0, a.table, a.rows);
    return cudaGetLastError();
}

// a.staged: the host verified 16-byte alignment of the rows (TMA bulk copies)
template <int L, bool EXACT, class IO>
cudaError_t launch_wide(const IO &io, const LaunchArgs &a, const float2 *v0)
{
    if constexpr (IoTraits<IO>::kRowPtr) {
        if (a.staged) return launch_wide_v<L, EXACT, IO, true>(io, a, v0);
    }
    return launch_wide_v<L, EXACT, IO, false>(io, a, v0);
}

template <int L, bool EXACT>
cudaError_t launch_kind(const LaunchArgs &a, const float2 *v0)
{
    const IoArgs &q = a.io;
    switch (a.kind) {
    case KIND_C2C_FWD: {
        IoC2C<false> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale};
        return launch_wide<L, EXACT>(io, a, v0);
    }
    case KIND_C2C_INV: {
        IoC2C<true> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale};
        return launch_wide<L, EXACT>(io, a, v0);
    }
    case KIND_GEN_FWD: {
        IoGeneric<false> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2,
                            q.p0, q.p1, q.p2, q.p3, q.scale};
        return launch_wide<L, EXACT>(io, a, v0);
    }
    case KIND_GEN_INV: {
        IoGeneric<true> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2,
                           q.p0, q.p1, q.p2, q.p3, q.scale};
        return launch_wide<L, EXACT>(io, a, v0);
    }
    case KIND_RFFT: {
        IoRfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n};
        return launch_wide<L, EXACT>(io, a, v0);
    }
    case KIND_IRFFT: {
        IoIrfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n, q.scale};
        return launch_wide<L, EXACT>(io, a, v0);
    }
    default:
        return cudaErrorNotSupported;
    }
}

} // namespace

cudaError_t launch_wide_fft(int L, const LaunchArgs &a, const float2 *v0)
{
    switch (L) {
    case 13: return a.exact ? launch_kind<13, true>(a, v0) : launch_kind<13, false>(a, v0);
    case 14: return a.exact ? launch_kind<14, true>(a, v0) : launch_kind<14, false>(a, v0);
    default: return cudaErrorNotSupported;
    }
}

} // namespace kofft
