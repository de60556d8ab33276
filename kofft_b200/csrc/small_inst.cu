// small_inst.cu -- instantiates the N <= 16 literal kernels for every I/O policy.
#include "launch.h"
#include "small_kernels.cuh"

namespace kofft {

namespace {

template <int N, bool EXACT, class IO>
cudaError_t launch_small_one(const IO &io, const LaunchArgs &a)
{
    const int threads = 128;
    long blocks = (a.rows + threads - 1) / threads;
    long cap = (long)a.num_sms * 16;
    int grid = (int)(blocks < cap ? blocks : cap);
    if (grid <= 0) return cudaSuccess;
    fft_small_kernel<N, EXACT, IO><<<grid, threads, 0, a.stream>>>(io, a.rows);
    return cudaGetLastError();
}

template <int N, bool EXACT>
cudaError_t launch_small_kind(const LaunchArgs &a)
{
    const IoArgs &q = a.io;
    switch (a.kind) {
    case KIND_C2C_FWD: {
        IoC2C<false> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale};
        return launch_small_one<N, EXACT>(io, a);
    }
    case KIND_C2C_INV: {
        IoC2C<true> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale};
        return launch_small_one<N, EXACT>(io, a);
    }
    case KIND_GEN_FWD: {
        IoGeneric<false> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2,
                            q.p0, q.p1, q.p2, q.p3, q.scale};
        return launch_small_one<N, EXACT>(io, a);
    }
    case KIND_GEN_INV: {
        IoGeneric<true> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2,
                           q.p0, q.p1, q.p2, q.p3, q.scale};
        return launch_small_one<N, EXACT>(io, a);
    }
    case KIND_STFT: {
        IoStft io{(const float *)q.in, (const float *)q.aux, (float2 *)q.out, q.p0, q.p1, q.p2, q.n};
        return launch_small_one<N, EXACT>(io, a);
    }
    case KIND_ISTFT: {
        IoIstft io{(const float2 *)q.in, (const float *)q.aux, (float *)q.out, q.n, q.scale};
        return launch_small_one<N, EXACT>(io, a);
    }
    case KIND_RFFT: {
        IoRfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n};
        return launch_small_one<N, EXACT>(io, a);
    }
    case KIND_IRFFT: {
        IoIrfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n, q.scale};
        return launch_small_one<N, EXACT>(io, a);
    }
    default:
        return cudaErrorInvalidValue;
    }
}

template <int N>
cudaError_t launch_small_n(const LaunchArgs &a)
{
    return a.exact ? launch_small_kind<N, true>(a) : launch_small_kind<N, false>(a);
}

} // namespace

cudaError_t launch_small_fft(int n, const LaunchArgs &a)
{
    switch (n) {
    case 1: return launch_small_n<1>(a);
    case 2: return launch_small_n<2>(a);
    case 4: return launch_small_n<4>(a);
    case 8: return launch_small_n<8>(a);
    case 16: return launch_small_n<16>(a);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_cta_fft(int L, const LaunchArgs &a)
{
    switch (L) {
#define KOFFT_CASE_L(L) case L: return launch_cta_fft_L##L(a);
    KOFFT_CASE_L(5) KOFFT_CASE_L(6) KOFFT_CASE_L(7) KOFFT_CASE_L(8) KOFFT_CASE_L(9)
    KOFFT_CASE_L(10) KOFFT_CASE_L(11) KOFFT_CASE_L(12) KOFFT_CASE_L(13) KOFFT_CASE_L(14)
#undef KOFFT_CASE_L
    default: return cudaErrorInvalidValue;
    }
}

} // namespace kofft
