// fft_f64_inst.cu -- instantiates the f64 kernels (fft_f64.cuh; N <= 16 through small_kernels.cuh).
#include "fft_f64.cuh"
#include "launch.h"
#include "small_kernels.cuh"

namespace kofft {

namespace {

template <int L, class IO, bool STAGED>
cudaError_t launch_f64_L_v(const IO &io, const LaunchF64Args &a)
{
    using P = PlanD<L>;
    auto kern = fft_f64_kernel<L, IO, STAGED>;
    static PerDevice occ_pd;
    int &occ = occ_pd.get();
    if (occ == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        int o = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, P::CTA, P::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        occ = o > 0 ? o : 1;
    }
    const long groups = (a.rows + P::TPC - 1) / P::TPC;
    long cap = a.max_ctas > 0 ? a.max_ctas : (long)occ * a.num_sms;
    int grid = (int)(groups < cap ? groups : cap);
    if (grid <= 0) return cudaSuccess;
    kern<<<grid, P::CTA, P::SMEM_BYTES, a.stream>>>(io, a.tw0, a.table, a.rows);
    return cudaGetLastError();
}

// a.staged: TMA input prefetch (the rows are 16-byte aligned, which double2 rows always are; the switch
// exists for A/B measurements, kofft_cuda_set_tma_staging)
template <int L, class IO>
cudaError_t launch_f64_L(const IO &io, const LaunchF64Args &a)
{
    if constexpr (IO::kStageable) {
        if (a.staged) return launch_f64_L_v<L, IO, true>(io, a);
    }
    return launch_f64_L_v<L, IO, false>(io, a);
}

template <int N, class IO>
cudaError_t launch_f64_small(const IO &io, const LaunchF64Args &a)
{
    const int threads = 128;
    long blocks = (a.rows + threads - 1) / threads;
    long cap = (long)a.num_sms * 16;
    int grid = (int)(blocks < cap ? blocks : cap);
    if (grid <= 0) return cudaSuccess;
    fft_small_kernel<N, true, IO><<<grid, threads, 0, a.stream>>>(io, a.rows);
    return cudaGetLastError();
}

template <class IO>
cudaError_t launch_f64_io(const IO &io, const LaunchF64Args &a)
{
    switch (a.n) {
    case 1: return launch_f64_small<1>(io, a);
    case 2: return launch_f64_small<2>(io, a);
    case 4: return launch_f64_small<4>(io, a);
    case 8: return launch_f64_small<8>(io, a);
    case 16: return launch_f64_small<16>(io, a);
    case 32: return launch_f64_L<5>(io, a);
    case 64: return launch_f64_L<6>(io, a);
    case 128: return launch_f64_L<7>(io, a);
    case 256: return launch_f64_L<8>(io, a);
    case 512: return launch_f64_L<9>(io, a);
    case 1024: return launch_f64_L<10>(io, a);
    case 2048: return launch_f64_L<11>(io, a);
    case 4096: return launch_f64_L<12>(io, a);
    case 8192: return launch_f64_L<13>(io, a);
    default: return cudaErrorNotSupported;
    }
}

} // namespace

cudaError_t launch_fft_f64(const LaunchF64Args &a)
{
    if (a.generic) { // strided / split rows
        if (a.inverse) {
            IoGenericD<true> io{a.in_re, a.in_im, a.out_re, a.out_im, a.in_es, a.in_rs, a.out_es, a.out_rs, a.scale};
            return launch_f64_io(io, a);
        }
        IoGenericD<false> io{a.in_re, a.in_im, a.out_re, a.out_im, a.in_es, a.in_rs, a.out_es, a.out_rs, a.scale};
        return launch_f64_io(io, a);
    }
    if (a.real == 1) { // rfft: n = m
        IoRfftD io{a.in, a.out, a.rtw, a.n};
        return launch_f64_io(io, a);
    }
    if (a.real == 2) { // irfft
        IoIrfftD io{a.in, a.out, a.rtw, a.n, a.scale};
        return launch_f64_io(io, a);
    }
    if (a.inverse) {
        IoC2CD<true> io{a.in, a.out, a.n, a.scale};
        return launch_f64_io(io, a);
    }
    IoC2CD<false> io{a.in, a.out, a.n, a.scale};
    return launch_f64_io(io, a);
}

} // namespace kofft
