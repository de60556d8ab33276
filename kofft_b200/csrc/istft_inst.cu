// istft_inst.cu -- instantiates the fused single-kernel istft (istft_fused.cuh) for N = 512 .. 4096.
#include "istft_fused.cuh"
#include "launch.h"

namespace kofft {

namespace {

template <int L, bool EXACT>
cudaError_t launch_one(const IstftFusedArgs &f, const LaunchArgs &a)
{
    using K = IstftFused<L, EXACT>;
    auto kern = istft_fused_kernel<L, EXACT>;
    static PerDevice occ_pd;
    int &occ = occ_pd.get();
    const int smem = K::smem_bytes(f.hop);
    static PerDevice smem_pd; // largest dynamic shared memory size opted into so far
    int &smem_set = smem_pd.get();
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        smem_set = smem;
    }
    {
        int o = 0; // occupancy depends on hop through the shared-memory size
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, K::CTA, smem);
        if (e != cudaSuccess) return e;
        occ = o > 0 ? o : 1;
    }
    long nfe = (f.out_len + f.hop - 1) / f.hop;
    if (nfe > f.nframes) nfe = f.nframes;
    const long runs = f.channels * ((nfe + f.run_frames - 1) / f.run_frames);
    long cap = a.max_ctas > 0 ? a.max_ctas : (long)occ * a.num_sms;
    int grid = (int)(runs < cap ? runs : cap);
    if (grid <= 0) return cudaSuccess;
    kern<<<grid, K::CTA, smem, a.stream>>>(f, a.tw0, a.table);
    return cudaGetLastError();
}

} // namespace

cudaError_t launch_istft_fused(int L, const LaunchArgs &a, const IstftFusedArgs &f)
{
    switch (L) {
    case 9: return a.exact ? launch_one<9, true>(f, a) : launch_one<9, false>(f, a);
    case 10: return a.exact ? launch_one<10, true>(f, a) : launch_one<10, false>(f, a);
    case 11: return a.exact ? launch_one<11, true>(f, a) : launch_one<11, false>(f, a);
    case 12: return a.exact ? launch_one<12, true>(f, a) : launch_one<12, false>(f, a);
    default: return cudaErrorNotSupported;
    }
}

} // namespace kofft
