// fft_split32_inst.cu -- instantiates the warp-specialised split kernel (fft_split32.cuh), N = 2^13 .. 2^15.
#include <cstdio>
#include <cstdlib>

#include <cuda.h>
#include <dlfcn.h>

#include "fft_split32.cuh"
#include "launch.h"

namespace kofft {

namespace {

// The [rows * 2^LA][32 complex] view of the input rows as a 2-D tensor of f32 (64 per row), boxes of 256 rows x
// `cols` complex.  cuTensorMapEncodeTiled is a driver entry point: resolved through the runtime, no -lcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
cudaError_t encode_rows_map(TmaMap *out, const void *base, unsigned long long total_rows, unsigned cols)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) { // the driver library is already in the process (the runtime loaded it)
        void *h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libcuda.so.1", RTLD_NOW);
        void *p = h ? dlsym(h, "cuTensorMapEncodeTiled") : nullptr;
        if (getenv("KOFFT_CUDA_VERBOSE")) fprintf(stderr, "[kofft_cuda] cuTensorMapEncodeTiled at %p\n", p);
        if (!p) return cudaErrorNotSupported;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    static_assert(sizeof(TmaMap) == sizeof(CUtensorMap), "opaque tensor map");
    const cuuint64_t dims[2] = {64, total_rows};
    const cuuint64_t strides[1] = {256};
    const cuuint32_t box[2] = {2 * cols, 256};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(reinterpret_cast<CUtensorMap *>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims,
                    strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (getenv("KOFFT_CUDA_VERBOSE")) fprintf(stderr, "[kofft_cuda] tensor map over %llu rows of 256 B: CUresult %d\n", total_rows, (int)r);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int LA, bool EXACT, class IO, int EPI, bool STAGED, bool PRE = false>
cudaError_t launch_split_v(const IO &io, const LaunchArgs &a, SplitArgs &g);

// a.staged: the host verified 16-byte alignment of the rows (TMA tile loads)
template <int LA, bool EXACT, class IO, int EPI>
cudaError_t launch_split(const IO &io, const LaunchArgs &a, SplitArgs &g)
{
    if constexpr (LA == 10 && IoTraits<IO>::kRowPtr) {
        if (a.staged && (a.rows << LA) < (1L << 31)) return launch_split_v<LA, EXACT, IO, EPI, true>(io, a, g);
    }
    return launch_split_v<LA, EXACT, IO, EPI, false>(io, a, g);
}

// one persistent cooperative launch, one 512-thread CTA per SM, teams of NT CTAs
template <int LA, bool EXACT, class IO, int EPI, bool STAGED, bool PRE>
cudaError_t launch_split_v(const IO &io, const LaunchArgs &a, SplitArgs &g)
{
    using F = Split32<LA, EXACT, IO, EPI, STAGED, PRE>;
    static_assert(F::SLOTS == kSplitSlots && F::FLAG_STRIDE == kPipeFlagStride && F::ZSLOTS == kSplitZSlots, "host-side sizes");
    auto kern = split32_kernel<LA, EXACT, IO, EPI, STAGED, PRE>;
    TmaMap map = {};
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
    if constexpr (STAGED) {
        // the tiles come from the caller's rows, or (PRE) from the teams' untwisted rows behind the intermediate
        const void *base = PRE ? static_cast<const void *>(scratch + (size_t(teams) * F::SLOTS << F::L)) : static_cast<const void *>(io.in);
        const unsigned long long trows = PRE ? (unsigned long long)(teams * F::ZSLOTS) : (unsigned long long)a.rows;
        cudaError_t em = encode_rows_map(&map, base, trows << LA, F::COLS);
        if (em != cudaSuccess) return em;
    }
    cudaError_t e = cudaMemsetAsync(flags, 0, sizeof(unsigned) * F::FLAG_STRIDE * teams, a.stream);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.numAttrs = 1;
    if (g.persist_l2) { // the intermediate as a persisting-L2 window (the caller raised cudaLimitPersistingL2CacheSize)
        attr[1].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[1].val.accessPolicyWindow.base_ptr = scratch;
        attr[1].val.accessPolicyWindow.num_bytes = sizeof(float2) * (size_t(1) << F::L) * F::SLOTS * teams;
        attr[1].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[1].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.numAttrs = 2;
    }
    cfg.gridDim = dim3((unsigned)(teams * F::NT));
    cfg.blockDim = dim3(F::CTA);
    cfg.dynamicSmemBytes = F::SMEM_BYTES;
    cfg.stream = a.stream;
    cfg.attrs = attr;
    return cudaLaunchKernelEx(&cfg, kern, io, tw0, table, rows, scratch, flags, map);
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
        if constexpr (LA == 10) { // the B warps untwist the rows ahead of pass A, which then stages tiles as for C2C
            if (g.pre_rows && (long(kSplitZSlots) * g.max_teams << LA) < (1L << 31))
                return launch_split_v<LA, EXACT, IoIrfft<EXACT>, SPLIT_STORE, true, true>(io, a, g);
        }
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
