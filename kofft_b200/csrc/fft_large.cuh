// fft_large.cuh -- N = 2^15 .. 2^16 (and the half-length core of rfft 2^16 .. 2^17): the
// faithful two-pass split of kofft's radix-2 Stockham.
//
//   pass A ("column pass") = stages 0 .. 7: for every column j' (low L-8 index bits) a
//       256-point transform over the high 8 bits with stride 2^(L-8).  It starts at stage 0,
//       so its twiddles are the ordinary ones of a 256-point transform read from the big
//       table with stride 2^(L-8).  A CTA takes 16 adjacent columns; threads are mapped
//       column-fastest so that every global access is a 128-byte run.
//   pass B ("row pass") = stages 8 .. L-1: for every k (the 8 output bits produced by pass A)
//       a contiguous 2^(L-8)-point transform whose twiddles depend on k:
//       T[(k + (k_local + c_low 2^s) 2^8) << (L-9-s-t)], and whose bin c lands at K = k + 256 c.
//       A CTA takes 16 or 32 values of k, so after a transpose through shared memory every
//       global store is again a 128-byte run.  The inter-step twiddle of a textbook four-step
//       FFT does not exist here: it is already inside pass B's stage twiddles.
//
// The intermediate lives in a scratch buffer that the host sizes to stay L2-resident (the
// batch is processed in chunks), so HBM sees one read and one write of the data.
//
// rfft: the Hermitian twist needs bins K and m-K together.  With K = k + 256 c the mirror of
// (k, c) is (256-k, NB-1-c), so the row-pass CTA is given k-blocks paired with their mirrors
// ({0..h-1} with {128, 255..257-h}, {bh..bh+h-1} with {256-bh .. 257-bh-h}) and twists in
// shared memory before the store (src/rfft.rs:450-463).
#pragma once
#include "fft_kernels.cuh"

namespace kofft {

constexpr int LARGE_S1 = 8; // stages in pass A

// ---------------------------------------------------------------------------------------------
// pass A
// ---------------------------------------------------------------------------------------------
template <bool EXACT, class IO>
struct ColPass {
    using P = Plan<LARGE_S1>; // 256-point engine, passes (4, 4), 16 threads per column
    using P0 = Pass<P, 0, EXACT>;
    using P1 = Pass<P, 1, EXACT>;
    static constexpr int COLS = 16;            // columns per CTA tile
    static constexpr int RS = P::PADN + 1;     // odd region stride: column-fastest threads hit distinct banks
    static constexpr int SMEM_BYTES = 2 * COLS * RS * 8;
    static_assert(P::T == 16 && P::CTA == 256 && P::TPC == COLS, "tile shape");

    // n = 2^L elements per transform, lsub = L - 8, tiles = batch * (2^lsub / 16)
    // row0: index of the chunk's first transform in the caller's batch (scratch is chunk-local)
    static KD void run(const IO &io, const Tw0 &tw0, const float2 *__restrict__ table, int lsub, long tiles,
                       long row0, float2 *__restrict__ scratch, float2 *smem)
    {
        const int tid = threadIdx.x;
        const int slot = tid & (COLS - 1); // column within the tile (fastest)
        const int t = tid >> 4;            // thread within the column's 256-point transform
        float2 *buf0 = smem + slot * RS;
        float2 *buf1 = buf0 + COLS * RS;
        int par = 0;
        TwMap map;
        map.sh2 = lsub; // 256-point twiddles = every 2^lsub-th entry of the big table
        float2 tw1[P1::NTW];
        P1::load_tw(table, t, tw1, map);
        const int ltiles = lsub - 4; // log2 tiles per transform
        const long n = 1L << (LARGE_S1 + lsub);
        for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const long b = tile >> ltiles;
            const long j = ((tile & ((1L << ltiles) - 1)) << 4) + slot;
            float2 x[EPT];
#pragma unroll
            for (int q = 0; q < P0::R; q++) x[q] = io.load(row0 + b, (int)(((long)P0::src_index(t, 0, q) << lsub) + j));
            P0::compute(x, tw0.v);
            float2 *bf = par ? buf1 : buf0;
            par ^= 1;
#pragma unroll
            for (int w = 0; w < P0::R; w++) bf[P0::dst_pad(P0::dst_base(t, 0), w)] = x[w];
            __syncthreads();
#pragma unroll
            for (int q = 0; q < P1::R; q++) x[q] = bf[P1::src_pad(P1::src_base(t, 0), q)];
            P1::compute(x, tw1);
            float2 *o = scratch + b * n + j;
#pragma unroll
            for (int w = 0; w < P1::R; w++) o[(long)P1::dst_index(t, 0, w) << lsub] = x[w];
        }
    }
};

// ---------------------------------------------------------------------------------------------
// pass B
// ---------------------------------------------------------------------------------------------
enum RowEpilogue : int { ROW_STORE = 0, ROW_TWIST = 1 };

template <int LB, bool EXACT, class IO, int EPI>
struct RowPass {
    using P = Plan<LB>;
    static_assert(P::NP == 2, "row pass covers sub-lengths 128 and 256");
    using P0 = Pass<P, 0, EXACT, false>; // k != 0: no unit twiddles, per-thread pass-0 twiddles
    using P1 = Pass<P, 1, EXACT, false>;
    static constexpr int NB = P::N;
    static constexpr int TPC = P::TPC;          // sub-transforms (values of k) per CTA: 32 or 16
    static constexpr int NKB = 256 / TPC;       // k-blocks per transform
    static constexpr int RS = P::PADN + 1;      // odd region stride for the transposed read-back
    static constexpr int SMEM_BYTES = 2 * TPC * RS * 8;
    static constexpr int HALF = TPC / 2;

    // which k the CTA's slot handles in block kb
    static KHD int kmap(int kb, int slot)
    {
        if (EPI == ROW_TWIST) {
            if (slot < HALF) return HALF * kb + slot;
            int k = 256 - HALF * kb - (slot - HALF);
            return k == 256 ? 128 : k;
        }
        return kb * TPC + slot;
    }
    // slot holding the mirror 256 - k of the slot's k (twist only)
    static KHD int mirror_slot(int kb, int slot)
    {
        if (slot < HALF) return (kb == 0 && slot == 0) ? 0 : slot + HALF;
        return (kb == 0 && slot == HALF) ? HALF : slot - HALF;
    }

    // tiles = batch * NKB; the grid is a multiple of NKB so a CTA keeps its k-block (and with it
    // its twiddles, which live in registers) for every tile it processes.
    static KD void run(const IO &io, const float2 *__restrict__ table, long tiles, long row0,
                       const float2 *__restrict__ scratch, float2 *smem)
    {
        const int tid = threadIdx.x;
        const int slot = tid / P::T;
        const int t = tid - slot * P::T;
        const int kb = blockIdx.x % NKB;
        const int k = kmap(kb, slot);
        float2 *buf0 = smem + slot * RS;
        float2 *buf1 = buf0 + TPC * RS;
        const float2 *all0 = smem;
        const float2 *all1 = smem + TPC * RS;
        int par = 0;
        TwMap map;
        map.k0 = k;
        map.sh1 = LARGE_S1;
        float2 tw0[P0::NTW], tw1[P1::NTW];
        P0::load_tw(table, t, tw0, map);
        P1::load_tw(table, t, tw1, map);
        const long n = (long)NB << LARGE_S1;
        for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const long b = tile / NKB;
            const float2 *in = scratch + b * n + (long)k * NB;
            float2 x[EPT];
#pragma unroll
            for (int u = 0; u < P0::U; u++)
#pragma unroll
                for (int q = 0; q < P0::R; q++) x[u * P0::R + q] = in[P0::src_index(t, u, q)];
            P0::compute(x, tw0);
            float2 *bf = par ? buf1 : buf0;
            par ^= 1;
#pragma unroll
            for (int u = 0; u < P0::U; u++)
#pragma unroll
                for (int w = 0; w < P0::R; w++) bf[P0::dst_pad(P0::dst_base(t, u), w)] = x[u * P0::R + w];
            __syncthreads();
#pragma unroll
            for (int q = 0; q < P1::R; q++) x[q] = bf[P1::src_pad(P1::src_base(t, 0), q)];
            P1::compute(x, tw1);
            // transpose through shared memory: bins of all TPC sub-transforms, k fastest
            bf = par ? buf1 : buf0;
            const float2 *all = par ? all1 : all0;
            par ^= 1;
#pragma unroll
            for (int w = 0; w < P1::R; w++) bf[P1::dst_pad(P1::dst_base(t, 0), w)] = x[w];
            __syncthreads();
#pragma unroll
            for (int e = 0; e < EPT; e++) {
                const int flat = e * P::CTA + tid;
                const int s2 = flat % TPC, c = flat / TPC;
                const long K = kmap(kb, s2) + ((long)c << LARGE_S1);
                const float2 a = all[s2 * RS + pad(c)];
                if constexpr (EPI == ROW_TWIST) {
                    const int kk = kmap(kb, s2);
                    float2 ym;
                    if (kk == 0)
                        ym = c == 0 ? a : all[s2 * RS + pad(NB - c)]; // m - K = 256 (NB - c)
                    else
                        ym = all[mirror_slot(kb, s2) * RS + pad(NB - 1 - c)];
                    io.twist_store(row0 + b, K, a, ym);
                } else {
                    io.store(row0 + b, (int)K, a);
                }
            }
        }
    }
};

#ifdef __CUDACC__
template <bool EXACT, class IO>
__global__ void __launch_bounds__(256, 2)
    colpass_kernel(const __grid_constant__ IO io, const __grid_constant__ Tw0 tw0, const float2 *__restrict__ table,
                   int lsub, long tiles, long row0, float2 *__restrict__ scratch)
{
    extern __shared__ __align__(128) float2 smem[];
    ColPass<EXACT, IO>::run(io, tw0, table, lsub, tiles, row0, scratch, smem);
}

template <int LB, bool EXACT, class IO, int EPI>
__global__ void __launch_bounds__(256, 2)
    rowpass_kernel(const __grid_constant__ IO io, const float2 *__restrict__ table, long tiles, long row0,
                   const float2 *__restrict__ scratch)
{
    extern __shared__ __align__(128) float2 smem[];
    RowPass<LB, EXACT, IO, EPI>::run(io, table, tiles, row0, scratch, smem);
}
#endif

} // namespace kofft
