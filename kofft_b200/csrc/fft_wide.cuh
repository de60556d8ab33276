// fft_wide.cuh -- N = 8192 / 16384 complex in one CTA with 32 elements per thread: the same faithful radix-2
// Stockham stages (src/fft.rs:789-912) as three fused register passes (5 + 5 + (L - 10) stages) with two
// shared-memory exchanges in ONE buffer, so that N = 8192 (64 KB + padding) leaves room for TWO independent CTAs
// per SM: while one CTA sits at a barrier or waits for its row, the other computes.  (The 16-elements-per-thread
// engine in fft_kernels.cuh needs a stage plus two exchange buffers at this size: one CTA per SM, whose barrier
// phases nothing fills -- 45 % of the HBM roofline; the warp-specialised split kernel is at 54 %.)
//
//   pass 0 (stages 0..4)   : thread t loads x[t + q N/32] (coalesced), group k = 0: thread-independent twiddles
//   exchange 1             : layout pad32 (conflict-free for pass 1's stride-2^(L-10) reads)
//   pass 1 (stages 5..9)   : twiddles per k = b >> (L-10) (32 values of k) from a shared-memory table
//   exchange 2             : same buffer, layout i + (i >> (L-10)) (conflict-free for pass 2's stride-2^(L-10) reads)
//   pass 2 (stages 10..L-1): 32 >> (L-10) butterflies of radix 2^(L-10) per thread, twiddles from the L1-resident
//                            device table; thread t stores bins t + u N/32 + c 1024: coalesced
#pragma once
#include "fft_split32.cuh"

#ifndef KOFFT_WIDE_UNROLL
#define KOFFT_WIDE_UNROLL 1
#endif

namespace kofft {

template <class IO> struct IsIrfftIo { static constexpr bool value = false; };
template <bool E> struct IsIrfftIo<IoIrfft<E>> { static constexpr bool value = true; };

// STAGED: the row arrives by TMA bulk copies in the exchange buffer itself while the previous row's last pass runs; pass 0
// is then in place in index space (thread t reads and writes positions t + q N/32).  At N = 8192 pass 1 reads 8 adjacent
// elements per group k, so two k would share banks: output c of thread t is stored in the slot of thread t ^ 8 when c is
// odd (a warp-level swap: __syncwarp between the warp's reads and writes), and pass 1 reads q ^ (k & 1).
template <int L, bool EXACT, class IO, bool STAGED = false>
struct WideCta {
    static_assert(L == 13 || L == 14, "N = 8192, 16384");
    static constexpr int N = 1 << L, CTA = N / WIDE;
    static constexpr int R2 = L - 10;
    using P0 = WidePass<L, 0, 5, EXACT, true>;
    using P1 = WidePass<L, 5, 5, EXACT, false>;
    using P2 = WidePass<L, 10, R2, EXACT, false>;
    static_assert(P0::U == 1 && P1::U == 1 && P2::LJ == 0, "pass shapes");
    static constexpr int MIN_BLOCKS = L == 13 ? 2 : 1;
    KHD static constexpr int pad_a(int i) { return i + (i >> 5); }
    // second exchange: an XOR swizzle instead of padding (pass 1 stores 16 adjacent elements, pass 2 reads elements
    // 2^R2 apart: both hit 16 distinct 8-byte banks), so the buffer is exactly one row
    KHD static constexpr int pad_b(int i) { return i ^ ((i >> 4) & ((1 << R2) - 1)); }
    static constexpr int BUF = N + (N >> 5) + 8;       // float2; pad_a (plain loads) is the larger layout
    static constexpr int TW1 = 32 * 33;                // [k][33]: 31 twiddles per k, rows on different banks
    // rfft: the Hermitian twist pairs bin K with bin N - K, which another thread holds.  Every thread holds 16 bins of the
    // lower half (c < R/2 in K = b + 1024 c) and the 16 mirrors of other threads' lower bins: the upper bins go through a
    // side buffer of N/2 elements, then each thread twists both bins of its 16 pairs (src/rfft.rs:450-463).  The main buffer
    // is free as soon as pass 2 has read its inputs, so the next row's copy overlaps the last pass and the epilogue.
    static constexpr bool TWIST = IO::kEpilogueExchange;
    static constexpr int SIDE = TWIST ? N / 2 : 0;
    static constexpr int SMEM_BYTES = (BUF + TW1 + SIDE + 2) * 8; // + the mbarrier of the staged row
    static constexpr int ROW_UNROLL = KOFFT_WIDE_UNROLL; // copies of the row loop body
    static constexpr bool SWAP = P1::LJ == 3;
    // irfft: the rows of N + 1 bins are only 8-byte aligned and the untwist (src/rfft.rs:485-498) pairs bin e with bin N - e,
    // which another thread loads: the raw row is staged in the buffer by 8-byte asynchronous copies (issued, like the bulk
    // copies of the STAGED variant, as soon as pass 2 has read its inputs), every thread untwists its 32 elements from the
    // staged bins, and pass 0 then runs in place as in the STAGED variant
    static constexpr bool UNTW = IsIrfftIo<IO>::value;
    static constexpr unsigned ROW_BYTES = N * 8u, PIECE = 16384u;

    static KD void issue_row(float2 *buf, const float2 *src, unsigned long long *bar)
    {
        mbar_expect_tx(bar, ROW_BYTES);
#pragma unroll
        for (unsigned o = 0; o < ROW_BYTES; o += PIECE)
            bulk_copy_g2s(reinterpret_cast<unsigned char *>(buf) + o, reinterpret_cast<const unsigned char *>(src) + o, PIECE, bar);
    }

    static KD void run(const IO &io, const Tw0W &tw0, const float2 *__restrict__ table, long rows, float2 *smem)
    {
        const int t = threadIdx.x;
        float2 *buf = smem;
        float2 *tw1s = smem + BUF;
        // pass-1 twiddles: entry e = (2^tl - 1) + c of group k is T[(k + (c << 5)) << (L - 6 - tl)]
        for (int i = t; i < 32 * 31; i += CTA) {
            const int k = i / 31, e = i - k * 31;
            int tl = 0;
            while ((2 << tl) - 1 <= e) tl++;
            const int c = e + 1 - (1 << tl);
            tw1s[k * 33 + e] = table[(long)(k + (c << 5)) << (L - 6 - tl)];
        }
        const float2 *tw1 = tw1s + (t >> P1::LJ) * 33;
        float2 *side = smem + BUF + TW1;
        unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + BUF + TW1 + SIDE);
        unsigned phase = 0;
        if constexpr (STAGED && !UNTW) {
            if (t == 0) {
                mbar_init(bar, 1);
                fence_mbar_init();
            }
        }
        __syncthreads();
        if constexpr (STAGED && !UNTW) {
            if (t == 0 && (long)blockIdx.x < rows) issue_row(buf, io.row_ptr(blockIdx.x), bar);
        }
        auto fetch_bins = [&](long r) {
            if constexpr (UNTW) {
                const float2 *X = io.in + r * (N + 1);
#pragma unroll
                for (int q = 0; q < 32; q++) cp_async8(buf + t + q * CTA, X + t + q * CTA);
                if (t == 0) cp_async8(buf + N, X + N);
                cp_async_commit();
            }
        };
        if (UNTW && (long)blockIdx.x < rows) fetch_bins(blockIdx.x);
        // STAGED, pass 1: element q of group k = t >> LJ sits at k N/32 + j + ((q ^ (k & 1)) << LJ) when SWAP
        const int k1 = t >> P1::LJ, sb = SWAP ? (k1 & 1) : 0;
        const float2 *s1e = buf + (k1 << (L - 5)) + (t & (P1::J - 1)) + (sb << P1::LJ);
        const float2 *s1o = buf + (k1 << (L - 5)) + (t & (P1::J - 1)) - (sb << P1::LJ);
#pragma unroll(ROW_UNROLL)
        for (long row = blockIdx.x; row < rows; row += gridDim.x) {
            float2 x[WIDE];
            if constexpr (STAGED || UNTW) {
                if constexpr (UNTW) {
                    cp_async_wait_all();
                    __syncthreads(); // every thread's bins have landed
#pragma unroll
                    for (int q = 0; q < 32; q++) {
                        const int e = t + q * CTA;
                        const float2 a = buf[e], xm = buf[N - e];
                        const float2 v = e == 0 ? io.untwist0(a, xm) : io.untwist(a, xm, KOFFT_LDG(io.rtw + e));
                        x[q] = io.from_raw(v);
                    }
                    __syncthreads(); // the raw bins are consumed: pass 0 may write
                } else {
                    mbar_wait(bar, phase);
                    phase ^= 1;
#pragma unroll
                    for (int q = 0; q < 32; q++) x[q] = io.from_raw(buf[t + q * CTA]);
                }
                P0::compute(x, tw0.v);
                if (SWAP && !UNTW) warp_sync(); // the warp has read its positions: they may be rewritten
#pragma unroll
                for (int w = 0; w < 32; w++) {
                    const int c = bitrev(w, 5);
                    buf[(SWAP && (c & 1) ? (t ^ 8) : t) + c * CTA] = x[w];
                }
                __syncthreads();
#pragma unroll
                for (int q = 0; q < 32; q++) x[q] = ((q & 1) ? s1o : s1e)[q << P1::LJ];
            } else {
#pragma unroll
                for (int q = 0; q < 32; q++) x[q] = io.load(row, P0::src_index(t, 0, q));
                P0::compute(x, tw0.v);
                __syncthreads(); // the previous row's pass-2 reads of the buffer are complete
#pragma unroll
                for (int w = 0; w < 32; w++) buf[pad_a(P0::dst_index(t, 0, w))] = x[w];
                __syncthreads();
#pragma unroll
                for (int q = 0; q < 32; q++) x[q] = buf[pad_a(P1::src_index(t, 0, q))];
            }
            __syncthreads(); // everyone has read: the buffer takes the second layout
            P1::compute(x, tw1);
#pragma unroll
            for (int w = 0; w < 32; w++) buf[pad_b(P1::dst_index(t, 0, w))] = x[w];
            // pass-2 twiddles of this thread's butterflies k = t + u CTA: T[(k + (c << 10)) << (L - 11 - tl)]
            float2 tw2[P2::NTW];
#pragma unroll
            for (int u = 0; u < P2::U; u++)
#pragma unroll
                for (int tl = 0; tl < R2; tl++)
#pragma unroll
                    for (int c = 0; c < (1 << tl); c++)
                        tw2[u * (P2::R - 1) + (1 << tl) - 1 + c] =
                            KOFFT_LDG(table + ((long)(P2::bfly(t, u) + (c << 10)) << (L - 11 - tl)));
            __syncthreads();
#pragma unroll
            for (int u = 0; u < P2::U; u++)
#pragma unroll
                for (int q = 0; q < P2::R; q++) x[u * P2::R + q] = buf[pad_b(P2::src_index(t, u, q))];
            if constexpr (UNTW) {
                __syncthreads(); // everyone has read: the next row's bins may land while the last pass runs
                if (row + gridDim.x < rows) fetch_bins(row + gridDim.x);
            } else if constexpr (STAGED) {
                __syncthreads(); // everyone has read: the next row may land while the last pass runs
                if (t == 0 && row + gridDim.x < rows) {
                    fence_proxy_async();
                    issue_row(buf, io.row_ptr(row + gridDim.x), bar);
                }
            }
            P2::compute(x, tw2);
            if constexpr (TWIST) {
                constexpr int HR = P2::R / 2;
#pragma unroll
                for (int u = 0; u < P2::U; u++)
#pragma unroll
                    for (int w = 0; w < P2::R; w++) {
                        const int c = bitrev(w, R2);
                        if (c >= HR) side[(c - HR) * 1024 + P2::bfly(t, u)] = x[u * P2::R + w];
                    }
                __syncthreads();
#pragma unroll
                for (int u = 0; u < P2::U; u++) {
                    const int b = P2::bfly(t, u);
#pragma unroll
                    for (int w = 0; w < P2::R; w++) {
                        const int c = bitrev(w, R2);
                        if (c >= HR) continue;
                        const float2 y = x[u * P2::R + w];
                        const int K = b + 1024 * c;
                        if (b == 0 && c == 0) { // bins 0 and N from Y[0]; bin N/2 is its own mirror
                            io.twist_store(row, 0, y, y);
                            const float2 h = side[0];
                            io.twist_store(row, N / 2, h, h);
                        } else {
                            const float2 ym = b == 0 ? side[(HR - c) * 1024] : side[(HR - 1 - c) * 1024 + (1024 - b)];
                            io.twist_store(row, K, y, ym);
                            io.twist_store(row, N - K, ym, y);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < P2::U; u++)
#pragma unroll
                    for (int w = 0; w < P2::R; w++) io.store(row, P2::dst_index(t, u, w), x[u * P2::R + w]);
            }
        }
    }
};

#ifdef __CUDACC__
template <int L, bool EXACT, class IO, bool STAGED>
__global__ void __launch_bounds__((WideCta<L, EXACT, IO>::CTA), (WideCta<L, EXACT, IO>::MIN_BLOCKS))
    fft_wide_kernel(const __grid_constant__ IO io, const __grid_constant__ Tw0W tw0, const float2 *__restrict__ table, long rows)
{
    extern __shared__ __align__(128) float2 smem[];
    WideCta<L, EXACT, IO, STAGED>::run(io, tw0, table, rows, smem);
}
#endif

} // namespace kofft
