// fft_engine.cuh -- single-CTA batched power-of-two C2C f32 FFT for N = 32 .. 16384,
// arithmetically faithful to kofft's radix-2 Stockham autosort
// (reference: src/fft.rs:789-912 fft_split_simd; stage structure :836-898).
//
// The reference runs L = log2 N radix-2 stages; stage s (n1 = 2^s groups, n2 = N/2^(s+1)):
//     u = src[2k*n2 + j]; v = src[(2k+1)*n2 + j] * T[k*n2];
//     dst[k*n2 + j] = u + v; dst[(k+n1)*n2 + j] = u - v          (k < n1, j < n2)
// where T is FftPlanner::get_twiddles(N) (src/fft.rs:391-405), an *inexact* f32 table, so a
// mathematically different factorisation would miss the 1e-5 parity budget.
//
// Here r consecutive stages s..s+r-1 are fused into one register "pass": a thread owns the
// 2^r elements  i = k*2^(L-s) + q*2^(L-s-r) + j'  (q = 0..2^r-1) of one (k, j') pair, runs r
// layers of radix-2 butterflies on them with exactly the reference's operands and table
// entries  T[(k + c_low*2^s) * 2^(L-1-s-t)]  (layer t, c_low = the t output bits already
// produced), and scatters them to  i' = (k + c*2^s)*2^(L-s-r) + j'.  Between passes the
// elements are exchanged through shared memory; data makes one trip from HBM and one back.
//
// Pass schedule: the last two passes are radix 16 (J = 16 and J = 1 elements between a
// thread's consecutive q), the first pass(es) take the remaining stages.  With the padded
// layout P(i) = i + (i >> 4) every 64-bit shared access of a half-warp hits 16 distinct
// 8-byte bank pairs (tests/test_engine_layout.py enumerates this).
#pragma once
#include "hostdev.h"

namespace kofft {

constexpr int ilog2c(unsigned v) { return v <= 1 ? 0 : 1 + ilog2c(v >> 1); }

constexpr int bitrev(int w, int bits)
{
    int r = 0;
    for (int b = 0; b < bits; b++)
        if (w & (1 << b)) r |= 1 << (bits - 1 - b);
    return r;
}

constexpr int EPT = 16; // elements per thread

// MINCTA_: smallest CTA.  Transforms needing fewer threads share a CTA (TPC > 1).  256 keeps two
// 8-warp CTAs per SM; 128 gives four 4-warp CTAs per SM (finer-grained barriers).
template <int L_, int MINCTA_ = 256>
struct Plan {
    static constexpr int L = L_;
    static constexpr int MINCTA = MINCTA_;
    static constexpr int N = 1 << L;
    static_assert(L >= 5 && L <= 14, "single-CTA engine covers N = 32 .. 16384");
    static constexpr int NP = L <= 8 ? 2 : (L <= 12 ? 3 : 4);
    static constexpr int R0 = L - 4 * (NP - 1);      // log2 radix of pass 0 (1..4)
    static constexpr int T = N / EPT;                // threads per transform
    static constexpr int CTA = T < MINCTA ? MINCTA : T; // threads per CTA
    static constexpr int TPC = CTA / T;              // transforms per CTA
    static constexpr int PADN = N + (N >> 4);        // padded float2 per exchange buffer
    // two exchange buffers (one __syncthreads per exchange) when they fit, else one
    static constexpr int NBUF = (2 * TPC * PADN * 8 <= 160 * 1024) ? 2 : 1;
    static constexpr int XCHG_BYTES = NBUF * TPC * PADN * 8;
    static constexpr int STAGE_BYTES = TPC * N * 8;  // TMA landing zone for the next row group
    static constexpr bool CAN_STAGE = XCHG_BYTES + STAGE_BYTES <= 226 * 1024;
    static constexpr int SMEM_BYTES = XCHG_BYTES;
    static constexpr int SMEM_BYTES_STAGED = XCHG_BYTES + STAGE_BYTES;
    static constexpr bool TW_REGS = L <= 12;         // later-pass twiddles live in registers

    static constexpr int r(int p) { return p == 0 ? R0 : 4; }
    static constexpr int s(int p) { return p == 0 ? 0 : R0 + 4 * (p - 1); }
    static constexpr int logJ(int p) { return L - s(p) - r(p); }
};

KHD constexpr int pad(int i) { return i + (i >> 4); }

// Twiddles of pass 0 (k = 0): thread-independent, passed in the kernel parameter block so they
// sit in the constant bank.  v[(2^t - 1) + c_low] = T[c_low << (L-1-t)].
struct Tw0 {
    float2 v[16];
};
// the same for the f64 twin (fft_f64.cuh)
struct Tw0D {
    double2 v[16];
};

// Where a sub-transform sits inside a larger transform of length 2^Lbig (large-N two-pass
// path, fft_large.cuh).  The engine's stage s of the sub-transform is stage s + sh1 of the big
// one, k0 holds the sh1 output bits produced before this sub-transform started, and sh2 is the
// number of stages that follow it:  index = (k0 + ((k + c_low*2^s) << sh1)) << (L-1-s-t + sh2).
// A stand-alone transform has {0, 0, 0}.
struct TwMap {
    int k0 = 0, sh1 = 0, sh2 = 0;
};

// index into the table for pass p, layer t, group k, produced bits c_low
template <class P>
KHD int tw_index(int p, int t, int k, int c_low, const TwMap &m = TwMap())
{
    return (m.k0 + ((k + (c_low << P::s(p))) << m.sh1)) << (P::L - 1 - P::s(p) - t + m.sh2);
}

// ------------------------------------------------------------------------------------------
// One thread's share of pass p: U = 16 >> r butterflies of radix 2^r on x[u*R + w].
// ------------------------------------------------------------------------------------------
// UNIT0: the sub-transform starts at stage 0 of the whole transform, so the group-0 twiddle of
// pass 0 is table entry 0 == (1, 0) and its other pass-0 twiddles are thread-independent (Tw0).
// REAL0 (pass 0 only, with UNIT0): the inputs are real (imag == +0): butterflies on elements that
// have met only unit twiddles so far take the real-input shortcuts of hostdev.h.
template <class P, int p, bool EXACT, bool UNIT0 = true, bool REAL0 = false>
struct Pass {
    static constexpr int r = P::r(p);
    static constexpr int R = 1 << r;
    static constexpr int U = EPT >> r;
    static constexpr int s = P::s(p);
    static constexpr int LJ = P::logJ(p);
    static constexpr int J = 1 << LJ;
    static constexpr int NTW = U * (R - 1); // twiddles per thread in this pass

    // butterfly index of sub-butterfly u of thread t
    static KHD int bfly(int t, int u) { return t + u * P::T; }

    // Element indices are split into a per-thread runtime base and a compile-time offset:
    //   read  x[u*R + q]  from  src_base(t,u) + src_off(q)
    //   write x[u*R + w]  to    dst_base(t,u) + dst_off(w)
    // By construction of the schedule (logJ is 0 or >= 4) either the offset or the base is a
    // multiple of 16, so pad(base + off) == pad(base) + pad(off) and shared-memory addresses
    // are "register + immediate".
    static KHD int src_base(int t, int u)
    {
        int b = bfly(t, u);
        int k = b >> LJ, j = b & (J - 1);
        return (k << (P::L - s)) + j;
    }
    static constexpr int src_off(int q) { return q << LJ; }
    static KHD int dst_base(int t, int u)
    {
        int b = bfly(t, u);
        int k = b >> LJ, j = b & (J - 1);
        return (k << LJ) + j;
    }
    static constexpr int dst_off(int w) { return bitrev(w, r) << (s + LJ); }
    static_assert(LJ == 0 || LJ >= 4, "schedule invariant for additive padding");
    // additive padding is valid when every offset, or every base, is a multiple of 16
    static constexpr bool SRC_ADDITIVE = LJ >= 4 || (P::L - s) >= 4;
    static constexpr bool DST_ADDITIVE = (s + LJ) >= 4;
    static KHD int src_pad(int base, int q)
    {
        return SRC_ADDITIVE ? pad(base) + pad(src_off(q)) : pad(base + src_off(q));
    }
    static KHD int dst_pad(int base, int w)
    {
        return DST_ADDITIVE ? pad(base) + pad(dst_off(w)) : pad(base + dst_off(w));
    }

    static KHD int src_index(int t, int u, int q) { return src_base(t, u) + src_off(q); }
    static KHD int dst_index(int t, int u, int w) { return dst_base(t, u) + dst_off(w); }

    // gather this thread's twiddles for the pass from the device-resident table
    static KHD void load_tw(const float2 *__restrict__ table, int t, float2 *tw /*[NTW]*/, const TwMap &m = TwMap())
    {
#pragma unroll
        for (int u = 0; u < U; u++) {
            int k = bfly(t, u) >> LJ;
#pragma unroll
            for (int tl = 0; tl < r; tl++)
#pragma unroll
                for (int c = 0; c < (1 << tl); c++)
                    tw[u * (R - 1) + (1 << tl) - 1 + c] = table[tw_index<P>(p, tl, k, c, m)];
        }
    }

    // r layers of radix-2 butterflies.  tw: per-thread twiddles (p >= 1) or Tw0::v (p == 0).
    // RE_LAST (p >= 1): the caller uses only the real parts of the results: the last layer computes only those.
    template <bool RE_LAST = false>
    static KHD void compute(float2 *x, const float2 *tw)
    {
#pragma unroll
        for (int u = 0; u < U; u++) {
#pragma unroll
            for (int tl = 0; tl < r; tl++) {
                const int bit = 1 << (r - 1 - tl);
#pragma unroll
                for (int w0 = 0; w0 < R; w0++) {
                    if (w0 & bit) continue;
                    // bits above `bit` in w0 hold the tl output bits produced so far
                    const int c_low = bitrev(w0 >> (r - tl), tl);
                    float2 &a = x[u * R + w0];
                    float2 &b = x[u * R + (w0 | bit)];
                    if (p == 0 && UNIT0) {
                        // still real: every earlier layer used the unit twiddle (its c_low was 0)
                        const bool real_in = REAL0 && (tl == 0 || (w0 >> (r - tl + 1)) == 0);
                        if (c_low == 0) {
                            if (real_in)
                                butterfly_unit_real(a, b);
                            else
                                butterfly_unit(a, b); // T[0] == (1, 0) exactly
                        } else if (real_in) {
                            butterfly_real<EXACT>(a, b, tw[(1 << tl) - 1 + c_low]);
                        } else {
                            butterfly<EXACT>(a, b, tw[(1 << tl) - 1 + c_low]);
                        }
                    } else if (RE_LAST && tl == r - 1) {
                        butterfly_re<EXACT>(a, b, tw[u * (R - 1) + (1 << tl) - 1 + c_low]);
                    } else {
                        butterfly<EXACT>(a, b, tw[u * (R - 1) + (1 << tl) - 1 + c_low]);
                    }
                }
            }
        }
    }
};

} // namespace kofft
