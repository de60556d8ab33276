// fft_f64.cuh -- the f64 twin of the single-CTA engine: ScalarFftImpl<f64>::fft / ifft for
// power-of-two N = 32 .. 8192 (reference: src/fft.rs:914-1051 fft_split_simd for f64; dispatch
// :642-706 and :1054-1082; ifft :1134-1174; table :391-405 in f64).
//
// Same radix-2 Stockham stage structure, same fused register passes and the same index algebra
// as the f32 engine (Plan / Pass in fft_engine.cuh are reused for every address), with double2
// elements, unfused f64 arithmetic and the planner's f64 table.  What differs from the f32
// kernel is sized by the element: 16 elements per thread are 64 registers, so the later-pass
// twiddles are read from the (L1/L2-resident) table per pass instead of living in registers,
// and the exchange goes through ONE padded shared-memory buffer (two barriers per exchange).
// STAGED: that buffer is idle from the last exchange of a row group until the first exchange of
// the next one, so the next group's rows are fetched into it by one TMA bulk copy while the last
// register pass runs and the results are stored -- input prefetch without a byte of extra shared
// memory (two CTAs of N = 4096 per SM either way).
// An HBM-bound kernel like its f32 twin (32 B per point), measured in profiles/.
#pragma once
#include "fft_kernels.cuh"

namespace kofft {

template <int L_>
struct PlanD : Plan<L_, 64> {
    using Base = Plan<L_, 64>;
    static_assert(L_ >= 5 && L_ <= 13, "f64 single-CTA engine covers N = 32 .. 8192");
    static constexpr int SMEM_BYTES = Base::TPC * Base::PADN * 16; // one exchange buffer
};

// Pass<P, p> supplies the index functions; compute / load_tw are restated for double2
template <class P, int p>
struct PassD : Pass<P, p, true> {
    using B = Pass<P, p, true>;
    static KHD void load_tw(const double2 *__restrict__ table, int t, double2 *tw /*[NTW]*/)
    {
#pragma unroll
        for (int u = 0; u < B::U; u++) {
            int k = B::bfly(t, u) >> B::LJ;
#pragma unroll
            for (int tl = 0; tl < B::r; tl++)
#pragma unroll
                for (int c = 0; c < (1 << tl); c++)
                    tw[u * (B::R - 1) + (1 << tl) - 1 + c] = table[tw_index<P>(p, tl, k, c)];
        }
    }
    // r layers of radix-2 butterflies; tw: per-thread twiddles (p >= 1) or Tw0D::v (p == 0)
    static KHD void compute(double2 *x, const double2 *tw)
    {
#pragma unroll
        for (int u = 0; u < B::U; u++) {
#pragma unroll
            for (int tl = 0; tl < B::r; tl++) {
                const int bit = 1 << (B::r - 1 - tl);
#pragma unroll
                for (int w0 = 0; w0 < B::R; w0++) {
                    if (w0 & bit) continue;
                    const int c_low = bitrev(w0 >> (B::r - tl), tl);
                    double2 &a = x[u * B::R + w0];
                    double2 &b = x[u * B::R + (w0 | bit)];
                    if (p == 0) {
                        if (c_low == 0)
                            butterfly_unit_f64(a, b); // T[0] == (1, 0) exactly
                        else
                            butterfly_f64(a, b, tw[(1 << tl) - 1 + c_low]);
                    } else {
                        butterfly_f64(a, b, tw[u * (B::R - 1) + (1 << tl) - 1 + c_low]);
                    }
                }
            }
        }
    }
};

// FftImpl<f64>::fft / ifft on contiguous rows; ifft = conj -> fft -> conj, * 1/n (src/fft.rs:1163-1172)
template <bool INV>
struct IoC2CD {
    typedef void is_f64;
    static constexpr bool kEpilogueExchange = false;
    static constexpr bool kStageable = true; // dense 16-byte aligned rows: TMA prefetch possible
    const double2 *__restrict__ in;
    double2 *__restrict__ out;
    long n;
    double scale; // 1/n, computed on the host as 1.0 / (double)(float)n
    KHD double2 from_raw(double2 v) const
    {
        if (INV) v.y = -v.y;
        return v;
    }
    KHD double2 load(long row, int i) const { return from_raw(in[row * n + i]); }
    KHD void store(long row, int i, double2 v) const
    {
        if (INV) {
            v.y = -v.y;
            v.x = dmul(v.x, scale);
            v.y = dmul(v.y, scale);
        }
        out[row * n + i] = v;
    }
};

// strided and split (SoA) rows, FftImpl<f64>::fft_strided / fft_out_of_place_strided / fft_split / ifft_split
// (src/fft.rs:1175-1336, 556-586 -> 1365-1439): element e of row r lives at re[r*row_stride + e*elem_stride]
// (strides in doubles); interleaved data passes im = re + 1 and elem_stride = 2 * stride
template <bool INV>
struct IoGenericD {
    typedef void is_f64;
    static constexpr bool kEpilogueExchange = false;
    static constexpr bool kStageable = false;
    const double *__restrict__ in_re;
    const double *__restrict__ in_im;
    double *__restrict__ out_re;
    double *__restrict__ out_im;
    long in_es, in_rs, out_es, out_rs;
    double scale;
    KHD double2 load(long row, int i) const
    {
        const long o = row * in_rs + (long)i * in_es;
        return make_double2(in_re[o], INV ? -in_im[o] : in_im[o]);
    }
    KHD void store(long row, int i, double2 v) const
    {
        if (INV) {
            v.y = -v.y;
            v.x = dmul(v.x, scale);
            v.y = dmul(v.y, scale);
        }
        const long o = row * out_rs + (long)i * out_es;
        out_re[o] = v.x;
        out_im[o] = v.y;
    }
};

// rfft for f64 (rfft_direct, src/rfft.rs:425-465 with T = f64): the row of 2m reals is read as m complex
// (the pack is a reinterpretation); after the length-m FFT the bins go once more through the exchange
// buffer and are twisted.  T'[k] = exp(-i pi k / m) from build_twiddle_table::<f64> (:172-183).
struct IoRfftD {
    typedef void is_f64;
    static constexpr bool kEpilogueExchange = true;
    static constexpr bool kStageable = true;
    const double2 *__restrict__ in;  // [rows][m]
    double2 *__restrict__ out;       // [rows][m + 1]
    const double2 *__restrict__ rtw; // m entries
    long n;                          // m (the engine's transform length; `n` so that staging is generic)
    KHD double2 from_raw(double2 v) const { return v; }
    KHD double2 load(long row, int i) const { return in[row * n + i]; }
    KHD void store(long, int, double2) const {}
    // Y: the m FFT bins of this row, padded layout (or a plain array for m <= 16, where pad(k) == k)
    KHD void epilogue(long row, int k, const double2 *Y) const
    {
        const long m = n;
        double2 *o = out + row * (m + 1);
        const double2 a = Y[pad(k)];
        if (k == 0) { // :451-453
            o[0] = make_double2(dadd(a.x, a.y), 0.0);
            o[m] = make_double2(dsub(a.x, a.y), 0.0);
            return;
        }
        const double2 ym = Y[pad((int)m - k)];
        const double2 b = make_double2(ym.x, -ym.y);
        const double2 sum = add2(a, b), diff = sub2(a, b);
        const double2 t = cmul<true>(rtw[k], diff);
        const double2 temp = make_double2(dadd(sum.x, t.y), dsub(sum.y, t.x)); // sum + (t.im, -t.re)
        o[k] = make_double2(dmul(temp.x, 0.5), dmul(temp.y, 0.5));
    }
};

// irfft for f64 (irfft_direct, src/rfft.rs:468-508): the untwist is evaluated while loading element i
// (it needs X[i] and X[m-i]), then ifft (conj, fft, conj, * 1/m), then the unpack is a reinterpretation
struct IoIrfftD {
    typedef void is_f64;
    static constexpr bool kEpilogueExchange = false;
    static constexpr bool kStageable = false;
    const double2 *__restrict__ in;  // [rows][m + 1]
    double2 *__restrict__ out;       // [rows][m] == 2m reals
    const double2 *__restrict__ rtw;
    long n;                          // m
    double scale;                    // 1/m
    KHD double2 load(long row, int i) const
    {
        const long m = n;
        const double2 *X = in + row * (m + 1);
        double2 v;
        if (i == 0) {
            const double2 x0 = X[0], xm = X[m];
            v = make_double2(dmul(dadd(x0.x, xm.x), 0.5), dmul(dsub(x0.x, xm.x), 0.5));
        } else {
            const double2 a = X[i], xm = X[m - i];
            const double2 b = make_double2(xm.x, -xm.y);
            const double2 sum = add2(a, b), diff = sub2(a, b);
            const double2 tw = rtw[i];
            const double2 t = cmul<true>(make_double2(tw.x, -tw.y), diff);
            const double2 temp = make_double2(dsub(sum.x, t.y), dadd(sum.y, t.x)); // sum - (t.im, -t.re)
            v = make_double2(dmul(temp.x, 0.5), dmul(temp.y, 0.5));
        }
        v.y = -v.y; // ifft's leading conjugation
        return v;
    }
    KHD void store(long row, int i, double2 v) const
    {
        v.y = -v.y;
        out[row * n + i] = make_double2(dmul(v.x, scale), dmul(v.y, scale));
    }
};

template <int L, class IO>
struct CtaFftD {
    using P = PlanD<L>;
    using P0 = PassD<P, 0>;
    using P1 = PassD<P, 1>;
    using P2 = PassD<P, (P::NP > 2 ? 2 : 1)>;
    using P3 = PassD<P, (P::NP > 3 ? 3 : 1)>;

#if defined(__CUDACC__) || defined(KOFFT_EMU)
    template <class PA, class PB>
    static KD void xchg(double2 *buf, int t, double2 *x)
    {
#pragma unroll
        for (int u = 0; u < PA::U; u++)
#pragma unroll
            for (int w = 0; w < PA::R; w++) buf[PA::dst_pad(PA::dst_base(t, u), w)] = x[u * PA::R + w];
        __syncthreads();
#pragma unroll
        for (int u = 0; u < PB::U; u++)
#pragma unroll
            for (int q = 0; q < PB::R; q++) x[u * PB::R + q] = buf[PB::src_pad(PB::src_base(t, u), q)];
        __syncthreads(); // one buffer: everyone has read before the next exchange overwrites it
    }

    template <bool STAGED>
    static KD void run(const IO &io, const Tw0D &tw0, const double2 *__restrict__ table, long rows, double2 *smem)
    {
        const int tid = threadIdx.x;
        const int slot = tid / P::T;
        const int t = tid - slot * P::T;
        double2 *buf = smem + slot * P::PADN;
        const long groups = (rows + P::TPC - 1) / P::TPC;
        __shared__ __align__(8) unsigned long long mbar;
        unsigned phase = 0;
        // the rows of group g are one contiguous byte range; it lands unpadded at the start of the buffer
        auto stage_issue = [&](long g) {
            if constexpr (STAGED) {
                long nr = rows - g * P::TPC;
                if (nr > P::TPC) nr = P::TPC;
                const unsigned bytes = (unsigned)(nr * P::N * 16);
                fence_proxy_async(); // the exchange's generic-proxy traffic precedes the bulk copy's writes
                mbar_expect_tx(&mbar, bytes);
                bulk_copy_g2s(smem, io.in + g * P::TPC * io.n, bytes, &mbar);
            }
        };
        if constexpr (STAGED) {
            if (tid == 0) {
                mbar_init(&mbar, 1);
                fence_mbar_init();
            }
            __syncthreads();
            if (tid == 0 && (long)blockIdx.x < groups) stage_issue(blockIdx.x);
        }
        for (long g = blockIdx.x; g < groups; g += gridDim.x) {
            const long row = g * P::TPC + slot;
            const bool active = row < rows;
            double2 x[EPT];
            if constexpr (STAGED) {
                mbar_wait(&mbar, phase);
                phase ^= 1;
#pragma unroll
                for (int u = 0; u < P0::U; u++)
#pragma unroll
                    for (int q = 0; q < P0::R; q++)
                        x[u * P0::R + q] = io.from_raw(smem[slot * P::N + P0::src_index(t, u, q)]);
                __syncthreads(); // everyone has read the stage before the exchange overwrites it
            } else if (active) {
#pragma unroll
                for (int u = 0; u < P0::U; u++)
#pragma unroll
                    for (int q = 0; q < P0::R; q++) x[u * P0::R + q] = io.load(row, P0::src_index(t, u, q));
            } else {
#pragma unroll
                for (int e = 0; e < EPT; e++) x[e] = make_double2(0.0, 0.0);
            }
            const long gn = g + gridDim.x;
            // after the group's LAST exchange the buffer is idle: request the next group's rows
            auto prefetch = [&](int pass_after) {
                if (STAGED && !IO::kEpilogueExchange && pass_after == P::NP - 1 && tid == 0 && gn < groups) stage_issue(gn);
            };
            P0::compute(x, tw0.v);
            xchg<P0, P1>(buf, t, x);
            prefetch(1);
            {
                double2 tw[P1::NTW];
                P1::load_tw(table, t, tw);
                P1::compute(x, tw);
            }
            if (P::NP > 2) {
                xchg<P1, P2>(buf, t, x);
                prefetch(2);
                double2 tw[P2::NTW];
                P2::load_tw(table, t, tw);
                P2::compute(x, tw);
            }
            if (P::NP > 3) {
                xchg<P2, P3>(buf, t, x);
                prefetch(3);
                double2 tw[P3::NTW];
                P3::load_tw(table, t, tw);
                P3::compute(x, tw);
            }
            using PL = PassD<P, P::NP - 1>;
            if constexpr (IO::kEpilogueExchange) {
                // rfft: all bins of the row back into the (padded) buffer, then thread t twists bins t + u T.
                // The staged prefetch (issued after the last exchange) has to wait for this one instead.
#pragma unroll
                for (int u = 0; u < PL::U; u++)
#pragma unroll
                    for (int w = 0; w < PL::R; w++) buf[PL::dst_pad(PL::dst_base(t, u), w)] = x[u * PL::R + w];
                __syncthreads();
                if (active) {
#pragma unroll
                    for (int u = 0; u < EPT; u++) io.epilogue(row, t + u * P::T, buf);
                }
                __syncthreads();
                if (STAGED && tid == 0 && gn < groups) stage_issue(gn);
            } else if (active) {
#pragma unroll
                for (int u = 0; u < PL::U; u++)
#pragma unroll
                    for (int w = 0; w < PL::R; w++) io.store(row, PL::dst_index(t, u, w), x[u * PL::R + w]);
            }
        }
    }
#endif
};

#ifdef __CUDACC__
template <int L, class IO, bool STAGED>
__global__ void __launch_bounds__((PlanD<L>::CTA)) fft_f64_kernel(const __grid_constant__ IO io, const __grid_constant__ Tw0D tw0,
                                                                 const double2 *__restrict__ table, long rows)
{
    extern __shared__ __align__(128) double2 smem_d[];
    CtaFftD<L, IO>::template run<STAGED>(io, tw0, table, rows, smem_d);
}
#endif

} // namespace kofft
