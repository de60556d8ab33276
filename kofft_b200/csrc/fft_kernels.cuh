// fft_kernels.cuh -- the CTA body around fft_engine.cuh plus the I/O policies that fuse
// kofft's per-call pre/post passes into the first load / last store of the transform:
//   IoC2C      FftImpl::fft / ifft on contiguous rows        (src/fft.rs:1054-1174)
//   IoGeneric  strided and split (SoA) entry points           (src/fft.rs:1175-1336, 1365-1439)
//   IoStft     stft framing + windowing                        (src/stft.rs:91-102)
//   IoIstft    istft: ifft + real*window                       (src/stft.rs:141-147)
//   IoRfft     rfft pack + Hermitian twist                     (src/rfft.rs:444-463)
//   IoIrfft    irfft untwist + unpack                          (src/rfft.rs:485-503)
#pragma once
#include "async_copy.cuh"
#include "fft_engine.cuh"

namespace kofft {

#ifdef __CUDACC__
#define KOFFT_LDG(p) __ldg(p)
#define KOFFT_LDCG(p) __ldcg(p) // L2 only: data written by other CTAs of the same launch
#else
#define KOFFT_LDG(p) (*(p))
#define KOFFT_LDCG(p) (*(p))
#endif

// ------------------------------------------------------------------------------------------
// I/O policies.  load(row, i) returns element i of the length-N complex sequence fed to the
// forward Stockham core for transform `row`; store(row, i, v) receives output bin i.
// ------------------------------------------------------------------------------------------

// ifft = conj -> fft -> conj, re*scale, im*scale  (src/fft.rs:1163-1172)
template <bool INV>
KHD float2 pre_conj(float2 v)
{
    if (INV) v.y = -v.y;
    return v;
}
template <bool INV>
KHD float2 post_conj_scale(float2 v, float scale)
{
    if (INV) {
        v.y = -v.y;
        v.x = mul_rn(v.x, scale);
        v.y = mul_rn(v.y, scale);
    }
    return v;
}

template <bool INV>
struct IoC2C {
    static constexpr bool kStoreAux = false;
    static constexpr bool kEpilogueExchange = false;
    const float2 *__restrict__ in;
    float2 *__restrict__ out;
    long n;
    float scale; // 1/n, computed on the host as 1.0f / (float)n
    KHD void init(int) {}
    KHD float2 load(long row, int i) const { return pre_conj<INV>(KOFFT_LDG(in + row * n + i)); }
    KHD void store(long row, int i, float2 v) const { out[row * n + i] = post_conj_scale<INV>(v, scale); }
    static constexpr bool kStageable = true;
    static constexpr bool kLoadAux = false;
    KHD float load_aux(int) const { return 0.0f; }
    // staging: the TPC rows of a group are one contiguous byte range.  The group a CTA works on
    // is part of the IO state (group_init / group_next) so the row loop needs no divisions.
    long cur_g = 0, g_step = 0;
    KHD void group_init(long g0, long step, int) { cur_g = g0; g_step = step; }
    KHD void group_next(int) { cur_g += g_step; }
    KHD unsigned stage_bytes(int tpc, long rows) const
    {
        long nr = rows - cur_g * tpc;
        return (unsigned)((nr < tpc ? nr : tpc) * n * 8);
    }
    KHD const void *stage_src(int tpc) const { return in + cur_g * tpc * n; }
    KHD int row_begin(int) const { return 0; }
    template <bool FULL>
    KHD float2 load_staged(const unsigned char *stage, int, int slot, int i, float) const
    {
        return pre_conj<INV>(reinterpret_cast<const float2 *>(stage)[slot * n + i]);
    }
    // large-N column pass: raw row pointer for asynchronous tile copies + the load-side transform
    KHD const float2 *row_ptr(long row) const { return in + row * n; }
    KHD float2 from_raw(float2 v) const { return pre_conj<INV>(v); }
    KHD bool row_full(int) const { return true; }
#if defined(__CUDACC__) || defined(KOFFT_EMU)
    // L2-hinted accessors of the pipelined large-N kernel (fft_large.cuh LargePipe): the rows stream
    // through L2 with evict_first so they do not displace the pinned intermediate
    KD float2 load_hint(long row, int i, unsigned long long pol) const { return pre_conj<INV>(ldg_hint(in + row * n + i, pol)); }
    KD void store_hint(long row, int i, float2 v, unsigned long long pol) const
    {
        stg_hint(out + row * n + i, post_conj_scale<INV>(v, scale), pol);
    }
#endif
};

// element e of row r lives at  re[r*row_stride + e*elem_stride]  (strides in floats);
// interleaved data passes im = re + 1 and elem_stride = 2*stride.
template <bool INV>
struct IoGeneric {
    static constexpr bool kStoreAux = false;
    static constexpr bool kEpilogueExchange = false;
    const float *__restrict__ in_re;
    const float *__restrict__ in_im;
    float *__restrict__ out_re;
    float *__restrict__ out_im;
    long in_es, in_rs, out_es, out_rs;
    float scale;
    static constexpr bool kStageable = false;
    static constexpr bool kLoadAux = false;
    KHD void init(int) {}
    KHD float2 load(long row, int i) const
    {
        long o = row * in_rs + (long)i * in_es;
        return pre_conj<INV>(make_float2(KOFFT_LDG(in_re + o), KOFFT_LDG(in_im + o)));
    }
    KHD void store(long row, int i, float2 v) const
    {
        v = post_conj_scale<INV>(v, scale);
        long o = row * out_rs + (long)i * out_es;
        out_re[o] = v.x;
        out_im[o] = v.y;
    }
};

// stft: frame f of channel c, x[i] = signal[c][f*hop + i] * window[i], 0 past the end, imag 0
struct IoStft {
    static constexpr bool kStoreAux = false;
    static constexpr bool kEpilogueExchange = false;
    const float *__restrict__ signal;
    const float *__restrict__ window;
    float2 *__restrict__ frames;
    long len;     // samples per channel
    long nframes; // frames per channel
    long hop;
    long n; // win_len
    KHD void init(int) {}
    KHD float2 load(long row, int i) const
    {
        long c = row / nframes, f = row - c * nframes;
        long pos = f * hop + i;
        float x = 0.0f;
        if (pos < len) x = mul_rn(KOFFT_LDG(signal + c * len + pos), KOFFT_LDG(window + i));
        return make_float2(x, 0.0f);
    }
    KHD void store(long row, int i, float2 v) const { frames[row * n + i] = v; }
    // staging: the TPC consecutive frames of group g (same channel: the host only picks the
    // staged kernel when nframes % TPC == 0) read one contiguous run of (TPC-1)*hop + n samples
    static constexpr bool kStageable = true;
    static constexpr bool kLoadAux = true;
    KHD float load_aux(int i) const { return KOFFT_LDG(window + i); } // window value, kept in a register
    // Group state of the staged kernel, advanced by ADDITIONS only as the persistent CTA strides over the
    // groups (two divisions and a few products per CTA lifetime; the per-group bookkeeping used to be
    // ~100 instructions of 64-bit multiplies and compares, a sixth of the issue slots of a frame):
    //   cur_f0  first frame of the group within its channel
    //   sig_off element offset of the group's first sample in `signal`  (= c*len + f0*hop)
    //   rem     samples from there to the end of the channel            (= len - f0*hop, may be <= 0)
    long cur_f0 = 0, adv_f = 0;
    long sig_off = 0, rem = 0, adv_off = 0, adv_rem = 0, wrap_off = 0, wrap_rem = 0;
    KHD void group_init(long g0, long step, int tpc)
    {
        const long r0 = g0 * tpc;
        const long c0 = r0 / nframes;
        cur_f0 = r0 - c0 * nframes;
        const long a = step * tpc;
        const long adv_c = a / nframes;
        adv_f = a - adv_c * nframes;
        sig_off = c0 * len + cur_f0 * hop;
        rem = len - cur_f0 * hop;
        adv_off = adv_c * len + adv_f * hop;
        adv_rem = adv_f * hop;
        wrap_off = len - nframes * hop; // passing the end of a channel: one more channel, nframes fewer frames
        wrap_rem = nframes * hop;
    }
    KHD void group_next(int)
    {
        cur_f0 += adv_f;
        sig_off += adv_off;
        rem -= adv_rem;
        if (cur_f0 >= nframes) {
            cur_f0 -= nframes;
            sig_off += wrap_off;
            rem += wrap_rem;
        }
    }
    KHD unsigned stage_bytes(int tpc, long) const
    {
        const long want = (long)(tpc - 1) * hop + n;
        if (rem <= 0) return 0u;
        return (unsigned)((rem < want ? rem : want) * 4);
    }
    KHD const void *stage_src(int) const { return signal + sig_off; }
    // samples of this slot's frame that lie inside the signal, clamped to [0, n]
    KHD int row_begin(int slot) const
    {
        const long r = rem - (long)slot * hop;
        return r <= 0 ? 0 : (r >= n ? (int)n : (int)r);
    }
    // FULL: the whole frame lies inside the signal (rem == n), no per-element bound check
    template <bool FULL>
    KHD float2 load_staged(const unsigned char *stage, int rem, int slot, int i, float w) const
    {
        float x = 0.0f;
        if (FULL || i < rem) x = mul_rn(reinterpret_cast<const float *>(stage)[(long)slot * hop + i], w);
        return make_float2(x, 0.0f);
    }
    KHD bool row_full(int rem) const { return rem >= (int)n; }
};

// stft_magnitudes (src/visual/spectrogram.rs:52-76): the spectrogram consumer keeps only
// |X[k]| for k < win_len/2 and the largest magnitude.  Fused behind the last FFT stage: the
// complex frames never reach HBM (4*(N/2) instead of 8*N bytes per frame).  mag is
// sqrt(re*re + im*im) with every operation rounded as in the reference; the maximum is exact and
// order-independent (non-negative floats order like their bit patterns -> integer atomicMax).
struct IoStftMag : IoStft {
    float *__restrict__ mags; // [rows][n/2]
    int *max_bits;            // running maximum of this launch, as float bits (initialised to 0)
    mutable float tmax;       // this thread's running maximum
    KHD void store(long row, int i, float2 v) const
    {
        const int half = (int)(n >> 1);
        if (i < half) {
            const float mag = sqrt_rn(add_rn(mul_rn(v.x, v.x), mul_rn(v.y, v.y)));
            mags[row * half + i] = mag;
            if (mag > tmax) tmax = mag; // NaN never becomes the maximum (src/visual/spectrogram.rs:70)
        }
    }
    KHD void finish() const
    {
        if (tmax > 0.0f) atomicMax(max_bits, float_bits(tmax));
    }
};

// istft stage 1: time[row][i] = (ifft(frame).re) * window[i]; the overlap-add is a second,
// order-preserving gather kernel (ola_kernel below).
struct IoIstft {
    static constexpr bool kEpilogueExchange = false;
    const float2 *__restrict__ frames;
    const float *__restrict__ window;
    float *__restrict__ time; // [rows][n] f32
    long n;
    float scale;
    KHD void init(int) {}
    KHD float2 load(long row, int i) const { return pre_conj<true>(KOFFT_LDG(frames + row * n + i)); }
    KHD void store(long row, int i, float2 v) const
    {
        float re = mul_rn(v.x, scale);
        time[row * n + i] = mul_rn(re, KOFFT_LDG(window + i));
    }
    static constexpr bool kStoreAux = true;
    KHD float store_aux_value(int i) const { return KOFFT_LDG(window + i); }
    KHD void store_aux(long row, int i, float2 v, float w) const
    {
        float re = mul_rn(v.x, scale);
        time[row * n + i] = mul_rn(re, w);
    }
    static constexpr bool kStageable = true;
    static constexpr bool kLoadAux = false;
    KHD float load_aux(int) const { return 0.0f; }
    // staging: the TPC rows of a group are one contiguous byte range.  The group a CTA works on
    // is part of the IO state (group_init / group_next) so the row loop needs no divisions.
    long cur_g = 0, g_step = 0;
    KHD void group_init(long g0, long step, int) { cur_g = g0; g_step = step; }
    KHD void group_next(int) { cur_g += g_step; }
    KHD unsigned stage_bytes(int tpc, long rows) const
    {
        long nr = rows - cur_g * tpc;
        return (unsigned)((nr < tpc ? nr : tpc) * n * 8);
    }
    KHD const void *stage_src(int tpc) const { return frames + cur_g * tpc * n; }
    KHD int row_begin(int) const { return 0; }
    template <bool FULL>
    KHD float2 load_staged(const unsigned char *stage, int, int slot, int i, float) const
    {
        return pre_conj<true>(reinterpret_cast<const float2 *>(stage)[slot * n + i]);
    }
    KHD bool row_full(int) const { return true; }
};

// rfft: input row = 2m reals viewed as m complex (pack is a reinterpretation); after the
// length-m FFT the bins are exchanged once more through shared memory and twisted.
template <bool EXACT>
struct IoRfft {
    static constexpr bool kStoreAux = false;
    static constexpr bool kEpilogueExchange = true;
    const float2 *__restrict__ in;     // [rows][m]
    float2 *__restrict__ out;          // [rows][m + 1]
    const float2 *__restrict__ rtw;    // T'[k] = exp(-i pi k / m), m entries (src/rfft.rs:172-183)
    long m;
    KHD void init(int) {}
    KHD float2 load(long row, int i) const { return KOFFT_LDG(in + row * m + i); }
    KHD void store(long, int, float2) const {}
    static constexpr bool kStageable = true;
    static constexpr bool kLoadAux = false;
    KHD float load_aux(int) const { return 0.0f; }
    // staging: the TPC rows of a group are one contiguous byte range.  The group a CTA works on
    // is part of the IO state (group_init / group_next) so the row loop needs no divisions.
    long cur_g = 0, g_step = 0;
    KHD void group_init(long g0, long step, int) { cur_g = g0; g_step = step; }
    KHD void group_next(int) { cur_g += g_step; }
    KHD unsigned stage_bytes(int tpc, long rows) const
    {
        long nr = rows - cur_g * tpc;
        return (unsigned)((nr < tpc ? nr : tpc) * m * 8);
    }
    KHD const void *stage_src(int tpc) const { return in + cur_g * tpc * m; }
    KHD int row_begin(int) const { return 0; }
    template <bool FULL>
    KHD float2 load_staged(const unsigned char *stage, int, int slot, int i, float) const
    {
        return reinterpret_cast<const float2 *>(stage)[slot * m + i];
    }
    KHD const float2 *row_ptr(long row) const { return in + row * m; }
    KHD float2 from_raw(float2 v) const { return v; }
    KHD bool row_full(int) const { return true; }
    // Hermitian twist of bin k (src/rfft.rs:450-463): a = Y[k], ym = Y[m-k] (for k = 0: a = Y[0])
    KHD void twist_store(long row, long k, float2 a, float2 ym) const
    {
        twist_store_tw(row, k, a, ym, k == 0 ? make_float2(1.0f, 0.0f) : KOFFT_LDG(rtw + k));
    }
    // same, with the table entry T'[k] supplied by the caller
    KHD void twist_store_tw(long row, long k, float2 a, float2 ym, float2 tw) const
    {
        float2 *o = out + row * (m + 1);
        if (k == 0) {
            o[0] = make_float2(add_rn(a.x, a.y), 0.0f);
            o[m] = make_float2(sub_rn(a.x, a.y), 0.0f);
            return;
        }
        o[k] = twist(a, ym, tw);
    }
    // out[k] for 0 < k < m (src/rfft.rs:453-461)
    KHD float2 twist(float2 a, float2 ym, float2 tw) const
    {
        float2 b = make_float2(ym.x, -ym.y);
        float2 sum = add2(a, b), diff = sub2(a, b);
        if (EXACT) {
            // the same individually rounded operations in packed form (8 issue slots instead of 14): the four
            // products of tw * diff as two FMUL2 combined by scalar adds (a packed product must not feed a packed
            // add: ptxas would contract it), -t.re obtained as q - p, which is -(p - q) exactly
            float2 p = mul2(make_float2(tw.x, tw.x), diff);                     // (tw.re d.re, tw.re d.im)
            float2 q = mul2(make_float2(tw.y, tw.y), make_float2(diff.y, diff.x)); // (tw.im d.im, tw.im d.re)
            float2 r = make_float2(add_rn(p.y, q.y), sub_rn(q.x, p.x));         // (t.im, -t.re)
            return mul2(add2(sum, r), make_float2(0.5f, 0.5f));                  // (sum + (t.im, -t.re)) / 2
        }
        float2 t = cmul<EXACT>(tw, diff);
        float2 temp = make_float2(add_rn(sum.x, t.y), sub_rn(sum.y, t.x)); // sum + (t.im, -t.re)
        return make_float2(mul_rn(temp.x, 0.5f), mul_rn(temp.y, 0.5f));
    }
#if defined(__CUDACC__) || defined(KOFFT_EMU)
    // L2-hinted accessors of the pipelined large-N kernel (see IoC2C)
    KD float2 load_hint(long row, int i, unsigned long long pol) const { return ldg_hint(in + row * m + i, pol); }
    KD void twist_store_tw_hint(long row, long k, float2 a, float2 ym, float2 tw, unsigned long long pol) const
    {
        float2 *o = out + row * (m + 1);
        if (k == 0) {
            stg_hint(o, make_float2(add_rn(a.x, a.y), 0.0f), pol);
            stg_hint(o + m, make_float2(sub_rn(a.x, a.y), 0.0f), pol);
            return;
        }
        stg_hint(o + k, twist(a, ym, tw), pol);
    }
#endif
    // Y: padded shared copy of the m FFT bins of this row
    KHD void epilogue(long row, int k, const float2 *Y) const
    {
        twist_store(row, k, Y[pad(k)], k == 0 ? Y[0] : Y[pad((int)m - k)]);
    }
};

// irfft: the untwist (src/rfft.rs:485-498) is evaluated straight from global memory while
// loading element i (it needs X[i] and X[m-i]), then ifft, then unpack to 2m reals.
template <bool EXACT>
struct IoIrfft {
    static constexpr bool kStoreAux = false;
    static constexpr bool kEpilogueExchange = false;
    const float2 *__restrict__ in;   // [rows][m + 1]
    float2 *__restrict__ out;        // [rows][m] == 2m reals
    const float2 *__restrict__ rtw;
    long m;
    float scale; // 1/m
    static constexpr bool kStageable = false;
    static constexpr bool kLoadAux = false;
    KHD void init(int) {}
    // element 0 of the packed spectrum (src/rfft.rs:485-488) and element i from X[i], X[m-i], T'[i] (:489-498)
    KHD float2 untwist0(float2 x0, float2 xm) const
    {
        return make_float2(mul_rn(add_rn(x0.x, xm.x), 0.5f), mul_rn(sub_rn(x0.x, xm.x), 0.5f));
    }
    KHD float2 untwist(float2 a, float2 xm, float2 tw) const
    {
        float2 b = make_float2(xm.x, -xm.y);
        float2 sum = add2(a, b), diff = sub2(a, b);
        float2 t = cmul<EXACT>(make_float2(tw.x, -tw.y), diff);
        float2 temp = make_float2(sub_rn(sum.x, t.y), add_rn(sum.y, t.x)); // sum - (t.im, -t.re)
        return make_float2(mul_rn(temp.x, 0.5f), mul_rn(temp.y, 0.5f));
    }
    KHD float2 load(long row, int i) const
    {
        const float2 *X = in + row * (m + 1);
        float2 v;
        if (i == 0)
            v = untwist0(KOFFT_LDG(X), KOFFT_LDG(X + m));
        else
            v = untwist(KOFFT_LDG(X + i), KOFFT_LDG(X + (m - i)), KOFFT_LDG(rtw + i));
        return pre_conj<true>(v);
    }
    KHD float2 from_raw(float2 v) const { return pre_conj<true>(v); } // an untwisted row (split kernel, PRE)
    KHD void store(long row, int i, float2 v) const { out[row * m + i] = post_conj_scale<true>(v, scale); }
};

// Compile-time properties of an IO policy that shape the kernel around it.
//   kRealInput: load() returns imag == +0 -> pass 0 takes the real-input butterfly shortcuts
//   kMinCta   : Plan<L, kMinCta> (see fft_engine.cuh)
#ifndef KOFFT_STFT_MIN_CTA
#define KOFFT_STFT_MIN_CTA 128
#endif
#ifndef KOFFT_STFT_REAL
#define KOFFT_STFT_REAL 1
#endif
// STFT occupancy knobs (scripts/build_variants.sh): pass-1 twiddles from a small shared-memory table
// instead of 30 registers, the stage sized for real samples (N*4 instead of N*8 bytes), and the
// number of CTAs per SM the kernel is compiled for (0 = the generic 512 / CTA)
#ifndef KOFFT_STFT_TW1_SMEM
#define KOFFT_STFT_TW1_SMEM 0
#endif
#ifndef KOFFT_STFT_BLOCKS
#define KOFFT_STFT_BLOCKS 0
#endif
// copies of the row loop body the compiler lays out (build knob for the STFT: with two copies the per-frame output base
// is rewritten every other frame, so the stores still reading it delay the loop less)
#ifndef KOFFT_STFT_UNROLL
#define KOFFT_STFT_UNROLL 3 // measured (profiles/r05g): 1: 3.86 ms, 2: 3.79, 3: 3.73, 4: 4.00, 5: 4.48 (config 4 shape on 16 channels)
#endif
template <class IO> struct RowUnroll { static constexpr int value = 1; };
template <> struct RowUnroll<IoStft> { static constexpr int value = KOFFT_STFT_UNROLL; };
#ifndef KOFFT_STFT_MAG_UNROLL
#define KOFFT_STFT_MAG_UNROLL 1
#endif
template <> struct RowUnroll<IoStftMag> { static constexpr int value = KOFFT_STFT_MAG_UNROLL; };
#ifndef KOFFT_C2C_UNROLL
#define KOFFT_C2C_UNROLL 1
#endif
template <bool INV> struct RowUnroll<IoC2C<INV>> { static constexpr int value = KOFFT_C2C_UNROLL; };

template <class IO>
struct IoTraits {
    static constexpr bool kRealInput = false;
    static constexpr int kMinCta = 256;
    static constexpr bool kRowPtr = false; // has row_ptr()/from_raw(): rows are plain contiguous float2
    static constexpr bool kFinish = false; // has finish(): called once per thread when the CTA is done
    static constexpr bool kHint = false;   // has load_hint()/store_hint(): L2-hinted row accessors
    static constexpr bool kTw1Smem = false; // pass-1 twiddles from shared memory (staged kernel only)
    static constexpr bool kStageHalf = false; // the staged input is real: N*4 bytes per transform
    static constexpr int kMinBlocks = 0;    // CTAs per SM to compile for (0 = 512 / CTA)
};
template <bool INV>
struct IoTraits<IoC2C<INV>> {
    static constexpr bool kRealInput = false;
    static constexpr int kMinCta = 256;
    static constexpr bool kRowPtr = true;
    static constexpr bool kFinish = false;
    static constexpr bool kHint = true;
    static constexpr bool kTw1Smem = false;
    static constexpr bool kStageHalf = false;
    static constexpr int kMinBlocks = 0;
};
template <bool EXACT>
struct IoTraits<IoRfft<EXACT>> {
    static constexpr bool kRealInput = false;
    static constexpr int kMinCta = 256;
    static constexpr bool kRowPtr = true;
    static constexpr bool kFinish = false;
    static constexpr bool kHint = true;
    static constexpr bool kTw1Smem = false;
    static constexpr bool kStageHalf = false;
    static constexpr int kMinBlocks = 0;
};
template <>
struct IoTraits<IoStft> {
    static constexpr bool kRealInput = KOFFT_STFT_REAL != 0;
    static constexpr int kMinCta = KOFFT_STFT_MIN_CTA;
    static constexpr bool kRowPtr = false;
    static constexpr bool kFinish = false;
    static constexpr bool kHint = false;
    static constexpr bool kTw1Smem = KOFFT_STFT_TW1_SMEM != 0;
    static constexpr bool kStageHalf = KOFFT_STFT_TW1_SMEM != 0;
    static constexpr int kMinBlocks = KOFFT_STFT_BLOCKS;
};
template <>
struct IoTraits<IoStftMag> {
    static constexpr bool kRealInput = KOFFT_STFT_REAL != 0;
    static constexpr int kMinCta = KOFFT_STFT_MIN_CTA;
    static constexpr bool kRowPtr = false;
    static constexpr bool kFinish = true;
    static constexpr bool kHint = false;
    static constexpr bool kTw1Smem = KOFFT_STFT_TW1_SMEM != 0;
    static constexpr bool kStageHalf = KOFFT_STFT_TW1_SMEM != 0;
    static constexpr int kMinBlocks = KOFFT_STFT_BLOCKS;
};

// ------------------------------------------------------------------------------------------
// The CTA body, written as per-thread phase functions so that tests/emu can run the very
// same code on the CPU (phase by phase over all threads) -- see tests/emu/emu_engine.cpp.
// ------------------------------------------------------------------------------------------
template <class P, bool EXACT, class IO>
struct CtaFft {
    using P0 = Pass<P, 0, EXACT, true, IoTraits<IO>::kRealInput>;
    using P1 = Pass<P, 1, EXACT>;
    using P2 = Pass<P, (P::NP > 2 ? 2 : 1), EXACT>;
    using P3 = Pass<P, (P::NP > 3 ? 3 : 1), EXACT>;

    // shared-memory layout of the staged kernel: [stage][exchange buffers][pass-1 twiddle table]
    static constexpr bool TW1S = IoTraits<IO>::kTw1Smem && P::TW_REGS && P::NP == 3;
    static constexpr int STAGE_B = IoTraits<IO>::kStageHalf ? P::STAGE_BYTES / 2 : P::STAGE_BYTES;
    static constexpr int NK1 = P::T >> P1::LJ;             // distinct groups k of pass 1 per transform
    static constexpr int TW1S_BYTES = TW1S ? NK1 * 16 * 8 : 0;
    static constexpr int SMEM_STAGED = STAGE_B + P::XCHG_BYTES + TW1S_BYTES;

    template <class PS>
    static KHD void load_global(const IO &io, long row, int t, float2 *x)
    {
#pragma unroll
        for (int u = 0; u < PS::U; u++)
#pragma unroll
            for (int q = 0; q < PS::R; q++) x[u * PS::R + q] = io.load(row, PS::src_index(t, u, q));
    }
    template <class PS>
    static KHD void store_global(const IO &io, long row, int t, const float2 *x)
    {
#pragma unroll
        for (int u = 0; u < PS::U; u++)
#pragma unroll
            for (int w = 0; w < PS::R; w++) io.store(row, PS::dst_index(t, u, w), x[u * PS::R + w]);
    }
    // same, with per-element constants (the ISTFT window) preloaded into registers
    template <class PS>
    static KHD void store_global_aux(const IO &io, long row, int t, const float2 *x, const float *aux)
    {
#pragma unroll
        for (int u = 0; u < PS::U; u++)
#pragma unroll
            for (int w = 0; w < PS::R; w++)
                io.store_aux(row, PS::dst_index(t, u, w), x[u * PS::R + w], aux[u * PS::R + w]);
    }
    template <class PS>
    static KHD void store_smem(float2 *buf, int t, const float2 *x)
    {
#pragma unroll
        for (int u = 0; u < PS::U; u++)
#pragma unroll
            for (int w = 0; w < PS::R; w++) buf[PS::dst_pad(PS::dst_base(t, u), w)] = x[u * PS::R + w];
    }
    template <class PS>
    static KHD void load_smem(const float2 *buf, int t, float2 *x)
    {
#pragma unroll
        for (int u = 0; u < PS::U; u++)
#pragma unroll
            for (int q = 0; q < PS::R; q++) x[u * PS::R + q] = buf[PS::src_pad(PS::src_base(t, u), q)];
    }
    // rfft-style epilogue: thread t handles bins k = t + u*T
    static KHD void epilogue(const IO &io, long row, int t, const float2 *buf)
    {
#pragma unroll
        for (int u = 0; u < EPT; u++) io.epilogue(row, t + u * P::T, buf);
    }

#if defined(__CUDACC__) || defined(KOFFT_EMU)
    // rows: number of transforms.  smem: [stage (STAGED only)] [NBUF exchange buffers].
    //
    // STAGED: the next row group's raw input (one contiguous byte range) is fetched by a single
    // TMA bulk copy (cp.async.bulk, completion on an mbarrier) into the stage while the CTA is
    // still computing passes 1.. of the current group, so the HBM read latency of group g+1
    // overlaps the butterflies of group g.  Pass 0 then reads the stage instead of HBM.
    template <bool STAGED>
    static KD void run(IO io, const Tw0 &tw0, const float2 *__restrict__ table, long rows, float2 *smem)
    {
        const int tid = threadIdx.x;
        // which of the CTA's TPC transforms.  A compile-time 0 for the one-frame STFT CTAs (it removes a
        // per-thread 64-bit multiply from the staged load); the other kernels keep the division: with the
        // constant the C2C N = 4096 kernel measured 1.5 % slower (different schedule, profiles/r02i)
        const int slot = (P::TPC == 1 && IoTraits<IO>::kMinCta < 256) ? 0 : tid / P::T;
        const int t = tid - slot * P::T;
        unsigned char *smem_stage = reinterpret_cast<unsigned char *>(smem);
        float2 *xch = STAGED ? smem + STAGE_B / 8 : smem;
        float2 *buf0 = xch + slot * P::PADN;
        float2 *buf1 = P::NBUF == 2 ? buf0 + P::TPC * P::PADN : buf0;
        int par = 0;
        __shared__ __align__(8) unsigned long long mbar;
        unsigned phase = 0;
        const long groups = (rows + P::TPC - 1) / P::TPC;

        if constexpr (STAGED) {
            if (tid == 0) {
                mbar_init(&mbar, 1);
                fence_mbar_init();
            }
            __syncthreads();
        }

        io.init(t);
        float2 tw1[P1::NTW], tw2[P2::NTW], tw3[P3::NTW];
        // TW1S: the pass-1 twiddles depend only on the group k = t >> logJ (a handful of values), so they
        // are read from a [k][16] table in shared memory (broadcast loads) instead of holding 30 registers
        constexpr bool TW1_FROM_SMEM = STAGED && TW1S;
        float2 *tw1s = xch + P::XCHG_BYTES / 8;
        if constexpr (TW1_FROM_SMEM) {
            for (int i = tid; i < NK1 * 15; i += P::CTA) {
                const int k = i / 15, e = i - k * 15;
                int tl = 0;
                while ((2 << tl) - 1 <= e) tl++;
                tw1s[k * 16 + e] = table[tw_index<P>(1, tl, k, e + 1 - (1 << tl))];
            }
            __syncthreads();
            if (P::NP > 2) P2::load_tw(table, t, tw2);
        } else if (P::TW_REGS) {
            P1::load_tw(table, t, tw1);
            if (P::NP > 2) P2::load_tw(table, t, tw2);
        }
        const float2 *tw1p = TW1_FROM_SMEM ? tw1s + (t >> P1::LJ) * 16 : tw1;
        float aux[EPT]; // per-element constants of pass-0 loads (the STFT window), hoisted out of the row loop
        if constexpr (STAGED && IO::kLoadAux) {
#pragma unroll
            for (int u = 0; u < P0::U; u++)
#pragma unroll
                for (int q = 0; q < P0::R; q++) aux[u * P0::R + q] = io.load_aux(P0::src_index(t, u, q));
        }

        using PLast = Pass<P, P::NP - 1, EXACT>;
        float auxo[EPT]; // per-element constants of the final stores (the ISTFT window)
        if constexpr (IO::kStoreAux) {
#pragma unroll
            for (int u = 0; u < PLast::U; u++)
#pragma unroll
                for (int w = 0; w < PLast::R; w++) auxo[u * PLast::R + w] = io.store_aux_value(PLast::dst_index(t, u, w));
        }

        if constexpr (STAGED) {
            io.group_init(blockIdx.x, gridDim.x, P::TPC);
            if (tid == 0 && (long)blockIdx.x < groups) stage_issue(io, smem_stage, &mbar, rows);
        }
#pragma unroll(RowUnroll<IO>::value)
        for (long g = blockIdx.x; g < groups; g += gridDim.x) {
            const long row = g * P::TPC + slot;
            const bool active = row < rows;
            float2 x[EPT];
            if constexpr (STAGED) {
                const unsigned char *stage = smem_stage;
                if (io.stage_bytes(P::TPC, rows) != 0) {
                    mbar_wait(&mbar, phase);
                    phase ^= 1;
                }
                const int rctx = io.row_begin(slot);
                if (io.row_full(rctx)) { // warp-uniform: a slot's threads share the frame
#pragma unroll
                    for (int u = 0; u < P0::U; u++)
#pragma unroll
                        for (int q = 0; q < P0::R; q++)
                            x[u * P0::R + q] = io.template load_staged<true>(stage, rctx, slot, P0::src_index(t, u, q),
                                                                             IO::kLoadAux ? aux[u * P0::R + q] : 0.0f);
                } else {
#pragma unroll
                    for (int u = 0; u < P0::U; u++)
#pragma unroll
                        for (int q = 0; q < P0::R; q++)
                            x[u * P0::R + q] = io.template load_staged<false>(stage, rctx, slot, P0::src_index(t, u, q),
                                                                              IO::kLoadAux ? aux[u * P0::R + q] : 0.0f);
                }
            } else {
                if (active) {
                    load_global<P0>(io, row, t, x);
                } else {
#pragma unroll
                    for (int e = 0; e < EPT; e++) x[e] = make_float2(0.f, 0.f);
                }
            }
            P0::compute(x, tw0.v);

            float2 *b = par ? buf1 : buf0;
            par ^= 1;
            store_smem<P0>(b, t, x);
            __syncthreads();
            if constexpr (STAGED) { // every thread has consumed the stage: refill it for this CTA's next group
                io.group_next(P::TPC);
                if (tid == 0 && g + gridDim.x < groups) stage_issue(io, smem_stage, &mbar, rows);
            }
            load_smem<P1>(b, t, x);
            if (P::NBUF == 1) __syncthreads();
            if (!P::TW_REGS) P1::load_tw(table, t, tw1);
            P1::compute(x, tw1p);

            if (P::NP > 2) {
                b = par ? buf1 : buf0;
                par ^= 1;
                store_smem<P1>(b, t, x);
                __syncthreads();
                load_smem<P2>(b, t, x);
                if (P::NBUF == 1) __syncthreads();
                if (!P::TW_REGS) P2::load_tw(table, t, tw2);
                P2::compute(x, tw2);
            }
            if (P::NP > 3) {
                b = par ? buf1 : buf0;
                par ^= 1;
                store_smem<P2>(b, t, x);
                __syncthreads();
                load_smem<P3>(b, t, x);
                if (P::NBUF == 1) __syncthreads();
                P3::load_tw(table, t, tw3);
                P3::compute(x, tw3);
            }

            using PL = Pass<P, P::NP - 1, EXACT>;
            if constexpr (IO::kEpilogueExchange) {
                b = par ? buf1 : buf0;
                par ^= 1;
                store_smem<PL>(b, t, x);
                __syncthreads();
                if (active) epilogue(io, row, t, b);
                if (P::NBUF == 1) __syncthreads();
            } else if constexpr (IO::kStoreAux) {
                if (active) store_global_aux<PL>(io, row, t, x, auxo);
            } else {
                if (active) store_global<PL>(io, row, t, x);
            }
        }
        if constexpr (IoTraits<IO>::kFinish) io.finish();
    }

    // one thread: arm the barrier with the byte count and start the bulk copy of io's current group
    static KD void stage_issue(const IO &io, unsigned char *stage, unsigned long long *bar, long rows)
    {
        const unsigned bytes = io.stage_bytes(P::TPC, rows);
        if (bytes == 0) return;
        mbar_expect_tx(bar, bytes);
        bulk_copy_g2s(stage, io.stage_src(P::TPC), bytes, bar);
    }
#endif
};

#ifdef __CUDACC__
template <int L, bool EXACT, class IO, bool STAGED>
__global__ void __launch_bounds__((Plan<L, IoTraits<IO>::kMinCta>::CTA),
                                  (IoTraits<IO>::kMinBlocks > 0 && Plan<L, IoTraits<IO>::kMinCta>::CTA == 128
                                       ? IoTraits<IO>::kMinBlocks
                                       : (Plan<L, IoTraits<IO>::kMinCta>::CTA <= 512 ? 512 / Plan<L, IoTraits<IO>::kMinCta>::CTA : 1)))
    fft_cta_kernel(const __grid_constant__ IO io, const __grid_constant__ Tw0 tw0,
                   const float2 *__restrict__ table, long rows)
{
    extern __shared__ __align__(128) float2 smem[];
    CtaFft<Plan<L, IoTraits<IO>::kMinCta>, EXACT, IO>::template run<STAGED>(io, tw0, table, rows, smem);
}
#endif

} // namespace kofft
