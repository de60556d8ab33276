// istft_fused.cuh -- istft in ONE kernel: inverse FFT, x 1/N, x window, ordered overlap-add and
// normalisation fused behind the last FFT stage (reference: src/stft.rs:117-156; the
// `inverse_parallel` variant :289-343 via zero_uncovered).
//
// One CTA = one frame at a time (N/16 threads, 16 bins per thread); a CTA owns a run of G
// consecutive frames of one channel and the output samples [F0*hop, Fend*hop) they start (the
// channel's last run also owns the tail up to out_len).  It walks the frames in increasing
// order -- first the H = ceil(N/hop)-1 halo frames before F0, whose tails reach into the owned
// samples (recomputed, not communicated) -- and ADDS each frame's windowed real part into an
// N-sample accumulator ring in shared memory, every thread adding its own 16 samples.  After
// frame f has been added no later frame touches samples [f*hop, (f+1)*hop): they are
// normalised by the summed window power, stored, and their ring slots recycled
// (re-initialised from the caller's `output`, which the reference accumulates into).  Per sample
// this is exactly the reference's sequence of f32 additions, in the same (frame) order, so the
// result is bit-identical and deterministic; no atomics, no intermediate frame buffer in HBM.
//
// Pipeline inside the frame loop (no barrier beyond the two the FFT exchanges need anyway):
//   iteration f:  P0 | sync A | finalise + recycle region f-1 | P1 | sync B | P2 | ring += frame f
// The ring add of frame f-1 precedes sync A of iteration f, the recycling of region f-1
// precedes sync B, and the ring add of frame f follows it.  The next frame is fetched by one TMA
// bulk copy while the current one is transformed; the initial-output values for the recycled
// slots are loaded at the top of the iteration so their latency hides behind pass 0.
//
// Window power: sample p = f*hop + r is covered by K(r) = floor((N-1-r)/hop)+1 frames in steady
// state; normtab[r] = w[r+(K-1)hop]^2 + ... + w[r]^2 summed in the reference's (frame) order is
// built once per CTA in shared memory.  The first ceil(N/hop)-1 regions of a channel and the
// tail after the last frame have fewer covering frames and take the generic loop.
//
// HBM traffic per frame: 8N (frame) * (1 + H/G) + 4 hop (initial output) + 4 hop (result).
#pragma once
#include "fft_kernels.cuh"

// registers per thread the fused kernel is compiled for (168 = three 128-thread CTAs per SM, spill-free;
// 128 = four CTAs per SM) -- a build-time knob for scripts/build_variants.sh
#ifndef KOFFT_ISTFT_REGS
#define KOFFT_ISTFT_REGS 168
#endif
// 1: ONE exchange buffer (a second barrier per exchange) -> 43 KB of shared memory per CTA at N = 2048,
// which together with KOFFT_ISTFT_REGS = 128 puts four CTAs on an SM instead of three
// 1: the recycled slots' initial values are loaded a whole iteration before their use (0: at the top of the iteration);
// 1: the ring accumulators of a frame are all read before any is updated
// (PREPIPE measured slower, HOIST 1.5 % faster: profiles/r04u); 1: the last butterfly layer computes real parts only
#ifndef KOFFT_ISTFT_PREPIPE
#define KOFFT_ISTFT_PREPIPE 0
#endif
#ifndef KOFFT_ISTFT_HOIST
#define KOFFT_ISTFT_HOIST 1
#endif
// copies of the frame loop body (measured, profiles/r05g: 2 is neutral in EXACT mode and 3 % faster in FAST mode, 3 is slower)
#ifndef KOFFT_ISTFT_UNROLL
#define KOFFT_ISTFT_UNROLL 2
#endif
#ifndef KOFFT_ISTFT_RE_LAST
#define KOFFT_ISTFT_RE_LAST 1
#endif
#ifndef KOFFT_ISTFT_ONEBUF
#define KOFFT_ISTFT_ONEBUF 0
#endif

namespace kofft {

struct IstftFusedArgs {
    const float2 *frames; // [channels][nframes][N]
    const float *window;  // [N]
    float *output;        // [channels][out_len], accumulated into
    float *norm;          // optional [channels][out_len]
    long channels, nframes, hop, out_len;
    int run_frames;       // G
    int zero_uncovered;
    float scale;          // 1/N
};

template <int L, bool EXACT>
struct IstftFused {
    using P = Plan<L, 32>; // one frame per CTA: N/16 threads
    static_assert(P::NP == 3 && P::NBUF == 2 && P::TPC == 1, "fused istft covers N = 512 .. 4096");
    using IO = IoIstft;
    using H = CtaFft<P, EXACT, IO>;
    using P0 = Pass<P, 0, EXACT>;
    using P1 = Pass<P, 1, EXACT>;
    using P2 = Pass<P, 2, EXACT>;
    static constexpr int N = P::N;
    static constexpr int CTA = P::CTA;
    static constexpr int FRAME_UNROLL = KOFFT_ISTFT_UNROLL; // copies of the frame loop body
    // CTAs per SM are bounded by shared memory (estimated at hop = N/4); give the registers that leaves
    static constexpr bool ONEBUF = KOFFT_ISTFT_ONEBUF != 0;
    static constexpr int XCHG_BYTES = ONEBUF ? P::PADN * 8 : P::XCHG_BYTES;
    static constexpr int SMEM_EST = P::STAGE_BYTES + XCHG_BYTES + N * 4 + N + 16;
    static constexpr int BY_SMEM = 227 * 1024 / (SMEM_EST + 1024);
    static constexpr int BY_REGS = 65536 / (CTA * KOFFT_ISTFT_REGS); // registers per thread the body is given
    static constexpr int BY_BOTH = BY_SMEM < BY_REGS ? BY_SMEM : BY_REGS;
#ifdef KOFFT_ISTFT_FORCE_BLOCKS
    static constexpr int MIN_BLOCKS = KOFFT_ISTFT_FORCE_BLOCKS;
#else
    static constexpr int MIN_BLOCKS = BY_BOTH > 8 ? 8 : (BY_BOTH < 1 ? 1 : BY_BOTH);
#endif
    // [stage N*8][exchange 2*PADN*8][ring N*4][normtab hop*4][mbarrier 8]
    static constexpr int smem_bytes(long hop) { return P::STAGE_BYTES + XCHG_BYTES + N * 4 + (int)((hop * 4 + 15) & ~15L) + 16; }

    // generic window-power sum of sample p (frames f_lo..f_hi in increasing order, src/stft.rs:146-148)
    static KD float norm_generic(const IstftFusedArgs &a, long p)
    {
        const long f_hi = (p / a.hop < a.nframes - 1) ? p / a.hop : a.nframes - 1;
        const long f_lo = p >= N ? (p - N) / a.hop + 1 : 0;
        float nrm = 0.0f;
        for (long f = f_lo; f <= f_hi; f++) {
            const float w = KOFFT_LDG(a.window + (p - f * a.hop));
            nrm = add_rn(nrm, mul_rn(w, w));
        }
        return nrm;
    }
    // normalise (or leave / zero) one finished sample and write it back (src/stft.rs:150-154, 335-341)
    static KD void emit(const IstftFusedArgs &a, long c, long p, float acc, float nrm)
    {
        if (nrm > 1e-8f)
            acc = div_rn(acc, nrm);
        else if (a.zero_uncovered)
            acc = 0.0f;
        a.output[c * a.out_len + p] = acc;
        if (a.norm) a.norm[c * a.out_len + p] = nrm;
    }

    static KD void run(const IstftFusedArgs &a, const Tw0 &tw0, const float2 *__restrict__ table, float2 *smem)
    {
        const int t = threadIdx.x;
        unsigned char *stage = reinterpret_cast<unsigned char *>(smem);
        float2 *xch = smem + P::STAGE_BYTES / 8;
        float2 *buf0 = xch;
        float2 *buf1 = ONEBUF ? buf0 : buf0 + P::PADN;
        float *ring = reinterpret_cast<float *>(xch + XCHG_BYTES / 8);
        float *normtab = ring + N;
        const long hop = a.hop;
        unsigned long long *mbar = reinterpret_cast<unsigned long long *>(normtab + ((hop + 3) & ~3L));
        unsigned phase = 0;
        if (t == 0) {
            mbar_init(mbar, 1);
            fence_mbar_init();
        }
        float2 tw1[P1::NTW], tw2[P2::NTW];
        P1::load_tw(table, t, tw1);
        P2::load_tw(table, t, tw2);
        float wv[EPT]; // window at this thread's output positions
#pragma unroll
        for (int w = 0; w < P2::R; w++) wv[w] = KOFFT_LDG(a.window + P2::dst_index(t, 0, w));
        // steady-state window power per residue, summed in frame order (earliest frame = largest offset first)
        for (long r = t; r < hop; r += CTA) {
            float nrm = 0.0f;
            for (long i = (N - 1 - r) / hop; i >= 0; i--) {
                const float w = KOFFT_LDG(a.window + r + i * hop);
                nrm = add_rn(nrm, mul_rn(w, w));
            }
            normtab[r] = nrm;
        }
        __syncthreads();
        // Steady-state fast path (hop == PRE * CTA, the common 4x overlap with one frame per CTA): every
        // thread finalises exactly PRE samples per frame, at fixed residues j = t + i CTA, so their
        // window power lives in registers and every address is "per-frame base + compile-time offset".
        constexpr int PRE = 4;
        const bool fast_shape = hop == (long)PRE * CTA;
        float nreg[PRE];
#pragma unroll
        for (int i = 0; i < PRE; i++) nreg[i] = fast_shape ? normtab[t + i * CTA] : 0.0f;

        long nfe = (a.out_len + hop - 1) / hop; // frames that can reach the output
        if (nfe > a.nframes) nfe = a.nframes;
        const long G = a.run_frames;
        const long runs_per_ch = (nfe + G - 1) / G;
        const long total = a.channels * runs_per_ch;
        const long halo = (N + hop - 1) / hop - 1; // == frames before f that still reach region f

        for (long run = blockIdx.x; run < total; run += gridDim.x) {
            const long c = run / runs_per_ch;
            const long F0 = (run - c * runs_per_ch) * G;
            const long Fend = F0 + G < nfe ? F0 + G : nfe;
            const long fs = F0 - halo > 0 ? F0 - halo : 0;
            const long own_lo = F0 * hop;
            const long own_hi = Fend == nfe ? a.out_len : Fend * hop;
            const float2 *fr_base = a.frames + c * a.nframes * N;
            const float *out0 = a.output + c * a.out_len;

            if (t == 0) { // first frame of the run
                mbar_expect_tx(mbar, (unsigned)(N * 8));
                bulk_copy_g2s(stage, fr_base + fs * N, (unsigned)(N * 8), mbar);
            }
            // ring <- initial output for the owned samples among [fs*hop, fs*hop + N)
            for (int j = t; j < N; j += CTA) {
                const long p = fs * hop + j;
                ring[p & (N - 1)] = (p >= own_lo && p < own_hi) ? out0[p] : 0.0f;
            }
            __syncthreads();

            // finalise region fr = [fr*hop, (fr+1)*hop) and recycle its slots for the samples N later.
            // pre: the recycled slots' initial values, loaded by the caller ahead of time (hop <= PRE*CTA)
            const bool use_pre = hop <= (long)PRE * CTA;
            auto recycled_init = [&](long p2) { return (p2 >= own_lo && p2 < own_hi) ? out0[p2] : 0.0f; };
            // fast(fr): region fr is owned, in steady state and lies wholly inside the owned samples.  The bounds are
            // frame indices computed once per run, so the per-frame test is two comparisons.
            const long fast_lo = fast_shape ? (F0 > halo ? F0 : halo) : 1;
            const long fast_hi = fast_shape ? own_hi / hop : 0; // (fr + 1) * hop <= own_hi
            auto fast = [&](long fr) { return fr >= fast_lo && fr < fast_hi; };
            // the slots recycled after frame fr get the samples [fr*hop + N, (fr+1)*hop + N): wholly inside the owned
            // samples for pre_in_lo <= fr < pre_in_hi, wholly beyond them for fr >= pre_out_lo
            const long pre_in_lo = fast_shape ? (own_lo > N ? (own_lo - N + hop - 1) / hop : 0) : 1;
            const long pre_in_hi = fast_shape ? (own_hi >= N + hop ? (own_hi - N - hop) / hop + 1 : 0) : 0;
            const long pre_out_lo = own_hi > N ? (own_hi - N + hop - 1) / hop : 0;
            // initial values of the slots recycled after frame fr (loaded one whole iteration before they are used)
            auto load_pre = [&](long fr, float *pre) {
                if (fr >= pre_in_lo && fr < pre_in_hi) {
                    const float *src = out0 + fr * hop + N + t;
#pragma unroll
                    for (int i = 0; i < PRE; i++) pre[i] = src[i * CTA];
                } else if (fast_shape && fr >= pre_out_lo) {
#pragma unroll
                    for (int i = 0; i < PRE; i++) pre[i] = 0.0f;
                } else {
#pragma unroll
                    for (int i = 0; i < PRE; i++) {
                        const long j = t + (long)i * CTA;
                        pre[i] = j < hop ? recycled_init(fr * hop + j + N) : 0.0f;
                    }
                }
            };
            auto finalize_region = [&](long fr, const float *pre) {
                const bool steady = fr >= halo; // every residue has its full set of covering frames
                if (fast(fr)) { // warp-uniform
                    const int rb = (int)((fr * hop) & (N - 1)) + t;
                    float *o = a.output + c * a.out_len + fr * hop + t;
                    float *no = a.norm ? a.norm + c * a.out_len + fr * hop + t : nullptr;
#pragma unroll
                    for (int i = 0; i < PRE; i++) {
                        float *r = ring + ((rb + i * CTA) & (N - 1));
                        float acc = *r;
                        const float nrm = nreg[i];
                        if (nrm > 1e-8f)
                            acc = div_rn(acc, nrm);
                        else if (a.zero_uncovered)
                            acc = 0.0f;
                        o[i * CTA] = acc;
                        if (no) no[i * CTA] = nrm;
                        *r = pre[i];
                    }
                } else if (use_pre) {
#pragma unroll
                    for (int i = 0; i < PRE; i++) {
                        const long j = t + (long)i * CTA;
                        if (j < hop) {
                            const long p = fr * hop + j;
                            float *r = ring + (p & (N - 1));
                            if (fr >= F0 && p < own_hi) emit(a, c, p, *r, steady ? normtab[j] : norm_generic(a, p));
                            *r = pre[i];
                        }
                    }
                } else {
                    for (long j = t; j < hop; j += CTA) {
                        const long p = fr * hop + j;
                        float *r = ring + (p & (N - 1));
                        if (fr >= F0 && p < own_hi) emit(a, c, p, *r, steady ? normtab[j] : norm_generic(a, p));
                        *r = recycled_init(p + N);
                    }
                }
            };

            int par = 0;
            float pre[PRE]; // for the region finalised in the current iteration
            if (KOFFT_ISTFT_PREPIPE && use_pre && fs + 1 < Fend) load_pre(fs, pre);
#pragma unroll(FRAME_UNROLL)
            for (long f = fs; f < Fend; f++) {
                if (!KOFFT_ISTFT_PREPIPE && use_pre && f > fs) load_pre(f - 1, pre);
                mbar_wait(mbar, phase);
                phase ^= 1;
                float2 x[EPT];
#pragma unroll
                for (int u = 0; u < P0::U; u++)
#pragma unroll
                    for (int q = 0; q < P0::R; q++)
                        x[u * P0::R + q] =
                            pre_conj<true>(reinterpret_cast<const float2 *>(stage)[P0::src_index(t, u, q)]);
                P0::compute(x, tw0.v);
                float2 *b = par ? buf1 : buf0;
                par ^= 1;
                H::template store_smem<P0>(b, t, x);
                __syncthreads(); // A
                if (t == 0 && f + 1 < Fend) { // stage consumed: prefetch the run's next frame
                    mbar_expect_tx(mbar, (unsigned)(N * 8));
                    bulk_copy_g2s(stage, fr_base + (f + 1) * N, (unsigned)(N * 8), mbar);
                }
                if (f > fs) finalize_region(f - 1, pre);
                if (KOFFT_ISTFT_PREPIPE && use_pre && f + 1 < Fend) load_pre(f, pre); // used after the next frame's first barrier
                H::template load_smem<P1>(b, t, x);
                if (ONEBUF) __syncthreads(); // everyone has read the buffer before the next exchange overwrites it
                P1::compute(x, tw1);
                b = par ? buf1 : buf0;
                par ^= 1;
                H::template store_smem<P1>(b, t, x);
                __syncthreads(); // B
                H::template load_smem<P2>(b, t, x);
                if (ONEBUF) __syncthreads();
                P2::template compute<KOFFT_ISTFT_RE_LAST != 0>(x, tw2); // only frame.re is used below
                // ordered overlap-add of frame f: ifft = conj, re*scale (src/fft.rs:1163-1172),
                // then frame.re * window (src/stft.rs:144)
                const int fb = (int)((f * hop) & (N - 1)) + P2::dst_base(t, 0); // slot of the thread's first sample
                if (KOFFT_ISTFT_HOIST) {
                    float acc[P2::R]; // the accumulator values first: one shared-memory round trip for all of them
#pragma unroll
                    for (int w = 0; w < P2::R; w++) acc[w] = ring[(fb + P2::dst_off(w)) & (N - 1)];
#pragma unroll
                    for (int w = 0; w < P2::R; w++) {
                        const float v = mul_rn(mul_rn(x[w].x, a.scale), wv[w]);
                        ring[(fb + P2::dst_off(w)) & (N - 1)] = add_rn(acc[w], v);
                    }
                } else {
#pragma unroll
                    for (int w = 0; w < P2::R; w++) {
                        const float v = mul_rn(mul_rn(x[w].x, a.scale), wv[w]);
                        float *r = ring + ((fb + P2::dst_off(w)) & (N - 1));
                        *r = add_rn(*r, v);
                    }
                }
            }
            __syncthreads();
            if (Fend > fs) {
                const long fr = Fend - 1;
                const bool steady = fr >= halo;
                for (long j = t; j < hop; j += CTA) {
                    const long p = fr * hop + j;
                    if (fr >= F0 && p < own_hi) emit(a, c, p, ring[p & (N - 1)], steady ? normtab[j] : norm_generic(a, p));
                }
            }
            // the channel's last run also owns everything after its last frame's hop
            if (Fend == nfe) {
                const long covered = nfe > 0 ? (nfe - 1) * hop + N : 0;
                for (long p = Fend * hop + t; p < a.out_len; p += CTA) {
                    if (p < covered) {
                        emit(a, c, p, ring[p & (N - 1)], norm_generic(a, p));
                    } else { // no frame reaches this sample
                        if (a.zero_uncovered) a.output[c * a.out_len + p] = 0.0f;
                        if (a.norm) a.norm[c * a.out_len + p] = 0.0f;
                    }
                }
            }
            __syncthreads(); // the ring is re-initialised by the next run
        }
    }
};

#ifdef __CUDACC__
template <int L, bool EXACT>
__global__ void __launch_bounds__((IstftFused<L, EXACT>::CTA), (IstftFused<L, EXACT>::MIN_BLOCKS))
    istft_fused_kernel(const __grid_constant__ IstftFusedArgs a, const __grid_constant__ Tw0 tw0,
                       const float2 *__restrict__ table)
{
    extern __shared__ __align__(128) float2 smem[];
    IstftFused<L, EXACT>::run(a, tw0, table, smem);
}
#endif

} // namespace kofft
