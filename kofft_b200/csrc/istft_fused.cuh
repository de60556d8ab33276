// istft_fused.cuh -- istft in ONE kernel: inverse FFT, x 1/N, x window, ordered overlap-add and
// normalisation fused behind the last FFT stage (reference: src/stft.rs:117-156; the
// `inverse_parallel` variant :289-343 via zero_uncovered).
//
// A CTA owns a run of G consecutive frames of one channel and the output samples
// [F0*hop, Fend*hop) they start (the channel's last run also owns the tail up to out_len).
// It walks the frames in increasing order -- first the H = ceil(N/hop)-1 halo frames before F0,
// whose tails reach into the owned samples (recomputed, not communicated) -- and adds each
// frame's windowed real part into an N-sample ring buffer in shared memory.  After frame f has
// been added no later frame touches samples [f*hop, (f+1)*hop), so they are normalised by the
// summed window power, stored, and their ring slots recycled (re-initialised from the caller's
// `output`, which the reference accumulates into).  Per sample this performs exactly the
// reference's sequence of f32 additions, in the same (frame) order, so the result is
// bit-identical and deterministic; no atomics, no intermediate frame buffer in HBM.
//
// HBM traffic per frame: 8N (frame) * (1 + H/G) + 4 hop (initial output) + 4 hop (result).
#pragma once
#include "fft_kernels.cuh"

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
    using P = Plan<L>;
    static_assert(P::NP == 3 && P::NBUF == 2, "fused istft covers N = 512 .. 4096");
    using IO = IoIstft;
    using H = CtaFft<P, EXACT, IO>;
    using P0 = Pass<P, 0, EXACT>;
    using P1 = Pass<P, 1, EXACT>;
    using P2 = Pass<P, 2, EXACT>;
    static constexpr int N = P::N;
    static constexpr int TPC = P::TPC;
    static constexpr int SMEM_BYTES = P::STAGE_BYTES + P::XCHG_BYTES + N * 4 + 16;

    // normalise (or leave / zero) one finished sample and write it back
    static KD void finalize(const IstftFusedArgs &a, long c, long p, float acc)
    {
        const long f_hi = (p / a.hop < a.nframes - 1) ? p / a.hop : a.nframes - 1;
        const long f_lo = p >= N ? (p - N) / a.hop + 1 : 0;
        float nrm = 0.0f;
        for (long f = f_lo; f <= f_hi; f++) {
            const float w = KOFFT_LDG(a.window + (p - f * a.hop));
            nrm = add_rn(nrm, mul_rn(w, w));
        }
        if (nrm > 1e-8f)
            acc = div_rn(acc, nrm);
        else if (a.zero_uncovered)
            acc = 0.0f;
        a.output[c * a.out_len + p] = acc;
        if (a.norm) a.norm[c * a.out_len + p] = nrm;
    }

    static KD void run(const IstftFusedArgs &a, const Tw0 &tw0, const float2 *__restrict__ table, float2 *smem)
    {
        const int tid = threadIdx.x;
        const int slot = tid / P::T;
        const int t = tid - slot * P::T;
        unsigned char *stage = reinterpret_cast<unsigned char *>(smem);
        float2 *xch = smem + P::STAGE_BYTES / 8;
        float2 *buf0 = xch + slot * P::PADN;
        float2 *buf1 = buf0 + TPC * P::PADN;
        float *ring = reinterpret_cast<float *>(xch + P::XCHG_BYTES / 8);
        unsigned long long *mbar = reinterpret_cast<unsigned long long *>(ring + N); // 8-byte aligned: N*4 % 8 == 0
        unsigned phase = 0;
        if (tid == 0) {
            mbar_init(mbar, 1);
            fence_mbar_init();
        }
        float2 tw1[P1::NTW], tw2[P2::NTW];
        P1::load_tw(table, t, tw1);
        P2::load_tw(table, t, tw2);
        float wv[EPT]; // window at this thread's output positions
#pragma unroll
        for (int w = 0; w < P2::R; w++) wv[w] = KOFFT_LDG(a.window + P2::dst_index(t, 0, w));
        __syncthreads();

        const long hop = a.hop;
        long nfe = (a.out_len + hop - 1) / hop; // frames that can reach the output
        if (nfe > a.nframes) nfe = a.nframes;
        const long G = a.run_frames;
        const long runs_per_ch = (nfe + G - 1) / G;
        const long total = a.channels * runs_per_ch;
        const long halo = (N + hop - 1) / hop - 1;
        int par = 0;

        for (long run = blockIdx.x; run < total; run += gridDim.x) {
            const long c = run / runs_per_ch;
            const long F0 = (run - c * runs_per_ch) * G;
            const long Fend = F0 + G < nfe ? F0 + G : nfe;
            const long fs = F0 - halo > 0 ? F0 - halo : 0;
            const long own_lo = F0 * hop;
            const long own_hi = Fend == nfe ? a.out_len : Fend * hop;
            const float2 *fr_base = a.frames + c * a.nframes * N;
            const float *out0 = a.output + c * a.out_len;

            // first group of the run
            if (tid == 0) {
                const long nf = Fend - fs < TPC ? Fend - fs : TPC;
                mbar_expect_tx(mbar, (unsigned)(nf * N * 8));
                bulk_copy_g2s(stage, fr_base + fs * N, (unsigned)(nf * N * 8), mbar);
            }
            // ring <- initial output for the owned samples among [fs*hop, fs*hop + N)
            for (int j = tid; j < N; j += P::CTA) {
                const long p = fs * hop + j;
                ring[p & (N - 1)] = (p >= own_lo && p < own_hi) ? out0[p] : 0.0f;
            }
            __syncthreads();

            for (long f = fs; f < Fend; f += TPC) {
                mbar_wait(mbar, phase);
                phase ^= 1;
                float2 x[EPT];
#pragma unroll
                for (int u = 0; u < P0::U; u++)
#pragma unroll
                    for (int q = 0; q < P0::R; q++)
                        x[u * P0::R + q] = pre_conj<true>(
                            reinterpret_cast<const float2 *>(stage)[slot * N + P0::src_index(t, u, q)]);
                P0::compute(x, tw0.v);
                float2 *b = par ? buf1 : buf0;
                par ^= 1;
                H::template store_smem<P0>(b, t, x);
                __syncthreads();
                if (tid == 0 && f + TPC < Fend) { // stage consumed: prefetch the run's next group
                    const long nf = Fend - (f + TPC) < TPC ? Fend - (f + TPC) : TPC;
                    mbar_expect_tx(mbar, (unsigned)(nf * N * 8));
                    bulk_copy_g2s(stage, fr_base + (f + TPC) * N, (unsigned)(nf * N * 8), mbar);
                }
                H::template load_smem<P1>(b, t, x);
                P1::compute(x, tw1);
                b = par ? buf1 : buf0;
                par ^= 1;
                H::template store_smem<P1>(b, t, x);
                __syncthreads();
                H::template load_smem<P2>(b, t, x);
                P2::compute(x, tw2);

                // ordered overlap-add: one frame of the group at a time
#pragma unroll 1
                for (int sl = 0; sl < TPC; sl++) {
                    const long fr = f + sl;
                    if (fr >= Fend) break;
                    if (slot == sl) {
#pragma unroll
                        for (int w = 0; w < P2::R; w++) {
                            const long p = fr * hop + P2::dst_index(t, 0, w);
                            // ifft: conj, re*scale (src/fft.rs:1163-1172); then frame.re * window (src/stft.rs:144)
                            const float v = mul_rn(mul_rn(x[w].x, a.scale), wv[w]);
                            float *r = ring + (p & (N - 1));
                            *r = add_rn(*r, v);
                        }
                    }
                    __syncthreads();
                    // samples [fr*hop, (fr+1)*hop) are complete: write them, recycle their slots
                    for (long j = tid; j < hop; j += P::CTA) {
                        const long p = fr * hop + j;
                        float *r = ring + (p & (N - 1));
                        if (fr >= F0 && p < own_hi) finalize(a, c, p, *r);
                        const long p2 = p + N;
                        *r = (p2 >= own_lo && p2 < own_hi) ? out0[p2] : 0.0f;
                    }
                    __syncthreads();
                }
            }
            // the channel's last run also owns everything after its last frame's hop
            if (Fend == nfe) {
                const long covered = nfe > 0 ? (nfe - 1) * hop + N : 0;
                for (long p = Fend * hop + tid; p < a.out_len; p += P::CTA) {
                    if (p < covered) {
                        finalize(a, c, p, ring[p & (N - 1)]);
                    } else { // no frame reaches this sample
                        if (a.zero_uncovered) a.output[c * a.out_len + p] = 0.0f;
                        if (a.norm) a.norm[c * a.out_len + p] = 0.0f;
                    }
                }
            }
            __syncthreads();
        }
    }
};

#ifdef __CUDACC__
template <int L, bool EXACT>
__global__ void __launch_bounds__(Plan<L>::CTA, 2)
    istft_fused_kernel(const __grid_constant__ IstftFusedArgs a, const __grid_constant__ Tw0 tw0,
                       const float2 *__restrict__ table)
{
    extern __shared__ __align__(128) float2 smem[];
    IstftFused<L, EXACT>::run(a, tw0, table, smem);
}
#endif

} // namespace kofft
