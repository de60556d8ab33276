// ola.cu -- istft stage 2: ordered overlap-add + normalisation.
//
// Reference (src/stft.rs:136-154): frames are visited in increasing order and each adds
//   output[start+i] += frame[i].re * window[i];  scratch[start+i] += window[i]*window[i]
// then output[i] /= scratch[i] where scratch[i] > 1e-8.  A gather over the (at most
// ceil(win_len/hop)) frames covering an output sample, visited in increasing frame order,
// performs exactly the same sequence of f32 additions per sample, so the result is
// bit-identical and deterministic without atomics.
#include "hostdev.h"
#include "launch.h"

namespace kofft {

__global__ void __launch_bounds__(256) ola_kernel(const OlaArgs a)
{
    const long total = a.channels * a.out_len;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long c = idx / a.out_len;
        const long p = idx - c * a.out_len;
        // frames f with f*hop <= p < f*hop + win_len
        long f_hi = p / a.hop;
        if (f_hi > a.nframes - 1) f_hi = a.nframes - 1;
        long f_lo = p >= a.win_len ? (p - a.win_len) / a.hop + 1 : 0;
        float acc = a.output[idx];
        float nrm = 0.0f;
        const float *tc = a.time + c * a.nframes * a.win_len;
        for (long f = f_lo; f <= f_hi; f++) {
            const long i = p - f * a.hop;
            const float w = __ldg(a.window + i);
            acc = add_rn(acc, __ldg(tc + f * a.win_len + i));
            nrm = add_rn(nrm, mul_rn(w, w));
        }
        if (nrm > 1e-8f)
            acc = div_rn(acc, nrm);
        else if (a.zero_uncovered)
            acc = 0.0f;
        a.output[idx] = acc;
        if (a.norm) a.norm[idx] = nrm;
    }
}

cudaError_t launch_ola(const OlaArgs &a, cudaStream_t stream)
{
    const long total = a.channels * a.out_len;
    if (total <= 0) return cudaSuccess;
    const int threads = 256;
    long blocks = (total + threads - 1) / threads;
    const long cap = 148L * 32;
    int grid = (int)(blocks < cap ? blocks : cap);
    ola_kernel<<<grid, threads, 0, stream>>>(a);
    return cudaGetLastError();
}

} // namespace kofft
