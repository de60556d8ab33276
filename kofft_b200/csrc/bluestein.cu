// bluestein.cu -- the element-wise steps of kofft's Bluestein path for non-power-of-two lengths
// (reference: ScalarFftImpl::fft src/fft.rs:1083-1132, FftPlanner::get_bluestein :411-433).  The
// two length-m transforms in between are the ordinary power-of-two kernels, so the arithmetic
// (and in EXACT mode every bit) matches the reference:
//   pre : a[i] = x[i] * chirp[i] (i < n), 0 (n <= i < m)                       :1099-1104
//   mid : a[i] = conj(a[i] * fft_b[i])                                         :1106-1111
//   post: a[i] = conj(a[i]) * (1/m); out[i] = a[i] * chirp[i] (i < n)          :1113-1123
// ifft (src/fft.rs:1163-1172) wraps fft in conj / conj * (1/n): folded into pre and post.
#include "hostdev.h"
#include "launch.h"

namespace kofft {

namespace {

template <bool EXACT>
__global__ void __launch_bounds__(256) blue_pre_kernel(const float2 *__restrict__ x, const float2 *__restrict__ chirp,
                                                       float2 *__restrict__ a, long n, long m, long rows, int inverse)
{
    const long total = rows * m;
    for (long idx = blockIdx.x * 256L + threadIdx.x; idx < total; idx += gridDim.x * 256L) {
        const long r = idx / m, i = idx - r * m;
        float2 v = make_float2(0.0f, 0.0f);
        if (i < n) {
            v = x[r * n + i];
            if (inverse) v.y = -v.y;
            v = cmul<EXACT>(v, __ldg(chirp + i));
        }
        a[idx] = v;
    }
}

template <bool EXACT>
__global__ void __launch_bounds__(256) blue_mid_kernel(float2 *__restrict__ a, const float2 *__restrict__ bfft, long m,
                                                       long rows)
{
    const long total = rows * m;
    for (long idx = blockIdx.x * 256L + threadIdx.x; idx < total; idx += gridDim.x * 256L) {
        const long i = idx & (m - 1);
        float2 v = cmul<EXACT>(a[idx], __ldg(bfft + i));
        v.y = -v.y;
        a[idx] = v;
    }
}

template <bool EXACT>
__global__ void __launch_bounds__(256) blue_post_kernel(const float2 *__restrict__ a, const float2 *__restrict__ chirp,
                                                        float2 *__restrict__ out, long n, long m, long rows,
                                                        float scale_m, int inverse, float scale_n)
{
    const long total = rows * n;
    for (long idx = blockIdx.x * 256L + threadIdx.x; idx < total; idx += gridDim.x * 256L) {
        const long r = idx / n, i = idx - r * n;
        float2 v = a[r * m + i];
        v.y = -v.y;
        v.x = mul_rn(v.x, scale_m);
        v.y = mul_rn(v.y, scale_m);
        v = cmul<EXACT>(v, __ldg(chirp + i));
        if (inverse) {
            v.y = -v.y;
            v.x = mul_rn(v.x, scale_n);
            v.y = mul_rn(v.y, scale_n);
        }
        out[idx] = v;
    }
}

int grid_for(long total, int num_sms)
{
    long g = (total + 255) / 256;
    const long cap = (long)num_sms * 16;
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

} // namespace

cudaError_t launch_bluestein_step(int step, const BluesteinArgs &b, bool exact, int num_sms, cudaStream_t s)
{
    if (b.rows == 0) return cudaSuccess;
    switch (step) {
    case 0:
        if (exact)
            blue_pre_kernel<true><<<grid_for(b.rows * b.m, num_sms), 256, 0, s>>>(b.x, b.chirp, b.a, b.n, b.m, b.rows, b.inverse);
        else
            blue_pre_kernel<false><<<grid_for(b.rows * b.m, num_sms), 256, 0, s>>>(b.x, b.chirp, b.a, b.n, b.m, b.rows, b.inverse);
        break;
    case 1:
        if (exact)
            blue_mid_kernel<true><<<grid_for(b.rows * b.m, num_sms), 256, 0, s>>>(b.a, b.bfft, b.m, b.rows);
        else
            blue_mid_kernel<false><<<grid_for(b.rows * b.m, num_sms), 256, 0, s>>>(b.a, b.bfft, b.m, b.rows);
        break;
    default:
        if (exact)
            blue_post_kernel<true><<<grid_for(b.rows * b.n, num_sms), 256, 0, s>>>(b.a, b.chirp, b.out, b.n, b.m, b.rows, b.scale_m,
                                                                                  b.inverse, b.scale_n);
        else
            blue_post_kernel<false><<<grid_for(b.rows * b.n, num_sms), 256, 0, s>>>(b.a, b.chirp, b.out, b.n, b.m, b.rows, b.scale_m,
                                                                                   b.inverse, b.scale_n);
        break;
    }
    return cudaGetLastError();
}

} // namespace kofft
