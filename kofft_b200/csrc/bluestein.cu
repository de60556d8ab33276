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

// ---- T = f64: the same three steps on double2 (every product and sum rounded individually) ----
__global__ void __launch_bounds__(256) blue_pre_kernel_d(const double2 *__restrict__ x, const double2 *__restrict__ chirp,
                                                         double2 *__restrict__ a, long n, long m, long rows, int inverse)
{
    const long total = rows * m;
    for (long idx = blockIdx.x * 256L + threadIdx.x; idx < total; idx += gridDim.x * 256L) {
        const long r = idx / m, i = idx - r * m;
        double2 v = make_double2(0.0, 0.0);
        if (i < n) {
            v = x[r * n + i];
            if (inverse) v.y = -v.y;
            v = cmul<true>(v, chirp[i]);
        }
        a[idx] = v;
    }
}
__global__ void __launch_bounds__(256) blue_mid_kernel_d(double2 *__restrict__ a, const double2 *__restrict__ bfft, long m, long rows)
{
    const long total = rows * m;
    for (long idx = blockIdx.x * 256L + threadIdx.x; idx < total; idx += gridDim.x * 256L) {
        double2 v = cmul<true>(a[idx], bfft[idx & (m - 1)]);
        v.y = -v.y;
        a[idx] = v;
    }
}
__global__ void __launch_bounds__(256) blue_post_kernel_d(const double2 *__restrict__ a, const double2 *__restrict__ chirp,
                                                          double2 *__restrict__ out, long n, long m, long rows, double scale_m,
                                                          int inverse, double scale_n)
{
    const long total = rows * n;
    for (long idx = blockIdx.x * 256L + threadIdx.x; idx < total; idx += gridDim.x * 256L) {
        const long r = idx / n, i = idx - r * n;
        double2 v = a[r * m + i];
        v.y = -v.y;
        v.x = dmul(v.x, scale_m);
        v.y = dmul(v.y, scale_m);
        v = cmul<true>(v, chirp[i]);
        if (inverse) {
            v.y = -v.y;
            v.x = dmul(v.x, scale_n);
            v.y = dmul(v.y, scale_n);
        }
        out[idx] = v;
    }
}

// f64 gather / scatter (src/fft.rs:1191-1197, 921-933; ifft_split :1414-1425), untwist (src/rfft.rs:485-498), twist (:450-463)
__global__ void __launch_bounds__(256) gather_kernel_d(const double *__restrict__ re, const double *__restrict__ im, long es, long rs,
                                                       double2 *__restrict__ a, long n, long rows, int neg_im)
{
    const long total = rows * n;
    for (long idx = blockIdx.x * 256L + threadIdx.x; idx < total; idx += gridDim.x * 256L) {
        const long r = idx / n, i = idx - r * n;
        const double y = im[r * rs + i * es];
        a[idx] = make_double2(re[r * rs + i * es], neg_im ? -y : y);
    }
}
__global__ void __launch_bounds__(256) scatter_kernel_d(const double2 *__restrict__ a, double *__restrict__ re, double *__restrict__ im,
                                                        long es, long rs, long n, long rows, int neg_im, double scale)
{
    const long total = rows * n;
    for (long idx = blockIdx.x * 256L + threadIdx.x; idx < total; idx += gridDim.x * 256L) {
        const long r = idx / n, i = idx - r * n;
        double2 v = a[idx];
        if (neg_im) { // *i = -*i; *r *= scale; *i *= scale
            v.y = -v.y;
            v.x = dmul(v.x, scale);
            v.y = dmul(v.y, scale);
        }
        re[r * rs + i * es] = v.x;
        im[r * rs + i * es] = v.y;
    }
}
__global__ void __launch_bounds__(256) untwist_kernel_d(const double2 *__restrict__ x, const double2 *__restrict__ rtw,
                                                        double2 *__restrict__ y, long m, long rows)
{
    const long total = rows * m;
    for (long idx = blockIdx.x * 256L + threadIdx.x; idx < total; idx += gridDim.x * 256L) {
        const long r = idx / m, k = idx - r * m;
        const double2 *X = x + r * (m + 1);
        double2 v;
        if (k == 0) {
            const double2 x0 = X[0], xm = X[m];
            v = make_double2(dmul(dadd(x0.x, xm.x), 0.5), dmul(dsub(x0.x, xm.x), 0.5));
        } else {
            const double2 a = X[k], q = X[m - k];
            const double2 b = make_double2(q.x, -q.y);
            const double2 sum = add2(a, b), diff = sub2(a, b);
            const double2 tw = rtw[k];
            const double2 t = cmul<true>(make_double2(tw.x, -tw.y), diff);
            v = make_double2(dmul(dsub(sum.x, t.y), 0.5), dmul(dadd(sum.y, t.x), 0.5));
        }
        y[idx] = v;
    }
}
__global__ void __launch_bounds__(256) twist_kernel_d(const double2 *__restrict__ y, const double2 *__restrict__ rtw,
                                                      double2 *__restrict__ out, long m, long rows)
{
    const long total = rows * (m + 1);
    for (long idx = blockIdx.x * 256L + threadIdx.x; idx < total; idx += gridDim.x * 256L) {
        const long r = idx / (m + 1), k = idx - r * (m + 1);
        const double2 *Y = y + r * m;
        double2 v;
        if (k == 0 || k == m) {
            const double2 a = Y[0];
            v = make_double2(k == 0 ? dadd(a.x, a.y) : dsub(a.x, a.y), 0.0);
        } else {
            const double2 a = Y[k], q = Y[m - k];
            const double2 b = make_double2(q.x, -q.y);
            const double2 sum = add2(a, b), diff = sub2(a, b);
            const double2 t = cmul<true>(rtw[k], diff);
            v = make_double2(dmul(dadd(sum.x, t.y), 0.5), dmul(dsub(sum.y, t.x), 0.5));
        }
        out[idx] = v;
    }
}

// ---- the element-wise steps around a non-power-of-two core in rfft / irfft / stft / istft / strided / split ----
// (the reference reaches Bluestein from all of them because they call fft.fft(): src/rfft.rs:447, 502,
// src/stft.rs:102, 141, src/fft.rs:797-809, 1191-1197)

// gather (src/fft.rs:1191-1193, 800-804): a[r*n + i] = (re[r*rs + i*es], im[r*rs + i*es])
__global__ void __launch_bounds__(256) gather_kernel(const float *__restrict__ re, const float *__restrict__ im, long es, long rs,
                                                     float2 *__restrict__ a, long n, long rows)
{
    const long total = rows * n;
    for (long idx = blockIdx.x * 256L + threadIdx.x; idx < total; idx += gridDim.x * 256L) {
        const long r = idx / n, i = idx - r * n;
        a[idx] = make_float2(re[r * rs + i * es], im[r * rs + i * es]);
    }
}
// scatter (src/fft.rs:1195-1197, 806-809)
__global__ void __launch_bounds__(256) scatter_kernel(const float2 *__restrict__ a, float *__restrict__ re, float *__restrict__ im,
                                                      long es, long rs, long n, long rows)
{
    const long total = rows * n;
    for (long idx = blockIdx.x * 256L + threadIdx.x; idx < total; idx += gridDim.x * 256L) {
        const long r = idx / n, i = idx - r * n;
        const float2 v = a[idx];
        re[r * rs + i * es] = v.x;
        im[r * rs + i * es] = v.y;
    }
}
// stft framing + windowing (src/stft.rs:91-101): frames[row][i] = (signal[c][f*hop + i] * w[i] or 0, 0)
__global__ void __launch_bounds__(256) frame_kernel(const float *__restrict__ signal, const float *__restrict__ window,
                                                    float2 *__restrict__ frames, long len, long nframes, long hop, long n, long rows)
{
    const long total = rows * n;
    for (long idx = blockIdx.x * 256L + threadIdx.x; idx < total; idx += gridDim.x * 256L) {
        const long row = idx / n, i = idx - row * n;
        const long c = row / nframes, f = row - c * nframes;
        const long pos = f * hop + i;
        float x = 0.0f;
        if (pos < len) x = mul_rn(signal[c * len + pos], __ldg(window + i));
        frames[idx] = make_float2(x, 0.0f);
    }
}
// istft: time[row][i] = ifft(frame)[i].re * w[i] (src/stft.rs:141-145); z already carries the 1/n of the ifft
__global__ void __launch_bounds__(256) time_kernel(const float2 *__restrict__ z, const float *__restrict__ window,
                                                   float *__restrict__ time, long n, long rows)
{
    const long total = rows * n;
    for (long idx = blockIdx.x * 256L + threadIdx.x; idx < total; idx += gridDim.x * 256L)
        time[idx] = mul_rn(z[idx].x, __ldg(window + (idx % n)));
}
// irfft untwist (src/rfft.rs:488-499): y[r][k] from X[r][0..m]
template <bool EXACT>
__global__ void __launch_bounds__(256) untwist_kernel(const float2 *__restrict__ x, const float2 *__restrict__ rtw,
                                                      float2 *__restrict__ y, long m, long rows)
{
    const long total = rows * m;
    for (long idx = blockIdx.x * 256L + threadIdx.x; idx < total; idx += gridDim.x * 256L) {
        const long r = idx / m, k = idx - r * m;
        const float2 *X = x + r * (m + 1);
        float2 v;
        if (k == 0) {
            const float2 x0 = X[0], xm = X[m];
            v = make_float2(mul_rn(add_rn(x0.x, xm.x), 0.5f), mul_rn(sub_rn(x0.x, xm.x), 0.5f));
        } else {
            const float2 a = X[k], q = X[m - k];
            const float2 b = make_float2(q.x, -q.y);
            const float2 sum = make_float2(add_rn(a.x, b.x), add_rn(a.y, b.y));
            const float2 diff = make_float2(sub_rn(a.x, b.x), sub_rn(a.y, b.y));
            const float2 tw = __ldg(rtw + k);
            const float2 t = cmul<EXACT>(make_float2(tw.x, -tw.y), diff);
            v = make_float2(mul_rn(sub_rn(sum.x, t.y), 0.5f), mul_rn(add_rn(sum.y, t.x), 0.5f));
        }
        y[idx] = v;
    }
}
// rfft twist (src/rfft.rs:450-463): out[r][k] from y = fft(m) of the packed row
template <bool EXACT>
__global__ void __launch_bounds__(256) twist_kernel(const float2 *__restrict__ y, const float2 *__restrict__ rtw,
                                                    float2 *__restrict__ out, long m, long rows)
{
    const long total = rows * (m + 1);
    for (long idx = blockIdx.x * 256L + threadIdx.x; idx < total; idx += gridDim.x * 256L) {
        const long r = idx / (m + 1), k = idx - r * (m + 1);
        const float2 *Y = y + r * m;
        float2 v;
        if (k == 0 || k == m) {
            const float2 a = Y[0];
            v = make_float2(k == 0 ? add_rn(a.x, a.y) : sub_rn(a.x, a.y), 0.0f);
        } else {
            const float2 a = Y[k], q = Y[m - k];
            const float2 b = make_float2(q.x, -q.y);
            const float2 sum = make_float2(add_rn(a.x, b.x), add_rn(a.y, b.y));
            const float2 diff = make_float2(sub_rn(a.x, b.x), sub_rn(a.y, b.y));
            const float2 t = cmul<EXACT>(__ldg(rtw + k), diff);
            v = make_float2(mul_rn(add_rn(sum.x, t.y), 0.5f), mul_rn(sub_rn(sum.y, t.x), 0.5f));
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

cudaError_t launch_bluestein_step_f64(int step, const BluesteinArgsD &b, int num_sms, cudaStream_t s)
{
    if (b.rows == 0) return cudaSuccess;
    switch (step) {
    case 0:
        blue_pre_kernel_d<<<grid_for(b.rows * b.m, num_sms), 256, 0, s>>>(b.x, b.chirp, b.a, b.n, b.m, b.rows, b.inverse);
        break;
    case 1:
        blue_mid_kernel_d<<<grid_for(b.rows * b.m, num_sms), 256, 0, s>>>(b.a, b.bfft, b.m, b.rows);
        break;
    default:
        blue_post_kernel_d<<<grid_for(b.rows * b.n, num_sms), 256, 0, s>>>(b.a, b.chirp, b.out, b.n, b.m, b.rows, b.scale_m, b.inverse,
                                                                         b.scale_n);
        break;
    }
    return cudaGetLastError();
}

// stft_magnitudes on frames of any length (src/visual/spectrogram.rs:60-73): mags[row][i] = |frame[row][i]| for i < n/2,
// and the running maximum of all of them (NaN never becomes the maximum)
__global__ void __launch_bounds__(256) mag_kernel(const float2 *__restrict__ frames, float *__restrict__ mags, int *max_bits, long n,
                                                  long rows)
{
    const long half = n >> 1, total = rows * half;
    float tmax = 0.0f;
    for (long idx = blockIdx.x * 256L + threadIdx.x; idx < total; idx += gridDim.x * 256L) {
        const long r = idx / half, i = idx - r * half;
        const float2 v = frames[r * n + i];
        const float mag = sqrt_rn(add_rn(mul_rn(v.x, v.x), mul_rn(v.y, v.y)));
        mags[idx] = mag;
        if (mag > tmax) tmax = mag;
    }
    if (tmax > 0.0f) atomicMax(max_bits, float_bits(tmax));
}

cudaError_t launch_elementwise_f64(const ElementwiseArgsD &e, int num_sms, cudaStream_t s)
{
    if (e.rows == 0 || e.n == 0) return cudaSuccess;
    switch (e.op) {
    case EW_GATHER:
        gather_kernel_d<<<grid_for(e.rows * e.n, num_sms), 256, 0, s>>>(e.re, e.im, e.es, e.rs, e.a, e.n, e.rows, e.neg_im);
        break;
    case EW_SCATTER:
        scatter_kernel_d<<<grid_for(e.rows * e.n, num_sms), 256, 0, s>>>(e.a, e.out_re, e.out_im, e.es, e.rs, e.n, e.rows, e.neg_im, e.scale);
        break;
    case EW_UNTWIST:
        untwist_kernel_d<<<grid_for(e.rows * e.n, num_sms), 256, 0, s>>>(e.x, e.rtw, e.a, e.n, e.rows);
        break;
    case EW_TWIST:
        twist_kernel_d<<<grid_for(e.rows * (e.n + 1), num_sms), 256, 0, s>>>(e.x, e.rtw, e.a, e.n, e.rows);
        break;
    default:
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_elementwise(const ElementwiseArgs &e, bool exact, int num_sms, cudaStream_t s)
{
    if (e.rows == 0 || e.n == 0) return cudaSuccess;
    const int g = grid_for(e.rows * (e.op == EW_TWIST ? e.n + 1 : e.n), num_sms);
    switch (e.op) {
    case EW_GATHER:
        gather_kernel<<<g, 256, 0, s>>>(e.re, e.im, e.es, e.rs, e.a, e.n, e.rows);
        break;
    case EW_SCATTER:
        scatter_kernel<<<g, 256, 0, s>>>(e.a, e.out_re, e.out_im, e.es, e.rs, e.n, e.rows);
        break;
    case EW_FRAME:
        frame_kernel<<<g, 256, 0, s>>>(e.re, e.aux_f, e.a, e.len, e.nframes, e.hop, e.n, e.rows);
        break;
    case EW_TIME:
        time_kernel<<<g, 256, 0, s>>>(e.a, e.aux_f, e.out_re, e.n, e.rows);
        break;
    case EW_MAG:
        if (e.n < 2) return cudaSuccess;
        mag_kernel<<<g, 256, 0, s>>>(e.x, e.out_re, e.max_bits, e.n, e.rows);
        break;
    case EW_UNTWIST:
        if (exact)
            untwist_kernel<true><<<g, 256, 0, s>>>(e.x, e.rtw, e.a, e.n, e.rows);
        else
            untwist_kernel<false><<<g, 256, 0, s>>>(e.x, e.rtw, e.a, e.n, e.rows);
        break;
    case EW_TWIST:
        if (exact)
            twist_kernel<true><<<g, 256, 0, s>>>(e.x, e.rtw, e.a, e.n, e.rows);
        else
            twist_kernel<false><<<g, 256, 0, s>>>(e.x, e.rtw, e.a, e.n, e.rows);
        break;
    default:
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

} // namespace kofft
