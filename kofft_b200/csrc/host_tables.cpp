// host_tables.cpp -- host-side generators for the tables the GPU kernels consume.
//
// kofft derives its twiddle tables with an f32 recurrence whose rounding errors (1e-5 ..
// 1e-3 relative by N = 4096 .. 32768) are far above the 1e-5 parity budget, so the GPU path
// must use *kofft's* tables, bit for bit, not mathematically exact ones.
// They are therefore produced here on the host with the same libm calls and the same
// operation order, then uploaded once per context ("FftPlanner's twiddle cache becomes
// device-resident twiddle tables").
//
// MUST be compiled with -ffp-contract=off: the only fused operations are the explicit fmaf
// calls that restate the reference's mul_add.
#include <cmath>
#include <cstddef>

#include "host_tables.h"

namespace kofft {

static const float kPi32 = 3.14159274101257324219f; // core::f32::consts::PI

// FftPlanner::get_twiddles, reference src/fft.rs:391-405
void host_fft_twiddles(size_t n, float *out)
{
    const size_t half = n / 2;
    const float angle = -2.0f * kPi32 / static_cast<float>(n);
    const float s = sinf(angle), c = cosf(angle); // f32::sin_cos -> libm
    float re = 1.0f, im = 0.0f;
    for (size_t k = 0; k < half; ++k) {
        out[2 * k] = re;
        out[2 * k + 1] = im;
        const float re_old = re;
        re = fmaf(re, c, -(im * s));
        im = fmaf(im, c, re_old * s);
    }
}

// FftPlanner<f64>::get_twiddles, the same lines for T = f64: angle = -from_f32(2.0) * PI / from_f32(n as f32),
// f64::sin_cos and f64::mul_add (libm sin / cos / fma)
void host_fft_twiddles_f64(size_t n, double *out)
{
    const size_t half = n / 2;
    const double pi = 3.14159265358979323846; // core::f64::consts::PI
    const double angle = -static_cast<double>(2.0f) * pi / static_cast<double>(static_cast<float>(n));
    const double s = sin(angle), c = cos(angle);
    double re = 1.0, im = 0.0;
    for (size_t k = 0; k < half; ++k) {
        out[2 * k] = re;
        out[2 * k + 1] = im;
        const double re_old = re;
        re = fma(re, c, -(im * s));
        im = fma(im, c, re_old * s);
    }
}

// Correctly rounded roots of unity exp(-2 pi i k stride / n), k < count, evaluated in f64.  NOT the
// reference's table: used only where the reference has no usable output (the multi-GPU
// transform of BASELINE configs[4]) and for the inter-step twiddles of that path.
void host_accurate_twiddles(size_t n, size_t stride, size_t count, float *out)
{
    const double step = -2.0 * 3.14159265358979323846 / static_cast<double>(n);
    for (size_t k = 0; k < count; ++k) {
        // reduce k*stride mod n exactly before scaling so large exponents keep full accuracy
        const size_t e = (k * stride) % n;
        // use the octant symmetry-free direct form: |step*e| <= 2 pi, f64 error ~1e-16
        out[2 * k] = static_cast<float>(cos(step * static_cast<double>(e)));
        out[2 * k + 1] = static_cast<float>(sin(step * static_cast<double>(e)));
    }
}

// FftPlanner::get_bluestein, reference src/fft.rs:411-428: chirp[i] = expi(-angle), b[i] = expi(angle)
// for i < n, b[m - i] = b[i], zero elsewhere; angle = pi * ((i*i) as f32) / (n as f32) in f32,
// expi = (cos, sin) through libm.  (The reference then transforms b with its own FFT: done on the
// device with the bit-exact kernels.)
void host_bluestein_chirp(size_t n, size_t m, float *chirp, float *b)
{
    for (size_t i = 0; i < 2 * m; ++i) b[i] = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        const float angle = kPi32 * static_cast<float>(i * i) / static_cast<float>(n);
        chirp[2 * i] = cosf(-angle);
        chirp[2 * i + 1] = sinf(-angle);
        b[2 * i] = cosf(angle);
        b[2 * i + 1] = sinf(angle);
    }
    for (size_t i = 1; i < n; ++i) {
        b[2 * (m - i)] = b[2 * i];
        b[2 * (m - i) + 1] = b[2 * i + 1];
    }
}

// the same for T = f64 (src/fft.rs:411-433): angle = T::pi() * T::from_f32((i*i) as f32) / T::from_f32(n as f32) -- the square
// passes through f32 before it is widened -- and expi = f64::sin_cos
void host_bluestein_chirp_f64(size_t n, size_t m, double *chirp, double *b)
{
    for (size_t i = 0; i < 2 * m; ++i) b[i] = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const double angle = 3.14159265358979323846 * static_cast<double>(static_cast<float>(i * i)) /
                             static_cast<double>(static_cast<float>(n));
        chirp[2 * i] = cos(-angle);
        chirp[2 * i + 1] = sin(-angle);
        b[2 * i] = cos(angle);
        b[2 * i + 1] = sin(angle);
    }
    for (size_t i = 1; i < n; ++i) {
        b[2 * (m - i)] = b[2 * i];
        b[2 * (m - i) + 1] = b[2 * i + 1];
    }
}

// build_twiddle_table, reference src/rfft.rs:172-183; `current = current.mul(w)` with
// Complex::mul unfused (src/num.rs:160-165) or fused under +fma (src/num.rs:173-178)
void host_rfft_twiddles(size_t m, float *out, bool fma_mul)
{
    const float angle = -kPi32 / static_cast<float>(m);
    const float ws = sinf(angle), wc = cosf(angle);
    float re = 1.0f, im = 0.0f;
    for (size_t k = 0; k < m; ++k) {
        out[2 * k] = re;
        out[2 * k + 1] = im;
        float nre, nim;
        if (fma_mul) {
            nre = fmaf(re, wc, -(im * ws));
            nim = fmaf(re, ws, im * wc);
        } else {
            nre = re * wc - im * ws;
            nim = re * ws + im * wc;
        }
        re = nre;
        im = nim;
    }
}

// build_twiddle_table::<f64>, reference src/rfft.rs:172-183: angle = -PI / from_f32(m as f32), current *= w with
// the unfused Complex::mul (src/num.rs:160-165)
void host_rfft_twiddles_f64(size_t m, double *out)
{
    const double angle = -3.14159265358979323846 / static_cast<double>(static_cast<float>(m));
    const double ws = sin(angle), wc = cos(angle);
    double re = 1.0, im = 0.0;
    for (size_t k = 0; k < m; ++k) {
        out[2 * k] = re;
        out[2 * k + 1] = im;
        const double nre = re * wc - im * ws;
        const double nim = re * ws + im * wc;
        re = nre;
        im = nim;
    }
}

// I0 series, reference src/window.rs:9-21
static float bessel_i0(float x)
{
    float sum = 1.0f;
    const float y = x * x / 4.0f;
    float t = y;
    float k = 1.0f;
    for (int n = 1; n < 20; ++n) {
        k *= static_cast<float>(n);
        sum += t / (k * k);
        t *= y;
    }
    return sum;
}

// hann / hamming / blackman / kaiser, reference src/window.rs:24-61 (periodic: divide by len)
int host_window(int kind, size_t len, float beta, float *out)
{
    switch (kind) {
    case 0:
        for (size_t i = 0; i < len; ++i)
            out[i] = 0.5f - 0.5f * cosf(2.0f * kPi32 * static_cast<float>(i) / static_cast<float>(len));
        return 0;
    case 1:
        for (size_t i = 0; i < len; ++i)
            out[i] = 0.54f - 0.46f * cosf(2.0f * kPi32 * static_cast<float>(i) / static_cast<float>(len));
        return 0;
    case 2:
        for (size_t i = 0; i < len; ++i) {
            const float x = static_cast<float>(i) / static_cast<float>(len);
            out[i] = 0.42f - 0.5f * cosf(2.0f * kPi32 * x) + 0.08f * cosf(4.0f * kPi32 * x);
        }
        return 0;
    case 3: {
        const float denom = bessel_i0(beta);
        const float mid = static_cast<float>(len - 1) / 2.0f;
        for (size_t i = 0; i < len; ++i) {
            const float r = (static_cast<float>(i) - mid) / mid;
            out[i] = bessel_i0(beta * sqrtf(1.0f - r * r)) / denom;
        }
        return 0;
    }
    default:
        return -1;
    }
}

} // namespace kofft
