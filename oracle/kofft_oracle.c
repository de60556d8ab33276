/*
 * kofft_oracle.c -- CPU restatement of okian/kofft's f32 FFT / rfft / STFT hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (libkofft_cuda.so, kofft_b200/) never links, imports or calls anything in oracle/.
 *
 * Why a restatement: the reference is a pure-Rust crate and this image has no
 * rustc/cargo, so the reference itself cannot be compiled or run here
 * (oracle/_ref does not exist for that reason).  Every function below follows the
 * cited reference lines operation by operation: same operand order, same points of
 * rounding, fmaf only where the reference calls mul_add.  It must be compiled with
 *     gcc -O2 -ffp-contract=off -fno-fast-math
 * so that the compiler neither fuses nor re-associates anything.
 *
 * Third-party arithmetic: f32::sin_cos / f32::cos / f32::mul_add in the reference are
 * Rust std wrappers over the system libm (sinf, cosf or sincosf, fmaf).  On this image
 * that is glibc 2.39 -- the same libm this file links -- and glibc's sinf, cosf and
 * sincosf share one implementation, so the table values agree.  libm::sqrtf (the Rust
 * `libm` crate, Cargo.toml `libm = "0.2"`, un-pinned: no Cargo.lock) is only used by
 * kaiser(); sqrt is correctly rounded in every conforming implementation.
 *
 * Parity pinning: checked against every value-pinning test the reference holds for
 * this path (restated in tests/test_oracle_golden.py).
 * Those tests pin tables only at n=8 index 1 and transforms only up to N=32 against
 * an independent DFT; above that the reference pins nothing but self-consistency, so
 * at BASELINE sizes the oracle's authority is the line-by-line restatement.
 *
 * Build flavour modelled: default cargo features, RUSTFLAGS empty (the flavour the
 * reference's published benchmarks use, benchmarks/README.md:13).  `fma_mul != 0`
 * selects the `-C target-feature=+fma` flavour of Complex::mul (src/num.rs:145-188)
 * where it matters (rfft table, src/rfft.rs:172-183).
 *
 * Layout: complex data is interleaved {re, im} f32 pairs == kofft's #[repr(C)]
 * Complex<f32> (src/num.rs:105-110).
 */
#include <math.h>
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KO_API __attribute__((visibility("default")))

/* FftError, src/fft.rs:446-454, in declaration order; 0 = Ok. */
enum {
    KO_OK = 0,
    KO_EMPTY_INPUT = 1,
    KO_NON_POW2_NO_STD = 2,
    KO_MISMATCHED_LENGTHS = 3,
    KO_INVALID_STRIDE = 4,
    KO_INVALID_HOP_SIZE = 5,
    KO_INVALID_VALUE = 6
};

typedef struct { float re, im; } c32;

/* core::f32::consts::PI */
static const float KO_PI32 = 3.14159274101257324219f;

static inline c32 c_new(float re, float im) { c32 r = { re, im }; return r; }
static inline c32 c_add(c32 a, c32 b) { return c_new(a.re + b.re, a.im + b.im); } /* num.rs:128-133 */
static inline c32 c_sub(c32 a, c32 b) { return c_new(a.re - b.re, a.im - b.im); } /* num.rs:136-141 */
/* Complex::mul, default build (no target_feature=fma): src/num.rs:160-165 */
static inline c32 c_mul(c32 a, c32 b)
{
    return c_new(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
/* Complex::mul under target_feature=fma: src/num.rs:173-178 */
static inline c32 c_mul_fma(c32 a, c32 b)
{
    return c_new(fmaf(a.re, b.re, -(a.im * b.im)), fmaf(a.re, b.im, a.im * b.re));
}

static int is_pow2(size_t n) { return n != 0 && (n & (n - 1)) == 0; }

/* ------------------------------------------------------------------------- */
/* Twiddle tables                                                            */
/* ------------------------------------------------------------------------- */

/* FftPlanner::get_twiddles, src/fft.rs:391-405.  out: n/2 complex. */
KO_API void kofft_oracle_twiddles_f32(size_t n, float *out)
{
    size_t half = n / 2;
    float angle = -2.0f * KO_PI32 / (float)n; /* (-2*pi)/n, left to right */
    float sin_step = sinf(angle), cos_step = cosf(angle);
    float w_re = 1.0f, w_im = 0.0f;
    for (size_t i = 0; i < half; i++) {
        out[2 * i] = w_re;
        out[2 * i + 1] = w_im;
        float tmp = w_re;
        w_re = fmaf(w_re, cos_step, -(w_im * sin_step)); /* :402 */
        w_im = fmaf(w_im, cos_step, tmp * sin_step);     /* :403 */
    }
}

/* build_twiddle_table, src/rfft.rs:172-183.  out: m complex, T'[k] = exp(-i pi k / m). */
KO_API void kofft_oracle_rfft_twiddles_f32(size_t m, float *out, int fma_mul)
{
    float angle = -KO_PI32 / (float)m;
    float sin_step = sinf(angle), cos_step = cosf(angle);
    c32 w = c_new(cos_step, sin_step);
    c32 cur = c_new(1.0f, 0.0f);
    for (size_t i = 0; i < m; i++) {
        out[2 * i] = cur.re;
        out[2 * i + 1] = cur.im;
        cur = fma_mul ? c_mul_fma(cur, w) : c_mul(cur, w);
    }
}

/* ------------------------------------------------------------------------- */
/* N <= 16 literal kernels, src/fft_kernels.rs                               */
/* ------------------------------------------------------------------------- */

static void k_fft2(c32 *x) /* :4-11 */
{
    c32 a = x[0], b = x[1];
    x[0] = c_add(a, b);
    x[1] = c_sub(a, b);
}

static void k_fft4(c32 *x) /* :13-30 */
{
    c32 a0 = x[0], a1 = x[1], a2 = x[2], a3 = x[3];
    c32 even0 = c_add(a0, a2), even1 = c_sub(a0, a2);
    c32 odd0 = c_add(a1, a3), odd1 = c_sub(a1, a3);
    c32 w1 = c_new(0.0f, -1.0f);
    c32 t1 = c_mul(odd1, w1);
    x[0] = c_add(even0, odd0);
    x[2] = c_sub(even0, odd0);
    x[1] = c_add(even1, t1);
    x[3] = c_sub(even1, t1);
}

static void k_fft8(c32 *x) /* :32-88 */
{
    c32 w1 = c_new(0.0f, -1.0f);
    float s = 0.70710677f;
    c32 x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3], x4 = x[4], x5 = x[5], x6 = x[6], x7 = x[7];
    c32 a0 = c_add(x0, x4), a1 = c_sub(x0, x4), a2 = c_add(x2, x6), a3 = c_sub(x2, x6);
    c32 t = c_mul(a3, w1);
    x[0] = c_add(a0, a2);
    x[2] = c_sub(a0, a2);
    x[1] = c_add(a1, t);
    x[3] = c_sub(a1, t);
    c32 b0 = c_add(x1, x5), b1 = c_sub(x1, x5), b2 = c_add(x3, x7), b3 = c_sub(x3, x7);
    t = c_mul(b3, w1);
    x[4] = c_add(b0, b2);
    x[6] = c_sub(b0, b2);
    x[5] = c_add(b1, t);
    x[7] = c_sub(b1, t);
    c32 t1 = c_mul(x[5], c_new(s, -s));
    c32 t2 = c_mul(x[6], w1);
    c32 t3 = c_mul(x[7], c_new(-s, -s));
    c32 o0 = x[4], e0 = x[0], e1 = x[1], e2 = x[2], e3 = x[3];
    x[0] = c_add(e0, o0);
    x[4] = c_sub(e0, o0);
    x[1] = c_add(e1, t1);
    x[5] = c_sub(e1, t1);
    x[2] = c_add(e2, t2);
    x[6] = c_sub(e2, t2);
    x[3] = c_add(e3, t3);
    x[7] = c_sub(e3, t3);
}

/* one "FFT8 of 8 inputs" half of fft16, writing out[0..8): src/fft_kernels.rs:114-147 / 149-182 */
static void k_fft16_half(c32 y0, c32 y1, c32 y2, c32 y3, c32 y4, c32 y5, c32 y6, c32 y7, c32 *out)
{
    /* inputs arrive in the order the reference names them: (x0,x8,x4,x12, x2,x10,x6,x14) */
    c32 w1 = c_new(0.0f, -1.0f);
    float s = 0.70710677f;
    c32 a0 = c_add(y0, y1), a1 = c_sub(y0, y1), a2 = c_add(y2, y3), a3 = c_sub(y2, y3);
    c32 t = c_mul(a3, w1);
    c32 ea0 = c_add(a0, a2), ea2 = c_sub(a0, a2), ea1 = c_add(a1, t), ea3 = c_sub(a1, t);
    c32 b0 = c_add(y4, y5), b1 = c_sub(y4, y5), b2 = c_add(y6, y7), b3 = c_sub(y6, y7);
    t = c_mul(b3, w1);
    c32 eb0 = c_add(b0, b2), eb2 = c_sub(b0, b2), eb1 = c_add(b1, t), eb3 = c_sub(b1, t);
    c32 t0 = eb0;
    c32 t1 = c_mul(eb1, c_new(s, -s));
    c32 t2 = c_mul(eb2, w1);
    c32 t3 = c_mul(eb3, c_new(-s, -s));
    out[0] = c_add(ea0, t0);
    out[1] = c_add(ea1, t1);
    out[2] = c_add(ea2, t2);
    out[3] = c_add(ea3, t3);
    out[4] = c_sub(ea0, t0);
    out[5] = c_sub(ea1, t1);
    out[6] = c_sub(ea2, t2);
    out[7] = c_sub(ea3, t3);
}

static void k_fft16(c32 *x) /* :90-224 */
{
    c32 in[16];
    memcpy(in, x, sizeof in);
    k_fft16_half(in[0], in[8], in[4], in[12], in[2], in[10], in[6], in[14], x);
    k_fft16_half(in[1], in[9], in[5], in[13], in[3], in[11], in[7], in[15], x + 8);
    float c1 = 0.9238795f, s1 = -0.38268343f, c2 = 0.70710677f, s2 = -0.70710677f;
    float c3 = 0.38268343f, s3 = -0.9238795f, c4 = 0.0f, s4 = -1.0f;
    c32 o[8], e[8];
    o[0] = x[8];
    o[1] = c_mul(x[9], c_new(c1, s1));
    o[2] = c_mul(x[10], c_new(c2, s2));
    o[3] = c_mul(x[11], c_new(c3, s3));
    o[4] = c_mul(x[12], c_new(c4, s4));
    o[5] = c_mul(x[13], c_new(-c3, s3));
    o[6] = c_mul(x[14], c_new(-c2, s2));
    o[7] = c_mul(x[15], c_new(-c1, s1));
    for (int i = 0; i < 8; i++) e[i] = x[i];
    for (int i = 0; i < 8; i++) {
        x[i] = c_add(e[i], o[i]);
        x[i + 8] = c_sub(e[i], o[i]);
    }
}

/* ------------------------------------------------------------------------- */
/* Radix-2 Stockham autosort, f32 SoA: fft_split_simd, src/fft.rs:789-912    */
/* ------------------------------------------------------------------------- */

/* The SSE body (:845-862) and the scalar tail (:881-893) perform the same unfused
 * mul/mul/sub, mul/mul/add, add, sub per lane, so one scalar loop restates both. */
static void stockham_soa(float *re, float *im, float *sre, float *sim, size_t n, const c32 *tw)
{
    float *src_re = re, *src_im = im, *dst_re = sre, *dst_im = sim;
    size_t n1 = 1, n2 = n;
    while (n1 < n) {
        n2 >>= 1;
        for (size_t k = 0; k < n1; k++) {
            c32 w = tw[k * n2];
            size_t even_base = 2 * k * n2, odd_base = even_base + n2;
            size_t dst0 = k * n2, dst1 = (k + n1) * n2;
            for (size_t j = 0; j < n2; j++) {
                float even_re = src_re[even_base + j], even_im = src_im[even_base + j];
                float odd_re = src_re[odd_base + j], odd_im = src_im[odd_base + j];
                float t_re = odd_re * w.re - odd_im * w.im;
                float t_im = odd_re * w.im + odd_im * w.re;
                dst_re[dst0 + j] = even_re + t_re;
                dst_im[dst0 + j] = even_im + t_im;
                dst_re[dst1 + j] = even_re - t_re;
                dst_im[dst1 + j] = even_im - t_im;
            }
        }
        float *t;
        t = src_re; src_re = dst_re; dst_re = t;
        t = src_im; src_im = dst_im; dst_im = t;
        n1 <<= 1;
    }
    if (src_re != re) { /* :899-904 */
        memcpy(re, src_re, n * sizeof(float));
        memcpy(im, src_im, n * sizeof(float));
    }
}

/* A planner: cached table + scratch for one size (one ScalarFftImpl reused, as in
 * kofft-bench/benches/bench_fft.rs:112-153). */
typedef struct {
    size_t n;
    c32 *tw;
    float *re, *im, *sre, *sim;
} ko_plan;

static void plan_free(ko_plan *p)
{
    free(p->tw); free(p->re); free(p->im); free(p->sre); free(p->sim);
    memset(p, 0, sizeof *p);
}

static int plan_ensure(ko_plan *p, size_t n)
{
    if (p->n == n) return 0;
    plan_free(p);
    p->tw = (c32 *)malloc((n / 2 + 1) * sizeof(c32));
    p->re = (float *)malloc(n * sizeof(float));
    p->im = (float *)malloc(n * sizeof(float));
    p->sre = (float *)malloc(n * sizeof(float));
    p->sim = (float *)malloc(n * sizeof(float));
    if (!p->tw || !p->re || !p->im || !p->sre || !p->sim) return -1;
    kofft_oracle_twiddles_f32(n, (float *)p->tw);
    p->n = n;
    return 0;
}

/* ScalarFftImpl::fft, src/fft.rs:1054-1082 + stockham_fft_with_threshold :642-706 */
static int fft_with_plan(ko_plan *p, c32 *x, size_t n);

/* Bluestein for non-power-of-two n (std builds): FftPlanner::get_bluestein src/fft.rs:411-433 and
 * ScalarFftImpl::fft src/fft.rs:1083-1132.  chirp[i] = expi(-angle), b[i] = expi(angle) with
 * angle = pi * ((i*i) as f32) / (n as f32) evaluated in f32 (unreduced, so large i are inexact --
 * deterministically, through the same libm sinf/cosf); b is transformed with a fresh
 * ScalarFftImpl (same algorithm); Complex::mul is the unfused form (src/num.rs:160-165). */
static int bluestein_with_plan(ko_plan *p, c32 *x, size_t n)
{
    size_t m = 1;
    while (m < 2 * n - 1) m <<= 1; /* (2n-1).next_power_of_two() */
    c32 *chirp = (c32 *)malloc(n * sizeof(c32));
    c32 *b = (c32 *)calloc(m, sizeof(c32));
    c32 *a = (c32 *)calloc(m, sizeof(c32));
    if (!chirp || !b || !a) { free(chirp); free(b); free(a); return -1; }
    for (size_t i = 0; i < n; i++) {
        float angle = KO_PI32 * (float)(i * i) / (float)n;
        chirp[i].re = cosf(-angle); chirp[i].im = sinf(-angle);
        b[i].re = cosf(angle); b[i].im = sinf(angle);
    }
    for (size_t i = 1; i < n; i++) b[m - i] = b[i];
    ko_plan fresh; memset(&fresh, 0, sizeof fresh);
    int rc = fft_with_plan(&fresh, b, m);
    plan_free(&fresh);
    if (!rc) {
        for (size_t i = 0; i < n; i++) a[i] = c_mul(x[i], chirp[i]);
        rc = fft_with_plan(p, a, m);
    }
    if (!rc) {
        for (size_t i = 0; i < m; i++) { a[i] = c_mul(a[i], b[i]); a[i].im = -a[i].im; }
        rc = fft_with_plan(p, a, m);
    }
    if (!rc) {
        float scale = 1.0f / (float)m;
        for (size_t i = 0; i < m; i++) { a[i].im = -a[i].im; a[i].re = a[i].re * scale; a[i].im = a[i].im * scale; }
        for (size_t i = 0; i < n; i++) x[i] = c_mul(a[i], chirp[i]);
    }
    free(chirp); free(b); free(a);
    return rc;
}

static int fft_with_plan(ko_plan *p, c32 *x, size_t n)
{
    if (n == 0) return KO_EMPTY_INPUT;
    if (n == 1) return KO_OK;
    if (n <= 16 && is_pow2(n)) {
        switch (n) {
        case 2: k_fft2(x); break;
        case 4: k_fft4(x); break;
        case 8: k_fft8(x); break;
        default: k_fft16(x); break;
        }
        return KO_OK;
    }
    if (!is_pow2(n)) return bluestein_with_plan(p, x, n); /* std build (src/fft.rs:1083-1132) */
    if (plan_ensure(p, n)) return -1;
    for (size_t i = 0; i < n; i++) { p->re[i] = x[i].re; p->im[i] = x[i].im; } /* :690-693 */
    stockham_soa(p->re, p->im, p->sre, p->sim, n, p->tw);
    for (size_t i = 0; i < n; i++) { x[i].re = p->re[i]; x[i].im = p->im[i]; } /* :695-698 */
    return KO_OK;
}

/* ScalarFftImpl::ifft, src/fft.rs:1134-1174 (serial branch; the rayon branch is the same math) */
static int ifft_with_plan(ko_plan *p, c32 *x, size_t n)
{
    if (n == 0) return KO_EMPTY_INPUT;
    if (n == 1) return KO_OK;
    for (size_t i = 0; i < n; i++) x[i].im = -x[i].im;
    int rc = fft_with_plan(p, x, n);
    if (rc) return rc;
    float scale = 1.0f / (float)n;
    for (size_t i = 0; i < n; i++) {
        x[i].im = -x[i].im;
        x[i].re = x[i].re * scale;
        x[i].im = x[i].im * scale;
    }
    return KO_OK;
}

KO_API int kofft_oracle_fft_f32(float *data, size_t n)
{
    ko_plan p; memset(&p, 0, sizeof p);
    int rc = fft_with_plan(&p, (c32 *)data, n);
    plan_free(&p);
    return rc;
}

KO_API int kofft_oracle_ifft_f32(float *data, size_t n)
{
    ko_plan p; memset(&p, 0, sizeof p);
    int rc = ifft_with_plan(&p, (c32 *)data, n);
    plan_free(&p);
    return rc;
}

/* ScalarFftImpl::fft_split / ifft_split (f32), src/fft.rs:1365-1439 */
KO_API int kofft_oracle_fft_split_f32(float *re, float *im, size_t n_re, size_t n_im)
{
    if (n_re != n_im) return KO_MISMATCHED_LENGTHS;
    size_t n = n_re;
    if (n == 0) return KO_EMPTY_INPUT;
    if (!is_pow2(n) || n <= 16) { /* fft_split_simd :797-810 routes through fft() */
        c32 *buf = (c32 *)malloc(n * sizeof(c32));
        for (size_t i = 0; i < n; i++) buf[i] = c_new(re[i], im[i]);
        int rc = kofft_oracle_fft_f32((float *)buf, n);
        if (!rc) for (size_t i = 0; i < n; i++) { re[i] = buf[i].re; im[i] = buf[i].im; }
        free(buf);
        return rc;
    }
    ko_plan p; memset(&p, 0, sizeof p);
    if (plan_ensure(&p, n)) return -1;
    stockham_soa(re, im, p.sre, p.sim, n, p.tw);
    plan_free(&p);
    return KO_OK;
}

KO_API int kofft_oracle_ifft_split_f32(float *re, float *im, size_t n_re, size_t n_im)
{
    if (n_re != n_im) return KO_MISMATCHED_LENGTHS;
    size_t n = n_re;
    for (size_t i = 0; i < n; i++) im[i] = -im[i];           /* :1409-1411 */
    int rc = kofft_oracle_fft_split_f32(re, im, n, n);
    if (rc) return rc;
    float scale = 1.0f / (float)n;                            /* :1413 */
    for (size_t i = 0; i < n; i++) {
        im[i] = -im[i];
        re[i] *= scale;
        im[i] *= scale;
    }
    return KO_OK;
}

/* fft_strided / ifft_strided, src/fft.rs:1175-1199, 1235-1259.
 * input_len = number of complex elements in `input`, n = scratch.len(). */
KO_API int kofft_oracle_fft_strided_f32(float *input, size_t input_len, size_t stride, size_t n, int inverse)
{
    if (stride == 0) return KO_INVALID_STRIDE;
    if (n == 0) return KO_OK;
    if (input_len < (n - 1) * stride + 1) return KO_MISMATCHED_LENGTHS;
    c32 *in = (c32 *)input;
    c32 *scratch = (c32 *)malloc(n * sizeof(c32));
    for (size_t i = 0; i < n; i++) scratch[i] = in[i * stride];
    int rc = inverse ? kofft_oracle_ifft_f32((float *)scratch, n) : kofft_oracle_fft_f32((float *)scratch, n);
    if (!rc) for (size_t i = 0; i < n; i++) in[i * stride] = scratch[i];
    free(scratch);
    return rc;
}

/* fft_out_of_place_strided / ifft_..., src/fft.rs:1260-1336 */
KO_API int kofft_oracle_fft_out_of_place_strided_f32(const float *input, size_t input_len, size_t in_stride,
                                                     float *output, size_t output_len, size_t out_stride,
                                                     int inverse)
{
    if (in_stride == 0 || out_stride == 0) return KO_INVALID_STRIDE;
    if (input_len % in_stride != 0 || output_len % out_stride != 0) return KO_INVALID_STRIDE;
    size_t n = input_len / in_stride;
    if (output_len / out_stride != n) return KO_MISMATCHED_LENGTHS;
    const c32 *in = (const c32 *)input;
    c32 *out = (c32 *)output;
    c32 *scratch = (c32 *)malloc((n ? n : 1) * sizeof(c32));
    for (size_t i = 0; i < n; i++) scratch[i] = in[i * in_stride];
    int rc = inverse ? kofft_oracle_ifft_f32((float *)scratch, n) : kofft_oracle_fft_f32((float *)scratch, n);
    if (!rc) for (size_t i = 0; i < n; i++) out[i * out_stride] = scratch[i];
    free(scratch);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* Real FFT: rfft_direct / irfft_direct, src/rfft.rs:425-508                 */
/* (rfft_direct_f32_avx :515-574 under feature x86_64 is the same unfused    */
/*  math lane 0 only, and + is commutative, so one restatement covers both.) */
/* ------------------------------------------------------------------------- */

typedef struct {
    ko_plan fft;
    size_t m;
    int fma_mul;
    c32 *tw; /* m entries */
    c32 *scratch;
} ko_rplan;

static void rplan_free(ko_rplan *p)
{
    plan_free(&p->fft);
    free(p->tw); free(p->scratch);
    memset(p, 0, sizeof *p);
}

static int rplan_ensure(ko_rplan *p, size_t m, int fma_mul)
{
    if (p->tw && p->m == m && p->fma_mul == fma_mul) return 0;
    free(p->tw); free(p->scratch);
    p->tw = (c32 *)malloc((m ? m : 1) * sizeof(c32));
    p->scratch = (c32 *)malloc((m ? m : 1) * sizeof(c32));
    if (!p->tw || !p->scratch) return -1;
    kofft_oracle_rfft_twiddles_f32(m, (float *)p->tw, fma_mul);
    p->m = m;
    p->fma_mul = fma_mul;
    return 0;
}

static int rfft_with_plan(ko_rplan *p, const float *input, size_t n, c32 *output, size_t out_len, int fma_mul)
{
    if (n == 0) return KO_EMPTY_INPUT;
    if (n % 2 != 0) return KO_INVALID_VALUE;
    size_t m = n / 2;
    if (out_len != m + 1) return KO_MISMATCHED_LENGTHS;
    if (rplan_ensure(p, m, fma_mul)) return -1;
    for (size_t i = 0; i < m; i++) output[i] = c_new(input[2 * i], input[2 * i + 1]); /* :444-446 */
    int rc = fft_with_plan(&p->fft, output, m);                                        /* :447 */
    if (rc) return rc;
    c32 *scratch = p->scratch;
    memcpy(scratch, output, m * sizeof(c32));                                          /* :449 */
    c32 y0 = scratch[0];
    output[0] = c_new(y0.re + y0.im, 0.0f);
    output[m] = c_new(y0.re - y0.im, 0.0f);
    float half = 0.5f;
    for (size_t k = 1; k < m; k++) { /* :454-463 */
        c32 a = scratch[k];
        c32 b = c_new(scratch[m - k].re, -scratch[m - k].im);
        c32 sum = c_add(a, b), diff = c_sub(a, b);
        c32 w = p->tw[k];
        /* default build: w.mul(diff) unfused.  With +fma AND without cargo feature x86_64
         * it would be fused; xtask's +fma build enables feature x86_64 (unfused AVX path),
         * so the twist is unfused in both modelled flavours. */
        c32 t = c_mul(w, diff);
        c32 temp = c_add(sum, c_new(t.im, -t.re));
        output[k] = c_new(temp.re * half, temp.im * half);
    }
    return KO_OK;
}

static int irfft_with_plan(ko_rplan *p, const c32 *input, size_t in_len, float *output, size_t n, int fma_mul)
{
    if (n == 0) return KO_EMPTY_INPUT;
    if (n % 2 != 0) return KO_INVALID_VALUE;
    size_t m = n / 2;
    if (in_len != m + 1) return KO_MISMATCHED_LENGTHS;
    if (rplan_ensure(p, m, fma_mul)) return -1;
    c32 *scratch = p->scratch;
    float half = 0.5f;
    scratch[0] = c_new((input[0].re + input[m].re) * half, (input[0].re - input[m].re) * half); /* :485-488 */
    for (size_t k = 1; k < m; k++) { /* :489-498 */
        c32 a = input[k];
        c32 b = c_new(input[m - k].re, -input[m - k].im);
        c32 sum = c_add(a, b), diff = c_sub(a, b);
        c32 w = c_new(p->tw[k].re, -p->tw[k].im);
        c32 t = c_mul(w, diff);
        c32 temp = c_sub(sum, c_new(t.im, -t.re));
        scratch[k] = c_new(temp.re * half, temp.im * half);
    }
    int rc = ifft_with_plan(&p->fft, scratch, m); /* :499 */
    if (rc) return rc;
    for (size_t i = 0; i < m; i++) { /* :500-503 */
        output[2 * i] = scratch[i].re;
        output[2 * i + 1] = scratch[i].im;
    }
    return KO_OK;
}

KO_API int kofft_oracle_rfft_f32(const float *input, size_t n, float *output, size_t out_len, int fma_mul)
{
    ko_rplan p; memset(&p, 0, sizeof p);
    int rc = rfft_with_plan(&p, input, n, (c32 *)output, out_len, fma_mul);
    rplan_free(&p);
    return rc;
}

KO_API int kofft_oracle_irfft_f32(const float *input, size_t in_len, float *output, size_t n, int fma_mul)
{
    ko_rplan p; memset(&p, 0, sizeof p);
    int rc = irfft_with_plan(&p, (const c32 *)input, in_len, output, n, fma_mul);
    rplan_free(&p);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* Windows, src/window.rs:9-61                                               */
/* ------------------------------------------------------------------------- */

KO_API void kofft_oracle_hann_f32(size_t len, float *out) /* :24-28 */
{
    for (size_t i = 0; i < len; i++)
        out[i] = 0.5f - 0.5f * cosf(2.0f * KO_PI32 * (float)i / (float)len);
}

KO_API void kofft_oracle_hamming_f32(size_t len, float *out) /* :31-35 */
{
    for (size_t i = 0; i < len; i++)
        out[i] = 0.54f - 0.46f * cosf(2.0f * KO_PI32 * (float)i / (float)len);
}

KO_API void kofft_oracle_blackman_f32(size_t len, float *out) /* :38-48 */
{
    for (size_t i = 0; i < len; i++) {
        float a0 = 0.42f, a1 = 0.5f, a2 = 0.08f;
        float x = (float)i / (float)len;
        out[i] = a0 - a1 * cosf(2.0f * KO_PI32 * x) + a2 * cosf(4.0f * KO_PI32 * x);
    }
}

static float bessel0(float x) /* :9-21 */
{
    float sum = 1.0f;
    float y = x * x / 4.0f;
    float t = y;
    float k = 1.0f;
    for (int n = 1; n < 20; n++) {
        k *= (float)n;
        sum += t / (k * k);
        t *= y;
    }
    return sum;
}

KO_API void kofft_oracle_kaiser_f32(size_t len, float beta, float *out) /* :52-61 */
{
    float denom = bessel0(beta);
    float m = (float)(len - 1) / 2.0f;
    for (size_t i = 0; i < len; i++) {
        float r = ((float)i - m) / m;
        out[i] = bessel0(beta * sqrtf(1.0f - r * r)) / denom;
    }
}

/* ------------------------------------------------------------------------- */
/* STFT / ISTFT, src/stft.rs:76-156, 289-343                                 */
/* ------------------------------------------------------------------------- */

static void stft_frame(ko_plan *p, const float *signal, size_t len, const float *window, size_t win_len,
                       size_t start, c32 *frame)
{
    for (size_t i = 0; i < win_len; i++) { /* :94-101 */
        float x = (start + i < len) ? signal[start + i] * window[i] : 0.0f;
        frame[i] = c_new(x, 0.0f);
    }
    (void)p;
}

/* stft(): frames is [nframes][win_len] complex, dense. */
KO_API int kofft_oracle_stft_f32(const float *signal, size_t len, const float *window, size_t win_len,
                                 size_t hop, float *frames, size_t nframes)
{
    if (hop == 0) return KO_INVALID_HOP_SIZE;
    size_t required = (len + hop - 1) / hop; /* div_ceil :86 */
    if (nframes < required) return KO_MISMATCHED_LENGTHS;
    ko_plan p; memset(&p, 0, sizeof p);
    int rc = KO_OK;
    for (size_t f = 0; f < nframes && !rc; f++) {
        c32 *frame = (c32 *)frames + f * win_len;
        stft_frame(&p, signal, len, window, win_len, f * hop, frame);
        rc = fft_with_plan(&p, frame, win_len); /* :102 */
    }
    plan_free(&p);
    return rc;
}

/* istft(): src/stft.rs:117-156.  frames are transformed in place (as the reference does);
 * `output` is accumulated into, `scratch` is zeroed first. */
KO_API int kofft_oracle_istft_f32(float *frames, size_t nframes, const float *window, size_t win_len,
                                  size_t hop, float *output, size_t out_len, float *scratch, size_t scratch_len)
{
    if (hop == 0) return KO_INVALID_HOP_SIZE;
    if (scratch_len != out_len) return KO_MISMATCHED_LENGTHS;
    for (size_t i = 0; i < scratch_len; i++) scratch[i] = 0.0f;
    ko_plan p; memset(&p, 0, sizeof p);
    int rc = KO_OK;
    for (size_t f = 0; f < nframes && !rc; f++) {
        size_t start = f * hop;
        c32 *frame = (c32 *)frames + f * win_len;
        rc = ifft_with_plan(&p, frame, win_len); /* :141 */
        if (rc) break;
        for (size_t i = 0; i < win_len; i++) { /* :142-147 */
            if (start + i < out_len) {
                output[start + i] += frame[i].re * window[i];
                scratch[start + i] += window[i] * window[i];
            }
        }
    }
    plan_free(&p);
    if (rc) return rc;
    for (size_t i = 0; i < out_len; i++) /* :150-154 */
        if (scratch[i] > 1e-8f) output[i] /= scratch[i];
    return KO_OK;
}

/* inverse_parallel(): src/stft.rs:289-343 -- frames not modified, own norm buffer,
 * output zeroed where norm <= 1e-8. */
KO_API int kofft_oracle_istft_parallel_f32(const float *frames, size_t nframes, const float *window,
                                           size_t win_len, size_t hop, float *output, size_t out_len)
{
    if (hop == 0) return KO_INVALID_HOP_SIZE;
    float *norm = (float *)calloc(out_len ? out_len : 1, sizeof(float));
    c32 *tb = (c32 *)malloc((win_len ? win_len : 1) * sizeof(c32));
    ko_plan p; memset(&p, 0, sizeof p);
    int rc = KO_OK;
    for (size_t f = 0; f < nframes && !rc; f++) {
        size_t start = f * hop;
        memcpy(tb, (const c32 *)frames + f * win_len, win_len * sizeof(c32));
        rc = ifft_with_plan(&p, tb, win_len);
        if (rc) break;
        for (size_t i = 0; i < win_len; i++) {
            if (start + i < out_len) {
                float acc = tb[i].re * window[i];
                float nf = window[i] * window[i];
                output[start + i] += acc;
                norm[start + i] += nf;
            }
        }
    }
    plan_free(&p);
    if (!rc)
        for (size_t i = 0; i < out_len; i++) {
            if (norm[i] > 1e-8f) output[i] /= norm[i];
            else output[i] = 0.0f;
        }
    free(norm); free(tb);
    return rc;
}

/* IstftStream::push_frame + flush, src/stft.rs:407-520, restated as a whole-stream helper:
 * pushes every frame, concatenates what push_frame returns, then appends flush().
 * out must hold nframes*hop + (win_len - hop) floats; returns number written via *written. */
KO_API int kofft_oracle_istft_stream_f32(const float *frames, size_t nframes, const float *window,
                                         size_t win_len, size_t hop, float *out, size_t *written)
{
    if (hop == 0) return KO_INVALID_HOP_SIZE;
    size_t buflen = win_len + hop * 2;
    float *buffer = (float *)calloc(buflen, sizeof(float));
    float *norm = (float *)calloc(buflen, sizeof(float));
    c32 *tb = (c32 *)malloc((win_len ? win_len : 1) * sizeof(c32));
    ko_plan p; memset(&p, 0, sizeof p);
    size_t buf_pos = 0, out_pos = 0, nout = 0;
    int rc = KO_OK;
    for (size_t f = 0; f < nframes; f++) {
        memcpy(tb, (const c32 *)frames + f * win_len, win_len * sizeof(c32));
        rc = ifft_with_plan(&p, tb, win_len);
        if (rc) break;
        for (size_t i = 0; i < win_len; i++) {
            float win = window[i];
            float val = tb[i].re * win;
            buffer[buf_pos + i] += val;
            norm[buf_pos + i] += win * win;
        }
        size_t out_start = out_pos, out_end = out_pos + hop;
        for (size_t i = out_start; i < out_end; i++) {
            if (norm[i] > 1e-8f) buffer[i] /= norm[i];
            norm[i] = 0.0f;
        }
        out_pos += hop;
        buf_pos += hop;
        if (buf_pos + win_len > buflen) {
            size_t nl = buf_pos + win_len;
            buffer = (float *)realloc(buffer, nl * sizeof(float));
            norm = (float *)realloc(norm, nl * sizeof(float));
            for (size_t i = buflen; i < nl; i++) { buffer[i] = 0.0f; norm[i] = 0.0f; }
            buflen = nl;
        }
        for (size_t i = 0; i < hop; i++) {
            size_t idx = buf_pos + win_len - hop + i;
            buffer[idx] = 0.0f;
            norm[idx] = 0.0f;
        }
        memcpy(out + nout, buffer + out_start, hop * sizeof(float));
        nout += hop;
    }
    if (!rc && nframes > 0) { /* flush :500-519 */
        size_t out_start = out_pos, out_end = buf_pos + win_len - hop;
        if (out_start < out_end) {
            for (size_t i = out_start; i < out_end; i++) {
                if (norm[i] > 1e-8f) buffer[i] /= norm[i];
                norm[i] = 0.0f;
            }
            memcpy(out + nout, buffer + out_start, (out_end - out_start) * sizeof(float));
            nout += out_end - out_start;
        }
    }
    *written = nout;
    plan_free(&p);
    free(buffer); free(norm); free(tb);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* f64 reference DFT for error measurement (not part of the reference).      */
/* ------------------------------------------------------------------------- */

/* out[k] for the listed bins only; in: n complex f32 interleaved; out: nbins complex f64. */
KO_API void kofft_oracle_dft_bins_f64(const float *in, size_t n, const size_t *bins, size_t nbins, double *out)
{
    const double two_pi = 6.283185307179586476925286766559;
    for (size_t b = 0; b < nbins; b++) {
        size_t k = bins[b];
        long double sr = 0.0L, si = 0.0L;
        for (size_t j = 0; j < n; j++) {
            size_t ph = (size_t)(((unsigned __int128)k * j) % n);
            double ang = -two_pi * (double)ph / (double)n;
            double c = cos(ang), s = sin(ang);
            double xr = in[2 * j], xi = in[2 * j + 1];
            sr += (long double)(xr * c - xi * s);
            si += (long double)(xr * s + xi * c);
        }
        out[2 * b] = (double)sr;
        out[2 * b + 1] = (double)si;
    }
}

/* ------------------------------------------------------------------------- */
/* Batched, threaded drivers: the timed CPU baseline ("kofft CPU path,        */
/* restated").  Rows / frames / channels are split evenly across threads,    */
/* the analogue of stft::parallel's par_iter_mut (src/stft.rs:246-262).      */
/* ------------------------------------------------------------------------- */

typedef struct {
    int kind; /* 0 c2c, 1 rfft, 2 stft, 3 irfft, 4 stft with a fresh planner per frame */
    int inverse, fma_mul;
    float *a; const float *b; float *c;
    size_t n, lo, hi;
    /* stft */
    const float *window; size_t win_len, hop, len, nframes;
    int rc;
} ko_job;

static void *job_main(void *arg)
{
    ko_job *j = (ko_job *)arg;
    j->rc = 0;
    if (j->kind == 0) {
        ko_plan p; memset(&p, 0, sizeof p);
        for (size_t r = j->lo; r < j->hi && !j->rc; r++) {
            c32 *row = (c32 *)j->a + r * j->n;
            j->rc = j->inverse ? ifft_with_plan(&p, row, j->n) : fft_with_plan(&p, row, j->n);
        }
        plan_free(&p);
    } else if (j->kind == 1) {
        ko_rplan p; memset(&p, 0, sizeof p);
        size_t m = j->n / 2;
        for (size_t r = j->lo; r < j->hi && !j->rc; r++)
            j->rc = rfft_with_plan(&p, j->b + r * j->n, j->n, (c32 *)j->c + r * (m + 1), m + 1, j->fma_mul);
        rplan_free(&p);
    } else if (j->kind == 3) {
        ko_rplan p; memset(&p, 0, sizeof p);
        size_t m = j->n / 2;
        for (size_t r = j->lo; r < j->hi && !j->rc; r++)
            j->rc = irfft_with_plan(&p, (const c32 *)j->b + r * (m + 1), m + 1, j->c + r * j->n, j->n, j->fma_mul);
        rplan_free(&p);
    } else { /* stft: [lo,hi) indexes channel*nframes + frame */
        ko_plan p; memset(&p, 0, sizeof p);
        for (size_t q = j->lo; q < j->hi && !j->rc; q++) {
            size_t ch = q / j->nframes, f = q % j->nframes;
            c32 *frame = (c32 *)j->c + q * j->win_len;
            if (j->kind == 4) plan_free(&p); /* ScalarFftImpl::default() per frame, src/stft.rs:260 */
            stft_frame(&p, j->b + ch * j->len, j->len, j->window, j->win_len, f * j->hop, frame);
            j->rc = fft_with_plan(&p, frame, j->win_len);
        }
        plan_free(&p);
    }
    return NULL;
}

static int run_jobs(ko_job proto, size_t total, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > total && total > 0) nthreads = (int)total;
    ko_job *jobs = (ko_job *)malloc(sizeof(ko_job) * (size_t)nthreads);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = proto;
        jobs[t].lo = total * (size_t)t / (size_t)nthreads;
        jobs[t].hi = total * (size_t)(t + 1) / (size_t)nthreads;
    }
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, job_main, &jobs[t]);
    job_main(&jobs[0]);
    int rc = jobs[0].rc;
    for (int t = 1; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        if (!rc) rc = jobs[t].rc;
    }
    free(jobs); free(th);
    return rc;
}

/* batch() / batch_inverse(), src/fft.rs:2156-2175, rows dense [batch][n]. */
KO_API int kofft_oracle_fft_batch_f32(float *data, size_t n, size_t batch, int inverse, int nthreads)
{
    ko_job j; memset(&j, 0, sizeof j);
    j.kind = 0; j.inverse = inverse; j.a = data; j.n = n;
    return run_jobs(j, batch, nthreads);
}

KO_API int kofft_oracle_rfft_batch_f32(const float *in, size_t n, size_t batch, float *out, int fma_mul, int nthreads)
{
    if (n == 0) return KO_EMPTY_INPUT;
    if (n % 2) return KO_INVALID_VALUE;
    ko_job j; memset(&j, 0, sizeof j);
    j.kind = 1; j.b = in; j.c = out; j.n = n; j.fma_mul = fma_mul;
    return run_jobs(j, batch, nthreads);
}

KO_API int kofft_oracle_irfft_batch_f32(const float *in, size_t n, size_t batch, float *out, int fma_mul, int nthreads)
{
    if (n == 0) return KO_EMPTY_INPUT;
    if (n % 2) return KO_INVALID_VALUE;
    ko_job j; memset(&j, 0, sizeof j);
    j.kind = 3; j.b = in; j.c = out; j.n = n; j.fma_mul = fma_mul;
    return run_jobs(j, batch, nthreads);
}

/* channels x stft(); signal [channels][len], frames [channels][nframes][win_len] complex.
 * fresh_planner != 0 re-derives the twiddle table per frame like stft::parallel does. */
KO_API int kofft_oracle_stft_batch_f32(const float *signal, size_t len, size_t channels, const float *window,
                                       size_t win_len, size_t hop, float *frames, size_t nframes,
                                       int fresh_planner, int nthreads)
{
    if (hop == 0) return KO_INVALID_HOP_SIZE;
    if (nframes < (len + hop - 1) / hop) return KO_MISMATCHED_LENGTHS;
    ko_job j; memset(&j, 0, sizeof j);
    j.kind = fresh_planner ? 4 : 2;
    j.b = signal; j.c = frames; j.len = len; j.window = window; j.win_len = win_len; j.hop = hop;
    j.nframes = nframes;
    return run_jobs(j, channels * nframes, nthreads);
}
