/* kofft_oracle_f64.c -- CPU restatement of kofft's f64 FFT path (TEST INFRASTRUCTURE, like
 * kofft_oracle.c: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load it).
 *
 * ScalarFftImpl<f64>: the generic dispatch src/fft.rs:1054-1082, the literal kernels
 * src/fft_kernels.rs:4-224 instantiated for T = f64 (constants are T::from_f32(<f32 literal>)),
 * stockham_fft_with_threshold's f64 branch :707-740, fft_split_simd for f64 :914-1051 (the AVX /
 * NEON bodies :972-1016 and the scalar tail :1017-1030 are the same unfused mul/mul/sub,
 * mul/mul/add, add, sub per lane), ifft :1134-1174 and FftPlanner::get_twiddles :391-405 with
 * T = f64 (f64::sin_cos, f64::mul_add -> libm sin / cos / fma).  Compiled with -ffp-contract=off.
 *
 * Parity status: pinned by the reference's own f64 tests (tests/split64.rs: split == AoS, ifft
 * round trip; src/lib.rs Complex64 checks) restated in tests/test_oracle_golden.py, and by the
 * line-by-line restatement; not by reference execution (no Rust toolchain in this image).
 * The N <= 16 kernels below are the f32 restatement of kofft_oracle.c retyped, nothing else. */
#include <math.h>
#include <pthread.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define KO_API __attribute__((visibility("default")))
enum { KO_OK = 0, KO_EMPTY_INPUT = 1, KO_NON_POW2_NO_STD = 2 };

typedef struct { double re, im; } c64;
static inline c64 z_new(double re, double im) { c64 r = { re, im }; return r; }
static inline c64 z_add(c64 a, c64 b) { return z_new(a.re + b.re, a.im + b.im); } /* num.rs:128-133 */
static inline c64 z_sub(c64 a, c64 b) { return z_new(a.re - b.re, a.im - b.im); } /* num.rs:136-141 */
/* Complex::mul, default build (no target_feature=fma): src/num.rs:160-165 */
static inline c64 z_mul(c64 a, c64 b) { return z_new(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
static int is_pow2(size_t n) { return n != 0 && (n & (n - 1)) == 0; }

/* FftPlanner::<f64>::get_twiddles, src/fft.rs:391-405.  out: n/2 complex. */
KO_API void kofft_oracle_twiddles_f64(size_t n, double *out)
{
    size_t half = n / 2;
    double angle = -(double)2.0f * 3.14159265358979323846 / (double)(float)n; /* -from_f32(2.0) * PI / from_f32(n as f32) */
    double sin_step = sin(angle), cos_step = cos(angle);
    double w_re = 1.0, w_im = 0.0;
    for (size_t i = 0; i < half; i++) {
        out[2 * i] = w_re;
        out[2 * i + 1] = w_im;
        double tmp = w_re;
        w_re = fma(w_re, cos_step, -(w_im * sin_step)); /* :402 */
        w_im = fma(w_im, cos_step, tmp * sin_step);     /* :403 */
    }
}

/* ---- N <= 16 literal kernels, src/fft_kernels.rs, T = f64 ---------------------------------- */

static void kd_fft2(c64 *x) /* :4-11 */
{
    c64 a = x[0], b = x[1];
    x[0] = z_add(a, b);
    x[1] = z_sub(a, b);
}

static void kd_fft4(c64 *x) /* :13-30 */
{
    c64 a0 = x[0], a1 = x[1], a2 = x[2], a3 = x[3];
    c64 even0 = z_add(a0, a2), even1 = z_sub(a0, a2);
    c64 odd0 = z_add(a1, a3), odd1 = z_sub(a1, a3);
    c64 w1 = z_new(0.0, -1.0);
    c64 t1 = z_mul(odd1, w1);
    x[0] = z_add(even0, odd0);
    x[2] = z_sub(even0, odd0);
    x[1] = z_add(even1, t1);
    x[3] = z_sub(even1, t1);
}

static void kd_fft8(c64 *x) /* :32-88 */
{
    c64 w1 = z_new(0.0, -1.0);
    double s = (double)0.70710677f; /* T::from_f32(0.70710677) */
    c64 x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3], x4 = x[4], x5 = x[5], x6 = x[6], x7 = x[7];
    c64 a0 = z_add(x0, x4), a1 = z_sub(x0, x4), a2 = z_add(x2, x6), a3 = z_sub(x2, x6);
    c64 t = z_mul(a3, w1);
    x[0] = z_add(a0, a2);
    x[2] = z_sub(a0, a2);
    x[1] = z_add(a1, t);
    x[3] = z_sub(a1, t);
    c64 b0 = z_add(x1, x5), b1 = z_sub(x1, x5), b2 = z_add(x3, x7), b3 = z_sub(x3, x7);
    t = z_mul(b3, w1);
    x[4] = z_add(b0, b2);
    x[6] = z_sub(b0, b2);
    x[5] = z_add(b1, t);
    x[7] = z_sub(b1, t);
    c64 t1 = z_mul(x[5], z_new(s, -s));
    c64 t2 = z_mul(x[6], w1);
    c64 t3 = z_mul(x[7], z_new(-s, -s));
    c64 o0 = x[4], e0 = x[0], e1 = x[1], e2 = x[2], e3 = x[3];
    x[0] = z_add(e0, o0);
    x[4] = z_sub(e0, o0);
    x[1] = z_add(e1, t1);
    x[5] = z_sub(e1, t1);
    x[2] = z_add(e2, t2);
    x[6] = z_sub(e2, t2);
    x[3] = z_add(e3, t3);
    x[7] = z_sub(e3, t3);
}

/* one "FFT8 of 8 inputs" half of fft16, writing out[0..8): src/fft_kernels.rs:114-147 / 149-182 */
static void kd_fft16_half(c64 y0, c64 y1, c64 y2, c64 y3, c64 y4, c64 y5, c64 y6, c64 y7, c64 *out)
{
    /* inputs arrive in the order the reference names them: (x0,x8,x4,x12, x2,x10,x6,x14) */
    c64 w1 = z_new(0.0, -1.0);
    double s = (double)0.70710677f; /* T::from_f32(0.70710677) */
    c64 a0 = z_add(y0, y1), a1 = z_sub(y0, y1), a2 = z_add(y2, y3), a3 = z_sub(y2, y3);
    c64 t = z_mul(a3, w1);
    c64 ea0 = z_add(a0, a2), ea2 = z_sub(a0, a2), ea1 = z_add(a1, t), ea3 = z_sub(a1, t);
    c64 b0 = z_add(y4, y5), b1 = z_sub(y4, y5), b2 = z_add(y6, y7), b3 = z_sub(y6, y7);
    t = z_mul(b3, w1);
    c64 eb0 = z_add(b0, b2), eb2 = z_sub(b0, b2), eb1 = z_add(b1, t), eb3 = z_sub(b1, t);
    c64 t0 = eb0;
    c64 t1 = z_mul(eb1, z_new(s, -s));
    c64 t2 = z_mul(eb2, w1);
    c64 t3 = z_mul(eb3, z_new(-s, -s));
    out[0] = z_add(ea0, t0);
    out[1] = z_add(ea1, t1);
    out[2] = z_add(ea2, t2);
    out[3] = z_add(ea3, t3);
    out[4] = z_sub(ea0, t0);
    out[5] = z_sub(ea1, t1);
    out[6] = z_sub(ea2, t2);
    out[7] = z_sub(ea3, t3);
}

static void kd_fft16(c64 *x) /* :90-224 */
{
    c64 in[16];
    memcpy(in, x, sizeof in);
    kd_fft16_half(in[0], in[8], in[4], in[12], in[2], in[10], in[6], in[14], x);
    kd_fft16_half(in[1], in[9], in[5], in[13], in[3], in[11], in[7], in[15], x + 8);
    double c1 = (double)0.9238795f, s1 = -(double)0.38268343f, c2 = (double)0.70710677f, s2 = -(double)0.70710677f;
    double c3 = (double)0.38268343f, s3 = -(double)0.9238795f, c4 = 0.0, s4 = -1.0;
    c64 o[8], e[8];
    o[0] = x[8];
    o[1] = z_mul(x[9], z_new(c1, s1));
    o[2] = z_mul(x[10], z_new(c2, s2));
    o[3] = z_mul(x[11], z_new(c3, s3));
    o[4] = z_mul(x[12], z_new(c4, s4));
    o[5] = z_mul(x[13], z_new(-c3, s3));
    o[6] = z_mul(x[14], z_new(-c2, s2));
    o[7] = z_mul(x[15], z_new(-c1, s1));
    for (int i = 0; i < 8; i++) e[i] = x[i];
    for (int i = 0; i < 8; i++) {
        x[i] = z_add(e[i], o[i]);
        x[i + 8] = z_sub(e[i], o[i]);
    }
}

/* ---- radix-2 Stockham autosort, f64 SoA: fft_split_simd, src/fft.rs:914-1051 ---------------- */
static void stockham_soa_f64(double *re, double *im, double *sre, double *sim, size_t n, const c64 *tw)
{
    double *src_re = re, *src_im = im, *dst_re = sre, *dst_im = sim;
    size_t n1 = 1, n2 = n;
    while (n1 < n) { /* :962-1036 */
        n2 >>= 1;
        for (size_t k = 0; k < n1; k++) {
            c64 w = tw[k * n2];
            size_t even_base = 2 * k * n2, odd_base = even_base + n2;
            size_t dst0 = k * n2, dst1 = (k + n1) * n2;
            for (size_t j = 0; j < n2; j++) {
                double even_re = src_re[even_base + j], even_im = src_im[even_base + j];
                double odd_re = src_re[odd_base + j], odd_im = src_im[odd_base + j];
                double t_re = odd_re * w.re - odd_im * w.im;
                double t_im = odd_re * w.im + odd_im * w.re;
                dst_re[dst0 + j] = even_re + t_re;
                dst_im[dst0 + j] = even_im + t_im;
                dst_re[dst1 + j] = even_re - t_re;
                dst_im[dst1 + j] = even_im - t_im;
            }
        }
        double *t;
        t = src_re; src_re = dst_re; dst_re = t;
        t = src_im; src_im = dst_im; dst_im = t;
        n1 <<= 1;
    }
    if (src_re != re) { /* :1037-1042 */
        memcpy(re, src_re, n * sizeof(double));
        memcpy(im, src_im, n * sizeof(double));
    }
}

typedef struct {
    size_t n;
    c64 *tw;
    double *re, *im, *sre, *sim;
} kd_plan;

static void kd_plan_free(kd_plan *p)
{
    free(p->tw); free(p->re); free(p->im); free(p->sre); free(p->sim);
    memset(p, 0, sizeof *p);
}

static int kd_plan_ensure(kd_plan *p, size_t n)
{
    if (p->n == n) return 0;
    kd_plan_free(p);
    p->tw = (c64 *)malloc((n / 2 + 1) * sizeof(c64));
    p->re = (double *)malloc(n * sizeof(double));
    p->im = (double *)malloc(n * sizeof(double));
    p->sre = (double *)malloc(n * sizeof(double));
    p->sim = (double *)malloc(n * sizeof(double));
    if (!p->tw || !p->re || !p->im || !p->sre || !p->sim) return -1;
    kofft_oracle_twiddles_f64(n, (double *)p->tw);
    p->n = n;
    return 0;
}

static int kd_fft(kd_plan *p, c64 *x, size_t n);
static void kd_plan_free(kd_plan *p);

/* Bluestein for non-power-of-two n with T = f64 (std builds): FftPlanner::get_bluestein src/fft.rs:411-433 and
 * ScalarFftImpl::fft src/fft.rs:1083-1132.  angle = T::pi() * T::from_f32((i*i) as f32) / T::from_f32(n as f32): the
 * square goes through f32 (inexact above 2^24) before it is widened; expi = f64::sin_cos -> libm sin / cos. */
static int kd_bluestein(kd_plan *p, c64 *x, size_t n)
{
    size_t m = 1;
    while (m < 2 * n - 1) m <<= 1; /* (2n-1).next_power_of_two() */
    c64 *chirp = (c64 *)malloc(n * sizeof(c64));
    c64 *b = (c64 *)calloc(m, sizeof(c64));
    c64 *a = (c64 *)calloc(m, sizeof(c64));
    if (!chirp || !b || !a) { free(chirp); free(b); free(a); return -1; }
    for (size_t i = 0; i < n; i++) {
        double angle = 3.14159265358979323846 * (double)(float)(i * i) / (double)(float)n;
        chirp[i] = z_new(cos(-angle), sin(-angle));
        b[i] = z_new(cos(angle), sin(angle));
    }
    for (size_t i = 1; i < n; i++) b[m - i] = b[i];
    kd_plan fresh; memset(&fresh, 0, sizeof fresh);
    int rc = kd_fft(&fresh, b, m);
    kd_plan_free(&fresh);
    if (!rc) {
        for (size_t i = 0; i < n; i++) a[i] = z_mul(x[i], chirp[i]);
        rc = kd_fft(p, a, m);
    }
    if (!rc) {
        for (size_t i = 0; i < m; i++) { a[i] = z_mul(a[i], b[i]); a[i].im = -a[i].im; }
        rc = kd_fft(p, a, m);
    }
    if (!rc) {
        double scale = 1.0 / (double)(float)m; /* T::one() / T::from_f32(m as f32) */
        for (size_t i = 0; i < m; i++) { a[i].im = -a[i].im; a[i].re = a[i].re * scale; a[i].im = a[i].im * scale; }
        for (size_t i = 0; i < n; i++) x[i] = z_mul(a[i], chirp[i]);
    }
    free(chirp); free(b); free(a);
    return rc;
}

/* ScalarFftImpl::<f64>::fft, src/fft.rs:1054-1082 + stockham_fft_with_threshold :642-740; Bluestein for the other
 * lengths (std build, :1083-1132). */
static int kd_fft(kd_plan *p, c64 *x, size_t n)
{
    if (n == 0) return KO_EMPTY_INPUT;
    if (n == 1) return KO_OK;
    if (!is_pow2(n)) return kd_bluestein(p, x, n);
    if (n <= 16) {
        switch (n) {
        case 2: kd_fft2(x); break;
        case 4: kd_fft4(x); break;
        case 8: kd_fft8(x); break;
        default: kd_fft16(x); break;
        }
        return KO_OK;
    }
    if (kd_plan_ensure(p, n)) return -1;
    for (size_t i = 0; i < n; i++) { p->re[i] = x[i].re; p->im[i] = x[i].im; } /* :714-717 */
    stockham_soa_f64(p->re, p->im, p->sre, p->sim, n, p->tw);
    for (size_t i = 0; i < n; i++) { x[i].re = p->re[i]; x[i].im = p->im[i]; } /* :719-722 */
    return KO_OK;
}

/* ScalarFftImpl::<f64>::ifft, src/fft.rs:1134-1174 (the parallel branch is f32-only) */
static int kd_ifft(kd_plan *p, c64 *x, size_t n)
{
    if (n == 0) return KO_EMPTY_INPUT;
    if (n == 1) return KO_OK;
    for (size_t i = 0; i < n; i++) x[i].im = -x[i].im;
    int rc = kd_fft(p, x, n);
    if (rc) return rc;
    double scale = 1.0 / (double)(float)n; /* T::one() / T::from_f32(n as f32) */
    for (size_t i = 0; i < n; i++) {
        x[i].im = -x[i].im;
        x[i].re = x[i].re * scale;
        x[i].im = x[i].im * scale;
    }
    return KO_OK;
}

KO_API int kofft_oracle_fft_f64(double *data, size_t n, int inverse)
{
    kd_plan p; memset(&p, 0, sizeof p);
    int rc = inverse ? kd_ifft(&p, (c64 *)data, n) : kd_fft(&p, (c64 *)data, n);
    kd_plan_free(&p);
    return rc;
}

/* batch() / batch_inverse() (src/fft.rs:2156-2175) over dense rows, rows split evenly over threads,
 * one planner per thread: the timed CPU baseline of the f64 bench line */
typedef struct { double *data; size_t n, r0, r1; int inverse, rc; } kd_job;
static void *kd_job_main(void *arg)
{
    kd_job *j = (kd_job *)arg;
    kd_plan p; memset(&p, 0, sizeof p);
    for (size_t r = j->r0; r < j->r1 && !j->rc; r++) {
        c64 *x = (c64 *)j->data + r * j->n;
        j->rc = j->inverse ? kd_ifft(&p, x, j->n) : kd_fft(&p, x, j->n);
    }
    kd_plan_free(&p);
    return NULL;
}

KO_API int kofft_oracle_fft_batch_f64(double *data, size_t n, size_t batch, int inverse, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > batch) nthreads = batch ? (int)batch : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    kd_job *jobs = (kd_job *)malloc(sizeof(kd_job) * (size_t)nthreads);
    if (!th || !jobs) { free(th); free(jobs); return -1; }
    for (int t = 0; t < nthreads; t++) {
        jobs[t].data = data; jobs[t].n = n; jobs[t].inverse = inverse; jobs[t].rc = 0;
        jobs[t].r0 = batch * (size_t)t / (size_t)nthreads;
        jobs[t].r1 = batch * (size_t)(t + 1) / (size_t)nthreads;
        if (t + 1 < nthreads) pthread_create(&th[t], NULL, kd_job_main, &jobs[t]);
    }
    kd_job_main(&jobs[nthreads - 1]);
    int rc = jobs[nthreads - 1].rc;
    for (int t = 0; t + 1 < nthreads; t++) { pthread_join(th[t], NULL); if (jobs[t].rc) rc = jobs[t].rc; }
    free(th); free(jobs);
    return rc;
}

/* ---- real transforms for f64: src/rfft.rs:172-183 (table), 425-465 (rfft_direct), 468-508 (irfft_direct) ---- */

/* build_twiddle_table::<f64>.  out: m complex, T'[k] = exp(-i pi k / m) by current = current.mul(w). */
KO_API void kofft_oracle_rfft_twiddles_f64(size_t m, double *out)
{
    double angle = -3.14159265358979323846 / (double)(float)m; /* -T::pi() / T::from_f32(m as f32) */
    c64 w = z_new(cos(angle), sin(angle));
    c64 cur = z_new(1.0, 0.0);
    for (size_t k = 0; k < m; k++) {
        out[2 * k] = cur.re;
        out[2 * k + 1] = cur.im;
        cur = z_mul(cur, w);
    }
}

static int kd_rfft(kd_plan *p, const double *input, size_t n, c64 *output, const c64 *tw)
{
    if (n == 0) return KO_EMPTY_INPUT;
    if (n % 2 != 0) return 6; /* InvalidValue */
    size_t m = n / 2;
    c64 *scratch = (c64 *)malloc((m ? m : 1) * sizeof(c64));
    if (!scratch) return -1;
    for (size_t i = 0; i < m; i++) output[i] = z_new(input[2 * i], input[2 * i + 1]); /* :444-446 */
    int rc = kd_fft(p, output, m);
    if (rc) { free(scratch); return rc; }
    memcpy(scratch, output, m * sizeof(c64));
    c64 y0 = scratch[0];
    output[0] = z_new(y0.re + y0.im, 0.0);
    output[m] = z_new(y0.re - y0.im, 0.0);
    double half = (double)0.5f;
    for (size_t k = 1; k < m; k++) { /* :454-462 */
        c64 a = scratch[k];
        c64 b = z_new(scratch[m - k].re, -scratch[m - k].im);
        c64 sum = z_add(a, b), diff = z_sub(a, b);
        c64 t = z_mul(tw[k], diff);
        c64 temp = z_add(sum, z_new(t.im, -t.re));
        output[k] = z_new(temp.re * half, temp.im * half);
    }
    free(scratch);
    return KO_OK;
}

static int kd_irfft(kd_plan *p, const c64 *input, double *output, size_t n, const c64 *tw)
{
    if (n == 0) return KO_EMPTY_INPUT;
    if (n % 2 != 0) return 6;
    size_t m = n / 2;
    c64 *scratch = (c64 *)malloc((m ? m : 1) * sizeof(c64));
    if (!scratch) return -1;
    double half = (double)0.5f;
    scratch[0] = z_new((input[0].re + input[m].re) * half, (input[0].re - input[m].re) * half); /* :486-489 */
    for (size_t k = 1; k < m; k++) { /* :490-498 */
        c64 a = input[k];
        c64 b = z_new(input[m - k].re, -input[m - k].im);
        c64 sum = z_add(a, b), diff = z_sub(a, b);
        c64 w = z_new(tw[k].re, -tw[k].im);
        c64 t = z_mul(w, diff);
        c64 temp = z_sub(sum, z_new(t.im, -t.re));
        scratch[k] = z_new(temp.re * half, temp.im * half);
    }
    int rc = kd_ifft(p, scratch, m);
    if (rc) { free(scratch); return rc; }
    for (size_t i = 0; i < m; i++) { output[2 * i] = scratch[i].re; output[2 * i + 1] = scratch[i].im; }
    free(scratch);
    return KO_OK;
}

/* rows dense: in [batch][n] doubles -> out [batch][n/2+1] complex (inverse = 0), or the reverse (inverse = 1) */
KO_API int kofft_oracle_rfft_batch_f64(const double *in, size_t n, size_t batch, double *out, int inverse)
{
    if (n == 0) return KO_EMPTY_INPUT;
    if (n % 2 != 0) return 6;
    size_t m = n / 2;
    c64 *tw = (c64 *)malloc((m ? m : 1) * sizeof(c64));
    if (!tw) return -1;
    kofft_oracle_rfft_twiddles_f64(m, (double *)tw);
    kd_plan p; memset(&p, 0, sizeof p);
    int rc = KO_OK;
    for (size_t r = 0; r < batch && !rc; r++) {
        if (inverse)
            rc = kd_irfft(&p, (const c64 *)in + r * (m + 1), out + r * n, n, tw);
        else
            rc = kd_rfft(&p, in + r * n, n, (c64 *)out + r * (m + 1), tw);
    }
    kd_plan_free(&p);
    free(tw);
    return rc;
}
