// hostdev.h -- lets the FFT engine (fft_engine.cuh) compile both under nvcc (the product)
// and under plain g++ (tests/emu: a CPU emulation of the CTA used to check index math and
// bit-exactness before spending GPU time).  The g++ side is test scaffolding only.
#pragma once

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define KHD __device__ __forceinline__
#define KD __device__ __forceinline__
#elif defined(KOFFT_EMU)
// whole-kernel CPU emulation (tests/emu/cuda_emu.h): CUDA threads are coroutines
#include "cuda_emu.h"
#define KHD inline
#define KD inline
#else
#include <cmath>
#define KHD inline
#define KD inline
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
// phase-by-phase host replay (tests/emu/emu_engine.cpp) is single-threaded
static inline int atomicMax(int *a, int v) { int old = *a; if (v > old) *a = v; return old; }
#endif

namespace kofft {

// ---- exactly-rounded scalar ops (never contracted) ------------------------------------------
#if defined(__CUDA_ARCH__)
KD float mul_rn(float a, float b) { return __fmul_rn(a, b); }
KD float add_rn(float a, float b) { return __fadd_rn(a, b); }
KD float sub_rn(float a, float b) { return __fsub_rn(a, b); }
KD float div_rn(float a, float b) { return __fdiv_rn(a, b); }
KD float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
KD float sqrt_rn(float a) { return __fsqrt_rn(a); }
KD int float_bits(float a) { return __float_as_int(a); }
#else
// host: translation units including this header are compiled with -ffp-contract=off
KHD float mul_rn(float a, float b) { return a * b; }
KHD float add_rn(float a, float b) { return a + b; }
KHD float sub_rn(float a, float b) { return a - b; }
KHD float div_rn(float a, float b) { return a / b; }
KHD float fma_rn(float a, float b, float c) { return fmaf(a, b, c); }
KHD float sqrt_rn(float a) { return sqrtf(a); }
KHD int float_bits(float a)
{
    int i;
    __builtin_memcpy(&i, &a, 4);
    return i;
}
#endif

// ---- packed f32x2 ops (sm_100a FADD2 / FMUL2 / FFMA2) -----------------------------------------
// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with explicit .rn, so in EXACT
// code a packed product is never fed to a packed add (scalar adds of FMUL2 halves are left
// alone); packed adds of non-product operands are safe everywhere.
#if defined(__CUDA_ARCH__)
KD float2 add2(float2 a, float2 b)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; "
        "mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
KD float2 sub2(float2 a, float2 b)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rc, ra, rb; "
        "mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
KD float2 mul2(float2 a, float2 b)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; "
        "mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
KD float2 fma2(float2 a, float2 b, float2 c)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; "
        "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
#else
KHD float2 add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
KHD float2 sub2(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
KHD float2 mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
KHD float2 fma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif

// ---- the radix-2 butterfly of kofft's Stockham stage (src/fft.rs:845-862 / 881-893) ---------
//   t = v * w ;  u' = u + t ;  v' = u - t
// Measured on B200 (scripts/microbench/fp32x2.cu, profiles/r01e_fp32x2.txt): a packed f32x2
// instruction occupies the FMA pipe for two cycles, i.e. packing saves issue slots, not pipe
// cycles.  What counts is lane-operations per butterfly:
// EXACT: every product and sum individually rounded, in the reference's operand order
//        -> bit-identical to the reference / oracle.  10 lane-ops: 2 FMUL2 (v.x*(w.x,w.y),
//        v.y*(w.y,w.x)) + 2 FADD + 2 FADD2 = 6 issue slots.  The products are combined with
//        SCALAR adds: ptxas contracts mul.f32x2 -> add.f32x2 into FFMA2 even with .rn, but it
//        leaves FMUL2 -> FADD alone (checked in SASS, and by the bit-exact GPU tests).
// FAST : 6 lane-ops: a = u + v*w as FFMA2 + 2 FFMA, then b = 2u - a as one FFMA2 (immediate 2,
//        negated addend) = 4 issue slots.  ~2e-7 rel-L2 from EXACT.
template <bool EXACT>
KHD void butterfly(float2 &u, float2 &v, const float2 w)
{
    if (EXACT) {
        float2 p = mul2(make_float2(v.x, v.x), w);
        float2 q = mul2(make_float2(v.y, v.y), make_float2(w.y, w.x));
        float2 t = make_float2(sub_rn(p.x, q.x), add_rn(p.y, q.y));
        float2 a = add2(u, t);
        v = sub2(u, t);
        u = a;
    } else {
        float2 a = fma2(make_float2(v.x, v.x), w, u);
        a.x = fma_rn(-v.y, w.y, a.x);
        a.y = fma_rn(v.y, w.x, a.y);
        v = fma2(u, make_float2(2.0f, 2.0f), make_float2(-a.x, -a.y));
        u = a;
    }
}

// the real parts only (the last stage of an inverse transform whose caller keeps frame.re, src/stft.rs:144): the same
// individually rounded products and sums as the .x components above; the .y components are left stale
template <bool EXACT>
KHD void butterfly_re(float2 &u, float2 &v, const float2 w)
{
    if (EXACT) {
        float2 p = mul2(v, w); // (v.re w.re, v.im w.im)
        float t = sub_rn(p.x, p.y);
        float a = add_rn(u.x, t);
        v.x = sub_rn(u.x, t);
        u.x = a;
    } else {
        float a = fma_rn(-v.y, w.y, fma_rn(v.x, w.x, u.x));
        v.x = fma_rn(u.x, 2.0f, -a);
        u.x = a;
    }
}

// twiddle == (1, 0) exactly (table entry 0): v*1 is the identity for finite v
KHD void butterfly_unit(float2 &u, float2 &v)
{
    float2 a = add2(u, v);
    v = sub2(u, v);
    u = a;
}

// Real-input shortcuts (stft frames have imag == +0 exactly, src/stft.rs:97-100).  While an element
// has only met unit twiddles it stays real, and the reference's operations on its zero imaginary
// part reduce to identities: v.y*w = +-0 drops out of v*w, and 0 +- t.y = +-t.y.  The results are
// the same f32 values (only the sign of an exact zero can differ, which no later operation of
// the transform can turn into a different value).
KHD void butterfly_unit_real(float2 &u, float2 &v)
{
    float a = add_rn(u.x, v.x);
    v.x = sub_rn(u.x, v.x);
    u.x = a;
}
template <bool EXACT>
KHD void butterfly_real(float2 &u, float2 &v, const float2 w)
{
    if (EXACT) {
        float2 t = mul2(make_float2(v.x, v.x), w);
        float ax = add_rn(u.x, t.x), bx = sub_rn(u.x, t.x);
        u = make_float2(ax, t.y);
        v = make_float2(bx, -t.y);
    } else {
        float ax = fma_rn(v.x, w.x, u.x);
        float ty = mul_rn(v.x, w.y);
        float bx = fma_rn(u.x, 2.0f, -ax);
        u = make_float2(ax, ty);
        v = make_float2(bx, -ty);
    }
}

// Complex::mul, unfused (src/num.rs:160-165), or contracted in FAST mode
template <bool EXACT>
KHD float2 cmul(const float2 a, const float2 b)
{
    if (EXACT)
        return make_float2(sub_rn(mul_rn(a.x, b.x), mul_rn(a.y, b.y)), add_rn(mul_rn(a.x, b.y), mul_rn(a.y, b.x)));
    return make_float2(fma_rn(a.x, b.x, -(a.y * b.y)), fma_rn(a.x, b.y, a.y * b.x));
}

// ---- f64 twin (ScalarFftImpl<f64>, src/fft.rs:914-1051): the same radix-2 butterfly in double, every
// product and sum rounded individually (the reference's f64 path has no fused form either).  There is
// no FAST variant: B200's FP64 pipe is not the limit of this HBM-bound kernel.
#if defined(__CUDA_ARCH__)
KD double dmul(double a, double b) { return __dmul_rn(a, b); }
KD double dadd(double a, double b) { return __dadd_rn(a, b); }
KD double dsub(double a, double b) { return __dsub_rn(a, b); }
#else
KHD double dmul(double a, double b) { return a * b; }
KHD double dadd(double a, double b) { return a + b; }
KHD double dsub(double a, double b) { return a - b; }
#endif
KHD double2 add2(double2 a, double2 b) { return make_double2(dadd(a.x, b.x), dadd(a.y, b.y)); }
KHD double2 sub2(double2 a, double2 b) { return make_double2(dsub(a.x, b.x), dsub(a.y, b.y)); }
// Complex::mul, unfused (src/num.rs:160-165)
template <bool EXACT>
KHD double2 cmul(const double2 a, const double2 b)
{
    return make_double2(dsub(dmul(a.x, b.x), dmul(a.y, b.y)), dadd(dmul(a.x, b.y), dmul(a.y, b.x)));
}
// t = v * w ;  u' = u + t ;  v' = u - t   (src/fft.rs:1019-1030)
KHD void butterfly_f64(double2 &u, double2 &v, const double2 w)
{
    const double2 t = make_double2(dsub(dmul(v.x, w.x), dmul(v.y, w.y)), dadd(dmul(v.x, w.y), dmul(v.y, w.x)));
    const double2 a = add2(u, t);
    v = sub2(u, t);
    u = a;
}
KHD void butterfly_unit_f64(double2 &u, double2 &v)
{
    const double2 a = add2(u, v);
    v = sub2(u, v);
    u = a;
}
// element constructors by type, so that the N <= 16 literal kernels are written once for both precisions;
// a float literal widened to double is exactly the reference's T::from_f32(literal)
template <class C2> struct Cx;
template <> struct Cx<float2> {
    typedef float real;
    static KHD float2 make(float x, float y) { return make_float2(x, y); }
};
template <> struct Cx<double2> {
    typedef double real;
    static KHD double2 make(double x, double y) { return make_double2(x, y); }
};

} // namespace kofft
