// small_kernels.cuh -- N = 1, 2, 4, 8, 16: kofft's unrolled literal-constant kernels
// (reference: src/fft_kernels.rs:4-224, dispatched from src/fft.rs:1062-1071 when the whole
// transform has N <= 16).  One thread per transform; not a throughput path, present so the
// FftImpl surface is complete and bit-faithful for every power of two.
#pragma once
#include "fft_kernels.cuh"

namespace kofft {

template <bool EXACT, class C2 = float2>
KHD void small_fft2(C2 *x) // fft_kernels.rs:4-11
{
    C2 a = x[0], b = x[1];
    x[0] = add2(a, b);
    x[1] = sub2(a, b);
}

template <bool EXACT, class C2 = float2>
KHD void small_fft4(C2 *x) // :13-30
{
    C2 a0 = x[0], a1 = x[1], a2 = x[2], a3 = x[3];
    C2 even0 = add2(a0, a2), even1 = sub2(a0, a2);
    C2 odd0 = add2(a1, a3), odd1 = sub2(a1, a3);
    C2 t1 = cmul<EXACT>(odd1, Cx<C2>::make(0.0f, -1.0f));
    x[0] = add2(even0, odd0);
    x[2] = sub2(even0, odd0);
    x[1] = add2(even1, t1);
    x[3] = sub2(even1, t1);
}

// the 8-point sub-transform shared by fft8's two halves and fft16's two halves is NOT the
// same code in the reference (fft8 combines in place, fft16 names temporaries), but both
// evaluate the same expression tree; restated once per caller below to keep operand order.
template <bool EXACT, class C2 = float2>
KHD void small_fft8(C2 *x) // :32-88
{
    const C2 w1 = Cx<C2>::make(0.0f, -1.0f);
    const float s = 0.70710677f;
    C2 x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3], x4 = x[4], x5 = x[5], x6 = x[6], x7 = x[7];
    C2 a0 = add2(x0, x4), a1 = sub2(x0, x4), a2 = add2(x2, x6), a3 = sub2(x2, x6);
    C2 t = cmul<EXACT>(a3, w1);
    C2 e0 = add2(a0, a2), e2 = sub2(a0, a2), e1 = add2(a1, t), e3 = sub2(a1, t);
    C2 b0 = add2(x1, x5), b1 = sub2(x1, x5), b2 = add2(x3, x7), b3 = sub2(x3, x7);
    t = cmul<EXACT>(b3, w1);
    C2 o0 = add2(b0, b2), o2 = sub2(b0, b2), o1 = add2(b1, t), o3 = sub2(b1, t);
    C2 t1 = cmul<EXACT>(o1, Cx<C2>::make(s, -s));
    C2 t2 = cmul<EXACT>(o2, w1);
    C2 t3 = cmul<EXACT>(o3, Cx<C2>::make(-s, -s));
    x[0] = add2(e0, o0);
    x[4] = sub2(e0, o0);
    x[1] = add2(e1, t1);
    x[5] = sub2(e1, t1);
    x[2] = add2(e2, t2);
    x[6] = sub2(e2, t2);
    x[3] = add2(e3, t3);
    x[7] = sub2(e3, t3);
}

template <bool EXACT, class C2 = float2>
KHD void small_fft16_half(C2 y0, C2 y1, C2 y2, C2 y3, C2 y4, C2 y5, C2 y6,
                          C2 y7, C2 *out) // :114-147 / :149-182
{
    const C2 w1 = Cx<C2>::make(0.0f, -1.0f);
    const float s = 0.70710677f;
    C2 a0 = add2(y0, y1), a1 = sub2(y0, y1), a2 = add2(y2, y3), a3 = sub2(y2, y3);
    C2 t = cmul<EXACT>(a3, w1);
    C2 ea0 = add2(a0, a2), ea2 = sub2(a0, a2), ea1 = add2(a1, t), ea3 = sub2(a1, t);
    C2 b0 = add2(y4, y5), b1 = sub2(y4, y5), b2 = add2(y6, y7), b3 = sub2(y6, y7);
    t = cmul<EXACT>(b3, w1);
    C2 eb0 = add2(b0, b2), eb2 = sub2(b0, b2), eb1 = add2(b1, t), eb3 = sub2(b1, t);
    C2 t0 = eb0;
    C2 t1 = cmul<EXACT>(eb1, Cx<C2>::make(s, -s));
    C2 t2 = cmul<EXACT>(eb2, w1);
    C2 t3 = cmul<EXACT>(eb3, Cx<C2>::make(-s, -s));
    out[0] = add2(ea0, t0);
    out[1] = add2(ea1, t1);
    out[2] = add2(ea2, t2);
    out[3] = add2(ea3, t3);
    out[4] = sub2(ea0, t0);
    out[5] = sub2(ea1, t1);
    out[6] = sub2(ea2, t2);
    out[7] = sub2(ea3, t3);
}

template <bool EXACT, class C2 = float2>
KHD void small_fft16(C2 *x) // :90-224
{
    C2 in[16];
#pragma unroll
    for (int i = 0; i < 16; i++) in[i] = x[i];
    small_fft16_half<EXACT, C2>(in[0], in[8], in[4], in[12], in[2], in[10], in[6], in[14], x);
    small_fft16_half<EXACT, C2>(in[1], in[9], in[5], in[13], in[3], in[11], in[7], in[15], x + 8);
    const float c1 = 0.9238795f, s1 = -0.38268343f, c2 = 0.70710677f, s2 = -0.70710677f;
    const float c3 = 0.38268343f, s3 = -0.9238795f, c4 = 0.0f, s4 = -1.0f;
    C2 o[8];
    o[0] = x[8];
    o[1] = cmul<EXACT>(x[9], Cx<C2>::make(c1, s1));
    o[2] = cmul<EXACT>(x[10], Cx<C2>::make(c2, s2));
    o[3] = cmul<EXACT>(x[11], Cx<C2>::make(c3, s3));
    o[4] = cmul<EXACT>(x[12], Cx<C2>::make(c4, s4));
    o[5] = cmul<EXACT>(x[13], Cx<C2>::make(-c3, s3));
    o[6] = cmul<EXACT>(x[14], Cx<C2>::make(-c2, s2));
    o[7] = cmul<EXACT>(x[15], Cx<C2>::make(-c1, s1));
#pragma unroll
    for (int i = 0; i < 8; i++) {
        C2 e = x[i];
        x[i] = add2(e, o[i]);
        x[i + 8] = sub2(e, o[i]);
    }
}

template <int N, bool EXACT, class C2 = float2>
KHD void small_fft(C2 *x)
{
    if (N == 2) small_fft2<EXACT, C2>(x);
    if (N == 4) small_fft4<EXACT, C2>(x);
    if (N == 8) small_fft8<EXACT, C2>(x);
    if (N == 16) small_fft16<EXACT, C2>(x);
}

// element type of an I/O policy: float2 unless the policy says otherwise (the f64 twin, fft_f64.cuh)
template <class IO, class = void>
struct IoElem {
    typedef float2 type;
};
template <class IO>
struct IoElem<IO, typename IO::is_f64> {
    typedef double2 type;
};

// one thread, one transform
template <int N, bool EXACT, class IO>
KHD void small_transform(const IO &io, long row)
{
    typename IoElem<IO>::type x[N < 16 ? 16 : N];
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = io.load(row, i);
    small_fft<N, EXACT, typename IoElem<IO>::type>(x);
    if constexpr (IO::kEpilogueExchange) {
#pragma unroll
        for (int k = 0; k < N; k++) io.epilogue(row, k, x);
    } else {
#pragma unroll
        for (int i = 0; i < N; i++) io.store(row, i, x[i]);
    }
}

#ifdef __CUDACC__
template <int N, bool EXACT, class IO>
__global__ void __launch_bounds__(128) fft_small_kernel(const __grid_constant__ IO io, long rows)
{
    for (long row = blockIdx.x * (long)blockDim.x + threadIdx.x; row < rows; row += (long)gridDim.x * blockDim.x)
        small_transform<N, EXACT, IO>(io, row);
}
#endif

} // namespace kofft
