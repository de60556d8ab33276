// Microbenchmark: issue throughput of scalar vs packed FP32 ops on sm_100a (per SMSP, warp-instr / clk).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2 fp32x2.cu && ./fp32x2
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
#define NACC 16
template <int OP>
__global__ void __launch_bounds__(1024) k(float *out, float seed, long long *cyc)
{
    float a[NACC], b[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) { a[i] = seed + i + threadIdx.x; b[i] = seed * 0.5f + i; }
    float m0 = seed * 1.0001f, m1 = seed * 0.9999f;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (OP == 0) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(m0));
            if (OP == 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(m0));
            if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(m0), "f"(m1));
            if (OP == 3) asm volatile("{.reg .b64 x, y; mov.b64 x, {%0,%1}; mov.b64 y, {%2,%3}; mul.rn.f32x2 x, x, y; mov.b64 {%0,%1}, x;}" : "+f"(a[i]), "+f"(b[i]) : "f"(m0), "f"(m1));
            if (OP == 4) asm volatile("{.reg .b64 x, y; mov.b64 x, {%0,%1}; mov.b64 y, {%2,%3}; add.rn.f32x2 x, x, y; mov.b64 {%0,%1}, x;}" : "+f"(a[i]), "+f"(b[i]) : "f"(m0), "f"(m1));
            if (OP == 5) asm volatile("{.reg .b64 x, y; mov.b64 x, {%0,%1}; mov.b64 y, {%2,%3}; fma.rn.f32x2 x, x, y, y; mov.b64 {%0,%1}, x;}" : "+f"(a[i]), "+f"(b[i]) : "f"(m0), "f"(m1));
            if (OP == 6) { // exact butterfly mix A: 4 FMUL + 2 FADD + 2 FADD2 (current EXACT)
                float p0, p1, p2, p3;
                asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(p0) : "f"(a[i]), "f"(m0));
                asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(p1) : "f"(b[i]), "f"(m1));
                asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(p2) : "f"(a[i]), "f"(m1));
                asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(p3) : "f"(b[i]), "f"(m0));
                asm volatile("sub.rn.f32 %0, %0, %1;" : "+f"(p0) : "f"(p1));
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(p2) : "f"(p3));
                asm volatile("{.reg .b64 x, y, z; mov.b64 x, {%0,%1}; mov.b64 y, {%2,%3}; add.rn.f32x2 z, x, y; sub.rn.f32x2 x, x, y; mov.b64 {%0,%1}, x; mov.b64 {%2,%3}, z;}" : "+f"(a[i]), "+f"(b[i]), "+f"(p0), "+f"(p2));
            }
            if (OP == 7) { // exact butterfly mix B: 2 FMUL2 + 2 FADD + 2 FADD2
                float p0, p1, p2, p3;
                asm volatile("{.reg .b64 x, y; mov.b64 x, {%2,%2}; mov.b64 y, {%3,%4}; mul.rn.f32x2 x, x, y; mov.b64 {%0,%1}, x;}" : "=f"(p0), "=f"(p2) : "f"(a[i]), "f"(m0), "f"(m1));
                asm volatile("{.reg .b64 x, y; mov.b64 x, {%2,%2}; mov.b64 y, {%3,%4}; mul.rn.f32x2 x, x, y; mov.b64 {%0,%1}, x;}" : "=f"(p1), "=f"(p3) : "f"(b[i]), "f"(m1), "f"(m0));
                asm volatile("sub.rn.f32 %0, %0, %1;" : "+f"(p0) : "f"(p1));
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(p2) : "f"(p3));
                asm volatile("{.reg .b64 x, y, z; mov.b64 x, {%0,%1}; mov.b64 y, {%2,%3}; add.rn.f32x2 z, x, y; sub.rn.f32x2 x, x, y; mov.b64 {%0,%1}, x; mov.b64 {%2,%3}, z;}" : "+f"(a[i]), "+f"(b[i]), "+f"(p0), "+f"(p2));
            }
            if (OP == 8) { // fast butterfly: FMUL2 + 2 FFMA + 2 FADD2
                float p1, p3, p0, p2;
                asm volatile("{.reg .b64 x, y; mov.b64 x, {%2,%2}; mov.b64 y, {%3,%4}; mul.rn.f32x2 x, x, y; mov.b64 {%0,%1}, x;}" : "=f"(p1), "=f"(p3) : "f"(b[i]), "f"(m1), "f"(m0));
                asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(p0) : "f"(a[i]), "f"(m0), "f"(p1));
                asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(p2) : "f"(a[i]), "f"(m1), "f"(p3));
                asm volatile("{.reg .b64 x, y, z; mov.b64 x, {%0,%1}; mov.b64 y, {%2,%3}; add.rn.f32x2 z, x, y; sub.rn.f32x2 x, x, y; mov.b64 {%0,%1}, x; mov.b64 {%2,%3}, z;}" : "+f"(a[i]), "+f"(b[i]), "+f"(p0), "+f"(p2));
            }
            if (OP == 9) { // 1 FMUL + 1 IADD-ish alu op (do they dual-issue across pipes?)
                asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(m0));
                int t = __float_as_int(b[i]);
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(t) : "r"(it));
                b[i] = __int_as_float(t);
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char *name, int instr_per_unit, int threads)
{
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    k<OP><<<148, threads>>>(out, 1.0f, cyc);
    k<OP><<<148, threads>>>(out, 1.0f, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
    double winstr = (double)ITER * NACC * instr_per_unit * (threads / 32) / 4.0; // per SMSP
    printf("%-44s warps/SMSP=%2d  %.3f warp-instr/clk/SMSP  (%.1f clk per unit per warp-slot)\n", name, threads / 128, winstr / c, c / ((double)ITER * NACC * (threads / 128)));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int threads : {256, 512, 1024}) {
        run<0>("FMUL", 1, threads); run<1>("FADD", 1, threads); run<2>("FFMA", 1, threads);
        run<3>("FMUL2", 1, threads); run<4>("FADD2", 1, threads); run<5>("FFMA2", 1, threads);
        run<6>("butterfly 4FMUL+2FADD+2FADD2 (8 instr)", 8, threads);
        run<7>("butterfly 2FMUL2+2FADD+2FADD2 (6 instr)", 6, threads);
        run<8>("butterfly FMUL2+2FFMA+2FADD2 (5 instr)", 5, threads);
        run<9>("FMUL + LOP3 (2 instr)", 2, threads);
    }
    return 0;
}
