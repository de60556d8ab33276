// host_tables.h -- see host_tables.cpp
#pragma once
#include <cstddef>

namespace kofft {
void host_fft_twiddles(size_t n, float *out);                  // n/2 complex
void host_fft_twiddles_f64(size_t n, double *out);             // n/2 complex, FftPlanner<f64>
void host_accurate_twiddles(size_t n, size_t stride, size_t count, float *out); // exp(-2 pi i k stride / n)
void host_bluestein_chirp(size_t n, size_t m, float *chirp, float *b); // n and m complex
void host_bluestein_chirp_f64(size_t n, size_t m, double *chirp, double *b);
void host_rfft_twiddles(size_t m, float *out, bool fma_mul);   // m complex
void host_rfft_twiddles_f64(size_t m, double *out);            // m complex, build_twiddle_table::<f64>
int host_window(int kind, size_t len, float beta, float *out); // 0 ok, -1 unknown kind
} // namespace kofft
