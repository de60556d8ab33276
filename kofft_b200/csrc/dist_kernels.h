// dist_kernels.h -- see dist_kernels.cu
#pragma once
#include <cuda_runtime.h>

namespace kofft {

constexpr int kMaxDistWorld = 16;

struct ScatterArgs {
    const float2 *src = nullptr;      // [rows][world * cb] row-major
    float2 *dst[kMaxDistWorld] = {};  // destination buffers (local or peer device memory)
    long rows = 0, cb = 0;            // both multiples of 32
    int world = 1;
    int rank = 0;                     // destinations are visited starting at rank + 1 (no incast)
    long dst_pitch = 0;               // elements between consecutive destination rows
    long dst_off = 0;                 // column offset at the destination
    int twiddle = 0;                  // 1: multiply by W_N^{(row0 + r) * c}; 2: by its conjugate
    long row0 = 0;
    int log2n = 0, llo = 0;           // N = 2^log2n; exponent split at bit llo
    const float2 *tlo = nullptr;      // exp(-2 pi i j / N),        j < 2^llo
    const float2 *thi = nullptr;      // exp(-2 pi i j 2^llo / N),  j < 2^(log2n - llo)
};

cudaError_t launch_transpose_scatter(const ScatterArgs &a, int num_sms, cudaStream_t stream);

} // namespace kofft
