// dist_kernels.cu -- the data-movement kernel of the multi-GPU four-step transform (BASELINE
// configs[4]: one C2C transform too large for one GPU).  There is no reference
// code for this path (kofft is single-process CPU code and its planner table is numerically
// meaningless at 2^30).
//
// transpose_scatter: the local matrix S [rows][G*cb] (row-major complex) is cut into G column
// blocks; block d is transposed and stored into destination d's buffer
//     D_d[c * dst_pitch + dst_off + r] = S[r][d*cb + c] * W_N^{(row0 + r) * (d*cb + c)}   (twiddle optional)
// D_d may be a peer GPU's memory (NVLink P2P stores): the all-to-all exchange of the four-step
// FFT *is* this kernel's store, no staging buffer and no separate collective.  Tiles are 32 x 32
// through padded shared memory, so loads and (peer) stores are both 256-byte runs.
#include "dist_kernels.h"

namespace kofft {

namespace {

__device__ __forceinline__ float2 cmulf(float2 a, float2 b)
{
    return make_float2(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x));
}

__global__ void __launch_bounds__(256) transpose_scatter_kernel(const __grid_constant__ ScatterArgs a)
{
    __shared__ float2 tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5; // 32 x 8
    const long cols = (long)a.world * a.cb;
    const long tiles_c = cols >> 5, tiles_r = a.rows >> 5;
    const long ntiles = tiles_c * tiles_r;
    const unsigned lomask = (1u << a.llo) - 1u;
    for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        // Consecutive tiles go to DIFFERENT destinations, starting with this rank's right neighbour:
        // every GPU keeps all peers busy at once and no peer is the target of everybody at the same
        // time (all ranks run this loop in lock step; destination-major order makes the all-to-all an
        // 8-to-1 incast, measured 249 GB/s per GPU instead of the link rate).  Within a destination,
        // consecutive tiles walk down a tile column, extending each other's runs at the destination.
        const int di = (int)(t % a.world);
        const long u = t / a.world;
        const long tcd = u / tiles_r, tr = u - tcd * tiles_r;
        const int dsel = (a.rank + 1 + di) % a.world;
        const long r0 = tr << 5, c0 = (long)dsel * a.cb + (tcd << 5);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const long r = r0 + ty + 8 * i, c = c0 + tx;
            float2 v = a.src[r * cols + c];
            if (a.twiddle) {
                // exponent (row0 + r) * c < N; W_N^p = T_hi[p >> llo] * T_lo[p & mask], both tables f64-rounded
                const unsigned long long p = (unsigned long long)(a.row0 + r) * (unsigned long long)c;
                const unsigned pm = (unsigned)(p & ((1ull << a.log2n) - 1ull));
                float2 w = cmulf(__ldg(a.thi + (pm >> a.llo)), __ldg(a.tlo + (pm & lomask)));
                if (a.twiddle == 2) w.y = -w.y; // inverse transform: conjugate twiddle
                v = cmulf(v, w);
            }
            tile[ty + 8 * i][tx] = v;
        }
        __syncthreads();
        const int d = dsel;
        const long cin = tcd << 5; // column inside the destination block
        float2 *dst = a.dst[d];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const long c = cin + ty + 8 * i, r = r0 + tx;
            dst[c * a.dst_pitch + a.dst_off + r] = tile[tx][ty + 8 * i];
        }
        __syncthreads();
    }
}

} // namespace

cudaError_t launch_transpose_scatter(const ScatterArgs &a, int num_sms, cudaStream_t stream)
{
    if (a.rows % 32 != 0 || a.cb % 32 != 0 || a.world < 1 || a.world > kMaxDistWorld) return cudaErrorInvalidValue;
    const long ntiles = (a.rows >> 5) * (((long)a.world * a.cb) >> 5);
    if (ntiles == 0) return cudaSuccess;
    long grid = (long)num_sms * 8;
    if (grid > ntiles) grid = ntiles;
    transpose_scatter_kernel<<<(int)grid, 256, 0, stream>>>(a);
    return cudaGetLastError();
}

} // namespace kofft
