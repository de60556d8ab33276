// cuda_emu.h -- a tiny cooperative CUDA-block emulator for CPU-only CI (test scaffolding).
//
// Lets a __global__ kernel body written for nvcc be compiled by g++ and executed on the host:
// every CUDA thread of a block is a ucontext coroutine, __syncthreads() is a counting barrier
// that yields to the scheduler, threadIdx/blockIdx/blockDim/gridDim are globals the scheduler
// sets before resuming a thread, and `__shared__` variables become statics (one block runs at
// a time).  The mbarrier / TMA-bulk-copy helpers used by the kernels have host stand-ins here.
//
// Only tests/emu/*.cpp includes this; it is never part of the product build.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
inline emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __grid_constant__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __ldg(p) (*(p))
// threads are cooperative coroutines: a plain read-modify-write is atomic
inline int atomicMax(int *a, int v) { int old = *a; if (v > old) *a = v; return old; }

namespace cuda_emu {

struct Thread {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = false;
    emu_dim3 tid;
    unsigned block = 0; // index of the thread's block inside the cluster
};

struct NamedBar {
    unsigned long long generation = 0;
    int arrived = 0;
};
struct WarpState {
    unsigned long long generation = 0;
    int arrived = 0;
    int alive = 0;
    float xchg[32] = {};
};
struct BlockState {
    unsigned long long generation = 0;
    int arrived = 0;
    int alive = 0;
    emu_dim3 bid;
    NamedBar named[16];           // bar.sync id, n
    std::vector<WarpState> warps; // __syncwarp / shuffles
};

struct Scheduler {
    ucontext_t main_ctx;
    std::vector<Thread> threads;
    std::vector<BlockState> blocks; // the blocks of one cluster run concurrently
    int current = -1;
    unsigned long long cluster_generation = 0;
    int cluster_arrived = 0;
    int alive = 0;
    std::function<void()> body;
};

inline Scheduler *g_sched = nullptr;
inline size_t g_stack_bytes = 256 * 1024; // per coroutine; kernels with many threads per team lower it
inline unsigned g_late_from = ~0u;      // see run_cluster
inline unsigned long long g_late_ratio = 1;

inline void yield_to_scheduler()
{
    Scheduler *s = g_sched;
    Thread &t = s->threads[s->current];
    swapcontext(&t.ctx, &s->main_ctx);
}

inline void syncthreads()
{
    Scheduler *s = g_sched;
    BlockState &b = s->blocks[s->threads[s->current].block];
    const unsigned long long gen = b.generation;
    if (++b.arrived == b.alive) {
        b.arrived = 0;
        b.generation++;
        return;
    }
    while (b.generation == gen) yield_to_scheduler();
}

// barrier.cluster.arrive + wait: every thread of every block of the cluster
inline void cluster_sync()
{
    Scheduler *s = g_sched;
    const unsigned long long gen = s->cluster_generation;
    if (++s->cluster_arrived == s->alive) {
        s->cluster_arrived = 0;
        s->cluster_generation++;
        return;
    }
    while (s->cluster_generation == gen) yield_to_scheduler();
}

inline unsigned cluster_rank() { return g_sched->threads[g_sched->current].block; }

// bar.sync id, nthreads: the nthreads threads that name the barrier (all of them must be alive)
inline void named_barrier(int id, int nthreads)
{
    Scheduler *s = g_sched;
    NamedBar &b = s->blocks[s->threads[s->current].block].named[id & 15];
    const unsigned long long gen = b.generation;
    if (++b.arrived == nthreads) {
        b.arrived = 0;
        b.generation++;
        return;
    }
    while (b.generation == gen) yield_to_scheduler();
}

// bar.arrive id, nthreads: counts without waiting
inline void named_barrier_arrive(int id, int nthreads)
{
    Scheduler *s = g_sched;
    NamedBar &b = s->blocks[s->threads[s->current].block].named[id & 15];
    if (++b.arrived == nthreads) {
        b.arrived = 0;
        b.generation++;
    }
}

inline WarpState &my_warp()
{
    Scheduler *s = g_sched;
    Thread &t = s->threads[s->current];
    return s->blocks[t.block].warps[t.tid.x >> 5];
}
// __syncwarp(): the live threads of the calling thread's warp
inline void syncwarp()
{
    WarpState &w = my_warp();
    const unsigned long long gen = w.generation;
    if (++w.arrived == w.alive) {
        w.arrived = 0;
        w.generation++;
        return;
    }
    while (w.generation == gen) yield_to_scheduler();
}
// __all_sync over the full warp
inline bool warp_all(bool p)
{
    WarpState &w = my_warp();
    const unsigned lane = g_sched->threads[g_sched->current].tid.x & 31u;
    w.xchg[lane] = p ? 1.0f : 0.0f;
    syncwarp();
    bool r = true;
    for (int i = 0; i < 32; i++) r = r && w.xchg[i] != 0.0f;
    syncwarp();
    return r;
}
// __shfl_xor_sync over the full warp
inline float shfl_xor(float v, int mask)
{
    WarpState &w = my_warp();
    const unsigned lane = g_sched->threads[g_sched->current].tid.x & 31u;
    w.xchg[lane] = v;
    syncwarp();
    const float r = w.xchg[(lane ^ (unsigned)mask) & 31u];
    syncwarp();
    return r;
}

inline void trampoline()
{
    Scheduler *s = g_sched;
    s->body();
    Thread &t = s->threads[s->current];
    BlockState &b = s->blocks[t.block];
    t.done = true;
    s->alive--;
    b.alive--;
    {
        WarpState &w = b.warps[t.tid.x >> 5];
        w.alive--;
        if (w.alive > 0 && w.arrived == w.alive) {
            w.arrived = 0;
            w.generation++;
        }
    }
    // an exited thread no longer takes part in barriers
    if (b.alive > 0 && b.arrived == b.alive) {
        b.arrived = 0;
        b.generation++;
    }
    if (s->alive > 0 && s->cluster_arrived == s->alive) {
        s->cluster_arrived = 0;
        s->cluster_generation++;
    }
    swapcontext(&t.ctx, &s->main_ctx);
}

// run `body` once per thread of `cluster` consecutive blocks starting at first_block
inline void run_cluster(unsigned first_block, unsigned cluster, unsigned grid_x, unsigned threads_x,
                        const std::function<void()> &body)
{
    Scheduler s;
    s.body = body;
    s.threads.resize(static_cast<size_t>(threads_x) * cluster);
    s.blocks.resize(cluster);
    s.alive = static_cast<int>(threads_x * cluster);
    g_sched = &s;
    gridDim.x = grid_x;
    blockDim.x = threads_x;
    const size_t stack_bytes = g_stack_bytes;
    for (unsigned c = 0; c < cluster; c++) {
        s.blocks[c].alive = static_cast<int>(threads_x);
        s.blocks[c].bid.x = first_block + c;
        s.blocks[c].warps.resize((threads_x + 31) / 32);
        for (unsigned i = 0; i < threads_x; i++) s.blocks[c].warps[i >> 5].alive++;
        for (unsigned i = 0; i < threads_x; i++) {
            Thread &t = s.threads[c * threads_x + i];
            t.stack.resize(stack_bytes);
            t.tid.x = i;
            t.block = c;
            getcontext(&t.ctx);
            t.ctx.uc_stack.ss_sp = t.stack.data();
            t.ctx.uc_stack.ss_size = stack_bytes;
            t.ctx.uc_link = &s.main_ctx;
            makecontext(&t.ctx, reinterpret_cast<void (*)()>(trampoline), 0);
        }
    }
    // g_late_from / g_late_ratio: blocks with cluster rank >= g_late_from get one scheduling pass for
    // every g_late_ratio passes of the others -- a crude model of CTAs that start (or run) late, to
    // shake out flag protocols that only work while the CTAs of a team advance in lockstep
    unsigned long long pass = 0;
    while (s.alive > 0) {
        const bool late_turn = g_late_ratio <= 1 || pass % g_late_ratio == g_late_ratio - 1;
        bool ran = false;
        for (size_t i = 0; i < s.threads.size(); i++) {
            Thread &t = s.threads[i];
            if (t.done) continue;
            if (!late_turn && t.block >= g_late_from) continue;
            ran = true;
            s.current = static_cast<int>(i);
            threadIdx = t.tid;
            blockIdx = s.blocks[t.block].bid;
            swapcontext(&s.main_ctx, &t.ctx);
        }
        pass++;
        (void)ran;
    }
    g_sched = nullptr;
}

// launch a whole grid, block after block (or cluster after cluster)
inline void launch(unsigned grid_x, unsigned threads_x, const std::function<void()> &body, unsigned cluster = 1)
{
    for (unsigned b = 0; b < grid_x; b += cluster) run_cluster(b, cluster, grid_x, threads_x, body);
}

} // namespace cuda_emu

#define __syncthreads() cuda_emu::syncthreads()
