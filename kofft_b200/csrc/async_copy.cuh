// async_copy.cuh -- mbarrier + TMA bulk-copy (cp.async.bulk) helpers shared by the kernels.
// SASS on sm_100a: SYNCS.* for the mbarrier ops, UBLKCP for the bulk copy.
// Under KOFFT_EMU (tests/emu, CPU-only CI) the same names are host stand-ins that perform the
// copy synchronously and enforce the hardware's 16-byte alignment / size rules.
#pragma once
#include "hostdev.h"

namespace kofft {

// L2 eviction-priority policies of one thread (see make_l2_policy below)
struct L2Policy {
    unsigned long long first = 0, last = 0;
};

#if defined(__CUDACC__)

KD unsigned smem_u32(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

KD void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
KD void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// spin until the phase with the given parity has completed
KD void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// generic-proxy accesses to shared memory (ld/st) ordered before later async-proxy (TMA) writes to it
KD void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// the same for global memory: rows written with ordinary stores (and acquired) before a tensor-map copy reads them
KD void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
// one arrival + the number of bytes the bulk copies of this phase will deliver
KD void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy; dst, src and bytes must be multiples of 16
KD void bulk_copy_g2s(void *dst_smem, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 16-byte asynchronous global -> shared copy (LDGSTS, L2 only), grouped completion per thread
KD void cp_async16(void *dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
// same with an L2 eviction-priority hint.  NOTE: check the SASS of every kernel using this
// (scripts/check_sass.sh): ptxas 12.9 can allocate an ODD uniform-register pair for the LDGSTS
// descriptor (desc[UR1]), which the hardware rejects as an illegal instruction.
KD void cp_async16_hint(void *dst_smem, const void *src, unsigned long long pol)
{
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "l"(pol)
                 : "memory");
}
// 8-byte variant (through L1; source rows that are only 8-byte aligned)
KD void cp_async8(void *dst_smem, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
KD void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
KD void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// all but the most recent PENDING committed groups are complete
template <int PENDING> KD void cp_async_wait_but() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

// thread-block cluster barrier with release/acquire semantics (all threads of all CTAs)
KD void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
KD unsigned cluster_ctarank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

// ---- L2 eviction-priority hints (createpolicy + .L2::cache_hint) ------------------------------
// The pipelined large-N kernel streams its input and output through L2 (evict_first) and pins the
// pass-A -> pass-B intermediate (evict_last) so that it never makes the round trip to HBM.
KD L2Policy make_l2_policy()
{
    // the encodings createpolicy.fractional.L2::evict_first / evict_last (fraction 1.0) produce
    L2Policy p;
    p.first = 0x12F0000000000000ull;
    p.last = 0x14F0000000000000ull;
    return p;
}
// read-only streaming load (never written during the launch)
KD float2 ldg_hint(const float2 *p, unsigned long long pol)
{
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;"
                 : "=f"(v.x), "=f"(v.y)
                 : "l"(p), "l"(pol));
    return v;
}
// L2-only load of data written by other CTAs of the same launch
KD float2 ldcg_hint(const float2 *p, unsigned long long pol)
{
    float2 v;
    asm volatile("ld.global.cg.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
    return v;
}
// (volatile, no memory clobber: ordered against the other hinted accesses, __syncthreads and
// __threadfence, but the compiler may still schedule shared-memory traffic around it; the
// kernels never touch these addresses with plain loads or stores)
KD void stg_hint(float2 *p, float2 v, unsigned long long pol)
{
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol));
}

// ---- dependency flags between co-resident CTAs of a persistent kernel (cooperative launch) ------
// A counter is incremented by one thread of a CTA after a barrier that orders the CTA's global
// stores before it (release: __threadfence + atomicAdd, the pattern cooperative_groups' grid sync
// uses); a consumer polls with relaxed loads (an acquire load would invalidate L1 on every
// iteration) and fences once when the count is reached, then releases its CTA through a barrier.
KD unsigned flag_load(const unsigned *c)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
    return v;
}
KD void flag_acquire() { __threadfence(); }
KD void flag_arrive(unsigned *c)
{
    // release: fence.acq_rel (MEMBAR.ALL.GPU) + relaxed add; __threadfence() is a sequentially consistent fence
    // (MEMBAR.SC.GPU), which the arrival does not need
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(c) : "memory");
}
// the arriving CTA only READ the guarded data (its loads have completed): no fence needed
KD void flag_arrive_relaxed(unsigned *c) { atomicAdd(c, 1u); }

// ---- warp-level and partial-CTA primitives of the warp-specialised large-N kernel (fft_split32.cuh) ----
// bar.sync id, n: barrier among the n threads (a multiple of 32) that name it; id 0 is __syncthreads
KD void named_barrier(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// bar.arrive: counts towards barrier `id` without waiting (the waiting threads use named_barrier)
KD void named_barrier_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
KD void warp_sync() { __syncwarp(); }
KD float shfl_xor_f(float v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }
// warp-group register reallocation (setmaxnreg, sm_90+): executed by all four warps of a warp group
template <int REGS> KD void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS> KD void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }
// 16-byte L2-only load of data written by other CTAs of the same launch
KD float4 ldcg_hint4(const void *p, unsigned long long pol)
{
    float4 v;
    asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
KD void nano_sleep(unsigned ns) { __nanosleep(ns); }

// one acquire load after a relaxed polling loop has seen the count: LDG.STRONG.GPU + CCTL.IVALL, no MEMBAR
KD unsigned flag_load_acquire(const unsigned *c)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
    return v;
}
// warp vote: true iff the predicate holds on every lane (the result is warp-uniform by construction)
KD bool warp_all(bool p) { return __all_sync(0xffffffffu, p) != 0; }
// arrival on a CTA-local counter in shared memory (release before, acquire after); returns the previous count
KD unsigned smem_count_arrive(unsigned *c)
{
    __threadfence_block();
    const unsigned old = atomicAdd(c, 1u);
    __threadfence_block();
    return old;
}

// ---- TMA tensor-map loads (cp.async.bulk.tensor, SASS UTMALDG) -----------------------------------
// The map is a CUtensorMap encoded on the host (cuTensorMapEncodeTiled) and passed as a __grid_constant__
// kernel parameter.  c0 = element offset in the innermost dimension, c1 = row.
struct alignas(64) TmaMap {
    unsigned long long opaque[16];
};
KD void tma_load_2d(void *dst_smem, const TmaMap *map, int c0, int c1, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst_smem)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

#elif defined(KOFFT_EMU)

inline void named_barrier(int id, int nthreads) { cuda_emu::named_barrier(id, nthreads); }
inline void named_barrier_arrive(int id, int nthreads) { cuda_emu::named_barrier_arrive(id, nthreads); }
inline void warp_sync() { cuda_emu::syncwarp(); }
inline float shfl_xor_f(float v, int mask) { return cuda_emu::shfl_xor(v, mask); }
template <int REGS> inline void setmaxnreg_inc() {}
template <int REGS> inline void setmaxnreg_dec() {}
inline float4 ldcg_hint4(const void *p, unsigned long long)
{
    if (reinterpret_cast<uintptr_t>(p) & 15) abort();
    return *reinterpret_cast<const float4 *>(p);
}
inline void nano_sleep(unsigned) {}

inline L2Policy make_l2_policy() { return L2Policy{1ull, 2ull}; }
inline float2 ldg_hint(const float2 *p, unsigned long long) { return *p; }
inline float2 ldcg_hint(const float2 *p, unsigned long long) { return *p; }
inline void stg_hint(float2 *p, float2 v, unsigned long long) { *p = v; }
// the emulator runs the CTAs of one "cluster" (here: a team) as interleaved coroutines: a poll yields
inline unsigned flag_load(const unsigned *c)
{
    cuda_emu::yield_to_scheduler();
    return *c;
}
inline void flag_acquire() {}
inline void flag_arrive(unsigned *c) { ++*c; }
inline void flag_arrive_relaxed(unsigned *c) { ++*c; }

inline void cp_async16(void *dst_smem, const void *src)
{
    if ((reinterpret_cast<uintptr_t>(dst_smem) & 15) || (reinterpret_cast<uintptr_t>(src) & 15)) abort();
    memcpy(dst_smem, src, 16);
}
inline void cp_async16_hint(void *dst_smem, const void *src, unsigned long long) { cp_async16(dst_smem, src); }
inline void cp_async8(void *dst_smem, const void *src)
{
    if ((reinterpret_cast<uintptr_t>(dst_smem) & 7) || (reinterpret_cast<uintptr_t>(src) & 7)) abort();
    memcpy(dst_smem, src, 8);
}
inline void cp_async_commit() {}
inline void cp_async_wait_all() {}
template <int PENDING> inline void cp_async_wait_but() {}

inline void cluster_sync() { cuda_emu::cluster_sync(); }
inline unsigned cluster_ctarank() { return cuda_emu::cluster_rank(); }

// *bar counts completed phases; a phase completes when the bytes announced by expect_tx have
// all been delivered (copies are synchronous here).
struct EmuMbar {
    unsigned completed;
    int pending;
};
static_assert(sizeof(EmuMbar) == 8, "fits the kernels' 8-byte mbarrier word");
inline EmuMbar &emu_mbar(unsigned long long *bar) { return *reinterpret_cast<EmuMbar *>(bar); }
inline void mbar_init(unsigned long long *bar, unsigned) { emu_mbar(bar) = EmuMbar{0u, 0}; }
inline void fence_mbar_init() {}
inline void fence_proxy_async() {}
inline void fence_proxy_async_global() {}
inline void mbar_wait(unsigned long long *bar, unsigned parity)
{
    while ((emu_mbar(bar).completed & 1u) == parity) cuda_emu::yield_to_scheduler();
}
inline void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    emu_mbar(bar).pending += static_cast<int>(bytes);
    if (bytes == 0) emu_mbar(bar).completed++;
}
inline void bulk_copy_g2s(void *dst_smem, const void *src, unsigned bytes, unsigned long long *bar)
{
    if ((reinterpret_cast<uintptr_t>(dst_smem) & 15) || (reinterpret_cast<uintptr_t>(src) & 15) || (bytes & 15) || !bytes)
        abort(); // the hardware would fault / hang on these
    memcpy(dst_smem, src, bytes);
    EmuMbar &m = emu_mbar(bar);
    m.pending -= static_cast<int>(bytes);
    if (m.pending == 0) m.completed++;
}

inline unsigned flag_load_acquire(const unsigned *c) { return *c; }
inline bool warp_all(bool p) { return cuda_emu::warp_all(p); }
inline unsigned smem_count_arrive(unsigned *c) { return (*c)++; }
// host stand-in of a 2-D tiled tensor map over f32 elements: dims / strides as cuTensorMapEncodeTiled takes them
struct TmaMap {
    const void *base;
    unsigned long long dim0, dim1;   // elements
    unsigned long long stride1;      // bytes between rows
    unsigned box0, box1;             // elements
    unsigned swizzle;                // 0 none, 64: CU_TENSOR_MAP_SWIZZLE_64B (address bits [5:4] ^= bits [8:7])
    unsigned pad_[21];
};
inline void tma_load_2d(void *dst_smem, const TmaMap *map, int c0, int c1, unsigned long long *bar)
{
    if ((reinterpret_cast<uintptr_t>(dst_smem) & 127) % 16) abort();
    if ((map->box0 * 4) % 16 || (map->stride1 % 16) || (reinterpret_cast<uintptr_t>(map->base) & 15)) abort();
    if (c0 < 0 || c1 < 0 || (unsigned long long)c0 + map->box0 > map->dim0 || (unsigned long long)c1 + map->box1 > map->dim1) abort();
    char *d = static_cast<char *>(dst_smem);
    if (map->swizzle == 64 && ((reinterpret_cast<uintptr_t>(dst_smem) & 511) || map->box0 * 4 != 64)) abort();
    for (unsigned r = 0; r < map->box1; r++) {
        const char *src = static_cast<const char *>(map->base) + (size_t)(c1 + r) * map->stride1 + (size_t)c0 * 4;
        if (map->swizzle == 64) { // 16-byte chunk index ^= bits [8:7] of the destination address
            for (unsigned ch = 0; ch < 4; ch++) {
                const size_t off = (size_t)r * 64 + ch * 16;
                memcpy(d + (off ^ (((off >> 7) & 3) << 4)), src + ch * 16, 16);
            }
        } else {
            memcpy(d + (size_t)r * map->box0 * 4, src, (size_t)map->box0 * 4);
        }
    }
    EmuMbar &m = emu_mbar(bar);
    m.pending -= static_cast<int>(map->box0 * map->box1 * 4);
    if (m.pending == 0) m.completed++;
}

#endif

} // namespace kofft
