// fft_split32.cuh -- N = 2^13 .. 2^15 complex (rfft / irfft 2^14 .. 2^16) as ONE persistent,
// warp-specialised kernel: the faithful two-pass split of kofft's radix-2 Stockham
// (src/fft.rs:789-912) with 32 elements per thread.
//
//   N = 2^LA * 32.   Index i = n * 32 + j'  (n: LA bits, j': 5 bits).
//   pass A = stages 0 .. LA-1: for every column j' a 2^LA-point transform over n (stride 32).  It
//       starts at stage 0, so its twiddles are those of a 2^LA-point transform read from the big
//       table with stride 32.  A tile = COLS = 8192 / 2^LA adjacent columns = 8192 elements, done by
//       the CTA's 8 "A warps" (256 threads x 32 elements) as two register passes with ONE exchange
//       through shared memory.  Results go to the intermediate (global memory, pinned in L2).
//   pass B = stages LA .. LA+4: for every k (the LA output bits produced by pass A) the 32
//       contiguous elements k*32 .. k*32+31 form a 32-point transform done IN ONE THREAD's registers
//       (twiddles T[(k + c_low 2^LA) << (4-t)], 31 per thread, resident for the kernel's lifetime);
//       bin c lands at K = k + c 2^LA, so 16 lanes holding adjacent k store 128-byte runs directly.
//       The CTA's 8 "B warps" work independently of each other: a warp owns 16 adjacent k and their
//       16 mirrors 2^LA - k (the rfft twist pairs bin K with m - K = (2^LA - k) + (31 - c) 2^LA, i.e.
//       lane l with lane 31 - l and register w with register 31 - w: two shuffles per bin), reads its
//       32 rows with 16-byte coalesced loads, transposes them through a private shared-memory
//       region (__syncwarp only) and never meets a block-wide barrier.
//
// Both roles run at the same time on every SM (A warps of transform i + 1 .. i + SLOTS - 1 next to
// the B warps of transform i), so the barrier waits and memory latencies of one role are filled by
// the other.  Compared with the 16-elements-per-thread pipeline in fft_large.cuh this halves the
// shared-memory exchanges per element and takes the block barriers per element from six per 8192
// to two per 8192 (both inside pass A).
//
// Dependencies: a "team" of NT = 32 / COLS consecutive CTAs owns the transforms team, team + teams,
// ...; member kb runs column tile kb and the warp tiles 8 kb .. 8 kb + 7.  Per team and intermediate
// slot one arrival counter per pass (see LargePipe in fft_large.cuh for why per slot).  The launch
// is cooperative: every CTA is resident, which the flag waits rely on.
#pragma once
#include "fft_kernels.cuh"

namespace kofft {

constexpr int WIDE = 32; // elements per thread

KHD constexpr int pad32(int i) { return i + (i >> 5); }

// pass-0 twiddles (group k = 0) of a radix-32 first pass: v[(2^t - 1) + c] = T[c << (L-1-t)]
struct Tw0W {
    float2 v[32];
};

// Stages S .. S+RR-1 of a 2^LS-point (sub-)transform, 32 elements per thread: U = 32 >> RR
// butterflies of radix 2^RR on x[u*R + w].  Index algebra as Pass<> in fft_engine.cuh.
template <int LS, int S, int RR, bool EXACT, bool UNIT0>
struct WidePass {
    static constexpr int r = RR, R = 1 << RR, U = WIDE >> RR;
    static constexpr int LJ = LS - S - RR, J = 1 << LJ;
    static constexpr int T = (1 << LS) / WIDE; // threads per transform
    static constexpr int NTW = U * (R - 1);
    static_assert(RR >= 1 && RR <= 5 && LJ >= 0, "pass shape");
    static KHD int bfly(int t, int u) { return t + u * T; }
    static KHD int src_index(int t, int u, int q)
    {
        const int b = bfly(t, u);
        return ((b >> LJ) << (LS - S)) + (b & (J - 1)) + (q << LJ);
    }
    static KHD int dst_index(int t, int u, int w)
    {
        const int b = bfly(t, u);
        return ((b >> LJ) << LJ) + (b & (J - 1)) + (bitrev(w, r) << (S + LJ));
    }
    // tw: UNIT0 -> thread-independent [R-1] (k = 0); else per thread [U][R-1]
    template <class TW>
    static KHD void compute(float2 *x, const TW *tw)
    {
#pragma unroll
        for (int u = 0; u < U; u++) {
#pragma unroll
            for (int tl = 0; tl < r; tl++) {
                const int bit = 1 << (r - 1 - tl);
#pragma unroll
                for (int w0 = 0; w0 < R; w0++) {
                    if (w0 & bit) continue;
                    const int c_low = bitrev(w0 >> (r - tl), tl);
                    float2 &a = x[u * R + w0];
                    float2 &b = x[u * R + (w0 | bit)];
                    if (UNIT0) {
                        if (c_low == 0)
                            butterfly_unit(a, b); // T[0] == (1, 0) exactly
                        else
                            butterfly<EXACT>(a, b, tw[(1 << tl) - 1 + c_low]);
                    } else {
                        butterfly<EXACT>(a, b, tw[u * (R - 1) + (1 << tl) - 1 + c_low]);
                    }
                }
            }
        }
    }
};

enum SplitEpilogue : int { SPLIT_STORE = 0, SPLIT_TWIST = 1 };

#ifndef KOFFT_SPLIT_SLOTS
#define KOFFT_SPLIT_SLOTS 4
#endif
// 1: every A warp polls / arrives by itself.  Measured slower (rfft 2^16 x 16384: 2.94 vs 2.61 ms, profiles/r03g):
// a warp's release fence right behind its 32 stores waits for all of them, eight times per tile; the block-level
// scheme arrives behind the NEXT tile's first barrier, when the stores have long landed.
// timing experiments only (scripts/build_variants.sh): 1 = the A warps alone (B warps return, slots are never waited
// for), 2 = the B warps alone (they do not wait for pass A and transform whatever the intermediate holds)
#ifndef KOFFT_SPLIT_ONLY_ROLE
#define KOFFT_SPLIT_ONLY_ROLE 0
#endif
// 1: the A warps' second block barrier per tile becomes bar.arrive for seven warps (only the warp that issues the
// next tile's copy waits); the B warps' polling interval in ns
#ifndef KOFFT_SPLIT_LAZY_BAR2
#define KOFFT_SPLIT_LAZY_BAR2 0
#endif
#ifndef KOFFT_SPLIT_POLL_NS
#define KOFFT_SPLIT_POLL_NS 32
#endif
#ifndef KOFFT_SPLIT_ZSLOTS
#define KOFFT_SPLIT_ZSLOTS 4
#endif
#ifndef KOFFT_SPLIT_UNROLL_A
#define KOFFT_SPLIT_UNROLL_A 0 // 0: per kind (see UNROLL_A)
#endif
#ifndef KOFFT_SPLIT_UNROLL_B
#define KOFFT_SPLIT_UNROLL_B 1
#endif
// irfft (PRE) tuning: transforms the untwist runs ahead of pass A; where a warp announces its untwisted bins (its
// release fence waits for the stores): 0 right behind them, 2 behind the tile's transposition reads, 1 behind the
// 32-point rows; 1: evict_last on those stores
#ifndef KOFFT_SPLIT_ZAHEAD
#define KOFFT_SPLIT_ZAHEAD 3
#endif
#ifndef KOFFT_SPLIT_ZDEFER
#define KOFFT_SPLIT_ZDEFER 0
#endif
#ifndef KOFFT_SPLIT_ZHINT
#define KOFFT_SPLIT_ZHINT 1
#endif
#ifndef KOFFT_SPLIT_WARP_FLAGS
#define KOFFT_SPLIT_WARP_FLAGS 0
#endif
// register budgets of the two roles (setmaxnreg): 256 * A + 256 * B == 512 * 128
#ifndef KOFFT_SPLIT_REGS_A
#define KOFFT_SPLIT_REGS_A 96
#endif
#ifndef KOFFT_SPLIT_REGS_B
#define KOFFT_SPLIT_REGS_B 160
#endif

// STAGED (LA = 10, plain contiguous rows): pass A's tile arrives by TMA tensor-map loads (four boxes of 256 rows x 8
// columns, SASS UTMALDG) in the exchange buffer itself while the previous tile's second register pass runs and
// stores -- the HBM latency of a tile is hidden behind the previous tile's arithmetic at no cost in shared
// memory.  The first register pass is "in place" in index space (a thread reads and writes the same 32
// positions of its column), so the landed tile doubles as the exchange buffer without an extra barrier.
template <int LA, bool EXACT, class IO, int EPI, bool STAGED = false, bool PRE = false>
struct Split32 {
    static_assert(LA >= 8 && LA <= 10, "N = 2^13 .. 2^15");
    static constexpr int L = LA + 5;
    static constexpr int NR = 1 << LA;            // rows of pass B per transform
    static constexpr int COLS = 8192 >> LA;       // columns per pass-A tile: 8, 16, 32
    static constexpr int NT = 32 / COLS;          // CTAs per team: 4, 2, 1
    static constexpr int LOG_NT = ilog2c(NT);
    static constexpr int RA0 = LA - 5;            // pass A = (RA0, 5) stages
    using A0 = WidePass<LA, 0, RA0, EXACT, true>;
    using A1 = WidePass<LA, RA0, 5, EXACT, false>;
    using PB = WidePass<5, 0, 5, EXACT, false>;
    static constexpr int TA = A0::T;              // threads per column: 32, 16, 8
    static_assert(TA * COLS == 256 && A1::U == 1 && A1::LJ == 0, "pass-A tile shape");
    static constexpr int A_THREADS = 256, B_THREADS = 256, CTA = A_THREADS + B_THREADS;
    static constexpr int B_WARPS = B_THREADS / 32;
    static constexpr int NTILES = NR / 32;        // warp tiles (16 rows + 16 mirrors) per transform
    static_assert(NTILES == NT * B_WARPS, "a team's B warps cover one transform");
    static constexpr int SLOTS = KOFFT_SPLIT_SLOTS;
    static constexpr int FLAG_STRIDE = 32;        // unsigned per team: cntA[SLOTS], cntB[SLOTS] (PRE: + cntZ, cntZf[ZSLOTS])
    // PRE (irfft): the B warps also untwist the rows (src/rfft.rs:485-498) ZAHEAD transforms ahead of pass A, into
    // ZSLOTS L2-resident rows per team that pass A's tile loads read instead of the caller's input
    static constexpr int ZSLOTS = KOFFT_SPLIT_ZSLOTS, ZAHEAD = KOFFT_SPLIT_ZAHEAD;
    static_assert(2 * SLOTS + (PRE ? 2 * ZSLOTS : 0) <= FLAG_STRIDE && ZAHEAD < ZSLOTS, "flag line");

    // ---- shared memory (float2 units) ----
    // pass-A exchange: one padded region per column; the region stride makes the 16 lanes of a half
    // warp (COLS columns x 16/COLS consecutive t) hit 16 distinct 8-byte banks
    static constexpr int PADN = NR + (NR >> 5);
    static constexpr int RSA = PADN + (COLS == 8 ? 2 : 1);
    // STAGED: the tile as the TMA delivers it, row-major [NR rows][COLS] in NBOX boxes of 256 rows
    static constexpr int NBOX = NR / 256;
    static_assert(!STAGED || (LA == 10 && (IoTraits<IO>::kRowPtr || PRE)), "staging: 2^15 points, plain contiguous rows");
    static_assert(!PRE || STAGED, "the untwisted rows are fetched as staged tiles");
    static constexpr int XA = STAGED ? NR * COLS : COLS * RSA;
    static constexpr unsigned TILE_BYTES = 8192u * 8u;
    // pass-A second-pass twiddles: [TA][33] (k1-dependent, shared by the columns)
    static constexpr int TWA = TA * 33;
    // pass-B transposition: per warp two halves of 16 rows x 34 (16-byte accesses, conflict-free)
    static constexpr int RSB = 34, HB = 16 * RSB, XB = 2 * HB;
    // rfft: the CTA's slice of T' (src/rfft.rs:172-183), [warp][register][lane]
    // (PRE: the same room holds the input bins a warp untwists next, [warp][k | n - k][16][lane])
    static constexpr int RTW = (EPI == SPLIT_TWIST || PRE) ? B_WARPS * 32 * 32 : 0;
    static constexpr int OFF_TWA = (XA + 1) & ~1;
    static constexpr int OFF_XB = (OFF_TWA + TWA + 1) & ~1;
    static constexpr int OFF_RTW = OFF_XB + B_WARPS * XB;
    static constexpr int OFF_BAR = OFF_RTW + RTW; // the mbarrier of the staged tile loads
    static constexpr int SMEM_BYTES = (OFF_BAR + 2) * 8;
    static constexpr bool HINT = IoTraits<IO>::kHint;

    static constexpr int A_WARPS = A_THREADS / 32;
    // copies of the roles' tile loops.  Measured (profiles/r05h, 2^15): two copies of the A warps' loop take C2C from 1.23 to
    // 1.18 ms, leave rfft where it is and slow the irfft variant down (4.7 ms); copies of the B warps' loop never help
    static constexpr int UNROLL_A = KOFFT_SPLIT_UNROLL_A > 0 ? KOFFT_SPLIT_UNROLL_A : ((EPI == SPLIT_STORE && !PRE && LA == 10) ? 2 : 1);
    static constexpr int UNROLL_B = KOFFT_SPLIT_UNROLL_B;
    static constexpr bool WARP_FLAGS = KOFFT_SPLIT_WARP_FLAGS != 0;
    static constexpr bool LAZY_BAR2 = KOFFT_SPLIT_LAZY_BAR2 != 0;
    static KD unsigned goal_a(long i) { return (unsigned)(NT * (WARP_FLAGS ? A_WARPS : 1) * (i / SLOTS + 1)); }
    static KD unsigned goal_b(long i) { return (unsigned)(NTILES * (i / SLOTS + 1)); }
    static KD unsigned goal_z(long j) { return (unsigned)(NTILES * (j / ZSLOTS + 1)); } // every B warp of the team has written row j
    static KD unsigned goal_zfree(long j) { return (unsigned)(NT * (j / ZSLOTS)); }     // the tiles of row j - ZSLOTS have landed

    // A warps, per tile: wait (whole warp polls: no block barrier involved) until every B warp of the team has consumed
    // transform i - SLOTS, whose slot this tile's stores overwrite; after the stores each warp arrives by itself
    // (a write-after-read dependency: the B warps arrive once their loads have RETURNED, so observing the count is
    // enough -- no acquire.)  `seen` is a value of the counter loaded earlier: the L2 round trip of the flag load
    // is not on the critical path when the slot is already free.
    static KD void a_wait_slot(unsigned *cntB, long i, unsigned seen = 0)
    {
        if (KOFFT_SPLIT_ONLY_ROLE == 1) return;
        if (i >= SLOTS) {
            const unsigned want = goal_b(i - SLOTS);
            while (seen < want) {
                nano_sleep(32);
                seen = flag_load(cntB + i % SLOTS);
            }
        }
    }
    static KD void a_arrive(unsigned *cntA, long i, int tid)
    {
        warp_sync(); // orders the warp's stores before lane 0's release
        if ((tid & 31) == 0) flag_arrive(cntA + i % SLOTS);
    }

    // ------------------------------------------------------------------------------------------
    // A warps: tid 0 .. 255
    // ------------------------------------------------------------------------------------------
    static KD void a_role(const IO &io, const Tw0W &tw0, const float2 *__restrict__ table, long cnt, long team,
                          long teams, int kb, float2 *__restrict__ slots, float2 *smem, unsigned *cntA, unsigned *cntB,
                          int tid)
    {
        const long n = 1L << L;
        const int col = tid % COLS, t = tid / COLS;
        float2 *bf = smem + col * RSA;
        float2 *twa = smem + OFF_TWA;
        const L2Policy pol = make_l2_policy();
        // second-pass twiddles of the column transform: entry e = (2^tl - 1) + c of row k1 is
        // T[(k1 + (c << RA0)) << (LA-1-RA0-tl + 5)]
        for (int i = tid; i < TA * 31; i += A_THREADS) {
            const int k1 = i / 31, e = i - k1 * 31;
            int tl = 0;
            while ((2 << tl) - 1 <= e) tl++;
            const int c = e + 1 - (1 << tl);
            twa[k1 * 33 + e] = table[(long)(k1 + (c << RA0)) << (LA - 1 - RA0 - tl + 5)];
        }
        named_barrier(1, A_THREADS);
        const float2 *tw1 = twa + t * 33;
        const long j0 = (long)kb * COLS + col;
#pragma unroll(UNROLL_A)
        for (long i = 0; i < cnt; i++) {
            const long row = team + i * teams;
            unsigned seen = 0;
            if (!WARP_FLAGS && tid == 0 && i >= SLOTS) seen = flag_load(cntB + i % SLOTS); // consumed before the second barrier
            float2 x[WIDE];
#pragma unroll
            for (int u = 0; u < A0::U; u++)
#pragma unroll
                for (int q = 0; q < A0::R; q++) {
                    const int idx = (int)(((long)A0::src_index(t, u, q) << 5) + j0);
                    if constexpr (HINT)
                        x[u * A0::R + q] = io.load_hint(row, idx, pol.first);
                    else
                        x[u * A0::R + q] = io.load(row, idx);
                }
            A0::compute(x, tw0.v);
#pragma unroll
            for (int u = 0; u < A0::U; u++)
#pragma unroll
                for (int w = 0; w < A0::R; w++) bf[pad32(A0::dst_index(t, u, w))] = x[u * A0::R + w];
            named_barrier(1, A_THREADS); // exchange; also: every thread is past the previous tile's stores
            if (!WARP_FLAGS && tid == 32 && i > 0) flag_arrive(cntA + (i - 1) % SLOTS); // (not the polling thread's warp)
#pragma unroll
            for (int q = 0; q < 32; q++) x[q] = bf[pad32(A1::src_index(t, 0, q))];
            if (!WARP_FLAGS && tid == 0) a_wait_slot(cntB, i, seen); // released to the others by the barrier
            named_barrier(1, A_THREADS); // the exchange buffer is free again
            A1::compute(x, tw1);
            if (WARP_FLAGS) a_wait_slot(cntB, i);
            float2 *o = slots + (i % SLOTS) * n + j0;
#pragma unroll
            for (int w = 0; w < 32; w++) stg_hint(o + ((long)A1::dst_index(t, 0, w) << 5), x[w], pol.last);
            if (WARP_FLAGS) a_arrive(cntA, i, tid);
        }
        if (!WARP_FLAGS && cnt > 0) {
            named_barrier(1, A_THREADS);
            if (tid == 0) flag_arrive(cntA + (cnt - 1) % SLOTS);
        }
    }

    // ------------------------------------------------------------------------------------------
    // A warps, STAGED variant (see the struct comment).  The tile lands row-major, [1024 rows][8 columns] (TMA
    // destinations are 128-byte aligned, so the four boxes are contiguous).  Thread mappings (tid 0 .. 255):
    //   first pass : column tid & 7, t0 = tid >> 3; reads rows q 32 + t0 (a half warp = 8 columns x 2 consecutive
    //                rows = 16 consecutive banks).  A WARP owns, for every q, the four rows q 32 + 4w .. + 3 of
    //                the 8 columns, reads them and writes its results back into the same set of positions
    //                (__syncwarp in between): output c of thread t0 goes to row c 32 + (t0 ^ bit3(c)).
    //   second pass: column tid & 7, k1 = ((tid >> 3) & 3) 8 + (tid >> 5); reads rows k1 32 + (q ^ bit3(k1)).  The
    //                two k1 of a half warp differ in bit 3, so their rows have different parity = different banks
    //                (rows are 64 bytes: without the swap every row of the same q would share a bank).
    // ------------------------------------------------------------------------------------------
    static KD void issue_tile(const TmaMap *map, float2 *stage, unsigned long long *bar, long row, int kb)
    {
        mbar_expect_tx(bar, TILE_BYTES);
#pragma unroll
        for (int b = 0; b < NBOX; b++)
            tma_load_2d(stage + b * 256 * COLS, map, kb * COLS * 2, (int)(row * NR + b * 256), bar);
    }
    static KD void a_role_staged(const IO &io, const Tw0W &tw0, const float2 *__restrict__ table, long cnt, long team,
                                 long teams, int kb, float2 *__restrict__ slots, float2 *smem, unsigned *cntA,
                                 unsigned *cntB, int tid, const TmaMap *map, unsigned long long *bar)
    {
        unsigned *cntZ = cntB + SLOTS, *cntZf = cntZ + ZSLOTS;
        // the tile's row in the tensor map: the caller's row, or (PRE) the team's slot of untwisted rows
        auto tile_row = [&](long i) { return PRE ? team * ZSLOTS + i % ZSLOTS : team + i * teams; };
        // PRE: the row exists once every B warp of the team has arrived; their (generic-proxy) stores are then
        // ordered before the tile copy's (async-proxy) reads
        auto wait_row = [&](long i) {
            if constexpr (PRE) {
                const unsigned want = goal_z(i);
                while (flag_load(cntZ + i % ZSLOTS) < want) nano_sleep(32);
                (void)flag_load_acquire(cntZ + i % ZSLOTS);
                fence_proxy_async_global();
            }
        };
        const long n = 1L << L;
        const int col = tid & 7, t0 = tid >> 3;
        const int k1 = ((tid >> 3) & 3) * 8 + (tid >> 5);
        float2 *stage = smem;
        float2 *twa = smem + OFF_TWA;
        const L2Policy pol = make_l2_policy();
        if (tid == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        for (int i = tid; i < TA * 31; i += A_THREADS) {
            const int kk = i / 31, e = i - kk * 31;
            int tl = 0;
            while ((2 << tl) - 1 <= e) tl++;
            const int c = e + 1 - (1 << tl);
            twa[kk * 33 + e] = table[(long)(kk + (c << RA0)) << (LA - 1 - RA0 - tl + 5)];
        }
        named_barrier(1, A_THREADS);
        if (tid == 0 && cnt > 0) {
            wait_row(0);
            issue_tile(map, stage, bar, tile_row(0), kb);
        }
        const float2 *tw1 = twa + k1 * 33;
        float2 *s0 = stage + t0 * COLS + col;        // row t0 (+ q 32 rows)
        float2 *s0x = stage + (t0 ^ 1) * COLS + col; // the neighbour's row, for outputs with bit 3 set
        const int sbit = (k1 >> 3) & 1;
        const float2 *s1e = stage + (k1 * 32 + sbit) * COLS + col; // even q: row k1 32 + q + sbit
        const float2 *s1o = stage + (k1 * 32 - sbit) * COLS + col; // odd q:  row k1 32 + q - sbit
        const long j0 = (long)kb * COLS + col;
        unsigned phase = 0;
#pragma unroll(UNROLL_A)
        for (long i = 0; i < cnt; i++) {
            unsigned seen = 0;
            if (!WARP_FLAGS && tid == 0 && i >= SLOTS) seen = flag_load(cntB + i % SLOTS); // consumed before the second barrier
            float2 x[WIDE];
            mbar_wait(bar, phase);
            phase ^= 1;
            if (PRE && tid == 64) flag_arrive_relaxed(cntZf + i % ZSLOTS); // the row's slot may be rewritten (by this CTA's share)
#pragma unroll
            for (int q = 0; q < 32; q++) x[q] = io.from_raw(s0[q * 32 * COLS]);
            A0::compute(x, tw0.v);
            warp_sync(); // the warp has read its rows: they may be rewritten
#pragma unroll
            for (int w = 0; w < 32; w++) {
                const int c = bitrev(w, 5);
                (((c >> 3) & 1) ? s0x : s0)[c * 32 * COLS] = x[w];
            }
            if (LAZY_BAR2 && !WARP_FLAGS && tid == 0) a_wait_slot(cntB, i, seen); // released to the others by the barrier
            named_barrier(1, A_THREADS); // exchange; also: every thread is past the previous tile's stores
            if (!WARP_FLAGS && tid == 32 && i > 0) flag_arrive(cntA + (i - 1) % SLOTS); // (not the polling thread's warp)
#pragma unroll
            for (int q = 0; q < 32; q++) x[q] = ((q & 1) ? s1o : s1e)[q * COLS];
            if (LAZY_BAR2) {
                // only the warp that issues the next tile's copy waits for the others' reads
                if (tid < 32)
                    named_barrier(2, A_THREADS);
                else
                    named_barrier_arrive(2, A_THREADS);
            } else {
                if (!WARP_FLAGS && tid == 0) a_wait_slot(cntB, i, seen); // released to the others by the barrier
                named_barrier(1, A_THREADS); // the buffer is free: the next tile may land
            }
            if (tid == 0 && i + 1 < cnt) {
                wait_row(i + 1);
                fence_proxy_async();
                issue_tile(map, stage, bar, tile_row(i + 1), kb);
            }
            A1::compute(x, tw1);
            if (WARP_FLAGS) a_wait_slot(cntB, i);
            float2 *o = slots + (i % SLOTS) * n + j0;
#pragma unroll
            for (int w = 0; w < 32; w++) stg_hint(o + ((long)(k1 + (bitrev(w, 5) << 5)) << 5), x[w], pol.last);
            if (WARP_FLAGS) a_arrive(cntA, i, tid);
        }
        if (!WARP_FLAGS && cnt > 0) {
            named_barrier(1, A_THREADS);
            if (tid == 0) flag_arrive(cntA + (cnt - 1) % SLOTS);
        }
    }

    // ------------------------------------------------------------------------------------------
    // B warps: tid 256 .. 511.  SPECIAL: the transform's last warp tile, whose lane 16 would repeat
    // row 2^(LA-1) (lane 15's, which mirrors itself) and takes row 0 (which mirrors itself) instead.
    // ------------------------------------------------------------------------------------------
    template <bool SPECIAL>
    static KD void b_role(const IO &io, const float2 *__restrict__ table, long cnt, long team, long teams, int kb,
                          const float2 *__restrict__ slots, float2 *smem, unsigned *cntA, unsigned *cntB, int wl,
                          float2 *__restrict__ zs = nullptr)
    {
        const long n = 1L << L;
        const int w = wl >> 5, lane = wl & 31, h = lane >> 4, lp = lane & 15;
        const int wt = kb * B_WARPS + w; // warp tile of the transform
        const int k0 = 1 + 16 * wt;      // low half: rows k0 .. k0 + 15
        const int kh0 = NR - k0 - 15;    // high half: their mirrors, ascending (lane l mirrors lane 31 - l)
        int k = h ? kh0 + lp : k0 + lp;
        if (SPECIAL && lane == 16) k = 0;
        float2 twb[31];
#pragma unroll
        for (int tl = 0; tl < 5; tl++)
#pragma unroll
            for (int c = 0; c < (1 << tl); c++) twb[(1 << tl) - 1 + c] = KOFFT_LDG(table + ((long)(k + (c << LA)) << (4 - tl)));
        float2 *tb = smem + OFF_XB + w * XB + h * HB;
        float2 *rtws = smem + OFF_RTW + w * (32 * 32) + lane;
        if constexpr (EPI == SPLIT_TWIST) {
#pragma unroll
            for (int wi = 0; wi < 32; wi++) rtws[wi * 32] = KOFFT_LDG(io.rtw + k + ((long)bitrev(wi, 5) << LA));
        }
        warp_sync();
        // 16-byte asynchronous copies (LDGSTS, L2 only): instruction j fetches row j of each half (256 bytes per
        // half warp) straight into the transposition region
        auto fetch = [&](long i) {
            const float2 *slot = slots + (i % SLOTS) * n;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                int rj = h ? kh0 + j : k0 + j;
                if (SPECIAL && j == 0 && h) rj = 0;
                cp_async16(tb + j * RSB + 2 * lp, slot + (long)rj * 32 + 2 * lp);
            }
            cp_async_commit();
        };
        // PRE: this warp's share of row j's untwist -- bins k = 512 wt + lane + 32 jj (jj < 16) and their mirrors n - k,
        // both from the same two inputs; coalesced 256-byte requests.  Bin 0 (X[0], X[n]) and bin n/2 (its own mirror)
        // belong to the team's first lane.  The input bins arrive by 8-byte asynchronous copies (the rows of n + 1
        // bins are only 8-byte aligned) issued one transform earlier, so their HBM latency is not waited for.
        float2 *xr = smem + OFF_RTW + w * (32 * 32) + lane;
        const int kfirst = wt * 512 + lane;
        auto fetch_bins = [&](long j) {
            if constexpr (PRE) {
                const float2 *X = io.in + (team + j * teams) * (n + 1);
#pragma unroll
                for (int jj = 0; jj < 16; jj++) {
                    const int kk = kfirst + jj * 32;
                    cp_async8(xr + jj * 32, X + kk);
                    cp_async8(xr + 512 + jj * 32, X + ((int)n - kk));
                }
                cp_async_commit();
            }
        };
        // tile_pending: the copies of the current tile were committed after the bins' (they are the newest group)
        auto untwist_row = [&](long j, bool tile_pending) {
            if constexpr (PRE) {
                unsigned *cntZ = cntB + SLOTS, *cntZf = cntZ + ZSLOTS;
                L2Policy pol = make_l2_policy();
                if (!KOFFT_SPLIT_ZHINT) pol.last = 0x1000000000000000ull; // evict_normal
                float2 *Z = zs + (j % ZSLOTS) * n;
                // the warp's 32 entries of T' (L2): requested before anything is waited for
                float2 ta[16], tm[16];
#pragma unroll
                for (int jj = 0; jj < 16; jj++) {
                    const int kk = kfirst + jj * 32;
                    ta[jj] = KOFFT_LDG(io.rtw + kk);
                    tm[jj] = KOFFT_LDG(io.rtw + (kk == 0 ? 0 : (int)n - kk)); // T' has n entries
                }
                if (j >= ZSLOTS) {
                    const unsigned want = goal_zfree(j);
                    while (flag_load(cntZf + j % ZSLOTS) < want) nano_sleep(64);
                }
                if (tile_pending)
                    cp_async_wait_but<1>();
                else
                    cp_async_wait_all();
#pragma unroll
                for (int jj = 0; jj < 16; jj++) {
                    const int kk = kfirst + jj * 32;
                    const float2 a = xr[jj * 32], b = xr[512 + jj * 32];
                    if (kk != 0) {
                        stg_hint(Z + kk, io.untwist(a, b, ta[jj]), pol.last);
                        stg_hint(Z + ((int)n - kk), io.untwist(b, a, tm[jj]), pol.last);
                    } else {
                        const float2 xh = KOFFT_LDG(io.in + (team + j * teams) * (n + 1) + n / 2);
                        stg_hint(Z, io.untwist0(a, b), pol.last);
                        stg_hint(Z + n / 2, io.untwist(xh, xh, KOFFT_LDG(io.rtw + n / 2)), pol.last);
                    }
                }
            }
        };
        // the release fence of the announcement waits for the warp's stores: `when` picks where in the iteration it sits
        auto announce_row = [&](long j, int when) {
            if constexpr (PRE) {
                if (when == KOFFT_SPLIT_ZDEFER) {
                    warp_sync(); // orders the warp's stores before lane 0's release
                    if (lane == 0) flag_arrive(cntB + SLOTS + j % ZSLOTS);
                }
            }
        };
        if constexpr (PRE) {
            for (long j = 0; j < ZAHEAD && j < cnt; j++) {
                fetch_bins(j);
                untwist_row(j, false);
                announce_row(j, KOFFT_SPLIT_ZDEFER);
            }
            if (ZAHEAD < cnt) fetch_bins(ZAHEAD);
        }
        bool fetched = false;
#pragma unroll(UNROLL_B)
        for (long i = 0; i < cnt; i++) {
            const long row = team + i * teams;
            const bool ahead = PRE && i + ZAHEAD < cnt;
            if (ahead) {
                // the tile's copies first if pass A is already done with it: their latency hides behind the untwist
                if (!fetched && warp_all(flag_load(cntA + i % SLOTS) >= goal_a(i))) {
                    (void)flag_load_acquire(cntA + i % SLOTS);
                    fetch(i);
                    fetched = true;
                }
                untwist_row(i + ZAHEAD, fetched);
                announce_row(i + ZAHEAD, 0);
            }
            if (!fetched) {
                // every lane polls (one broadcast request per iteration): the warp stays converged, which keeps the
                // shuffles below plain SHFL instructions (behind a one-lane polling loop the compiler guards every
                // shuffle with a convergence sequence: 2600 extra instructions per tile, profiles/r03a)
                const unsigned want = goal_a(i);
                while (KOFFT_SPLIT_ONLY_ROLE != 2 && flag_load(cntA + i % SLOTS) < want) nano_sleep(KOFFT_SPLIT_POLL_NS);
                (void)flag_load_acquire(cntA + i % SLOTS);
                fetch(i);
            }
            if (PRE && i + ZAHEAD + 1 < cnt) { // behind the tile's copies, so that only those are waited for
                fetch_bins(i + ZAHEAD + 1);
                cp_async_wait_but<1>();
            } else {
                cp_async_wait_all();
            }
            warp_sync();
            if (lane == 0) flag_arrive_relaxed(cntB + i % SLOTS); // the warp's reads of the slot are complete
            float2 x[WIDE];
#pragma unroll
            for (int q2 = 0; q2 < 16; q2++) {
                const float4 e = *reinterpret_cast<const float4 *>(tb + lp * RSB + 2 * q2);
                x[2 * q2] = make_float2(e.x, e.y);
                x[2 * q2 + 1] = make_float2(e.z, e.w);
            }
            warp_sync(); // the region is free again
            if (ahead) announce_row(i + ZAHEAD, 2);
            // The next tile's rows, if its pass A is already complete: their L2 latency hides behind this tile's epilogue.
            // The flag is loaded here and looked at after the register pass, so its own L2 round trip is hidden too.
            const unsigned nxt = i + 1 < cnt ? flag_load(cntA + (i + 1) % SLOTS) : 0u;
            PB::compute(x, twb);
            if (ahead) announce_row(i + ZAHEAD, 1);
            fetched = false;
            if (i + 1 < cnt) {
                if (warp_all(nxt >= goal_a(i + 1))) {
                    (void)flag_load_acquire(cntA + (i + 1) % SLOTS);
                    fetch(i + 1);
                    fetched = true;
                }
            }
            // register wi holds bin K = k + bitrev(wi) 2^LA
            if constexpr (EPI == SPLIT_TWIST) {
                float2 *o = io.out + row * (io.m + 1) + k;
#pragma unroll
                for (int wi = 0; wi < 32; wi++) {
                    const float2 p = x[31 - wi];
                    float2 ym = make_float2(shfl_xor_f(p.x, 31), shfl_xor_f(p.y, 31));
                    if (SPECIAL) { // selects, not branches: the warp stays converged for the shuffles
                        // row 2^(LA-1) (lane 15): m - K = 2^(LA-1) + (31 - c) 2^LA, its own register 31 - wi;
                        // row 0 (lane 16): m - K = (32 - c) 2^LA
                        const float2 q = x[bitrev((32 - bitrev(wi, 5)) & 31, 5)];
                        ym.x = lane == 15 ? p.x : (lane == 16 ? q.x : ym.x);
                        ym.y = lane == 15 ? p.y : (lane == 16 ? q.y : ym.y);
                    }
                    const float2 tw = rtws[wi * 32];
                    if (SPECIAL && wi == 0) {
                        // lane 16 holds row 0: bins 0 and m (src/rfft.rs:450-452)
                        float2 v0 = io.twist(x[0], ym, tw);
                        if (lane == 16) v0 = make_float2(add_rn(x[0].x, x[0].y), 0.0f);
                        o[0] = v0;
                        if (lane == 16) o[io.m] = make_float2(sub_rn(x[0].x, x[0].y), 0.0f);
                    } else {
                        o[(long)bitrev(wi, 5) << LA] = io.twist(x[wi], ym, tw);
                    }
                }
            } else {
#pragma unroll
                for (int wi = 0; wi < 32; wi++) {
                    const int K = k + (bitrev(wi, 5) << LA);
                    io.store(row, K, x[wi]);
                }
            }
        }
    }

    // rows: transforms in the batch; scratch: teams * SLOTS * 2^L complex; flags: teams * FLAG_STRIDE
    // counters, zero at launch.  gridDim.x is a multiple of NT and every CTA is resident.
    static KD void run(const IO &io, const Tw0W &tw0, const float2 *__restrict__ table, long rows,
                       float2 *__restrict__ scratch, float2 *smem, unsigned *flags, const TmaMap *map = nullptr)
    {
        const int tid = threadIdx.x;
        const long n = 1L << L;
        const int kb = blockIdx.x % NT;
        const long team = blockIdx.x / NT, teams = gridDim.x / NT;
        unsigned *cntA = flags + team * FLAG_STRIDE, *cntB = cntA + SLOTS;
        const long cnt = team < rows ? (rows - team + teams - 1) / teams : 0;
        float2 *slots = scratch + team * SLOTS * n;
        float2 *zs = PRE ? scratch + teams * SLOTS * n + team * ZSLOTS * n : nullptr; // behind every team's slots
        if (tid < A_THREADS) {
            setmaxnreg_dec<KOFFT_SPLIT_REGS_A>();
            if (KOFFT_SPLIT_ONLY_ROLE == 2) return;
            if constexpr (STAGED)
                a_role_staged(io, tw0, table, cnt, team, teams, kb, slots, smem, cntA, cntB, tid, map,
                              reinterpret_cast<unsigned long long *>(smem + OFF_BAR));
            else
                a_role(io, tw0, table, cnt, team, teams, kb, slots, smem, cntA, cntB, tid);
        } else {
            setmaxnreg_inc<KOFFT_SPLIT_REGS_B>();
            if (KOFFT_SPLIT_ONLY_ROLE == 1) return;
            const int wl = tid - A_THREADS;
            if (kb == NT - 1 && (wl >> 5) == B_WARPS - 1)
                b_role<true>(io, table, cnt, team, teams, kb, slots, smem, cntA, cntB, wl, zs);
            else
                b_role<false>(io, table, cnt, team, teams, kb, slots, smem, cntA, cntB, wl, zs);
        }
    }
};

#ifdef __CUDACC__
template <int LA, bool EXACT, class IO, int EPI, bool STAGED, bool PRE = false>
__global__ void __launch_bounds__(512, 1)
    split32_kernel(const __grid_constant__ IO io, const __grid_constant__ Tw0W tw0, const float2 *__restrict__ table,
                   long rows, float2 *__restrict__ scratch, unsigned *flags, const __grid_constant__ TmaMap map)
{
    extern __shared__ __align__(128) float2 smem[];
    Split32<LA, EXACT, IO, EPI, STAGED, PRE>::run(io, tw0, table, rows, scratch, smem, flags, &map);
}
#endif

} // namespace kofft
