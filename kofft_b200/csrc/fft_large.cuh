// fft_large.cuh -- N = 2^15 .. 2^16 (and the half-length core of rfft 2^16 .. 2^17): the
// faithful two-pass split of kofft's radix-2 Stockham.
//
//   pass A ("column pass") = stages 0 .. 7: for every column j' (low L-8 index bits) a
//       256-point transform over the high 8 bits with stride 2^(L-8).  It starts at stage 0,
//       so its twiddles are the ordinary ones of a 256-point transform read from the big
//       table with stride 2^(L-8).  A CTA takes 16 adjacent columns; threads are mapped
//       column-fastest so that every global access is a 128-byte run.
//   pass B ("row pass") = stages 8 .. L-1: for every k (the 8 output bits produced by pass A)
//       a contiguous 2^(L-8)-point transform whose twiddles depend on k:
//       T[(k + (k_local + c_low 2^s) 2^8) << (L-9-s-t)], and whose bin c lands at K = k + 256 c.
//       A CTA takes 16 or 32 values of k, so after a transpose through shared memory every
//       global store is again a 128-byte run.  The inter-step twiddle of a textbook four-step
//       FFT does not exist here: it is already inside pass B's stage twiddles.
//
// The intermediate lives in a scratch buffer that the host sizes to stay L2-resident (the
// batch is processed in chunks), so HBM sees one read and one write of the data.
//
// rfft: the Hermitian twist needs bins K and m-K together.  With K = k + 256 c the mirror of
// (k, c) is (256-k, NB-1-c), so the row-pass CTA is given k-blocks paired with their mirrors
// ({0..h-1} with {128, 255..257-h}, {bh..bh+h-1} with {256-bh .. 257-bh-h}) and twists in
// shared memory before the store (src/rfft.rs:450-463).
#pragma once
#include "fft_kernels.cuh"

namespace kofft {

constexpr int LARGE_S1 = 8; // stages in pass A

// ---------------------------------------------------------------------------------------------
// pass A
// ---------------------------------------------------------------------------------------------
template <bool EXACT, class IO>
struct ColPass {
    using P = Plan<LARGE_S1>; // 256-point engine, passes (4, 4), 16 threads per column
    using P0 = Pass<P, 0, EXACT>;
    using P1 = Pass<P, 1, EXACT>;
    static constexpr int COLS = 16;            // columns per CTA tile
    static constexpr int RS = P::PADN + 1;     // odd region stride: column-fastest threads hit distinct banks
    static constexpr int SMEM_BYTES = 2 * COLS * RS * 8;
    static constexpr int STAGE_BYTES = 256 * COLS * 8; // next tile's raw input: [256 rows][16 columns]
    static constexpr int SMEM_BYTES_STAGED = SMEM_BYTES + STAGE_BYTES;
    static_assert(P::T == 16 && P::CTA == 256 && P::TPC == COLS, "tile shape");

    // one tile: 16 adjacent columns starting at j0 of transform `row` (io) -> scratch_row.
    // bf: this thread's exchange region (column-fastest layout), tw1: pass-1 twiddles.
    static KD void tile(const IO &io, const Tw0 &tw0, const float2 *tw1, int lsub, long row, long j0,
                        float2 *__restrict__ scratch_row, float2 *bf, int t, int slot)
    {
        const long j = j0 + slot;
        float2 x[EPT];
#pragma unroll
        for (int q = 0; q < P0::R; q++) x[q] = io.load(row, (int)(((long)P0::src_index(t, 0, q) << lsub) + j));
        P0::compute(x, tw0.v);
#pragma unroll
        for (int w = 0; w < P0::R; w++) bf[P0::dst_pad(P0::dst_base(t, 0), w)] = x[w];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < P1::R; q++) x[q] = bf[P1::src_pad(P1::src_base(t, 0), q)];
        P1::compute(x, tw1);
        float2 *o = scratch_row + j;
#pragma unroll
        for (int w = 0; w < P1::R; w++) o[(long)P1::dst_index(t, 0, w) << lsub] = x[w];
    }

    // n = 2^L elements per transform, lsub = L - 8, tiles = batch * (2^lsub / 16)
    // row0: index of the chunk's first transform in the caller's batch (scratch is chunk-local)
    // STAGED (plain contiguous rows only): the CTA's next tile -- 256 segments of 128 bytes -- is
    // fetched with 16-byte asynchronous copies (cp.async / LDGSTS) while the current tile is transformed.
    template <bool STAGED = false>
    static KD void run(const IO &io, const Tw0 &tw0, const float2 *__restrict__ table, int lsub, long tiles,
                       long row0, float2 *__restrict__ scratch, float2 *smem)
    {
        const int tid = threadIdx.x;
        const int slot = tid & (COLS - 1); // column within the tile (fastest)
        const int t = tid >> 4;            // thread within the column's 256-point transform
        float2 *stage = smem;
        float2 *xch = STAGED ? smem + STAGE_BYTES / 8 : smem;
        float2 *buf0 = xch + slot * RS;
        float2 *buf1 = buf0 + COLS * RS;
        int par = 0;
        TwMap map;
        map.sh2 = lsub; // 256-point twiddles = every 2^lsub-th entry of the big table
        float2 tw1[P1::NTW];
        P1::load_tw(table, t, tw1, map);
        const int ltiles = lsub - 4; // log2 tiles per transform
        const long n = 1L << (LARGE_S1 + lsub);
        if constexpr (STAGED && IoTraits<IO>::kRowPtr) {
            // thread i copies 16-byte piece (i & 7) of rows (i >> 3) + 32 m, m = 0..7
            auto prefetch = [&](long tl) {
                const long b = tl >> ltiles;
                const long j0 = (tl & ((1L << ltiles) - 1)) << 4;
                const float2 *src = io.row_ptr(row0 + b) + j0 + 2 * (tid & 7);
#pragma unroll
                for (int m = 0; m < 8; m++) {
                    const int r = (tid >> 3) + 32 * m;
                    cp_async16(stage + r * COLS + 2 * (tid & 7), src + ((long)r << lsub));
                }
                cp_async_commit();
            };
            if ((long)blockIdx.x < tiles) prefetch(blockIdx.x);
            for (long tl = blockIdx.x; tl < tiles; tl += gridDim.x) {
                const long b = tl >> ltiles;
                const long j0 = (tl & ((1L << ltiles) - 1)) << 4;
                cp_async_wait_all();
                __syncthreads(); // every thread's pieces of this tile have landed
                float2 *bf = par ? buf1 : buf0;
                par ^= 1;
                const long j = j0 + slot;
                float2 x[EPT];
#pragma unroll
                for (int q = 0; q < P0::R; q++) x[q] = io.from_raw(stage[P0::src_index(t, 0, q) * COLS + slot]);
                P0::compute(x, tw0.v);
#pragma unroll
                for (int w = 0; w < P0::R; w++) bf[P0::dst_pad(P0::dst_base(t, 0), w)] = x[w];
                __syncthreads(); // exchange; also: the stage has been consumed by everyone
                if (tl + gridDim.x < tiles) prefetch(tl + gridDim.x);
#pragma unroll
                for (int q = 0; q < P1::R; q++) x[q] = bf[P1::src_pad(P1::src_base(t, 0), q)];
                P1::compute(x, tw1);
                float2 *o = scratch + b * n + j;
#pragma unroll
                for (int w = 0; w < P1::R; w++) o[(long)P1::dst_index(t, 0, w) << lsub] = x[w];
            }
        } else {
            for (long tl = blockIdx.x; tl < tiles; tl += gridDim.x) {
                const long b = tl >> ltiles;
                const long j0 = (tl & ((1L << ltiles) - 1)) << 4;
                tile(io, tw0, tw1, lsub, row0 + b, j0, scratch + b * n, par ? buf1 : buf0, t, slot);
                par ^= 1;
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------
// pass B
// ---------------------------------------------------------------------------------------------
enum RowEpilogue : int { ROW_STORE = 0, ROW_TWIST = 1 };

template <int LB, bool EXACT, class IO, int EPI>
struct RowPass {
    using P = Plan<LB>;
    static_assert(P::NP == 2, "row pass covers sub-lengths 128 and 256");
    using P0 = Pass<P, 0, EXACT, false>; // k != 0: no unit twiddles, per-thread pass-0 twiddles
    using P1 = Pass<P, 1, EXACT, false>;
    static constexpr int NB = P::N;
    static constexpr int TPC = P::TPC;          // sub-transforms (values of k) per CTA: 32 or 16
    static constexpr int NKB = 256 / TPC;       // k-blocks per transform
    static constexpr int RS = P::PADN + 1;      // odd region stride for the transposed read-back
    // Exchange regions: slot s starts at slot_off(s).  Sub-transforms of 128 points use 8 threads, so
    // a half-warp spans two slots: threads are mapped so that these are slots s and s + 16, whose
    // regions differ by 8 (mod 16) eight-byte banks -> the per-slot padded layout stays conflict-free
    // across the pair, while 16 consecutive slots (the transposed read-back) still differ by the odd RS.
    static constexpr bool PAIRED = P::T < 16;
    static_assert(P::T == 8 || P::T == 16, "row pass sub-transforms are 128 or 256 points");
    static KHD int slot_of(int tid) { return PAIRED ? (tid >> 4) + 16 * ((tid >> 3) & 1) : tid / P::T; }
    static KHD int slot_off(int s) { return s * RS + (PAIRED ? 8 * (s >> 4) : 0); }
    static constexpr int BUF = TPC * RS + 16;       // float2 per exchange buffer
    static constexpr int NBUFS = EPI == ROW_TWIST ? 3 : 2; // twist: + the CTA's rfft twiddles, same layout
    static constexpr int SMEM_BYTES = NBUFS * BUF * 8;
    static constexpr int STAGE_BYTES = TPC * NB * 8; // next tile's TPC sub-transforms, unpadded
    static constexpr int SMEM_BYTES_STAGED = SMEM_BYTES + STAGE_BYTES;
    static constexpr int HALF = TPC / 2;

    // STAGED: where the slot's sub-transform sits in the stage.  Plain store: slot order.  Twist:
    // the mirrored half {256 - HALF kb - i} is one ascending run of k, so it is staged ascending
    // (one bulk copy) and read back reversed; k = 128 of block 0 takes the last stage row.
    static KHD int stage_row(int slot) { return (EPI == ROW_TWIST && slot >= HALF) ? TPC - 1 - (slot - HALF) : slot; }

    // which k the CTA's slot handles in block kb
    static KHD int kmap(int kb, int slot)
    {
        if (EPI == ROW_TWIST) {
            if (slot < HALF) return HALF * kb + slot;
            int k = 256 - HALF * kb - (slot - HALF);
            return k == 256 ? 128 : k;
        }
        return kb * TPC + slot;
    }
    // slot holding the mirror 256 - k of the slot's k (twist only)
    static KHD int mirror_slot(int kb, int slot)
    {
        if (slot < HALF) return (kb == 0 && slot == 0) ? 0 : slot + HALF;
        return (kb == 0 && slot == HALF) ? HALF : slot - HALF;
    }

    // one tile: the TPC sub-transforms of k-block kb of one transform, read from scratch_row
    // (LDCG: the intermediate was written by other CTAs), stored / twisted through io as `row`.
    // bfa/bfb: this thread's two exchange regions, alla/allb: the same buffers seen CTA-wide.
    // staged_in != nullptr: the slot's sub-transform is already in shared memory (TMA bulk copy);
    // after_sync(): called by every thread right after the first barrier (the stage is free again)
    template <class AfterSync>
    static KD void tile(const IO &io, const float2 *tw0, const float2 *tw1, long row, int kb, int k,
                        const float2 *__restrict__ scratch_row, float2 *bfa, float2 *bfb, const float2 *allb, int t,
                        int tid, const float2 *staged_in, AfterSync after_sync, const float2 *rtwb = nullptr)
    {
        float2 x[EPT];
        if (staged_in) {
#pragma unroll
            for (int u = 0; u < P0::U; u++)
#pragma unroll
                for (int q = 0; q < P0::R; q++) x[u * P0::R + q] = staged_in[P0::src_index(t, u, q)];
        } else {
            const float2 *in = scratch_row + (long)k * NB;
#pragma unroll
            for (int u = 0; u < P0::U; u++)
#pragma unroll
                for (int q = 0; q < P0::R; q++) x[u * P0::R + q] = KOFFT_LDCG(in + P0::src_index(t, u, q));
        }
        P0::compute(x, tw0);
#pragma unroll
        for (int u = 0; u < P0::U; u++)
#pragma unroll
            for (int w = 0; w < P0::R; w++) bfa[P0::dst_pad(P0::dst_base(t, u), w)] = x[u * P0::R + w];
        __syncthreads();
        after_sync();
#pragma unroll
        for (int q = 0; q < P1::R; q++) x[q] = bfa[P1::src_pad(P1::src_base(t, 0), q)];
        P1::compute(x, tw1);
        // transpose through shared memory: bins of all TPC sub-transforms, k fastest
#pragma unroll
        for (int w = 0; w < P1::R; w++) bfb[P1::dst_pad(P1::dst_base(t, 0), w)] = x[w];
        __syncthreads();
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            const int flat = e * P::CTA + tid;
            const int s2 = flat % TPC, c = flat / TPC;
            const long K = kmap(kb, s2) + ((long)c << LARGE_S1);
            const float2 a = allb[slot_off(s2) + pad(c)];
            if constexpr (EPI == ROW_TWIST) {
                const int kk = kmap(kb, s2);
                float2 ym;
                if (kk == 0)
                    ym = c == 0 ? a : allb[slot_off(s2) + pad(NB - c)]; // m - K = 256 (NB - c)
                else
                    ym = allb[slot_off(mirror_slot(kb, s2)) + pad(NB - 1 - c)];
                io.twist_store_tw(row, K, a, ym, rtwb[slot_off(s2) + pad(c)]);
            } else {
                io.store(row, (int)K, a);
            }
        }
    }

    // the CTA's share of the rfft table T' (src/rfft.rs:172-183), laid out like the transposed bins:
    // a CTA keeps its k-block, so this is loaded once per CTA, not once per element per tile
    static KD void load_rtw(const IO &io, int kb, int tid, float2 *rtwb)
    {
        if constexpr (EPI == ROW_TWIST) {
#pragma unroll
            for (int e = 0; e < EPT; e++) {
                const int flat = e * P::CTA + tid;
                const int s2 = flat % TPC, c = flat / TPC;
                rtwb[slot_off(s2) + pad(c)] = KOFFT_LDG(io.rtw + kmap(kb, s2) + ((long)c << LARGE_S1));
            }
        }
    }

    // tiles = batch * NKB; the grid is a multiple of NKB so a CTA keeps its k-block (and with it
    // its twiddles, which live in registers) for every tile it processes.
    // one thread: bulk copies of the TPC sub-transforms of k-block kb of one transform into the stage
    static KD void stage_issue(const float2 *scratch_row, int kb, float2 *stage, unsigned long long *bar)
    {
        mbar_expect_tx(bar, (unsigned)STAGE_BYTES);
        if (EPI == ROW_TWIST) {
            bulk_copy_g2s(stage, scratch_row + (long)(HALF * kb) * NB, HALF * NB * 8, bar);
            if (kb == 0) { // k = 128, then 255 .. 257 - HALF (ascending: 257 - HALF .. 255)
                bulk_copy_g2s(stage + (long)HALF * NB, scratch_row + (long)(257 - HALF) * NB, (HALF - 1) * NB * 8, bar);
                bulk_copy_g2s(stage + (long)(TPC - 1) * NB, scratch_row + 128L * NB, NB * 8, bar);
            } else {
                bulk_copy_g2s(stage + (long)HALF * NB, scratch_row + (long)(257 - HALF * kb - HALF) * NB, HALF * NB * 8, bar);
            }
        } else {
            bulk_copy_g2s(stage, scratch_row + (long)(kb * TPC) * NB, STAGE_BYTES, bar);
        }
    }

    // STAGED: the CTA's next tile is fetched from the (L2-resident) intermediate with TMA bulk
    // copies while the current one is transformed, twisted and stored.
    template <bool STAGED = false>
    static KD void run(const IO &io, const float2 *__restrict__ table, long tiles, long row0,
                       const float2 *__restrict__ scratch, float2 *smem_all)
    {
        const int tid = threadIdx.x;
        const int slot = slot_of(tid);
        const int t = tid & (P::T - 1);
        const int kb = blockIdx.x % NKB;
        const int k = kmap(kb, slot);
        float2 *stage = smem_all;
        float2 *smem = STAGED ? smem_all + STAGE_BYTES / 8 : smem_all;
        __shared__ __align__(8) unsigned long long mbar;
        unsigned phase = 0;
        if constexpr (STAGED) {
            if (tid == 0) {
                mbar_init(&mbar, 1);
                fence_mbar_init();
            }
            __syncthreads();
        }
        float2 *buf0 = smem + slot_off(slot);
        float2 *buf1 = buf0 + BUF;
        const float2 *rtwb = smem + 2 * BUF;
        load_rtw(io, kb, tid, smem + 2 * BUF); // made visible by the first tile's barriers
        TwMap map;
        map.k0 = k;
        map.sh1 = LARGE_S1;
        float2 tw0[P0::NTW], tw1[P1::NTW];
        P0::load_tw(table, t, tw0, map);
        P1::load_tw(table, t, tw1, map);
        const long n = (long)NB << LARGE_S1;
        if constexpr (STAGED) {
            if (tid == 0 && (long)blockIdx.x < tiles) stage_issue(scratch + (blockIdx.x / NKB) * n, kb, stage, &mbar);
        }
        for (long tl = blockIdx.x; tl < tiles; tl += gridDim.x) {
            const long b = tl / NKB;
            // buffer A of this tile was last read before the previous tile's second barrier and
            // buffer B before this tile's first barrier, so two buffers suffice
            if constexpr (STAGED) {
                mbar_wait(&mbar, phase);
                phase ^= 1;
                const long nxt = tl + gridDim.x;
                tile(io, tw0, tw1, row0 + b, kb, k, scratch + b * n, buf0, buf1, smem + BUF, t, tid,
                     stage + (long)stage_row(slot) * NB, [&] {
                         if (tid == 0 && nxt < tiles) stage_issue(scratch + (nxt / NKB) * n, kb, stage, &mbar);
                     }, rtwb);
            } else {
                tile(io, tw0, tw1, row0 + b, kb, k, scratch + b * n, buf0, buf1, smem + BUF, t, tid,
                     (const float2 *)nullptr, [] {}, rtwb);
            }
            __syncthreads();
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Both passes in ONE persistent kernel: a thread-block cluster of NKB CTAs owns one transform at
// a time.  Each CTA runs its share of the column tiles, the cluster synchronises
// (barrier.cluster, release/acquire), then CTA `rank` runs k-block `rank` of the row pass.  The
// intermediate goes through a small per-cluster global scratch (double-buffered, a few MB in
// total, so it never leaves L2) and is read back with ld.global.cg.  One launch per batch, no
// chunking, no launch gaps, and the row-pass twiddles stay in registers.
// ---------------------------------------------------------------------------------------------
template <int LB, bool EXACT, class IO, int EPI>
struct LargeFused {
    using C = ColPass<EXACT, IO>;
    using R = RowPass<LB, EXACT, IO, EPI>;
    static constexpr int CLUSTER = R::NKB;             // 8 (N = 2^15) or 16 (N = 2^16) CTAs
    static constexpr int NTA = (1 << LB) / C::COLS;    // column tiles per transform (== CLUSTER)
    static constexpr int TW_SMEM = 16 * 16;            // pass-A pass-1 twiddles: [t][15] float2
    static constexpr int XCHG = (C::SMEM_BYTES > R::SMEM_BYTES ? C::SMEM_BYTES : R::SMEM_BYTES);
    static constexpr int SMEM_BYTES = XCHG + TW_SMEM * 8;
    static_assert(NTA == CLUSTER, "one column tile and one k-block per CTA");

    // scratch: clusters * 2 * n complex.  rows: transforms in the batch.
    static KD void run(const IO &io, const Tw0 &tw0, const float2 *__restrict__ table, long rows,
                       float2 *__restrict__ scratch, float2 *smem, int rank, long cluster_id, long nclusters)
    {
        const int tid = threadIdx.x;
        const long n = 1L << (LARGE_S1 + LB);
        // pass A (column-fastest mapping)
        const int slotA = tid & (C::COLS - 1), tA = tid >> 4;
        float2 *twA = smem + XCHG / 8; // shared copy of the 16 x 15 pass-1 twiddles of the 256-point pass
        if (tid < 16) {
            TwMap map;
            map.sh2 = LB;
            float2 tmp[C::P1::NTW];
            C::P1::load_tw(table, tid, tmp, map);
#pragma unroll
            for (int i = 0; i < C::P1::NTW; i++) twA[tid * 16 + i] = tmp[i];
        }
        // pass B (sub-transform-major mapping); k-block == cluster rank, twiddles in registers
        const int slotB = R::slot_of(tid), tB = tid & (R::P::T - 1);
        const int kb = rank, k = R::kmap(kb, slotB);
        TwMap mapB;
        mapB.k0 = k;
        mapB.sh1 = LARGE_S1;
        float2 twB0[R::P0::NTW], twB1[R::P1::NTW];
        R::P0::load_tw(table, tB, twB0, mapB);
        R::P1::load_tw(table, tB, twB1, mapB);
        __syncthreads();

        float2 *bufA0 = smem + slotA * C::RS;
        float2 *bufB0 = smem + R::slot_off(slotB), *bufB1 = bufB0 + R::BUF;
        R::load_rtw(io, kb, tid, smem + 2 * R::BUF); // beyond both passes' exchange regions
        int it = 0;
        for (long b = cluster_id; b < rows; b += nclusters, it ^= 1) {
            float2 *sc = scratch + (cluster_id * 2 + it) * n;
            C::tile(io, tw0, twA + tA * 16, LB, b, (long)rank * C::COLS, sc, bufA0, tA, slotA);
            cluster_sync(); // every column tile of this transform is in `sc` (also a CTA barrier)
            R::tile(io, twB0, twB1, b, kb, k, sc, bufB0, bufB1, smem + R::BUF, tB, tid, (const float2 *)nullptr, [] {}, smem + 2 * R::BUF);
            __syncthreads(); // smem is reused by the next transform's column tile
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Both passes in ONE persistent kernel, software-pipelined per transform with dependency flags.
// A "team" of NKB consecutive CTAs (one per column tile / k-block) owns the transforms
// q, q + teams, q + 2 teams, ...; member kb runs, for its i-th transform,
//     step i:   pass A tile of transform i+2  (HBM -> intermediate slot (i+2) mod 4)  ... arrive A
//               pass B tile of transform i    (intermediate slot i mod 4 -> HBM)      ... arrive B
// Pass B of transform i needs the NKB pass-A tiles of that transform (its inputs are requested one
// step early, behind the previous tile's epilogue), pass A of transform i+2 overwrites the slot of
// transform i-2 and needs its NKB pass-B tiles to have consumed it: one arrival counter per team,
// slot and pass.  Both conditions were met about one step earlier unless a member lags by more than a step, and
// the polling thread issues its flag load a register pass ahead of the check, so neither the
// skew between CTAs nor the flag's L2 round trip is normally exposed.  There is no grid-wide
// synchronisation: a slow SM delays only its own team (the first version of this kernel used a
// grid barrier per batch chunk and lost 14 % of the warp time to it, profiles/r01x).  The
// cooperative launch guarantees that all CTAs are co-resident, which the flag waits rely on.
//
// The intermediate is teams * 4 * 2^L complex (37 MB for N = 2^15): written with the L2
// evict_last policy while the rows stream through with evict_first where the instruction allows,
// so it never makes the round trip to HBM (the two-kernel path above needs long launches to
// amortise launch + ramp + tail, and at those chunk sizes the intermediate spills: 2x the
// algorithmic DRAM traffic, profiles/r01u_rfft_*pass_kernel.json).
//
// CTA j keeps column tile / k-block  j mod NKB  for its lifetime, so its pass-B twiddles stay in
// registers and its rfft-table slice in shared memory.  Shared memory: two exchange buffers (+ the
// T' slice for the twist).  Pass A exchanges through buf0 and (STAGED) reads its tile from buf1,
// where 16-byte asynchronous copies put it while the previous pass-B tile ran its epilogue; pass B
// exchanges through buf1 (between its register passes) and buf0 (transposed bins).
// ---------------------------------------------------------------------------------------------
// pipeline depth of the persistent kernel: build-time knobs (scripts/build_variants.sh); the host sizes the
// intermediate from kLargePipeSlots (launch.h), which follows KOFFT_PIPE_SLOTS
#ifndef KOFFT_PIPE_AHEAD
#define KOFFT_PIPE_AHEAD 2
#endif
#ifndef KOFFT_PIPE_SLOTS
#define KOFFT_PIPE_SLOTS (KOFFT_PIPE_AHEAD + 2)
#endif

template <int LB, bool EXACT, class IO, int EPI, bool STAGED = false>
struct LargePipe {
    using C = ColPass<EXACT, IO>;
    using R = RowPass<LB, EXACT, IO, EPI>;
    static constexpr int NKB = R::NKB;                 // 8 (N = 2^15) or 16 (N = 2^16)
    static constexpr int LOG_NKB = ilog2c(NKB);
    static_assert((1 << LB) / C::COLS == NKB, "one column tile per k-block");
    static constexpr int BUFA = C::COLS * C::RS;
    static constexpr int BUF = (BUFA > R::BUF ? BUFA : R::BUF); // float2 per exchange buffer
    static_assert(BUF * 8 >= C::STAGE_BYTES, "the idle exchange buffer holds pass A's next tile");
    static constexpr int NBUFS = EPI == ROW_TWIST ? 3 : 2;      // twist: + the CTA's rfft twiddles
    static constexpr int TW_SMEM = 16 * 16;                     // pass-A pass-1 twiddles: [t][15] float2
    static constexpr int SMEM_BYTES = (NBUFS * BUF + TW_SMEM) * 8;
    static constexpr bool HINT = IoTraits<IO>::kHint;
    static_assert(!STAGED || IoTraits<IO>::kRowPtr, "staging needs plain contiguous rows");

    // thread i copies 16-byte piece (i & 7) of rows (i >> 3) + 32 m, m = 0..7, of the tile
    // [256 rows][16 columns] starting at column j0 of transform `row`
    static KD void prefetch_a(const IO &io, long row, long j0, float2 *stage, int tid, const L2Policy &pol)
    {
        if constexpr (STAGED) {
            const float2 *src = io.row_ptr(row) + j0 + 2 * (tid & 7);
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const int r = (tid >> 3) + 32 * m;
                cp_async16(stage + r * C::COLS + 2 * (tid & 7), src + ((long)r << LB)); // no L2 hint: see cp_async16_hint
            }
            cp_async_commit();
        }
    }

    // pass A tile: 16 adjacent columns of transform `row` -> scratch_row.  STAGED: the tile is in
    // `stage`.  before_store(): called by every thread before the exchange barrier -- the polling
    // thread returns from it only when the destination slot may be overwritten.
    template <class Before>
    static KD void tile_a(const IO &io, const Tw0 &tw0, const float2 *tw1, long row, long j0,
                          float2 *__restrict__ scratch_row, float2 *bf, const float2 *stage, int t, int slot,
                          const L2Policy &pol, Before before_store)
    {
        using P0 = typename C::P0;
        using P1 = typename C::P1;
        const long j = j0 + slot;
        float2 x[EPT];
        if constexpr (STAGED) {
            cp_async_wait_all();
            __syncthreads(); // every thread's pieces have landed
#pragma unroll
            for (int q = 0; q < P0::R; q++) x[q] = io.from_raw(stage[P0::src_index(t, 0, q) * C::COLS + slot]);
        } else {
#pragma unroll
            for (int q = 0; q < P0::R; q++) {
                const int idx = (int)(((long)P0::src_index(t, 0, q) << LB) + j);
                if constexpr (HINT)
                    x[q] = io.load_hint(row, idx, pol.first);
                else
                    x[q] = io.load(row, idx);
            }
        }
        P0::compute(x, tw0.v);
#pragma unroll
        for (int w = 0; w < P0::R; w++) bf[P0::dst_pad(P0::dst_base(t, 0), w)] = x[w];
        before_store();
        __syncthreads();
#pragma unroll
        for (int q = 0; q < P1::R; q++) x[q] = bf[P1::src_pad(P1::src_base(t, 0), q)];
        P1::compute(x, tw1);
        float2 *o = scratch_row + j;
#pragma unroll
        for (int w = 0; w < P1::R; w++) stg_hint(o + ((long)P1::dst_index(t, 0, w) << LB), x[w], pol.last);
    }

    // Per-thread constants of pass B's transposed epilogue.  Element e of thread tid is bin
    // c = c0 + STEP e of sub-transform slot s2, with s2 = tid mod TPC and c0 = tid / TPC fixed for the
    // thread's lifetime, so every shared-memory and output address is "base + compile-time offset".
    static constexpr int STEP = R::P::CTA / R::TPC;                        // 8 or 16 bins between a thread's elements
    static constexpr int cpad(int e) { return STEP * e + ((STEP * e) >> 4); } // pad(c0 + STEP e) - c0   (c0 < STEP <= 16)
    static constexpr int mpad(int e) { return R::NB - 2 + (R::NB >> 4) - cpad(e); } // pad(NB-1-c) + c0
    struct Epi {
        int own, mir; // float2 offsets into a buffer: slot_off(s2) + c0, slot_off(mirror) - c0
        int kk, c0;   // this thread's k and first bin
    };

    // this thread's 16 inputs of a pass-B tile (sub: the slot's contiguous sub-transform in the intermediate)
    static KD void load_b(float2 *x, const float2 *__restrict__ sub, int t, const L2Policy &pol)
    {
        using P0 = typename R::P0;
#pragma unroll
        for (int u = 0; u < P0::U; u++)
#pragma unroll
            for (int q = 0; q < P0::R; q++) x[u * P0::R + q] = ldcg_hint(sub + P0::src_index(t, u, q), pol.first);
    }

    // pass B tile.  x: the tile's inputs (load_b), replaced by the NEXT tile's inputs (next_sub, or
    // null) once the registers are free, so that their L2 latency hides behind the epilogue.
    // bfa / bfb are this thread's exchange regions (buf1 / buf0), allb is buf0 seen CTA-wide, rtwb the
    // CTA's slice of T' (twist only).  Hooks, called by every thread: after_first() right after the
    // first barrier (every thread's earlier global stores are ordered before it); before_second()
    // before the second barrier -- the polling thread returns from it only when next_sub is complete;
    // after_second() right after the second barrier (the tile's inputs are consumed, buf1 is free).
    template <class H1, class H2, class H3>
    static KD void tile_b(const IO &io, float2 *x, const float2 *tw0, const float2 *tw1, long row, int kb,
                          const float2 *__restrict__ next_sub, float2 *bfa, float2 *bfb, const float2 *allb, int t,
                          int tid, const Epi &ep, const float2 *rtwb, const L2Policy &pol, H1 after_first,
                          H2 before_second, H3 after_second)
    {
        using P0 = typename R::P0;
        using P1 = typename R::P1;
        constexpr int NB = R::NB;
        P0::compute(x, tw0);
#pragma unroll
        for (int u = 0; u < P0::U; u++)
#pragma unroll
            for (int w = 0; w < P0::R; w++) bfa[P0::dst_pad(P0::dst_base(t, u), w)] = x[u * P0::R + w];
        __syncthreads();
        after_first();
#pragma unroll
        for (int q = 0; q < P1::R; q++) x[q] = bfa[P1::src_pad(P1::src_base(t, 0), q)];
        P1::compute(x, tw1);
#pragma unroll
        for (int w = 0; w < P1::R; w++) bfb[P1::dst_pad(P1::dst_base(t, 0), w)] = x[w];
        before_second();
        __syncthreads();
        after_second();
        if (next_sub) load_b(x, next_sub, t, pol);
        // transposed read-back: bins of all TPC sub-transforms, k fastest -> 128-byte store runs
        const float2 *pa = allb + ep.own;
        if constexpr (EPI == ROW_TWIST) {
            const float2 *pr = rtwb + ep.own;
            const float2 *pm = allb + ep.mir;
            float2 *o = io.out + row * (io.m + 1) + ep.kk + ((long)ep.c0 << LARGE_S1);
            const bool mine = ep.kk != 0; // k = 0 mirrors inside its own sub-transform: fixed up below
#pragma unroll
            for (int e = 0; e < EPT; e++) {
                const float2 v = io.twist(pa[cpad(e)], pm[mpad(e)], pr[cpad(e)]);
                if (mine) {
                    if constexpr (HINT)
                        stg_hint(o + ((long)(STEP * e) << LARGE_S1), v, pol.first);
                    else
                        o[(long)(STEP * e) << LARGE_S1] = v;
                }
            }
            if (kb == 0 && tid < NB) { // k = 0 (slot 0): bin K = 256 c pairs with m - K = 256 (NB - c); bins 0 and m
                const int c = tid;
                const float2 a = allb[R::slot_off(0) + pad(c)];
                const float2 ym = c == 0 ? a : allb[R::slot_off(0) + pad(NB - c)];
                const float2 tw = rtwb[R::slot_off(0) + pad(c)];
                if constexpr (HINT)
                    io.twist_store_tw_hint(row, (long)c << LARGE_S1, a, ym, tw, pol.first);
                else
                    io.twist_store_tw(row, (long)c << LARGE_S1, a, ym, tw);
            }
        } else {
#pragma unroll
            for (int e = 0; e < EPT; e++) {
                const int K = ep.kk + ((ep.c0 + STEP * e) << LARGE_S1);
                if constexpr (HINT)
                    io.store_hint(row, K, pa[cpad(e)], pol.first);
                else
                    io.store(row, K, pa[cpad(e)]);
            }
        }
    }

    static constexpr int AHEAD = KOFFT_PIPE_AHEAD; // pass A runs this many transforms ahead of pass B
    static constexpr int SLOTS = KOFFT_PIPE_SLOTS; // intermediate slots per team
    static_assert(AHEAD >= 1 && SLOTS >= AHEAD + 2, "a slot is rewritten at least one full step after its last reader");
    static constexpr int FLAG_STRIDE = 32;  // unsigned per team: {cntA, cntB} in a 128-byte line of their own

    // rows: transforms in the batch; scratch: teams * SLOTS * n complex; flags: teams * FLAG_STRIDE
    // counters, zero at launch.  gridDim.x is a multiple of NKB and every CTA is resident.
    static KD void run(const IO &io, const Tw0 &tw0, const float2 *__restrict__ table, long rows,
                       float2 *__restrict__ scratch, float2 *smem, unsigned *flags)
    {
        const int tid = threadIdx.x;
        const long n = 1L << (LARGE_S1 + LB);
        const int kb = blockIdx.x % NKB; // column tile of pass A == k-block of pass B
        const long team = blockIdx.x / NKB, teams = gridDim.x / NKB;
        // per team and per intermediate slot: arrivals of pass-A tiles / of pass-B tiles.  One counter per
        // SLOT, not per team: a member may run up to two steps ahead of another, and a single running
        // count would let its early arrivals stand in for a late member's missing one (seen on the GPU as
        // a wrong first transform of the teams whose CTAs start in the second placement wave).  Per slot
        // this cannot happen: nobody arrives for transform i+3 before everybody has finished transform i.
        unsigned *cntA = flags + team * FLAG_STRIDE, *cntB = cntA + SLOTS;
        auto doneA = [&](long i) { return cntA + i % SLOTS; };
        auto doneB = [&](long i) { return cntB + i % SLOTS; };
        auto goal = [&](long i) { return (unsigned)(NKB * (i / SLOTS + 1)); };
        const L2Policy pol = make_l2_policy();
        // pass A (column-fastest mapping); its pass-1 twiddles are shared by the 16 columns -> shared memory
        const int slotA = tid & (C::COLS - 1), tA = tid >> 4;
        float2 *twA = smem + NBUFS * BUF;
        if (tid < 16) {
            TwMap map;
            map.sh2 = LB;
            float2 tmp[C::P1::NTW];
            C::P1::load_tw(table, tid, tmp, map);
#pragma unroll
            for (int i = 0; i < C::P1::NTW; i++) twA[tid * 16 + i] = tmp[i];
        }
        // pass B (sub-transform-major mapping); twiddles of the CTA's k-block in registers
        const int slotB = R::slot_of(tid), tB = tid & (R::P::T - 1);
        const int k = R::kmap(kb, slotB);
        TwMap mapB;
        mapB.k0 = k;
        mapB.sh1 = LARGE_S1;
        float2 twB0[R::P0::NTW], twB1[R::P1::NTW];
        R::P0::load_tw(table, tB, twB0, mapB);
        R::P1::load_tw(table, tB, twB1, mapB);
        float2 *buf0 = smem, *buf1 = smem + BUF;
        const float2 *rtwb = smem + 2 * BUF;
        R::load_rtw(io, kb, tid, smem + 2 * BUF);
        Epi ep;
        {
            const int s2 = tid % R::TPC;
            ep.c0 = tid / R::TPC;
            ep.kk = R::kmap(kb, s2);
            ep.own = R::slot_off(s2) + ep.c0;
            ep.mir = EPI == ROW_TWIST ? R::slot_off(R::mirror_slot(kb, s2)) - ep.c0 : 0;
        }
        __syncthreads();

        float2 *bufA = buf0 + slotA * C::RS;
        float2 *bfa = buf1 + R::slot_off(slotB), *bfb = buf0 + R::slot_off(slotB);
        const long j0 = (long)kb * C::COLS;
        const long cnt = team < rows ? (rows - team + teams - 1) / teams : 0; // transforms of this team
        float2 *slots = scratch + team * SLOTS * n;
        auto row_of = [&](long i) { return team + i * teams; };
        auto slot_of_i = [&](long i) { return slots + (i % SLOTS) * n; };
        // the polling thread's view of the two counters: loaded ahead of time, re-read only while short
        unsigned seenA = 0, seenB = 0;
        auto peek = [&](unsigned *c) -> unsigned { return tid == 0 ? flag_load(c) : 0u; };
        auto await = [&](unsigned *c, unsigned &seen, unsigned target) {
            if (tid == 0) {
                while (seen < target) seen = flag_load(c);
                flag_acquire();
            }
        };
        if (cnt == 0) return;

        // prologue: pass A of the team's first two transforms
        prefetch_a(io, row_of(0), j0, buf1, tid, pol);
        for (long a0 = 0; a0 < AHEAD && a0 < cnt; a0++) {
            tile_a(io, tw0, twA + tA * 16, row_of(a0), j0, slot_of_i(a0), bufA, buf1, tA, slotA, pol, [] {});
            __syncthreads();
            if (tid == 0) flag_arrive(doneA(a0));
            if (a0 + 1 < cnt) prefetch_a(io, row_of(a0 + 1), j0, buf1, tid, pol); // buf1 is idle until the next tile
        }
        float2 xb[EPT];
        bool have_xb = false;
        for (long i = 0; i < cnt; i++) {
            const bool nextA = i + AHEAD < cnt;
            if (nextA) {
                // pass A of transform i+AHEAD into the slot transform i+AHEAD-SLOTS occupied: every member
                // must have consumed that one; the flag load is issued before the tile's first pass
                const long victim = i + AHEAD - SLOTS;
                if (victim >= 0) seenB = peek(doneB(victim));
                tile_a(io, tw0, twA + tA * 16, row_of(i + AHEAD), j0, slot_of_i(i + AHEAD), bufA, buf1, tA, slotA, pol,
                       [&] { if (victim >= 0) await(doneB(victim), seenB, goal(victim)); });
            }
            // pass B of transform i
            if (!have_xb) { // first tile: its inputs could not be requested behind an epilogue
                if (nextA) __syncthreads(); // pass A's reads of buf0 / stores precede the arrival below
                if (nextA && tid == 0) flag_arrive(doneA(i + AHEAD));
                seenA = peek(doneA(i));
                await(doneA(i), seenA, goal(i));
                __syncthreads();
                load_b(xb, slot_of_i(i) + (long)k * R::NB, tB, pol);
            }
            const bool arrive_in_b = have_xb && nextA; // pass A's arrival rides on pass B's first barrier
            const bool nextB = i + 1 < cnt;
            if (nextB) seenA = peek(doneA(i + 1));
            tile_b(io, xb, twB0, twB1, row_of(i), kb, nextB ? slot_of_i(i + 1) + (long)k * R::NB : (const float2 *)nullptr,
                   bfa, bfb, buf0, tB, tid, ep, rtwb, pol,
                   [&] { if (arrive_in_b && tid == 0) flag_arrive(doneA(i + AHEAD)); },
                   [&] { if (nextB) await(doneA(i + 1), seenA, goal(i + 1)); },
                   [&] {
                       if (tid == 0) flag_arrive_relaxed(doneB(i)); // this CTA's reads of slot i are in registers
                       if (i + AHEAD + 1 < cnt) prefetch_a(io, row_of(i + AHEAD + 1), j0, buf1, tid, pol);
                   });
            have_xb = nextB;
            __syncthreads(); // both buffers are rewritten by the next tile
        }
    }
};

#ifdef __CUDACC__
template <bool EXACT, class IO, bool STAGED>
__global__ void __launch_bounds__(256, 2)
    colpass_kernel(const __grid_constant__ IO io, const __grid_constant__ Tw0 tw0, const float2 *__restrict__ table,
                   int lsub, long tiles, long row0, float2 *__restrict__ scratch)
{
    extern __shared__ __align__(128) float2 smem[];
    ColPass<EXACT, IO>::template run<STAGED>(io, tw0, table, lsub, tiles, row0, scratch, smem);
}

template <int LB, bool EXACT, class IO, int EPI, bool STAGED>
__global__ void __launch_bounds__(256, 2)
    rowpass_kernel(const __grid_constant__ IO io, const float2 *__restrict__ table, long tiles, long row0,
                   const float2 *__restrict__ scratch)
{
    extern __shared__ __align__(128) float2 smem[];
    RowPass<LB, EXACT, IO, EPI>::template run<STAGED>(io, table, tiles, row0, scratch, smem);
}
template <int LB, bool EXACT, class IO, int EPI>
__global__ void __launch_bounds__(256, 2)
    large_fused_kernel(const __grid_constant__ IO io, const __grid_constant__ Tw0 tw0, const float2 *__restrict__ table,
                       long rows, float2 *__restrict__ scratch)
{
    extern __shared__ __align__(128) float2 smem[];
    using F = LargeFused<LB, EXACT, IO, EPI>;
    const int rank = (int)cluster_ctarank();
    F::run(io, tw0, table, rows, scratch, smem, rank, blockIdx.x / F::CLUSTER, gridDim.x / F::CLUSTER);
}
template <int LB, bool EXACT, class IO, int EPI, bool STAGED>
__global__ void __launch_bounds__(256, 2)
    large_pipe_kernel(const __grid_constant__ IO io, const __grid_constant__ Tw0 tw0, const float2 *__restrict__ table,
                      long rows, float2 *__restrict__ scratch, unsigned *flags)
{
    extern __shared__ __align__(128) float2 smem[];
    LargePipe<LB, EXACT, IO, EPI, STAGED>::run(io, tw0, table, rows, scratch, smem, flags);
}
#endif

} // namespace kofft
