// emu_engine.cpp -- CPU emulation of the single-CTA FFT kernel (test scaffolding).
//
// Compiles kofft_b200/csrc/fft_kernels.cuh + small_kernels.cuh with plain g++ and replays
// CtaFft::run() phase by phase over all threads of a CTA (a __syncthreads() boundary becomes
// "finish the phase for every thread").  It exists because the build container has no GPU:
// it lets pytest check the engine's index math, twiddle selection, padding and every I/O
// policy bit-for-bit against the oracle before GPU time is spent.  It is never shipped, never
// imported by kofft_b200/, and is not a CPU fallback: only tests/ builds and loads it.
//
// Build: g++ -O1 -ffp-contract=off -std=c++17 -shared -fPIC (see tests/emu/build.py)
#include <cstring>
#include <vector>

#include "../../kofft_b200/csrc/small_kernels.cuh"

using namespace kofft;

namespace {

struct EmuArgs {
    const void *in, *in2;
    void *out, *out2;
    const void *aux;
    long n, p0, p1, p2, p3;
    float scale;
};

template <class P, bool EXACT, class IO>
void emu_cta(const IO &io_in, const Tw0 &tw0, const float2 *table, long rows, bool staged)
{
    using C = CtaFft<P, EXACT, IO>;
    using P0 = typename C::P0;
    using P1 = typename C::P1;
    using P2 = typename C::P2;
    using P3 = typename C::P3;
    using PL = Pass<P, P::NP - 1, EXACT>;
    const int NT = P::CTA;
    std::vector<float2> smem(2 * P::TPC * P::PADN);
    std::vector<float2> X(static_cast<size_t>(NT) * EPT);
    std::vector<float2> TW1(static_cast<size_t>(NT) * 16), TW2(static_cast<size_t>(NT) * 16), TW3(static_cast<size_t>(NT) * 16);
    IO io = io_in;
    IO io_prev[3] = {io_in, io_in, io_in};
    auto slot_of = [](int tid) { return tid / P::T; };
    auto t_of = [](int tid) { return tid % P::T; };
    auto buf_of = [&](int tid, int par) { return smem.data() + slot_of(tid) * P::PADN + (par ? P::TPC * P::PADN : 0); };
    for (int tid = 0; tid < NT; tid++) {
        P1::load_tw(table, t_of(tid), &TW1[tid * 16]);
        if (P::NP > 2) P2::load_tw(table, t_of(tid), &TW2[tid * 16]);
        if (P::NP > 3) P3::load_tw(table, t_of(tid), &TW3[tid * 16]);
    }
    int par = 0;
    const long groups = (rows + P::TPC - 1) / P::TPC;
    for (long g = 0; g < groups; g++) {
        auto row_of = [&](int tid) { return g * P::TPC + slot_of(tid); };
        int b = par; par ^= 1;
        std::vector<unsigned char> stage(P::STAGE_BYTES, 0xCD);
        if constexpr (IO::kStageable) {
            if (staged) {
                // grid of 3 CTAs: exercises group_next (adv_c / adv_f) as well as group_init
                if (g < 3) io.group_init(g, 3, P::TPC);
                else { io = io_prev[g % 3]; io.group_next(P::TPC); }
                io_prev[g % 3] = io;
                unsigned nb = io.stage_bytes(P::TPC, rows);
                if (nb > (unsigned)P::STAGE_BYTES || (nb % 16) != 0 || ((size_t)io.stage_src(P::TPC) % 16) != 0)
                    throw 1; // what the TMA bulk copy would reject
                if (nb) memcpy(stage.data(), io.stage_src(P::TPC), nb);
            }
        }
        for (int tid = 0; tid < NT; tid++) {
            float2 *x = &X[tid * EPT];
            bool done = false;
            if constexpr (IO::kStageable) {
                if (staged) {
                    int t = t_of(tid);
                    const int rctx = io.row_begin(slot_of(tid));
                    for (int u = 0; u < P0::U; u++)
                        for (int q = 0; q < P0::R; q++) {
                            int idx = P0::src_index(t, u, q);
                            const float w = IO::kLoadAux ? io.load_aux(idx) : 0.0f;
                            x[u * P0::R + q] = io.row_full(rctx)
                                                   ? io.template load_staged<true>(stage.data(), rctx, slot_of(tid), idx, w)
                                                   : io.template load_staged<false>(stage.data(), rctx, slot_of(tid), idx, w);
                        }
                    done = true;
                }
            }
            if (!done) {
                if (row_of(tid) < rows) C::template load_global<P0>(io, row_of(tid), t_of(tid), x);
                else for (int e = 0; e < EPT; e++) x[e] = make_float2(0.f, 0.f);
            }
            P0::compute(x, tw0.v);
            C::template store_smem<P0>(buf_of(tid, b), t_of(tid), x);
        }
        // -- sync --
        int b2 = par; if (P::NP > 2) par ^= 1;
        for (int tid = 0; tid < NT; tid++) {
            float2 *x = &X[tid * EPT];
            C::template load_smem<P1>(buf_of(tid, b), t_of(tid), x);
            P1::compute(x, &TW1[tid * 16]);
            if (P::NP > 2) C::template store_smem<P1>(buf_of(tid, b2), t_of(tid), x);
        }
        int b3 = par; if (P::NP > 3) par ^= 1;
        if (P::NP > 2) {
            for (int tid = 0; tid < NT; tid++) {
                float2 *x = &X[tid * EPT];
                C::template load_smem<P2>(buf_of(tid, b2), t_of(tid), x);
                P2::compute(x, &TW2[tid * 16]);
                if (P::NP > 3) C::template store_smem<P2>(buf_of(tid, b3), t_of(tid), x);
            }
        }
        if (P::NP > 3) {
            for (int tid = 0; tid < NT; tid++) {
                float2 *x = &X[tid * EPT];
                C::template load_smem<P3>(buf_of(tid, b3), t_of(tid), x);
                P3::compute(x, &TW3[tid * 16]);
            }
        }
        if constexpr (IO::kEpilogueExchange) {
            int be = par; par ^= 1;
            for (int tid = 0; tid < NT; tid++) C::template store_smem<PL>(buf_of(tid, be), t_of(tid), &X[tid * EPT]);
            for (int tid = 0; tid < NT; tid++)
                if (row_of(tid) < rows) C::epilogue(io, row_of(tid), t_of(tid), buf_of(tid, be));
        } else {
            for (int tid = 0; tid < NT; tid++)
                if (row_of(tid) < rows) C::template store_global<PL>(io, row_of(tid), t_of(tid), &X[tid * EPT]);
        }
    }
}

template <int N, bool EXACT, class IO>
void emu_small(const IO &io, long rows)
{
    for (long r = 0; r < rows; r++) small_transform<N, EXACT, IO>(io, r);
}

static bool g_staged = false;

template <bool EXACT, class IO>
int run_sized(int n, const IO &io, const Tw0 &tw0, const float2 *table, long rows)
{
    switch (n) {
    case 1: emu_small<1, EXACT>(io, rows); return 0;
    case 2: emu_small<2, EXACT>(io, rows); return 0;
    case 4: emu_small<4, EXACT>(io, rows); return 0;
    case 8: emu_small<8, EXACT>(io, rows); return 0;
    case 16: emu_small<16, EXACT>(io, rows); return 0;
#define CASE_L(L) case (1 << L): emu_cta<Plan<L>, EXACT>(io, tw0, table, rows, g_staged && Plan<L>::CAN_STAGE); return 0;
    CASE_L(5) CASE_L(6) CASE_L(7) CASE_L(8) CASE_L(9) CASE_L(10) CASE_L(11) CASE_L(12) CASE_L(13) CASE_L(14)
#undef CASE_L
    default: return -1;
    }
}

template <bool EXACT>
int run_kind(int kind, int n, const EmuArgs &q, const Tw0 &tw0, const float2 *table, long rows)
{
    switch (kind) {
    case 0: { IoC2C<false> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale}; return run_sized<EXACT>(n, io, tw0, table, rows); }
    case 1: { IoC2C<true> io{(const float2 *)q.in, (float2 *)q.out, q.n, q.scale}; return run_sized<EXACT>(n, io, tw0, table, rows); }
    case 2: { IoGeneric<false> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2, q.p0, q.p1, q.p2, q.p3, q.scale}; return run_sized<EXACT>(n, io, tw0, table, rows); }
    case 3: { IoGeneric<true> io{(const float *)q.in, (const float *)q.in2, (float *)q.out, (float *)q.out2, q.p0, q.p1, q.p2, q.p3, q.scale}; return run_sized<EXACT>(n, io, tw0, table, rows); }
    case 4: { IoStft io{(const float *)q.in, (const float *)q.aux, (float2 *)q.out, q.p0, q.p1, q.p2, q.n}; return run_sized<EXACT>(n, io, tw0, table, rows); }
    case 5: { IoIstft io{(const float2 *)q.in, (const float *)q.aux, (float *)q.out, q.n, q.scale}; return run_sized<EXACT>(n, io, tw0, table, rows); }
    case 6: { IoRfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n}; return run_sized<EXACT>(n, io, tw0, table, rows); }
    case 7: { IoIrfft<EXACT> io{(const float2 *)q.in, (float2 *)q.out, (const float2 *)q.aux, q.n, q.scale}; return run_sized<EXACT>(n, io, tw0, table, rows); }
    default: return -2;
    }
}

} // namespace

// kind: kofft::Kind numbering (launch.h).  table: the n/2-entry FftPlanner table (host).
extern "C" __attribute__((visibility("default"))) void kofft_emu_set_staged(int staged) { g_staged = staged != 0; }

extern "C" __attribute__((visibility("default"))) int
kofft_emu_run(int kind, int exact, long n, long rows, const void *in, const void *in2, void *out, void *out2,
              const void *aux, long p0, long p1, long p2, long p3, float scale, const float *table)
{
    EmuArgs q{in, in2, out, out2, aux, n, p0, p1, p2, p3, scale};
    Tw0 tw0;
    memset(&tw0, 0, sizeof tw0);
    if (n > 16) {
        int L = 0;
        while ((1L << L) < n) L++;
        const int NP = L <= 8 ? 2 : (L <= 12 ? 3 : 4);
        const int R0 = L - 4 * (NP - 1);
        for (int tl = 0; tl < R0; tl++)
            for (int c = 0; c < (1 << tl); c++) {
                long idx = (long)c << (L - 1 - tl);
                tw0.v[(1 << tl) - 1 + c] = make_float2(table[2 * idx], table[2 * idx + 1]);
            }
    }
    try {
        return exact ? run_kind<true>(kind, (int)n, q, tw0, (const float2 *)table, rows)
                     : run_kind<false>(kind, (int)n, q, tw0, (const float2 *)table, rows);
    } catch (int) {
        return -7; // staged copy would violate TMA alignment / size rules
    }
}

// shared-memory conflict audit: for every (L, exchange, access) return the worst number of
// distinct 8-byte bank pairs hit by more than one lane of a half-warp (1 = conflict-free).
template <class P, class PS, bool SRC>
int worst_conflict()
{
    int worst = 1;
    for (int hw = 0; hw < P::CTA / 16; hw++)
        for (int u = 0; u < PS::U; u++)
            for (int q = 0; q < PS::R; q++) {
                int cnt[16] = {0};
                for (int l = 0; l < 16; l++) {
                    int tid = hw * 16 + l, slot = tid / P::T, t = tid % P::T;
                    int a = slot * P::PADN + (SRC ? PS::src_pad(PS::src_base(t, u), q) : PS::dst_pad(PS::dst_base(t, u), q));
                    cnt[a & 15]++;
                }
                for (int b = 0; b < 16; b++) worst = cnt[b] > worst ? cnt[b] : worst;
            }
    return worst;
}

template <int L>
void audit_L(int *out) // out[0..5]: st0, ld1, st1, ld2, st2, ld3 (0 where absent)
{
    using P = Plan<L>;
    using P0 = Pass<P, 0, true>;
    using P1 = Pass<P, 1, true>;
    using P2 = Pass<P, (P::NP > 2 ? 2 : 1), true>;
    using P3 = Pass<P, (P::NP > 3 ? 3 : 1), true>;
    for (int i = 0; i < 6; i++) out[i] = 0;
    out[0] = worst_conflict<P, P0, false>();
    out[1] = worst_conflict<P, P1, true>();
    if (P::NP > 2) { out[2] = worst_conflict<P, P1, false>(); out[3] = worst_conflict<P, P2, true>(); }
    if (P::NP > 3) { out[4] = worst_conflict<P, P2, false>(); out[5] = worst_conflict<P, P3, true>(); }
}

extern "C" __attribute__((visibility("default"))) int kofft_emu_bank_audit(int L, int *out)
{
    switch (L) {
#define CASE_L(L) case L: audit_L<L>(out); return 0;
    CASE_L(5) CASE_L(6) CASE_L(7) CASE_L(8) CASE_L(9) CASE_L(10) CASE_L(11) CASE_L(12) CASE_L(13) CASE_L(14)
#undef CASE_L
    default: return -1;
    }
}
