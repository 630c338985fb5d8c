// tcgen05 / TMEM / TMA implicit-GEMM Conv1d (k=5, pad 2, stride 1; stride-2 gather / scatter; plus an optional
// 1x1 slab) with the fused U-Net epilogues, in error-compensated BF16x3 arithmetic.
//
// Replaces the cuDNN conv + native GroupNorm + SiLU + embedding/residual adds of
//   Conv1dBlock_dd / ResidualTemporalBlock_dd   diffuser/models/helpers.py:88-94, temporal_dd.py:71-74
//   Downsample1d / Upsample1d                   diffuser/models/helpers.py:38-52
// and their backward-data counterparts inside torch.autograd.grad (temporal_cond.py:321).
//
// GEMM view: M = (samples x horizon) rows, N = output channels, K = taps x input channels.
//   * one tile = S whole samples (S*H <= 128 rows) x NT channels, so GroupNorm slabs never straddle tiles and the
//     conv halo is produced by TMA out-of-bounds zero fill: the A tile of tap k is ONE 3-D box (K-block channels,
//     H rows, S samples) of the channels-last activation plane started at row k-2;
//   * operands are BF16 pairs (v ~= hi + lo, |error| ~ 2^-17 |v|); each K block issues hi*hi + hi*lo + lo*hi into
//     the same FP32 TMEM accumulator -> ~2^-16 relative product error, which keeps eps within the 1e-3 parity
//     budget where single-pass BF16/TF32 does not; algorithmic FLOPs are 1/3 of the issued;
//   * persistent CTAs, warp-specialised: warp 0 TMA producer, warp 1 MMA issuer (+TMEM owner), warps 2..9 epilogue;
//     shared-memory ring of NSTAGE stages, two TMEM accumulator stages so the epilogue of tile i overlaps the main
//     loop of tile i+1;
//   * template <NT, GS, ACC2, CG, TE, NG, TEAM, AR>: NT tile columns (64/128/192/256), GS GroupNorm group size (0: none), ACC2 a
//     separate accumulator for the forward residual 1x1 conv, CG = 2 CTA pairs (tcgen05 cta_group::2: M = 256 over
//     the two SMs of a TPC, each CTA stages its own A rows and half of the B rows; 256-column tiles use 32-channel
//     K blocks / SWIZZLE_64B), TE the TMA-store epilogue (results staged in swizzled shared-memory tiles and written
//     by cp.async.bulk.tensor stores; its inputs -- saved yhat, identity planes -- arrive by TMA loads); NG = 4: 16 epilogue
//     warps on 16-column chunks (path 4);
//     TEAM: the two epilogue warp groups take alternate tiles (64-column forward tiles); AR: A-operand reuse -- the
//     haloed, position-major activation tile of a channel block is loaded once and the taps are descriptor offsets
//     (stride-1 convs on the 256-column pair tiles).  DESIGN.md 3.1 has the measurements behind these choices.
// Diagnostics (none of them on by default): PMP_TC_TIMERS=1|2 (+PMP_TC_TIMER_MATCH=N:C1:g) per-role / per-phase cycle
// counters read through pmp_debug_counters; PMP_TC_DEBUG bit mask -- 1 no epilogue, 2 no direct stores, 4 no
// identity/yhat loads, 8 no embedding, 16 MMA only (no TMA fill), 32 N=256 issue-rate probe, 64/128 A / B loaded for one
// tap of five, 512 no TMA stores, 2048 all stores to the first rows: timing probes whose results are garbage; PMP_TC_WIDE=0 keeps 128-column tiles, PMP_TC_AR=0 / PMP_TC_TEAM=0 / PMP_TC_NG4_BWD=0 switch the respective variants off.
#include <cuda.h>

#include <map>
#include <tuple>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pmp {
namespace {
using namespace tc;

// Diagnostics (role / phase timers, work-skipping timing probes) exist only in -DPMP_DIAG builds
// (`make diag` -> libpmp_b200_diag.so); the release library compiles them out.
#ifdef PMP_DIAG
#define DBGF(bit) ((P.dbg & (bit)) != 0)
#define CTRP (P.ctr)
#else
#define DBGF(bit) (false)
#define CTRP (static_cast<unsigned long long*>(nullptr))
#endif

constexpr int TC_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int BK = 64;                      // channels per K block = one 128-byte swizzle row
constexpr int A_TILE_BYTES = 128 * BK * 2;  // 16 KB per BF16 plane
constexpr int MAXSEG = 5;                   // samples touched by 32 consecutive rows (H >= 8)

}  // namespace
unsigned long long* tc_timer_buffer();
namespace {

struct TcP {
    ConvP e;
    int tilesM, tilesN, ntiles;
    int npairs;     // CTA-pair schedule: ceil(tilesM / 2) * tilesN * nphase
    int kpt;        // K blocks per tap = C1 / 64
    int nkb2;       // K blocks of slab 2 (slab 1: ntaps[phase] * kpt)
    int nphase;     // 2 for transposed (stride-2 scatter) convs: output rows 2j+phase
    int mode4d;     // 1: stride-2 gather, A maps are 4-D (C, parity, H/2, R)
    unsigned long long* ctr;   // optional role timers (PMP_TC_TIMERS=1): see pmp_debug_counters
    int tmode;      // PMP_TC_TIMERS: 1 = role timers, 2 = epilogue phase laps (see pmp_debug_counters)
    int dbg;        // diagnostics (PMP_TC_DEBUG): bit 0 = epilogue releases the accumulator without reading it
    alignas(64) CUtensorMap mA1h;
    alignas(64) CUtensorMap mA1l;
    alignas(64) CUtensorMap mA2h;
    alignas(64) CUtensorMap mA2l;
    alignas(64) CUtensorMap mW1h;
    alignas(64) CUtensorMap mW1l;
    alignas(64) CUtensorMap mW2h;
    alignas(64) CUtensorMap mW2l;
    // TMA-store epilogue (TE kernels): fp32 maps of yhat / out / gy and BF16 maps of the operand planes,
    // box = (32 channels, HJ rows, S samples) of the tile (4-D with the output phase for stride-2 scatter)
    alignas(64) CUtensorMap mSy;
    alignas(64) CUtensorMap mSo;
    alignas(64) CUtensorMap mSoh;
    alignas(64) CUtensorMap mSol;
    alignas(64) CUtensorMap mSg;
    alignas(64) CUtensorMap mSgh;
    alignas(64) CUtensorMap mSgl;
    // TMA-LOAD side of the epilogue: the identity / residual term given as BF16 planes (box geometry of the store
    // maps); the saved yhat of a GroupNorm-backward launch is loaded through mSy
    alignas(64) CUtensorMap mIh;
    alignas(64) CUtensorMap mIl;
};

// small-K conv on the CUDA cores for the row owned by this thread (weights staged in smem [t][d][NT])
template <int NT, int CW>
__device__ __forceinline__ void inline_conv(const ConvP& p, const float* sinl, int ch, int rg, int h, int Hl,
                                            bool valid, float (&v)[CW]) {
    if (!valid) return;
    const int half = p.inl_taps >> 1;
    for (int t = 0; t < p.inl_taps; ++t) {
        const int hh = h + t - half;
        if (hh < 0 || hh >= Hl) continue;
        int xrow = rg + p.inl_row0;
        if (p.inl_wrap && xrow >= p.inl_wrap) xrow -= p.inl_wrap;
        const float* xr = p.inl_x + ((size_t)xrow * Hl + hh) * p.inl_ld;
        for (int d = 0; d < p.inl_D; ++d) {
            const float xd = __ldg(xr + d);
            const float4* w = reinterpret_cast<const float4*>(sinl + (t * p.inl_D + d) * NT + ch * CW);
#pragma unroll
            for (int k = 0; k < CW / 4; ++k) {
                const float4 w4 = w[k];
                v[4 * k] = fmaf(xd, w4.x, v[4 * k]);
                v[4 * k + 1] = fmaf(xd, w4.y, v[4 * k + 1]);
                v[4 * k + 2] = fmaf(xd, w4.z, v[4 * k + 2]);
                v[4 * k + 3] = fmaf(xd, w4.w, v[4 * k + 3]);
            }
        }
    }
}

template <int NT, int GS, bool ACC2, int CG, bool TE, int NG = 2, bool TEAM = false, bool AR = false>
struct Cfg {
    // AR (A-operand reuse, stride-1 convs on the 256-column pair tiles): the haloed activation tile of a 32-channel
    // block is loaded ONCE, position-major (row = h*S + s), and the taps are descriptor start offsets of S rows;
    // only the weight tiles stream per tap.  Cuts the L2->SM operand traffic of a 5-tap conv by 40 %.
    static_assert(!AR || ((NT == 256 && CG == 2 && !TEAM) || NT == 64), "A reuse: 256-column pair tiles, 64-column tiles");
    static_assert(!AR || !ACC2, "A reuse: one accumulator");
    static constexpr int NA = 2;                        // haloed A buffers
    static constexpr int A_BUF_ROWS = 192;              // (HJ + 4) * S <= 128 + 4 * 16
    static constexpr int A_BUF_BYTES = A_BUF_ROWS * 32 * 2;   // per plane (32-channel K blocks)
    // TEAM: the two epilogue warp groups work on ALTERNATE TILES (each walks all chunks of its own tile, with its own
    // barrier and scratch tables) instead of on alternate chunks of one tile: two per-tile latency chains (TMEM load ->
    // stats -> barrier -> reduce -> barrier -> output pass) overlap.  For the narrow, epilogue-bound 64-column tiles.
    static_assert(!TEAM || (NG == 2 && !ACC2 && NT == 64), "team mode: 64-column tiles, one accumulator");
    // NG epilogue warp groups of 4 warps (one per TMEM lane quarter): 2 groups walk 32-column chunks, 4 groups (16
    // epilogue warps, path 4) walk 16-column chunks with half the live registers per thread
    static constexpr int CW = NG == 4 ? 16 : 32;
    static constexpr int THREADS = 64 + NG * 128;
    static constexpr int ETHREADS = NG * 128;
    // CG = 2: CTA pair (cta_group::2); each CTA stages its own 128 A rows and NT/2 of the B rows
    // TE: TMA-store epilogue (64 KB of staging tiles: per half of the epilogue warps one fp32 and two BF16 tiles)
    // channels per K block: 64 (one 128-byte swizzle row) except for the 256-column tiles, which take 32-channel
    // blocks (64-byte rows, SWIZZLE_64B) so that four stages of the same total size give the TMA more lookahead
    static constexpr int BKC = (NT == 256 || AR) ? 32 : 64;
    static constexpr int A_BYTES = 128 * BKC * 2;
    static constexpr int B_TILE_BYTES = (NT / CG) * BKC * 2;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_TILE_BYTES;
    // PIPE (TE kernels on the 256-column tiles, and the 16-warp 64-column ones): the epilogue's staging is split into
    // OUTPUT tiles (F | H | L) and INPUT tiles for the identity planes (Ih | Il; yhat of a GroupNorm-backward launch
    // comes in through F, which is not an output there), so that the TMA load of the NEXT chunk's inputs and the drain
    // of the PREVIOUS chunk's bulk stores overlap the arithmetic of the current one (measured before: 0.9-1.9 k cycles
    // of exposed drain per chunk and 1.1-1.7 k per load, 30-45 % of the epilogue).  Paid for with one ring stage.
    static constexpr bool PIPE = TE && !TEAM && (NT == 256 || (NT == 64 && NG == 4));
    static constexpr int NSTAGE =
        AR ? (PIPE ? 4 : 5)
           : (CG == 2 ? (NT == 256 ? (PIPE ? 3 : 4) : (NT == 128 ? (TE ? 3 : 4) : 3)) : (NT == 64 ? (TE ? 2 : 3) : (NT == 128 ? 3 : 2)));
    // AR: the ring holds the weight tiles only (NSTAGE stages of 2 planes), the A buffers sit in front of it
    static constexpr int RING_BYTES = AR ? (NA * 2 * A_BUF_BYTES + NSTAGE * 2 * B_TILE_BYTES) : NSTAGE * STAGE_BYTES;
    static constexpr int STG_HALF_BYTES = 128 * CW * 4 + 2 * 128 * CW * 2 + (PIPE ? 2 * 128 * CW * 2 : 0);   // F | H | L (| Ih | Il) of one warp group
    static constexpr int STG_BYTES = TE ? NG * STG_HALF_BYTES : 0;
    static constexpr int INL_FLOATS = (NT == 64 && CG == 1) ? 5 * 16 * 64 : 0;   // inline small-K conv weights [taps][D<=16][64]
    static constexpr int ACC_COLS = NT * (ACC2 ? 2 : 1);
    // two accumulator stages when they fit next to the yhat stash of the GroupNorm-backward epilogue
    static constexpr int NACC = TEAM ? 4 : ((NT != 192 && 2 * ACC_COLS <= 512) ? 2 : 1);
    // the GroupNorm-backward epilogue parks yhat in spare TMEM columns between its two passes when there are
    // any (NT <= 128); the 256-column tiles fill TMEM with their two accumulator stages and re-read yhat (L2)
    static constexpr bool STASH = NACC * ACC_COLS + NACC * NT <= 512;
    static constexpr int G = GS ? NT / GS : 1;
    static constexpr int NCH = NT / CW;
    // chunks per GroupNorm group: consecutive chunks go to consecutive warp groups, so a group's row partials have
    // CPG distinct writers (needs CPG <= NG)
    static constexpr int CPG = (GS > CW) ? GS / CW : 1;
    static_assert(!GS || CPG <= NG || NG == 2, "a GroupNorm group spans more chunks than there are warp groups");
    // smem: stages | staging tiles | row/slab partials | slab stats | column params | row info | barriers
    static constexpr int TBUF_FLOATS = 0;
    // NG = 4: every warp group takes a CONTIGUOUS range of NCH / 4 chunks (64 columns >= one GroupNorm group), so a
    // (row, group) partial has one writer; NG = 2: the two groups alternate chunks (two writers)
    static constexpr bool CONTIG = NG == 4 && !TEAM;
    static_assert(!CONTIG || !GS || (NT / NG) % GS == 0 || GS % (NT / NG) == 0, "chunk ranges vs GroupNorm groups");
    static constexpr int PSLOTS = TEAM ? 1 : (NG == 2 ? (CPG >= 2 ? 2 : 1) : ((GS > NT / NG) ? GS / (NT / NG) : 1));
    static constexpr int PART_FLOATS = PSLOTS * 128 * G * 2;
    static constexpr int NTEAM = TEAM ? 2 : 1;     // copies of the per-tile scratch tables
    static constexpr int STAT_FLOATS = 384;
    static constexpr int COL_FLOATS = 4 * NT;
    static constexpr int ROW_INTS = 4 * 128;
    static constexpr int SMEM_BYTES = RING_BYTES + STG_BYTES +
                                      (TBUF_FLOATS + NTEAM * (PART_FLOATS + STAT_FLOATS + COL_FLOATS + INL_FLOATS) + ROW_INTS) * 4 +
                                      256 + 1024;
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

template <int NT, int GS, bool ACC2, int CG, bool TE, int NG, bool TEAM, bool AR>
__global__ void __launch_bounds__(64 + NG * 128, 1) conv_tc_kernel(const __grid_constant__ TcP P) {
    using C = Cfg<NT, GS, ACC2, CG, TE, NG, TEAM, AR>;
    constexpr int CW = C::CW, ETH = C::ETHREADS;
    constexpr int G = C::G, NCH = C::NCH, NSTAGE = C::NSTAGE, NACC = C::NACC;
    const ConvP& p = P.e;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer
    uint8_t* stg = smem + C::RING_BYTES;   // TE staging tiles (1024-byte aligned)
    float* tbuf = reinterpret_cast<float*>(stg + C::STG_BYTES);
    // per-tile scratch tables (one set per epilogue team): row/slab partials | slab stats | column params | inline weights
    const int team_of_warp = (TEAM && (threadIdx.x >> 5) >= 6) ? 1 : 0;
    constexpr int SCR_FLOATS = C::PART_FLOATS + C::STAT_FLOATS + C::COL_FLOATS + C::INL_FLOATS;
    float* spart = tbuf + C::TBUF_FLOATS + team_of_warp * SCR_FLOATS;
    float* sstat = spart + C::PART_FLOATS;
    float* scol = sstat + C::STAT_FLOATS;                 // bias | gamma | beta | bias2, NT each
    float* sinl = scol + C::COL_FLOATS;
    int* srow = reinterpret_cast<int*>(tbuf + C::TBUF_FLOATS + C::NTEAM * SCR_FLOATS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(srow + C::ROW_INTS);
    uint64_t* full = bars;
    uint64_t* empty = bars + NSTAGE;
    uint64_t* tfull = bars + 2 * NSTAGE;
    uint64_t* tempty = tfull + 4;
    uint64_t* afull = tempty + 4;          // AR: haloed A buffers
    uint64_t* aempty = afull + 2;
    uint64_t* lfull = aempty + 2;          // TE: epilogue input tiles (yhat, identity planes) landed, one per warp group
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lfull + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.S, HJ = p.HJ, N = p.N;
    constexpr int BKC = C::BKC, A_TILE = C::A_BYTES;
    const uint32_t bytesA = (uint32_t)(S * HJ * BKC * 2);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&P.mA1h); prefetch_tmap(&P.mA1l); prefetch_tmap(&P.mW1h); prefetch_tmap(&P.mW1l);
        if (P.nkb2) { prefetch_tmap(&P.mA2h); prefetch_tmap(&P.mA2l); prefetch_tmap(&P.mW2h); prefetch_tmap(&P.mW2l); }
        if (TE) {
            if (p.yhat && p.gn_fwd) prefetch_tmap(&P.mSy);
            if (p.out) prefetch_tmap(&P.mSo);
            if (p.out_hi) { prefetch_tmap(&P.mSoh); prefetch_tmap(&P.mSol); }
            if (p.gy) prefetch_tmap(&P.mSg);
            if (p.gy_hi) { prefetch_tmap(&P.mSgh); prefetch_tmap(&P.mSgl); }
            if (p.yhat && p.gn_bwd) prefetch_tmap(&P.mSy);
            if (p.ident_hi) { prefetch_tmap(&P.mIh); prefetch_tmap(&P.mIl); }
        }
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        // tempty: the epilogues of both CTAs of a pair release the accumulator stage to the leader's MMA warp
        for (int i = 0; i < 4; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], CG); }
        for (int i = 0; i < 2; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
        for (int i = 0; i < 4; ++i) mbar_init(&lfull[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();   // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // tile schedule: CG = 1: tile -> (phase, n tile, m tile); CG = 2: a cluster takes a PAIR of adjacent m
    // tiles of the same (phase, n tile); rank r owns m tile 2*pair + r (a pair's odd tile may lie past the
    // last row: its loads are zero-filled and its epilogue stores nothing)
    const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;
    const int sched0 = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int sched_step = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int nsched = CG == 2 ? P.npairs : P.ntiles;

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (lane == 0) {
            auto ld2 = [](void* d, const CUtensorMap* m, uint64_t* b, int c0, int c1) {
                if (CG == 2) tma2_load_2d(d, m, b, c0, c1); else tma_load_2d(d, m, b, c0, c1);
            };
            auto ld3 = [](void* d, const CUtensorMap* m, uint64_t* b, int c0, int c1, int c2) {
                if (CG == 2) tma2_load_3d(d, m, b, c0, c1, c2); else tma_load_3d(d, m, b, c0, c1, c2);
            };
            auto ld4 = [](void* d, const CUtensorMap* m, uint64_t* b, int c0, int c1, int c2, int c3) {
                if (CG == 2) tma2_load_4d(d, m, b, c0, c1, c2, c3); else tma_load_4d(d, m, b, c0, c1, c2, c3);
            };
            int s = 0;
            uint32_t ph = 0;
            if constexpr (AR) {
                // A reuse: per 32-channel block ONE haloed, position-major activation tile (both planes), then one
                // weight tile pair per tap through the ring
                int ai = 0;
                uint32_t aph = 0;
                uint8_t* const bbase = smem + C::NA * 2 * C::A_BUF_BYTES;
                const uint32_t bytesAh = (uint32_t)((HJ + 4) * S * BKC * 2);
                for (int tile = sched0; tile < nsched; tile += sched_step) {
                    const int r0 = ((tile / P.tilesN) * CG + (int)crank) * S, n0 = (tile % P.tilesN) * NT;
                    const int nb0 = n0 + (int)crank * (NT / CG);
                    for (int cb = 0; cb < P.kpt + P.nkb2; ++cb) {
                        const bool sl2 = cb >= P.kpt;
                        const int c0 = (sl2 ? cb - P.kpt : cb) * BKC;
                        mbar_wait(&aempty[ai], aph ^ 1);
                        uint8_t* ab = smem + ai * 2 * C::A_BUF_BYTES;
                        if (crank == 0) mbar_expect_tx(&afull[ai], CG * 2 * bytesAh);
                        ld3(ab, sl2 ? &P.mA2h : &P.mA1h, &afull[ai], c0, r0, -2);
                        ld3(ab + C::A_BUF_BYTES, sl2 ? &P.mA2l : &P.mA1l, &afull[ai], c0, r0, -2);
                        const int nt = sl2 ? 1 : p.ntaps[0];
                        for (int tap = 0; tap < nt; ++tap) {
                            mbar_wait(&empty[s], ph ^ 1);
                            uint8_t* st = bbase + s * 2 * C::B_TILE_BYTES;
                            if (crank == 0) mbar_expect_tx(&full[s], CG * 2 * C::B_TILE_BYTES);
                            const int wr = sl2 ? nb0 : p.tap_k[0][tap] * N + nb0;
                            ld2(st, sl2 ? &P.mW2h : &P.mW1h, &full[s], c0, wr);
                            ld2(st + C::B_TILE_BYTES, sl2 ? &P.mW2l : &P.mW1l, &full[s], c0, wr);
                            if (++s == NSTAGE) { s = 0; ph ^= 1; }
                        }
                        if (++ai == C::NA) { ai = 0; aph ^= 1; }
                    }
                }
            } else
            for (int tile = sched0; tile < nsched; tile += sched_step) {
                const int op = tile % P.nphase, t2 = tile / P.nphase;
                const int r0 = ((t2 / P.tilesN) * CG + (int)crank) * S, n0 = (t2 % P.tilesN) * NT;
                const int nb0 = n0 + (int)crank * (NT / CG);   // first B row staged by this CTA
                const int nkb1 = p.ntaps[op] * P.kpt;
                for (int kb = 0; kb < nkb1 + P.nkb2; ++kb) {
                    if (DBGF(16)) break;   // diagnostics: MMA-only runs read whatever is in shared memory
                    const long long tw0 = (CTRP && P.tmode == 1) ? clock64() : 0;
                    mbar_wait(&empty[s], ph ^ 1);
                    if (CTRP && P.tmode == 1) atomicAdd(CTRP + 5, (unsigned long long)(clock64() - tw0));
                    uint8_t* st = smem + s * C::STAGE_BYTES;
                    // dbg 64 (timing probe for operand reuse, garbage results): the A tile is only loaded for one tap of five
                    const bool skipA = DBGF(64) && kb < nkb1 && p.ntaps[op] == 5 && (kb / P.kpt) != 2;
                    const bool skipB = DBGF(128) && kb < nkb1 && p.ntaps[op] == 5 && (kb / P.kpt) != 2;   // same probe for B
                    if (CG == 1 || crank == 0)
                        mbar_expect_tx(&full[s], CG * ((skipA ? 0 : 2 * bytesA) + (skipB ? 0 : 2 * C::B_TILE_BYTES)));
                    if (kb < nkb1) {
                        const int tap = kb / P.kpt, c0 = (kb - tap * P.kpt) * BKC;
                        const int sh = p.tap_shift[op][tap], wr = p.tap_k[op][tap] * N + nb0;
                        if (skipA) {
                        } else if (P.mode4d) {
                            // input row 2j + sh = parity (sh & 1) of pair j + floor(sh / 2)
                            const int par = sh & 1, js = (sh - par) >> 1;
                            ld4(st, &P.mA1h, &full[s], c0, par, js, r0);
                            ld4(st + A_TILE, &P.mA1l, &full[s], c0, par, js, r0);
                        } else {
                            ld3(st, &P.mA1h, &full[s], c0, sh, r0);
                            ld3(st + A_TILE, &P.mA1l, &full[s], c0, sh, r0);
                        }
                        if (!skipB) {
                            ld2(st + 2 * A_TILE, &P.mW1h, &full[s], c0, wr);
                            ld2(st + 2 * A_TILE + C::B_TILE_BYTES, &P.mW1l, &full[s], c0, wr);
                        }
                    } else {
                        const int c0 = (kb - nkb1) * BKC;
                        ld3(st, &P.mA2h, &full[s], c0, 0, r0);
                        ld3(st + A_TILE, &P.mA2l, &full[s], c0, 0, r0);
                        ld2(st + 2 * A_TILE, &P.mW2h, &full[s], c0, nb0);
                        ld2(st + 2 * A_TILE + C::B_TILE_BYTES, &P.mW2l, &full[s], c0, nb0);
                    }
                    if (++s == NSTAGE) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ==============================
        if (lane == 0 && crank == 0) {
            // instruction descriptor: D=F32, A=B=BF16, both K-major, N=NT, M=128 per CTA
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) |
                                   ((uint32_t)((DBGF(32) && NT == 128 && !ACC2 ? 256 : NT) >> 3) << 17) |   // dbg 32: N=256 rate probe
                                   (((128u * CG) >> 4) << 24);
            auto mma = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
                if (CG == 2) tc_mma2(d, a, b, id, acc); else tc_mma(d, a, b, id, acc);
            };
            auto commit = [](uint64_t* bar) {
                if (CG == 2) tc_commit2(bar); else tc_commit(bar);
            };
            int s = 0, as = 0;
            uint32_t ph = 0, aph = 0;
            int ai = 0;            // AR: haloed A buffer ring
            uint32_t aphA = 0;
            const long long tm0 = (CTRP && P.tmode == 1) ? clock64() : 0;
            for (int tile = sched0; tile < nsched; tile += sched_step) {
                const long long te0 = (CTRP && P.tmode == 1) ? clock64() : 0;
                mbar_wait(&tempty[as], aph ^ 1);
                if (CTRP && P.tmode == 1) atomicAdd(CTRP + 1, (unsigned long long)(clock64() - te0));
                tc_fence_after();
                const uint32_t d1 = tmem_base + (uint32_t)(as * C::ACC_COLS);
                const uint32_t d2 = ACC2 ? d1 + NT : d1;
                uint32_t acc1 = 0, acc2 = ACC2 ? 0u : 1u;
                if constexpr (AR) {
                    uint8_t* const bbase = smem + C::NA * 2 * C::A_BUF_BYTES;
                    uint32_t accum = 0;
                    for (int cb = 0; cb < P.kpt + P.nkb2; ++cb) {
                        const bool sl2 = cb >= P.kpt;
                        const long long tf0 = (CTRP && P.tmode == 1) ? clock64() : 0;
                        mbar_wait(&afull[ai], aphA);
                        if (CTRP && P.tmode == 1) atomicAdd(CTRP + 7, (unsigned long long)(clock64() - tf0));   // haloed A tile
                        tc_fence_after();
                        const uint32_t abuf = smem_u32(smem + ai * 2 * C::A_BUF_BYTES);
                        const int nt = sl2 ? 1 : p.ntaps[0];
                        for (int tap = 0; tap < nt; ++tap) {
                            const long long tf1 = (CTRP && P.tmode == 1) ? clock64() : 0;
                            mbar_wait(&full[s], ph);
                            if (CTRP && P.tmode == 1) atomicAdd(CTRP + 0, (unsigned long long)(clock64() - tf1));
                            tc_fence_after();
                            // tap = start offset of (shift + 2) * S rows inside the haloed tile
                            const uint32_t aoff = (uint32_t)(((sl2 ? 0 : p.tap_shift[0][tap]) + 2) * S * BKC * 2);
                            const uint32_t sb = smem_u32(bbase + s * 2 * C::B_TILE_BYTES);
                            const uint64_t ah = make_desc<BKC>(abuf + aoff), al = make_desc<BKC>(abuf + C::A_BUF_BYTES + aoff);
                            const uint64_t bh = make_desc<BKC>(sb), bl = make_desc<BKC>(sb + C::B_TILE_BYTES);
#pragma unroll
                            for (int k = 0; k < BKC / 16; ++k) {
                                const uint64_t ko = (uint64_t)(k * 2);
                                mma(d1, ah + ko, bh + ko, idesc, accum);
                                mma(d1, ah + ko, bl + ko, idesc, 1u);
                                mma(d1, al + ko, bh + ko, idesc, 1u);
                                accum = 1u;
                            }
                            commit(&empty[s]);
                            if (++s == NSTAGE) { s = 0; ph ^= 1; }
                        }
                        commit(&aempty[ai]);      // the haloed tile may be overwritten (in both CTAs)
                        if (++ai == C::NA) { ai = 0; aphA ^= 1; }
                    }
                } else {
                const int nkb1 = p.ntaps[tile % P.nphase] * P.kpt;
                for (int kb = 0; kb < nkb1 + P.nkb2; ++kb) {
                    const long long tf0 = (CTRP && P.tmode == 1) ? clock64() : 0;
                    if (!DBGF(16)) mbar_wait(&full[s], ph);
                    if (CTRP && P.tmode == 1) atomicAdd(CTRP + 0, (unsigned long long)(clock64() - tf0));
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * C::STAGE_BYTES);
                    const uint64_t ah = make_desc<BKC>(sa), al = make_desc<BKC>(sa + A_TILE);
                    const uint64_t bh = make_desc<BKC>(sa + 2 * A_TILE), bl = make_desc<BKC>(sa + 2 * A_TILE + C::B_TILE_BYTES);
                    const bool slab2 = kb >= nkb1;
                    const uint32_t d = DBGF(32) ? tmem_base : (slab2 ? d2 : d1);
                    uint32_t accum = slab2 ? (ACC2 ? acc2 : 1u) : acc1;
#pragma unroll
                    for (int k = 0; k < BKC / 16; ++k) {
                        const uint64_t ko = (uint64_t)(k * 2);   // 32 bytes per K=16 step, in 16-byte units
                        mma(d, ah + ko, bh + ko, idesc, accum);
                        mma(d, ah + ko, bl + ko, idesc, 1u);
                        mma(d, al + ko, bh + ko, idesc, 1u);
                        accum = 1u;
                    }
                    if (slab2) acc2 = 1u; else acc1 = 1u;
                    if (!DBGF(16)) commit(&empty[s]);   // frees the smem stage (of both CTAs) when these MMAs have read it
                    if (++s == NSTAGE) { s = 0; ph ^= 1; }
                }
                }
                commit(&tfull[as]);      // accumulator complete
                if (++as == NACC) { as = 0; aph ^= 1; }
            }
            if (CTRP && P.tmode == 1) atomicAdd(CTRP + 2, (unsigned long long)(clock64() - tm0));
        }
    } else {
        // ============================== epilogue (warps 2..9) ==============================
        // Every thread owns ONE accumulator row (TMEM lane) and walks 32-column chunks; two warps
        // share each TMEM lane quarter and take alternating chunks.  All global traffic is
        // per-row 16-byte vectors (no shared-memory transposes); GroupNorm statistics are per-row
        // partial sums combined in a fixed order, so results are run-to-run deterministic.
        const int eq = warp & 3;                 // TMEM lane quarter this warp may access
        const int hf = (warp - 2) >> 2;          // warp group 0..NG-1: which chunks
        const int etid0 = threadIdx.x - 64;      // 0..ETH-1
        // cooperative loops, barriers and "thread 0" duties are per team in team mode, per CTA otherwise
        const int etid = TEAM ? (etid0 & 127) : etid0;
        constexpr int ESTEP = TEAM ? 128 : ETH;
        auto ebar = [&]() {
            if (TEAM) asm volatile("bar.sync %0, 128;" ::"r"(6 + hf) : "memory");
            else epi_barw<ETH>();
        };
        // chunks of this thread: [ch0, chE) in steps of CHS
        const int ch0 = TEAM ? 0 : (C::CONTIG ? hf * (NCH / NG) : hf);
        constexpr int CHS = (TEAM || C::CONTIG) ? 1 : NG;
        const int chE = C::CONTIG ? ch0 + NCH / NG : NCH;
        constexpr bool PIPE = C::PIPE;
        const int mrow = eq * 32 + lane;
        const int act = p.act;
        const float inv_cnt = 1.0f / (float)(HJ * (GS ? GS : 1));
        const int RS = AR ? S : 1;               // row stride between consecutive positions of one sample
        int as = 0;
        uint32_t aph = 0;
        long long tlap = 0;
        auto lap = [&](int k) {
            if (CTRP && P.tmode == 2 && etid == 0) {
                const long long now = clock64();
                atomicAdd(CTRP + k, (unsigned long long)(now - tlap));
                tlap = now;
            }
        };
        // writer slot of chunk ch in the row-partial table (NG = 2: the warp group; NG = 4: position of the chunk in its group)
        auto pslot = [&](int ch) {
            (void)ch;
            return (TEAM || C::PSLOTS == 1) ? 0 : hf % C::PSLOTS;
        };
        // TE: staging tiles of this half of the epilogue warps and the thread that issues its bulk stores
        uint8_t* const sF = stg + hf * C::STG_HALF_BYTES;
        uint8_t* const sH = sF + 128 * CW * 4;
        uint8_t* const sL = sH + 128 * CW * 2;
        uint8_t* const sIh = PIPE ? sL + 128 * CW * 2 : sH;    // identity planes in: own tiles (PIPE) or in place
        uint8_t* const sIl = PIPE ? sIh + 128 * CW * 2 : sL;
        const bool st_leader = (etid0 & 127) == 0;
        // free the staging tiles of this half (the bulk stores of the previous chunk have read them)
        auto stage_acquire = [&]() {
            const long long ta = (CTRP && P.tmode == 3 && etid0 == 0) ? clock64() : 0;
            if (st_leader) bulk_wait_read0();
            if (CTRP && P.tmode == 3 && etid0 == 0) atomicAdd(CTRP + 0, (unsigned long long)(clock64() - ta));   // store drain
            half_bar(hf);
            if (CTRP && P.tmode == 3 && etid0 == 0) { atomicAdd(CTRP + 1, (unsigned long long)(clock64() - ta)); atomicAdd(CTRP + 3, 1ull); }   // + barrier
        };
        // make the staged rows visible to the TMA engine and write them out as boxes of the tile
        auto stage_release = [&](const CUtensorMap* mf, const CUtensorMap* mh, const CUtensorMap* ml, int c0, int op, int r0) {
            if (DBGF(2048)) r0 = (int)blockIdx.x * S;   // timing probe: every CTA keeps storing to its own first tile (same TMA work, no new HBM lines)
            fence_async_smem();
            half_bar(hf);
            if (st_leader && !DBGF(512)) {   // dbg 512: results staged but never stored (timing probe)
                if (p.ostride == 2) {
                    if (mf) tma_store_4d(mf, sF, c0, op, 0, r0);
                    if (mh) { tma_store_4d(mh, sH, c0, op, 0, r0); tma_store_4d(ml, sL, c0, op, 0, r0); }
                } else {
                    const int c1 = AR ? r0 : 0, c2 = AR ? 0 : r0;   // position-major store maps are (C, R, H)
                    if (mf) tma_store_3d(mf, sF, c0, c1, c2);
                    if (mh) { tma_store_3d(mh, sH, c0, c1, c2); tma_store_3d(ml, sL, c0, c1, c2); }
                }
                bulk_commit();
            }
        };
        // TE: the epilogue's own INPUT tiles -- the saved yhat of a GroupNorm-backward launch and the identity /
        // residual planes -- come in through the same staging tiles by TMA loads (whole 128-byte lines, asynchronous)
        // instead of per-row 16-byte global loads, which cost one L1 wavefront per row and instruction; every thread
        // then reads (and later overwrites, in place) only its own row
        uint32_t lph = 0;
        const uint32_t box_rows = (uint32_t)(S * HJ);
        auto stage_load = [&](const CUtensorMap* mf, const CUtensorMap* mh, const CUtensorMap* ml, int c0, int r0) {
            if (st_leader) {
                mbar_expect_tx(&lfull[hf], box_rows * CW * ((mf ? 4u : 0u) + (mh ? 4u : 0u)));
                const int c1 = AR ? r0 : 0, c2 = AR ? 0 : r0;
                if (mf) tma_load_3d(sF, mf, &lfull[hf], c0, c1, c2);
                if (mh) { tma_load_3d(sIh, mh, &lfull[hf], c0, c1, c2); tma_load_3d(sIl, ml, &lfull[hf], c0, c1, c2); }
            }
        };
        auto stage_load_wait = [&]() {
            const long long ta = (CTRP && P.tmode == 3 && etid0 == 0) ? clock64() : 0;
            mbar_wait(&lfull[hf], lph);
            lph ^= 1;
            if (CTRP && P.tmode == 3 && etid0 == 0) { atomicAdd(CTRP + 2, (unsigned long long)(clock64() - ta)); atomicAdd(CTRP + 4, 1ull); }   // load landed
        };
        const bool tl_id = TE && p.ident_hi != nullptr;   // the host only selects TE kernels for unit-stride identity terms
        int it = 0;
        for (int tile = sched0; tile < nsched; tile += sched_step, ++it) {
            if (TEAM) {
                if ((it & 1) != hf) continue;
                as = it & 3;
                aph = (uint32_t)((it >> 2) & 1);
            }
            const int op = tile % P.nphase, t2 = tile / P.nphase;
            const int r0 = ((t2 / P.tilesN) * CG + (int)crank) * S, n0 = (t2 % P.tilesN) * NT;
            const int Sv = max(0, min(S, p.R - r0));
            const int Mv = Sv * HJ;
            for (int i = etid; i < NT; i += ESTEP) {
                const bool cv = n0 + i < N;
                scol[i] = (p.bias && cv) ? __ldg(p.bias + n0 + i) : 0.f;
                scol[NT + i] = (p.gamma && cv) ? __ldg(p.gamma + n0 + i) : 0.f;
                scol[2 * NT + i] = (p.beta && cv) ? __ldg(p.beta + n0 + i) : 0.f;
                scol[3 * NT + i] = (p.bias2 && cv) ? __ldg(p.bias2 + n0 + i) : 0.f;
            }
            if (C::INL_FLOATS && p.inl_x)
                for (int i = etid; i < p.inl_taps * p.inl_D * NT; i += ESTEP) {
                    const int td = i / NT, nn = i - td * NT;
                    sinl[i] = (n0 + nn < N) ? __ldg(p.inl_w + (size_t)td * N + n0 + nn) : 0.f;
                }
            if (GS)
                for (int i = etid; i < C::PART_FLOATS; i += ESTEP) spart[i] = 0.f;
            // per-row constants
            // accumulator row -> (sample, position): sample-major tiles, position-major (row = h*S + s) with A reuse
            const int s_loc = AR ? (mrow % S) : min(mrow / HJ, S - 1);
            const int hrow_ = AR ? (mrow / S) : (mrow - s_loc * HJ);
            const bool valid = AR ? (hrow_ < HJ && s_loc < Sv) : (mrow < Mv);
            const int rg = valid ? r0 + s_loc : 0;                              // global sample (row of R)
            const size_t orow = valid ? (size_t)(r0 + s_loc) * p.Hout + (size_t)(hrow_ * p.ostride + op) : 0;
            const int slab0 = s_loc * G;
            if (PIPE) {
                // inputs of this thread group's first chunk: in flight while the accumulator is still being computed
                // (the input tiles are free: every thread has passed the last input barrier of the previous tile)
                if (GS && p.gn_bwd) stage_load(&P.mSy, tl_id ? &P.mIh : nullptr, &P.mIl, n0 + ch0 * CW, r0);
                else if (tl_id) stage_load(nullptr, &P.mIh, &P.mIl, n0 + ch0 * CW, r0);
            }
            const long long tq0 = (CTRP && etid == 0) ? clock64() : 0;
            if (CTRP && P.tmode == 2 && etid == 0 && tlap) atomicAdd(CTRP + 5, (unsigned long long)(tq0 - tlap));   // preamble
            mbar_wait(&tfull[as], aph);
            const long long tq1 = (CTRP && etid == 0) ? clock64() : 0;
            tlap = tq1;
            if (CTRP && P.tmode == 2 && etid == 0) atomicAdd(CTRP + 0, (unsigned long long)(tq1 - tq0));
            tc_fence_after();
            ebar();
            if (DBGF(1)) {
                // main-loop-only timing runs: results are garbage
                tc_fence_before();
                ebar();
                if (etid == 0) {
                    if (CG == 2) mbar_arrive_cluster(&tempty[as], 0); else mbar_arrive(&tempty[as]);
                    if (CTRP) atomicAdd(CTRP + 6, 1ull);
                }
                if (!TEAM && ++as == NACC) { as = 0; aph ^= 1; }
                continue;
            }
            const uint32_t tacc = tmem_base + ((uint32_t)(eq * 32) << 16) + (uint32_t)(as * C::ACC_COLS);
            const uint32_t tstash = tmem_base + ((uint32_t)(eq * 32) << 16) + (uint32_t)(NACC * C::ACC_COLS + as * NT);
            float v[CW];
            const int hrow = hrow_;
            if (C::INL_FLOATS && p.inl_mode == 1) {
                // the convolution itself has K = taps*D <= 70: compute it here and park it in the (unused)
                // accumulator so that the GroupNorm passes below run unchanged
                for (int ch = ch0; ch < chE; ch += CHS) {
#pragma unroll
                    for (int j = 0; j < CW; ++j) v[j] = 0.f;
                    inline_conv<NT, CW>(p, sinl, ch, rg, hrow, HJ, valid, v);
                    tmem_stw<CW>(tacc + ch * CW, v);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }

            if (GS && p.gn_fwd) {
                // ---- pass 1: per-row sums of y = acc + bias and y^2 for every group
#pragma unroll 1
                for (int ch = ch0; ch < chE; ch += CHS) {
                    tmem_ldw<CW>(tacc + ch * CW, v);
                    constexpr int GPC = (GS && GS < CW) ? CW / GS : 1;   // groups per chunk
                    float s1[GPC], s2[GPC];
#pragma unroll
                    for (int g = 0; g < GPC; ++g) s1[g] = s2[g] = 0.f;
                    add_ldsw<CW>(scol + ch * CW, v);
#pragma unroll
                    for (int j = 0; j < CW; ++j) {
                        const float x = v[j];
                        s1[GPC > 1 ? j / (CW / GPC) : 0] += x;
                        s2[GPC > 1 ? j / (CW / GPC) : 0] += x * x;
                    }
                    const int g0 = (ch * CW) / (GS ? GS : 1);
#pragma unroll
                    for (int g = 0; g < GPC; ++g) {
                        float* d = spart + ((pslot(ch) * 128 + mrow) * G + g0 + g) * 2;
                        d[0] += s1[g];
                        d[1] += s2[g];
                    }
                }
                lap(1);
                ebar();
                {
                    // slab (s, g): the 2*HJ row partials are summed by one thread in 4 independent
                    // chains (fixed order -> deterministic) instead of one dependent LDS+FADD chain
                }
                if (etid < S * G) {
                    const int s = etid / G, g = etid - s * G;
                    float a4[4] = {0.f, 0.f, 0.f, 0.f}, b4[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int h2 = 0; h2 < C::PSLOTS; ++h2) {
                        const float* base = spart + ((h2 * 128 + (AR ? s : s * HJ)) * G + g) * 2;   // + j * RS rows
                        int j = 0;
                        for (; j + 4 <= HJ; j += 4) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                a4[u] += base[(j + u) * RS * G * 2];
                                b4[u] += base[(j + u) * RS * G * 2 + 1];
                            }
                        }
                        for (; j < HJ; ++j) {
                            a4[0] += base[j * RS * G * 2];
                            b4[0] += base[j * RS * G * 2 + 1];
                        }
                    }
                    const float a = (a4[0] + a4[1]) + (a4[2] + a4[3]), b = (b4[0] + b4[1]) + (b4[2] + b4[3]);
                    const float mean = a * inv_cnt;
                    const float var = fmaxf(b * inv_cnt - mean * mean, 0.f);
                    const float rs = 1.0f / sqrtf(var + 1e-5f);
                    sstat[etid] = mean;
                    sstat[128 + etid] = rs;
                    if (s < Sv) p.rstd[(size_t)(r0 + s) * 8 + n0 / (GS ? GS : 1) + g] = rs;
                }
                ebar();
                lap(2);
            }

            if (!(GS && p.gn_bwd)) {
                // ---- forward (with or without GroupNorm) and plain backward: one output pass
                int tt = 0;
                if (p.embT && valid) {
                    const int64_t t64 = p.tidx[rg];
                    tt = (int)(t64 < 0 ? 0 : (t64 > p.tmax ? p.tmax : t64));
                }
                const float* eT = p.embT ? p.embT + (size_t)tt * p.ldeT + p.eoff + n0 : nullptr;
                const float* eR = (p.embR && valid && rg < p.n_cond) ? p.embR + (size_t)rg * p.ldeR + p.eoff + n0 : nullptr;
                const float* idp = p.ident ? p.ident + orow * p.ldi + n0 : nullptr;
                const __nv_bfloat16* idh = (!TE && p.ident_hi) ? p.ident_hi + orow * p.ldi + n0 : nullptr;
                const __nv_bfloat16* idl = (!TE && p.ident_hi) ? p.ident_lo + orow * p.ldi + n0 : nullptr;
                const float* adp = p.addend ? p.addend + orow * p.ldad + n0 : nullptr;
#pragma unroll 1
                for (int ch = ch0; ch < chE; ch += CHS) {
                    tmem_ldw<CW>(tacc + ch * CW, v);
                    add_ldsw<CW>(scol + ch * CW, v);   // + bias
                    if (GS && p.gn_fwd) {
                        constexpr int GPC2 = (GS && GS < CW) ? CW / GS : 1;
                        float mu[GPC2], rsd[GPC2];
#pragma unroll
                        for (int g = 0; g < GPC2; ++g) {
                            const int sl = slab0 + (ch * CW) / (GS ? GS : 1) + g;
                            mu[g] = sstat[sl];
                            rsd[g] = sstat[128 + sl];
                        }
#pragma unroll
                        for (int j = 0; j < CW; ++j)
                            v[j] = (v[j] - mu[GPC2 > 1 ? j / (CW / GPC2) : 0]) * rsd[GPC2 > 1 ? j / (CW / GPC2) : 0];   // yhat
                        if constexpr (PIPE) {
                            // yhat is parked over the accumulator chunk it came from until the output tiles are free
                            // (after the arithmetic below): no extra live registers
                            tmem_stw<CW>(tacc + ch * CW, v);
                        } else if (TE) {
                            stage_acquire();
                            if (tl_id) stage_load(nullptr, &P.mIh, &P.mIl, n0 + ch * CW, r0);
                            stage_f32w<CW>(sF, mrow, v);
                        } else if (valid && !DBGF(2)) {
                            st_roww<CW>(p.yhat + orow * p.ldy + n0 + ch * CW, v);
                        }
                        act_affinew<CW>(scol + NT + ch * CW, scol + 2 * NT + ch * CW, v, act);
                    } else if (TE && !PIPE) {
                        stage_acquire();
                        if (tl_id) stage_load(nullptr, &P.mIh, &P.mIl, n0 + ch * CW, r0);
                    }
                    if (ACC2) {
                        float w[CW];
                        tmem_ldw<CW>(tacc + NT + ch * CW, w);
                        add_ldsw<CW>(scol + 3 * NT + ch * CW, w);
#pragma unroll
                        for (int j = 0; j < CW; ++j) v[j] += w[j];
                    }
                    if (C::INL_FLOATS && p.inl_mode == 2) {
                        add_ldsw<CW>(scol + 3 * NT + ch * CW, v);
                        inline_conv<NT, CW>(p, sinl, ch, rg, hrow, HJ, valid, v);
                    }
                    if (N < n0 + ch * CW + CW) {
                        // narrow output (N = trajectory dimension): masked scalar IO, optional energy
                        if (valid) {
                            float e = 0.f;
                            for (int j = 0; j < CW; ++j) {
                                const int n = n0 + ch * CW + j;
                                if (n >= N) break;
                                float x = v[j];
                                if (idp) x += __ldg(idp + ch * CW + j);
                                if (adp) x += __ldg(adp + ch * CW + j);
                                if (p.out) p.out[orow * p.ldo + n] = x;
                                e = fmaf(0.5f * x, x, e);
                            }
                            if (p.energy) atomicAdd(p.energy + rg, e);
                        }
                        continue;
                    }
                    if (eT && !DBGF(8)) add_roww<CW>(eT + ch * CW, v);
                    if (eR && !DBGF(8)) add_roww<CW>(eR + ch * CW, v);
                    if (idp && !DBGF(4)) add_roww<CW>(idp + ch * CW, v);
                    if constexpr (TE) {
                        if (tl_id) {
                            stage_load_wait();
                            add_unstage_planesw<CW>(sIh, sIl, mrow, v);
                        }
                    } else if (idh && !DBGF(4)) {
                        add_planesw<CW>(idh + ch * CW, idl + ch * CW, v);
                    }
                    if (adp && !DBGF(4)) add_roww<CW>(adp + ch * CW, v);
                    if (TE) {
                        // with GroupNorm the fp32 tile carries yhat, so an fp32 copy of the result (only kept for
                        // identity residuals / SIMT consumers) goes out directly; without it the tile is free
                        const bool gnf = GS && p.gn_fwd;
                        if (PIPE) {
                            // one barrier: the previous chunk's stores have drained (output tiles free) and every
                            // thread has consumed this chunk's identity rows (input tiles free for the next load)
                            stage_acquire();
                            if (tl_id && ch + CHS < chE) stage_load(nullptr, &P.mIh, &P.mIl, n0 + (ch + CHS) * CW, r0);
                        }
                        if (p.out) {
                            if (gnf) { if (valid) st_roww<CW>(p.out + orow * p.ldo + n0 + ch * CW, v); }
                            else stage_f32w<CW>(sF, mrow, v);
                        }
                        if (p.out_hi) stage_planesw<CW>(sH, sL, mrow, v);
                        if constexpr (PIPE) {
                            if (gnf) {
                                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                                tmem_ldw<CW>(tacc + ch * CW, v);
                                stage_f32w<CW>(sF, mrow, v);
                            }
                        }
                        stage_release(gnf ? &P.mSy : (p.out ? &P.mSo : nullptr), p.out_hi ? &P.mSoh : nullptr, &P.mSol,
                                      n0 + ch * CW, op, r0);
                    } else if (valid && !DBGF(2)) {
                        const size_t o = orow * p.ldo + n0 + ch * CW;
                        if (p.out) st_roww<CW>(p.out + o, v);
                        if (p.out_hi) st_planesw<CW>(p.out_hi + o, p.out_lo + o, v);
                    }
                }
            } else {
                // ---- backward with GroupNorm/activation backward.
                // pass 0: x = acc (+ident +addend) -> raw output; g_yhat = x*act'(z)*gamma is written
                // back over the accumulator and yhat is stashed in spare TMEM columns, so pass 1
                // (after the slab means are known) touches no global input again.
                const float* idp = p.ident ? p.ident + orow * p.ldi + n0 : nullptr;
                const __nv_bfloat16* idh = (!TE && p.ident_hi) ? p.ident_hi + orow * p.ldi + n0 : nullptr;
                const __nv_bfloat16* idl = (!TE && p.ident_hi) ? p.ident_lo + orow * p.ldi + n0 : nullptr;
                const float* adp = p.addend ? p.addend + orow * p.ldad + n0 : nullptr;
                const float* yp = TE ? nullptr : p.yhat + orow * p.ldy + n0;
#pragma unroll 1
                for (int ch = ch0; ch < chE; ch += CHS) {
                    if (TE && !PIPE) {
                        // yhat (fp32 tile) and the identity planes of this chunk arrive by TMA
                        stage_acquire();
                        stage_load(&P.mSy, tl_id ? &P.mIh : nullptr, &P.mIl, n0 + ch * CW, r0);
                    }
                    tmem_ldw<CW>(tacc + ch * CW, v);
                    if (idp && !DBGF(4)) add_roww<CW>(idp + ch * CW, v);
                    if constexpr (!TE) {
                        if (idh && !DBGF(4)) add_planesw<CW>(idh + ch * CW, idl + ch * CW, v);
                    }
                    if (adp && !DBGF(4)) add_roww<CW>(adp + ch * CW, v);
                    float y[CW];
                    if constexpr (TE) {
                        stage_load_wait();
                        unstage_f32w<CW>(sF, mrow, y);
                        if (tl_id) add_unstage_planesw<CW>(sIh, sIl, mrow, v);
                    } else if (!DBGF(4)) ld_roww<CW>(yp + ch * CW, y);
                    else {
#pragma unroll
                        for (int j = 0; j < CW; ++j) y[j] = 0.5f;
                    }
                    uint32_t ph[PIPE ? CW / 2 : 1], pl[PIPE ? CW / 2 : 1];   // raw output planes, packed, until the tiles are free
                    if (PIPE) {
                        // every thread holds its inputs: the next chunk's (or, re-reading yhat, pass 1's first) load starts
                        half_bar(hf);
                        if (ch + CHS < chE) stage_load(&P.mSy, tl_id ? &P.mIh : nullptr, &P.mIl, n0 + (ch + CHS) * CW, r0);
                        else if (!C::STASH) stage_load(&P.mSy, nullptr, nullptr, n0 + ch0 * CW, r0);
                        if (p.out && valid) st_roww<CW>(p.out + orow * p.ldo + n0 + ch * CW, v);   // fp32 copy: not on the default path
                        if constexpr (PIPE) {
                            if (p.out_hi) {
#pragma unroll
                                for (int k = 0; k < CW / 2; ++k) split_bf16x2(v[2 * k], v[2 * k + 1], ph[k], pl[k]);
                            }
                        }
                    } else if (TE) {
                        if (p.out || p.out_hi) {
                            // in place: every thread overwrites the row it has just read
                            if (p.out) stage_f32w<CW>(sF, mrow, v);
                            if (p.out_hi) stage_planesw<CW>(sH, sL, mrow, v);
                            stage_release(p.out ? &P.mSo : nullptr, p.out_hi ? &P.mSoh : nullptr, &P.mSol, n0 + ch * CW, op, r0);
                        }
                    } else if (valid && !DBGF(2)) {
                        const size_t o = orow * p.ldo + n0 + ch * CW;
                        if (p.out) st_roww<CW>(p.out + o, v);
                        if (p.out_hi) st_planesw<CW>(p.out_hi + o, p.out_lo + o, v);
                    }
                    constexpr int GPC = (GS && GS < CW) ? CW / GS : 1;
                    float q1[GPC], q2[GPC];
#pragma unroll
                    for (int g = 0; g < GPC; ++g) q1[g] = q2[g] = 0.f;
                    if (!valid) {
#pragma unroll
                        for (int j = 0; j < CW; ++j) { v[j] = 0.f; y[j] = 0.f; }
                    }
                    act_affine_bwdw<CW>(scol + NT + ch * CW, scol + 2 * NT + ch * CW, y, v, act);
#pragma unroll
                    for (int j = 0; j < CW; ++j) {
                        q1[GPC > 1 ? j / (CW / GPC) : 0] += v[j];
                        q2[GPC > 1 ? j / (CW / GPC) : 0] += v[j] * y[j];
                    }
                    tmem_stw<CW>(tacc + ch * CW, v);
                    if (C::STASH) tmem_stw<CW>(tstash + ch * CW, y);
                    if constexpr (PIPE) {
                        if (p.out_hi) {
                            // the previous chunk's stores drained while the arithmetic above ran
                            stage_acquire();
                            stage_packedw<CW>(sH, sL, mrow, ph, pl);
                            stage_release(nullptr, &P.mSoh, &P.mSol, n0 + ch * CW, op, r0);
                        }
                    }
                    const int g0 = (ch * CW) / (GS ? GS : 1);
#pragma unroll
                    for (int g = 0; g < GPC; ++g) {
                        float* d = spart + ((pslot(ch) * 128 + mrow) * G + g0 + g) * 2;
                        d[0] += q1[g];
                        d[1] += q2[g];
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                lap(1);
                ebar();
                if (etid < S * G) {
                    const int s = etid / G, g = etid - s * G;
                    float a4[4] = {0.f, 0.f, 0.f, 0.f}, b4[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int h2 = 0; h2 < C::PSLOTS; ++h2) {
                        const float* base = spart + ((h2 * 128 + (AR ? s : s * HJ)) * G + g) * 2;   // + j * RS rows
                        int j = 0;
                        for (; j + 4 <= HJ; j += 4) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                a4[u] += base[(j + u) * RS * G * 2];
                                b4[u] += base[(j + u) * RS * G * 2 + 1];
                            }
                        }
                        for (; j < HJ; ++j) {
                            a4[0] += base[j * RS * G * 2];
                            b4[0] += base[j * RS * G * 2 + 1];
                        }
                    }
                    const float a = (a4[0] + a4[1]) + (a4[2] + a4[3]), b = (b4[0] + b4[1]) + (b4[2] + b4[3]);
                    sstat[etid] = a * inv_cnt;
                    sstat[128 + etid] = b * inv_cnt;
                    sstat[256 + etid] = (s < Sv) ? __ldg(p.rstd + (size_t)(r0 + s) * 8 + n0 / (GS ? GS : 1) + g) : 0.f;
                }
                ebar();
                lap(2);
#pragma unroll 1
                for (int ch = ch0; ch < chE; ch += CHS) {
                    float y[CW];
                    if (TE && !PIPE) {
                        stage_acquire();
                        if (!C::STASH) stage_load(&P.mSy, nullptr, nullptr, n0 + ch * CW, r0);
                    }
                    tmem_ldw<CW>(tacc + ch * CW, v);
                    if (C::STASH) {
                        tmem_ldw<CW>(tstash + ch * CW, y);
                    } else {
                        if constexpr (TE) {
                            stage_load_wait();
                            unstage_f32w<CW>(sF, mrow, y);
                            if (PIPE) {
                                half_bar(hf);
                                if (ch + CHS < chE) stage_load(&P.mSy, nullptr, nullptr, n0 + (ch + CHS) * CW, r0);
                            }
                        } else {
                            ld_roww<CW>(yp + ch * CW, y);
                        }
                        if (!valid) {
#pragma unroll
                            for (int j = 0; j < CW; ++j) y[j] = 0.f;
                        }
                    }
                    constexpr int GPC3 = (GS && GS < CW) ? CW / GS : 1;
                    float m1[GPC3], m2[GPC3], rsd[GPC3];
#pragma unroll
                    for (int g = 0; g < GPC3; ++g) {
                        const int sl = slab0 + (ch * CW) / (GS ? GS : 1) + g;
                        m1[g] = sstat[sl]; m2[g] = sstat[128 + sl]; rsd[g] = sstat[256 + sl];
                    }
#pragma unroll
                    for (int j = 0; j < CW; ++j) {
                        const int g = GPC3 > 1 ? j / (CW / GPC3) : 0;
                        v[j] = rsd[g] * (v[j] - m1[g] - y[j] * m2[g]);
                    }
                    if (PIPE) {
                        stage_acquire();
                        if (p.gy && valid) st_roww<CW>(p.gy + orow * p.ldg + n0 + ch * CW, v);   // fp32 copy: not on the default path
                        if (p.gy_hi) stage_planesw<CW>(sH, sL, mrow, v);
                        stage_release(nullptr, p.gy_hi ? &P.mSgh : nullptr, &P.mSgl, n0 + ch * CW, op, r0);
                    } else if (TE) {
                        if (p.gy) stage_f32w<CW>(sF, mrow, v);
                        if (p.gy_hi) stage_planesw<CW>(sH, sL, mrow, v);
                        stage_release(p.gy ? &P.mSg : nullptr, p.gy_hi ? &P.mSgh : nullptr, &P.mSgl, n0 + ch * CW, op, r0);
                    } else if (valid && !DBGF(2)) {
                        const size_t o = orow * p.ldg + n0 + ch * CW;
                        if (p.gy) st_roww<CW>(p.gy + o, v);
                        if (p.gy_hi) st_planesw<CW>(p.gy_hi + o, p.gy_lo + o, v);
                    }
                }
            }
            // release the accumulator stage
            lap(3);
            tc_fence_before();
            ebar();
            if (etid == 0) {
                if (CG == 2) mbar_arrive_cluster(&tempty[as], 0); else mbar_arrive(&tempty[as]);
            }
            if (CTRP && etid == 0) {
                if (P.tmode == 1) {
                    atomicAdd(CTRP + 3, (unsigned long long)(tq1 - tq0));
                    atomicAdd(CTRP + 4, (unsigned long long)(clock64() - tq1));
                } else {
                    lap(4);
                }
                atomicAdd(CTRP + 6, 1ull);
            }
            if (!TEAM && ++as == NACC) { as = 0; aph ^= 1; }
        }
        if (TE && st_leader) bulk_wait_read0();   // shared memory stays valid until the last stores have read it
    }

    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();   // no CTA of a pair exits (or frees TMEM) while its peer may still signal it
    if (warp == 1) {
        tc_fence_after();
        if (CG == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
    static EncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(ptr);
    }
    return fn;
}

using MapKey = std::tuple<const void*, int, int, int, int, int, int>;
std::map<MapKey, CUtensorMap> g_maps;
// encoded tensor maps are cached per (address, shape): a steady workload re-uses a few hundred of them; a service
// that sees many distinct batch sizes would grow the cache without bound, so it is simply dropped past a cap
void remember_map(const MapKey& key, const CUtensorMap& m) {
    if (g_maps.size() > 65536) g_maps.clear();
    g_maps[key] = m;
}

// activation plane [R][H][ld] BF16 -> 3-D map (C, H, R), box (64, H, S)
int act_map(CUtensorMap* out, const __nv_bfloat16* base, int C, int H, int R, int ld, int S, int bk) {
    MapKey key(base, C, H, R, ld, S, 3 + 100 * bk);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
    EncodeFn enc = get_encode();
    PMP_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)R};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)H};
    cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)H, (cuuint32_t)S};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PMP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation C=%d H=%d R=%d ld=%d S=%d) failed: %d", C, H, R, ld, S, (int)r);
    remember_map(key, *out);
    return 0;
}
// stride-2 gather: rows (2j + parity) -> 4-D map (C, parity, H/2, R), box (64, 1, HJ, S)
int act_map4(CUtensorMap* out, const __nv_bfloat16* base, int C, int Hin, int R, int ld, int S, int HJ, int bk) {
    MapKey key(base, C, Hin, R, ld, S * 1000 + HJ, 4 + 100 * bk);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
    EncodeFn enc = get_encode();
    PMP_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[4] = {(cuuint64_t)C, 2, (cuuint64_t)(Hin / 2), (cuuint64_t)R};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 4, (cuuint64_t)ld * 2 * (cuuint64_t)Hin};
    cuuint32_t box[4] = {(cuuint32_t)bk, 1, (cuuint32_t)HJ, (cuuint32_t)S};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PMP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(strided activation C=%d H=%d R=%d ld=%d) failed: %d", C, Hin, R, ld, (int)r);
    remember_map(key, *out);
    return 0;
}
// A reuse: haloed, position-major activation tile.  Map (C, R, H) -- the sample index runs faster than the
// position in the box, so shared-memory row = h*S + s -- box (bk, S, HJ + 4) started at position -2.
int act_map_pm(CUtensorMap* out, const __nv_bfloat16* base, int C, int H, int R, int ld, int S, int HJ, int bk) {
    MapKey key(base, C, H, R, ld, S * 1000 + HJ, 5 + 100 * bk);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
    EncodeFn enc = get_encode();
    PMP_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)R, (cuuint64_t)H};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2 * (cuuint64_t)H, (cuuint64_t)ld * 2};
    cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)S, (cuuint32_t)(HJ + 4)};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PMP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(position-major activation C=%d H=%d R=%d ld=%d S=%d) failed: %d", C, H, R,
                ld, S, (int)r);
    remember_map(key, *out);
    return 0;
}
// weight plane [taps*N][C] BF16 -> 2-D map (C, taps*N), box (64, NT)
int w_map(CUtensorMap* out, const __nv_bfloat16* base, int C, int rows, int NT, int bk) {
    MapKey key(base, C, rows, NT, 0, 0, 2 + 100 * bk);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
    EncodeFn enc = get_encode();
    PMP_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)NT};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PMP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weight C=%d rows=%d NT=%d) failed: %d", C, rows, NT, (int)r);
    remember_map(key, *out);
    return 0;
}

// output tensor [R][Hout][ld] (fp32 or BF16) -> store map with box (32, HJ, S): 3-D (C, H, R) for unit output
// stride, 4-D (C, phase, H/2, R) for the stride-2 scatter of the transposed convs
int st_map(CUtensorMap* out, const void* base, bool f32, int C, int Hout, int R, int ld, int S, int HJ, int ostride,
           int cw, bool pm = false) {
    MapKey key(base, C, Hout, R, ld, S * 1000 + HJ, (f32 ? 10 : 20) + ostride + 1000 * cw + (pm ? 100000 : 0));
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
    EncodeFn enc = get_encode();
    PMP_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t es = f32 ? 4 : 2;
    const CUtensorMapDataType dt = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const int rowb = cw * (int)es;   // bytes per staged row: 128 / 64 / 32 -> the matching swizzle span
    const CUtensorMapSwizzle sw = rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (rowb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r;
    if (pm) {
        // position-major staging tiles (A-reuse kernels): box (cw, S, HJ) of the (C, R, H) view
        cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)R, (cuuint64_t)Hout};
        cuuint64_t strides[2] = {(cuuint64_t)ld * es * (cuuint64_t)Hout, (cuuint64_t)ld * es};
        cuuint32_t box[3] = {(cuuint32_t)cw, (cuuint32_t)S, (cuuint32_t)HJ};
        cuuint32_t el[3] = {1, 1, 1};
        r = enc(out, dt, 3, const_cast<void*>(base), dims, strides, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else if (ostride == 1) {
        cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)Hout, (cuuint64_t)R};
        cuuint64_t strides[2] = {(cuuint64_t)ld * es, (cuuint64_t)ld * es * (cuuint64_t)Hout};
        cuuint32_t box[3] = {(cuuint32_t)cw, (cuuint32_t)HJ, (cuuint32_t)S};
        cuuint32_t el[3] = {1, 1, 1};
        r = enc(out, dt, 3, const_cast<void*>(base), dims, strides, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t dims[4] = {(cuuint64_t)C, 2, (cuuint64_t)(Hout / 2), (cuuint64_t)R};
        cuuint64_t strides[3] = {(cuuint64_t)ld * es, (cuuint64_t)ld * es * 2, (cuuint64_t)ld * es * (cuuint64_t)Hout};
        cuuint32_t box[4] = {(cuuint32_t)cw, 1, (cuuint32_t)HJ, (cuuint32_t)S};
        cuuint32_t el[4] = {1, 1, 1, 1};
        r = enc(out, dt, 4, const_cast<void*>(base), dims, strides, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    PMP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(store C=%d H=%d R=%d ld=%d S=%d HJ=%d) failed: %d", C, Hout, R, ld, S,
                HJ, (int)r);
    remember_map(key, *out);
    return 0;
}

int g_num_sms = 0;

// A/B switches of the kernel variants: read from the environment in -DPMP_DIAG builds only
int diag_env(const char* name, int dflt) {
#ifdef PMP_DIAG
    if (const char* e = getenv(name)) return atoi(e);
#endif
    (void)name;
    return dflt;
}
}  // namespace
// role timers of conv_tc_kernel, enabled with PMP_TC_TIMERS=1: [0] MMA warp waiting for TMA data,
// [1] MMA warp waiting for a free accumulator (epilogue behind), [2] MMA warp total, [3] epilogue waiting
// for the accumulator, [4] epilogue busy, [5] TMA producer waiting for a free stage, [6] tiles
unsigned long long* tc_timer_buffer() {
#ifndef PMP_DIAG
    return nullptr;
#else
    static unsigned long long* buf = nullptr;
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("PMP_TC_TIMERS");
        on = (e && atoi(e)) ? 1 : 0;
        if (on && cudaMalloc((void**)&buf, 8 * sizeof(unsigned long long)) == cudaSuccess)
            cudaMemset(buf, 0, 8 * sizeof(unsigned long long));
    }
    return buf;
#endif
}
namespace {

template <int NT, int GS, bool ACC2, int CG, bool TE, int NG = 2, bool TEAM = false, bool AR = false>
int launch_t(TcP& P, cudaStream_t st) {
    using C = Cfg<NT, GS, ACC2, CG, TE, NG, TEAM, AR>;
    static bool configured = false;
    if (!configured) {
        PMP_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<NT, GS, ACC2, CG, TE, NG, TEAM, AR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES));
        configured = true;
    }
    if (!g_num_sms) {
        int dev = 0;
        PMP_CUDA_OK(cudaGetDevice(&dev));
        PMP_CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    if (CG == 1) {
        const int grid = P.ntiles < g_num_sms ? P.ntiles : g_num_sms;
        conv_tc_kernel<NT, GS, ACC2, CG, TE, NG, TEAM, AR><<<grid, C::THREADS, C::SMEM_BYTES, st>>>(P);
    } else {
        // one cluster of two CTAs (the two SMs of a TPC) per tile pair, persistent over the pairs
        const int nclusters = P.npairs < g_num_sms / 2 ? P.npairs : g_num_sms / 2;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(2 * nclusters, 1, 1);
        cfg.blockDim = dim3(C::THREADS, 1, 1);
        cfg.dynamicSmemBytes = C::SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        PMP_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc_kernel<NT, GS, ACC2, CG, TE, NG, TEAM, AR>, P));
    }
    PMP_LAUNCH_OK();
    return 0;
}

int pick_nt(const ConvP& p, bool wide = false) {
    const bool gn = p.gn_fwd || p.gn_bwd;
    if (gn) {
        const int gs = p.N / 8;
        if (gs == 8) return 64;
        if (gs == 32 || gs == 64) return wide ? 256 : 128;
        if (gs == 96) return 192;
        return 0;
    }
    if (wide && p.N % 256 == 0) return 256;
    if (p.N % 128 == 0) return 128;
    if (p.N % 64 == 0 || p.N < 64) return 64;
    return 0;
}
}  // namespace

bool conv_tc_eligible(const ConvP& p) {
    const bool s1 = p.nphase == 1 && p.istride == 1 && p.ostride == 1;             // conv k5 / 1x1
    const bool s2 = p.nphase == 1 && p.istride == 2 && p.ostride == 1;             // stride-2 gather
    const bool t2 = p.nphase == 2 && p.istride == 1 && p.ostride == 2;             // stride-2 scatter (2 phases)
    if (!s1 && !s2 && !t2) return false;
    if (!s1 && (p.A2 || p.A2_hi)) return false;
    if (t2 && (p.gn_fwd || p.gn_bwd)) return false;
    if (s2 && (p.Hin & 1)) return false;
    const bool inl_pre = p.inl_x && p.inl_mode == 1, inl_post = p.inl_x && p.inl_mode == 2;
    if ((inl_pre || inl_post) && (!s1 || p.inl_D > 16 || p.inl_taps > 5 || p.N != 64)) return false;
    if (!inl_pre) {
        if (!p.A1_hi || !p.A1_lo || !p.W1_hi || !p.W1_lo) return false;
        if (p.C1 % BK != 0 || (p.lda1 % 8) != 0) return false;
    }
    if ((p.A2 || p.A2_hi) && !inl_post &&
        (!p.A2_hi || !p.A2_lo || !p.W2_hi || !p.W2_lo || p.C2 % BK != 0 || (p.lda2 % 8) != 0))
        return false;
    if (p.N < 64 && (p.gn_fwd || p.gn_bwd || p.out_hi || p.embT || p.embR || p.ident_hi || !s1)) return false;
    if (p.HJ < 8 || p.HJ > 128 || p.HJ > 256) return false;
    const int nt = pick_nt(p);
    if (!nt || (p.N % nt != 0 && p.N >= 64)) return false;
    if ((p.gn_fwd || p.gn_bwd) && (128 / p.HJ) * (nt / (p.N / 8)) > 128) return false;
    return true;
}

// mode 3 (gemm path 3): TMA-store epilogue, on CTA pairs (cta_group::2) for the 128-channel tiles -- the pair
// halves the B operand staged per CTA, which frees the shared memory the staging tiles need; the narrow
// level-0 tiles (NT = 64) keep single CTAs; the 192-channel tiles (Dual Kuka) use pairs with the direct epilogue
static int launch_conv_tc_mode(const ConvP& p, cudaStream_t st, int mode) {
    if (!conv_tc_eligible(p)) return 1;
    TcP P;
    memset(&P, 0, sizeof P);
    P.e = p;
    const bool inl_pre = p.inl_x && p.inl_mode == 1, inl_post = p.inl_x && p.inl_mode == 2;
    const bool acc2 = (p.A2 != nullptr || p.A2_hi != nullptr) && !inl_post && (p.bias2 != nullptr || p.gn_fwd);
    // 256-column tiles (CTA pairs only) feed the tensor pipe best; a separate second accumulator does not fit next to them
    static int wide = -1;
    if (wide < 0) wide = diag_env("PMP_TC_WIDE", 1);
    const int NT = pick_nt(p, wide && (mode == 3 || mode == 4) && !acc2 && !p.inl_x);
    const bool m34 = mode == 3 || mode == 4;
    // CTA pairs also for the 64-column tiles (no inline small-K conv): a tcgen05.mma costs ~140 cycles here whatever its
    // N (measured, r2l), so M = 256 per instruction halves the main-loop time of the narrow layers
    static int pair64 = -1;
    if (pair64 < 0) pair64 = diag_env("PMP_TC_PAIR64", 1);
    const int cg = (m34 && !p.inl_x && (NT >= 128 || (NT == 64 && pair64 && p.N == 64))) ? 2 : 1;
    const bool te = m34 && (NT == 64 || ((NT == 128 || NT == 256) && cg == 2)) && p.N >= 64 &&
                    !(p.ident_hi && p.ostride != 1);   // identity planes come in by TMA load (unit-stride box maps)
    // 16 epilogue warps on 16-column chunks (path 4 forces them everywhere; A/B switch).  Round 1 used them for the
    // GroupNorm-backward launches (5-8 % faster while those epilogues were the bottleneck); since the epilogue inputs
    // arrive by TMA and its arrays stay in registers the launches are main-loop bound and 8 warps are 4.5 % faster
    // over the whole evaluation (same-box A/B, r2j), so path 3 no longer uses them by default
    static int ng4_bwd = -1;
    if (ng4_bwd < 0) ng4_bwd = diag_env("PMP_TC_NG4_BWD", 0);
    const bool ng4 = te && !(NT == 64 && cg == 2) && (mode == 4 || (mode == 3 && ng4_bwd && p.gn_bwd));   // the 64-column pair tiles only exist with 8 epilogue warps
    const int cw = ng4 ? 16 : 32;
    const int S = 128 / p.HJ;
    P.e.S = S;
    P.tilesM = (p.R + S - 1) / S;
    P.tilesN = (p.N + NT - 1) / NT;
    P.nphase = p.nphase;
    P.mode4d = p.istride == 2 ? 1 : 0;
#ifdef PMP_DIAG
    P.ctr = tc_timer_buffer();
    if (P.ctr) {
        // PMP_TC_TIMERS=1|2, optional PMP_TC_TIMER_MATCH="N:C1:g" (g = gn_fwd + 2*gn_bwd) restricts the timers to one shape
        static int tmode = 0, mN = -1, mC = -1, mG = -1;
        if (!tmode) {
            tmode = atoi(getenv("PMP_TC_TIMERS"));
            if (const char* e = getenv("PMP_TC_TIMER_MATCH")) sscanf(e, "%d:%d:%d", &mN, &mC, &mG);
        }
        P.tmode = tmode;
        if (mN >= 0 && (p.N != mN || p.C1 != mC || (p.gn_fwd + 2 * p.gn_bwd) != mG)) P.ctr = nullptr;
    }
    {
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("PMP_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
        P.dbg = dbg;
    }
#endif
    P.ntiles = P.tilesM * P.tilesN * P.nphase;
    P.npairs = ((P.tilesM + 1) / 2) * P.tilesN * P.nphase;
    static int ar_on = -1;
    if (ar_on < 0) ar_on = diag_env("PMP_TC_AR", 3);   // bit 1: 64-column tiles too (+2 % over the evaluation once their main loop had become the limit, r2i)
    // A-operand reuse for the stride-1 convs on the 256-column pair tiles and (bit 1 of PMP_TC_AR) the 64-column tiles
    const bool s1conv = p.nphase == 1 && p.istride == 1 && p.ostride == 1;
    const bool ar = m34 && te && s1conv &&
                    (((ar_on & 1) && NT == 256) || ((ar_on & 2) && NT == 64 && !p.inl_x && !acc2));
    const int bk = (NT == 256 || ar) ? 32 : 64;   // Cfg::BKC
    P.kpt = inl_pre ? 0 : p.C1 / bk;
    P.nkb2 = ((p.A2 || p.A2_hi) && !inl_post) ? p.C2 / bk : 0;
    const int brows = NT / cg;   // B rows staged per CTA
    int rc;
    if (ar) {
        if ((rc = act_map_pm(&P.mA1h, p.A1_hi, p.C1, p.Hin, p.R, p.lda1, S, p.HJ, bk))) return rc;
        if ((rc = act_map_pm(&P.mA1l, p.A1_lo, p.C1, p.Hin, p.R, p.lda1, S, p.HJ, bk))) return rc;
        if ((rc = w_map(&P.mW1h, p.W1_hi, p.C1, p.ntaps[0] * p.N, brows, bk))) return rc;
        if ((rc = w_map(&P.mW1l, p.W1_lo, p.C1, p.ntaps[0] * p.N, brows, bk))) return rc;
    } else if (!inl_pre) {
        if (P.mode4d) {
            if ((rc = act_map4(&P.mA1h, p.A1_hi, p.C1, p.Hin, p.R, p.lda1, S, p.HJ, bk))) return rc;
            if ((rc = act_map4(&P.mA1l, p.A1_lo, p.C1, p.Hin, p.R, p.lda1, S, p.HJ, bk))) return rc;
        } else {
            if ((rc = act_map(&P.mA1h, p.A1_hi, p.C1, p.Hin, p.R, p.lda1, S, bk))) return rc;
            if ((rc = act_map(&P.mA1l, p.A1_lo, p.C1, p.Hin, p.R, p.lda1, S, bk))) return rc;
        }
        int maxk = 0;
        for (int ph = 0; ph < p.nphase; ++ph)
            for (int t = 0; t < p.ntaps[ph]; ++t) maxk = p.tap_k[ph][t] > maxk ? p.tap_k[ph][t] : maxk;
        if ((rc = w_map(&P.mW1h, p.W1_hi, p.C1, (maxk + 1) * p.N, brows, bk))) return rc;
        if ((rc = w_map(&P.mW1l, p.W1_lo, p.C1, (maxk + 1) * p.N, brows, bk))) return rc;
    }
    if ((p.A2 || p.A2_hi) && !inl_post && ar) {
        if ((rc = act_map_pm(&P.mA2h, p.A2_hi, p.C2, p.Hin, p.R, p.lda2, S, p.HJ, bk))) return rc;
        if ((rc = act_map_pm(&P.mA2l, p.A2_lo, p.C2, p.Hin, p.R, p.lda2, S, p.HJ, bk))) return rc;
        if ((rc = w_map(&P.mW2h, p.W2_hi, p.C2, p.N, brows, bk))) return rc;
        if ((rc = w_map(&P.mW2l, p.W2_lo, p.C2, p.N, brows, bk))) return rc;
    } else if ((p.A2 || p.A2_hi) && !inl_post) {
        if ((rc = act_map(&P.mA2h, p.A2_hi, p.C2, p.Hin, p.R, p.lda2, S, bk))) return rc;
        if ((rc = act_map(&P.mA2l, p.A2_lo, p.C2, p.Hin, p.R, p.lda2, S, bk))) return rc;
        if ((rc = w_map(&P.mW2h, p.W2_hi, p.C2, p.N, brows, bk))) return rc;
        if ((rc = w_map(&P.mW2l, p.W2_lo, p.C2, p.N, brows, bk))) return rc;
    }
    if (te) {
        const int os = p.ostride;
        if (p.yhat && p.gn_fwd && (rc = st_map(&P.mSy, p.yhat, true, p.N, p.Hout, p.R, p.ldy, S, p.HJ, os, cw, ar))) return rc;
        if (p.out && (rc = st_map(&P.mSo, p.out, true, p.N, p.Hout, p.R, p.ldo, S, p.HJ, os, cw, ar))) return rc;
        if (p.out_hi) {
            if ((rc = st_map(&P.mSoh, p.out_hi, false, p.N, p.Hout, p.R, p.ldo, S, p.HJ, os, cw, ar))) return rc;
            if ((rc = st_map(&P.mSol, p.out_lo, false, p.N, p.Hout, p.R, p.ldo, S, p.HJ, os, cw, ar))) return rc;
        }
        if (p.gy && (rc = st_map(&P.mSg, p.gy, true, p.N, p.Hout, p.R, p.ldg, S, p.HJ, os, cw, ar))) return rc;
        if (p.gy_hi) {
            if ((rc = st_map(&P.mSgh, p.gy_hi, false, p.N, p.Hout, p.R, p.ldg, S, p.HJ, os, cw, ar))) return rc;
            if ((rc = st_map(&P.mSgl, p.gy_lo, false, p.N, p.Hout, p.R, p.ldg, S, p.HJ, os, cw, ar))) return rc;
        }
        // epilogue inputs that come in by TMA load: saved yhat (GroupNorm backward), identity planes
        if (p.yhat && p.gn_bwd && (rc = st_map(&P.mSy, p.yhat, true, p.N, p.Hout, p.R, p.ldy, S, p.HJ, os, cw, ar))) return rc;
        if (p.ident_hi) {
            if ((rc = st_map(&P.mIh, p.ident_hi, false, p.N, p.Hout, p.R, p.ldi, S, p.HJ, os, cw, ar))) return rc;
            if ((rc = st_map(&P.mIl, p.ident_lo, false, p.N, p.Hout, p.R, p.ldi, S, p.HJ, os, cw, ar))) return rc;
        }
    }
    const bool gn = p.gn_fwd || p.gn_bwd;
    const int gs = gn ? p.N / 8 : 0;
    if (ar && NT == 64 && cg == 2) {
        if (gs == 8) return launch_t<64, 8, false, 2, true, 2, true, true>(P, st);
        return launch_t<64, 0, false, 2, true, 2, true, true>(P, st);
    }
    if (ar && NT == 64) {
        if (ng4) {
            if (gs == 8) return launch_t<64, 8, false, 1, true, 4, false, true>(P, st);
            return launch_t<64, 0, false, 1, true, 4, false, true>(P, st);
        }
        if (gs == 8) return launch_t<64, 8, false, 1, true, 2, true, true>(P, st);
        return launch_t<64, 0, false, 1, true, 2, true, true>(P, st);
    }
    if (ar) {
        if (ng4) {
            if (gs == 32) return launch_t<256, 32, false, 2, true, 4, false, true>(P, st);
            if (gs == 64) return launch_t<256, 64, false, 2, true, 4, false, true>(P, st);
            return launch_t<256, 0, false, 2, true, 4, false, true>(P, st);
        }
        if (gs == 32) return launch_t<256, 32, false, 2, true, 2, false, true>(P, st);
        if (gs == 64) return launch_t<256, 64, false, 2, true, 2, false, true>(P, st);
        return launch_t<256, 0, false, 2, true, 2, false, true>(P, st);
    }
    if (ng4) {
        if (NT == 256) {
            if (gs == 32) return launch_t<256, 32, false, 2, true, 4>(P, st);
            if (gs == 64) return launch_t<256, 64, false, 2, true, 4>(P, st);
            return launch_t<256, 0, false, 2, true, 4>(P, st);
        }
        if (NT == 128) {
            if (gs == 32) return acc2 ? launch_t<128, 32, true, 2, true, 4>(P, st) : launch_t<128, 32, false, 2, true, 4>(P, st);
            if (gs == 64) return acc2 ? launch_t<128, 64, true, 2, true, 4>(P, st) : launch_t<128, 64, false, 2, true, 4>(P, st);
            return acc2 ? launch_t<128, 0, true, 2, true, 4>(P, st) : launch_t<128, 0, false, 2, true, 4>(P, st);
        }
        if (gs == 8) return acc2 ? launch_t<64, 8, true, 1, true, 4>(P, st) : launch_t<64, 8, false, 1, true, 4>(P, st);
        return acc2 ? launch_t<64, 0, true, 1, true, 4>(P, st) : launch_t<64, 0, false, 1, true, 4>(P, st);
    }
    if (NT == 256) {
        if (gs == 32) return launch_t<256, 32, false, 2, true>(P, st);
        if (gs == 64) return launch_t<256, 64, false, 2, true>(P, st);
        return launch_t<256, 0, false, 2, true>(P, st);
    }
    if (NT == 64 && cg == 2) {
        if (!acc2) {
            if (gs == 8) return launch_t<64, 8, false, 2, true, 2, true>(P, st);
            return launch_t<64, 0, false, 2, true, 2, true>(P, st);
        }
        if (gs == 8) return launch_t<64, 8, true, 2, true>(P, st);
        return launch_t<64, 0, true, 2, true>(P, st);
    }
    if (NT == 64) {
        if (te) {
            static int team = -1;
            if (team < 0) team = diag_env("PMP_TC_TEAM", 1);
            if (team && !acc2) {
                if (gs == 8) return launch_t<64, 8, false, 1, true, 2, true>(P, st);
                return launch_t<64, 0, false, 1, true, 2, true>(P, st);
            }
            if (gs == 8) return acc2 ? launch_t<64, 8, true, 1, true>(P, st) : launch_t<64, 8, false, 1, true>(P, st);
            return acc2 ? launch_t<64, 0, true, 1, true>(P, st) : launch_t<64, 0, false, 1, true>(P, st);
        }
        if (gs == 8) return acc2 ? launch_t<64, 8, true, 1, false>(P, st) : launch_t<64, 8, false, 1, false>(P, st);
        return acc2 ? launch_t<64, 0, true, 1, false>(P, st) : launch_t<64, 0, false, 1, false>(P, st);
    }
    if (cg == 2) {
        if (NT == 128) {
            if (gs == 32) return acc2 ? launch_t<128, 32, true, 2, true>(P, st) : launch_t<128, 32, false, 2, true>(P, st);
            if (gs == 64) return acc2 ? launch_t<128, 64, true, 2, true>(P, st) : launch_t<128, 64, false, 2, true>(P, st);
            return acc2 ? launch_t<128, 0, true, 2, true>(P, st) : launch_t<128, 0, false, 2, true>(P, st);
        }
        if (NT == 192) return acc2 ? launch_t<192, 96, true, 2, false>(P, st) : launch_t<192, 96, false, 2, false>(P, st);
        return 1;
    }
    if (NT == 128) {
        if (gs == 32) return acc2 ? launch_t<128, 32, true, 1, false>(P, st) : launch_t<128, 32, false, 1, false>(P, st);
        if (gs == 64) return acc2 ? launch_t<128, 64, true, 1, false>(P, st) : launch_t<128, 64, false, 1, false>(P, st);
        return acc2 ? launch_t<128, 0, true, 1, false>(P, st) : launch_t<128, 0, false, 1, false>(P, st);
    }
    if (NT == 192) return acc2 ? launch_t<192, 96, true, 1, false>(P, st) : launch_t<192, 96, false, 1, false>(P, st);
    return 1;
}

int launch_conv_tc(const ConvP& p, cudaStream_t st) { return launch_conv_tc_mode(p, st, 1); }
// gemm path 3: CTA pairs + TMA-store epilogue where they apply, the single-CTA kernel otherwise
int launch_conv_tc_pair(const ConvP& p, cudaStream_t st) { return launch_conv_tc_mode(p, st, 3); }
// gemm path 4: path 3 with 16 epilogue warps walking 16-column chunks
int launch_conv_tc_pair16(const ConvP& p, cudaStream_t st) { return launch_conv_tc_mode(p, st, 4); }

}  // namespace pmp
