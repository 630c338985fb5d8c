// conv_tc2: operand-reuse variant of the tcgen05 implicit-GEMM Conv1d (stride 1, k = 5 or 1,
// optional 1x1 slab in the same accumulator).  Same math, epilogues and BF16x3 arithmetic as
// conv_tc.cu (see there for the reference citations); what changes is the main loop, which in
// conv_tc.cu is L2->smem bandwidth bound (every tap re-reads its A tile, every M tile re-reads B):
//
//   * rows of a tile are POSITION-major (m = h*S + s).  The haloed A tile of a 32-channel chunk
//     (positions -2 .. H+1 of S samples) is loaded ONCE by a single 3-D TMA box (channels, samples,
//     positions; out-of-bounds positions and samples are zero filled) and the 5 taps read it through
//     shared-memory descriptors whose start address is advanced by tap*S rows;
//   * one CTA tile is TWO such M tiles (2*S samples) that share every B (weight) tile, i.e. two
//     TMEM accumulators per tile, double buffered (4 x NT columns);
//   => L2->smem bytes per issued MMA cycle drop ~2.5x (A: 5x fewer loads, B: shared by 256 rows).
//
// K chunk = 32 channels (64-byte rows, SWIZZLE_64B) so that two A stages (2 tiles x hi/lo planes)
// and four B stages fit next to the epilogue scratch in 227 KB.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <tuple>

#include "common.cuh"
#include "tc_common.cuh"

namespace pmp {
namespace {
using namespace tc;

constexpr int T2_THREADS = 320;              // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int BK2 = 32;                      // channels per chunk = one 64-byte swizzle row
constexpr int A_ROWS = 192;                  // smem rows per A plane (>= 4*S + 128)
constexpr int A_PLANE = A_ROWS * BK2 * 2;    // 12 KB
constexpr int A_STAGE = 4 * A_PLANE;         // 2 tiles x (hi, lo)
constexpr int NSTA = 2;

struct Tc2P {
    ConvP e;
    int tilesM, tilesN, ntiles;
    int nch1, nch2;   // 32-channel chunks of slab 1 / slab 2
    int dbg;          // PMP_TC2_DBG: 1 = skip epilogue global stores, 2 = also skip epilogue global loads (timing experiments)
    alignas(64) CUtensorMap mA1h;
    alignas(64) CUtensorMap mA1l;
    alignas(64) CUtensorMap mA2h;
    alignas(64) CUtensorMap mA2l;
    alignas(64) CUtensorMap mW1h;
    alignas(64) CUtensorMap mW1l;
    alignas(64) CUtensorMap mW2h;
    alignas(64) CUtensorMap mW2l;
};

// K-major, 64-byte swizzle, rows of 32 BF16: 8-row groups 512 B apart (SBO = 32 x 16 B), LBO = 1
__device__ __forceinline__ uint64_t make_desc64(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}

template <int NT, int GS>
struct Cfg2 {
    static constexpr int NSTB = NT <= 128 ? 4 : 2;                  // B ring depth (smem budget)
    static constexpr int B_PLANE = NT * BK2 * 2;
    static constexpr int B_STAGE = 2 * B_PLANE;
    static constexpr int ACC_COLS = 2 * NT;                       // two M tiles
    static constexpr int NACC = (2 * ACC_COLS <= 512) ? 2 : 1;
    static constexpr int G = GS ? NT / GS : 1;
    static constexpr int NCH = NT / 32;
    static constexpr int PART_FLOATS = 2 * 128 * 8 * 2;
    static constexpr int STAT_FLOATS = 2 * 3 * 128;
    static constexpr int COL_FLOATS = 4 * NT;
    static constexpr int STG_FLOATS = 0;
    static constexpr int SMEM_BYTES = NSTA * A_STAGE + NSTB * B_STAGE + (STG_FLOATS + PART_FLOATS + STAT_FLOATS + COL_FLOATS) * 4 + 256 + 1024;
};

template <int NT, int GS>
__global__ void __launch_bounds__(T2_THREADS, 1) conv_tc2_kernel(const __grid_constant__ Tc2P P) {
    using C = Cfg2<NT, GS>;
    constexpr int G = C::G, NCH = C::NCH, NACC = C::NACC, NSTB = C::NSTB;
    const ConvP& p = P.e;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer
    uint8_t* sA = smem;
    uint8_t* sB = smem + NSTA * A_STAGE;
    float* spart = reinterpret_cast<float*>(sB + NSTB * C::B_STAGE);
    float* sstat = spart + C::PART_FLOATS;
    float* scol = sstat + C::STAT_FLOATS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(scol + C::COL_FLOATS);
    uint64_t* afull = bars;
    uint64_t* aempty = afull + NSTA;
    uint64_t* bfull = aempty + NSTA;
    uint64_t* bempty = bfull + NSTB;
    uint64_t* tfull = bempty + NSTB;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.S, Hl = p.HJ, N = p.N;
    const uint32_t bytesA = (uint32_t)((Hl + 4) * S * BK2 * 2);
    const int nch = P.nch1 + P.nch2;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&P.mA1h); prefetch_tmap(&P.mA1l); prefetch_tmap(&P.mW1h); prefetch_tmap(&P.mW1l);
        if (P.nch2) { prefetch_tmap(&P.mA2h); prefetch_tmap(&P.mA2l); prefetch_tmap(&P.mW2h); prefetch_tmap(&P.mW2l); }
        for (int i = 0; i < NSTA; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
        for (int i = 0; i < NSTB; ++i) { mbar_init(&bfull[i], 1); mbar_init(&bempty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (lane == 0) {
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
                const int r0 = (tile / P.tilesN) * 2 * S, n0 = (tile % P.tilesN) * NT;
                for (int c = 0; c < nch; ++c) {
                    const bool s2 = c >= P.nch1;
                    const int c0 = (s2 ? c - P.nch1 : c) * BK2;
                    mbar_wait(&aempty[sa], pa ^ 1);
                    mbar_expect_tx(&afull[sa], 4 * bytesA);
                    uint8_t* st = sA + sa * A_STAGE;
                    // box (32 channels, S samples, H+4 positions) starting at position -2
                    tma_load_3d(st, s2 ? &P.mA2h : &P.mA1h, &afull[sa], c0, r0, -2);
                    tma_load_3d(st + A_PLANE, s2 ? &P.mA2l : &P.mA1l, &afull[sa], c0, r0, -2);
                    tma_load_3d(st + 2 * A_PLANE, s2 ? &P.mA2h : &P.mA1h, &afull[sa], c0, r0 + S, -2);
                    tma_load_3d(st + 3 * A_PLANE, s2 ? &P.mA2l : &P.mA1l, &afull[sa], c0, r0 + S, -2);
                    if (++sa == NSTA) { sa = 0; pa ^= 1; }
                    const int nt = s2 ? 1 : p.ntaps[0];
                    for (int t = 0; t < nt; ++t) {
                        mbar_wait(&bempty[sb], pb ^ 1);
                        mbar_expect_tx(&bfull[sb], C::B_STAGE);
                        uint8_t* sbp = sB + sb * C::B_STAGE;
                        const int wr = (s2 ? 0 : p.tap_k[0][t] * N) + n0;
                        tma_load_2d(sbp, s2 ? &P.mW2h : &P.mW1h, &bfull[sb], c0, wr);
                        tma_load_2d(sbp + C::B_PLANE, s2 ? &P.mW2l : &P.mW1l, &bfull[sb], c0, wr);
                        if (++sb == NSTB) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ==============================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
            int sa = 0, sb = 0, as = 0;
            uint32_t pa = 0, pb = 0, aph = 0;
            for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
                mbar_wait(&tempty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)(as * C::ACC_COLS);
                uint32_t accum = 0;
                for (int c = 0; c < nch; ++c) {
                    const bool s2 = c >= P.nch1;
                    mbar_wait(&afull[sa], pa);
                    tc_fence_after();
                    const uint32_t abase = smem_u32(sA + sa * A_STAGE);
                    const int nt = s2 ? 1 : p.ntaps[0];
                    for (int t = 0; t < nt; ++t) {
                        mbar_wait(&bfull[sb], pb);
                        tc_fence_after();
                        const uint32_t bbase = smem_u32(sB + sb * C::B_STAGE);
                        const uint64_t bh = make_desc64(bbase), bl = make_desc64(bbase + C::B_PLANE);
                        // tap shift sh in [-2,2] -> the tile's row m reads smem row m + (sh+2)*S
                        const uint32_t roff = (uint32_t)(((s2 ? 0 : p.tap_shift[0][t]) + 2) * S * BK2 * 2);
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const uint64_t ah = make_desc64(abase + (2 * i) * A_PLANE + roff);
                            const uint64_t al = make_desc64(abase + (2 * i + 1) * A_PLANE + roff);
                            const uint32_t d = d0 + (uint32_t)(i * NT);
#pragma unroll
                            for (int k = 0; k < BK2 / 16; ++k) {
                                const uint64_t ko = (uint64_t)(k * 2);
                                tc_mma(d, ah + ko, bh + ko, idesc, (k == 0) ? accum : 1u);
                                tc_mma(d, ah + ko, bl + ko, idesc, 1u);
                                tc_mma(d, al + ko, bh + ko, idesc, 1u);
                            }
                        }
                        accum = 1u;
                        tc_commit(&bempty[sb]);
                        if (++sb == NSTB) { sb = 0; pb ^= 1; }
                    }
                    tc_commit(&aempty[sa]);
                    if (++sa == NSTA) { sa = 0; pa ^= 1; }
                }
                tc_commit(&tfull[as]);
                if (++as == NACC) { as = 0; aph ^= 1; }
            }
        }
    } else {
        // ============================== epilogue (warps 2..9) ==============================
        // warps 2..5 own M tile 0, warps 6..9 M tile 1; each thread owns one accumulator row
        // (position-major: m = h*S + s) and walks all 32-column chunks of its tile.
        // sstat planes (256 slabs each: tile*128 + s*G + g): [0] mean | m1, [1] rstd | m2, [2] rstd (bwd)
        const int eq = warp & 3;
        const int ti = (warp - 2) >> 2;          // which of the two M tiles
        const int etid = threadIdx.x - 64;       // 0..255
        const int mrow = eq * 32 + lane;
        const int act = p.act;
        const float inv_cnt = 1.0f / (float)(Hl * (GS ? GS : 1));
        constexpr int GPC = (GS && GS < 32) ? 32 / GS : 1;   // groups per 32-column chunk
        int as = 0;
        uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
            const int rbase = (tile / P.tilesN) * 2 * S, n0 = (tile % P.tilesN) * NT;
            const int r0 = rbase + ti * S;
            for (int i = etid; i < NT; i += 256) {
                scol[i] = p.bias ? __ldg(p.bias + n0 + i) : 0.f;
                scol[NT + i] = p.gamma ? __ldg(p.gamma + n0 + i) : 0.f;
                scol[2 * NT + i] = p.beta ? __ldg(p.beta + n0 + i) : 0.f;
            }
            const int h = mrow / S, s_loc = mrow - h * S;
            const bool valid = (h < Hl) && (r0 + s_loc < p.R) 
#ifdef PMP_DIAG
                               && !(P.dbg & 1)
#endif
                ;
            const int rg = valid ? r0 + s_loc : 0;
            const size_t orow = valid ? (size_t)rg * p.Hout + h : 0;
            const int slab0 = ti * 128 + s_loc * G;
            float* mypart = spart + ((ti * 128 + mrow) * G) * 2;
            if (GS) {
#pragma unroll
                for (int g = 0; g < G; ++g) { mypart[2 * g] = 0.f; mypart[2 * g + 1] = 0.f; }
            }
            mbar_wait(&tfull[as], aph);
            tc_fence_after();
            epi_bar();
            const uint32_t tacc = tmem_base + ((uint32_t)(eq * 32) << 16) + (uint32_t)(as * C::ACC_COLS + ti * NT);
            float v[32];

            if (GS && p.gn_fwd) {
                // ---- pass 1: per-row sums of y = acc + bias and y^2 for every group
#pragma unroll 1
                for (int ch = 0; ch < NCH; ++ch) {
                    tmem_ld32(tacc + ch * 32, v);
                    float s1[GPC], s2[GPC];
#pragma unroll
                    for (int g = 0; g < GPC; ++g) s1[g] = s2[g] = 0.f;
                    add_lds32(scol + ch * 32, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float x = v[j];
                        s1[GPC > 1 ? j / (32 / GPC) : 0] += x;
                        s2[GPC > 1 ? j / (32 / GPC) : 0] += x * x;
                    }
                    const int g0 = (ch * 32) / (GS ? GS : 1);
#pragma unroll
                    for (int g = 0; g < GPC; ++g) {
                        mypart[2 * (g0 + g)] += s1[g];
                        mypart[2 * (g0 + g) + 1] += s2[g];
                    }
                }
                epi_bar();
                if (etid < 2 * S * G) {
                    const int t2 = etid / (S * G), rem = etid - t2 * S * G;
                    const int s = rem / G, g = rem - s * G;
                    float a = 0.f, b = 0.f;
                    for (int hh = 0; hh < Hl; ++hh) {
                        a += spart[((t2 * 128 + hh * S + s) * G + g) * 2];
                        b += spart[((t2 * 128 + hh * S + s) * G + g) * 2 + 1];
                    }
                    const float mean = a * inv_cnt;
                    const float var = fmaxf(b * inv_cnt - mean * mean, 0.f);
                    const float rs = 1.0f / sqrtf(var + 1e-5f);
                    const int idx = t2 * 128 + s * G + g;
                    sstat[idx] = mean;
                    sstat[256 + idx] = rs;
                    const int rr = rbase + t2 * S + s;
                    if (rr < p.R) p.rstd[(size_t)rr * 8 + n0 / (GS ? GS : 1) + g] = rs;
                }
                epi_bar();
            }

            if (!(GS && p.gn_bwd)) {
                // ---- forward (with or without GroupNorm) and plain backward: one output pass
                int tt = 0;
                if (p.embT && valid) {
                    const int64_t t64 = p.tidx[rg];
                    tt = (int)(t64 < 0 ? 0 : (t64 > p.tmax ? p.tmax : t64));
                }
                const int tt_i = (p.embT && valid) ? tt : -1;
                const int er_i = (p.embR && valid && rg < p.n_cond) ? rg : -1;
#pragma unroll 1
                for (int ch = 0; ch < NCH; ++ch) {
                    tmem_ld32(tacc + ch * 32, v);
                    add_lds32(scol + ch * 32, v);   // + bias
                    if (GS && p.gn_fwd) {
                        float mu[GPC], rsd[GPC];
#pragma unroll
                        for (int g = 0; g < GPC; ++g) {
                            const int sl = slab0 + (ch * 32) / (GS ? GS : 1) + g;
                            mu[g] = sstat[sl];
                            rsd[g] = sstat[256 + sl];
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            v[j] = (v[j] - mu[GPC > 1 ? j / (32 / GPC) : 0]) * rsd[GPC > 1 ? j / (32 / GPC) : 0];   // yhat
                        if (valid) st_row32(p.yhat + orow * p.ldy + n0 + ch * 32, v);
                        float ga[32], be[32];
                        lds32(scol + NT + ch * 32, ga);
                        lds32(scol + 2 * NT + ch * 32, be);
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = act_fwd_fast(ga[j] * v[j] + be[j], act);
                    }
                    if (tt_i >= 0) add_row32(p.embT + (size_t)tt_i * p.ldeT + p.eoff + n0 + ch * 32, v);
                    if (er_i >= 0) add_row32(p.embR + (size_t)er_i * p.ldeR + p.eoff + n0 + ch * 32, v);
                    if (p.ident) add_row32(p.ident + orow * p.ldi + n0 + ch * 32, v);
                    if (p.addend) add_row32(p.addend + orow * p.ldad + n0 + ch * 32, v);
                    if (valid) {
                        const size_t o = orow * p.ldo + n0 + ch * 32;
                        if (p.out) st_row32(p.out + o, v);
                        if (p.out_hi) st_planes32(p.out_hi + o, p.out_lo + o, v);
                    }
                }
            } else {
                // ---- backward with GroupNorm/activation backward.
                // pass 0 writes the raw gradient, overwrites the accumulator with g_yhat and reduces the
                // slab means; pass 1 re-reads yhat (L2 resident) and emits g_y.
#pragma unroll 1
                for (int ch = 0; ch < NCH; ++ch) {
                    tmem_ld32(tacc + ch * 32, v);
                    if (p.ident) add_row32(p.ident + orow * p.ldi + n0 + ch * 32, v);
                    if (p.addend) add_row32(p.addend + orow * p.ldad + n0 + ch * 32, v);
                    if (valid) {
                        const size_t o = orow * p.ldo + n0 + ch * 32;
                        if (p.out) st_row32(p.out + o, v);
                        if (p.out_hi) st_planes32(p.out_hi + o, p.out_lo + o, v);
                    }
                    float y[32];
                    ld_row32(p.yhat + orow * p.ldy + n0 + ch * 32, y);
                    float q1[GPC], q2[GPC];
#pragma unroll
                    for (int g = 0; g < GPC; ++g) q1[g] = q2[g] = 0.f;
                    float ga[32], be[32];
                    lds32(scol + NT + ch * 32, ga);
                    lds32(scol + 2 * NT + ch * 32, be);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float gam = ga[j];
                        const float gyh = valid ? v[j] * act_bwd_fast(gam * y[j] + be[j], act) * gam : 0.f;
                        v[j] = gyh;
                        q1[GPC > 1 ? j / (32 / GPC) : 0] += gyh;
                        q2[GPC > 1 ? j / (32 / GPC) : 0] += valid ? gyh * y[j] : 0.f;
                    }
                    tmem_st32(tacc + ch * 32, v);
                    const int g0 = (ch * 32) / (GS ? GS : 1);
#pragma unroll
                    for (int g = 0; g < GPC; ++g) {
                        mypart[2 * (g0 + g)] += q1[g];
                        mypart[2 * (g0 + g) + 1] += q2[g];
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                epi_bar();
                if (etid < 2 * S * G) {
                    const int t2 = etid / (S * G), rem = etid - t2 * S * G;
                    const int s = rem / G, g = rem - s * G;
                    float a = 0.f, b = 0.f;
                    for (int hh = 0; hh < Hl; ++hh) {
                        a += spart[((t2 * 128 + hh * S + s) * G + g) * 2];
                        b += spart[((t2 * 128 + hh * S + s) * G + g) * 2 + 1];
                    }
                    const int idx = t2 * 128 + s * G + g;
                    sstat[idx] = a * inv_cnt;
                    sstat[256 + idx] = b * inv_cnt;
                    const int rr = rbase + t2 * S + s;
                    sstat[512 + idx] = (rr < p.R) ? __ldg(p.rstd + (size_t)rr * 8 + n0 / (GS ? GS : 1) + g) : 0.f;
                }
                epi_bar();
#pragma unroll 1
                for (int ch = 0; ch < NCH; ++ch) {
                    float y[32];
                    tmem_ld32(tacc + ch * 32, v);
                    ld_row32(p.yhat + orow * p.ldy + n0 + ch * 32, y);
                    float m1[GPC], m2[GPC], rsd[GPC];
#pragma unroll
                    for (int g = 0; g < GPC; ++g) {
                        const int sl = slab0 + (ch * 32) / (GS ? GS : 1) + g;
                        m1[g] = sstat[sl]; m2[g] = sstat[256 + sl]; rsd[g] = sstat[512 + sl];
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int g = GPC > 1 ? j / (32 / GPC) : 0;
                        v[j] = rsd[g] * (v[j] - m1[g] - y[j] * m2[g]);
                    }
                    if (valid) {
                        const size_t o = orow * p.ldg + n0 + ch * 32;
                        if (p.gy) st_row32(p.gy + o, v);
                        if (p.gy_hi) st_planes32(p.gy_hi + o, p.gy_lo + o, v);
                    }
                }
            }
            // release the accumulator stage
            tc_fence_before();
            epi_bar();
            if (etid == 0) mbar_arrive(&tempty[as]);
            if (++as == NACC) { as = 0; aph ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode2() {
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
std::map<MapKey, CUtensorMap> g_maps2;

// activation plane [R][H][ld] BF16 -> 3-D map ordered (C, R, H), box (32, S, H+4): rows land
// position-major in shared memory with the conv halo zero filled
int act_map2(CUtensorMap* out, const __nv_bfloat16* base, int C, int H, int R, int ld, int S) {
    MapKey key(base, C, H, R, ld, S, 30);
    auto it = g_maps2.find(key);
    if (it != g_maps2.end()) { *out = it->second; return 0; }
    EncodeFn enc = get_encode2();
    PMP_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)R, (cuuint64_t)H};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2 * (cuuint64_t)H, (cuuint64_t)ld * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK2, (cuuint32_t)S, (cuuint32_t)(H + 4)};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PMP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(tc2 activation C=%d H=%d R=%d ld=%d S=%d) failed: %d", C, H, R, ld, S, (int)r);
    g_maps2[key] = *out;
    return 0;
}
int w_map2(CUtensorMap* out, const __nv_bfloat16* base, int C, int rows, int NT) {
    MapKey key(base, C, rows, NT, 0, 0, 20);
    auto it = g_maps2.find(key);
    if (it != g_maps2.end()) { *out = it->second; return 0; }
    EncodeFn enc = get_encode2();
    PMP_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK2, (cuuint32_t)NT};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PMP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(tc2 weight C=%d rows=%d NT=%d) failed: %d", C, rows, NT, (int)r);
    g_maps2[key] = *out;
    return 0;
}

int g_sms2 = 0;

template <int NT, int GS>
int launch2_t(Tc2P& P, cudaStream_t st) {
    using C = Cfg2<NT, GS>;
    static bool configured = false;
    if (!configured) {
        PMP_CUDA_OK(cudaFuncSetAttribute(conv_tc2_kernel<NT, GS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured = true;
    }
    if (!g_sms2) {
        int dev = 0;
        PMP_CUDA_OK(cudaGetDevice(&dev));
        PMP_CUDA_OK(cudaDeviceGetAttribute(&g_sms2, cudaDevAttrMultiProcessorCount, dev));
    }
    const int grid = P.ntiles < g_sms2 ? P.ntiles : g_sms2;
    conv_tc2_kernel<NT, GS><<<grid, T2_THREADS, C::SMEM_BYTES, st>>>(P);
    PMP_LAUNCH_OK();
    return 0;
}

int pick_nt2(const ConvP& p) {
    if (p.gn_fwd || p.gn_bwd) {
        const int gs = p.N / 8;
        if (gs == 8) return 64;
        if (gs == 32 || gs == 64) return 128;
        if (gs == 96) return 192;
        return 0;
    }
    if (p.N % 128 == 0) return 128;
    if (p.N % 64 == 0) return 64;
    return 0;
}
}  // namespace

bool conv_tc2_eligible(const ConvP& p) {
    if (!(p.nphase == 1 && p.istride == 1 && p.ostride == 1)) return false;
    if (!p.A1_hi || !p.A1_lo || !p.W1_hi || !p.W1_lo) return false;
    if (p.C1 % BK2 != 0 || (p.lda1 % 8) != 0) return false;
    const bool acc2 = (p.A2 != nullptr) && (p.bias2 != nullptr || p.gn_fwd);
    if (acc2) return false;                                    // forward residual conv: conv_tc.cu
    if (p.A2 && (!p.A2_hi || !p.A2_lo || !p.W2_hi || !p.W2_lo || p.C2 % BK2 != 0 || (p.lda2 % 8) != 0)) return false;
    if (p.energy) return false;
    if (p.HJ < 8 || p.HJ > 124) return false;
    const int S = 128 / p.HJ;
    if (4 * S + 128 > A_ROWS || (p.HJ + 4) * S > A_ROWS) return false;
    for (int t = 0; t < p.ntaps[0]; ++t)
        if (p.tap_shift[0][t] < -2 || p.tap_shift[0][t] > 2) return false;
    const int nt = pick_nt2(p);
    if (!nt || p.N % nt != 0) return false;
    if ((p.gn_fwd || p.gn_bwd) && S * (nt / (p.N / 8)) > 128) return false;
    return true;
}

int launch_conv_tc2(const ConvP& p, cudaStream_t st) {
    if (!conv_tc2_eligible(p)) return 1;
    Tc2P P;
    memset(&P, 0, sizeof P);
    P.e = p;
    const int NT = pick_nt2(p);
    const int S = 128 / p.HJ;
    P.e.S = S;
    P.tilesM = (p.R + 2 * S - 1) / (2 * S);
    P.tilesN = p.N / NT;
    P.ntiles = P.tilesM * P.tilesN;
    P.nch1 = p.C1 / BK2;
    P.nch2 = p.A2 ? p.C2 / BK2 : 0;
#ifdef PMP_DIAG
    {
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("PMP_TC2_DBG"); dbg = e ? atoi(e) : 0; }
        P.dbg = dbg;
    }
#endif
    int rc;
    if ((rc = act_map2(&P.mA1h, p.A1_hi, p.C1, p.Hin, p.R, p.lda1, S))) return rc;
    if ((rc = act_map2(&P.mA1l, p.A1_lo, p.C1, p.Hin, p.R, p.lda1, S))) return rc;
    int maxk = 0;
    for (int t = 0; t < p.ntaps[0]; ++t) maxk = p.tap_k[0][t] > maxk ? p.tap_k[0][t] : maxk;
    if ((rc = w_map2(&P.mW1h, p.W1_hi, p.C1, (maxk + 1) * p.N, NT))) return rc;
    if ((rc = w_map2(&P.mW1l, p.W1_lo, p.C1, (maxk + 1) * p.N, NT))) return rc;
    if (p.A2) {
        if ((rc = act_map2(&P.mA2h, p.A2_hi, p.C2, p.Hin, p.R, p.lda2, S))) return rc;
        if ((rc = act_map2(&P.mA2l, p.A2_lo, p.C2, p.Hin, p.R, p.lda2, S))) return rc;
        if ((rc = w_map2(&P.mW2h, p.W2_hi, p.C2, p.N, NT))) return rc;
        if ((rc = w_map2(&P.mW2l, p.W2_lo, p.C2, p.N, NT))) return rc;
    }
    const int gs = (p.gn_fwd || p.gn_bwd) ? p.N / 8 : 0;
    if (NT == 64) return gs == 8 ? launch2_t<64, 8>(P, st) : launch2_t<64, 0>(P, st);
    if (NT == 128) {
        if (gs == 32) return launch2_t<128, 32>(P, st);
        if (gs == 64) return launch2_t<128, 64>(P, st);
        return launch2_t<128, 0>(P, st);
    }
    if (NT == 192) return launch2_t<192, 96>(P, st);
    return 1;
}

}  // namespace pmp
