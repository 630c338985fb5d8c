// PTX wrappers shared by the tcgen05 conv kernels (conv_tc.cu, conv_tc2.cu): mbarrier, TMA, tcgen05
// MMA / commit / TMEM load-store, and per-row vector IO helpers of the epilogues.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace pmp {
namespace tc {

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, BF16 inputs, FP32 accumulate, M=128
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// ---- CTA-pair (cta_group::2) variants: one MMA spans the two SMs of a TPC (M = 256, each CTA holds
// its own 128 A rows and half of the B rows), so every SM reads half of B from shared memory.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(rank)
        : "memory");
}
// TMA loads issued by either CTA of a pair; the transaction bytes are credited to the LEADER's barrier
// (clearing the peer bit of the shared::cluster address selects the even CTA of the pair)
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
        : "memory");
}
// completion of all prior MMAs of this thread arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void tc_commit2(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
// D[tmem of both CTAs] (+)= A * B^T, M = 256 (128 rows per CTA), issued by the leader CTA only
__device__ __forceinline__ void tc_mma2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// ---- TMA-store epilogue: results are staged in shared memory in the layout of a swizzled TMA box
// ([rows][32 columns]; fp32 rows of 128 B with SWIZZLE_128B, BF16 rows of 64 B with SWIZZLE_64B) and
// written out by bulk tensor stores, so the LSU only sees conflict-free st.shared and global memory only
// sees whole 128-byte lines.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk stores committed by this thread have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy shared-memory writes of this thread become visible to the async proxy (TMA)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void half_bar(int hf) { asm volatile("bar.sync %0, 128;" ::"r"(2 + hf) : "memory"); }
// row r (0..127) of a [128][32] fp32 staging tile, SWIZZLE_128B: 16-byte chunk c sits at c ^ (r & 7)
__device__ __forceinline__ void stage_f32(uint8_t* tile, int r, const float (&v)[32]) {
    uint8_t* row = tile + r * 128;
    const int x = r & 7;
#pragma unroll
    for (int c = 0; c < 8; ++c)
        *reinterpret_cast<float4*>(row + ((c ^ x) << 4)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
// row r of two [128][32] BF16 staging tiles (hi / lo planes), SWIZZLE_64B: chunk c sits at c ^ ((r >> 1) & 3)
__device__ __forceinline__ void stage_planes(uint8_t* thi, uint8_t* tlo, int r, const float (&v)[32]) {
    uint8_t* rh = thi + r * 64;
    uint8_t* rl = tlo + r * 64;
    const int x = (r >> 1) & 3;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) split_bf16x2(v[8 * c + 2 * k], v[8 * c + 2 * k + 1], h[k], l[k]);
        *reinterpret_cast<uint4*>(rh + ((c ^ x) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(rl + ((c ^ x) << 4)) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}
// K-major operand tile descriptors.  BKC = 64: rows of 64 BF16 = 128 bytes, SWIZZLE_128B, 8-row groups 1024 B
// apart (SBO = 64 x 16 B); BKC = 32: rows of 64 bytes, SWIZZLE_64B, 8-row groups 512 B apart.  LBO = 1 (unused).
template <int BKC = 64>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    constexpr uint64_t sbo = BKC == 64 ? 64ull : 32ull;
    constexpr uint64_t layout = BKC == 64 ? 2ull : 4ull;
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])),
        "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])),
        "r"(__float_as_uint(v[23])), "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
        "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])),
        "r"(__float_as_uint(v[31]))
        : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// 32 consecutive floats of a shared-memory table (same address for every lane: broadcast)
__device__ __forceinline__ void lds32(const float* p, float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 t = reinterpret_cast<const float4*>(p)[i];
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}
__device__ __forceinline__ void add_lds32(const float* p, float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 t = reinterpret_cast<const float4*>(p)[i];
        v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
    }
}
// per-row 128-byte segment helpers (every thread handles its own row: 8 x 16-byte vectors)
__device__ __forceinline__ void ld_row32(const float* __restrict__ p, float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}
__device__ __forceinline__ void add_row32(const float* __restrict__ p, float (&v)[32]) {
    float4 t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = __ldg(reinterpret_cast<const float4*>(p) + i);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        v[4 * i] += t[i].x; v[4 * i + 1] += t[i].y; v[4 * i + 2] += t[i].z; v[4 * i + 3] += t[i].w;
    }
}
// v[j] += hi[j] + lo[j] for 32 columns of a BF16x3 plane pair (4 + 4 16-byte vectors per row)
__device__ __forceinline__ void add_planes32(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                             float (&v)[32]) {
    uint4 h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = __ldg(reinterpret_cast<const uint4*>(hi) + i);
        l[i] = __ldg(reinterpret_cast<const uint4*>(lo) + i);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t hw[4] = {h[i].x, h[i].y, h[i].z, h[i].w}, lw[4] = {l[i].x, l[i].y, l[i].z, l[i].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[8 * i + 2 * k] += __uint_as_float(hw[k] << 16) + __uint_as_float(lw[k] << 16);
            v[8 * i + 2 * k + 1] += __uint_as_float(hw[k] & 0xffff0000u) + __uint_as_float(lw[k] & 0xffff0000u);
        }
    }
}
__device__ __forceinline__ void st_row32(float* __restrict__ p, const float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
        reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void st_planes32(__nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                            const float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) split_bf16x2(v[8 * i + 2 * k], v[8 * i + 2 * k + 1], h[k], l[k]);
        reinterpret_cast<uint4*>(hi)[i] = make_uint4(h[0], h[1], h[2], h[3]);
        reinterpret_cast<uint4*>(lo)[i] = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

// ---- activation blocks of the epilogues.  SiLU (every shipped config) runs branch-free on
// ex2.approx / rcp.approx (2 MUFU + 3 FP32 ops per element, ~3e-7 relative error, far below the BF16x3
// product error); the 32 elements are independent straight-line code so the MUFU latency overlaps.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float silu_fast(float z) { return z * rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * z)); }
__device__ __forceinline__ float silu_bwd_fast(float z) {
    const float s = rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * z));
    return s * fmaf(z, 1.0f - s, 1.0f);
}
// v[j] = act(gamma[j] * v[j] + beta[j]) for 32 columns; gamma / beta are shared-memory tables
__device__ __forceinline__ void act_affine32(const float* gam, const float* bet, float (&v)[32], int act) {
    if (act == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 g4 = reinterpret_cast<const float4*>(gam)[q], b4 = reinterpret_cast<const float4*>(bet)[q];
            v[4 * q] = silu_fast(fmaf(g4.x, v[4 * q], b4.x));
            v[4 * q + 1] = silu_fast(fmaf(g4.y, v[4 * q + 1], b4.y));
            v[4 * q + 2] = silu_fast(fmaf(g4.z, v[4 * q + 2], b4.z));
            v[4 * q + 3] = silu_fast(fmaf(g4.w, v[4 * q + 3], b4.w));
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = act_fwd(gam[j] * v[j] + bet[j], act);
    }
}
// g[j] = g[j] * act'(gamma[j] * y[j] + beta[j]) * gamma[j]
__device__ __forceinline__ void act_affine_bwd32(const float* gam, const float* bet, const float (&y)[32], float (&g)[32],
                                                 int act) {
    if (act == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 g4 = reinterpret_cast<const float4*>(gam)[q], b4 = reinterpret_cast<const float4*>(bet)[q];
            g[4 * q] *= silu_bwd_fast(fmaf(g4.x, y[4 * q], b4.x)) * g4.x;
            g[4 * q + 1] *= silu_bwd_fast(fmaf(g4.y, y[4 * q + 1], b4.y)) * g4.y;
            g[4 * q + 2] *= silu_bwd_fast(fmaf(g4.z, y[4 * q + 2], b4.z)) * g4.z;
            g[4 * q + 3] *= silu_bwd_fast(fmaf(g4.w, y[4 * q + 3], b4.w)) * g4.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) g[j] *= act_bwd(gam[j] * y[j] + bet[j], act) * gam[j];
    }
}


// ------------------------------------------------------------------ chunk-width generic forms (CW = 32 or 16 columns
// per thread and step).  The 16-column forms exist for the 16-warp epilogue: half the live registers per thread.
template <int CW>
__device__ __forceinline__ void tmem_ldw(uint32_t taddr, float (&v)[CW]) {
    static_assert(CW == 16 || CW == 32, "chunk width");
    uint32_t r[CW];
    if constexpr (CW == 32) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr)
            : "memory");
    } else {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr)
            : "memory");
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] = __uint_as_float(r[i]);
}
template <int CW>
__device__ __forceinline__ void tmem_stw(uint32_t taddr, const float (&v)[CW]) {
    if constexpr (CW == 32) {
        tmem_st32(taddr, v);
    } else {
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
            "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
            ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
            "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
            "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
            "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
            "r"(__float_as_uint(v[15]))
            : "memory");
    }
}
template <int CW>
__device__ __forceinline__ void add_ldsw(const float* p, float (&v)[CW]) {
#pragma unroll
    for (int i = 0; i < CW / 4; ++i) {
        const float4 t = reinterpret_cast<const float4*>(p)[i];
        v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
    }
}
template <int CW>
__device__ __forceinline__ void ld_roww(const float* __restrict__ p, float (&v)[CW]) {
#pragma unroll
    for (int i = 0; i < CW / 4; ++i) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}
template <int CW>
__device__ __forceinline__ void add_roww(const float* __restrict__ p, float (&v)[CW]) {
    float4 t[CW / 4];
#pragma unroll
    for (int i = 0; i < CW / 4; ++i) t[i] = __ldg(reinterpret_cast<const float4*>(p) + i);
#pragma unroll
    for (int i = 0; i < CW / 4; ++i) {
        v[4 * i] += t[i].x; v[4 * i + 1] += t[i].y; v[4 * i + 2] += t[i].z; v[4 * i + 3] += t[i].w;
    }
}
template <int CW>
__device__ __forceinline__ void st_roww(float* __restrict__ p, const float (&v)[CW]) {
#pragma unroll
    for (int i = 0; i < CW / 4; ++i)
        reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
template <int CW>
__device__ __forceinline__ void add_planesw(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                            float (&v)[CW]) {
    uint4 h[CW / 8], l[CW / 8];
#pragma unroll
    for (int i = 0; i < CW / 8; ++i) {
        h[i] = __ldg(reinterpret_cast<const uint4*>(hi) + i);
        l[i] = __ldg(reinterpret_cast<const uint4*>(lo) + i);
    }
#pragma unroll
    for (int i = 0; i < CW / 8; ++i) {
        const uint32_t hw[4] = {h[i].x, h[i].y, h[i].z, h[i].w}, lw[4] = {l[i].x, l[i].y, l[i].z, l[i].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[8 * i + 2 * k] += __uint_as_float(hw[k] << 16) + __uint_as_float(lw[k] << 16);
            v[8 * i + 2 * k + 1] += __uint_as_float(hw[k] & 0xffff0000u) + __uint_as_float(lw[k] & 0xffff0000u);
        }
    }
}
template <int CW>
__device__ __forceinline__ void st_planesw(__nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, const float (&v)[CW]) {
#pragma unroll
    for (int i = 0; i < CW / 8; ++i) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) split_bf16x2(v[8 * i + 2 * k], v[8 * i + 2 * k + 1], h[k], l[k]);
        reinterpret_cast<uint4*>(hi)[i] = make_uint4(h[0], h[1], h[2], h[3]);
        reinterpret_cast<uint4*>(lo)[i] = make_uint4(l[0], l[1], l[2], l[3]);
    }
}
template <int CW>
__device__ __forceinline__ void act_affinew(const float* gam, const float* bet, float (&v)[CW], int act) {
    if (act == 0) {
#pragma unroll
        for (int q = 0; q < CW / 4; ++q) {
            const float4 g4 = reinterpret_cast<const float4*>(gam)[q], b4 = reinterpret_cast<const float4*>(bet)[q];
            v[4 * q] = silu_fast(fmaf(g4.x, v[4 * q], b4.x));
            v[4 * q + 1] = silu_fast(fmaf(g4.y, v[4 * q + 1], b4.y));
            v[4 * q + 2] = silu_fast(fmaf(g4.z, v[4 * q + 2], b4.z));
            v[4 * q + 3] = silu_fast(fmaf(g4.w, v[4 * q + 3], b4.w));
        }
    } else {
#pragma unroll
        for (int j = 0; j < CW; ++j) v[j] = act_fwd(gam[j] * v[j] + bet[j], act);
    }
}
template <int CW>
__device__ __forceinline__ void act_affine_bwdw(const float* gam, const float* bet, const float (&y)[CW], float (&g)[CW], int act) {
    if (act == 0) {
#pragma unroll
        for (int q = 0; q < CW / 4; ++q) {
            const float4 g4 = reinterpret_cast<const float4*>(gam)[q], b4 = reinterpret_cast<const float4*>(bet)[q];
            g[4 * q] *= silu_bwd_fast(fmaf(g4.x, y[4 * q], b4.x)) * g4.x;
            g[4 * q + 1] *= silu_bwd_fast(fmaf(g4.y, y[4 * q + 1], b4.y)) * g4.y;
            g[4 * q + 2] *= silu_bwd_fast(fmaf(g4.z, y[4 * q + 2], b4.z)) * g4.z;
            g[4 * q + 3] *= silu_bwd_fast(fmaf(g4.w, y[4 * q + 3], b4.w)) * g4.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < CW; ++j) g[j] *= act_bwd(gam[j] * y[j] + bet[j], act) * gam[j];
    }
}
// staging rows of CW columns: fp32 rows of CW*4 bytes (128: SWIZZLE_128B, 64: SWIZZLE_64B), BF16 rows of CW*2 bytes
// (64: SWIZZLE_64B, 32: SWIZZLE_32B); the XOR term is the one the TMA store applies for that swizzle mode
template <int RB>
__device__ __forceinline__ int swz_xor(int r) {
    return RB == 128 ? (r & 7) : (RB == 64 ? ((r >> 1) & 3) : ((r >> 2) & 1));
}
template <int CW>
__device__ __forceinline__ void stage_f32w(uint8_t* tile, int r, const float (&v)[CW]) {
    constexpr int RB = CW * 4;
    uint8_t* row = tile + r * RB;
    const int x = swz_xor<RB>(r);
#pragma unroll
    for (int c = 0; c < RB / 16; ++c)
        *reinterpret_cast<float4*>(row + ((c ^ x) << 4)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
template <int CW>
__device__ __forceinline__ void stage_planesw(uint8_t* thi, uint8_t* tlo, int r, const float (&v)[CW]) {
    constexpr int RB = CW * 2;
    uint8_t* rh = thi + r * RB;
    uint8_t* rl = tlo + r * RB;
    const int x = swz_xor<RB>(r);
#pragma unroll
    for (int c = 0; c < RB / 16; ++c) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) split_bf16x2(v[8 * c + 2 * k], v[8 * c + 2 * k + 1], h[k], l[k]);
        *reinterpret_cast<uint4*>(rh + ((c ^ x) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(rl + ((c ^ x) << 4)) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}
// planes that were split earlier (packed BF16 pairs kept in registers) into the two plane tiles
template <int CW>
__device__ __forceinline__ void stage_packedw(uint8_t* thi, uint8_t* tlo, int r, const uint32_t (&h)[CW / 2], const uint32_t (&l)[CW / 2]) {
    constexpr int RB = CW * 2;
    uint8_t* rh = thi + r * RB;
    uint8_t* rl = tlo + r * RB;
    const int x = swz_xor<RB>(r);
#pragma unroll
    for (int c = 0; c < RB / 16; ++c) {
        *reinterpret_cast<uint4*>(rh + ((c ^ x) << 4)) = make_uint4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
        *reinterpret_cast<uint4*>(rl + ((c ^ x) << 4)) = make_uint4(l[4 * c], l[4 * c + 1], l[4 * c + 2], l[4 * c + 3]);
    }
}
// the reverse direction: a row of a staging tile that a TMA LOAD filled (same swizzled layouts)
template <int CW>
__device__ __forceinline__ void unstage_f32w(const uint8_t* tile, int r, float (&v)[CW]) {
    constexpr int RB = CW * 4;
    const uint8_t* row = tile + r * RB;
    const int x = swz_xor<RB>(r);
#pragma unroll
    for (int c = 0; c < RB / 16; ++c) {
        const float4 t = *reinterpret_cast<const float4*>(row + ((c ^ x) << 4));
        v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
    }
}
// v[j] += hi[j] + lo[j] from the two BF16 plane tiles
template <int CW>
__device__ __forceinline__ void add_unstage_planesw(const uint8_t* thi, const uint8_t* tlo, int r, float (&v)[CW]) {
    constexpr int RB = CW * 2;
    const uint8_t* rh = thi + r * RB;
    const uint8_t* rl = tlo + r * RB;
    const int x = swz_xor<RB>(r);
#pragma unroll
    for (int c = 0; c < RB / 16; ++c) {
        const uint4 h = *reinterpret_cast<const uint4*>(rh + ((c ^ x) << 4));
        const uint4 l = *reinterpret_cast<const uint4*>(rl + ((c ^ x) << 4));
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[8 * c + 2 * k] += __uint_as_float(hw[k] << 16) + __uint_as_float(lw[k] << 16);
            v[8 * c + 2 * k + 1] += __uint_as_float(hw[k] & 0xffff0000u) + __uint_as_float(lw[k] & 0xffff0000u);
        }
    }
}
template <int NTHREADS>
__device__ __forceinline__ void epi_barw() { asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory"); }

template <int GS>
__device__ __forceinline__ float group_reduce(float v) {
    // sum over the lanes of one GroupNorm group inside a 32-column chunk
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    if (GS >= 16) v += __shfl_xor_sync(0xffffffffu, v, 8);
    if (GS >= 32) v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
}


}  // namespace tc
}  // namespace pmp
