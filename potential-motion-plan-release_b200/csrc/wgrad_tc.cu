// Weight-gradient GEMMs of the training step on the tcgen05 tensor cores (train.cu has the fp32 CUDA-core version,
// which stays the exact-fp32 path and the fallback for strided / narrow convolutions).
//
//   dW[co][ci][k] = sum_terms sum_{r,h} O[r,h,co] * I[r,h+k-pad,ci]          (stride-1 conv, KS taps, pad = KS/2)
//
// GEMM view: M = Co, N = Ci (per tap), K = terms * R * (H + 4).  The activations are channels-last [R][H][C], i.e.
// MN-major for this product, so they are first TRANSPOSED into K-major BF16x3 planes
//   OT[co][j], IT[ci][j],   j = ((term * G + r / 8) * (H + 4) + 2 + h) * 8 + r % 8,   G = ceil(R / 8),
// zero in the 2 + 2 padding positions of every sample (transpose_planes_kernel).  Samples are interleaved in groups of 8 so
// that one horizon step is 8 columns = 16 bytes: tap k is then the same NT GEMM with the B operand started 8 (k - pad)
// columns later, which keeps the innermost TMA coordinate 16-byte aligned (an element-granular start faults), and the
// shift never leaves the sample's own padding, where O is zero.  Both operands are plain K-major tiles (rows of 64 BF16
// = 128 bytes, SWIZZLE_128B) filled by 2-D TMA loads, exactly like the weight tiles of conv_tc_kernel; every K block
// issues hi*hi + hi*lo + lo*hi into one FP32 TMEM accumulator (BF16x3, ~2^-16 relative product error).
// One CTA = one (128 x NT) output tile of one tap over a slice of K (split-K); the epilogue adds its tile into a
// tap-major fp32 workspace Wt[k][co][ci] with vector reductions (red.global.add.v4.f32: 128 contiguous bytes per
// thread), and wgrad_scatter_kernel adds the workspace into the flat gradient in the reference layout [co][ci][k].
#include <algorithm>
#include <map>
#include <tuple>

#include "tc_common.cuh"

namespace pmp {
namespace {
using namespace tc;

// columns of the transposed planes that `nterms` terms occupy, rounded up to whole K blocks
static size_t wg_cols(int nterms, int R, int H) { return ((size_t)nterms * ((R + 7) / 8) * (H + 4) * 8 + 63) / 64 * 64; }

constexpr int WG_THREADS = 192;      // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int WG_BK = 64;

struct WgP {
    CUtensorMap mAh, mAl, mBh, mBl;   // OT planes (K, Co), IT planes (K, Ci)
    int Co, Ci, KS, pad;
    int nkb;                          // K blocks of 64 in total
    int kb_per_split, nsplit;
    int mtiles, ntiles;
    float* Wt;                        // [KS][Co][Ci]
    int dbg;                          // test hook: 1 no tap shift, 2 no reductions, 4 no MMA, 8 no TMA
};

template <int NT>
struct WgCfg {
    static constexpr int A_BYTES = 128 * 128;          // one plane tile: 128 rows x 128 bytes
    static constexpr int B_BYTES = NT * 128;
    static constexpr int STAGE = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int NSTAGE = NT == 256 ? 2 : (NT == 128 ? 3 : 4);
    static constexpr int SMEM = NSTAGE * STAGE + 1024 + 256;
};

template <int NT>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgP P) {
    using C = WgCfg<NT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::NSTAGE * C::STAGE);
    uint64_t* full = bars;
    uint64_t* empty = bars + C::NSTAGE;
    uint64_t* tfull = bars + 2 * C::NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::NSTAGE + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // work item: blockIdx.x -> (split, tap, n tile, m tile)
    int w = blockIdx.x;
    const int mt = w % P.mtiles; w /= P.mtiles;
    const int nt = w % P.ntiles; w /= P.ntiles;
    const int tap = w % P.KS;
    const int split = w / P.KS;
    const int kb0 = split * P.kb_per_split;
    const int kb1 = min(P.nkb, kb0 + P.kb_per_split);
    const int m0 = mt * 128, n0 = nt * NT;
    const int shift = (P.dbg & 1) ? 0 : 8 * (tap - P.pad);

    if (threadIdx.x == 0) {
        prefetch_tmap(&P.mAh); prefetch_tmap(&P.mAl); prefetch_tmap(&P.mBh); prefetch_tmap(&P.mBl);
        for (int i = 0; i < C::NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(NT < 32 ? 32 : NT)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t* st = smem + s * C::STAGE;
                const int k0 = kb * WG_BK;
                if (P.dbg & 8) {
                    mbar_arrive(&full[s]);
                } else {
                    mbar_expect_tx(&full[s], C::STAGE);
                    tma_load_2d(st, &P.mAh, &full[s], k0, m0);
                    tma_load_2d(st + C::A_BYTES, &P.mAl, &full[s], k0, m0);
                    tma_load_2d(st + 2 * C::A_BYTES, &P.mBh, &full[s], k0 + shift, n0);
                    tma_load_2d(st + 2 * C::A_BYTES + C::B_BYTES, &P.mBl, &full[s], k0 + shift, n0);
                }
                if (++s == C::NSTAGE) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D = F32, A = B = BF16, both K-major, N = NT, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
            int s = 0;
            uint32_t ph = 0, accum = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * C::STAGE);
                const uint64_t ah = make_desc<64>(sa), al = make_desc<64>(sa + C::A_BYTES);
                const uint64_t bh = make_desc<64>(sa + 2 * C::A_BYTES), bl = make_desc<64>(sa + 2 * C::A_BYTES + C::B_BYTES);
#pragma unroll
                for (int k = 0; k < WG_BK / 16; ++k) {
                    if (P.dbg & 4) break;
                    const uint64_t ko = (uint64_t)(k * 2);      // 32 bytes per K = 16 step, in 16-byte units
                    tc_mma(tmem_base, ah + ko, bh + ko, idesc, accum);
                    tc_mma(tmem_base, ah + ko, bl + ko, idesc, 1u);
                    tc_mma(tmem_base, al + ko, bh + ko, idesc, 1u);
                    accum = 1u;
                }
                tc_commit(&empty[s]);
                if (++s == C::NSTAGE) { s = 0; ph ^= 1; }
            }
            tc_commit(tfull);
        }
    } else {
        // epilogue: thread = one accumulator row (output channel co); 32-column chunks -> 8 vector reductions of 16 bytes
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int co = m0 + q * 32 + lane;
        mbar_wait(tfull, 0);
        tc_fence_after();
        if (kb1 > kb0) {
            float* dst = P.Wt + ((size_t)tap * P.Co + co) * P.Ci + n0;
#pragma unroll 1
            for (int ch = 0; ch < NT / 32; ++ch) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ch * 32, v);
                if (co < P.Co && !(P.dbg & 2)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + ch * 32 + j), "f"(v[j]), "f"(v[j + 1]),
                                     "f"(v[j + 2]), "f"(v[j + 3])
                                     : "memory");
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(NT < 32 ? 32 : NT) : "memory");
    }
}

// fp32 channels-last [R][H][ld] -> K-major BF16x3 planes: dst[c][col0 + ((g*(H+4) + 2 + h)*8 + r8] = split(src[8g + r8][h][c])
// One launch transposes all operands of a weight gradient (up to 4 jobs: O and I of both terms, blockIdx.z); one block =
// (group of 8 samples, 16-channel tile): every channel row of the group is one contiguous run of (H+4)*8 columns, written
// whole -- the 2 + 2 padding positions and the missing samples of a ragged last group as zeros -- and the blocks of the last
// group of the last term also zero the columns up to `zero_end` (K rounding + tap-shift slack), so the planes need no memset.
struct TJob {
    const float* src; int ld, C;
    __nv_bfloat16 *hi, *lo;
    size_t col0;
    int last;                 // 1: this job's last group zeroes [col_end, zero_end)
};
struct TJobs { TJob j[4]; int n; };

__global__ void __launch_bounds__(256) transpose_planes_kernel(const __grid_constant__ TJobs jobs, int R, int H, size_t Kld,
                                                               size_t zero_end, int vec) {
    extern __shared__ float tsm[];                 // [8][H][17]
    const TJob& J = jobs.j[blockIdx.z];
    const int g = blockIdx.x, c0 = blockIdx.y * 16, G = gridDim.x;
    if (c0 >= J.C) return;
    const int nr = min(8, R - g * 8);
    if (vec) {
        // 16-byte loads: 4 lanes cover the 16 channels of one (sample, position) row
        for (int e = threadIdx.x; e < nr * H * 4; e += 256) {
            const int q4 = e & 3, rh = e >> 2;     // rh = r8 * H + h
            const float4 v = __ldg(reinterpret_cast<const float4*>(J.src + ((size_t)g * 8 * H + rh) * J.ld + c0 + 4 * q4));
            float* d = tsm + rh * 17 + 4 * q4;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    } else {
        for (int e = threadIdx.x; e < nr * H * 16; e += 256) {
            const int c = e & 15, rh = e >> 4;
            tsm[rh * 17 + c] = (c0 + c < J.C) ? J.src[((size_t)g * 8 * H + rh) * J.ld + c0 + c] : 0.f;
        }
    }
    __syncthreads();
    const int HP = H + 4;
    const size_t base = J.col0 + (size_t)g * HP * 8;
    // one thread = the 8 interleaved samples of one (channel, position): two 16-byte stores (hi, lo); consecutive threads
    // take consecutive positions, so a warp writes 512 contiguous bytes per plane
    for (int e = threadIdx.x; e < 16 * HP; e += 256) {
        const int hp = e % HP, c = e / HP;
        if (c0 + c >= J.C) continue;
        const int h = hp - 2;
        uint32_t ph[4] = {0u, 0u, 0u, 0u}, pl[4] = {0u, 0u, 0u, 0u};
        if (h >= 0 && h < H) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float a = (2 * k < nr) ? tsm[((2 * k) * H + h) * 17 + c] : 0.f;
                const float b = (2 * k + 1 < nr) ? tsm[((2 * k + 1) * H + h) * 17 + c] : 0.f;
                split_bf16x2(a, b, ph[k], pl[k]);
            }
        }
        const size_t o = (size_t)(c0 + c) * Kld + base + (size_t)hp * 8;
        *reinterpret_cast<uint4*>(J.hi + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        *reinterpret_cast<uint4*>(J.lo + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
    if (J.last && g == G - 1) {
        const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
        const size_t end = J.col0 + (size_t)G * HP * 8;
        const int nz = (int)(zero_end - end);
        for (int e = threadIdx.x; e < 16 * nz; e += 256) {
            const int c = e / nz, k = e - c * nz;
            if (c0 + c >= J.C) continue;
            const size_t o = (size_t)(c0 + c) * Kld + end + k;
            J.hi[o] = z; J.lo[o] = z;
        }
    }
}

// G[(co*Ci + ci)*KS + k] += Wt[(k*Co + co)*Ci + ci]
__global__ void wgrad_scatter_kernel(const float* __restrict__ Wt, float* __restrict__ G, int Co, int Ci, int KS, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;      // index into G
    if (i >= n) return;
    const int k = (int)(i % KS);
    const size_t cc = i / KS;                                             // co*Ci + ci
    G[i] += Wt[(size_t)k * Co * Ci + cc];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn wg_encode() {
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

std::map<std::tuple<const void*, size_t, int, int>, CUtensorMap> g_wg_maps;

// K-major plane [rows][Kld] BF16 -> 2-D map (K, rows), box (64, box_rows), zero fill outside
int plane_map(CUtensorMap* out, const __nv_bfloat16* base, size_t Kld, int rows, int box_rows) {
    auto key = std::make_tuple((const void*)base, Kld, rows, box_rows);
    auto it = g_wg_maps.find(key);
    if (it != g_wg_maps.end()) { *out = it->second; return 0; }
    EncodeFn enc = wg_encode();
    PMP_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)Kld, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Kld * 2};
    cuuint32_t box[2] = {(cuuint32_t)WG_BK, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PMP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(wgrad plane K=%zu rows=%d box=%d) failed: %d", Kld, rows, box_rows, (int)r);
    if (g_wg_maps.size() > 4096) g_wg_maps.clear();
    g_wg_maps[key] = *out;
    return 0;
}

template <int NT>
int launch_wg(const WgP& P, int grid, cudaStream_t st) {
    using C = WgCfg<NT>;
    static bool attr = false;
    if (!attr) {
        PMP_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr = true;
    }
    wgrad_tc_kernel<NT><<<grid, WG_THREADS, C::SMEM, st>>>(P);
    PMP_LAUNCH_OK();
    return 0;
}

}  // namespace

bool wgrad_tc_eligible(int H, int Co, int Ci, int stride, int KS) {
    return stride == 1 && (KS == 1 || KS == 5) && Co % 64 == 0 && Ci % 64 == 0 && H <= 128 && H >= 4;
}

// scratch floats launch_wgrad_tc needs for (R, H, Co, Ci, KS): planes of both operands (hi + lo) + the tap-major workspace
size_t wgrad_tc_scratch_floats(int R, int H, int Co, int Ci, int KS) {
    const size_t K = wg_cols(2, R, H);
    const size_t Kld = K + 64;                                   // tap shifts read up to 2 columns past the end
    return ((size_t)(Co + Ci) * Kld * 2 * 2 + 3) / 4 + (size_t)KS * Co * Ci + 64;
}

// dW[co][ci][k] += sum_{r,h} O1[r,h,co] I1[r,h+k-KS/2,ci] + O2[r,h,co] I2[r,h+k-KS/2,ci]   (stride 1)
int launch_wgrad_tc(const float* O1, int ldo1, const float* I1, int ldi1, const float* O2, int ldo2, const float* I2, int ldi2,
                    int H, int Co, int Ci, int KS, int R, float* dW, float* scratch, cudaStream_t st, int dbg) {
    PMP_REQUIRE(wgrad_tc_eligible(H, Co, Ci, 1, KS), "wgrad_tc: shape not eligible");
    const int nterms = O2 ? 2 : 1;
    const size_t K = wg_cols(2, R, H);
    const size_t Kld = K + 64;
    __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(scratch);
    __nv_bfloat16* ol = oh + (size_t)Co * Kld;
    __nv_bfloat16* ih = ol + (size_t)Co * Kld;
    __nv_bfloat16* il = ih + (size_t)Ci * Kld;
    float* Wt = scratch + ((size_t)(Co + Ci) * Kld * 2 * 2 + 3) / 4;
    Wt = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(Wt) + 15) & ~(uintptr_t)15);
    PMP_CUDA_OK(cudaMemsetAsync(Wt, 0, (size_t)KS * Co * Ci * sizeof(float), st));
    const int G = (R + 7) / 8;
    const size_t tsmem = (size_t)8 * H * 17 * sizeof(float);
    static bool tattr = false;
    if (!tattr) {
        PMP_CUDA_OK(cudaFuncSetAttribute(transpose_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 128 * 17 * 4));
        tattr = true;
    }
    {
        TJobs jobs;
        memset(&jobs, 0, sizeof jobs);
        const float* Os[2] = {O1, O2};
        const float* Is[2] = {I1, I2};
        const int ldos[2] = {ldo1, ldo2}, ldis[2] = {ldi1, ldi2};
        for (int t = 0; t < nterms; ++t) {
            const size_t col0 = (size_t)t * G * (H + 4) * 8;
            const int last = t == nterms - 1;
            jobs.j[jobs.n++] = TJob{Os[t], ldos[t], Co, oh, ol, col0, last};
            jobs.j[jobs.n++] = TJob{Is[t], ldis[t], Ci, ih, il, col0, last};
        }
        // every column a K block or a shifted tap can touch: the used columns rounded up to 64, plus the slack
        const size_t zero_end = wg_cols(nterms, R, H) + 64;
        int vec = 1;      // 16-byte loads when every operand allows them
        for (int i = 0; i < jobs.n; ++i)
            if ((reinterpret_cast<uintptr_t>(jobs.j[i].src) & 15) || (jobs.j[i].ld & 3) || (jobs.j[i].C & 15)) vec = 0;
        transpose_planes_kernel<<<dim3(G, (std::max(Co, Ci) + 15) / 16, jobs.n), 256, tsmem, st>>>(jobs, R, H, Kld, zero_end, vec);
        PMP_LAUNCH_OK();
    }
    WgP P;
    memset(&P, 0, sizeof P);
    const int NT = Ci % 256 == 0 ? 256 : (Ci % 128 == 0 ? 128 : 64);
    int rc;
    if ((rc = plane_map(&P.mAh, oh, Kld, Co, 128))) return rc;
    if ((rc = plane_map(&P.mAl, ol, Kld, Co, 128))) return rc;
    if ((rc = plane_map(&P.mBh, ih, Kld, Ci, NT))) return rc;
    if ((rc = plane_map(&P.mBl, il, Kld, Ci, NT))) return rc;
    P.Co = Co; P.Ci = Ci; P.KS = KS; P.pad = KS / 2;
    const size_t Kused = wg_cols(nterms, R, H);
    P.nkb = (int)(Kused / 64);
    P.mtiles = (Co + 127) / 128; P.ntiles = Ci / NT;
    const int tiles = P.mtiles * P.ntiles * KS;
    int nsplit = (2 * 148 + tiles - 1) / tiles;
    if (nsplit > P.nkb) nsplit = P.nkb;
    if (nsplit < 1) nsplit = 1;
    P.kb_per_split = (P.nkb + nsplit - 1) / nsplit;
    P.nsplit = (P.nkb + P.kb_per_split - 1) / P.kb_per_split;
    P.Wt = Wt;
    P.dbg = dbg;
    const int grid = tiles * P.nsplit;
    rc = NT == 256 ? launch_wg<256>(P, grid, st) : (NT == 128 ? launch_wg<128>(P, grid, st) : launch_wg<64>(P, grid, st));
    if (rc) return rc;
    const size_t n = (size_t)Co * Ci * KS;
    wgrad_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Wt, dW, Co, Ci, KS, n);
    PMP_LAUNCH_OK();
    return 0;
}

}  // namespace pmp
