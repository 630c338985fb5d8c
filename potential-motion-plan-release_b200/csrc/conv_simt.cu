// fp32 CUDA-core implicit-GEMM Conv1d / ConvTranspose1d with the fused U-Net epilogues.
//
// Replaces, for one CFG row batch (reference file:line):
//   nn.Conv1d k5 + GroupNorm(8) + SiLU/Mish      diffuser/models/helpers.py:88-94
//   + time/wall embedding add, residual 1x1 conv   diffuser/models/temporal_dd.py:71-74
//   Downsample1d / Upsample1d                     diffuser/models/helpers.py:38-52
//   and the matching backward-data ops of torch.autograd.grad (temporal_cond.py:321).
//
// This is the exact-fp32 core: it is used for the layers whose shapes do not fit the
// tensor-core kernel (first conv Cin = D, last conv Cout = D, strided convs) and as the
// numerical reference path of the tcgen05 BF16x3 kernel (conv_tc.cu).
#include "common.cuh"

namespace pmp {

namespace {
constexpr int KC = 16;    // input channels per smem chunk
constexpr int TM = 8;     // rows per thread
constexpr int PADH = 2;   // zero halo rows on each side of a sample in smem
constexpr int NTHREADS = 256;
constexpr int MAXSLAB = 128;

template <int TN>
__device__ __forceinline__ void load_bfrag(const float* src, float (&b)[TN]) {
    if constexpr (TN == 4) {
        float4 t = *reinterpret_cast<const float4*>(src);
        b[0] = t.x; b[1] = t.y; b[2] = t.z; b[3] = t.w;
    } else {
#pragma unroll
        for (int j = 0; j < TN; j += 2) {
            float2 t = *reinterpret_cast<const float2*>(src + j);
            b[j] = t.x; b[j + 1] = t.y;
        }
    }
}

// one operand slab: acc[i][j] += sum_taps sum_c A[row_i + shift][c] * W[tap][c][n_j]
template <int TN>
__device__ __forceinline__ void run_slab(float (&acc)[TM][TN], const ConvP& p, float* As, float* Bs,
                                         const float* __restrict__ A, int lda, int C,
                                         const float* __restrict__ W, int ntaps, const int* tap_k,
                                         const int* tap_shift, const int (&abase)[TM], int r0, int Sv,
                                         int n0, int tid, int tx) {
    constexpr int NT = 16 * TN;
    const int Hin = p.Hin, HP = p.HP, arows = p.arows, N = p.N;
    const bool vecA = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const bool vecB = ((N & 3) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
    const int rows_in = Sv * Hin;
    for (int c0 = 0; c0 < C; c0 += KC) {
        __syncthreads();  // previous chunk fully consumed
        // ---- A chunk: [rows_in][KC] -> As[kk][s*HP + h + PADH] (transposed)
        for (int idx = tid; idx < rows_in * (KC / 4); idx += NTHREADS) {
            const int row = idx >> 2, c4 = idx & 3;
            const int s = row / Hin, h = row - s * Hin;
            const int c = c0 + c4 * 4;
            const float* src = A + ((size_t)(r0 + s) * Hin + h) * lda + c;
            float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
            if (vecA && c + 3 < C) {
                float4 t = __ldg(reinterpret_cast<const float4*>(src));
                v0 = t.x; v1 = t.y; v2 = t.z; v3 = t.w;
            } else {
                if (c < C) v0 = __ldg(src);
                if (c + 1 < C) v1 = __ldg(src + 1);
                if (c + 2 < C) v2 = __ldg(src + 2);
                if (c + 3 < C) v3 = __ldg(src + 3);
            }
            float* dst = As + (c4 * 4) * arows + s * HP + h + PADH;
            dst[0] = v0; dst[arows] = v1; dst[2 * arows] = v2; dst[3 * arows] = v3;
        }
        // ---- B chunk: W[k][c0+kk][n0 .. n0+NT) for every tap of this phase
        for (int t = 0; t < ntaps; ++t) {
            const int k = tap_k[t];
            for (int idx = tid; idx < KC * (NT / 4); idx += NTHREADS) {
                const int kk = idx / (NT / 4), n4 = idx - kk * (NT / 4);
                const int c = c0 + kk, n = n0 + n4 * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < C) {
                    const float* src = W + ((size_t)k * C + c) * N + n;
                    if (vecB && n + 3 < N) {
                        v = __ldg(reinterpret_cast<const float4*>(src));
                    } else {
                        if (n < N) v.x = __ldg(src);
                        if (n + 1 < N) v.y = __ldg(src + 1);
                        if (n + 2 < N) v.z = __ldg(src + 2);
                        if (n + 3 < N) v.w = __ldg(src + 3);
                    }
                }
                *reinterpret_cast<float4*>(Bs + (t * KC + kk) * NT + n4 * 4) = v;
            }
        }
        __syncthreads();
        // ---- FMA
        for (int t = 0; t < ntaps; ++t) {
            const float* Ab = As + tap_shift[t];
            const float* Bb = Bs + t * KC * NT + tx * TN;
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) {
                float a[TM], b[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) a[i] = Ab[kk * arows + abase[i]];
                load_bfrag<TN>(Bb + kk * NT, b);
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
    }
}

// deterministic (fixed order) reduction of per-thread partial sums to one (sample, group) slab
__device__ __forceinline__ float slab_reduce(const float* part, int s_local, int g_local, int HJ,
                                             int gs, int TN) {
    const int ty_lo = (s_local * HJ) / TM;
    int ty_hi = ((s_local + 1) * HJ - 1) / TM;
    if (ty_hi > 15) ty_hi = 15;
    const int tx_lo = (g_local * gs) / TN, tx_hi = ((g_local + 1) * gs) / TN;
    float sum = 0.f;
    for (int ty = ty_lo; ty <= ty_hi; ++ty) {
        const int seg = s_local - (ty * TM) / HJ;
        if (seg == 0 || seg == 1) {
            const float* row = part + seg * NTHREADS + ty * 16;
            for (int tx = tx_lo; tx < tx_hi; ++tx) sum += row[tx];
        }
    }
    return sum;
}

template <int TN, bool ACC2>
__global__ void __launch_bounds__(NTHREADS) conv_simt_kernel(const __grid_constant__ ConvP p) {
    constexpr int NT = 16 * TN;
    extern __shared__ __align__(16) float smem[];
    float* As = smem;
    float* Bs = smem + KC * p.arows;

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int r0 = blockIdx.x * p.S, n0 = blockIdx.y * NT, ph = blockIdx.z;
    const int Sv = min(p.S, p.R - r0);
    const int HJ = p.HJ;
    const int Mv = Sv * HJ;

    int abase[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = ty * TM + i;
        if (m < Mv) {
            const int s = m / HJ, j = m - s * HJ;
            abase[i] = s * p.HP + j * p.istride + PADH;
        } else {
            abase[i] = PADH;
        }
    }
    // zero the A tile once: halo rows and rows of absent samples stay zero for the whole kernel
    for (int i = tid; i < KC * p.arows; i += NTHREADS) As[i] = 0.f;

    float acc[TM][TN];
    float acc2[ACC2 ? TM : 1][ACC2 ? TN : 1];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    if constexpr (ACC2) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc2[i][j] = 0.f;
    }

    run_slab<TN>(acc, p, As, Bs, p.A1, p.lda1, p.C1, p.W1, p.ntaps[ph], p.tap_k[ph], p.tap_shift[ph],
                 abase, r0, Sv, n0, tid, tx);
    if (p.A2 != nullptr) {
        const int k0[1] = {0}, s0[1] = {0};
        if constexpr (ACC2) {
            run_slab<TN>(acc2, p, As, Bs, p.A2, p.lda2, p.C2, p.W2, 1, k0, s0, abase, r0, Sv, n0, tid, tx);
        } else {
            run_slab<TN>(acc, p, As, Bs, p.A2, p.lda2, p.C2, p.W2, 1, k0, s0, abase, r0, Sv, n0, tid, tx);
        }
    }

    // ------------------------------------------------------------------ epilogue
    __syncthreads();  // As / Bs are dead: reuse As as reduction scratch
    float* part = As;                    // [2 quantities][2 segments][256]
    float* stat = As + 4 * NTHREADS;     // [2][MAXSLAB]
    const bool gn = p.gn_fwd || p.gn_bwd;
    const int gs = gn ? (p.N >> 3) : NT;
    const int G = gn ? (NT / gs) : 1;
    const int my_g = (tx * TN) / gs;
    const int sa = (ty * TM) / HJ;
    const float inv_cnt = 1.0f / (float)(HJ * gs);
    const int act = p.act;

    bool vr[TM];
    int seg[TM];
    size_t orow[TM];
    int rr[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = ty * TM + i;
        vr[i] = m < Mv;
        const int s = vr[i] ? m / HJ : sa;
        const int j = m - s * HJ;
        seg[i] = s - sa;
        rr[i] = r0 + s;
        orow[i] = (size_t)(r0 + s) * p.Hout + (size_t)(j * p.ostride + ph);
    }
    const int nb = n0 + tx * TN;
    bool vc[TN];
    float bia[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        vc[j] = nb + j < p.N;
        bia[j] = (p.bias && vc[j]) ? __ldg(p.bias + nb + j) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] += bia[j];

    if (p.gn_fwd) {
        // two-pass mean / variance over (HJ rows x gs channels) of every (sample, group) slab
        float ps[2] = {0.f, 0.f};
#pragma unroll
        for (int i = 0; i < TM; ++i)
            if (vr[i]) {
                float t = 0.f;
#pragma unroll
                for (int j = 0; j < TN; ++j) t += acc[i][j];
                ps[seg[i]] += t;
            }
        part[tid] = ps[0];
        part[NTHREADS + tid] = ps[1];
        __syncthreads();
        if (tid < p.S * G) stat[tid] = slab_reduce(part, tid / G, tid % G, HJ, gs, TN) * inv_cnt;
        __syncthreads();
        float mean[2];
        mean[0] = stat[min(sa, p.S - 1) * G + my_g];
        mean[1] = stat[min(sa + 1, p.S - 1) * G + my_g];
        ps[0] = ps[1] = 0.f;
#pragma unroll
        for (int i = 0; i < TM; ++i)
            if (vr[i]) {
                float t = 0.f;
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    const float d = acc[i][j] - mean[seg[i]];
                    t += d * d;
                }
                ps[seg[i]] += t;
            }
        part[tid] = ps[0];
        part[NTHREADS + tid] = ps[1];
        __syncthreads();
        if (tid < p.S * G) {
            const float var = slab_reduce(part, tid / G, tid % G, HJ, gs, TN) * inv_cnt;
            const float rs = 1.0f / sqrtf(var + 1e-5f);
            stat[MAXSLAB + tid] = rs;
            const int s = tid / G;
            if (s < Sv) p.rstd[(size_t)(r0 + s) * 8 + n0 / gs + tid % G] = rs;
        }
        __syncthreads();
        float rs[2];
        rs[0] = stat[MAXSLAB + min(sa, p.S - 1) * G + my_g];
        rs[1] = stat[MAXSLAB + min(sa + 1, p.S - 1) * G + my_g];
        float gam[TN], bet[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            gam[j] = vc[j] ? __ldg(p.gamma + nb + j) : 0.f;
            bet[j] = vc[j] ? __ldg(p.beta + nb + j) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
            if (vr[i]) {
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    const float yh = (acc[i][j] - mean[seg[i]]) * rs[seg[i]];
                    if (vc[j]) p.yhat[orow[i] * p.ldy + nb + j] = yh;
                    acc[i][j] = act_fwd(gam[j] * yh + bet[j], act);
                }
            }
    }

    // additive terms
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        if (!vr[i]) continue;
        const float* eT = nullptr;
        if (p.embT) {
            int64_t tt = p.tidx[rr[i]];
            tt = tt < 0 ? 0 : (tt > p.tmax ? p.tmax : tt);
            eT = p.embT + (size_t)tt * p.ldeT + p.eoff + nb;
        }
        const float* eR = (p.embR && rr[i] < p.n_cond) ? p.embR + (size_t)rr[i] * p.ldeR + p.eoff + nb : nullptr;
        const float* idp = p.ident ? p.ident + orow[i] * p.ldi + nb : nullptr;
        const float* adp = p.addend ? p.addend + orow[i] * p.ldad + nb : nullptr;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            if (!vc[j]) continue;
            float v = acc[i][j];
            if (eT) v += __ldg(eT + j);
            if (eR) v += __ldg(eR + j);
            if constexpr (ACC2) v += acc2[i][j] + (p.bias2 ? __ldg(p.bias2 + nb + j) : 0.f);
            if (idp) v += __ldg(idp + j);
            if (adp) v += __ldg(adp + j);
            acc[i][j] = v;
        }
    }

    if (p.out) {
        const bool vec = (TN == 4) && ((p.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) &&
                         (nb + TN <= p.N);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            if (!vr[i]) continue;
            float* o = p.out + orow[i] * p.ldo + nb;
            if (vec) {
                *reinterpret_cast<float4*>(o) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j)
                    if (vc[j]) o[j] = acc[i][j];
            }
            if (p.out_hi) {
#pragma unroll
                for (int j = 0; j < TN; ++j)
                    if (vc[j]) split_bf16(acc[i][j], p.out_hi[orow[i] * p.ldo + nb + j], p.out_lo[orow[i] * p.ldo + nb + j]);
            }
        }
    }

    if (p.energy) {
        float ps[2] = {0.f, 0.f};
#pragma unroll
        for (int i = 0; i < TM; ++i)
            if (vr[i]) {
#pragma unroll
                for (int j = 0; j < TN; ++j)
                    if (vc[j]) ps[seg[i]] += 0.5f * acc[i][j] * acc[i][j];
            }
        if (ps[0] != 0.f) atomicAdd(p.energy + r0 + sa, ps[0]);
        if (ps[1] != 0.f && sa + 1 < Sv) atomicAdd(p.energy + r0 + sa + 1, ps[1]);
    }

    if (p.gn_bwd) {
        // GroupNorm + activation backward of g_a = acc
        float yh[TM][TN];
        float gam[TN], bet[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            gam[j] = vc[j] ? __ldg(p.gamma + nb + j) : 0.f;
            bet[j] = vc[j] ? __ldg(p.beta + nb + j) : 0.f;
        }
        float q1[2] = {0.f, 0.f}, q2[2] = {0.f, 0.f};
#pragma unroll
        for (int i = 0; i < TM; ++i) {
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                yh[i][j] = 0.f;
                if (vr[i] && vc[j]) {
                    const float y = __ldg(p.yhat + orow[i] * p.ldy + nb + j);
                    const float z = gam[j] * y + bet[j];
                    const float gyh = acc[i][j] * act_bwd(z, act) * gam[j];
                    yh[i][j] = y;
                    acc[i][j] = gyh;
                    q1[seg[i]] += gyh;
                    q2[seg[i]] += gyh * y;
                } else {
                    acc[i][j] = 0.f;
                }
            }
        }
        __syncthreads();
        part[tid] = q1[0];
        part[NTHREADS + tid] = q1[1];
        part[2 * NTHREADS + tid] = q2[0];
        part[3 * NTHREADS + tid] = q2[1];
        __syncthreads();
        if (tid < p.S * G) {
            stat[tid] = slab_reduce(part, tid / G, tid % G, HJ, gs, TN) * inv_cnt;
            stat[MAXSLAB + tid] = slab_reduce(part + 2 * NTHREADS, tid / G, tid % G, HJ, gs, TN) * inv_cnt;
        }
        __syncthreads();
        const int gidx = n0 / gs + my_g;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            if (!vr[i]) continue;
            const int sl = min(sa + seg[i], p.S - 1) * G + my_g;
            const float m1 = stat[sl], m2 = stat[MAXSLAB + sl];
            const float rs = __ldg(p.rstd + (size_t)rr[i] * 8 + gidx);
            float* o = p.gy + orow[i] * p.ldg + nb;
#pragma unroll
            for (int j = 0; j < TN; ++j)
                if (vc[j]) {
                    const float gv = rs * (acc[i][j] - m1 - yh[i][j] * m2);
                    if (p.gy) o[j] = gv;
                    if (p.gy_hi) split_bf16(gv, p.gy_hi[orow[i] * p.ldg + nb + j], p.gy_lo[orow[i] * p.ldg + nb + j]);
                }
        }
    }
}

// standalone GroupNorm/activation backward for gradients that are not produced whole-slab by
// one conv tile (transposed-conv outputs, halves of a skip concat).  HBM-bound: one block per (sample, group)
// slab, float4 loads of the gradient and yhat (the slab, <= 6 KB, is re-read from L1 in the second pass),
// packed 8-byte stores of the BF16 hi / lo planes.
__device__ __forceinline__ float gn_act_bwd(float z, int act) {
    if (act == 0) {
        float e, s;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(1.0f + e));
        return s * fmaf(z, 1.0f - s, 1.0f);
    }
    return act_bwd(z, act);
}
__global__ void __launch_bounds__(128) gn_bwd_kernel(const float* __restrict__ g, int ldgi,
                                                     const float* __restrict__ yhat, int ldy,
                                                     const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, int act,
                                                     float* __restrict__ gy, int ldg,
                                                     __nv_bfloat16* __restrict__ gy_hi,
                                                     __nv_bfloat16* __restrict__ gy_lo, int Hl, int C) {
    const int r = blockIdx.x >> 3, grp = blockIdx.x & 7;
    const int gs = C >> 3, gs4 = gs >> 2, n4 = Hl * gs4;
    __shared__ float red[2][4];
    float q1 = 0.f, q2 = 0.f;
    for (int e = threadIdx.x; e < n4; e += 128) {
        const int h = e / gs4, c = grp * gs + 4 * (e - h * gs4);
        const size_t row = (size_t)r * Hl + h;
        const float4 y = __ldg(reinterpret_cast<const float4*>(yhat + row * ldy + c));
        const float4 gv = __ldg(reinterpret_cast<const float4*>(g + row * ldgi + c));
        const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c)), be = __ldg(reinterpret_cast<const float4*>(beta + c));
        const float a0 = gv.x * gn_act_bwd(fmaf(ga.x, y.x, be.x), act) * ga.x, a1 = gv.y * gn_act_bwd(fmaf(ga.y, y.y, be.y), act) * ga.y;
        const float a2 = gv.z * gn_act_bwd(fmaf(ga.z, y.z, be.z), act) * ga.z, a3 = gv.w * gn_act_bwd(fmaf(ga.w, y.w, be.w), act) * ga.w;
        q1 += (a0 + a1) + (a2 + a3);
        q2 += (a0 * y.x + a1 * y.y) + (a2 * y.z + a3 * y.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        q1 += __shfl_xor_sync(0xffffffffu, q1, o);
        q2 += __shfl_xor_sync(0xffffffffu, q2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = q1;
        red[1][threadIdx.x >> 5] = q2;
    }
    __syncthreads();
    const float inv = 1.0f / (float)(Hl * gs);
    const float m1 = (red[0][0] + red[0][1] + red[0][2] + red[0][3]) * inv;
    const float m2 = (red[1][0] + red[1][1] + red[1][2] + red[1][3]) * inv;
    const float rs = rstd[(size_t)r * 8 + grp];
    for (int e = threadIdx.x; e < n4; e += 128) {
        const int h = e / gs4, c = grp * gs + 4 * (e - h * gs4);
        const size_t row = (size_t)r * Hl + h;
        const float4 y = __ldg(reinterpret_cast<const float4*>(yhat + row * ldy + c));
        const float4 gv = __ldg(reinterpret_cast<const float4*>(g + row * ldgi + c));
        const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c)), be = __ldg(reinterpret_cast<const float4*>(beta + c));
        float4 o;
        o.x = rs * (gv.x * gn_act_bwd(fmaf(ga.x, y.x, be.x), act) * ga.x - m1 - y.x * m2);
        o.y = rs * (gv.y * gn_act_bwd(fmaf(ga.y, y.y, be.y), act) * ga.y - m1 - y.y * m2);
        o.z = rs * (gv.z * gn_act_bwd(fmaf(ga.z, y.z, be.z), act) * ga.z - m1 - y.z * m2);
        o.w = rs * (gv.w * gn_act_bwd(fmaf(ga.w, y.w, be.w), act) * ga.w - m1 - y.w * m2);
        if (gy) *reinterpret_cast<float4*>(gy + row * ldg + c) = o;
        if (gy_hi) {
            uint32_t h0, l0, h1, l1;
            split_bf16x2(o.x, o.y, h0, l0);
            split_bf16x2(o.z, o.w, h1, l1);
            *reinterpret_cast<uint2*>(gy_hi + row * ldg + c) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(gy_lo + row * ldg + c) = make_uint2(l0, l1);
        }
    }
}

template <int TN, bool ACC2>
int launch_t(const ConvP& p, cudaStream_t st) {
    const int smem = conv_simt_smem_bytes(p, TN);
    static int configured = 0;
    if (configured < smem) {
        PMP_CUDA_OK(cudaFuncSetAttribute(conv_simt_kernel<TN, ACC2>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        configured = 100 * 1024;
    }
    dim3 grid((p.R + p.S - 1) / p.S, (p.N + 16 * TN - 1) / (16 * TN), p.nphase);
    conv_simt_kernel<TN, ACC2><<<grid, NTHREADS, smem, st>>>(p);
    PMP_LAUNCH_OK();
    return 0;
}
}  // namespace

int conv_simt_smem_bytes(const ConvP& p, int TN) {
    int a = KC * p.arows;
    if (a < 4 * NTHREADS + 2 * MAXSLAB) a = 4 * NTHREADS + 2 * MAXSLAB;
    return (a + 5 * KC * 16 * TN) * (int)sizeof(float);
}

int launch_conv_simt(const ConvP& p, cudaStream_t st) {
    PMP_REQUIRE(p.S >= 1 && p.S * p.HJ <= 128, "conv tile: S*HJ = %d*%d exceeds 128 rows", p.S, p.HJ);
    PMP_REQUIRE(p.HJ >= TM, "conv tile: fewer than %d output rows per sample (%d)", TM, p.HJ);
    PMP_REQUIRE(p.arows >= p.S * p.HP, "conv tile: A tile rows too small");
    const bool gn = p.gn_fwd || p.gn_bwd;
    const bool acc2 = (p.A2 != nullptr) && (p.bias2 != nullptr || p.gn_fwd);
    int TN = 4;
    if (gn) {
        PMP_REQUIRE(p.N % 8 == 0, "GroupNorm(8) needs N %% 8 == 0 (N=%d)", p.N);
        const int gs = p.N / 8;
        if (64 % gs != 0) {
            PMP_REQUIRE(gs == 96, "unsupported GroupNorm group size %d", gs);
            TN = 6;
        }
        PMP_REQUIRE(p.S * ((16 * TN) / gs) <= MAXSLAB, "too many GroupNorm slabs per tile");
    } else if (p.N % 96 == 0 && p.N % 64 != 0) {
        TN = 6;
    }
    if (TN == 4) return acc2 ? launch_t<4, true>(p, st) : launch_t<4, false>(p, st);
    return acc2 ? launch_t<6, true>(p, st) : launch_t<6, false>(p, st);
}

int launch_gn_bwd(const float* g, int ldgi, const float* yhat, int ldy, const float* rstd,
                  const float* gamma, const float* beta, int act, float* gy, int ldg,
                  __nv_bfloat16* gy_hi, __nv_bfloat16* gy_lo, int R, int Hl, int C, cudaStream_t st) {
    gn_bwd_kernel<<<R * 8, 128, 0, st>>>(g, ldgi, yhat, ldy, rstd, gamma, beta, act, gy, ldg, gy_hi, gy_lo, Hl, C);
    PMP_LAUNCH_OK();
    return 0;
}

}  // namespace pmp
