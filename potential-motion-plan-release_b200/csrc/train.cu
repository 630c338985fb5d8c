// Kernels of the training step: the energy loss 0.5*mean(w*(dE/dx - noise)^2) needs the
// derivative of dE/dx with respect to every parameter, i.e. a second-order pass through the U-Net.
//
// Replaces (reference file:line):
//   GaussianDiffusionPB.p_losses / loss + loss.backward()     diffuser/models/diffusion_pb.py:480-545,
//                                                             diffuser/utils/training_pbdiff.py:118-131
//   (torch.autograd double backward through TemporalUnet_WCond.forward, temporal_cond.py:245-329)
//   torch.optim.Adam.step + EMA.update_model_average          training_pbdiff.py:133-140, train_utils.py:15-31
//
// Formulation: with v = dL/d(eps) held fixed, dL/dtheta = grad_theta <dE/dx, v> =
// grad_theta of the directional derivative of E along v.  So the step is: forward (F) and backward-data (B1, gives eps;
// both are the sampler's own fused conv launches, their intermediates kept), a TANGENT forward pass (T) that pushes v
// through the network, and one reverse pass (B2) over the pair (value, tangent) whose tangent-side adjoints are
// exactly B1's gradients.  This file holds what T and B2 need beyond plain convolutions: GroupNorm+SiLU tangent and
// second-order adjoint, weight-gradient GEMMs, the small dense layers of the embedding MLPs / ViT with their
// backward passes, the loss, Adam + EMA over the flat parameter arena, and the re-layout of updated weights.
#include "common.cuh"

namespace pmp {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sums of NV values (blockDim.x == 128), result broadcast to every thread
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red /*[4*NV]*/) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < NV; ++i) red[warp * NV + i] = v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = red[i] + red[NV + i] + red[2 * NV + i] + red[3 * NV + i];
}

// SiLU and its first two derivatives
__device__ __forceinline__ void silu_d012(float p, float& a0, float& a1, float& a2) {
    const float s = 1.0f / (1.0f + expf(-p));
    a0 = p * s;
    a1 = s * (1.0f + p * (1.0f - s));
    a2 = s * (1.0f - s) * (2.0f + p * (1.0f - 2.0f * s));
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm(8) + SiLU, tangent:  yd = d(conv output) -> zd = d(activation output)
//   yhd = r (yd - mean(yd) - yhat mean(yhat yd)),  zd = silu'(gamma yhat + beta) gamma yhd (+ addend)
// one block per (sample, group) slab of Hl x GS elements of the channels-last [R][Hl][C] tensors
template <bool CACHE>
__global__ void __launch_bounds__(128) gn_tangent_kernel(const float* __restrict__ yd, int ldyd, const float* __restrict__ yhat,
                                                         const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, const float* __restrict__ addend,
                                                         int ldad, float* __restrict__ zd, int ldz, __nv_bfloat16* __restrict__ zh,
                                                         __nv_bfloat16* __restrict__ zl, float* __restrict__ yhd,
                                                         float* __restrict__ sOut, int Hl, int C) {
    __shared__ float red[8];
    const int r = blockIdx.x >> 3, g = blockIdx.x & 7, GS = C >> 3, n = Hl * GS;
    const size_t row0 = (size_t)r * Hl;
    float s[2] = {0.f, 0.f};
    constexpr int NC = CACHE ? 12 : 1;      // CACHE: at most 12 elements per thread, kept in registers between the passes
    float cd[NC], cy[NC];
#pragma unroll
    for (int it = 0; it < (CACHE ? NC : 1 << 30); ++it) {
        const int e = threadIdx.x + 128 * it;
        if (e >= n) break;
        const int h = e / GS, c = g * GS + (e - h * GS);
        const float d = yd[(row0 + h) * ldyd + c], yh = yhat[(row0 + h) * C + c];
        if (CACHE) { cd[it] = d; cy[it] = yh; }
        s[0] += d;
        s[1] += d * yh;
    }
    block_sum<2>(s, red);
    const float inv = 1.0f / (float)n, md = s[0] * inv, ms = s[1] * inv, rs = rstd[r * 8 + g];
    if (threadIdx.x == 0) sOut[r * 8 + g] = ms;
#pragma unroll
    for (int it = 0; it < (CACHE ? NC : 1 << 30); ++it) {
        const int e = threadIdx.x + 128 * it;
        if (e >= n) break;
        const int h = e / GS, c = g * GS + (e - h * GS);
        const size_t i = (row0 + h) * C + c;
        const float yh = CACHE ? cy[it] : yhat[i];
        const float t = rs * ((CACHE ? cd[it] : yd[(row0 + h) * ldyd + c]) - md - yh * ms);
        yhd[i] = t;
        float a0, a1, a2;
        const float ga = gamma[c];
        silu_d012(ga * yh + beta[c], a0, a1, a2);
        float z = a1 * ga * t;
        if (addend) z += addend[(row0 + h) * ldad + c];
        const size_t o = (row0 + h) * ldz + c;
        zd[o] = z;
        if (zh) split_bf16(z, zh[o], zl[o]);
    }
}

// GroupNorm(8) + SiLU, reverse pass over (value, tangent):  given zbar = adjoint of the activation output and
// gz = adjoint of its tangent (= B1's gradient w.r.t. the activation output), returns ybar = adjoint of the conv output
//   q = gz a1 gamma, g_y = r (q - mean q - yhat mean(q yhat))            (B1's GroupNorm backward, recomputed)
//   Q = zbar a1 gamma + gz a2 gamma^2 yhd
//   ybar = r (Q - mean Q - yhat mean(Q yhat)) - r s g_y - r (yhd mean(q yhat) + yhat mean(q yhd))
// and accumulates d gamma, d beta
// CACHE: the slab has at most 12 elements per thread: q, Q, yhat, yhd stay in registers between the two passes (one global
// read pass, SiLU derivatives evaluated once)
template <bool CACHE>
__global__ void __launch_bounds__(128) gn_adjoint_kernel(const float* __restrict__ zbar, int ldzb, const float* __restrict__ gz,
                                                         int ldgz, const float* __restrict__ yhat, const float* __restrict__ yhd,
                                                         const float* __restrict__ rstd, const float* __restrict__ sIn,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         float* __restrict__ ybar, __nv_bfloat16* __restrict__ yh_,
                                                         __nv_bfloat16* __restrict__ yl_, float* __restrict__ dgamma,
                                                         float* __restrict__ dbeta, float* __restrict__ dbias, int Hl, int C) {
    __shared__ float red[20];
    __shared__ float chg[128], chb[128], chy[128];
    const int r = blockIdx.x >> 3, g = blockIdx.x & 7, GS = C >> 3, n = Hl * GS;
    const size_t row0 = (size_t)r * Hl;
    if (threadIdx.x < GS) { chg[threadIdx.x] = 0.f; chb[threadIdx.x] = 0.f; chy[threadIdx.x] = 0.f; }
    __syncthreads();
    float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    const bool fixed_ch = (128 % GS) == 0;      // e = tid + 128 i  =>  e % GS == tid % GS
    float acc_g = 0.f, acc_b = 0.f, acc_y = 0.f;
    constexpr int NC = CACHE ? 12 : 1;
    float cq[NC], cQ[NC], cyh[NC], cyd[NC];
#pragma unroll
    for (int it = 0; it < (CACHE ? NC : 1 << 30); ++it) {
        const int e = threadIdx.x + 128 * it;
        if (e >= n) break;
        const int h = e / GS, cl = e - h * GS, c = g * GS + cl;
        const size_t i = (row0 + h) * C + c;
        const float yh = yhat[i], yd = yhd[i], ga = gamma[c];
        const float zb = zbar[(row0 + h) * ldzb + c], gzz = gz[(row0 + h) * ldgz + c];
        float a0, a1, a2;
        silu_d012(ga * yh + beta[c], a0, a1, a2);
        const float q = gzz * a1 * ga;
        const float Q = zb * a1 * ga + gzz * a2 * ga * ga * yd;
        if (CACHE) { cq[it] = q; cQ[it] = Q; cyh[it] = yh; cyd[it] = yd; }
        s[0] += q; s[1] += q * yh; s[2] += Q; s[3] += Q * yh; s[4] += q * yd;
        const float dg = zb * a1 * yh + gzz * (a2 * yh * ga * yd + a1 * yd), db = zb * a1 + gzz * a2 * ga * yd;
        if (fixed_ch) { acc_g += dg; acc_b += db; }        // this thread always sees the same channel: sum in registers
        else { atomicAdd(&chg[cl], dg); atomicAdd(&chb[cl], db); }
    }
    if (fixed_ch) { atomicAdd(&chg[threadIdx.x % GS], acc_g); atomicAdd(&chb[threadIdx.x % GS], acc_b); }
    block_sum<5>(s, red);
    const float inv = 1.0f / (float)n;
    const float m1 = s[0] * inv, m2 = s[1] * inv, mQ = s[2] * inv, mQy = s[3] * inv, mqd = s[4] * inv;
    const float rs = rstd[r * 8 + g], ss = sIn[r * 8 + g];
#pragma unroll
    for (int it = 0; it < (CACHE ? NC : 1 << 30); ++it) {
        const int e = threadIdx.x + 128 * it;
        if (e >= n) break;
        const int h = e / GS, c = g * GS + (e - h * GS);
        const size_t i = (row0 + h) * C + c;
        float q, Q, yh, yd;
        if (CACHE) {
            q = cq[it]; Q = cQ[it]; yh = cyh[it]; yd = cyd[it];
        } else {
            yh = yhat[i]; yd = yhd[i];
            const float ga = gamma[c];
            const float zb = zbar[(row0 + h) * ldzb + c], gzz = gz[(row0 + h) * ldgz + c];
            float a0, a1, a2;
            silu_d012(ga * yh + beta[c], a0, a1, a2);
            q = gzz * a1 * ga;
            Q = zb * a1 * ga + gzz * a2 * ga * ga * yd;
        }
        const float gy = rs * (q - m1 - yh * m2);
        const float v = rs * (Q - mQ - yh * mQy) - rs * ss * gy - rs * (yd * m2 + yh * mqd);
        ybar[i] = v;
        if (yh_) split_bf16(v, yh_[i], yl_[i]);
        if (dbias) {
            if (fixed_ch) acc_y += v;
            else atomicAdd(&chy[e - h * GS], v);
        }
    }
    if (dbias && fixed_ch) atomicAdd(&chy[threadIdx.x % GS], acc_y);
    __syncthreads();
    if (threadIdx.x < GS) {
        atomicAdd(dgamma + g * GS + threadIdx.x, chg[threadIdx.x]);
        atomicAdd(dbeta + g * GS + threadIdx.x, chb[threadIdx.x]);
        if (dbias) atomicAdd(dbias + g * GS + threadIdx.x, chy[threadIdx.x]);     // conv bias gradient = column sums of ybar
    }
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradient of a 1-D convolution written as  out[r,ho,co] = sum_{k,ci} W[co,ci,k] in[r, ho*stride + k - pad, ci]:
//   dW[co][ci][k] += sum_terms sum_{r,ho} O[r,ho,co] * I[r, ho*stride + k - pad, ci]
// (a ConvTranspose1d weight [Cin][Cout][K] is the same sum with the roles of its input and output exchanged).
// GEMM view: M = Co, N = Ci*KS, K = R*Ho.  One block = 64 x 64 (co, ci) x all taps; the samples are split over
// blockIdx.z and the partial sums meet in dW through atomicAdd.
struct WgradP {
    const float* O[2]; int ldo[2];
    const float* I[2]; int ldi[2];
    int nterms;
    int Ho, Co, Hi, Ci, stride, pad, R;
    float* dW;
    int rows_per_split;
};

template <int KS>
__global__ void __launch_bounds__(256) wgrad_kernel(const __grid_constant__ WgradP p) {
    extern __shared__ float sm[];
    const int HiP = p.Hi + 8;                 // 4 zero rows in front, >= 4 behind
    float* sO = sm;                           // [Ho][64]
    float* sI = sm + (size_t)p.Ho * 64;       // [HiP][64]
    const int co0 = blockIdx.x * 64, ci0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[KS][4][4];
#pragma unroll
    for (int k = 0; k < KS; ++k)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[k][a][b] = 0.f;
    const int r0 = blockIdx.z * p.rows_per_split;
    const int r1 = min(p.R, r0 + p.rows_per_split);
    for (int e = threadIdx.x; e < 4 * 64; e += 256) sI[e] = 0.f;
    for (int e = threadIdx.x; e < (HiP - 4 - p.Hi) * 64; e += 256) sI[(size_t)(4 + p.Hi) * 64 + e] = 0.f;
    for (int r = r0; r < r1; ++r) {
        for (int t = 0; t < p.nterms; ++t) {
            __syncthreads();
            const float* O = p.O[t] + (size_t)r * p.Ho * p.ldo[t];
            const float* I = p.I[t] + (size_t)r * p.Hi * p.ldi[t];
            for (int e = threadIdx.x; e < p.Ho * 64; e += 256) {
                const int h = e >> 6, c = e & 63;
                sO[e] = (co0 + c < p.Co) ? O[(size_t)h * p.ldo[t] + co0 + c] : 0.f;
            }
            for (int e = threadIdx.x; e < p.Hi * 64; e += 256) {
                const int h = e >> 6, c = e & 63;
                sI[(size_t)(4 + h) * 64 + c] = (ci0 + c < p.Ci) ? I[(size_t)h * p.ldi[t] + ci0 + c] : 0.f;
            }
            __syncthreads();
            for (int ho = 0; ho < p.Ho; ++ho) {
                const float4 o4 = *reinterpret_cast<const float4*>(sO + ho * 64 + ty * 4);
                const float ov[4] = {o4.x, o4.y, o4.z, o4.w};
                const int hb = ho * p.stride - p.pad + 4;
#pragma unroll
                for (int k = 0; k < KS; ++k) {
                    const float4 i4 = *reinterpret_cast<const float4*>(sI + (size_t)(hb + k) * 64 + tx * 4);
                    const float iv[4] = {i4.x, i4.y, i4.z, i4.w};
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) acc[k][a][b] = fmaf(ov[a], iv[b], acc[k][a][b]);
                }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int co = co0 + ty * 4 + a;
        if (co >= p.Co) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int ci = ci0 + tx * 4 + b;
            if (ci >= p.Ci) continue;
#pragma unroll
            for (int k = 0; k < KS; ++k) atomicAdd(p.dW + ((size_t)co * p.Ci + ci) * KS + k, acc[k][a][b]);
        }
    }
}

// out[c] += sum_rows X[row][c]
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int ld, int rows, int C, float* __restrict__ out,
                                                     int rows_per_block) {
    const int c = blockIdx.x * 64 + (threadIdx.x & 63), part = threadIdx.x >> 6;
    __shared__ float red[256];
    float s = 0.f;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    if (c < C)
        for (int r = r0 + part; r < r1; r += 4) s += X[(size_t)r * ld + c];
    red[threadIdx.x] = s;
    __syncthreads();
    if (part == 0 && c < C) atomicAdd(out + c, red[threadIdx.x] + red[threadIdx.x + 64] + red[threadIdx.x + 128] + red[threadIdx.x + 192]);
}

// out[r][c] (+)= sum_h X[r][h][c]      (gradient of a per-sample embedding broadcast over the horizon)
__global__ void hsum_kernel(const float* __restrict__ X, int ld, int Hl, int C, float* __restrict__ out, int ldo, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = (int)(i % C);
    const size_t r = i / C;
    float s = 0.f;
    for (int h = 0; h < Hl; ++h) s += X[(r * Hl + h) * ld + c];
    out[r * ldo + c] = s;
}

// ---------------------------------------------------------------------------------------------------------------
// small dense layers (rows = samples or wall tokens; a few hundred to a few thousand)
// C[m][n] (=|+=) sum_k A(m,k) B(k,n) with arbitrary element strides: the dense layers' backward passes
//   input gradient   din[r][i]  = sum_o dout[r][o] W[o][i]        (A = dout, B = W)
//   weight gradient  dW[o][i]  += sum_r dout[r][o] in[r][i]       (A = dout^T, B = in)
// These products are small (a few hundred rows); what matters is filling the GPU: 32 x 32 tiles, 16-deep K slices, 2 x 2
// register tiles, K split over blockIdx.z (atomic accumulation) until a few hundred blocks exist.
__global__ void __launch_bounds__(256) small_gemm_kernel(const float* __restrict__ A, long sam, long sak, const float* __restrict__ B,
                                                         long sbk, long sbn, float* __restrict__ Cm, int ldc, int M, int N, int K,
                                                         int k_per_split, int mode /*0 store, 1 add, 2 atomic add*/,
                                                         float* __restrict__ colsumA /*[M] += sum_k A(m,k), or null*/,
                                                         const float* __restrict__ biasN /*[N] added once, or null*/) {
    __shared__ float As[16][32 + 1];
    __shared__ float Bs[16][32 + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    const int kbeg = blockIdx.z * k_per_split, kend = min(K, kbeg + k_per_split);
    float acc[2][2] = {};
    float rs[2] = {0.f, 0.f};
    for (int k0 = kbeg; k0 < kend; k0 += 16) {
        __syncthreads();
        for (int i = tid; i < 32 * 16; i += 256) {
            // the faster-varying smem index follows the operand's unit-stride direction
            int mm, kk;
            if (sam == 1) { mm = i & 31; kk = i >> 5; } else { kk = i & 15; mm = i >> 4; }
            As[kk][mm] = (m0 + mm < M && k0 + kk < kend) ? A[(long)(m0 + mm) * sam + (long)(k0 + kk) * sak] : 0.f;
            int nn, kb;
            if (sbn == 1) { nn = i & 31; kb = i >> 5; } else { kb = i & 15; nn = i >> 4; }
            Bs[kb][nn] = (n0 + nn < N && k0 + kb < kend) ? B[(long)(k0 + kb) * sbk + (long)(n0 + nn) * sbn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float a0 = As[kk][ty * 2], a1 = As[kk][ty * 2 + 1], b0 = Bs[kk][tx * 2], b1 = Bs[kk][tx * 2 + 1];
            rs[0] += a0; rs[1] += a1;
            acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
            acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + ty * 2 + i;
        if (m >= M) continue;
        if (colsumA && blockIdx.y == 0 && tx == 0) atomicAdd(colsumA + m, rs[i]);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int n = n0 + tx * 2 + j;
            if (n >= N) continue;
            float* o = Cm + (size_t)m * ldc + n;
            const float v = acc[i][j] + ((biasN && blockIdx.z == 0) ? biasN[n] : 0.f);
            if (mode == 2) atomicAdd(o, v);
            else *o = mode ? *o + v : v;
        }
    }
}

__device__ __forceinline__ float small_act(float x, int kind) {
    if (kind == 0) return x / (1.0f + expf(-x));
    if (kind == 2) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
    const float sp = x > 20.0f ? x : log1pf(expf(x));
    return x * tanhf(sp);
}
__device__ __forceinline__ float small_act_d(float x, int kind) {
    if (kind == 0) {
        const float s = 1.0f / (1.0f + expf(-x));
        return s * (1.0f + x * (1.0f - s));
    }
    if (kind == 2) return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
    const float sp = x > 20.0f ? x : log1pf(expf(x));
    const float th = tanhf(sp), sg = 1.0f / (1.0f + expf(-x));
    return th + x * (1.0f - th * th) * sg;
}
// kind: 0 SiLU, 1 Mish, 2 GELU (erf)
__global__ void act_fwd_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, int cols, size_t n, int kind) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t r = i / cols;
    const int c = (int)(i % cols);
    y[r * ldy + c] = small_act(x[r * ldx + c], kind);
}
__global__ void act_bwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int lddy, float* __restrict__ dx,
                               int lddx, int cols, size_t n, int kind, int accumulate) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t r = i / cols;
    const int c = (int)(i % cols);
    const float v = dy[r * lddy + c] * small_act_d(x[r * ldx + c], kind);
    float* d = dx + r * lddx + c;
    *d = accumulate ? *d + v : v;
}

// LayerNorm over the last dim n (<= 128), one warp per row; saves xhat and rstd
__global__ void ln_fwd_kernel(const float* __restrict__ x, int ldx, int n, const float* __restrict__ g, const float* __restrict__ b,
                              float* __restrict__ y, int ldy, float* __restrict__ xhat, float* __restrict__ rstd, int rows) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    float s = 0.f;
    for (int j = lane; j < n; j += 32) s += x[(size_t)row * ldx + j];
    const float mean = warp_sum(s) / (float)n;
    float q = 0.f;
    for (int j = lane; j < n; j += 32) {
        const float d = x[(size_t)row * ldx + j] - mean;
        q += d * d;
    }
    const float rs = 1.0f / sqrtf(warp_sum(q) / (float)n + 1e-5f);
    if (lane == 0) rstd[row] = rs;
    for (int j = lane; j < n; j += 32) {
        const float xh = (x[(size_t)row * ldx + j] - mean) * rs;
        xhat[(size_t)row * n + j] = xh;
        y[(size_t)row * ldy + j] = xh * g[j] + b[j];
    }
}
// dx (+)= rstd (dyh - mean(dyh) - xhat mean(dyh xhat)), dyh = dy g ;  dg += dy xhat ; db += dy
__global__ void ln_bwd_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ xhat, const float* __restrict__ rstd,
                              const float* __restrict__ g, int n, float* __restrict__ dx, int lddx, float* __restrict__ dg,
                              float* __restrict__ db, int rows, int accumulate) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    float s0 = 0.f, s1 = 0.f;
    for (int j = lane; j < n; j += 32) {
        const float d = dy[(size_t)row * lddy + j], xh = xhat[(size_t)row * n + j];
        s0 += d * g[j];
        s1 += d * g[j] * xh;
        atomicAdd(dg + j, d * xh);
        atomicAdd(db + j, d);
    }
    const float m0 = warp_sum(s0) / (float)n, m1 = warp_sum(s1) / (float)n, rs = rstd[row];
    for (int j = lane; j < n; j += 32) {
        const float xh = xhat[(size_t)row * n + j];
        const float v = rs * (dy[(size_t)row * lddy + j] * g[j] - m0 - xh * m1);
        float* d = dx + (size_t)row * lddx + j;
        *d = accumulate ? *d + v : v;
    }
}

// single-head attention without output projection (vit_vanilla.py:41-66), one block per sample, nt <= 16 tokens, dim 64:
// out = x + softmax(q k^T scale) v ; qkv [R][nt][192]
__global__ void __launch_bounds__(256) attn_fwd_kernel(const float* __restrict__ qkv, const float* __restrict__ x, float* __restrict__ att,
                                                       float* __restrict__ out, int nt, float scale) {
    __shared__ float Q[16 * 192];
    __shared__ float A[16 * 16];
    const int b = blockIdx.x;
    for (int e = threadIdx.x; e < nt * 192; e += 256) Q[e] = qkv[(size_t)b * nt * 192 + e];
    __syncthreads();
    for (int e = threadIdx.x; e < nt * nt; e += 256) {
        const int i = e / nt, j = e - i * nt;
        float s = 0.f;
        for (int d = 0; d < 64; ++d) s = fmaf(Q[i * 192 + d], Q[j * 192 + 64 + d], s);
        A[i * 16 + j] = s * scale;
    }
    __syncthreads();
    if (threadIdx.x < nt) {
        const int i = threadIdx.x;
        float mx = -1e30f;
        for (int j = 0; j < nt; ++j) mx = fmaxf(mx, A[i * 16 + j]);
        float sum = 0.f;
        for (int j = 0; j < nt; ++j) { const float e_ = expf(A[i * 16 + j] - mx); A[i * 16 + j] = e_; sum += e_; }
        for (int j = 0; j < nt; ++j) { A[i * 16 + j] /= sum; att[((size_t)b * nt + i) * nt + j] = A[i * 16 + j]; }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nt * 64; e += 256) {
        const int i = e >> 6, d = e & 63;
        float s = 0.f;
        for (int j = 0; j < nt; ++j) s = fmaf(A[i * 16 + j], Q[j * 192 + 128 + d], s);
        out[(size_t)b * nt * 64 + e] = x[(size_t)b * nt * 64 + e] + s;
    }
}
// dqkv from dout (the residual branch dx += dout is the caller's)
__global__ void __launch_bounds__(256) attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ att,
                                                       const float* __restrict__ dout, float* __restrict__ dqkv, int nt, float scale) {
    __shared__ float Q[16 * 192];
    __shared__ float A[16 * 16], dS[16 * 16];
    __shared__ float dO[16 * 64];
    const int b = blockIdx.x;
    for (int e = threadIdx.x; e < nt * 192; e += 256) Q[e] = qkv[(size_t)b * nt * 192 + e];
    for (int e = threadIdx.x; e < nt * 64; e += 256) dO[e] = dout[(size_t)b * nt * 64 + e];
    for (int e = threadIdx.x; e < nt * nt; e += 256) A[(e / nt) * 16 + e % nt] = att[(size_t)b * nt * nt + e];
    __syncthreads();
    for (int e = threadIdx.x; e < nt * nt; e += 256) {   // dA = dO v^T
        const int i = e / nt, j = e - i * nt;
        float s = 0.f;
        for (int d = 0; d < 64; ++d) s = fmaf(dO[i * 64 + d], Q[j * 192 + 128 + d], s);
        dS[i * 16 + j] = s;
    }
    __syncthreads();
    if (threadIdx.x < nt) {   // softmax backward per row
        const int i = threadIdx.x;
        float dot = 0.f;
        for (int j = 0; j < nt; ++j) dot = fmaf(dS[i * 16 + j], A[i * 16 + j], dot);
        for (int j = 0; j < nt; ++j) dS[i * 16 + j] = A[i * 16 + j] * (dS[i * 16 + j] - dot) * scale;
    }
    __syncthreads();
    float* o = dqkv + (size_t)b * nt * 192;
    for (int e = threadIdx.x; e < nt * 64; e += 256) {
        const int i = e >> 6, d = e & 63;
        float dq = 0.f, dk = 0.f, dv = 0.f;
        for (int j = 0; j < nt; ++j) {
            dq = fmaf(dS[i * 16 + j], Q[j * 192 + 64 + d], dq);     // dQ_i = sum_j dS_ij K_j
            dk = fmaf(dS[j * 16 + i], Q[j * 192 + d], dk);          // dK_i = sum_j dS_ji Q_j
            dv = fmaf(A[j * 16 + i], dO[j * 64 + d], dv);           // dV_i = sum_j A_ji dO_j
        }
        o[i * 192 + d] = dq;
        o[i * 192 + 64 + d] = dk;
        o[i * 192 + 128 + d] = dv;
    }
}

// sinusoidal wall-coordinate features (helpers.py:300-328): f[tok][d*oct+o] = sin(c/norm * 2^o pi), cos part after td*oct
__global__ void pe_fwd_kernel(const float* __restrict__ walls, size_t n_coord, int td, int oct, float norm, float* __restrict__ f) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n_coord) return;
    const size_t tok = i / td;
    const int d = (int)(i % td);
    const float cn = walls[i] / norm;
    float* row = f + tok * (size_t)(td * oct * 2);
    for (int o = 0; o < oct; ++o) {
        const float sc = cn * ldexpf(3.14159274101257324f, o);
        row[d * oct + o] = sinf(sc);
        row[td * oct + d * oct + o] = cosf(sc);
    }
}
// SinusoidalPosEmb (helpers.py:16-36): out[r] = [sin(t f_i), cos(t f_i)], f_i = exp(i * neg_k)
__global__ void sinus_kernel(const int64_t* __restrict__ t, int R, int dim, float neg_k, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = dim / 2;
    if (i >= R * half) return;
    const int r = i / half, k = i - r * half;
    const float e = (float)t[r] * expf((float)k * neg_k);
    out[r * dim + k] = sinf(e);
    out[r * dim + half + k] = cosf(e);
}
// dst[r][c] (=|+=) scale[r] * src[r][c]   (scale may be null); src may broadcast one row (src_rows == 1)
__global__ void copy2d_kernel(const float* __restrict__ src, int lds, int src_rows, float* __restrict__ dst, int ldd, int cols, size_t n,
                              const float* __restrict__ rowscale, int accumulate) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t r = i / cols;
    const int c = (int)(i % cols);
    float v = src[(src_rows == 1 ? 0 : r) * lds + c];
    if (rowscale) v *= rowscale[r];
    float* d = dst + r * ldd + c;
    *d = accumulate ? *d + v : v;
}

// ---------------------------------------------------------------------------------------------------------------
// loss (diffusion_pb.py:505-545, helpers.py:201-214): L = 0.5 mean(w (eps - noise)^2), v = dL/d eps = w (eps - noise) / n
__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ eps, const float* __restrict__ noise,
                                                   const float* __restrict__ w, int HD, size_t n, float inv_n, float* __restrict__ v,
                                                   __nv_bfloat16* vh, __nv_bfloat16* vl, float* __restrict__ loss) {
    __shared__ float red[8];
    float s = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float d = eps[i] - noise[i], ww = w[i % HD];
        s += ww * d * d;
        const float g = ww * d * inv_n;
        v[i] = g;
        if (vh) split_bf16(g, vh[i], vl[i]);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        atomicAdd(loss, 0.5f * t * inv_n);
    }
}
// q_sample + apply_conditioning (+ set_loss_noise): x_noisy = a[t] x0 + b[t] noise, rows 0 and H-1 of x_noisy = x0's
__global__ void qsample_kernel(const float* __restrict__ x0, float* __restrict__ noise, const int64_t* __restrict__ t,
                               const float* __restrict__ sa, const float* __restrict__ sb, int H, int D, size_t n,
                               int zero_cond_noise, float* __restrict__ xn) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t r = i / ((size_t)H * D);
    const int h = (int)((i / D) % H);
    const bool cond = h == 0 || h == H - 1;
    const int tt = (int)t[r];
    xn[i] = cond ? x0[i] : sa[tt] * x0[i] + sb[tt] * noise[i];
    if (cond && zero_cond_noise) noise[i] = 0.f;
}

// ---------------------------------------------------------------------------------------------------------------
// torch.optim.Adam (no weight decay / amsgrad) + the EMA copy of training_pbdiff.py:108-113 over the flat arena.
// ema_mode 0: none, 1: ema = beta ema + (1-beta) p, 2: ema = p (reset_parameters before step_start_ema)
__global__ void adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                float* __restrict__ ema, size_t n, float lr_over_bc1, float beta1, float beta2, float inv_sqrt_bc2,
                                float eps, float grad_scale, int ema_mode, float ema_beta) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i] * grad_scale;
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float pi = p[i] - lr_over_bc1 * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    p[i] = pi;
    if (ema_mode == 1) ema[i] = ema[i] * ema_beta + (1.0f - ema_beta) * pi;
    else if (ema_mode == 2) ema[i] = pi;
}

// re-layout of the updated parameters for the conv kernels: packed fp32 arena element i = flat parameter map[i]
__global__ void repack_kernel(const float* __restrict__ P, const uint32_t* __restrict__ map, float* __restrict__ arena, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = map[i];
    if (s != 0xffffffffu) arena[i] = P[s];
}
// BF16x3 planes [K][N][C] (K-major B operand) of a packed fp32 weight [K][C][N]
__global__ void planes_kernel(const float* __restrict__ w, int K, int C, int N, __nv_bfloat16* __restrict__ hi,
                              __nv_bfloat16* __restrict__ lo) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= (size_t)K * C * N) return;
    const int c = (int)(i % C);
    const int nn = (int)((i / C) % N);
    const int k = (int)(i / ((size_t)C * N));
    const float v = w[((size_t)k * C + c) * N + nn];
    split_bf16(v, hi[i], lo[i]);
}
// all plane pairs of the model in ONE launch: blockIdx.y = weight, blockIdx.x strides over its elements
struct PlaneJob { const float* w; __nv_bfloat16* hi; __nv_bfloat16* lo; int K, C, N; };
__global__ void __launch_bounds__(256) planes_all_kernel(const PlaneJob* __restrict__ jobs) {
    const PlaneJob j = jobs[blockIdx.y];
    const size_t n = (size_t)j.K * j.C * j.N;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % j.C);
        const int nn = (int)((i / j.C) % j.N);
        const int k = (int)(i / ((size_t)j.C * j.N));
        split_bf16(j.w[((size_t)k * j.C + c) * j.N + nn], j.hi[i], j.lo[i]);
    }
}
__global__ void split_planes_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) split_bf16(x[i], hi[i], lo[i]);
}

inline unsigned nblk(size_t n, int b = 256) { return (unsigned)((n + b - 1) / b); }

}  // namespace

int launch_gn_tangent(const float* yd, int ldyd, const float* yhat, const float* rstd, const float* gamma, const float* beta,
                      const float* addend, int ldad, float* zd, int ldz, __nv_bfloat16* zh, __nv_bfloat16* zl, float* yhd, float* s,
                      int R, int Hl, int C, cudaStream_t st) {
    PMP_REQUIRE(C % 8 == 0 && C / 8 <= 128, "GroupNorm(8) tangent: C = %d unsupported", C);
    if (Hl * (C / 8) <= 12 * 128)
        gn_tangent_kernel<true><<<R * 8, 128, 0, st>>>(yd, ldyd, yhat, rstd, gamma, beta, addend, ldad, zd, ldz, zh, zl, yhd, s, Hl, C);
    else
        gn_tangent_kernel<false><<<R * 8, 128, 0, st>>>(yd, ldyd, yhat, rstd, gamma, beta, addend, ldad, zd, ldz, zh, zl, yhd, s, Hl, C);
    PMP_LAUNCH_OK();
    return 0;
}
int launch_gn_adjoint(const float* zbar, int ldzb, const float* gz, int ldgz, const float* yhat, const float* yhd, const float* rstd,
                      const float* s, const float* gamma, const float* beta, float* ybar, __nv_bfloat16* yh, __nv_bfloat16* yl,
                      float* dgamma, float* dbeta, float* dbias, int R, int Hl, int C, cudaStream_t st) {
    PMP_REQUIRE(C % 8 == 0 && C / 8 <= 128, "GroupNorm(8) adjoint: C = %d unsupported", C);
    if (Hl * (C / 8) <= 12 * 128)
        gn_adjoint_kernel<true><<<R * 8, 128, 0, st>>>(zbar, ldzb, gz, ldgz, yhat, yhd, rstd, s, gamma, beta, ybar, yh, yl, dgamma, dbeta,
                                                       dbias, Hl, C);
    else
        gn_adjoint_kernel<false><<<R * 8, 128, 0, st>>>(zbar, ldzb, gz, ldgz, yhat, yhd, rstd, s, gamma, beta, ybar, yh, yl, dgamma, dbeta,
                                                        dbias, Hl, C);
    PMP_LAUNCH_OK();
    return 0;
}

int launch_wgrad(const float* O1, int ldo1, const float* I1, int ldi1, const float* O2, int ldo2, const float* I2, int ldi2,
                 int Ho, int Co, int Hi, int Ci, int stride, int pad, int KS, int R, float* dW, cudaStream_t st) {
    PMP_REQUIRE(KS == 1 || KS == 3 || KS == 4 || KS == 5, "wgrad: kernel size %d unsupported", KS);
    PMP_REQUIRE(pad <= 2 && (Ho - 1) * stride - pad + KS - 1 < Hi + 4, "wgrad: geometry outside the padded tile");
    WgradP p;
    p.O[0] = O1; p.ldo[0] = ldo1; p.I[0] = I1; p.ldi[0] = ldi1;
    p.O[1] = O2; p.ldo[1] = ldo2; p.I[1] = I2; p.ldi[1] = ldi2;
    p.nterms = O2 ? 2 : 1;
    p.Ho = Ho; p.Co = Co; p.Hi = Hi; p.Ci = Ci; p.stride = stride; p.pad = pad; p.R = R; p.dW = dW;
    const int tiles = ((Co + 63) / 64) * ((Ci + 63) / 64);
    int nsplit = (3 * 148 + tiles - 1) / tiles;
    if (nsplit > R) nsplit = R;
    if (nsplit < 1) nsplit = 1;
    p.rows_per_split = (R + nsplit - 1) / nsplit;
    nsplit = (R + p.rows_per_split - 1) / p.rows_per_split;
    dim3 grid((Co + 63) / 64, (Ci + 63) / 64, nsplit);
    const size_t smem = ((size_t)Ho + Hi + 8) * 64 * sizeof(float);
    PMP_REQUIRE(smem <= 100 * 1024, "wgrad: horizon too long for the shared-memory tile");
    static bool attr = false;
    if (!attr) {
        PMP_CUDA_OK(cudaFuncSetAttribute(wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        PMP_CUDA_OK(cudaFuncSetAttribute(wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        PMP_CUDA_OK(cudaFuncSetAttribute(wgrad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        PMP_CUDA_OK(cudaFuncSetAttribute(wgrad_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr = true;
    }
    if (KS == 5) wgrad_kernel<5><<<grid, 256, smem, st>>>(p);
    else if (KS == 4) wgrad_kernel<4><<<grid, 256, smem, st>>>(p);
    else if (KS == 3) wgrad_kernel<3><<<grid, 256, smem, st>>>(p);
    else wgrad_kernel<1><<<grid, 256, smem, st>>>(p);
    PMP_LAUNCH_OK();
    return 0;
}

int launch_colsum(const float* X, int ld, int rows, int C, float* out, cudaStream_t st) {
    if (rows == 0 || C == 0) return 0;
    int nb = (rows + 63) / 64;
    if (nb > 256) nb = 256;
    const int rpb = (rows + nb - 1) / nb;
    dim3 grid((C + 63) / 64, (rows + rpb - 1) / rpb);
    colsum_kernel<<<grid, 256, 0, st>>>(X, ld, rows, C, out, rpb);
    PMP_LAUNCH_OK();
    return 0;
}
int launch_hsum(const float* X, int ld, int R, int Hl, int C, float* out, int ldo, cudaStream_t st) {
    const size_t n = (size_t)R * C;
    if (!n) return 0;
    hsum_kernel<<<nblk(n), 256, 0, st>>>(X, ld, Hl, C, out, ldo, n);
    PMP_LAUNCH_OK();
    return 0;
}
// K slices per block such that (tiles x splits) fills the GPU; returns the number of splits
static int small_gemm_splits(int M, int N, int K, int* kps) {
    const int tiles = ((M + 31) / 32) * ((N + 31) / 32);
    int nsplit = (2 * 148 + tiles - 1) / tiles;
    const int maxs = (K + 31) / 32;                 // at least two 16-deep slices per block
    if (nsplit > maxs) nsplit = maxs;
    if (nsplit < 1) nsplit = 1;
    *kps = ((K + nsplit - 1) / nsplit + 15) / 16 * 16;
    return (K + *kps - 1) / *kps;
}
int launch_lin_dgrad(const float* dout, int lddo, int nout, const float* W, int ldw, float* din, int lddi, int nin, int rows,
                     int accumulate, cudaStream_t st) {
    if (!rows || !nin) return 0;
    int kps;
    const int ns = small_gemm_splits(rows, nin, nout, &kps);
    int mode = accumulate ? 1 : 0;
    if (ns > 1) {
        if (!accumulate) PMP_CUDA_OK(cudaMemset2DAsync(din, (size_t)lddi * 4, 0, (size_t)nin * 4, rows, st));
        mode = 2;
    }
    dim3 grid((rows + 31) / 32, (nin + 31) / 32, ns);
    small_gemm_kernel<<<grid, 256, 0, st>>>(dout, lddo, 1, W, ldw, 1, din, lddi, rows, nin, nout, kps, mode, nullptr, nullptr);
    PMP_LAUNCH_OK();
    return 0;
}
// out[r][c] (=|+=) bias[c] + sum_j in[r][j] W[c][j]
int launch_lin_fwd(const float* in, int ldin, int nin, const float* W, int ldw, const float* bias, float* out, int ldo, int nout,
                   int rows, int accumulate, cudaStream_t st) {
    if (!rows || !nout) return 0;
    int kps;
    const int ns = small_gemm_splits(rows, nout, nin, &kps);
    int mode = accumulate ? 1 : 0;
    if (ns > 1) {
        if (!accumulate) PMP_CUDA_OK(cudaMemset2DAsync(out, (size_t)ldo * 4, 0, (size_t)nout * 4, rows, st));
        mode = 2;
    }
    dim3 grid((rows + 31) / 32, (nout + 31) / 32, ns);
    small_gemm_kernel<<<grid, 256, 0, st>>>(in, ldin, 1, W, 1, ldw, out, ldo, rows, nout, nin, kps, mode, nullptr, bias);
    PMP_LAUNCH_OK();
    return 0;
}
int launch_lin_wgrad(const float* dout, int lddo, int nout, const float* in, int ldin, int nin, int rows, float* dW, int lddw,
                     float* db, cudaStream_t st) {
    if (!nout || !nin || !rows) return 0;
    int kps;
    const int ns = small_gemm_splits(nout, nin, rows, &kps);
    dim3 grid((nout + 31) / 32, (nin + 31) / 32, ns);
    small_gemm_kernel<<<grid, 256, 0, st>>>(dout, 1, lddo, in, ldin, 1, dW, lddw, nout, nin, rows, kps, 2, db, nullptr);
    PMP_LAUNCH_OK();
    return 0;
}
int launch_act_fwd(const float* x, int ldx, float* y, int ldy, int rows, int cols, int kind, cudaStream_t st) {
    const size_t n = (size_t)rows * cols;
    if (!n) return 0;
    act_fwd_kernel<<<nblk(n), 256, 0, st>>>(x, ldx, y, ldy, cols, n, kind);
    PMP_LAUNCH_OK();
    return 0;
}
int launch_act_bwd(const float* x, int ldx, const float* dy, int lddy, float* dx, int lddx, int rows, int cols, int kind,
                   int accumulate, cudaStream_t st) {
    const size_t n = (size_t)rows * cols;
    if (!n) return 0;
    act_bwd_kernel<<<nblk(n), 256, 0, st>>>(x, ldx, dy, lddy, dx, lddx, cols, n, kind, accumulate);
    PMP_LAUNCH_OK();
    return 0;
}
int launch_ln_fwd(const float* x, int ldx, int n, const float* g, const float* b, float* y, int ldy, float* xhat, float* rstd,
                  int rows, cudaStream_t st) {
    if (!rows) return 0;
    PMP_REQUIRE(n <= 1024, "LayerNorm width %d", n);
    ln_fwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, ldx, n, g, b, y, ldy, xhat, rstd, rows);
    PMP_LAUNCH_OK();
    return 0;
}
int launch_ln_bwd(const float* dy, int lddy, const float* xhat, const float* rstd, const float* g, int n, float* dx, int lddx,
                  float* dg, float* db, int rows, int accumulate, cudaStream_t st) {
    if (!rows) return 0;
    ln_bwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(dy, lddy, xhat, rstd, g, n, dx, lddx, dg, db, rows, accumulate);
    PMP_LAUNCH_OK();
    return 0;
}
int launch_attn_fwd(const float* qkv, const float* x, float* att, float* out, int R, int nt, float scale, cudaStream_t st) {
    PMP_REQUIRE(nt <= 16, "attention: more than 16 tokens (%d)", nt);
    if (!R) return 0;
    attn_fwd_kernel<<<R, 256, 0, st>>>(qkv, x, att, out, nt, scale);
    PMP_LAUNCH_OK();
    return 0;
}
int launch_attn_bwd(const float* qkv, const float* att, const float* dout, float* dqkv, int R, int nt, float scale, cudaStream_t st) {
    PMP_REQUIRE(nt <= 16, "attention: more than 16 tokens (%d)", nt);
    if (!R) return 0;
    attn_bwd_kernel<<<R, 256, 0, st>>>(qkv, att, dout, dqkv, nt, scale);
    PMP_LAUNCH_OK();
    return 0;
}
int launch_pe_fwd(const float* walls, size_t n_coord, int td, int oct, float norm, float* f, cudaStream_t st) {
    if (!n_coord) return 0;
    pe_fwd_kernel<<<nblk(n_coord), 256, 0, st>>>(walls, n_coord, td, oct, norm, f);
    PMP_LAUNCH_OK();
    return 0;
}
int launch_sinus(const int64_t* t, int R, int dim, float neg_k, float* out, cudaStream_t st) {
    if (!R) return 0;
    sinus_kernel<<<nblk((size_t)R * dim / 2), 256, 0, st>>>(t, R, dim, neg_k, out);
    PMP_LAUNCH_OK();
    return 0;
}
int launch_copy2d(const float* src, int lds, int src_rows, float* dst, int ldd, int rows, int cols, const float* rowscale,
                  int accumulate, cudaStream_t st) {
    const size_t n = (size_t)rows * cols;
    if (!n) return 0;
    copy2d_kernel<<<nblk(n), 256, 0, st>>>(src, lds, src_rows, dst, ldd, cols, n, rowscale, accumulate);
    PMP_LAUNCH_OK();
    return 0;
}
int launch_loss(const float* eps, const float* noise, const float* w, int HD, size_t n, float* v, __nv_bfloat16* vh,
                __nv_bfloat16* vl, float* loss, cudaStream_t st) {
    PMP_CUDA_OK(cudaMemsetAsync(loss, 0, sizeof(float), st));
    unsigned g = nblk(n);
    if (g > 1184) g = 1184;
    loss_kernel<<<g, 256, 0, st>>>(eps, noise, w, HD, n, 1.0f / (float)n, v, vh, vl, loss);
    PMP_LAUNCH_OK();
    return 0;
}
int launch_qsample(const float* x0, float* noise, const int64_t* t, const float* sa, const float* sb, int R, int H, int D,
                   int zero_cond_noise, float* xn, cudaStream_t st) {
    const size_t n = (size_t)R * H * D;
    if (!n) return 0;
    qsample_kernel<<<nblk(n), 256, 0, st>>>(x0, noise, t, sa, sb, H, D, n, zero_cond_noise, xn);
    PMP_LAUNCH_OK();
    return 0;
}
int launch_adam_ema(float* p, const float* g, float* m, float* v, float* ema, size_t n, float lr, float beta1, float beta2,
                    float eps, int step, float grad_scale, int ema_mode, float ema_beta, cudaStream_t st) {
    if (!n) return 0;
    const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    adam_ema_kernel<<<nblk(n), 256, 0, st>>>(p, g, m, v, ema, n, (float)(lr / bc1), beta1, beta2, (float)(1.0 / sqrt(bc2)), eps,
                                             grad_scale, ema_mode, ema_beta);
    PMP_LAUNCH_OK();
    return 0;
}
int launch_repack(const float* P, const uint32_t* map, float* arena, size_t n, cudaStream_t st) {
    if (!n) return 0;
    repack_kernel<<<nblk(n), 256, 0, st>>>(P, map, arena, n);
    PMP_LAUNCH_OK();
    return 0;
}
int launch_planes(const float* w, int K, int C, int N, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t st) {
    const size_t n = (size_t)K * C * N;
    if (!n) return 0;
    planes_kernel<<<nblk(n), 256, 0, st>>>(w, K, C, N, hi, lo);
    PMP_LAUNCH_OK();
    return 0;
}
int launch_planes_all(const void* jobs_dev, int njobs, size_t max_elems, cudaStream_t st) {
    if (!njobs) return 0;
    unsigned gx = nblk(max_elems / 4 + 1);
    if (gx > 592) gx = 592;
    planes_all_kernel<<<dim3(gx, njobs), 256, 0, st>>>(reinterpret_cast<const PlaneJob*>(jobs_dev));
    PMP_LAUNCH_OK();
    return 0;
}
int launch_split_planes(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t n, cudaStream_t st) {
    if (!n) return 0;
    split_planes_kernel<<<nblk(n), 256, 0, st>>>(x, hi, lo, n);
    PMP_LAUNCH_OK();
    return 0;
}

}  // namespace pmp
