// Wall encoder (ViT1D) and time / block embedding kernels.
//
// Replaces (reference file:line):
//   ViT1D.forward + Transformer/Attention/FeedForward   diffuser/models/vit_vanilla.py:12-230
//   PositionalEncoding2D.forward_3Dinput                diffuser/models/helpers.py:300-328
//   SinusoidalPosEmb + TemporalUnet_WCond.time_mlp      diffuser/models/helpers.py:16-36,
//                                                       diffuser/models/temporal_cond.py:126-131
//   ResidualTemporalBlock_dd.time_mlp                   diffuser/models/temporal_dd.py:31-49
// All of it is time-step independent or row independent, so the sampler evaluates the ViT
// once per plan and turns the per-block embedding MLPs into table lookups added inside the
// conv epilogue (SURVEY.md Appendix B "embedding shortcut").
#include "common.cuh"

namespace pmp {
namespace {

constexpr int MAXTOK = 16;
constexpr int VT = 256;

__device__ __forceinline__ float vit_act(float x, int kind, bool ff) {
    if (kind == 0) return x / (1.0f + expf(-x));  // SiLU
    if (ff) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));  // GELU (FeedForward, mish=True)
    float sp = x > 20.0f ? x : log1pf(expf(x));
    return x * tanhf(sp);  // Mish
}

// out[tok][o] = (res? out : 0) + b[o] + sum_j W[o][j] * in[tok][j]
__device__ void tok_linear(const float* in, int ldin, int nin, const float* __restrict__ W,
                           const float* __restrict__ b, int nout, float* out, int ldout, int ntok,
                           int actk, bool ff, bool residual) {
    for (int o = threadIdx.x; o < nout; o += VT) {
        float acc[MAXTOK];
#pragma unroll
        for (int t = 0; t < MAXTOK; ++t) acc[t] = 0.f;
        const float* w = W + (size_t)o * nin;
        for (int j = 0; j < nin; ++j) {
            const float wv = __ldg(w + j);
#pragma unroll
            for (int t = 0; t < MAXTOK; ++t)
                if (t < ntok) acc[t] = fmaf(wv, in[t * ldin + j], acc[t]);
        }
        const float bv = b ? __ldg(b + o) : 0.f;
#pragma unroll
        for (int t = 0; t < MAXTOK; ++t)
            if (t < ntok) {
                float v = acc[t] + bv;
                if (actk >= 0) v = vit_act(v, actk, ff);
                if (residual) v += out[t * ldout + o];
                out[t * ldout + o] = v;
            }
    }
}

// LayerNorm over the last dim (n <= 64) of every token; one warp per token
__device__ void tok_layernorm(const float* in, int ldin, int n, const float* __restrict__ g,
                              const float* __restrict__ b, float* out, int ldout, int ntok) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = warp; t < ntok; t += VT / 32) {
        float s = 0.f;
        for (int j = lane; j < n; j += 32) s += in[t * ldin + j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / (float)n;
        float q = 0.f;
        for (int j = lane; j < n; j += 32) {
            const float d = in[t * ldin + j] - mean;
            q += d * d;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rs = 1.0f / sqrtf(q / (float)n + 1e-5f);
        for (int j = lane; j < n; j += 32)
            out[t * ldout + j] = (in[t * ldin + j] - mean) * rs * __ldg(g + j) + __ldg(b + j);
    }
}

__global__ void __launch_bounds__(VT) vit_kernel(const __grid_constant__ VitW w, const float* __restrict__ walls,
                                                 int n_tok, float* __restrict__ out) {
    __shared__ float X[MAXTOK * 64];
    __shared__ float Y[MAXTOK * 64];
    __shared__ float Hd[MAXTOK * 256];
    __shared__ float att[MAXTOK * MAXTOK];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int td = w.token_dim, dim = w.dim;
    const int wall_len = n_tok * td;
    const int fdim = w.octaves ? td * w.octaves * 2 : td;
    // ---- token features (into Hd, ld 32)
    float* F = Hd;
    for (int i = tid; i < n_tok * td; i += VT) {
        const int n = i / td, d = i - n * td;
        const float c = walls[(size_t)b * wall_len + i];
        if (w.octaves) {
            const float cn = c / w.norm_factor;
            for (int o = 0; o < w.octaves; ++o) {
                const float sc = cn * ldexpf(3.14159274101257324f, o);
                F[n * 32 + d * w.octaves + o] = sinf(sc);
                F[n * 32 + td * w.octaves + d * w.octaves + o] = cosf(sc);
            }
        } else {
            F[n * 32 + d] = c;
        }
    }
    __syncthreads();
    if (w.octaves) {
        tok_layernorm(F, 32, fdim, w.ln0_w, w.ln0_b, F, 32, n_tok);
        __syncthreads();
    }
    // Linear(fdim->dim) -> Linear(dim->4dim) -> act -> Linear(4dim->dim) -> LayerNorm
    tok_linear(F, 32, fdim, w.l1_w, w.l1_b, dim, Y, 64, n_tok, -1, false, false);
    __syncthreads();
    tok_linear(Y, 64, dim, w.l2_w, w.l2_b, w.mlp, Hd, 256, n_tok, w.act, false, false);
    __syncthreads();
    tok_linear(Hd, 256, w.mlp, w.l3_w, w.l3_b, dim, Y, 64, n_tok, -1, false, false);
    __syncthreads();
    tok_layernorm(Y, 64, dim, w.ln1_w, w.ln1_b, X + 64, 64, n_tok);  // tokens 1..n
    for (int j = tid; j < dim; j += VT) X[j] = __ldg(w.cls + j);     // token 0 = cls
    __syncthreads();
    const int nt = n_tok + 1;
    const float scale = 1.0f / sqrtf((float)dim);
    for (int l = 0; l < w.depth; ++l) {
        tok_layernorm(X, 64, dim, w.an_w[l], w.an_b[l], Y, 64, nt);
        __syncthreads();
        tok_linear(Y, 64, dim, w.qkv[l], nullptr, 3 * dim, Hd, 256, nt, -1, false, false);
        __syncthreads();
        for (int i = tid; i < nt * nt; i += VT) {
            const int q = i / nt, k = i - q * nt;
            float s = 0.f;
            for (int j = 0; j < dim; ++j) s = fmaf(Hd[q * 256 + j], Hd[k * 256 + dim + j], s);
            att[q * MAXTOK + k] = s * scale;
        }
        __syncthreads();
        if (tid < nt) {
            float mx = -INFINITY;
            for (int k = 0; k < nt; ++k) mx = fmaxf(mx, att[tid * MAXTOK + k]);
            float sum = 0.f;
            for (int k = 0; k < nt; ++k) {
                const float e = expf(att[tid * MAXTOK + k] - mx);
                att[tid * MAXTOK + k] = e;
                sum += e;
            }
            for (int k = 0; k < nt; ++k) att[tid * MAXTOK + k] /= sum;
        }
        __syncthreads();
        for (int i = tid; i < nt * dim; i += VT) {
            const int q = i / dim, j = i - q * dim;
            float s = 0.f;
            for (int k = 0; k < nt; ++k) s = fmaf(att[q * MAXTOK + k], Hd[k * 256 + 2 * dim + j], s);
            X[q * 64 + j] += s;
        }
        __syncthreads();
        tok_layernorm(X, 64, dim, w.fn_w[l], w.fn_b[l], Y, 64, nt);
        __syncthreads();
        tok_linear(Y, 64, dim, w.f0_w[l], w.f0_b[l], w.mlp, Hd, 256, nt, w.act, true, false);
        __syncthreads();
        tok_linear(Hd, 256, w.mlp, w.f3_w[l], w.f3_b[l], dim, X, 64, nt, -1, false, true);
        __syncthreads();
    }
    for (int j = tid; j < dim; j += VT) out[(size_t)b * dim + j] = X[j];
}

// temb = Linear(4dim->dim)(act(Linear(dim->4dim)(sinusoidal(t))))
__global__ void __launch_bounds__(256) time_mlp_kernel(const float* __restrict__ w1, const float* __restrict__ b1,
                                                       const float* __restrict__ w3, const float* __restrict__ b3,
                                                       int dim, int act, float neg_k,
                                                       const int64_t* __restrict__ t, float* __restrict__ out) {
    __shared__ float s[128];
    __shared__ float h[512];
    const int r = blockIdx.x, tid = threadIdx.x;
    const int half = dim / 2;
    const float tv = (float)t[r];
    for (int i = tid; i < half; i += 256) {
        const float f = expf((float)i * neg_k);
        const float e = tv * f;
        s[i] = sinf(e);
        s[half + i] = cosf(e);
    }
    __syncthreads();
    for (int o = tid; o < 4 * dim; o += 256) {
        float a = __ldg(b1 + o);
        for (int j = 0; j < dim; ++j) a = fmaf(__ldg(w1 + (size_t)o * dim + j), s[j], a);
        h[o] = act_fwd(a, act);
    }
    __syncthreads();
    for (int o = tid; o < dim; o += 256) {
        float a = __ldg(b3 + o);
        for (int j = 0; j < 4 * dim; ++j) a = fmaf(__ldg(w3 + (size_t)o * 4 * dim + j), h[j], a);
        out[(size_t)r * dim + o] = a;
    }
}

// out[r][c] (+)= bias[c] + sum_j W[c*ldw + j0 + j] * f(in[r*ldin + j]);   64x64 tiles, 4x4 per thread
__global__ void __launch_bounds__(256) small_linear_kernel(const float* __restrict__ in, int ldin, int nin,
                                                           const float* __restrict__ W, int ldw, int j0,
                                                           const float* __restrict__ bias, int pre_act, int act,
                                                           float* __restrict__ out, int ldo, int nout, int rows,
                                                           int accumulate) {
    __shared__ float Is[16][64 + 4];
    __shared__ float Ws[16][64 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < nin; k0 += 16) {
        __syncthreads();
        for (int i = tid; i < 64 * 16; i += 256) {
            const int rr = i >> 4, kk = i & 15;
            float v = 0.f;
            if (r0 + rr < rows && k0 + kk < nin) {
                v = in[(size_t)(r0 + rr) * ldin + k0 + kk];
                if (pre_act) v = act_fwd(v, act);
            }
            Is[kk][rr] = v;
            float wv = 0.f;
            if (c0 + rr < nout && k0 + kk < nin) wv = W[(size_t)(c0 + rr) * ldw + j0 + k0 + kk];
            Ws[kk][rr] = wv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = Is[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Ws[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty * 4 + i;
        if (r >= rows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + tx * 4 + j;
            if (c >= nout) continue;
            float v = acc[i][j] + (bias ? bias[c] : 0.f);
            float* o = out + (size_t)r * ldo + c;
            *o = accumulate ? (*o + v) : v;
        }
    }
}

__global__ void cfg3_hidden_kernel(const float* __restrict__ Ht, const int64_t* __restrict__ t, int tmax,
                                   const float* __restrict__ Hw, int n_cond, int R, int hid, int act,
                                   float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)R * hid) return;
    const int r = (int)(i / hid), j = (int)(i - (size_t)r * hid);
    const int64_t tt = t[r] < 0 ? 0 : (t[r] > tmax ? tmax : t[r]);
    float v = Ht[(size_t)tt * hid + j];
    if (r < n_cond) v += Hw[(size_t)r * hid + j];
    out[i] = act_fwd(v, act);
}

__global__ void fill_i64_kernel(int64_t* p, int64_t v, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace

int launch_fill_i64(int64_t* p, int64_t v, size_t n, cudaStream_t st) {
    if (n == 0) return 0;
    fill_i64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, v, n);
    PMP_LAUNCH_OK();
    return 0;
}

int launch_vit(const VitW& w, const float* walls, int B, int n_tok, float* out, cudaStream_t st) {
    PMP_REQUIRE(n_tok >= 1 && n_tok + 1 <= MAXTOK, "wall encoder supports 1..%d walls (got %d)", MAXTOK - 1, n_tok);
    PMP_REQUIRE(w.dim == 64 && w.mlp == 256, "wall encoder: only embed_dim 64 / mlp 256 is built (got %d/%d)", w.dim, w.mlp);
    PMP_REQUIRE((w.octaves ? w.token_dim * w.octaves * 2 : w.token_dim) <= 32, "wall feature dim > 32");
    PMP_REQUIRE(w.depth <= 8, "wall encoder depth > 8");
    if (B == 0) return 0;
    vit_kernel<<<B, VT, 0, st>>>(w, walls, n_tok, out);
    PMP_LAUNCH_OK();
    return 0;
}

int launch_time_mlp(const float* w1, const float* b1, const float* w3, const float* b3, int dim, int act,
                    float neg_k, const int64_t* t, int n, float* out, cudaStream_t st) {
    PMP_REQUIRE(dim <= 128 && dim % 2 == 0, "time embedding dim %d unsupported", dim);
    time_mlp_kernel<<<n, 256, 0, st>>>(w1, b1, w3, b3, dim, act, neg_k, t, out);
    PMP_LAUNCH_OK();
    return 0;
}

int launch_small_linear(const float* in, int ldin, int nin, const float* W, int ldw, int j0, const float* bias,
                        int pre_act, int act, float* out, int ldo, int nout, int rows, int accumulate,
                        cudaStream_t st) {
    if (rows == 0 || nout == 0) return 0;
    dim3 grid((rows + 63) / 64, (nout + 63) / 64);
    small_linear_kernel<<<grid, 256, 0, st>>>(in, ldin, nin, W, ldw, j0, bias, pre_act, act, out, ldo, nout, rows,
                                              accumulate);
    PMP_LAUNCH_OK();
    return 0;
}

int launch_cfg3_hidden(const float* Ht, const int64_t* t, int tmax, const float* Hw, int n_cond, int R, int hid,
                       int act, float* out, cudaStream_t st) {
    const size_t n = (size_t)R * hid;
    if (n == 0) return 0;
    cfg3_hidden_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Ht, t, tmax, Hw, n_cond, R, hid, act, out);
    PMP_LAUNCH_OK();
    return 0;
}

}  // namespace pmp
