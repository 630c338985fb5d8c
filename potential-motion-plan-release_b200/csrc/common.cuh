// Shared declarations of the pmp_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/pmp_b200.h"

namespace pmp {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define PMP_CUDA_OK(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            pmp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                           __LINE__);                                                       \
            return PMP_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

#define PMP_REQUIRE(cond, ...)          \
    do {                                \
        if (!(cond)) {                  \
            pmp::set_error(__VA_ARGS__); \
            return PMP_ERR_ARG;         \
        }                               \
    } while (0)

#define PMP_LAUNCH_OK()                                                                      \
    do {                                                                                     \
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e != cudaSuccess) {                                                             \
            pmp::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),       \
                           __FILE__, __LINE__);                                              \
            return PMP_ERR_CUDA;                                                             \
        }                                                                                    \
        pmp::count_launch();                                                                 \
    } while (0)

// ----------------------------------------------------------------------------------------
// activation (SiLU in every shipped energy-mode config, Mish kept as a switch)
__device__ __forceinline__ float act_fwd(float z, int act) {
    if (act == 0) return z / (1.0f + expf(-z));
    // mish(z) = z * tanh(softplus(z)), softplus with the PyTorch threshold of 20
    float sp = z > 20.0f ? z : log1pf(expf(z));
    return z * tanhf(sp);
}
__device__ __forceinline__ float act_bwd(float z, int act) {
    if (act == 0) {
        float s = 1.0f / (1.0f + expf(-z));
        return s * (1.0f + z * (1.0f - s));
    }
    float sp = z > 20.0f ? z : log1pf(expf(z));
    float th = tanhf(sp);
    float sg = 1.0f / (1.0f + expf(-z));
    return th + z * (1.0f - th * th) * sg;
}

// fast variants for the tensor-core epilogues: ex2.approx / rcp.approx based, ~1e-6 relative error,
// an order of magnitude below the BF16x3 product error of that path (the fp32 SIMT path keeps the
// IEEE versions above)
__device__ __forceinline__ float act_fwd_fast(float z, int act) {
    if (act == 0) return z * __frcp_rn(1.0f + __expf(-z));
    return act_fwd(z, act);
}
__device__ __forceinline__ float act_bwd_fast(float z, int act) {
    if (act == 0) {
        const float s = __frcp_rn(1.0f + __expf(-z));
        return s * fmaf(z, 1.0f - s, 1.0f);
    }
    return act_bwd(z, act);
}
// two floats -> packed BF16x3 words: hi pair and lo pair (one cvt.rn.bf16x2 each)
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ----------------------------------------------------------------------------------------
// implicit-GEMM convolution descriptor shared by the SIMT (fp32) and tcgen05 (bf16x3) cores.
//
// Activations are channels-last [R][Hrows][C] (the reference's own "b h t" layout, so the
// U-Net input/output need no transpose).  One tile = S whole samples x HJ output rows per
// sample (x one output phase) x NT output channels, so GroupNorm slabs (sample, group) never
// straddle a tile.
//   output row  ho = j*ostride + phase            (j < HJ)
//   input  row  hi = j*istride + tap_shift[phase][tap]   (zero outside [0,Hin))
//   weight slab    = tap_k[phase][tap]            of W1 laid out [KS][C1][N]
struct ConvP {
    // geometry
    int R, Hout, Hin;
    int S, HJ, HP;              // samples per tile, out rows per sample per phase, padded in rows (Hin+4)
    int ostride, istride, nphase;
    int ntaps[2];
    int tap_k[2][5];
    int tap_shift[2][5];
    int arows;                  // smem rows allotted to the A tile (>= S*HP, == 2 mod 8)
    // operand slab 1 (KS taps) and optional slab 2 (1x1, shift 0)
    const float* A1; int lda1; int C1; const float* W1;
    const float* A2; int lda2; int C2; const float* W2;
    int N;                      // output channels
    // ---- epilogue (all optional unless stated)
    const float* bias;          // [N]
    const float* bias2;         // [N] bias of the separate slab-2 accumulator
    // forward GroupNorm(8) + activation
    int gn_fwd;                 // 1: v = act(gamma*yhat+beta), saves yhat / rstd
    int gn_bwd;                 // 1: gy = GroupNorm/act backward of v using saved yhat / rstd
    const float* gamma; const float* beta;
    float* yhat; int ldy;       // [R*Hout][ldy]  written by gn_fwd, read by gn_bwd
    float* rstd;                // [R][8]
    int act;
    // embeddings added after the activation (temporal_dd.py:71)
    const float* embT; const int64_t* tidx; int ldeT; int tmax;  // table row = clamp(t of the sample, 0, tmax)
    const float* embR; int ldeR; int n_cond;           // per-row table, rows < n_cond only
    int eoff;
    // residual / skip terms
    const float* ident; int ldi;    // + ident[row][n]
    // the same term given as BF16x3 planes (hi + lo, same [row][ldi] indexing) when the producer kept no fp32
    // copy (tcgen05 paths 1 and 3 only)
    const __nv_bfloat16* ident_hi; const __nv_bfloat16* ident_lo;
    const float* addend; int ldad;  // + addend[row][n]
    float* out; int ldo;            // result after all adds (may be null)
    float* gy; int ldg;             // gn_bwd result
    float* energy;                  // [R] += 0.5*sum v^2   (final 1x1 conv only)
    // small-K convolution evaluated on the CUDA cores inside the tensor kernel's epilogue (K = taps*D,
    // D = trajectory dimension 2..14): out[row][n] += sum_t sum_d inl_x[(row + t - taps/2)][d] * inl_w[t][d][n]
    // inl_mode 1: it IS the conv (no MMA slab; first conv / 1x1 seed), added before GroupNorm;
    // inl_mode 2: residual 1x1 conv of the first block, added after GroupNorm (+ bias2)
    const float* inl_x; int inl_ld; int inl_D; int inl_taps; int inl_mode;
    // the two halves of a CFG batch share one copy of the state: inl_x row = (inl_row0 + row) wrapped at inl_wrap (0: no wrap)
    int inl_row0; int inl_wrap;
    const float* inl_w;
    // BF16x3 operand planes (v ~= hi + lo) of `out` / `gy`, same [row][ld] indexing, for consumers
    // that run on the tcgen05 core (null when the consumer is a SIMT conv)
    __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
    __nv_bfloat16* gy_hi; __nv_bfloat16* gy_lo;
    // operand planes of the inputs (A1 / A2) and BF16x3 weights laid out [tap][N][C] (K-major),
    // used only by the tcgen05 core
    const __nv_bfloat16* A1_hi; const __nv_bfloat16* A1_lo;
    const __nv_bfloat16* A2_hi; const __nv_bfloat16* A2_lo;
    const __nv_bfloat16* W1_hi; const __nv_bfloat16* W1_lo;
    const __nv_bfloat16* W2_hi; const __nv_bfloat16* W2_lo;
};

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

int launch_conv_simt(const ConvP& p, cudaStream_t st);
// tcgen05 core; returns 1 (and launches nothing) when the op does not fit it
int launch_conv_tc(const ConvP& p, cudaStream_t st);
bool conv_tc_eligible(const ConvP& p);
// same kernel with CTA pairs (cta_group::2, M = 256 across the two SMs of a TPC) for tiles of >= 128 channels
int launch_conv_tc_pair(const ConvP& p, cudaStream_t st);
int launch_conv_tc_pair16(const ConvP& p, cudaStream_t st);
// operand-reuse main loop (conv_tc2.cu): stride-1 convs without a separate second accumulator
int launch_conv_tc2(const ConvP& p, cudaStream_t st);
bool conv_tc2_eligible(const ConvP& p);
int launch_gn_bwd(const float* g, int ldgi, const float* yhat, int ldy, const float* rstd,
                  const float* gamma, const float* beta, int act, float* gy, int ldg,
                  __nv_bfloat16* gy_hi, __nv_bfloat16* gy_lo, int R, int Hl, int C, cudaStream_t st);
int conv_simt_smem_bytes(const ConvP& p, int TN);

// embeddings / wall encoder (embed.cu)
struct VitW {
    int token_dim, octaves, depth, dim, mlp, act;
    float norm_factor;
    const float *ln0_w, *ln0_b;         // LayerNorm(pe_dim) (only with positional encoding)
    const float *l1_w, *l1_b;           // Linear(pe_dim -> dim)
    const float *l2_w, *l2_b;           // Linear(dim -> 4 dim)
    const float *l3_w, *l3_b;           // Linear(4 dim -> dim)
    const float *ln1_w, *ln1_b;         // LayerNorm(dim)
    const float *cls;                   // [dim]
    const float *an_w[8], *an_b[8], *qkv[8];            // attention pre-norm + to_qkv [3dim,dim]
    const float *fn_w[8], *fn_b[8], *f0_w[8], *f0_b[8], *f3_w[8], *f3_b[8];
};
int launch_vit(const VitW& w, const float* walls, int B, int n_tok, float* out, cudaStream_t st);
// time MLP of the U-Net top level: temb[t] for a list of timesteps
int launch_time_mlp(const float* w1, const float* b1, const float* w3, const float* b3, int dim,
                    int act, float neg_k, const int64_t* t, int n, float* out /*[n][dim]*/, cudaStream_t st);
// out[r][c] = (bias? bias[c]:0) + sum_j W[c][j0+j] * f(in[r][j]),  j < nin ; f = act or identity
int launch_small_linear(const float* in, int ldin, int nin, const float* W, int ldw, int j0,
                        const float* bias, int pre_act, int act, float* out, int ldo, int nout,
                        int rows, int accumulate, cudaStream_t st);
// hidden = act(Ht[t_r] + (r<n_cond ? Hw[r] : 0)) for time_mlp_config 3
int launch_cfg3_hidden(const float* Ht, const int64_t* t, int tmax, const float* Hw, int n_cond, int R,
                       int hid, int act, float* out, cudaStream_t st);

// sampler (step.cu)
// device-resident state of a running pmp_sample loop (see DynStep in step.cu)
struct SamplerState {
    int s;              // index of the current reverse step
    int pad;
    uint64_t seed;      // Philox seed
    uint64_t elem0;     // global element index of the row chunk's first element
};
// `state` + `coef_tab` (device, [n_steps][8]) instead of `coef` / seed / draw / elem0: the per-step values are read
// from device memory (graph replay); the alignment-dependent kernel choice then assumes elem0 % 4 == 0
int launch_step(int kind, const float* x, const float* const* ec, const float* const* eu,
                const float* w, int K, int base, const float* coef, const float* noise,
                uint64_t seed, uint64_t draw, uint64_t elem0, const float* start, const float* goal,
                float* xo, float* trace, long trace_stride, int B, int H, int D, cudaStream_t st,
                const SamplerState* state = nullptr, const float* coef_tab = nullptr);
int launch_sampler_begin(SamplerState* state, uint64_t seed, uint64_t elem0, cudaStream_t st);
int launch_sampler_advance(SamplerState* state, cudaStream_t st);
int launch_sampler_set_t(const SamplerState* state, const int32_t* t_tab, int64_t* t2, size_t n, cudaStream_t st);
int launch_init_noise(float* x, const float* noise, uint64_t seed, uint64_t draw, uint64_t elem0,
                      float var_temp, const float* start, const float* goal, float* trace,
                      long trace_stride, int B, int H, int D, cudaStream_t st);
int launch_unnormalize(const float* x, const float* lo, const float* hi, float* out, int64_t rows,
                       int D, cudaStream_t st);
int launch_pick_first_valid(const uint8_t* valid, const int32_t* nchk, int n_prob, int n_samp,
                            int32_t* chosen, uint8_t* success, int32_t* total, cudaStream_t st);
int launch_fill_i64(int64_t* p, int64_t v, size_t n, cudaStream_t st);

// training step (train.cu); every pointer is a device pointer, tensors channels-last [R][H][C] unless noted
int launch_gn_tangent(const float* yd, int ldyd, const float* yhat, const float* rstd, const float* gamma, const float* beta,
                      const float* addend, int ldad, float* zd, int ldz, __nv_bfloat16* zh, __nv_bfloat16* zl, float* yhd, float* s,
                      int R, int Hl, int C, cudaStream_t st);
int launch_gn_adjoint(const float* zbar, int ldzb, const float* gz, int ldgz, const float* yhat, const float* yhd, const float* rstd,
                      const float* s, const float* gamma, const float* beta, float* ybar, __nv_bfloat16* yh, __nv_bfloat16* yl,
                      float* dgamma, float* dbeta, float* dbias, int R, int Hl, int C, cudaStream_t st);
// dW[co][ci][k] += sum_{r,ho} O1[r,ho,co] I1[r,ho*stride+k-pad,ci] (+ the same with O2, I2 when O2 != null)
int launch_wgrad(const float* O1, int ldo1, const float* I1, int ldi1, const float* O2, int ldo2, const float* I2, int ldi2,
                 int Ho, int Co, int Hi, int Ci, int stride, int pad, int KS, int R, float* dW, cudaStream_t st);
// the same on the tcgen05 tensor cores (wgrad_tc.cu; stride-1 convs with 64-multiple channel counts): BF16x3, K-major
// transposed planes in `scratch` (wgrad_tc_scratch_floats floats, 16-byte aligned)
bool wgrad_tc_eligible(int H, int Co, int Ci, int stride, int KS);
size_t wgrad_tc_scratch_floats(int R, int H, int Co, int Ci, int KS);
int launch_wgrad_tc(const float* O1, int ldo1, const float* I1, int ldi1, const float* O2, int ldo2, const float* I2, int ldi2,
                    int H, int Co, int Ci, int KS, int R, float* dW, float* scratch, cudaStream_t st, int dbg = 0);
int launch_colsum(const float* X, int ld, int rows, int C, float* out, cudaStream_t st);
int launch_hsum(const float* X, int ld, int R, int Hl, int C, float* out, int ldo, cudaStream_t st);
int launch_lin_fwd(const float* in, int ldin, int nin, const float* W, int ldw, const float* bias, float* out, int ldo, int nout,
                   int rows, int accumulate, cudaStream_t st);
int launch_lin_dgrad(const float* dout, int lddo, int nout, const float* W, int ldw, float* din, int lddi, int nin, int rows,
                     int accumulate, cudaStream_t st);
int launch_lin_wgrad(const float* dout, int lddo, int nout, const float* in, int ldin, int nin, int rows, float* dW, int lddw,
                     float* db, cudaStream_t st);
int launch_act_fwd(const float* x, int ldx, float* y, int ldy, int rows, int cols, int kind, cudaStream_t st);
int launch_act_bwd(const float* x, int ldx, const float* dy, int lddy, float* dx, int lddx, int rows, int cols, int kind,
                   int accumulate, cudaStream_t st);
int launch_ln_fwd(const float* x, int ldx, int n, const float* g, const float* b, float* y, int ldy, float* xhat, float* rstd,
                  int rows, cudaStream_t st);
int launch_ln_bwd(const float* dy, int lddy, const float* xhat, const float* rstd, const float* g, int n, float* dx, int lddx,
                  float* dg, float* db, int rows, int accumulate, cudaStream_t st);
int launch_attn_fwd(const float* qkv, const float* x, float* att, float* out, int R, int nt, float scale, cudaStream_t st);
int launch_attn_bwd(const float* qkv, const float* att, const float* dout, float* dqkv, int R, int nt, float scale, cudaStream_t st);
int launch_pe_fwd(const float* walls, size_t n_coord, int td, int oct, float norm, float* f, cudaStream_t st);
int launch_sinus(const int64_t* t, int R, int dim, float neg_k, float* out, cudaStream_t st);
int launch_copy2d(const float* src, int lds, int src_rows, float* dst, int ldd, int rows, int cols, const float* rowscale,
                  int accumulate, cudaStream_t st);
int launch_loss(const float* eps, const float* noise, const float* w, int HD, size_t n, float* v, __nv_bfloat16* vh,
                __nv_bfloat16* vl, float* loss, cudaStream_t st);
int launch_qsample(const float* x0, float* noise, const int64_t* t, const float* sa, const float* sb, int R, int H, int D,
                   int zero_cond_noise, float* xn, cudaStream_t st);
int launch_adam_ema(float* p, const float* g, float* m, float* v, float* ema, size_t n, float lr, float beta1, float beta2,
                    float eps, int step, float grad_scale, int ema_mode, float ema_beta, cudaStream_t st);
int launch_repack(const float* P, const uint32_t* map, float* arena, size_t n, cudaStream_t st);
int launch_planes(const float* w, int K, int C, int N, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t st);
// jobs_dev: device array of {const float* w; bf16* hi; bf16* lo; int K, C, N} (train.cu PlaneJob)
int launch_planes_all(const void* jobs_dev, int njobs, size_t max_elems, cudaStream_t st);
int launch_split_planes(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t n, cudaStream_t st);

}  // namespace pmp
