// Fused sampler update: classifier-free guidance + K-way potential composition + DDPM / DDIM
// posterior step + noise injection + start/goal conditioning in one elementwise kernel.
//
// Replaces (reference file:line):
//   p_mean_variance :177-191, q_posterior :153-159, p_sample :236-246   diffuser/models/diffusion_pb.py
//   pred_x_tm1_luo :197-232, ddim_p_sample :559-612, pred_x_tm1_ddim :670-721
//   apply_conditioning                                                  diffuser/models/helpers.py:168-172
//   PolicyCompose.ddpm/ddim_sample_loop composition                      diffuser/guides/policies_compose.py:286-297,327-342
//   LimitsNormalizer.unnormalize                                         diffuser/datasets/normalization.py:151-162
// Every float op is an explicitly rounded intrinsic in the reference's order (no FMA
// contraction), so with identical eps the result is bit-identical to PyTorch.
// HBM-bound: algorithmic bytes per element = 4*(1 + 2K + [noise] + 1).
#include <stdint.h>

#include "common.cuh"

namespace pmp {
namespace {

struct StepP {
    const float* ec[PMP_MAX_COMPOSE];
    const float* eu[PMP_MAX_COMPOSE];
    float w[PMP_MAX_COMPOSE];
    float coef[8];
    int K, base, kind;
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// Four N(0,1) draws per Philox call, addressed by (global element index / 4, draw index): the value of an
// element depends only on its GLOBAL index, so results are invariant to how the batch is sharded or chunked.
// Both Box-Muller outputs of each uniform pair are used.
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t draw, uint64_t group) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)group, (uint32_t)(group >> 32), (uint32_t)draw, (uint32_t)(draw >> 32)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float u1 = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = ((float)(r.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u3 = ((float)(r.z >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u4 = ((float)(r.w >> 8) + 0.5f) * (1.0f / 16777216.0f);
    // fast intrinsics (lg2/sin/cos.approx): the noise stream only has to be N(0,1)-distributed and reproducible,
    // it has no bit-level counterpart in the reference (torch.randn streams are injected instead when parity matters)
    // -2 ln u = (-2 ln 2) lg2 u; lg2 / sqrt / sin / cos are single MUFU operations
    float ra, rb;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(-1.3862943611198906f * __log2f(u1)));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(-1.3862943611198906f * __log2f(u3)));
    float sa, ca, sb, cb;
    __sincosf(6.283185307179586f * (u2 - 0.5f), &sa, &ca);
    __sincosf(6.283185307179586f * (u4 - 0.5f), &sb, &cb);
    return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
}
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t draw, uint64_t elem) {
    const float4 z = philox_normal4(seed, draw, elem >> 2);
    const int c = (int)(elem & 3);
    return c == 0 ? z.x : (c == 1 ? z.y : (c == 2 ? z.z : z.w));
}

// the update of one element (explicitly rounded, reference op order): shared by the scalar and the float4 kernel
__device__ __forceinline__ float step_elem(const float (&cf)[8], int kind, float xv, float eps, float z) {
    float x0 = __fsub_rn(__fmul_rn(cf[0], xv), __fmul_rn(cf[1], eps));
    x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
    if (kind == PMP_STEP_DDPM) {
        const float mean = __fadd_rn(__fmul_rn(cf[2], x0), __fmul_rn(cf[3], xv));
        return __fadd_rn(mean, __fmul_rn(cf[4], z));
    }
    const float mo = __fdiv_rn(__fsub_rn(xv, __fmul_rn(cf[2], x0)), cf[3]);
    float out = __fadd_rn(__fmul_rn(cf[4], x0), __fmul_rn(cf[5], mo));
    if (cf[7] != 0.0f) out = __fadd_rn(out, __fmul_rn(cf[6], z));
    return out;
}

// Per-step quantities of the whole-loop sampler that live in device memory, so that ONE captured CUDA graph of a
// reverse-diffusion step can be replayed for every step: the step index s (advanced by a one-thread kernel at the end
// of each replay), from which the kernels derive the timestep fed to the U-Net (t_tab[s]), the coefficient row
// (coef_tab[s]) and the Philox draw index (s + 1); seed and global element offset are set per call / row chunk.
struct DynStep {
    const pmp::SamplerState* state;
    const float* coef_tab;       // [n_steps][8]
};
__device__ __forceinline__ void step_inputs(const StepP& p, const DynStep& dyn, float (&cf)[8], uint64_t& seed, uint64_t& draw,
                                            uint64_t& elem0) {
    if (dyn.state) {
        const int s = dyn.state->s;
        seed = dyn.state->seed;
        elem0 = dyn.state->elem0;
        draw = (uint64_t)(s + 1);
#pragma unroll
        for (int i = 0; i < 8; ++i) cf[i] = __ldg(dyn.coef_tab + 8 * s + i);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) cf[i] = p.coef[i];
    }
}

// float4 variant (H*D and the global element offset are multiples of 4): 16-byte loads / stores, one Philox call
// per thread.  HBM-bound: 4*(2 + 2K [+1 injected noise]) bytes per element.
template <int KIND, bool K1>
__global__ void __launch_bounds__(256) step_kernel4(const __grid_constant__ StepP p, const DynStep dyn,
                                                    const float* __restrict__ x,
                                                    const float* __restrict__ noise, uint64_t seed, uint64_t draw,
                                                    uint64_t elem0, const float* __restrict__ start,
                                                    const float* __restrict__ goal, float* __restrict__ xo,
                                                    float* __restrict__ trace, long trace_stride, int H, int D,
                                                    size_t n4, unsigned magic) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n4) return;
    float cf[8];
    step_inputs(p, dyn, cf, seed, draw, elem0);
    const size_t e = g * 4;
    const int HD = H * D;
    // plan index of this float4: a multiply-high by the launcher's reciprocal (exact for the whole launch, see launch_step),
    // else one 32-bit division (the launcher guarantees n4 < 2^32); the conditioning rows are found by comparison
    const unsigned b = magic ? __umulhi((unsigned)g, magic) : (unsigned)g / (unsigned)(HD >> 2);
    const int rem = (int)(((unsigned)g - b * (unsigned)(HD >> 2)) << 2);
    const float4 xv4 = __ldg(reinterpret_cast<const float4*>(x + e));
    const float xv[4] = {xv4.x, xv4.y, xv4.z, xv4.w};
    float eps[4];
    if (K1) {
        const float4 c4 = __ldg(reinterpret_cast<const float4*>(p.ec[0] + e)), u4 = __ldg(reinterpret_cast<const float4*>(p.eu[0] + e));
        const float c[4] = {c4.x, c4.y, c4.z, c4.w}, u[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) eps[j] = __fadd_rn(u[j], __fmul_rn(p.w[0], __fsub_rn(c[j], u[j])));
    } else {
        float s[4], ub[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < p.K; ++k) {
            const float4 c4 = __ldg(reinterpret_cast<const float4*>(p.ec[k] + e)), u4 = __ldg(reinterpret_cast<const float4*>(p.eu[k] + e));
            const float c[4] = {c4.x, c4.y, c4.z, c4.w}, u[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float t = __fmul_rn(p.w[k], __fsub_rn(c[j], u[j]));
                s[j] = k == 0 ? t : __fadd_rn(s[j], t);
                if (k == p.base) ub[j] = u[j];
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) eps[j] = __fadd_rn(ub[j], s[j]);
    }
    const bool need_noise = (KIND == PMP_STEP_DDPM) || (cf[7] != 0.0f);
    float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (need_noise) z4 = noise ? __ldg(reinterpret_cast<const float4*>(noise + e)) : philox_normal4(seed, draw, (elem0 + e) >> 2);
    const float z[4] = {z4.x, z4.y, z4.z, z4.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = step_elem(cf, KIND, xv[j], eps[j], z[j]);
    if (rem < D || rem + 4 > HD - D) {          // only the first / last float4s of a plan touch a conditioned waypoint
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = rem + j;
            if (start && r < D) o[j] = start[(size_t)b * D + r];                        // h == 0
            if (goal && r >= HD - D) o[j] = goal[(size_t)b * D + (r - (HD - D))];       // h == H - 1
        }
    }
    const float4 o4 = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(xo + e) = o4;
    if (trace) *reinterpret_cast<float4*>(trace + (size_t)b * trace_stride + rem) = o4;
}

__global__ void __launch_bounds__(256) step_kernel(const __grid_constant__ StepP p, const DynStep dyn,
                                                   const float* __restrict__ x,
                                                   const float* __restrict__ noise, uint64_t seed, uint64_t draw,
                                                   uint64_t elem0, const float* __restrict__ start,
                                                   const float* __restrict__ goal, float* __restrict__ xo,
                                                   float* __restrict__ trace, long trace_stride, int H, int D,
                                                   size_t n) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float cf[8];
    step_inputs(p, dyn, cf, seed, draw, elem0);
    const int HD = H * D;
    const size_t b = e / HD;
    const int rem = (int)(e - b * HD);
    const int h = rem / D, d = rem - h * D;
    const float xv = x[e];
    float eps;
    if (p.K == 1) {
        const float c = p.ec[0][e], u = p.eu[0][e];
        eps = __fadd_rn(u, __fmul_rn(p.w[0], __fsub_rn(c, u)));
    } else {
        float s = __fmul_rn(p.w[0], __fsub_rn(p.ec[0][e], p.eu[0][e]));
        for (int k = 1; k < p.K; ++k) s = __fadd_rn(s, __fmul_rn(p.w[k], __fsub_rn(p.ec[k][e], p.eu[k][e])));
        eps = __fadd_rn(p.eu[p.base][e], s);
    }
    const bool need_noise = (p.kind == PMP_STEP_DDPM) || (cf[7] != 0.0f);
    float z = 0.0f;
    if (need_noise) z = noise ? noise[e] : philox_normal(seed, draw, elem0 + e);
    float out = step_elem(cf, p.kind, xv, eps, z);
    if (start && h == 0) out = start[b * D + d];
    if (goal && h == H - 1) out = goal[b * D + d];
    xo[e] = out;
    if (trace) trace[b * trace_stride + rem] = out;
}

__global__ void __launch_bounds__(256) init_noise_kernel(float* __restrict__ x, const float* __restrict__ noise,
                                                         uint64_t seed, uint64_t draw, uint64_t elem0, float var_temp,
                                                         const float* __restrict__ start, const float* __restrict__ goal,
                                                         float* __restrict__ trace, long trace_stride, int H, int D,
                                                         size_t n) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int HD = H * D;
    const size_t b = e / HD;
    const int rem = (int)(e - b * HD);
    const int h = rem / D, d = rem - h * D;
    const float z = noise ? noise[e] : philox_normal(seed, draw, elem0 + e);
    float out = __fmul_rn(var_temp, z);
    if (start && h == 0) out = start[b * D + d];
    if (goal && h == H - 1) out = goal[b * D + d];
    x[e] = out;
    if (trace) trace[b * trace_stride + rem] = out;
}

// clip to [-1,1], (x+1)/2, *(hi-lo)+lo : three separately rounded float32 ops
__global__ void __launch_bounds__(256) unnormalize_kernel(const float* __restrict__ x, const float* __restrict__ lo,
                                                          const float* __restrict__ hi, float* __restrict__ out,
                                                          size_t n, int D) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int d = (int)(e % D);
    float v = fminf(fmaxf(x[e], -1.0f), 1.0f);
    v = __fdiv_rn(__fadd_rn(v, 1.0f), 2.0f);
    out[e] = __fadd_rn(__fmul_rn(v, __fsub_rn(hi[d], lo[d])), lo[d]);
}

__global__ void pick_first_valid_kernel(const uint8_t* __restrict__ valid, const int32_t* __restrict__ nchk,
                                        int n_prob, int n_samp, int32_t* __restrict__ chosen,
                                        uint8_t* __restrict__ success, int32_t* __restrict__ total) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_prob) return;
    int idx = n_samp - 1, tot = 0;
    uint8_t ok = 0;
    for (int s = 0; s < n_samp; ++s) {
        tot += nchk[(size_t)p * n_samp + s];
        if (valid[(size_t)p * n_samp + s]) {
            idx = s;
            ok = 1;
            break;
        }
    }
    chosen[p] = idx;
    success[p] = ok;
    total[p] = tot;
}

__global__ void sampler_begin_kernel(SamplerState* st, uint64_t seed, uint64_t elem0) {
    st->s = 0;
    st->seed = seed;
    st->elem0 = elem0;
}
__global__ void sampler_advance_kernel(SamplerState* st) { st->s += 1; }
// t2[i] = timestep of the current step for every row of the CFG batch
__global__ void sampler_set_t_kernel(const SamplerState* __restrict__ st, const int32_t* __restrict__ t_tab,
                                     int64_t* __restrict__ t2, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) t2[i] = (int64_t)t_tab[st->s];
}

}  // namespace

int launch_sampler_begin(SamplerState* state, uint64_t seed, uint64_t elem0, cudaStream_t st) {
    sampler_begin_kernel<<<1, 1, 0, st>>>(state, seed, elem0);
    PMP_LAUNCH_OK();
    return 0;
}
int launch_sampler_advance(SamplerState* state, cudaStream_t st) {
    sampler_advance_kernel<<<1, 1, 0, st>>>(state);
    PMP_LAUNCH_OK();
    return 0;
}
int launch_sampler_set_t(const SamplerState* state, const int32_t* t_tab, int64_t* t2, size_t n, cudaStream_t st) {
    if (n == 0) return 0;
    sampler_set_t_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(state, t_tab, t2, n);
    PMP_LAUNCH_OK();
    return 0;
}

int launch_step(int kind, const float* x, const float* const* ec, const float* const* eu, const float* w, int K,
                int base, const float* coef, const float* noise, uint64_t seed, uint64_t draw, uint64_t elem0,
                const float* start, const float* goal, float* xo, float* trace, long trace_stride, int B, int H,
                int D, cudaStream_t st, const SamplerState* state, const float* coef_tab) {
    PMP_REQUIRE(K >= 1 && K <= PMP_MAX_COMPOSE, "K=%d composed potentials unsupported", K);
    PMP_REQUIRE(base >= 0 && base < K, "uncond base index %d out of range", base);
    PMP_REQUIRE(coef || (state && coef_tab), "step coefficients missing");
    DynStep dyn;
    dyn.state = state;
    dyn.coef_tab = coef_tab;
    StepP p;
    for (int k = 0; k < K; ++k) {
        p.ec[k] = ec[k];
        p.eu[k] = eu[k];
        p.w[k] = w[k];
    }
    for (int i = 0; i < 8; ++i) p.coef[i] = coef ? coef[i] : 0.f;
    p.K = K;
    p.base = base;
    p.kind = kind;
    const size_t n = (size_t)B * H * D;
    if (n == 0) return 0;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    bool vec = ((H * D) % 4 == 0) && (n / 4 < 0xffffffffull) && (elem0 % 4 == 0) && al16(x) && al16(xo) && al16(noise) && al16(trace) && (trace_stride % 4 == 0);
    for (int k = 0; k < K; ++k) vec = vec && al16(ec[k]) && al16(eu[k]);
    if (vec) {
        // b = floor(g / q) as umulhi(g, ceil(2^32 / q)): exact while g * (magic * q - 2^32) < 2^32, i.e. for g < 2^32 / q
        const unsigned q = (unsigned)(H * D) >> 2;
        const unsigned magic = (q > 1 && n / 4 < (1ull << 32) / q) ? (unsigned)(((1ull << 32) + q - 1) / q) : 0u;
        const unsigned grid = (unsigned)((n / 4 + 255) / 256);
#define PMP_STEP4(KIND, K1) step_kernel4<KIND, K1><<<grid, 256, 0, st>>>(p, dyn, x, noise, seed, draw, elem0, start, goal, xo, trace, \
                                                                          trace_stride, H, D, n / 4, magic)
        if (kind == PMP_STEP_DDPM) { if (K == 1) PMP_STEP4(PMP_STEP_DDPM, true); else PMP_STEP4(PMP_STEP_DDPM, false); }
        else { if (K == 1) PMP_STEP4(PMP_STEP_DDIM, true); else PMP_STEP4(PMP_STEP_DDIM, false); }
#undef PMP_STEP4
    }
    else
        step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, dyn, x, noise, seed, draw, elem0, start, goal, xo, trace,
                                                                trace_stride, H, D, n);
    PMP_LAUNCH_OK();
    return 0;
}

int launch_init_noise(float* x, const float* noise, uint64_t seed, uint64_t draw, uint64_t elem0, float var_temp,
                      const float* start, const float* goal, float* trace, long trace_stride, int B, int H, int D,
                      cudaStream_t st) {
    const size_t n = (size_t)B * H * D;
    if (n == 0) return 0;
    init_noise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, noise, seed, draw, elem0, var_temp, start, goal,
                                                                  trace, trace_stride, H, D, n);
    PMP_LAUNCH_OK();
    return 0;
}

int launch_unnormalize(const float* x, const float* lo, const float* hi, float* out, int64_t rows, int D,
                       cudaStream_t st) {
    const size_t n = (size_t)rows * D;
    if (n == 0) return 0;
    unnormalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, lo, hi, out, n, D);
    PMP_LAUNCH_OK();
    return 0;
}

int launch_pick_first_valid(const uint8_t* valid, const int32_t* nchk, int n_prob, int n_samp, int32_t* chosen,
                            uint8_t* success, int32_t* total, cudaStream_t st) {
    if (n_prob == 0) return 0;
    pick_first_valid_kernel<<<(n_prob + 127) / 128, 128, 0, st>>>(valid, nchk, n_prob, n_samp, chosen, success, total);
    PMP_LAUNCH_OK();
    return 0;
}

}  // namespace pmp
