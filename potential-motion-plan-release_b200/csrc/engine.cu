// Host-side engine: model handle, weight re-layout, the U-Net forward + analytic backward
// program, the reverse-diffusion loop and the C ABI (include/pmp_b200.h).
//
// Replaces (reference file:line):
//   TemporalUnet_WCond.__init__/forward      diffuser/models/temporal_cond.py:67-334
//   ResidualTemporalBlock_dd                 diffuser/models/temporal_dd.py:9-74
//   torch.autograd.grad(E, x)                diffuser/models/temporal_cond.py:321  (hand-written VJP)
//   GaussianDiffusionPB.p_sample_loop etc.   diffuser/models/diffusion_pb.py:249-281, 628-664
//   PolicyCompose.*_sample_loop              diffuser/guides/policies_compose.py:270-350
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace pmp {
unsigned long long* tc_timer_buffer();

static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_launches{0};
// pmp_set_option("uncond_dedup", 0) makes same-network compositions evaluate the unconditional half once per
// potential instead of once (more work, bit-identical results; the equivalence test uses it)
static std::atomic<int> g_uncond_dedup{1};
// pmp_set_option("cuda_graph", 0): pmp_sample launches every kernel of every step from the host (A/B, debugging)
static std::atomic<int> g_cuda_graph{1};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n); }

// ---- optional per-launch CUDA-event profiler for the conv kernels (bench.py roofline)
struct ProfRec { cudaEvent_t a, b; double flops; int cls; int R, Hout, N, C1, C2, taps, flags; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_pool;
static double conv_flops(const ConvP& p) {
    double k = 0;
    for (int ph = 0; ph < p.nphase; ++ph) k += (double)p.ntaps[ph] * p.C1;
    k /= p.nphase;   // each output row belongs to exactly one phase
    if (p.A2 || p.A2_hi) k += p.C2;
    return 2.0 * (double)p.R * p.Hout * p.N * k;
}
static int prof_launch_conv(const ConvP& p, cudaStream_t st, int cls, int (*fn)(const ConvP&, cudaStream_t)) {
    if (!g_prof_on) return fn(p, st);
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    if (!g_prof_pool.empty()) {
        ev = g_prof_pool.back();
        g_prof_pool.pop_back();
    } else {
        PMP_CUDA_OK(cudaEventCreate(&ev.first));
        PMP_CUDA_OK(cudaEventCreate(&ev.second));
    }
    PMP_CUDA_OK(cudaEventRecord(ev.first, st));
    int rc = fn(p, st);
    PMP_CUDA_OK(cudaEventRecord(ev.second, st));
    g_prof.push_back(ProfRec{ev.first, ev.second, conv_flops(p), cls, p.R, p.Hout, p.N, p.C1, (p.A2 || p.A2_hi) ? p.C2 : 0,
                             p.ntaps[0] + (p.nphase > 1 ? p.ntaps[1] : 0),
                             p.gn_fwd | (p.gn_bwd << 1) | ((p.ident != nullptr || p.ident_hi != nullptr) << 2) | ((p.istride > 1) << 3) | ((p.ostride > 1) << 4)});
    return rc;
}

struct HostT {
    std::vector<int64_t> shape;
    std::vector<float> v;
    int64_t numel() const { return (int64_t)v.size(); }
};

// forward activation view
struct Act {
    float* p = nullptr;
    int ld = 0, C = 0, Hl = 0;
    int blk = -1;  // residual block that produced it (its CB1 GroupNorm must be differentiated), -1: none
    bool grad_f32 = false;   // its gradient is also consumed in fp32 (standalone GroupNorm backward / skip addend)
};
struct Grad {
    float* raw = nullptr;
    int ldr = 0;
    float* gy = nullptr;  // dense [R*Hl][C]
};

struct TcW {  // BF16x3 weight planes [tap][N][C] in the 16-bit arena (element offsets)
    bool ok = false;
    size_t hi = 0, lo = 0;
};
struct CBW {  // Conv1dBlock_dd: conv k5 + GroupNorm(8) + act
    int Cin = 0, Cout = 0;
    size_t Wf = 0, Wb = 0, bias = 0, gamma = 0, beta = 0;  // offsets into the packed arena
    TcW tcf, tcb;
    // saved for backward (per eps evaluation)
    float* yhat = nullptr;
    float* rstd = nullptr;
};
struct ResB {
    std::string pfx;
    int Cin = 0, Cout = 0;
    CBW cb[2];
    bool has_res = false;
    size_t Rf = 0, Rb = 0, Rbias = 0;
    TcW tcRf, tcRb;
    int eoff = 0;       // channel offset in the embedding tables
    size_t tmW = 0, tmB = 0;      // cfg 0: Linear(128->Cout); cfg 3: second Linear(256->Cout)
    size_t tm0W = 0, tm0B = 0;    // cfg 3: first Linear(128->256)
};
struct Packer16Rec { size_t src; int K, C, N; size_t hi, lo; };
struct TrainState;
struct PlainConv {  // down / up sampling
    int C = 0, KS = 0;
    size_t Wf = 0, Wb = 0, bias = 0;
    TcW tcf, tcb;
    bool present = false;
};

}  // namespace pmp

using namespace pmp;

struct pmp_model {
    pmp_arch arch;
    int device = 0;
    bool finalized = false;
    std::map<std::string, HostT> hw;  // raw reference-layout weights
    std::vector<int> dims;            // [D, dim*m0, dim*m1, ...]
    int nl = 0;
    std::vector<ResB> blocks;         // downs (2 per level), mid1, mid2, ups (2 per level)
    std::vector<PlainConv> downc, upc;
    CBW finalcb;
    size_t fin1Wf = 0, fin1Wb = 0, fin1b = 0;
    TcW fin1tc;
    size_t tW1 = 0, tB1 = 0, tW3 = 0, tB3 = 0;
    VitW vit;
    int Ctot = 0;   // sum of Cout over residual blocks (embedding table width)
    int hid = 0;    // cfg3 hidden width per block (2*time_dim)
    float* arena = nullptr;   // packed weights on device
    size_t arena_n = 0;
    __nv_bfloat16* arena16 = nullptr;   // BF16x3 weight planes for the tcgen05 core
    size_t arena16_n = 0;
    std::vector<Packer16Rec> plane_recs;   // how arena16 derives from arena (training re-layout after an optimizer step)
    pmp::TrainState* train = nullptr;    // bound flat parameter / gradient arenas (pmp_train_bind)
    float* temb_tab = nullptr;  // [T][dim]
    int64_t* t_iota = nullptr;  // [T] 0..T-1
    float* embT = nullptr;      // cfg0: [T][Ctot] ; cfg3: Ht [T][nblk*hid]
    // per-call tables (grown on demand)
    float* embR = nullptr; size_t embR_n = 0;   // cfg0: [B][Ctot]; cfg3: Hw [B][nblk*hid]
    float* embE = nullptr; size_t embE_n = 0;   // cfg3 per step: [R][Ctot]
    float* hidb = nullptr; size_t hidb_n = 0;   // cfg3 per step hidden [R][nblk*hid]
    float* ws = nullptr; size_t ws_n = 0;       // activation workspace (floats)
    int row_chunk = 8192;   // rows of the CFG batch per pass (workspace ~1.8 MB per row for the Maze2D U-Net, grown on demand)
#ifdef PMP_DIAG
    bool keep_fp32 = getenv("PMP_KEEP_FP32") != nullptr;   // diagnostics: fp32 copies of every activation / gradient
#else
    bool keep_fp32 = false;
#endif
    int gemm_path = 3;   // 0 fp32 SIMT, 1 tcgen05 single CTA, 2 operand reuse, 3 CTA pairs + TMA-store epilogue (default)
    uint64_t gen = 0;    // bumped whenever something a captured graph bakes in changes (buffers re-allocated, weights, path)
    // ---- whole-loop sampler (pmp_sample) state owned by the FIRST model of a call: grow-only scratch + graph cache
    float* s_eps = nullptr; size_t s_eps_n = 0;     // [K][2][Bc][H][D] eps of every potential
    float* s_x2 = nullptr; size_t s_x2_n = 0;       // duplicated state (only when the first conv cannot share it)
    float* s_xs = nullptr; size_t s_xs_n = 0;       // state of the row chunk [Bc][H][D]
    float* s_sg = nullptr; size_t s_sg_n = 0;       // start | goal of the row chunk [2][Bc][D]
    float* s_wtab = nullptr; size_t s_wtab_n = 0;   // wall tables of the K potentials, back to back
    float* s_tabs = nullptr; size_t s_tabs_n = 0;   // t2 [2 Bc] int64 | coef_tab [n_steps][8] | t_tab [n_steps] int32 | SamplerState
    struct GraphEnt { std::vector<uint64_t> key; cudaGraphExec_t exec; uint64_t launches; };   // kernels per replay
    std::vector<GraphEnt> graphs;
    cudaStream_t cap_stream = nullptr;   // capture happens on a private stream (the caller's may be the legacy default stream)
    std::map<std::string, std::pair<float*, int64_t>> dbg;
    bool has_down(int L) const { return !(L >= nl - 1 || L >= arch.down_times); }
    bool has_up(int ind) const { return !(ind >= nl - 1 || ind < (nl - arch.down_times - 1)); }
    const float* A(size_t off) const { return arena + off; }
};

namespace pmp {

// Entry points select the model's device and put the caller's current device back on return.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Per-model scratch buffers grow in STREAM ORDER (cudaFreeAsync / cudaMallocAsync on the caller's stream): work already
// queued on that stream keeps its buffer until it has run, and nothing synchronises the device.  pmp_model_reserve
// sizes them up front so that a steady workload never allocates inside a call at all.
static int grow(float** p, size_t* cur, size_t need, cudaStream_t st, uint64_t* gen = nullptr) {
    if (*cur >= need) return 0;
    if (gen) ++*gen;
    if (*p) PMP_CUDA_OK(cudaFreeAsync(*p, st));
    *p = nullptr;
    *cur = 0;
    PMP_CUDA_OK(cudaMallocAsync((void**)p, need * sizeof(float), st));
    *cur = need;
    return 0;
}

// ------------------------------------------------------------------ weight packing
struct Packer {
    std::vector<float> buf;
    size_t add(const std::vector<float>& v) {
        size_t off = (buf.size() + 3) & ~(size_t)3;  // 16-byte aligned
        buf.resize(off + v.size());
        memcpy(buf.data() + off, v.data(), v.size() * sizeof(float));
        return off;
    }
};

// BF16 hi/lo split of a [K][C][N] fp32 weight, transposed to the K-major B-operand layout [K][N][C]
struct Packer16 {
    std::vector<uint16_t> buf;
    static uint16_t bf16_rn(float f) {
        uint32_t u;
        memcpy(&u, &f, 4);
        if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
        u += 0x7fffu + ((u >> 16) & 1u);
        return (uint16_t)(u >> 16);
    }
    static float bf16_f(uint16_t h) {
        uint32_t u = (uint32_t)h << 16;
        float f;
        memcpy(&f, &u, 4);
        return f;
    }
    typedef Packer16Rec Rec;
    std::vector<Rec> recs;   // every plane pair with the fp32 arena offset of the [K][C][N] weight it was split from
    TcW add(const std::vector<float>& w /*[K][C][N]*/, int K, int C, int N, size_t src) {
        TcW t;
        if (C % 64 != 0) return t;   // N may be narrow (trajectory dimension): TMA zero-fills / over-reads unused columns
        const size_t n = (size_t)K * C * N;
        size_t off = (buf.size() + 63) & ~(size_t)63;   // 128-byte aligned planes
        const size_t stride = (n + 63) & ~(size_t)63;
        buf.resize(off + 2 * stride);
        for (int k = 0; k < K; ++k)
            for (int c = 0; c < C; ++c)
                for (int nn = 0; nn < N; ++nn) {
                    const float v = w[((size_t)k * C + c) * N + nn];
                    const uint16_t h = bf16_rn(v);
                    const uint16_t l = bf16_rn(v - bf16_f(h));
                    const size_t o = ((size_t)k * N + nn) * C + c;
                    buf[off + o] = h;
                    buf[off + stride + o] = l;
                }
        t.ok = true; t.hi = off; t.lo = off + stride;
        recs.push_back(Rec{src, K, C, N, t.hi, t.lo});
        return t;
    }
};

static const HostT* find(const pmp_model* m, const std::string& k) {
    auto it = m->hw.find(k);
    return it == m->hw.end() ? nullptr : &it->second;
}
#define NEED(var, key)                                                    \
    const HostT* var = find(m, key);                                      \
    if (!var) {                                                           \
        set_error("weight '%s' was not set", std::string(key).c_str());   \
        return PMP_ERR_STATE;                                             \
    }

// Conv1d weight [Cout][Cin][KS] -> forward [KS][Cin][Cout]; backward [KS][Cout][Cin] with flipped taps
static void conv_fwd_layout(const HostT& w, std::vector<float>& o) {
    const int Co = (int)w.shape[0], Ci = (int)w.shape[1], K = (int)w.shape[2];
    o.resize(w.v.size());
    for (int co = 0; co < Co; ++co)
        for (int ci = 0; ci < Ci; ++ci)
            for (int k = 0; k < K; ++k) o[((size_t)k * Ci + ci) * Co + co] = w.v[((size_t)co * Ci + ci) * K + k];
}
static void conv_bwd_layout(const HostT& w, std::vector<float>& o, bool flip) {
    const int Co = (int)w.shape[0], Ci = (int)w.shape[1], K = (int)w.shape[2];
    o.resize(w.v.size());
    for (int co = 0; co < Co; ++co)
        for (int ci = 0; ci < Ci; ++ci)
            for (int k = 0; k < K; ++k) {
                const int kk = flip ? (K - 1 - k) : k;
                o[((size_t)kk * Co + co) * Ci + ci] = w.v[((size_t)co * Ci + ci) * K + k];
            }
}
// ConvTranspose1d weight [Cin][Cout][KS] -> forward [KS][Cin][Cout]; backward [KS][Cout][Cin]
static void convT_layouts(const HostT& w, std::vector<float>& f, std::vector<float>& b) {
    const int Ci = (int)w.shape[0], Co = (int)w.shape[1], K = (int)w.shape[2];
    f.resize(w.v.size());
    b.resize(w.v.size());
    for (int ci = 0; ci < Ci; ++ci)
        for (int co = 0; co < Co; ++co)
            for (int k = 0; k < K; ++k) {
                const float v = w.v[((size_t)ci * Co + co) * K + k];
                f[((size_t)k * Ci + ci) * Co + co] = v;
                b[((size_t)k * Co + co) * Ci + ci] = v;
            }
}

static int pack_cb(pmp_model* m, Packer& pk, Packer16& pk16, const std::string& pfx, CBW& cb) {
    NEED(w, pfx + "block.0.weight");
    NEED(b, pfx + "block.0.bias");
    NEED(g, pfx + "block.2.weight");
    NEED(be, pfx + "block.2.bias");
    PMP_REQUIRE(w->shape.size() == 3 && w->shape[2] == 5, "%sblock.0.weight must be [Cout,Cin,5]", pfx.c_str());
    cb.Cout = (int)w->shape[0];
    cb.Cin = (int)w->shape[1];
    PMP_REQUIRE(cb.Cout % 8 == 0, "GroupNorm(8) needs Cout %% 8 == 0");
    std::vector<float> t;
    conv_fwd_layout(*w, t);
    cb.Wf = pk.add(t);
    cb.tcf = pk16.add(t, 5, cb.Cin, cb.Cout, cb.Wf);
    conv_bwd_layout(*w, t, true);
    cb.Wb = pk.add(t);
    cb.tcb = pk16.add(t, 5, cb.Cout, cb.Cin, cb.Wb);
    cb.bias = pk.add(b->v);
    cb.gamma = pk.add(g->v);
    cb.beta = pk.add(be->v);
    return 0;
}

static int pack_res(pmp_model* m, Packer& pk, Packer16& pk16, const std::string& pfx, int Cin, int Cout, ResB& rb) {
    rb.pfx = pfx;
    rb.Cin = Cin;
    rb.Cout = Cout;
    int rc;
    if ((rc = pack_cb(m, pk, pk16, pfx + "blocks.0.", rb.cb[0]))) return rc;
    if ((rc = pack_cb(m, pk, pk16, pfx + "blocks.1.", rb.cb[1]))) return rc;
    PMP_REQUIRE(rb.cb[0].Cin == Cin && rb.cb[0].Cout == Cout && rb.cb[1].Cin == Cout && rb.cb[1].Cout == Cout,
                "%s: conv shapes do not match the architecture (%d->%d)", pfx.c_str(), Cin, Cout);
    rb.has_res = find(m, pfx + "residual_conv.weight") != nullptr;
    PMP_REQUIRE(rb.has_res == (Cin != Cout), "%s: residual_conv presence does not match Cin/Cout", pfx.c_str());
    if (rb.has_res) {
        NEED(w, pfx + "residual_conv.weight");
        NEED(b, pfx + "residual_conv.bias");
        std::vector<float> t;
        conv_fwd_layout(*w, t);
        rb.Rf = pk.add(t);
        rb.tcRf = pk16.add(t, 1, Cin, Cout, rb.Rf);
        conv_bwd_layout(*w, t, false);
        rb.Rb = pk.add(t);
        rb.tcRb = pk16.add(t, 1, Cout, Cin, rb.Rb);
        rb.Rbias = pk.add(b->v);
    }
    const int tdim = m->arch.dim + m->arch.wall_embed_dim;
    if (m->arch.time_mlp_config == 3) {
        NEED(w0, pfx + "time_mlp.0.weight");
        NEED(b0, pfx + "time_mlp.0.bias");
        NEED(w2, pfx + "time_mlp.2.weight");
        NEED(b2, pfx + "time_mlp.2.bias");
        PMP_REQUIRE(w0->shape[0] == 2 * tdim && w0->shape[1] == tdim && w2->shape[0] == Cout && w2->shape[1] == 2 * tdim,
                    "%stime_mlp shapes unexpected", pfx.c_str());
        rb.tm0W = pk.add(w0->v);
        rb.tm0B = pk.add(b0->v);
        rb.tmW = pk.add(w2->v);
        rb.tmB = pk.add(b2->v);
    } else {
        NEED(w1, pfx + "time_mlp.1.weight");
        NEED(b1, pfx + "time_mlp.1.bias");
        PMP_REQUIRE(w1->shape[0] == Cout && w1->shape[1] == tdim, "%stime_mlp.1.weight must be [%d,%d]", pfx.c_str(), Cout, tdim);
        rb.tmW = pk.add(w1->v);
        rb.tmB = pk.add(b1->v);
    }
    return 0;
}

// time-embedding tables from the packed arena: temb[t] and the t-only part of every block embedding
static int build_time_tables(pmp_model* m, cudaStream_t st) {
    const pmp_arch& a = m->arch;
    int rc;
    const int T = a.n_timesteps;
    PMP_REQUIRE(T >= 1 && T <= 100000, "n_timesteps out of range");
    if (!m->t_iota) {
        PMP_CUDA_OK(cudaMalloc((void**)&m->t_iota, T * sizeof(int64_t)));
        std::vector<int64_t> th(T);
        for (int i = 0; i < T; ++i) th[i] = i;
        PMP_CUDA_OK(cudaMemcpy(m->t_iota, th.data(), T * sizeof(int64_t), cudaMemcpyHostToDevice));
    }
    int64_t* tdev = m->t_iota;
    const int tdim = a.dim + a.wall_embed_dim;
    const int nblk = (int)m->blocks.size();
    if (!m->temb_tab) PMP_CUDA_OK(cudaMalloc((void**)&m->temb_tab, (size_t)T * a.dim * sizeof(float)));
    const float neg_k = (float)(-(log(10000.0) / (double)(a.dim / 2 - 1)));
    if ((rc = launch_time_mlp(m->A(m->tW1), m->A(m->tB1), m->A(m->tW3), m->A(m->tB3), a.dim, a.act, neg_k, tdev, T,
                              m->temb_tab, st)))
        return rc;
    if (a.time_mlp_config == 0) {
        if (!m->embT) PMP_CUDA_OK(cudaMalloc((void**)&m->embT, (size_t)T * m->Ctot * sizeof(float)));
        for (auto& b : m->blocks)
            if ((rc = launch_small_linear(m->temb_tab, a.dim, a.dim, m->A(b.tmW), tdim, 0, m->A(b.tmB), 1, a.act,
                                          m->embT + b.eoff, m->Ctot, b.Cout, T, 0, st)))
                return rc;
    } else {
        if (!m->embT) PMP_CUDA_OK(cudaMalloc((void**)&m->embT, (size_t)T * nblk * m->hid * sizeof(float)));
        for (int i = 0; i < nblk; ++i)
            if ((rc = launch_small_linear(m->temb_tab, a.dim, a.dim, m->A(m->blocks[i].tm0W), tdim, 0,
                                          m->A(m->blocks[i].tm0B), 0, a.act, m->embT + (size_t)i * m->hid,
                                          nblk * m->hid, m->hid, T, 0, st)))
                return rc;
    }
    return 0;
}

// lays every weight out for the kernels (offsets are stored in the model; same offsets on every call)
static int pack_all(pmp_model* m, Packer& pk, Packer16& pk16, size_t (&v_top)[11], size_t (&v_l)[8][9]) {
    const pmp_arch& a = m->arch;
    PMP_REQUIRE(a.n_mults >= 2 && a.n_mults <= PMP_MAX_LEVELS, "n_mults must be 2..%d", PMP_MAX_LEVELS);
    PMP_REQUIRE(a.time_mlp_config == 0 || a.time_mlp_config == 3, "time_mlp_config %d is not built (0 or 3)", a.time_mlp_config);
    PMP_REQUIRE(a.wall_embed_dim == 64 && a.dim == 64, "only dim = wall_embed_dim = 64 is built");
    m->nl = a.n_mults;
    m->dims.assign(1, a.transition_dim);
    for (int i = 0; i < a.n_mults; ++i) m->dims.push_back(a.dim * a.dim_mults[i]);
    int rc;
    m->blocks.clear();
    char nm[128];
    for (int L = 0; L < m->nl; ++L) {
        for (int j = 0; j < 2; ++j) {
            snprintf(nm, sizeof nm, "downs.%d.%d.", L, j);
            ResB rb;
            if ((rc = pack_res(m, pk, pk16, nm, j == 0 ? m->dims[L] : m->dims[L + 1], m->dims[L + 1], rb))) return rc;
            m->blocks.push_back(rb);
        }
    }
    {
        ResB rb1, rb2;
        if ((rc = pack_res(m, pk, pk16, "mid_block1.", m->dims[m->nl], m->dims[m->nl], rb1))) return rc;
        if ((rc = pack_res(m, pk, pk16, "mid_block2.", m->dims[m->nl], m->dims[m->nl], rb2))) return rc;
        m->blocks.push_back(rb1);
        m->blocks.push_back(rb2);
    }
    for (int ind = 0; ind < m->nl - 1; ++ind) {
        const int Ls = m->nl - 1 - ind;            // skip level consumed
        const int Cs = m->dims[Ls + 1], Co = m->dims[Ls];
        for (int j = 0; j < 2; ++j) {
            snprintf(nm, sizeof nm, "ups.%d.%d.", ind, j);
            ResB rb;
            if ((rc = pack_res(m, pk, pk16, nm, j == 0 ? 2 * Cs : Co, Co, rb))) return rc;
            m->blocks.push_back(rb);
        }
    }
    int off = 0;
    for (auto& b : m->blocks) {
        b.eoff = off;
        off += b.Cout;
    }
    m->Ctot = off;
    m->hid = 2 * (a.dim + a.wall_embed_dim);
    // down / up sampling convs
    m->downc.assign(m->nl, PlainConv());
    m->upc.assign(m->nl, PlainConv());
    for (int L = 0; L < m->nl; ++L) {
        if (!m->has_down(L)) continue;
        snprintf(nm, sizeof nm, "downs.%d.2.conv.", L);
        NEED(w, std::string(nm) + "weight");
        NEED(b, std::string(nm) + "bias");
        PMP_REQUIRE(w->shape.size() == 3 && w->shape[2] == 3 && w->shape[0] == m->dims[L + 1], "%sweight shape", nm);
        PlainConv& pc = m->downc[L];
        pc.present = true; pc.C = m->dims[L + 1]; pc.KS = 3;
        std::vector<float> t;
        conv_fwd_layout(*w, t); pc.Wf = pk.add(t); pc.tcf = pk16.add(t, 3, pc.C, pc.C, pc.Wf);
        conv_bwd_layout(*w, t, false); pc.Wb = pk.add(t); pc.tcb = pk16.add(t, 3, pc.C, pc.C, pc.Wb);
        pc.bias = pk.add(b->v);
    }
    for (int ind = 0; ind < m->nl - 1; ++ind) {
        if (!m->has_up(ind)) continue;
        snprintf(nm, sizeof nm, "ups.%d.2.conv.", ind);
        NEED(w, std::string(nm) + "weight");
        NEED(b, std::string(nm) + "bias");
        const int C = m->dims[m->nl - 1 - ind];
        PMP_REQUIRE(w->shape.size() == 3 && w->shape[2] == 4 && w->shape[0] == C, "%sweight shape", nm);
        PlainConv& pc = m->upc[ind];
        pc.present = true; pc.C = C; pc.KS = 4;
        std::vector<float> f, bw;
        convT_layouts(*w, f, bw);
        pc.Wf = pk.add(f); pc.Wb = pk.add(bw);
        pc.tcf = pk16.add(f, 4, C, C, pc.Wf); pc.tcb = pk16.add(bw, 4, C, C, pc.Wb);
        pc.bias = pk.add(b->v);
    }
    if ((rc = pack_cb(m, pk, pk16, "final_conv.0.", m->finalcb))) return rc;
    {
        NEED(w, "final_conv.1.weight");
        NEED(b, "final_conv.1.bias");
        PMP_REQUIRE(w->shape[0] == a.transition_dim && w->shape[1] == a.dim, "final_conv.1.weight shape");
        std::vector<float> t;
        conv_fwd_layout(*w, t); m->fin1Wf = pk.add(t);
        m->fin1tc = pk16.add(t, 1, a.dim, a.transition_dim, m->fin1Wf);
        conv_bwd_layout(*w, t, false); m->fin1Wb = pk.add(t);
        m->fin1b = pk.add(b->v);
    }
    {
        NEED(w1, "time_mlp.1.weight"); NEED(b1, "time_mlp.1.bias");
        NEED(w3, "time_mlp.3.weight"); NEED(b3, "time_mlp.3.bias");
        m->tW1 = pk.add(w1->v); m->tB1 = pk.add(b1->v); m->tW3 = pk.add(w3->v); m->tB3 = pk.add(b3->v);
    }
    // ---- ViT
    size_t &v_ln0w = v_top[0], &v_ln0b = v_top[1], &v_l1w = v_top[2], &v_l1b = v_top[3], &v_l2w = v_top[4], &v_l2b = v_top[5],
           &v_l3w = v_top[6], &v_l3b = v_top[7], &v_ln1w = v_top[8], &v_ln1b = v_top[9], &v_cls = v_top[10];
    v_ln0w = v_ln0b = 0;
    {
        const std::string p = "wallLoc_encoder.";
        int i = 1;
        if (a.vit_octaves) {
            NEED(g, p + "to_patch_embedding.2.weight"); NEED(b, p + "to_patch_embedding.2.bias");
            v_ln0w = pk.add(g->v); v_ln0b = pk.add(b->v);
            i = 3;
        }
        snprintf(nm, sizeof nm, "to_patch_embedding.%d.", i);
        NEED(l1w, p + nm + "weight"); NEED(l1b, p + nm + "bias");
        snprintf(nm, sizeof nm, "to_patch_embedding.%d.", i + 1);
        NEED(l2w, p + nm + "weight"); NEED(l2b, p + nm + "bias");
        snprintf(nm, sizeof nm, "to_patch_embedding.%d.", i + 3);
        NEED(l3w, p + nm + "weight"); NEED(l3b, p + nm + "bias");
        snprintf(nm, sizeof nm, "to_patch_embedding.%d.", i + 4);
        NEED(n1w, p + nm + "weight"); NEED(n1b, p + nm + "bias");
        NEED(cls, p + "cls_token");
        const int fdim = a.vit_octaves ? a.vit_token_dim * a.vit_octaves * 2 : a.vit_token_dim;
        PMP_REQUIRE(l1w->shape[1] == fdim && l1w->shape[0] == 64, "wall encoder first Linear must be [64,%d]", fdim);
        v_l1w = pk.add(l1w->v); v_l1b = pk.add(l1b->v); v_l2w = pk.add(l2w->v); v_l2b = pk.add(l2b->v);
        v_l3w = pk.add(l3w->v); v_l3b = pk.add(l3b->v); v_ln1w = pk.add(n1w->v); v_ln1b = pk.add(n1b->v);
        v_cls = pk.add(cls->v);
        PMP_REQUIRE(a.vit_depth >= 1 && a.vit_depth <= 8, "vit depth");
        for (int l = 0; l < a.vit_depth; ++l) {
            snprintf(nm, sizeof nm, "transformer.layers.%d.", l);
            const std::string q = p + nm;
            NEED(anw, q + "0.norm.weight"); NEED(anb, q + "0.norm.bias"); NEED(qkv, q + "0.fn.to_qkv.weight");
            NEED(fnw, q + "1.norm.weight"); NEED(fnb, q + "1.norm.bias");
            NEED(f0w, q + "1.fn.net.0.weight"); NEED(f0b, q + "1.fn.net.0.bias");
            NEED(f3w, q + "1.fn.net.3.weight"); NEED(f3b, q + "1.fn.net.3.bias");
            if (find(m, q + "0.fn.to_out.0.weight")) {
                set_error("wall encoder with an attention output projection is not built");
                return PMP_ERR_ARG;
            }
            v_l[l][0] = pk.add(anw->v); v_l[l][1] = pk.add(anb->v); v_l[l][2] = pk.add(qkv->v);
            v_l[l][3] = pk.add(fnw->v); v_l[l][4] = pk.add(fnb->v); v_l[l][5] = pk.add(f0w->v);
            v_l[l][6] = pk.add(f0b->v); v_l[l][7] = pk.add(f3w->v); v_l[l][8] = pk.add(f3b->v);
        }
    }
    return 0;
}

static int finalize_model(pmp_model* m) {
    const pmp_arch& a = m->arch;
    Packer pk;
    Packer16 pk16;
    size_t v_top[11], v_l[8][9];
    int rc;
    char nm[128];
    (void)nm;
    if ((rc = pack_all(m, pk, pk16, v_top, v_l))) return rc;
    const size_t v_ln0w = v_top[0], v_ln0b = v_top[1], v_l1w = v_top[2], v_l1b = v_top[3], v_l2w = v_top[4], v_l2b = v_top[5],
                 v_l3w = v_top[6], v_l3b = v_top[7], v_ln1w = v_top[8], v_ln1b = v_top[9], v_cls = v_top[10];
    m->plane_recs = pk16.recs;
    m->arena16_n = pk16.buf.size();
    // ---- upload
    if (m->arena) cudaFree(m->arena);
    m->arena = nullptr;
    m->arena_n = pk.buf.size();
    PMP_CUDA_OK(cudaMalloc((void**)&m->arena, m->arena_n * sizeof(float)));
    PMP_CUDA_OK(cudaMemcpy(m->arena, pk.buf.data(), m->arena_n * sizeof(float), cudaMemcpyHostToDevice));
    if (m->arena16) cudaFree(m->arena16);
    m->arena16 = nullptr;
    if (!pk16.buf.empty()) {
        PMP_CUDA_OK(cudaMalloc((void**)&m->arena16, pk16.buf.size() * sizeof(uint16_t)));
        PMP_CUDA_OK(cudaMemcpy(m->arena16, pk16.buf.data(), pk16.buf.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    }
    VitW& v = m->vit;
    memset(&v, 0, sizeof v);
    v.token_dim = a.vit_token_dim; v.octaves = a.vit_octaves; v.depth = a.vit_depth; v.dim = 64;
    v.mlp = 64 * a.vit_mlp_ratio; v.act = a.vit_act; v.norm_factor = a.vit_norm_factor;
    if (a.vit_octaves) { v.ln0_w = m->A(v_ln0w); v.ln0_b = m->A(v_ln0b); }
    v.l1_w = m->A(v_l1w); v.l1_b = m->A(v_l1b); v.l2_w = m->A(v_l2w); v.l2_b = m->A(v_l2b);
    v.l3_w = m->A(v_l3w); v.l3_b = m->A(v_l3b); v.ln1_w = m->A(v_ln1w); v.ln1_b = m->A(v_ln1b);
    v.cls = m->A(v_cls);
    for (int l = 0; l < a.vit_depth; ++l) {
        v.an_w[l] = m->A(v_l[l][0]); v.an_b[l] = m->A(v_l[l][1]); v.qkv[l] = m->A(v_l[l][2]);
        v.fn_w[l] = m->A(v_l[l][3]); v.fn_b[l] = m->A(v_l[l][4]); v.f0_w[l] = m->A(v_l[l][5]);
        v.f0_b[l] = m->A(v_l[l][6]); v.f3_w[l] = m->A(v_l[l][7]); v.f3_b[l] = m->A(v_l[l][8]);
    }
    if (m->temb_tab) { cudaFree(m->temb_tab); m->temb_tab = nullptr; }
    if (m->embT) { cudaFree(m->embT); m->embT = nullptr; }
    if (m->t_iota) { cudaFree(m->t_iota); m->t_iota = nullptr; }
    cudaStream_t st = 0;
    if ((rc = build_time_tables(m, st))) return rc;
    PMP_CUDA_OK(cudaStreamSynchronize(st));
    m->finalized = true;
    ++m->gen;
    return 0;
}

// ------------------------------------------------------------------ conv geometry helpers
static void finish_geom(ConvP& p) {
    p.S = 128 / p.HJ;
    if (p.S < 1) p.S = 1;
    p.HP = p.Hin + 4;
    while (p.S > 1 && p.S * p.HP > 320) --p.S;
    int ar = p.S * p.HP;
    while ((ar & 7) != 2) ++ar;
    p.arows = ar;
}
static ConvP geom_s1(int R, int Hl, int KS) {
    ConvP p;
    memset(&p, 0, sizeof p);
    p.R = R; p.Hout = Hl; p.Hin = Hl; p.HJ = Hl; p.ostride = 1; p.istride = 1; p.nphase = 1;
    p.ntaps[0] = KS;
    for (int k = 0; k < KS; ++k) { p.tap_k[0][k] = k; p.tap_shift[0][k] = k - KS / 2; }
    finish_geom(p);
    return p;
}
// stride-2 gather: hi = 2*ho - 1 + k  (Downsample1d forward KS=3; Upsample1d backward KS=4)
static ConvP geom_s2(int R, int Hout, int KS) {
    ConvP p;
    memset(&p, 0, sizeof p);
    p.R = R; p.Hout = Hout; p.Hin = 2 * Hout; p.HJ = Hout; p.ostride = 1; p.istride = 2; p.nphase = 1;
    p.ntaps[0] = KS;
    for (int k = 0; k < KS; ++k) { p.tap_k[0][k] = k; p.tap_shift[0][k] = k - 1; }
    finish_geom(p);
    return p;
}
// stride-2 scatter as two output phases: ho = 2j+phase, hi = (ho + 1 - k)/2
// (Upsample1d forward KS=4; Downsample1d backward KS=3)
static ConvP geom_t2(int R, int Hin, int KS) {
    ConvP p;
    memset(&p, 0, sizeof p);
    p.R = R; p.Hout = 2 * Hin; p.Hin = Hin; p.HJ = Hin; p.ostride = 2; p.istride = 1; p.nphase = 2;
    if (KS == 4) {
        p.ntaps[0] = 2; p.tap_k[0][0] = 1; p.tap_shift[0][0] = 0; p.tap_k[0][1] = 3; p.tap_shift[0][1] = -1;
        p.ntaps[1] = 2; p.tap_k[1][0] = 0; p.tap_shift[1][0] = 1; p.tap_k[1][1] = 2; p.tap_shift[1][1] = 0;
    } else {
        p.ntaps[0] = 1; p.tap_k[0][0] = 1; p.tap_shift[0][0] = 0;
        p.ntaps[1] = 2; p.tap_k[1][0] = 0; p.tap_shift[1][0] = 1; p.tap_k[1][1] = 2; p.tap_shift[1][1] = 0;
    }
    finish_geom(p);
    return p;
}

// ------------------------------------------------------------------ one eps evaluation
typedef __nv_bfloat16 bf16;

// forward activation view with its operand planes
struct ActP : Act {
    bf16* hi = nullptr;
    bf16* lo = nullptr;
};
struct GradP {
    float* raw = nullptr;
    bf16 *raw_hi = nullptr, *raw_lo = nullptr;
    int ldr = 0;
    float* gy = nullptr;   // dense [R*Hl][C]
    bf16 *gy_hi = nullptr, *gy_lo = nullptr;
};
// what a training step keeps of the forward (F) and backward-data (B1) passes of one eps evaluation: every tensor the
// tangent pass and the second reverse pass read again (train_host.inc)
struct TapeRes {
    ActP x, mid;              // block input, CB0 output + embedding (input of CB1)
    GradP gout;               // B1 gradient w.r.t. the block output (raw) and w.r.t. CB1's conv output (gy)
    float* gmid = nullptr;    // B1 gradient w.r.t. mid (raw)
    float* gy0 = nullptr;     // B1 gradient w.r.t. CB0's conv output
    bf16 *gy0_hi = nullptr, *gy0_lo = nullptr;
};
struct TapePlain {
    ActP in;                  // input of the down / up sampling conv
    GradP gout;               // B1 gradient w.r.t. its output (raw)
};
struct Tape {
    std::vector<TapeRes> res;
    std::vector<TapePlain> down, up;
    std::vector<int> Hl;
    ActP fin_in;                         // input of final_conv.0
    float* f = nullptr; bf16 *f_hi = nullptr, *f_lo = nullptr;   // its activation output
    float* u = nullptr;                  // U-Net output
    float* gf = nullptr;                 // B1 gradient w.r.t. f (raw)
    float* gyf = nullptr; bf16 *gyf_hi = nullptr, *gyf_lo = nullptr;   // B1 gradient w.r.t. final_conv.0's conv output
};

struct Ctx {
    pmp_model* m;
    cudaStream_t st;
    bool dry;          // only count workspace
    size_t used = 0;   // floats
    int R, H, n_cond;
    const int64_t* t;
    const float* embRow;   // cfg0: embR rows of this chunk; cfg3: E rows of this chunk
    int x_row0 = 0, x_wrap = 0;   // the input holds x_wrap rows shared by both CFG halves: row r reads (x_row0 + r) mod x_wrap
    int rc = 0;
    Tape* tape = nullptr;   // training step: keep an fp32 copy of every activation / gradient and record where they are
    bool train() const { return tape != nullptr; }
    bool tc() const { return m->gemm_path >= 1; }
    float* alloc(size_t n) {
        n = (n + 3) & ~(size_t)3;
        float* p = dry ? nullptr : m->ws + used;
        used += n;
        return p;
    }
    // BF16x3 operand planes for an fp32 tensor of n elements (tensor-core path only)
    void planes(size_t n, bf16** hi, bf16** lo) {
        *hi = *lo = nullptr;
        if (!tc()) return;
        float* p = alloc(n);   // 2 planes x n x 2 bytes = n floats
        if (!dry) {
            *hi = reinterpret_cast<bf16*>(p);
            *lo = *hi + ((n + 7) & ~(size_t)7);
        }
        used += 8;             // slack for the 16-byte rounding of the second plane
    }
    void conv(const ConvP& p) {
        if (dry || rc) return;
        if (m->gemm_path == 2 && conv_tc2_eligible(p)) {
            rc = prof_launch_conv(p, st, 1, launch_conv_tc2);
            return;
        }
        if (tc() && conv_tc_eligible(p)) {
            rc = prof_launch_conv(p, st, 1, m->gemm_path == 4 ? launch_conv_tc_pair16 : (m->gemm_path == 3 ? launch_conv_tc_pair : launch_conv_tc));
            return;
        }
        if ((!p.A1 && !p.inl_x) || (p.A2_hi && !p.A2) || (p.ident_hi && !p.ident)) {
            // planes-only tensors are produced only where every consumer runs on the tensor-core kernel
            set_error("internal: op (N=%d C1=%d H=%d) has BF16-plane-only operands but does not fit the tcgen05 kernel", p.N,
                      p.C1, p.Hout);
            rc = PMP_ERR_ARG;
            return;
        }
        rc = prof_launch_conv(p, st, 0, launch_conv_simt);
    }
    void dbg(const char* kind, const std::string& name, float* p, int64_t n) {
        if (!dry && p) m->dbg[std::string(kind) + name] = std::make_pair(p, n);
    }
};

// tensor consumed only as the A operand of a tensor-core conv: BF16 planes suffice (no fp32 copy)
static bool planes_only(const Ctx& c, int C) { return !c.train() && c.tc() && (C % 64) == 0 && c.H <= 128; }
// on gemm paths 1 and 3 the epilogue also takes residual terms as planes, so no activation or gradient between
// 64-multiple-channel convs needs an fp32 copy at all (path 2's kernel still reads fp32 residuals)
static bool planes_all(const Ctx& c, int C) { return planes_only(c, C) && c.m->gemm_path != 2 && !c.m->keep_fp32; }

static ActP new_act(Ctx& c, int C, int Hl) {
    ActP a;
    a.p = planes_all(c, C) ? nullptr : c.alloc((size_t)c.R * Hl * C);
    c.planes((size_t)c.R * Hl * C, &a.hi, &a.lo);
    a.ld = C; a.C = C; a.Hl = Hl; a.blk = -1;
    return a;
}

static ActP view_act(const ActP& base, int off, int C) {
    ActP a = base;
    a.p = base.p ? base.p + off : nullptr;
    a.hi = base.hi ? base.hi + off : nullptr;
    a.lo = base.lo ? base.lo + off : nullptr;
    a.C = C;
    return a;
}

static void set_emb(Ctx& c, ConvP& p, const ResB& b) {
    const pmp_model* m = c.m;
    if (m->arch.time_mlp_config == 0 && !c.train()) {
        p.embT = m->embT; p.tidx = c.t; p.ldeT = m->Ctot; p.tmax = m->arch.n_timesteps - 1;
        p.embR = c.embRow; p.ldeR = m->Ctot; p.n_cond = c.n_cond;
    } else {
        p.embR = c.embRow; p.ldeR = m->Ctot; p.n_cond = c.R;
    }
    p.eoff = b.eoff;
}
static void set_A1(ConvP& p, const float* a, const bf16* hi, const bf16* lo, int ld, int C) {
    p.A1 = a; p.A1_hi = hi; p.A1_lo = lo; p.lda1 = ld; p.C1 = C;
}
static void set_A2(ConvP& p, const float* a, const bf16* hi, const bf16* lo, int ld, int C) {
    p.A2 = a; p.A2_hi = hi; p.A2_lo = lo; p.lda2 = ld; p.C2 = C;
}
static void set_W1(const pmp_model* m, ConvP& p, size_t off, const TcW& t) {
    p.W1 = m->A(off);
    if (t.ok && m->arena16) { p.W1_hi = m->arena16 + t.hi; p.W1_lo = m->arena16 + t.lo; }
}
static void set_W2(const pmp_model* m, ConvP& p, size_t off, const TcW& t) {
    p.W2 = m->A(off);
    if (t.ok && m->arena16) { p.W2_hi = m->arena16 + t.hi; p.W2_lo = m->arena16 + t.lo; }
}

// out = ResidualTemporalBlock_dd(x)   (temporal_dd.py:71-74); `out` is pre-allocated (maybe a concat view)
static void res_fwd(Ctx& c, int bi, const ActP& x, ActP& out) {
    pmp_model* m = c.m;
    ResB& b = m->blocks[bi];
    const int R = c.R, Hl = x.Hl, Co = b.Cout;
    for (int j = 0; j < 2; ++j) {
        b.cb[j].yhat = c.alloc((size_t)R * Hl * Co);
        b.cb[j].rstd = c.alloc((size_t)R * 8);
        c.dbg("yhat:", b.pfx + (j ? "blocks.1." : "blocks.0."), b.cb[j].yhat, (int64_t)R * Hl * Co);
    }
    ActP mid = new_act(c, Co, Hl);
    if (planes_only(c, Co)) mid.p = nullptr;   // only read by the second conv of this block
    ConvP p = geom_s1(R, Hl, 5);
    set_A1(p, x.p, x.hi, x.lo, x.ld, b.Cin);
    set_W1(m, p, b.cb[0].Wf, b.cb[0].tcf);
    p.N = Co;
    p.bias = m->A(b.cb[0].bias); p.gn_fwd = 1; p.gamma = m->A(b.cb[0].gamma); p.beta = m->A(b.cb[0].beta);
    p.yhat = b.cb[0].yhat; p.ldy = Co; p.rstd = b.cb[0].rstd; p.act = m->arch.act;
    set_emb(c, p, b);
    p.out = mid.p; p.ldo = Co; p.out_hi = mid.hi; p.out_lo = mid.lo;
    if (b.Cin <= 16) {   // first block: K = 5*D, evaluated inline in the tensor kernel's epilogue
        p.inl_x = x.p; p.inl_ld = x.ld; p.inl_D = b.Cin; p.inl_taps = 5; p.inl_mode = 1; p.inl_w = m->A(b.cb[0].Wf);
        p.inl_row0 = c.x_row0; p.inl_wrap = c.x_wrap;
    }
    c.conv(p);
    ConvP q = geom_s1(R, Hl, 5);
    set_A1(q, mid.p, mid.hi, mid.lo, Co, Co);
    set_W1(m, q, b.cb[1].Wf, b.cb[1].tcf);
    q.N = Co;
    q.bias = m->A(b.cb[1].bias); q.gn_fwd = 1; q.gamma = m->A(b.cb[1].gamma); q.beta = m->A(b.cb[1].beta);
    q.yhat = b.cb[1].yhat; q.ldy = Co; q.rstd = b.cb[1].rstd; q.act = m->arch.act;
    if (b.has_res) {
        set_A2(q, x.p, x.hi, x.lo, x.ld, b.Cin);
        set_W2(m, q, b.Rf, b.tcRf);
        q.bias2 = m->A(b.Rbias);
        if (b.Cin <= 16) {
            q.inl_x = x.p; q.inl_ld = x.ld; q.inl_D = b.Cin; q.inl_taps = 1; q.inl_mode = 2; q.inl_w = m->A(b.Rf);
            q.inl_row0 = c.x_row0; q.inl_wrap = c.x_wrap;
        }
    } else {
        q.ident = x.p; q.ldi = x.ld;
        if (!x.p) { q.ident_hi = x.hi; q.ident_lo = x.lo; }
    }
    q.out = out.p; q.ldo = out.ld; q.out_hi = out.hi; q.out_lo = out.lo;
    c.conv(q);
    out.C = Co; out.Hl = Hl; out.blk = bi;
    if (c.train()) { c.tape->res[bi].x = x; c.tape->res[bi].mid = mid; }
    c.dbg("out:", b.pfx, out.p, out.ld == Co ? (int64_t)R * Hl * Co : 0);
}

// standalone GroupNorm backward through the CB1 of the block that produced `x`
static void gn_bwd_of(Ctx& c, const Act& x, const float* raw, int ldr, float* gy, bf16* gy_hi, bf16* gy_lo) {
    if (c.dry || c.rc) return;
    pmp_model* m = c.m;
    const CBW& cb = m->blocks[x.blk].cb[1];
    c.rc = launch_gn_bwd(raw, ldr, cb.yhat, x.C, cb.rstd, m->A(cb.gamma), m->A(cb.beta), m->arch.act, gy, x.C, gy_hi,
                         gy_lo, c.R, x.Hl, x.C, c.st);
}

// launches a backward conv `p` whose result is the gradient w.r.t. activation `x`;
// emits raw (+addend) and, when `x` came out of a residual block, the GroupNorm-backward'ed
// gradient for that block's second conv (fused when one tile holds whole slabs)
static GradP emit_grad(Ctx& c, ConvP& p, const Act& x, const float* addend, int ldad, float* raw_dst, int ld_dst,
                       bool fusable) {
    GradP g;
    const size_t n = (size_t)c.R * x.Hl * x.C;
    if (raw_dst) {
        g.raw = raw_dst; g.ldr = ld_dst;
    } else {
        // fp32 copy of the raw gradient only for the standalone GroupNorm backward (non-fusable producers, concat
        // gradients, which also feed the skip addend); residual / operand consumers read the planes
        const bool need_f32 = !planes_all(c, x.C) || x.grad_f32 || (x.blk >= 0 && !(fusable && p.N == x.C));
        g.raw = need_f32 ? c.alloc(n) : nullptr; g.ldr = x.C;
        c.planes(n, &g.raw_hi, &g.raw_lo);
    }
    p.out = g.raw; p.ldo = g.ldr; p.out_hi = g.raw_hi; p.out_lo = g.raw_lo;
    p.addend = addend; p.ldad = ldad;
    p.act = c.m->arch.act;
    if (x.blk >= 0) {
        g.gy = planes_only(c, x.C) ? nullptr : c.alloc(n);      // only read by the next dgrad conv
        c.planes(n, &g.gy_hi, &g.gy_lo);
        if (fusable && p.N == x.C) {
            const CBW& cb = c.m->blocks[x.blk].cb[1];
            p.gn_bwd = 1; p.gamma = c.m->A(cb.gamma); p.beta = c.m->A(cb.beta);
            p.yhat = cb.yhat; p.ldy = x.C; p.rstd = cb.rstd; p.gy = g.gy; p.ldg = x.C;
            p.gy_hi = g.gy_hi; p.gy_lo = g.gy_lo;
            c.conv(p);
        } else {
            c.conv(p);
            gn_bwd_of(c, x, g.raw, g.ldr, g.gy, g.gy_hi, g.gy_lo);
        }
    } else {
        c.conv(p);
    }
    return g;
}

// gradient through a residual block: gout = dE/d(out) -> dE/d(x)   
static GradP res_bwd(Ctx& c, int bi, const GradP& gout, const Act& x, const float* addend, int ldad, float* raw_dst,
                     int ld_dst) {
    pmp_model* m = c.m;
    ResB& b = m->blocks[bi];
    const int R = c.R, Hl = x.Hl, Co = b.Cout;
    const size_t n = (size_t)R * Hl * Co;
    float* gy0 = planes_only(c, Co) ? nullptr : c.alloc(n);   // only read by the dgrad conv below
    bf16 *gy0_hi, *gy0_lo;
    c.planes(n, &gy0_hi, &gy0_lo);
    ConvP p = geom_s1(R, Hl, 5);
    set_A1(p, gout.gy, gout.gy_hi, gout.gy_lo, Co, Co);
    set_W1(m, p, b.cb[1].Wb, b.cb[1].tcb);
    p.N = Co;
    p.gn_bwd = 1; p.gamma = m->A(b.cb[0].gamma); p.beta = m->A(b.cb[0].beta);
    p.yhat = b.cb[0].yhat; p.ldy = Co; p.rstd = b.cb[0].rstd; p.gy = gy0; p.ldg = Co; p.act = m->arch.act;
    p.gy_hi = gy0_hi; p.gy_lo = gy0_lo;
    if (c.train()) {
        TapeRes& tr = c.tape->res[bi];
        tr.gout = gout; tr.gy0 = gy0; tr.gy0_hi = gy0_hi; tr.gy0_lo = gy0_lo;
        tr.gmid = c.alloc(n);
        p.out = tr.gmid; p.ldo = Co;
    }
    c.conv(p);
    ConvP q = geom_s1(R, Hl, 5);
    set_A1(q, gy0, gy0_hi, gy0_lo, Co, Co);
    set_W1(m, q, b.cb[0].Wb, b.cb[0].tcb);
    q.N = b.Cin;
    if (b.has_res) {
        set_A2(q, gout.raw, gout.raw_hi, gout.raw_lo, gout.ldr, Co);
        set_W2(m, q, b.Rb, b.tcRb);
    } else {
        q.ident = gout.raw; q.ldi = gout.ldr;
        if (!gout.raw) { q.ident_hi = gout.raw_hi; q.ident_lo = gout.raw_lo; }
    }
    GradP g = emit_grad(c, q, x, addend, ldad, raw_dst, ld_dst, true);
    c.dbg("gx:", b.pfx, g.raw, g.ldr == x.C ? (int64_t)R * Hl * x.C : 0);
    return g;
}

static int run_unet(Ctx& c, const float* x_in, float* eps_out, float* energy, float* u_out) {
    pmp_model* m = c.m;
    const int R = c.R, H = c.H, nl = m->nl, D = m->arch.transition_dim;
    std::vector<int> Hl(nl);
    Hl[0] = H;
    for (int L = 1; L < nl; ++L) Hl[L] = m->has_down(L - 1) ? Hl[L - 1] / 2 : Hl[L - 1];
    if (c.train()) {
        *c.tape = Tape();
        c.tape->res.resize(m->blocks.size());
        c.tape->down.resize(nl);
        c.tape->up.resize(nl);
        c.tape->Hl = Hl;
    }
    // concat buffers for skip levels 1..nl-1: [R][Hl][2*C]; level 0 skip is never consumed
    std::vector<ActP> cat(nl);
    for (int L = 1; L < nl; ++L) cat[L] = new_act(c, 2 * m->dims[L + 1], Hl[L]);
    std::vector<ActP> lvl_in(nl), lvl_a(nl), lvl_b(nl);
    ActP cur;
    cur.p = const_cast<float*>(x_in); cur.ld = D; cur.C = D; cur.Hl = H; cur.blk = -1;
    for (int L = 0; L < nl; ++L) {
        const int C = m->dims[L + 1];
        lvl_in[L] = cur;
        ActP a = new_act(c, C, Hl[L]);
        res_fwd(c, 2 * L, cur, a);
        ActP b = (L >= 1) ? view_act(cat[L], C, C) : new_act(c, C, Hl[L]);
        res_fwd(c, 2 * L + 1, a, b);
        lvl_a[L] = a; lvl_b[L] = b;
        if (m->has_down(L)) {
            ActP d = new_act(c, C, Hl[L] / 2);
            ConvP p = geom_s2(R, Hl[L] / 2, 3);
            set_A1(p, b.p, b.hi, b.lo, b.ld, C);
            set_W1(m, p, m->downc[L].Wf, m->downc[L].tcf);
            p.N = C; p.bias = m->A(m->downc[L].bias);
            p.out = d.p; p.ldo = d.ld; p.out_hi = d.hi; p.out_lo = d.lo;
            c.conv(p);
            if (c.train()) c.tape->down[L].in = b;
            cur = d;
        } else {
            cur = b;
        }
    }
    const int bmid = 2 * nl;
    const int Cm = m->dims[nl];
    ActP mid_in = cur;
    ActP m1 = new_act(c, Cm, Hl[nl - 1]);
    res_fwd(c, bmid, cur, m1);
    ActP m2 = view_act(cat[nl - 1], 0, Cm);
    res_fwd(c, bmid + 1, m1, m2);
    cur = m2;
    std::vector<ActP> up_in(nl), up_a(nl), up_b(nl), up_out(nl);
    for (int ind = 0; ind < nl - 1; ++ind) {
        const int Ls = nl - 1 - ind, Co = m->dims[Ls], h = Hl[Ls];
        up_in[ind] = cur;   // first half of the concat (view), with its producer
        ActP a = new_act(c, Co, h);
        res_fwd(c, bmid + 2 + 2 * ind, cat[Ls], a);
        const bool up = m->has_up(ind);
        const bool last = ind == nl - 2;
        ActP b = (!up && !last) ? view_act(cat[Ls - 1], 0, Co) : new_act(c, Co, h);
        res_fwd(c, bmid + 3 + 2 * ind, a, b);
        up_a[ind] = a; up_b[ind] = b;
        if (up) {
            ActP o = !last ? view_act(cat[Ls - 1], 0, Co) : new_act(c, Co, 2 * h);
            o.Hl = 2 * h; o.blk = -1;
            ConvP p = geom_t2(R, h, 4);
            set_A1(p, b.p, b.hi, b.lo, b.ld, Co);
            set_W1(m, p, m->upc[ind].Wf, m->upc[ind].tcf);
            p.N = Co; p.bias = m->A(m->upc[ind].bias);
            p.out = o.p; p.ldo = o.ld; p.out_hi = o.hi; p.out_lo = o.lo;
            c.conv(p);
            if (c.train()) c.tape->up[ind].in = b;
            cur = o;
        } else {
            cur = b;
        }
        up_out[ind] = cur;
    }
    if (!c.dry && cur.Hl != H) {
        set_error("U-Net output horizon %d != input horizon %d (down/up sampling mismatch)", cur.Hl, H);
        return PMP_ERR_ARG;
    }
    // final_conv: Conv1dBlock_dd(dim,dim,5) + Conv1d(dim, D, 1)
    const int C0 = m->dims[1];
    CBW& fcb = m->finalcb;
    fcb.yhat = c.alloc((size_t)R * H * C0);
    fcb.rstd = c.alloc((size_t)R * 8);
    c.dbg("yhat:", "final_conv.0.", fcb.yhat, (int64_t)R * H * C0);
    float* f = c.alloc((size_t)R * H * C0);
    bf16 *f_hi, *f_lo;
    c.planes((size_t)R * H * C0, &f_hi, &f_lo);
    {
        ConvP p = geom_s1(R, H, 5);
        set_A1(p, cur.p, cur.hi, cur.lo, cur.ld, C0);
        set_W1(m, p, fcb.Wf, fcb.tcf);
        p.N = C0; p.bias = m->A(fcb.bias);
        p.gn_fwd = 1; p.gamma = m->A(fcb.gamma); p.beta = m->A(fcb.beta); p.yhat = fcb.yhat; p.ldy = C0;
        p.rstd = fcb.rstd; p.act = m->arch.act; p.out = f; p.ldo = C0; p.out_hi = f_hi; p.out_lo = f_lo;
        c.conv(p);
    }
    float* u = u_out ? u_out : c.alloc((size_t)R * H * D);
    if (c.train()) { c.tape->fin_in = cur; c.tape->f = f; c.tape->f_hi = f_hi; c.tape->f_lo = f_lo; c.tape->u = u; }
    {
        if (energy && !c.dry && !c.rc) {
            cudaError_t e = cudaMemsetAsync(energy, 0, (size_t)R * sizeof(float), c.st);
            if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return PMP_ERR_CUDA; }
        }
        ConvP p = geom_s1(R, H, 1);
        set_A1(p, f, f_hi, f_lo, C0, C0);
        set_W1(m, p, m->fin1Wf, m->fin1tc);
        p.N = D; p.bias = m->A(m->fin1b);
        p.out = u; p.ldo = D; p.energy = energy;
        c.conv(p);
    }
    c.dbg("", "u", u, (int64_t)R * H * D);

    // ================= backward: seed g_u = u (E = 0.5*sum u^2) =================
    float* gyf = c.alloc((size_t)R * H * C0);
    bf16 *gyf_hi, *gyf_lo;
    c.planes((size_t)R * H * C0, &gyf_hi, &gyf_lo);
    {
        ConvP p = geom_s1(R, H, 1);
        set_A1(p, u, nullptr, nullptr, D, D);
        p.W1 = m->A(m->fin1Wb); p.N = C0;
        if (D <= 16) { p.inl_x = u; p.inl_ld = D; p.inl_D = D; p.inl_taps = 1; p.inl_mode = 1; p.inl_w = m->A(m->fin1Wb); }
        p.gn_bwd = 1; p.gamma = m->A(fcb.gamma); p.beta = m->A(fcb.beta); p.yhat = fcb.yhat; p.ldy = C0;
        p.rstd = fcb.rstd; p.gy = gyf; p.ldg = C0; p.gy_hi = gyf_hi; p.gy_lo = gyf_lo; p.act = m->arch.act;
        if (c.train()) {
            c.tape->gyf = gyf; c.tape->gyf_hi = gyf_hi; c.tape->gyf_lo = gyf_lo;
            c.tape->gf = c.alloc((size_t)R * H * C0);
            p.out = c.tape->gf; p.ldo = C0;
        }
        c.conv(p);
    }
    GradP g;
    {
        // gradient w.r.t. the input of final_conv (= up_out[nl-2])
        ConvP p = geom_s1(R, H, 5);
        set_A1(p, gyf, gyf_hi, gyf_lo, C0, C0);
        set_W1(m, p, fcb.Wb, fcb.tcb);
        p.N = C0;
        g = emit_grad(c, p, up_out[nl - 2], nullptr, 0, nullptr, 0, true);
    }
    std::vector<float*> skipg(nl, nullptr);
    std::vector<int> skipld(nl, 0);
    for (int ind = nl - 2; ind >= 0; --ind) {
        const int Ls = nl - 1 - ind, Cs = m->dims[Ls + 1], Co = m->dims[Ls], h = Hl[Ls];
        if (m->has_up(ind)) {
            if (c.train()) c.tape->up[ind].gout = g;
            // Upsample1d backward: strided conv with the same weight (App. G item 6)
            ConvP p = geom_s2(R, h, 4);
            set_A1(p, g.raw, g.raw_hi, g.raw_lo, g.ldr, Co);
            set_W1(m, p, m->upc[ind].Wb, m->upc[ind].tcb);
            p.N = Co;
            g = emit_grad(c, p, up_b[ind], nullptr, 0, nullptr, 0, true);
        }
        g = res_bwd(c, bmid + 3 + 2 * ind, g, up_a[ind], nullptr, 0, nullptr, 0);
        Act catv; catv.C = 2 * Cs; catv.Hl = h; catv.blk = -1; catv.grad_f32 = true;
        GradP gc = res_bwd(c, bmid + 2 + 2 * ind, g, catv, nullptr, 0, nullptr, 0);
        // split the concat gradient: [:Cs] continues along the up path, [Cs:] goes to the skip
        skipg[Ls] = gc.raw ? gc.raw + Cs : nullptr; skipld[Ls] = 2 * Cs;
        g = GradP();
        g.raw = gc.raw; g.raw_hi = gc.raw_hi; g.raw_lo = gc.raw_lo; g.ldr = 2 * Cs;
        if (up_in[ind].blk >= 0) {
            g.gy = planes_only(c, Cs) ? nullptr : c.alloc((size_t)R * h * Cs);
            c.planes((size_t)R * h * Cs, &g.gy_hi, &g.gy_lo);
            gn_bwd_of(c, up_in[ind], g.raw, g.ldr, g.gy, g.gy_hi, g.gy_lo);
        }
    }
    g = res_bwd(c, bmid + 1, g, m1, nullptr, 0, nullptr, 0);
    {
        // mid_block1 input == output of the last down level (also skip nl-1)
        const bool sk = nl - 1 >= 1;
        g = res_bwd(c, bmid, g, mid_in, sk ? skipg[nl - 1] : nullptr, sk ? skipld[nl - 1] : 0, nullptr, 0);
    }
    for (int L = nl - 1; L >= 0; --L) {
        g = res_bwd(c, 2 * L + 1, g, lvl_a[L], nullptr, 0, nullptr, 0);
        if (L == 0) {
            res_bwd(c, 0, g, lvl_in[0], nullptr, 0, eps_out, D);
            break;
        }
        const bool prev_down = m->has_down(L - 1);
        if (!prev_down) {
            // input of this level is the previous level's block output (also skip L-1 when L-1 >= 1)
            const bool sk = L - 1 >= 1;
            g = res_bwd(c, 2 * L, g, lvl_in[L], sk ? skipg[L - 1] : nullptr, sk ? skipld[L - 1] : 0, nullptr, 0);
        } else {
            g = res_bwd(c, 2 * L, g, lvl_in[L], nullptr, 0, nullptr, 0);
            if (c.train()) c.tape->down[L - 1].gout = g;
            // Downsample1d backward: transposed conv, output_padding 1 (App. G item 5)
            const int C = m->dims[L];
            ConvP p = geom_t2(R, Hl[L], 3);
            set_A1(p, g.raw, g.raw_hi, g.raw_lo, g.ldr, C);
            set_W1(m, p, m->downc[L - 1].Wb, m->downc[L - 1].tcb);
            p.N = C;
            const bool sk = L - 1 >= 1;
            g = emit_grad(c, p, lvl_b[L - 1], sk ? skipg[L - 1] : nullptr, sk ? skipld[L - 1] : 0, nullptr, 0, false);
        }
    }
    return c.rc;
}

static size_t wall_table_width(const pmp_model* m) {
    return m->arch.time_mlp_config == 0 ? (size_t)m->Ctot : (size_t)m->blocks.size() * m->hid;
}

// the wall-only part of every block embedding (constant over diffusion steps) for n_cond plans:
// cfg0: tab [n_cond][Ctot] = W[:,64:] act(wemb);  cfg3: tab [n_cond][nblk*hid] = W0[:,64:] wemb
static int prepare_wall_tables(pmp_model* m, const float* wemb, int n_cond, float* tab, cudaStream_t st) {
    const pmp_arch& a = m->arch;
    const int tdim = a.dim + a.wall_embed_dim;
    const int nblk = (int)m->blocks.size();
    int rc;
    if (n_cond == 0) return 0;
    if (a.time_mlp_config == 0) {
        for (auto& b : m->blocks)
            if ((rc = launch_small_linear(wemb, a.wall_embed_dim, a.wall_embed_dim, m->A(b.tmW), tdim, a.dim, nullptr, 1,
                                          a.act, tab + b.eoff, m->Ctot, b.Cout, n_cond, 0, st)))
                return rc;
    } else {
        for (int i = 0; i < nblk; ++i)
            if ((rc = launch_small_linear(wemb, a.wall_embed_dim, a.wall_embed_dim, m->A(m->blocks[i].tm0W), tdim, a.dim,
                                          nullptr, 0, a.act, tab + (size_t)i * m->hid, nblk * m->hid, m->hid, n_cond,
                                          0, st)))
                return rc;
    }
    return 0;
}

static size_t workspace_floats(pmp_model* m, int R, int H) {
    Ctx c;
    c.m = m; c.st = 0; c.dry = true; c.R = R; c.H = H; c.n_cond = 0; c.t = nullptr; c.embRow = nullptr;
    run_unet(c, nullptr, nullptr, nullptr, nullptr);
    return c.used;
}

// rows per pass: the requested chunk rounded down to a whole number of 148-tile waves of the deepest
// (most expensive) level, where one tile holds S = 128/H_deep samples and N is split in 4 tiles
static int aligned_chunk(const pmp_model* m, int H) {
    int chunk = m->row_chunk;
    int hd = H;
    for (int L = 1; L < m->nl; ++L)
        if (m->has_down(L - 1)) hd /= 2;
    if (hd < 1) hd = 1;
    int unit = (128 / hd) * 37 * (m->gemm_path >= 2 ? 2 : 1);
    if (unit & 1) unit *= 2;   // keep the CFG pair (2 rows per plan) inside one chunk
    if (chunk > unit) chunk = (chunk / unit) * unit;
    return chunk;
}

// true when the first conv of the U-Net runs inside the tensor kernel's epilogue, which can read a state shared by the
// two CFG halves (ConvP::inl_wrap); the other cores need the duplicated batch
static bool can_share_x(const pmp_model* m, int H) {
    return m->gemm_path >= 1 && m->gemm_path != 2 && m->arch.transition_dim <= 16 && H <= 128 && m->dims.size() > 1 &&
           m->dims[1] == 64;
}

// every allocation eps_rows would make for R rows, made up front (nothing may allocate while a graph is being captured)
static int eps_rows_reserve(pmp_model* m, int R, int H, cudaStream_t st) {
    PMP_REQUIRE(H % 4 == 0 && H >= 32 && H <= 128, "horizon %d unsupported (multiple of 4 in [32,128])", H);
    const int chunk = aligned_chunk(m, H);
    int rc;
    const int Rc_max = R < chunk ? R : chunk;
    const size_t need = workspace_floats(m, Rc_max, H);
    if ((rc = grow(&m->ws, &m->ws_n, need, st, &m->gen))) return rc;
    if (m->arch.time_mlp_config != 0) {
        const int nblk = (int)m->blocks.size();
        if ((rc = grow(&m->hidb, &m->hidb_n, (size_t)chunk * nblk * m->hid, st, &m->gen))) return rc;
        if ((rc = grow(&m->embE, &m->embE_n, (size_t)chunk * m->Ctot, st, &m->gen))) return rc;
    }
    return 0;
}

// eps for R rows given prepared wall tables (rows < n_cond conditional).  x_wrap > 0: x holds x_wrap rows and row r
// of the batch reads row r mod x_wrap (the CFG halves share the state)
static int eps_rows(pmp_model* m, const float* x, const int64_t* t, const float* wall_tab, int n_cond, int R, int H,
                    float* eps, float* energy, float* u, cudaStream_t st, int x_wrap = 0) {
    const int D = m->arch.transition_dim;
    const int chunk = aligned_chunk(m, H);
    int rc;
    if ((rc = eps_rows_reserve(m, R, H, st))) return rc;
    const int nblk = (int)m->blocks.size();
    for (int r0 = 0; r0 < R; r0 += chunk) {
        const int Rc = (R - r0) < chunk ? (R - r0) : chunk;
        int nc = n_cond - r0;
        nc = nc < 0 ? 0 : (nc > Rc ? Rc : nc);
        Ctx c;
        c.m = m; c.st = st; c.dry = false; c.R = Rc; c.H = H; c.n_cond = nc; c.t = t + r0;
        if (m->arch.time_mlp_config == 0) {
            c.embRow = wall_tab ? wall_tab + (size_t)r0 * m->Ctot : nullptr;
        } else {
            // per-step block embeddings: E = W2 * act(Ht[t] + Hw[r]) + b2
            const float* Hw = (wall_tab && nc > 0) ? wall_tab + (size_t)r0 * nblk * m->hid : nullptr;
            if ((rc = launch_cfg3_hidden(m->embT, c.t, m->arch.n_timesteps - 1, Hw, Hw ? nc : 0, Rc, nblk * m->hid,
                                         m->arch.act, m->hidb, st)))
                return rc;
            for (int i = 0; i < nblk; ++i) {
                const ResB& b = m->blocks[i];
                if ((rc = launch_small_linear(m->hidb + (size_t)i * m->hid, nblk * m->hid, m->hid, m->A(b.tmW), m->hid, 0,
                                              m->A(b.tmB), 0, m->arch.act, m->embE + b.eoff, m->Ctot, b.Cout, Rc, 0, st)))
                    return rc;
            }
            c.embRow = m->embE;
        }
        m->dbg.clear();
        const float* xin = x + (size_t)r0 * H * D;
        if (x_wrap > 0) {
            c.x_row0 = r0 % x_wrap;
            c.x_wrap = x_wrap;
            xin = x;
        }
        rc = run_unet(c, xin, eps + (size_t)r0 * H * D, energy ? energy + r0 : nullptr,
                      u ? u + (size_t)r0 * H * D : nullptr);
        if (rc) return rc;
    }
    return 0;
}

#include "train_host.inc"

}  // namespace pmp

// ====================================================================== C ABI
extern "C" {

const char* pmp_last_error(void) { return g_err; }
int pmp_abi_version(void) { return PMP_ABI_VERSION; }
uint64_t pmp_launch_count(void) { return g_launches.load(); }

int pmp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
    }
    return ok;
}

int pmp_model_create(const pmp_arch* arch, int device, pmp_model** out) {
    PMP_REQUIRE(arch && out, "null argument");
    int n = 0;
    PMP_CUDA_OK(cudaGetDeviceCount(&n));
    PMP_REQUIRE(device >= 0 && device < n, "device %d not available (%d visible)", device, n);
    cudaDeviceProp p;
    PMP_CUDA_OK(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, p.major, p.minor);
        return PMP_ERR_CUDA;
    }
    DeviceGuard guard(device);
    pmp_model* m = new pmp_model();
    m->arch = *arch;
    m->device = device;
    if (m->arch.down_times <= 0) m->arch.down_times = 1 << 20;
    *out = m;
    return 0;
}

void pmp_model_destroy(pmp_model* m) {
    if (!m) return;
    DeviceGuard guard(m->device);
    if (m->arena16) cudaFree(m->arena16);
    for (auto& g : m->graphs) cudaGraphExecDestroy(g.exec);
    if (m->cap_stream) cudaStreamDestroy(m->cap_stream);
    float* bufs[] = {m->arena, m->temb_tab, m->embT, m->embR, m->embE, m->hidb, m->ws,
                     m->s_eps, m->s_x2, m->s_xs, m->s_sg, m->s_wtab, m->s_tabs};
    for (float* b : bufs)
        if (b) cudaFree(b);
    if (m->t_iota) cudaFree(m->t_iota);
    if (m->train) {
        for (cudaEvent_t e : m->train->bucket_ev) cudaEventDestroy(e);
        if (m->train->plane_jobs_dev) cudaFree(m->train->plane_jobs_dev);
        if (m->train->ev_main) cudaEventDestroy(m->train->ev_main);
        if (m->train->ev_side) cudaEventDestroy(m->train->ev_side);
        if (m->train->side) cudaStreamDestroy(m->train->side);
        if (m->train->map_dev) cudaFree(m->train->map_dev);
        delete m->train;
    }
    delete m;
}

int pmp_model_set_weight(pmp_model* m, const char* key, const float* data, const int64_t* shape, int ndim,
                         int on_device) {
    PMP_REQUIRE(m && key && data && shape && ndim >= 1 && ndim <= 4, "bad argument");
    HostT t;
    int64_t n = 1;
    for (int i = 0; i < ndim; ++i) {
        PMP_REQUIRE(shape[i] > 0, "weight '%s': non-positive dimension", key);
        t.shape.push_back(shape[i]);
        n *= shape[i];
    }
    t.v.resize((size_t)n);
    if (on_device) {
        DeviceGuard guard(m->device);
        PMP_CUDA_OK(cudaMemcpy(t.v.data(), data, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    } else {
        memcpy(t.v.data(), data, (size_t)n * sizeof(float));
    }
    m->hw[key] = std::move(t);
    m->finalized = false;
    return 0;
}

int pmp_model_finalize(pmp_model* m) {
    PMP_REQUIRE(m, "null model");
    DeviceGuard guard(m->device);
    const int rc = finalize_model(m);
    if (m->train && m->train->plane_jobs_dev) {
        // the packed arenas were re-allocated: the device table of plane jobs points into the old ones
        cudaFree(m->train->plane_jobs_dev);
        m->train->plane_jobs_dev = nullptr;
        m->train->n_plane_jobs = 0;
        m->train->max_plane_elems = 0;
    }
    return rc;
}

int pmp_model_set_row_chunk(pmp_model* m, int rows) {
    PMP_REQUIRE(m && rows >= 2, "row chunk must be >= 2");
    m->row_chunk = rows;
    ++m->gen;
    return 0;
}

int pmp_model_set_gemm_path(pmp_model* m, int path) {
    PMP_REQUIRE(m && path >= 0 && path <= 4,
                "gemm path must be 0 (fp32 SIMT), 1 (tcgen05), 2 (operand reuse), 3 (CTA pairs) or 4 (CTA pairs, 16 epilogue warps)");
    m->gemm_path = path;
    ++m->gen;
    return 0;
}

size_t pmp_model_workspace_bytes(const pmp_model* m) { return m ? m->ws_n * sizeof(float) : 0; }

int pmp_wall_embed(pmp_model* m, const float* walls, int B, int wall_len, float* wemb, void* stream) {
    PMP_REQUIRE(m && m->finalized, "model not finalized");
    PMP_REQUIRE(wall_len > 0 && wall_len % m->arch.vit_token_dim == 0, "wall vector length %d is not a multiple of %d",
                wall_len, m->arch.vit_token_dim);
    DeviceGuard guard(m->device);
    return launch_vit(m->vit, walls, B, wall_len / m->arch.vit_token_dim, wemb, (cudaStream_t)stream);
}

int pmp_unet_eps(pmp_model* m, const float* x, const int64_t* t, const float* wemb, int n_cond, int R, int H,
                 float* eps, float* energy, float* u, void* stream) {
    PMP_REQUIRE(m && m->finalized, "model not finalized");
    PMP_REQUIRE(x && t && eps && R >= 1 && n_cond >= 0 && n_cond <= R, "bad argument");
    PMP_REQUIRE(n_cond == 0 || wemb, "wemb required for conditional rows");
    DeviceGuard guard(m->device);
    int rc;
    if ((rc = grow(&m->embR, &m->embR_n, (size_t)(n_cond > 0 ? n_cond : 1) * wall_table_width(m), (cudaStream_t)stream))) return rc;
    if ((rc = prepare_wall_tables(m, wemb, n_cond, m->embR, (cudaStream_t)stream))) return rc;
    return eps_rows(m, x, t, m->embR, n_cond, R, H, eps, energy, u, (cudaStream_t)stream);
}


// ---------------------------------------------------------------------------------------------- training step
int pmp_train_bind(pmp_model* m, int n, const char* const* keys, const int64_t* offsets, float* params_dev, float* grads_dev,
                   int64_t total) {
    PMP_REQUIRE(m && m->finalized, "model not finalized");
    PMP_REQUIRE(n > 0 && keys && offsets && params_dev && grads_dev && total > 0 && total < (int64_t)0xffffffff, "bad argument");
    PMP_REQUIRE(m->arch.act == 0, "the training step is built for SiLU (energy-mode) networks only");
    DeviceGuard guard(m->device);
    if (!m->train) m->train = new TrainState();
    TrainState* ts = m->train;
    ts->P = params_dev; ts->G = grads_dev; ts->n = (size_t)total;
    ts->off.clear();
    for (int i = 0; i < n; ++i) {
        PMP_REQUIRE(keys[i] && offsets[i] >= 0 && offsets[i] < total, "bad offset for parameter %d", i);
        ts->off[keys[i]] = (size_t)offsets[i];
    }
    for (auto& kv : m->hw) {
        auto it = ts->off.find(kv.first);
        PMP_REQUIRE(it != ts->off.end(), "parameter '%s' has no offset", kv.first.c_str());
        PMP_REQUIRE(it->second + kv.second.v.size() <= ts->n, "parameter '%s' exceeds the flat arena", kv.first.c_str());
    }
    return build_repack_map(m, ts);
}

int pmp_train_sync_weights(pmp_model* m, int rebuild_tables, void* stream) {
    PMP_REQUIRE(m && m->finalized && m->train, "pmp_train_bind has not been called");
    DeviceGuard guard(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if ((rc = launch_repack(m->train->P, m->train->map_dev, m->arena, m->arena_n, st))) return rc;
    if (m->arena16 && !m->plane_recs.empty()) {
        TrainState* ts = m->train;
        if (!ts->plane_jobs_dev) {
            struct PJ { const float* w; __nv_bfloat16* hi; __nv_bfloat16* lo; int K, C, N; };
            std::vector<PJ> jobs;
            for (const auto& r : m->plane_recs) {
                jobs.push_back(PJ{m->arena + r.src, m->arena16 + r.hi, m->arena16 + r.lo, r.K, r.C, r.N});
                ts->max_plane_elems = std::max(ts->max_plane_elems, (size_t)r.K * r.C * r.N);
            }
            PMP_CUDA_OK(cudaMalloc(&ts->plane_jobs_dev, jobs.size() * sizeof(PJ)));
            PMP_CUDA_OK(cudaMemcpy(ts->plane_jobs_dev, jobs.data(), jobs.size() * sizeof(PJ), cudaMemcpyHostToDevice));
            ts->n_plane_jobs = (int)jobs.size();
        }
        if ((rc = launch_planes_all(ts->plane_jobs_dev, ts->n_plane_jobs, ts->max_plane_elems, st))) return rc;
    }
    if (rebuild_tables && (rc = build_time_tables(m, st))) return rc;
    return 0;
}

int pmp_train_loss_grad(pmp_model* m, const pmp_train_args* a, void* stream) {
    PMP_REQUIRE(m && m->finalized && m->train, "pmp_train_bind has not been called");
    PMP_REQUIRE(a && a->x_noisy_dev && a->t_dev && a->walls_dev && a->noise_dev && a->keep_dev && a->loss_weights_dev && a->loss_dev,
                "bad argument");
    PMP_REQUIRE(a->B >= 1 && a->H % 4 == 0 && a->H >= 32 && a->H <= 128, "batch %d / horizon %d unsupported", a->B, a->H);
    PMP_REQUIRE(a->wall_len > 0 && a->wall_len % m->arch.vit_token_dim == 0 && a->wall_len / m->arch.vit_token_dim < 16,
                "wall vector length %d does not fit the wall encoder", a->wall_len);
    DeviceGuard guard(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    TrainIO io{a->x_noisy_dev, a->t_dev, a->walls_dev, a->wall_len, a->noise_dev, a->keep_dev, a->loss_weights_dev, a->loss_dev,
               a->eps_dev, a->energy_dev};
    Ctx dry;
    dry.m = m; dry.st = st; dry.dry = true; dry.R = a->B; dry.H = a->H; dry.n_cond = a->B; dry.t = nullptr; dry.embRow = nullptr;
    int rc = run_train(dry, m->train, io);
    if (rc) return rc;
    if ((rc = grow(&m->ws, &m->ws_n, dry.used + 64, st, &m->gen))) return rc;
    Ctx c;
    c.m = m; c.st = st; c.dry = false; c.R = a->B; c.H = a->H; c.n_cond = a->B; c.t = a->t_dev; c.embRow = nullptr;
    m->dbg.clear();
    return run_train(c, m->train, io);
}

int pmp_train_set_buckets(pmp_model* m, int n, const int64_t* lower_bounds) {
    PMP_REQUIRE(m && m->train, "pmp_train_bind has not been called");
    PMP_REQUIRE(n >= 0 && (n == 0 || lower_bounds), "bad argument");
    DeviceGuard guard(m->device);
    TrainState* ts = m->train;
    for (cudaEvent_t e : ts->bucket_ev) cudaEventDestroy(e);
    ts->bucket_ev.clear();
    ts->bucket_lo.clear();
    for (int i = 0; i < n; ++i) {
        PMP_REQUIRE(lower_bounds[i] >= 0 && (i == 0 || lower_bounds[i] < lower_bounds[i - 1]), "bucket bounds must descend");
        cudaEvent_t e;
        PMP_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ts->bucket_ev.push_back(e);
        ts->bucket_lo.push_back((size_t)lower_bounds[i]);
    }
    return 0;
}

int pmp_train_wait_bucket(pmp_model* m, int k, void* stream) {
    PMP_REQUIRE(m && m->train && k >= 0 && k < (int)m->train->bucket_ev.size(), "no such gradient bucket");
    DeviceGuard guard(m->device);
    PMP_CUDA_OK(cudaStreamWaitEvent((cudaStream_t)stream, m->train->bucket_ev[k], 0));
    return 0;
}

int pmp_debug_wgrad(const float* o1, const float* i1, const float* o2, const float* i2, int H, int Co, int Ci, int KS, int R,
                    float* dw_dev, float* scratch_dev, int use_tc, int flags, void* stream) {
    PMP_REQUIRE(o1 && i1 && dw_dev, "bad argument");
    if (use_tc) {
        PMP_REQUIRE(scratch_dev, "scratch required");
        return launch_wgrad_tc(o1, Co, i1, Ci, o2, Co, i2, Ci, H, Co, Ci, KS, R, dw_dev, scratch_dev, (cudaStream_t)stream, flags);
    }
    return launch_wgrad(o1, Co, i1, Ci, o2, Co, i2, Ci, H, Co, H, Ci, 1, KS / 2, KS, R, dw_dev, (cudaStream_t)stream);
}
size_t pmp_debug_wgrad_scratch_floats(int R, int H, int Co, int Ci, int KS) { return wgrad_tc_scratch_floats(R, H, Co, Ci, KS); }

int pmp_q_sample(const float* x0_dev, float* noise_dev, const int64_t* t_dev, const float* sqrt_acp_dev, const float* sqrt_1m_acp_dev,
                 int B, int H, int D, int zero_cond_noise, float* x_noisy_dev, void* stream) {
    PMP_REQUIRE(x0_dev && noise_dev && t_dev && sqrt_acp_dev && sqrt_1m_acp_dev && x_noisy_dev && B >= 1 && H >= 2 && D >= 1, "bad argument");
    return launch_qsample(x0_dev, noise_dev, t_dev, sqrt_acp_dev, sqrt_1m_acp_dev, B, H, D, zero_cond_noise, x_noisy_dev,
                          (cudaStream_t)stream);
}

int pmp_adam_ema_step(float* p_dev, const float* g_dev, float* m_dev, float* v_dev, float* ema_dev, int64_t n, float lr, float beta1,
                      float beta2, float eps, int step, float grad_scale, int ema_mode, float ema_beta, void* stream) {
    PMP_REQUIRE(p_dev && g_dev && m_dev && v_dev && n >= 0 && step >= 1, "bad argument");
    PMP_REQUIRE(ema_mode == 0 || ema_dev, "ema buffer required");
    return launch_adam_ema(p_dev, g_dev, m_dev, v_dev, ema_dev, (size_t)n, lr, beta1, beta2, eps, step, grad_scale, ema_mode, ema_beta,
                           (cudaStream_t)stream);
}

int pmp_step(int kind, const float* x, const float* const* ec, const float* const* eu, const float* w, int K,
             int base, const float* coef, const float* noise, uint64_t seed, uint64_t offset, const float* start,
             const float* goal, float* xo, int B, int H, int D, void* stream) {
    PMP_REQUIRE(x && ec && eu && w && coef && xo, "null argument");
    PMP_REQUIRE(kind == PMP_STEP_DDPM || kind == PMP_STEP_DDIM, "bad step kind");
    return launch_step(kind, x, ec, eu, w, K, base, coef, noise, seed, offset, 0, start, goal, xo, nullptr, 0, B, H, D,
                       (cudaStream_t)stream);
}

int pmp_init_noise(float* x, const float* noise, uint64_t seed, uint64_t offset, float var_temp, const float* start,
                   const float* goal, int B, int H, int D, void* stream) {
    PMP_REQUIRE(x, "null argument");
    return launch_init_noise(x, noise, seed, offset, 0, var_temp, start, goal, nullptr, 0, B, H, D,
                             (cudaStream_t)stream);
}

int pmp_unnormalize(const float* x, const float* lo, const float* hi, float* out, int64_t rows, int D, void* stream) {
    PMP_REQUIRE(x && lo && hi && out, "null argument");
    return launch_unnormalize(x, lo, hi, out, rows, D, (cudaStream_t)stream);
}

int pmp_pick_first_valid(const uint8_t* valid, const int32_t* nchk, int n_prob, int n_samp, int32_t* chosen,
                         uint8_t* success, int32_t* total, void* stream) {
    PMP_REQUIRE(valid && nchk && chosen && success && total && n_samp >= 1, "bad argument");
    return launch_pick_first_valid(valid, nchk, n_prob, n_samp, chosen, success, total, (cudaStream_t)stream);
}

// One reverse-diffusion step of pmp_sample for the row chunk held in the sampler scratch: timestep rows, the K
// potentials' CFG pairs, the fused update.  `dyn`: every per-step value is read from device memory (SamplerState),
// which makes the launch sequence identical for all steps -> captured once as a CUDA graph and replayed.
struct StepCtx {
    const pmp_sample_args* a;
    pmp_model* m0;
    cudaStream_t st;
    int Bn, Bc;               // plans in this chunk / chunk capacity (buffer strides)
    bool share_x, dedup;
    float* x;                 // state of the chunk [Bn][H][D]
    const float *start, *goal;
    int64_t* t2;
    float* coef_tab;
    int32_t* t_tab;
    SamplerState* state;
    float* wtab[PMP_MAX_COMPOSE];
    // direct mode only
    int s;
    uint64_t elem0;
    const float* noise_s;     // injected noise of this step or null
    float* trace_s;
    long tstride;
};

static int enqueue_step(const StepCtx& c, bool dyn) {
    const pmp_sample_args* a = c.a;
    const int K = a->K, H = a->H, D = a->D, Bn = c.Bn;
    const size_t n = (size_t)Bn * H * D, tile = (size_t)c.Bc * H * D;
    cudaStream_t st = c.st;
    int rc;
    if (dyn) rc = launch_sampler_set_t(c.state, c.t_tab, c.t2, (size_t)2 * Bn, st);
    else rc = launch_fill_i64(c.t2, (int64_t)a->t_host[c.s], (size_t)2 * Bn, st);
    if (rc) return rc;
    const float* xin = c.x;
    if (!c.share_x) {
        // cores that read the input as a GEMM operand need the duplicated CFG batch
        float* x2 = c.m0->s_x2;
        if (cudaMemcpyAsync(x2, c.x, n * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(x2 + n, c.x, n * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
            set_error("cudaMemcpyAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
            return PMP_ERR_CUDA;
        }
        xin = x2;
    }
    const int wrap = c.share_x ? Bn : 0;
    const float* ec[PMP_MAX_COMPOSE];
    const float* eu[PMP_MAX_COMPOSE];
    for (int k = 0; k < K; ++k) {
        float* e2 = c.m0->s_eps + (size_t)k * 2 * tile;
        // the unconditional half does not see the walls: potentials that are the SAME network (the reference's
        // "one model, K obstacle subsets" compositions, config/comp) share it bit for bit, so it is evaluated
        // once and only the conditional rows are run for the repeats
        int dup = -1;
        for (int j = 0; j < k && c.dedup; ++j)
            if (a->models[j] == a->models[k]) { dup = j; break; }
        if (dup >= 0) {
            if ((rc = eps_rows(a->models[k], xin, c.t2, c.wtab[k], Bn, Bn, H, e2, nullptr, nullptr, st, wrap))) return rc;
            ec[k] = e2;
            eu[k] = eu[dup];
            continue;
        }
        if ((rc = eps_rows(a->models[k], xin, c.t2, c.wtab[k], Bn, 2 * Bn, H, e2, nullptr, nullptr, st, wrap))) return rc;
        ec[k] = e2;
        eu[k] = e2 + n;
    }
    if (dyn) {
        if ((rc = launch_step(a->kind, c.x, ec, eu, a->w, K, a->base, nullptr, nullptr, 0, 0, c.elem0, c.start, c.goal, c.x,
                              nullptr, 0, Bn, H, D, st, c.state, c.coef_tab)))
            return rc;
        return launch_sampler_advance(c.state, st);
    }
    return launch_step(a->kind, c.x, ec, eu, a->w, K, a->base, a->coef_host + 8 * c.s, c.noise_s, a->seed, (uint64_t)(c.s + 1),
                       c.elem0, c.start, c.goal, c.x, c.trace_s, c.tstride, Bn, H, D, st);
}

int pmp_sample(const pmp_sample_args* a, void* stream) {
    PMP_REQUIRE(a, "null argument");
    PMP_REQUIRE(a->K >= 1 && a->K <= PMP_MAX_COMPOSE, "K out of range");
    PMP_REQUIRE(a->B >= 1 && a->n_steps >= 1 && a->t_host && a->coef_host && a->x_dev && a->start_dev && a->goal_dev,
                "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int K = a->K, B = a->B, H = a->H, D = a->D, n_steps = a->n_steps;
    for (int k = 0; k < K; ++k) {
        PMP_REQUIRE(a->models[k] && a->models[k]->finalized, "model %d not finalized", k);
        PMP_REQUIRE(a->models[k]->arch.transition_dim == D, "model %d has transition_dim %d, expected %d", k,
                    a->models[k]->arch.transition_dim, D);
        PMP_REQUIRE(a->wemb_dev[k], "wemb_dev[%d] is null", k);
        PMP_REQUIRE(a->models[k]->device == a->models[0]->device, "composed models live on different devices");
    }
    for (int s = 0; s < n_steps; ++s)
        for (int k = 0; k < K; ++k)
            PMP_REQUIRE(a->t_host[s] >= 0 && a->t_host[s] < a->models[k]->arch.n_timesteps,
                        "step %d: timestep %d outside model %d's schedule (n_timesteps %d)", s, a->t_host[s], k,
                        a->models[k]->arch.n_timesteps);
    pmp_model* m0 = a->models[0];
    DeviceGuard guard(m0->device);
    // plans per pass: the CFG pair doubles the rows
    int chunk = aligned_chunk(m0, H);
    for (int k = 1; k < K; ++k)
        if (aligned_chunk(a->models[k], H) < chunk) chunk = aligned_chunk(a->models[k], H);
    int Bc = chunk / 2;
    if (Bc < 1) Bc = 1;
    if (Bc > B) Bc = B;
    bool share_x = true;
    for (int k = 0; k < K; ++k) share_x = share_x && can_share_x(a->models[k], H);
    const size_t tile = (size_t)Bc * H * D;
    // ---- scratch of the sampler (grow-only, stream-ordered; a steady workload allocates nothing here)
    int rc;
    size_t wt_off[PMP_MAX_COMPOSE + 1];
    wt_off[0] = 0;
    for (int k = 0; k < K; ++k) wt_off[k + 1] = wt_off[k] + (((size_t)Bc * wall_table_width(a->models[k]) + 3) & ~(size_t)3);
    const size_t t2_f = (size_t)4 * Bc, coef_f = (size_t)8 * n_steps, tt_f = ((size_t)n_steps + 3) & ~(size_t)3;
    if ((rc = grow(&m0->s_eps, &m0->s_eps_n, (size_t)K * 2 * tile, st, &m0->gen))) return rc;
    if (!share_x && (rc = grow(&m0->s_x2, &m0->s_x2_n, 2 * tile, st, &m0->gen))) return rc;
    if ((rc = grow(&m0->s_xs, &m0->s_xs_n, tile, st, &m0->gen))) return rc;
    if ((rc = grow(&m0->s_sg, &m0->s_sg_n, (size_t)2 * Bc * D + 8, st, &m0->gen))) return rc;
    if ((rc = grow(&m0->s_wtab, &m0->s_wtab_n, wt_off[K], st, &m0->gen))) return rc;
    if ((rc = grow(&m0->s_tabs, &m0->s_tabs_n, t2_f + coef_f + tt_f + 8, st, &m0->gen))) return rc;
    for (int k = 0; k < K; ++k)
        if ((rc = eps_rows_reserve(a->models[k], 2 * Bc, H, st))) return rc;
    StepCtx c;
    memset(&c, 0, sizeof c);
    c.a = a; c.m0 = m0; c.st = st; c.Bc = Bc; c.share_x = share_x; c.dedup = g_uncond_dedup.load() != 0;
    c.t2 = reinterpret_cast<int64_t*>(m0->s_tabs);
    c.coef_tab = m0->s_tabs + t2_f;
    c.t_tab = reinterpret_cast<int32_t*>(m0->s_tabs + t2_f + coef_f);
    c.state = reinterpret_cast<SamplerState*>(m0->s_tabs + t2_f + coef_f + tt_f);
    for (int k = 0; k < K; ++k) c.wtab[k] = m0->s_wtab + wt_off[k];
    // A graph per (potentials, chunk size, horizon, step kind) is replayed for every step.  Not with injected noise or a
    // diffusion trace (their addresses change per step and per call; those are the parity / debugging modes), not while
    // the per-launch profiler is recording, and not for a single step.
    const bool use_graph = g_cuda_graph.load() != 0 && !a->noise_dev && !a->trace_dev && !g_prof_on && n_steps >= 2;
    if (use_graph) {
        // pageable host tables: the copies are staged by the runtime before the call returns
        PMP_CUDA_OK(cudaMemcpyAsync(c.coef_tab, a->coef_host, coef_f * sizeof(float), cudaMemcpyHostToDevice, st));
        PMP_CUDA_OK(cudaMemcpyAsync(c.t_tab, a->t_host, (size_t)n_steps * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    }
    const long tstride = (long)(n_steps + 1) * H * D;
    for (int b0 = 0; b0 < B; b0 += Bc) {
        const int Bn = (B - b0) < Bc ? (B - b0) : Bc;
        const size_t n = (size_t)Bn * H * D;
        const uint64_t elem0 = (a->row_offset + (uint64_t)b0) * (uint64_t)(H * D);
        float* tr = a->trace_dev ? a->trace_dev + (size_t)b0 * tstride : nullptr;
        const size_t nstride = (size_t)B * H * D;
        const float* nz = a->noise_dev ? a->noise_dev + (size_t)b0 * H * D : nullptr;
        c.Bn = Bn;
        c.elem0 = elem0;
        c.tstride = tstride;
        for (int k = 0; k < K; ++k)
            if ((rc = prepare_wall_tables(a->models[k], a->wemb_dev[k] + (size_t)b0 * a->models[k]->arch.wall_embed_dim, Bn,
                                          c.wtab[k], st)))
                return rc;
        if (!use_graph) {
            // direct launches, in place on the caller's buffers
            c.x = a->x_dev + (size_t)b0 * H * D;
            c.start = a->start_dev + (size_t)b0 * D;
            c.goal = a->goal_dev + (size_t)b0 * D;
            if ((rc = launch_init_noise(c.x, nz, a->seed, 0, elem0, a->var_temp, c.start, c.goal, tr, tstride, Bn, H, D, st)))
                return rc;
            for (int s = 0; s < n_steps; ++s) {
                c.s = s;
                c.noise_s = a->noise_dev ? a->noise_dev + (size_t)(s + 1) * nstride + (size_t)b0 * H * D : nullptr;
                c.trace_s = tr ? tr + (size_t)(s + 1) * H * D : nullptr;
                if ((rc = enqueue_step(c, false))) return rc;
            }
            continue;
        }
        // graph replay on the scratch copies of state / start / goal (fixed addresses), result copied out at the end
        float* sstart = m0->s_sg;
        float* sgoal = m0->s_sg + (((size_t)Bc * D + 3) & ~(size_t)3);
        PMP_CUDA_OK(cudaMemcpyAsync(sstart, a->start_dev + (size_t)b0 * D, (size_t)Bn * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
        PMP_CUDA_OK(cudaMemcpyAsync(sgoal, a->goal_dev + (size_t)b0 * D, (size_t)Bn * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
        c.x = m0->s_xs;
        c.start = sstart;
        c.goal = sgoal;
        if ((rc = launch_init_noise(c.x, nullptr, a->seed, 0, elem0, a->var_temp, c.start, c.goal, nullptr, 0, Bn, H, D, st)))
            return rc;
        if ((rc = launch_sampler_begin(c.state, a->seed, elem0, st))) return rc;
        std::vector<uint64_t> key;
        key.push_back((uint64_t)K); key.push_back((uint64_t)Bn); key.push_back((uint64_t)Bc); key.push_back((uint64_t)H);
        key.push_back((uint64_t)a->kind); key.push_back((uint64_t)a->base); key.push_back((uint64_t)c.dedup);
        key.push_back((uint64_t)n_steps);     // the tables' offsets inside the scratch depend on it
        for (int k = 0; k < K; ++k) {
            uint32_t wb;
            memcpy(&wb, &a->w[k], 4);
            key.push_back((uint64_t)(uintptr_t)a->models[k]);
            key.push_back(a->models[k]->gen);
            key.push_back((uint64_t)wb);
        }
        cudaGraphExec_t exec = nullptr;
        uint64_t per_step = 0;
        for (auto& g : m0->graphs)
            if (g.key == key) { exec = g.exec; per_step = g.launches; break; }
        if (!exec) {
            cudaGraph_t graph = nullptr;
            if (!m0->cap_stream) PMP_CUDA_OK(cudaStreamCreateWithFlags(&m0->cap_stream, cudaStreamNonBlocking));
            PMP_CUDA_OK(cudaStreamBeginCapture(m0->cap_stream, cudaStreamCaptureModeRelaxed));
            const uint64_t l0 = g_launches.load();
            StepCtx cc = c;
            cc.st = m0->cap_stream;              // recorded only; the graph is launched on the caller's stream
            rc = enqueue_step(cc, true);
            per_step = g_launches.load() - l0;
            g_launches.fetch_sub(per_step);      // recorded, not executed: replays are counted below
            const cudaError_t ce = cudaStreamEndCapture(m0->cap_stream, &graph);
            if (rc || ce != cudaSuccess || !graph) {
                if (graph) cudaGraphDestroy(graph);
                if (!rc) { set_error("CUDA graph capture of the sampler step failed: %s", cudaGetErrorString(ce)); rc = PMP_ERR_CUDA; }
                return rc;
            }
            const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ie != cudaSuccess) { set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ie)); return PMP_ERR_CUDA; }
            if (m0->graphs.size() >= 16) {     // a handful of (chunk size, horizon, steps) combinations in practice
                cudaGraphExecDestroy(m0->graphs.front().exec);
                m0->graphs.erase(m0->graphs.begin());
            }
            m0->graphs.push_back(pmp_model::GraphEnt{key, exec, per_step});
        }
        for (int s = 0; s < n_steps; ++s) PMP_CUDA_OK(cudaGraphLaunch(exec, st));
        g_launches.fetch_add((uint64_t)n_steps * per_step);
        PMP_CUDA_OK(cudaMemcpyAsync(a->x_dev + (size_t)b0 * H * D, c.x, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

int pmp_set_option(const char* name, int value) {
    PMP_REQUIRE(name, "null option name");
    if (!strcmp(name, "uncond_dedup")) { g_uncond_dedup.store(value ? 1 : 0); return 0; }
    if (!strcmp(name, "cuda_graph")) { g_cuda_graph.store(value ? 1 : 0); return 0; }
    set_error("unknown option '%s'", name);
    return PMP_ERR_ARG;
}

int pmp_profile_enable(int on) {
    g_prof_on = on != 0;
    return 0;
}

int pmp_profile_read(int n_classes, double* ms, double* flops, int64_t* launches) {
    PMP_REQUIRE(ms && flops && launches && n_classes >= 1, "bad argument");
    for (int i = 0; i < n_classes; ++i) { ms[i] = 0; flops[i] = 0; launches[i] = 0; }
    FILE* dump = nullptr;
#ifdef PMP_DIAG
    if (const char* path = getenv("PMP_PROFILE_DUMP")) dump = fopen(path, "w");
#endif
    if (dump) fprintf(dump, "cls,R,Hout,N,C1,C2,taps,flags,ms,algo_tflops\n");
    for (auto& r : g_prof) {
        PMP_CUDA_OK(cudaEventSynchronize(r.b));
        float t = 0;
        PMP_CUDA_OK(cudaEventElapsedTime(&t, r.a, r.b));
        if (dump)
            fprintf(dump, "%d,%d,%d,%d,%d,%d,%d,%d,%.4f,%.1f\n", r.cls, r.R, r.Hout, r.N, r.C1, r.C2, r.taps, r.flags, t,
                    r.flops / (t * 1e9));
        if (r.cls < n_classes) { ms[r.cls] += t; flops[r.cls] += r.flops; launches[r.cls] += 1; }
        g_prof_pool.push_back(std::make_pair(r.a, r.b));
    }
    g_prof.clear();
    if (dump) fclose(dump);
    return 0;
}

int pmp_debug_counters(double* out8, int reset) {
    unsigned long long* buf = pmp::tc_timer_buffer();
    unsigned long long h[8] = {0};
    if (buf) {
        cudaDeviceSynchronize();
        cudaMemcpy(h, buf, sizeof h, cudaMemcpyDeviceToHost);
        if (reset) cudaMemset(buf, 0, sizeof h);
    }
    for (int i = 0; i < 8; ++i) out8[i] = (double)h[i];
    return buf ? 0 : 1;
}

int64_t pmp_debug_fetch(pmp_model* m, const char* name, float* host_out, int64_t capacity) {
    if (!m || !name) return -1;
    auto it = m->dbg.find(name);
    if (it == m->dbg.end() || it->second.second == 0) return -1;
    const int64_t n = it->second.second;
    if (host_out) {
        if (capacity < n) return -2;
        DeviceGuard guard(m->device);
        if (cudaDeviceSynchronize() != cudaSuccess) return -3;
        if (cudaMemcpy(host_out, it->second.first, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess)
            return -3;
    }
    return n;
}

}  // extern "C"
