// Batched trajectory-vs-obstacle collision checkers (bandwidth/latency-bound scans).
//
// Replaces (reference file:line):
//   RM2DCollisionChecker.check_1_traj_eps / check_single_traj   diffuser/guides/rm2d_colchk.py:49-93
//   AbstractRobot._valid_state/_state_fp/_edge_fp/distance       pb_diff_envs/robot/abstract_robot.py:97-163
//   Point2DRobot.no_collision                                    pb_diff_envs/robot/point2d_robot.py:50-54
//   RectangleWallGroup.is_point_inside_wg                        pb_diff_envs/environment/rand_rec_group.py:216-237
//   check_single_traj / check_collision (scorer)                 diffuser/guides/kuka_plan_utils.py:87-154
//   KukaCollisionChecker.check_single_traj                       diffuser/guides/kuka_colchk.py:81-104
//     (pybullet contact query replaced by the analytic sphere-link model of oracle/colchk.c)
// Bit-exact contract: float32 interpolation without FMA contraction, float64 box test with
// strict inequalities, integer check counts with the reference's early-exit order.
// One warp owns one trajectory: lanes read consecutive waypoints (coalesced float2 / row
// loads), evaluate segments in parallel and combine with ballot / prefix sums.
#include "common.cuh"

namespace pmp {
namespace {

constexpr int WARPS = 8;
constexpr int MAXW = 32;  // walls per environment

// The reference tests a float32 point against float64 box bounds with strict inequalities
// (rand_rec_group.py:216-237).  For a float32 x and a float64 bound T:
//   (double)x > T  <=>  x >= up(T),   up(T)   = smallest float32 whose value exceeds T
//   (double)x < T  <=>  x <= down(T), down(T) = largest  float32 whose value is below T
// so the bounds are converted ONCE per wall and the per-point test is four float32 compares (exactly the
// reference's truth value, NaNs included) instead of eight float64 ones.
struct Maze {
    const float* lo_f;  // [nw][2]  up((c-h)-m)
    const float* hi_f;  // [nw][2]  down((c+h)+m)
    int nw;
    float lox, loy, hix, hiy;
};

// up(T) / down(T), branch free: round towards the side, then step one float32 further when the rounded value does not
// lie strictly beyond T.  The step is the integer successor of the bit pattern (+-0 -> the smallest denormal of the right
// sign; -inf steps up to -FLT_MAX, +inf down to FLT_MAX); +inf is never stepped up, -inf never down, NaN never compares.
__device__ __forceinline__ float strictly_above(double t) {
    const float f = __double2float_ru(t);
    const int b = __float_as_int(f);
    const int up = (b & 0x7fffffff) == 0 ? 1 : (b >= 0 ? b + 1 : b - 1);
    return (((double)f <= t) && b != 0x7f800000) ? __int_as_float(up) : f;
}
__device__ __forceinline__ float strictly_below(double t) {
    const float f = __double2float_rd(t);
    const int b = __float_as_int(f);
    const int dn = (b & 0x7fffffff) == 0 ? (int)0x80000001u : (b >= 0 ? b - 1 : b + 1);
    return (((double)f >= t) && b != (int)0xff800000u) ? __int_as_float(dn) : f;
}

__device__ __forceinline__ bool m2d_valid(const Maze& e, float x, float y) {
    return x >= e.lox && y >= e.loy && x <= e.hix && y <= e.hiy;
}
__device__ __forceinline__ bool m2d_collides(const Maze& e, float x, float y) {
    for (int i = 0; i < e.nw; ++i)
        if (x >= e.lo_f[2 * i] && y >= e.lo_f[2 * i + 1] && x <= e.hi_f[2 * i] && y <= e.hi_f[2 * i + 1]) return true;
    return false;
}
// _state_fp: returns ok, bumps cnt by one only when the state is inside the limits
__device__ __forceinline__ bool m2d_state_fp(const Maze& e, float x, float y, int& cnt) {
    if (!m2d_valid(e, x, y)) return false;
    cnt += 1;
    return !m2d_collides(e, x, y);
}

__device__ bool m2d_edge_fp(const Maze& e, float ax, float ay, float bx, float by, double eps, int legacy_div,
                            int& cnt) {
    if (!m2d_valid(e, ax, ay) || !m2d_valid(e, bx, by)) return false;
    if (!m2d_state_fp(e, ax, ay, cnt)) return false;
    if (!m2d_state_fp(e, bx, by, cnt)) return false;
    const float dx = __fsub_rn(bx, ax), dy = __fsub_rn(by, ay);
    const float cx = fminf(fmaxf(bx, e.lox), e.hix), cy = fminf(fmaxf(by, e.loy), e.hiy);
    const float fx = fabsf(__fsub_rn(cx, ax)), fy = fabsf(__fsub_rn(cy, ay));
    const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)));
    int K;
    if (legacy_div) K = (int)__ddiv_rn((double)d, eps);
    else K = (int)__fdiv_rn(d, (float)eps);
    for (int k = 1; k < K; ++k) {
        const float coef = (float)__ddiv_rn((double)k, (double)K);
        const float px = __fadd_rn(ax, __fmul_rn(coef, dx));
        const float py = __fadd_rn(ay, __fmul_rn(coef, dy));
        if (!m2d_state_fp(e, px, py, cnt)) return false;
    }
    return true;
}

__device__ __forceinline__ void load_maze(Maze& e, float* sm, const double* cen, const double* hext, int env,
                                          int nw, double margin, int lane) {
    // each lane prepares one wall: identical double ops to the reference ((c-h)-m, (c+h)+m), then the exact
    // float32 thresholds
    for (int i = lane; i < nw; i += 32) {
        const double* c = cen + ((size_t)env * nw + i) * 2;
        const double* h = hext + ((size_t)env * nw + i) * 2;
        sm[2 * i] = strictly_above(__dsub_rn(__dsub_rn(c[0], h[0]), margin));
        sm[2 * i + 1] = strictly_above(__dsub_rn(__dsub_rn(c[1], h[1]), margin));
        sm[2 * MAXW + 2 * i] = strictly_below(__dadd_rn(__dadd_rn(c[0], h[0]), margin));
        sm[2 * MAXW + 2 * i + 1] = strictly_below(__dadd_rn(__dadd_rn(c[1], h[1]), margin));
    }
    __syncwarp();
    e.lo_f = sm;
    e.hi_f = sm + 2 * MAXW;
    e.nw = nw;
}

__global__ void __launch_bounds__(WARPS * 32) maze2d_swept_kernel(
    const float* __restrict__ traj, const int32_t* __restrict__ env_id, int N, int H, const double* __restrict__ cen,
    const double* __restrict__ hext, int nw, float lox, float loy, float hix, float hiy, double eps, double margin,
    int legacy_div, uint8_t* __restrict__ valid, int32_t* __restrict__ nchk) {
    __shared__ float sm[WARPS][4 * MAXW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * WARPS + warp;
    if (n >= N) return;
    // the trajectory does not depend on the environment: its first 64 segments are requested before the
    // env_id -> walls chain of dependent loads, so the two DRAM round trips overlap
    const float2* t = reinterpret_cast<const float2*>(traj + (size_t)n * H * 2);
    float2 pa[2], pb[2];
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int seg = it * 32 + lane;
        if (seg < H - 1) { pa[it] = __ldg(t + seg); pb[it] = __ldg(t + seg + 1); }
    }
    Maze e;
    e.lox = lox; e.loy = loy; e.hix = hix; e.hiy = hiy;
    load_maze(e, sm[warp], cen, hext, env_id[n], nw, margin, lane);
    int total = 0;
    bool ok_all = true;
    for (int base = 0; base < H - 1 && ok_all; base += 32) {
        const int seg = base + lane;
        bool ok = true;
        int cnt = 0;
        if (seg < H - 1) {
            const float2 a = base < 64 ? pa[base >> 5] : __ldg(t + seg), b = base < 64 ? pb[base >> 5] : __ldg(t + seg + 1);
            ok = m2d_edge_fp(e, a.x, a.y, b.x, b.y, eps, legacy_div, cnt);
        }
        const unsigned fail = __ballot_sync(0xffffffffu, !ok);
        const int first = fail ? (__ffs(fail) - 1) : 32;
        int c = (lane <= first) ? cnt : 0;  // segments after the first failing one are never evaluated
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        total += c;
        if (fail) ok_all = false;
    }
    if (lane == 0) {
        valid[n] = ok_all ? 1 : 0;
        nchk[n] = total;
    }
}

// ------------------------------------------------------------------------------ swept check, horizons <= 64
// Half a warp per trajectory, W consecutive waypoints per lane (H <= 16 W): both trajectories of a warp are evaluated
// in ONE pass with every lane busy, each waypoint is limit- and wall-tested once (the reference tests interior
// waypoints twice, as the end of one segment and the start of the next: same truth value, both counted), and the
// converted walls stay in shared memory for as long as consecutive trajectories name the same environment (the
// `samples_perprob` candidates of a problem are adjacent).  Blocks walk the batch in a grid-stride loop.
// K = int(d / eps) interior points per segment: when |b - a|^2 < (1.8 eps)^2 the quotient is < 2 whatever the
// rounding, the reference's `for k in range(1, K)` is empty, and the square root and the (float64) division are
// skipped; otherwise the reference's exact operations run.  valid / n_checks are bit-identical to the sequential walk.
// a segment whose endpoints are inside the limits and collision free: the reference's interior points (K = int(d / eps))
__device__ __noinline__ bool m2d_interior(const float4* __restrict__ walls, int nw, float lox, float loy, float hix, float hiy,
                                          float2 a, float dx, float dy, float ss, double eps, int legacy_div, int& cnt) {
    const float d = __fsqrt_rn(ss);
    int K;
    if (legacy_div) K = (int)__ddiv_rn((double)d, eps);
    else K = (int)__fdiv_rn(d, (float)eps);
    for (int k = 1; k < K; ++k) {
        const float coef = (float)__ddiv_rn((double)k, (double)K);
        const float px = __fadd_rn(a.x, __fmul_rn(coef, dx));
        const float py = __fadd_rn(a.y, __fmul_rn(coef, dy));
        if (!(px >= lox && py >= loy && px <= hix && py <= hiy)) return false;      // _state_fp: outside the limits, not counted
        cnt += 1;
        for (int i = 0; i < nw; ++i) {
            const float4 q = walls[i];
            if (px >= q.x && py >= q.y && px <= q.z && py <= q.w) return false;
        }
    }
    return true;
}

template <int W>
__global__ void __launch_bounds__(WARPS * 32) maze2d_swept16_kernel(
    const float* __restrict__ traj, const int32_t* __restrict__ env_id, int N, int H, const double* __restrict__ cen,
    const double* __restrict__ hext, int nw, float lox, float loy, float hix, float hiy, double eps, float thr2,
    double margin, int legacy_div, uint8_t* __restrict__ valid, int32_t* __restrict__ nchk) {
    __shared__ float4 sm[2 * WARPS][MAXW];       // per half warp: (lo.x, lo.y, hi.x, hi.y) thresholds of every wall
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = lane >> 4, hl = lane & 15;
    float4* const walls = sm[2 * warp + half];
    int cur_env = -1;
    const int npair = (N + 1) >> 1;
    // the waypoints and the environment id of the NEXT trajectory are requested before the current one is evaluated
    float2 wn[W];
    int envn = -1;
    auto fetch = [&](int pair) {
        const int n = 2 * pair + half;
        const bool act = pair < npair && n < N;
        const float2* t = reinterpret_cast<const float2*>(traj) + (size_t)(act ? n : 0) * H + hl * W;
#pragma unroll
        for (int i = 0; i < W; ++i) wn[i] = (act && hl * W + i < H) ? __ldg(t + i) : make_float2(0.f, 0.f);
        envn = act ? __ldg(env_id + n) : -1;
    };
    fetch(blockIdx.x * WARPS + warp);
    // number of this lane's segments (segment s joins waypoints s, s + 1; s < H - 1)
    const int nseg = min(W, max(0, H - 1 - hl * W));
    for (int pair = blockIdx.x * WARPS + warp; pair < npair; pair += gridDim.x * WARPS) {
        const int n = 2 * pair + half;
        const bool active = n < N;
        float2 w[W + 1];
#pragma unroll
        for (int i = 0; i < W; ++i) w[i] = wn[i];
        const int env = active ? envn : cur_env;
        fetch(pair + gridDim.x * WARPS);
        if (env != cur_env) {
            // this half-warp's walls: identical double ops to the reference ((c-h)-m, (c+h)+m), then exact float32 thresholds
            for (int i = hl; i < nw; i += 16) {
                const double* c = cen + ((size_t)env * nw + i) * 2;
                const double* h = hext + ((size_t)env * nw + i) * 2;
                walls[i] = make_float4(strictly_above(__dsub_rn(__dsub_rn(c[0], h[0]), margin)),
                                       strictly_above(__dsub_rn(__dsub_rn(c[1], h[1]), margin)),
                                       strictly_below(__dadd_rn(__dadd_rn(c[0], h[0]), margin)),
                                       strictly_below(__dadd_rn(__dadd_rn(c[1], h[1]), margin)));
            }
            cur_env = env;
        }
        __syncwarp();
        // every waypoint once: inside the limits (bit 0), inside an inflated wall (bit 1); branch free, one broadcast
        // shared-memory load per wall for the lane's W waypoints
        bool inl[W + 1], col[W + 1];
#pragma unroll
        for (int i = 0; i < W; ++i) {
            inl[i] = w[i].x >= lox && w[i].y >= loy && w[i].x <= hix && w[i].y <= hiy;
            col[i] = false;
        }
        for (int k = 0; k < nw; ++k) {
            const float4 q = walls[k];
#pragma unroll
            for (int i = 0; i < W; ++i) col[i] = col[i] | ((w[i].x >= q.x) & (w[i].y >= q.y) & (w[i].x <= q.z) & (w[i].y <= q.w));
        }
        // the first waypoint of the next lane closes this lane's last segment
        w[W].x = __shfl_down_sync(0xffffffffu, w[0].x, 1, 16);
        w[W].y = __shfl_down_sync(0xffffffffu, w[0].y, 1, 16);
        const unsigned fn = __shfl_down_sync(0xffffffffu, (inl[0] ? 1u : 0u) | (col[0] ? 2u : 0u), 1, 16);
        inl[W] = (fn & 1u) != 0;
        col[W] = (fn & 2u) != 0;
        // _edge_fp per segment, in order, stopping at the lane's first failure: 0 checks when an endpoint is outside the
        // limits, 1 when the first endpoint collides, 2 when the second does, 2 + interior points otherwise
        int part = 0;
        bool failed = !active;
#pragma unroll
        for (int j = 0; j < W; ++j) {
            if (j < nseg && !failed) {
                if (!(inl[j] && inl[j + 1])) failed = true;
                else if (col[j]) { part += 1; failed = true; }
                else if (col[j + 1]) { part += 2; failed = true; }
                else {
                    part += 2;
                    // both endpoints are inside the limits, so the reference's clipped end point is the end point itself
                    const float dx = __fsub_rn(w[j + 1].x, w[j].x), dy = __fsub_rn(w[j + 1].y, w[j].y);
                    const float ss = __fadd_rn(__fmul_rn(fabsf(dx), fabsf(dx)), __fmul_rn(fabsf(dy), fabsf(dy)));
                    if (!(ss < thr2))            // K >= 2 possible: the reference's exact square root, division and points
                        failed = !m2d_interior(walls, nw, lox, loy, hix, hiy, w[j], dx, dy, ss, eps, legacy_div, part);
                }
            }
        }
        // segments after the first failing one are never evaluated by the reference: lanes past the lowest failing lane drop out
        const unsigned fm = (__ballot_sync(0xffffffffu, failed) >> (16 * half)) & 0xffffu;
        const int first = fm ? (__ffs(fm) - 1) : 16;
        int c = (hl <= first) ? part : 0;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o, 16);
        if (active && hl == 0) {
            valid[n] = fm ? 0 : 1;
            nchk[n] = c;
        }
        __syncwarp();          // the walls may be rewritten in the next iteration
    }
}

// ------------------------------------------------------------------------------ swept check, one LANE per trajectory
// The bandwidth-bound form (few walls, any horizon): a lane walks its own trajectory sequentially, exactly like the
// reference's loop (no cross-lane prefix sums, ballots or duplicated per-lane bookkeeping: ~35 instructions per waypoint for
// 32 trajectories at once), with the walls of its environment as float32 thresholds in registers.  The warp stages the 32
// trajectories it owns through shared memory in chunks of 8 waypoints: global loads stay fully coalesced (every row of a
// chunk is 64 contiguous bytes, one 16-byte load per lane and 8 rows), the next chunk is in flight while the current one is
// walked, and the per-lane reads of the staged rows are bank-conflict free (row stride 9 float2).  A warp stops loading once
// every lane has failed.  valid / n_checks are bit-identical to the sequential walk.
template <int NWR>
__device__ __forceinline__ int m2d_interior_reg(const float4 (&wl)[NWR], float lox, float loy, float hix, float hiy, float2 a,
                                                float dx, float dy, float ss, double eps, int legacy_div) {
    // returns (points counted << 1) | ok
    const float d = __fsqrt_rn(ss);
    int K;
    if (legacy_div) K = (int)__ddiv_rn((double)d, eps);
    else K = (int)__fdiv_rn(d, (float)eps);
    int cnt = 0;
    for (int k = 1; k < K; ++k) {
        const float coef = (float)__ddiv_rn((double)k, (double)K);
        const float px = __fadd_rn(a.x, __fmul_rn(coef, dx));
        const float py = __fadd_rn(a.y, __fmul_rn(coef, dy));
        if (!(px >= lox && py >= loy && px <= hix && py <= hiy)) return cnt << 1;      // outside the limits: not counted
        cnt += 1;
        bool c = false;
#pragma unroll
        for (int i = 0; i < NWR; ++i) c = c || (px >= wl[i].x && py >= wl[i].y && px <= wl[i].z && py <= wl[i].w);
        if (c) return cnt << 1;
    }
    return (cnt << 1) | 1;
}

template <int NWR>
__global__ void __launch_bounds__(WARPS * 32) maze2d_swept_lane_kernel(
    const float* __restrict__ traj, const int32_t* __restrict__ env_id, int N, int H, const double* __restrict__ cen,
    const double* __restrict__ hext, int nw, float lox, float loy, float hix, float hiy, double eps, float thr2,
    double margin, int legacy_div, uint8_t* __restrict__ valid, int32_t* __restrict__ nchk) {
    constexpr int CH = 8;
    __shared__ float2 stage[WARPS][32][CH + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = lane & 3, rq = lane >> 2;
    const int ngroups = (N + 31) >> 5, nchunks = (H + CH - 1) / CH;
    for (int grp = blockIdx.x * WARPS + warp; grp < ngroups; grp += gridDim.x * WARPS) {
        const int n = grp * 32 + lane;
        const bool active = n < N;
        // cooperative load of chunk c: lane (rq, q) fetches waypoints 2q, 2q+1 of rows rq, rq+8, rq+16, rq+24
        float4 pre[4];
        auto fetch = [&](int c) {
            const int h0 = c * CH + 2 * q;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int nr = grp * 32 + rq + 8 * it;
                pre[it] = (nr < N && h0 < H) ? __ldg(reinterpret_cast<const float4*>(traj + ((size_t)nr * H + h0) * 2))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        fetch(0);
        // this lane's walls: identical double ops to the reference ((c-h)-m, (c+h)+m), then exact float32 thresholds
        float4 wl[NWR];
        const int env = active ? __ldg(env_id + n) : 0;
#pragma unroll
        for (int i = 0; i < NWR; ++i) {
            if (active && i < nw) {
                const double2 c = *reinterpret_cast<const double2*>(cen + ((size_t)env * nw + i) * 2);
                const double2 h = *reinterpret_cast<const double2*>(hext + ((size_t)env * nw + i) * 2);
                wl[i] = make_float4(strictly_above(__dsub_rn(__dsub_rn(c.x, h.x), margin)),
                                    strictly_above(__dsub_rn(__dsub_rn(c.y, h.y), margin)),
                                    strictly_below(__dadd_rn(__dadd_rn(c.x, h.x), margin)),
                                    strictly_below(__dadd_rn(__dadd_rn(c.y, h.y), margin)));
            } else {
                wl[i] = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);      // never inside
            }
        }
        // Walk state, kept small so that the per-waypoint work is the reference's compares and little else.
        //   * _edge_fp of segment (h - 1, h) makes 0 checks when an endpoint is outside the limits, 1 when the first endpoint
        //     collides, 2 (+ interior points) when both endpoints are tested, and is clear iff both are inside and free.
        //   * While a trajectory is alive every earlier segment was clear, so from the second segment on the previous waypoint
        //     is known to be inside and free: the segment makes 2 checks iff waypoint h is inside the limits and is clear iff
        //     h is also free.  Only the FIRST segment needs the general rule (waypoint 0 may be outside or colliding).
        //   * llox = the lower x limit, replaced by NaN once the trajectory has failed (or for an idle lane): every later
        //     waypoint then tests "outside the limits", which counts nothing and stays failed -- the early exit of the
        //     reference's loop without a `done` flag; valid = llox is not NaN.
        const float qnan = __int_as_float(0x7fc00000);
        float llox = active ? lox : qnan;
        // the other limits and the threshold in ordinary (lane-dependent, so not re-materialised) registers: as kernel
        // parameters they are re-fetched into uniform registers around every rare-path branch, three issue slots per waypoint
        const float lloy = active ? loy : qnan, lhix = active ? hix : qnan, lhiy = active ? hiy : qnan;
        const float lthr2 = active ? thr2 : qnan;
        auto hits = [&](const float2 w) {
            bool col = false;
#pragma unroll
            for (int i = 0; i < NWR; ++i) col = col | ((w.x >= wl[i].x) & (w.y >= wl[i].y) & (w.x <= wl[i].z) & (w.y <= wl[i].w));
            return col;
        };
        int cnt = 0;
        float2 pw = make_float2(0.f, 0.f);
        for (int c = 0; c < nchunks; ++c) {
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                stage[warp][rq + 8 * it][2 * q] = make_float2(pre[it].x, pre[it].y);
                stage[warp][rq + 8 * it][2 * q + 1] = make_float2(pre[it].z, pre[it].w);
            }
            __syncwarp();
            if (c + 1 < nchunks) fetch(c + 1);
            const int jn = min(CH, H - c * CH);      // even: the dispatch requires an even H
            int j = 0;
            if (c == 0) {      // waypoints 0 and 1: the first segment, general rule
                const float2 w0 = stage[warp][lane][0];
                const bool in0 = (w0.x >= llox) & (w0.y >= lloy) & (w0.x <= lhix) & (w0.y <= lhiy);
                const int ps = in0 ? (hits(w0) ? 1 : 2) : 0;
                const float2 w = stage[warp][lane][1];
                const bool in = (w.x >= llox) & (w.y >= lloy) & (w.x <= lhix) & (w.y <= lhiy);
                cnt = in ? ps : 0;
                const bool clear = in & (ps == 2) & !hits(w);
                if (!clear) llox = qnan;
                const float dx = __fsub_rn(w.x, w0.x), dy = __fsub_rn(w.y, w0.y);
                const float ss = __fadd_rn(__fmul_rn(fabsf(dx), fabsf(dx)), __fmul_rn(fabsf(dy), fabsf(dy)));
                if (clear && !(ss < lthr2)) {
                    const int r = m2d_interior_reg<NWR>(wl, lox, loy, hix, hiy, w0, dx, dy, ss, eps, legacy_div);
                    cnt += r >> 1;
                    if (!(r & 1)) llox = qnan;
                }
                pw = w;
                j = 2;
            }
#pragma unroll 2
            for (; j < jn; ++j) {
                const float2 w = stage[warp][lane][j];
                const bool in = (w.x >= llox) & (w.y >= lloy) & (w.x <= lhix) & (w.y <= lhiy);
                const bool clear = in & !hits(w);
                cnt += in ? 2 : 0;
                if (!clear) llox = qnan;
                const float dx = __fsub_rn(w.x, pw.x), dy = __fsub_rn(w.y, pw.y);
                const float ss = __fadd_rn(__fmul_rn(fabsf(dx), fabsf(dx)), __fmul_rn(fabsf(dy), fabsf(dy)));
                if (clear && !(ss < lthr2)) {      // K >= 2 possible: the reference's exact square root, division and points
                    const int r = m2d_interior_reg<NWR>(wl, lox, loy, hix, hiy, pw, dx, dy, ss, eps, legacy_div);
                    cnt += r >> 1;
                    if (!(r & 1)) llox = qnan;
                }
                pw = w;
            }
            if (__all_sync(0xffffffffu, llox != llox)) break;
        }
        const bool ok = !(llox != llox);
        if (active) {
            valid[n] = ok ? 1 : 0;
            nchk[n] = cnt;
        }
    }
}

__global__ void __launch_bounds__(WARPS * 32) maze2d_points_kernel(
    const float* __restrict__ traj, const int32_t* __restrict__ env_id, int N, int H, const double* __restrict__ cen,
    const double* __restrict__ hext, int nw, float lox, float loy, float hix, float hiy, double margin,
    int32_t* __restrict__ ncol, uint8_t* __restrict__ ptcol) {
    __shared__ float sm[WARPS][4 * MAXW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * WARPS + warp;
    if (n >= N) return;
    Maze e;
    e.lox = lox; e.loy = loy; e.hix = hix; e.hiy = hiy;
    load_maze(e, sm[warp], cen, hext, env_id[n], nw, margin, lane);
    const float2* t = reinterpret_cast<const float2*>(traj + (size_t)n * H * 2);
    int cnt = 0;
    bool bad = false;
    for (int base = 0; base < H; base += 32) {
        const int i = base + lane;
        bool col = false, oob = false;
        if (i < H) {
            const float2 p = t[i];
            oob = !m2d_valid(e, p.x, p.y);
            col = m2d_collides(e, p.x, p.y);
            if (ptcol) ptcol[(size_t)n * H + i] = col ? 1 : 0;
        }
        cnt += __popc(__ballot_sync(0xffffffffu, col));
        bad = bad || (__ballot_sync(0xffffffffu, oob) != 0u);
    }
    if (lane == 0) ncol[n] = bad ? -1 : cnt;
}

// ------------------------------------------------------------------------------ Kuka spheres
constexpr int MAXSPH = 32;
struct KukaP {
    float fixed[7][12];
    float sph[MAXSPH][5];
    float base[2][3];
    int n_sph, n_arms, nv;
    float table_z;
};

__device__ __forceinline__ void fk_step(float (&T)[12], const float (&F)[12], float s, float c) {
    float M[12];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            float v = __fmul_rn(T[4 * a + 0], F[0 + b]);
            v = __fadd_rn(v, __fmul_rn(T[4 * a + 1], F[4 + b]));
            v = __fadd_rn(v, __fmul_rn(T[4 * a + 2], F[8 + b]));
            M[4 * a + b] = v;
        }
        M[4 * a + 3] = __fadd_rn(M[4 * a + 3], T[4 * a + 3]);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float m0 = M[4 * a + 0], m1 = M[4 * a + 1];
        T[4 * a + 0] = __fadd_rn(__fmul_rn(m0, c), __fmul_rn(m1, s));
        T[4 * a + 1] = __fsub_rn(__fmul_rn(m1, c), __fmul_rn(m0, s));
        T[4 * a + 2] = M[4 * a + 2];
        T[4 * a + 3] = M[4 * a + 3];
    }
}

__device__ bool kuka_waypoint(const KukaP& p, const float* __restrict__ q, const float* __restrict__ bc,
                              const float* __restrict__ bh) {
    float wc[2][MAXSPH][4];
    for (int a = 0; a < p.n_arms; ++a) {
        float T[12] = {1.f, 0.f, 0.f, p.base[a][0], 0.f, 1.f, 0.f, p.base[a][1], 0.f, 0.f, 1.f, p.base[a][2]};
        int k = 0;
        for (int l = 0; l <= 7; ++l) {
            if (l > 0) {
                const double qa = (double)q[7 * a + l - 1];
                fk_step(T, p.fixed[l - 1], (float)sin(qa), (float)cos(qa));
            }
            // sphere table rows are sorted by link
            while (k < p.n_sph && (int)p.sph[k][0] == l) {
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    float v = __fmul_rn(T[4 * ax + 0], p.sph[k][1]);
                    v = __fadd_rn(v, __fmul_rn(T[4 * ax + 1], p.sph[k][2]));
                    v = __fadd_rn(v, __fmul_rn(T[4 * ax + 2], p.sph[k][3]));
                    wc[a][k][ax] = __fadd_rn(v, T[4 * ax + 3]);
                }
                wc[a][k][3] = p.sph[k][4];
                ++k;
            }
        }
    }
    for (int a = 0; a < p.n_arms; ++a)
        for (int k = 0; k < p.n_sph; ++k) {
            const float r = wc[a][k][3], r2 = __fmul_rn(r, r);
            if ((int)p.sph[k][0] != 0 && __fsub_rn(wc[a][k][2], r) < p.table_z) return true;
            for (int v = 0; v < p.nv; ++v) {
                float d2 = 0.f;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    float dj = __fsub_rn(fabsf(__fsub_rn(wc[a][k][j], bc[3 * v + j])), bh[3 * v + j]);
                    dj = fmaxf(dj, 0.f);
                    d2 = __fadd_rn(d2, __fmul_rn(dj, dj));
                }
                if (d2 < r2) return true;
            }
        }
    if (p.n_arms == 2)
        for (int k = 0; k < p.n_sph; ++k)
            for (int m = 0; m < p.n_sph; ++m) {
                float d2 = 0.f;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float dj = __fsub_rn(wc[0][k][j], wc[1][m][j]);
                    d2 = __fadd_rn(d2, __fmul_rn(dj, dj));
                }
                const float rr = __fadd_rn(wc[0][k][3], wc[1][m][3]);
                if (d2 < __fmul_rn(rr, rr)) return true;
            }
    return false;
}

__global__ void __launch_bounds__(WARPS * 32) kuka_kernel(const __grid_constant__ KukaP p,
                                                          const float* __restrict__ traj,
                                                          const int32_t* __restrict__ env_id, int N, int H,
                                                          const float* __restrict__ boxc, const float* __restrict__ boxh,
                                                          uint8_t* __restrict__ valid, int32_t* __restrict__ nchk,
                                                          int32_t* __restrict__ ncol, uint8_t* __restrict__ ptcol) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * WARPS + warp;
    if (n >= N) return;
    const int D = 7 * p.n_arms;
    const float* bc = boxc + (size_t)env_id[n] * p.nv * 3;
    const float* bh = boxh + (size_t)env_id[n] * p.nv * 3;
    int cnt = 0, first = -1;
    for (int base = 0; base < H; base += 32) {
        const int i = base + lane;
        bool col = false;
        if (i < H) {
            col = kuka_waypoint(p, traj + ((size_t)n * H + i) * D, bc, bh);
            if (ptcol) ptcol[(size_t)n * H + i] = col ? 1 : 0;
        }
        const unsigned m = __ballot_sync(0xffffffffu, col);
        cnt += __popc(m);
        if (first < 0 && m) first = base + __ffs(m) - 1;
    }
    if (lane == 0) {
        valid[n] = cnt == 0 ? 1 : 0;
        nchk[n] = first < 0 ? H : first + 1;
        ncol[n] = cnt;
    }
}

void kuka_fixed_host(float out[7][12]) {
    static const double XYZ[7][3] = {{0, 0, 0.1575}, {0, 0, 0.2025}, {0, 0.2045, 0}, {0, 0, 0.2155},
                                     {0, 0.1845, 0}, {0, 0, 0.2155}, {0, 0.081, 0}};
    static const double RPY[7][3] = {{0, 0, 0},
                                     {1.57079632679, 0, 3.14159265359},
                                     {1.57079632679, 0, 3.14159265359},
                                     {1.57079632679, 0, 0},
                                     {-1.57079632679, 3.14159265359, 0},
                                     {1.57079632679, 0, 0},
                                     {-1.57079632679, 3.14159265359, 0}};
    for (int i = 0; i < 7; ++i) {
        const double r = RPY[i][0], pt = RPY[i][1], y = RPY[i][2];
        const double cr = cos(r), sr = sin(r), cp = cos(pt), sp = sin(pt), cy = cos(y), sy = sin(y);
        const double R[9] = {cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
                             sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
                             -sp,     cp * sr,                cp * cr};
        for (int a = 0; a < 3; ++a) {
            for (int b = 0; b < 3; ++b) out[i][4 * a + b] = (float)R[3 * a + b];
            out[i][4 * a + 3] = (float)XYZ[i][a];
        }
    }
}

}  // namespace
}  // namespace pmp

using namespace pmp;

extern "C" int pmp_colchk_maze2d_swept(const float* traj, const int32_t* env_id, int N, int H, const double* cen,
                                       const double* hext, int nw, float lox, float loy, float hix, float hiy,
                                       double eps, double margin, int legacy_div, uint8_t* valid, int32_t* nchk,
                                       void* stream) {
    PMP_REQUIRE(nw >= 0 && nw <= MAXW, "at most %d walls per environment (got %d)", MAXW, nw);
    PMP_REQUIRE(H >= 2, "trajectory needs >= 2 waypoints");
    PMP_REQUIRE(eps > 0, "collision eps must be > 0");
    if (N == 0) return 0;
    if (nw <= 6 && H % 2 == 0 && (reinterpret_cast<uintptr_t>(traj) & 15) == 0 && (reinterpret_cast<uintptr_t>(cen) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(hext) & 15) == 0) {
        // few walls: one lane per trajectory, walls in registers, trajectories staged through shared memory
        static int sms = 0;
        if (!sms) {
            int dev = 0;
            PMP_CUDA_OK(cudaGetDevice(&dev));
            PMP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        }
        const int ngroups = (N + 31) / 32;
        // one group of 32 trajectories per warp: early exits make the groups uneven, so the hardware block scheduler
        // balances better than a grid-stride walk of few persistent blocks (measured)
        const int blocks = (ngroups + WARPS - 1) / WARPS;
        const double t = 1.8 * eps;
        const float thr2 = (float)(t * t);
        cudaStream_t st = (cudaStream_t)stream;
        if (nw <= 3)
            maze2d_swept_lane_kernel<3><<<blocks, WARPS * 32, 0, st>>>(traj, env_id, N, H, cen, hext, nw, lox, loy, hix, hiy, eps, thr2,
                                                                      margin, legacy_div, valid, nchk);
        else
            maze2d_swept_lane_kernel<6><<<blocks, WARPS * 32, 0, st>>>(traj, env_id, N, H, cen, hext, nw, lox, loy, hix, hiy, eps, thr2,
                                                                      margin, legacy_div, valid, nchk);
        PMP_LAUNCH_OK();
        return 0;
    }
    if (H <= 64) {
        // half a warp per trajectory, ceil(H / 16) waypoints per lane, persistent blocks
        static int sms = 0;
        if (!sms) {
            int dev = 0;
            PMP_CUDA_OK(cudaGetDevice(&dev));
            PMP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        }
        const int npair = (N + 1) / 2;
        int blocks = (npair + WARPS - 1) / WARPS;
        if (blocks > sms * 8) blocks = sms * 8;
        const double t = 1.8 * eps;
        const float thr2 = (float)(t * t);
        const int W = (H + 15) / 16;
        cudaStream_t st = (cudaStream_t)stream;
        if (W <= 2)
            maze2d_swept16_kernel<2><<<blocks, WARPS * 32, 0, st>>>(traj, env_id, N, H, cen, hext, nw, lox, loy, hix, hiy, eps,
                                                                   thr2, margin, legacy_div, valid, nchk);
        else if (W == 3)
            maze2d_swept16_kernel<3><<<blocks, WARPS * 32, 0, st>>>(traj, env_id, N, H, cen, hext, nw, lox, loy, hix, hiy, eps,
                                                                   thr2, margin, legacy_div, valid, nchk);
        else
            maze2d_swept16_kernel<4><<<blocks, WARPS * 32, 0, st>>>(traj, env_id, N, H, cen, hext, nw, lox, loy, hix, hiy, eps,
                                                                   thr2, margin, legacy_div, valid, nchk);
        PMP_LAUNCH_OK();
        return 0;
    }
    maze2d_swept_kernel<<<(N + WARPS - 1) / WARPS, WARPS * 32, 0, (cudaStream_t)stream>>>(
        traj, env_id, N, H, cen, hext, nw, lox, loy, hix, hiy, eps, margin, legacy_div, valid, nchk);
    PMP_LAUNCH_OK();
    return 0;
}

extern "C" int pmp_colchk_maze2d_points(const float* traj, const int32_t* env_id, int N, int H, const double* cen,
                                        const double* hext, int nw, float lox, float loy, float hix, float hiy,
                                        double margin, int32_t* ncol, uint8_t* ptcol, void* stream) {
    PMP_REQUIRE(nw >= 0 && nw <= MAXW, "at most %d walls per environment (got %d)", MAXW, nw);
    if (N == 0) return 0;
    maze2d_points_kernel<<<(N + WARPS - 1) / WARPS, WARPS * 32, 0, (cudaStream_t)stream>>>(
        traj, env_id, N, H, cen, hext, nw, lox, loy, hix, hiy, margin, ncol, ptcol);
    PMP_LAUNCH_OK();
    return 0;
}

extern "C" int pmp_colchk_kuka(const float* traj, const int32_t* env_id, int N, int H, int n_arms,
                               const float* base, const float* sph, int n_sph, const float* boxc, const float* boxh,
                               int nv, float table_z, uint8_t* valid, int32_t* nchk, int32_t* ncol, uint8_t* ptcol,
                               void* stream) {
    PMP_REQUIRE(n_arms == 1 || n_arms == 2, "n_arms must be 1 or 2");
    PMP_REQUIRE(n_sph >= 1 && n_sph <= MAXSPH, "1..%d link spheres supported (got %d)", MAXSPH, n_sph);
    KukaP p;
    kuka_fixed_host(p.fixed);
    int prev = 0;
    for (int k = 0; k < n_sph; ++k) {
        for (int j = 0; j < 5; ++j) p.sph[k][j] = sph[5 * k + j];
        const int l = (int)sph[5 * k];
        PMP_REQUIRE(l >= prev && l <= 7, "sphere table must be sorted by link index 0..7");
        prev = l;
    }
    for (int a = 0; a < n_arms; ++a)
        for (int j = 0; j < 3; ++j) p.base[a][j] = base[3 * a + j];
    p.n_sph = n_sph;
    p.n_arms = n_arms;
    p.nv = nv;
    p.table_z = table_z;
    if (N == 0) return 0;
    kuka_kernel<<<(N + WARPS - 1) / WARPS, WARPS * 32, 0, (cudaStream_t)stream>>>(p, traj, env_id, N, H, boxc, boxh,
                                                                                valid, nchk, ncol, ptcol);
    PMP_LAUNCH_OK();
    return 0;
}
