/* CPU restatement of the trajectory-vs-obstacle checkers.  TEST INFRASTRUCTURE ONLY
 * (see oracle/__init__.py).  Build: `make -C oracle` (gcc -O2 -ffp-contract=off).
 *
 * Maze2D follows, op by op and dtype by dtype:
 *   RM2DCollisionChecker.check_1_traj_eps     diffuser/guides/rm2d_colchk.py:76-93
 *   AbsStaticMaze2DEnv.edge_fp                pb_diff_envs/environment/abs_maze2d_env.py:36-37
 *   AbstractRobot._valid_state/_state_fp/_edge_fp/distance
 *                                             pb_diff_envs/robot/abstract_robot.py:97-163
 *   Point2DRobot.set_config/no_collision      pb_diff_envs/robot/point2d_robot.py:42-54
 *   RectangleWall(Group).is_point_inside(_wg) pb_diff_envs/environment/rand_rec_group.py:197-237
 *   check_single_traj / check_collision       diffuser/guides/kuka_plan_utils.py:87-154
 *
 * Kuka / Dual-Kuka: the reference calls pybullet 3.2.5 (individual_robot.py:48-95,
 * grouping.py:49-83), which is not vendored -> this file DEFINES the analytic
 * sphere-link model (parity unpinned): forward kinematics of the iiwa chain
 * (pb_diff_envs/data/robot/kuka_iiwa/model_0.urdf:85-264) in float32 (sin/cos taken in
 * double and rounded to float), link spheres from a table, sphere-vs-box,
 * sphere-vs-table-plane and (dual arm) sphere-vs-sphere tests, strict inequalities.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* ------------------------------------------------------------------ Maze2D */
typedef struct {
    const double *cen;   /* [nw][2] */
    const double *hext;  /* [nw][2] */
    int nw;
    float lo[2], hi[2];
    double margin;
} maze_env;

static int m2d_valid(const maze_env *e, const float *s)
{
    return s[0] >= e->lo[0] && s[1] >= e->lo[1] && s[0] <= e->hi[0] && s[1] <= e->hi[1];
}

static int m2d_collides(const maze_env *e, const float *p)
{
    double px = (double)p[0], py = (double)p[1];
    for (int i = 0; i < e->nw; ++i) {
        double blx = e->cen[2 * i] - e->hext[2 * i], bly = e->cen[2 * i + 1] - e->hext[2 * i + 1];
        double trx = e->cen[2 * i] + e->hext[2 * i], try_ = e->cen[2 * i + 1] + e->hext[2 * i + 1];
        if (px > blx - e->margin && py > bly - e->margin && px < trx + e->margin && py < try_ + e->margin)
            return 1;
    }
    return 0;
}

/* _state_fp: 0 checks if out of limits, else 1 check */
static int m2d_state_fp(const maze_env *e, const float *s, int *nchk)
{
    if (!m2d_valid(e, s)) return 0;
    *nchk += 1;
    return !m2d_collides(e, s);
}

static int m2d_edge_fp(const maze_env *e, const float *a, const float *b, double eps, int legacy_div,
                       int *nchk)
{
    if (!m2d_valid(e, a) || !m2d_valid(e, b)) return 0;
    if (!m2d_state_fp(e, a, nchk)) return 0;
    if (!m2d_state_fp(e, b, nchk)) return 0;
    float disp[2] = { b[0] - a[0], b[1] - a[1] };
    float sq[2];
    for (int j = 0; j < 2; ++j) {
        float t = fmaxf(b[j], e->lo[j]);
        t = fminf(t, e->hi[j]);
        float df = fabsf(t - a[j]);
        sq[j] = df * df;
    }
    float d = sqrtf(sq[0] + sq[1]);
    int K;
    if (legacy_div) K = (int)((double)d / eps);           /* numpy 1.23.4: float32 scalar / python float -> float64 */
    else            K = (int)(d / (float)eps);            /* numpy >= 2 (NEP 50): stays float32 */
    for (int k = 1; k < K; ++k) {
        float coef = (float)(((double)k * 1.) / (double)K);
        float c[2];
        c[0] = a[0] + coef * disp[0];
        c[1] = a[1] + coef * disp[1];
        if (!m2d_state_fp(e, c, nchk)) return 0;
    }
    return 1;
}

/* swept-segment candidate check.  traj [N][H][2] f32 (unnormalised); env_id [N];
 * cen/hext [E][nw][2] f64.  valid [N] u8, nchk [N] i32 */
void pmp_oracle_maze2d_swept(const float *traj, const int32_t *env_id, int N, int H,
                             const double *cen, const double *hext, int nw,
                             const float *lo, const float *hi, double eps, double margin,
                             int legacy_div, uint8_t *valid, int32_t *nchk)
{
    for (int n = 0; n < N; ++n) {
        maze_env e;
        e.cen = cen + (size_t)env_id[n] * nw * 2;
        e.hext = hext + (size_t)env_id[n] * nw * 2;
        e.nw = nw; e.margin = margin;
        e.lo[0] = lo[0]; e.lo[1] = lo[1]; e.hi[0] = hi[0]; e.hi[1] = hi[1];
        const float *t = traj + (size_t)n * H * 2;
        int cnt = 0, ok = 1;
        for (int i = 0; i + 1 < H; ++i) {
            if (!m2d_edge_fp(&e, t + 2 * i, t + 2 * (i + 1), eps, legacy_div, &cnt)) { ok = 0; break; }
        }
        valid[n] = (uint8_t)ok;
        nchk[n] = cnt;
    }
}

/* per-waypoint scorer (kuka_plan_utils.py:87-103): ncol [N] i32, ptcol [N][H] u8 (may be NULL).
 * Waypoints outside the limits would raise in the reference (point2d_robot.py:43-44);
 * they are reported as ncol = -1. */
void pmp_oracle_maze2d_points(const float *traj, const int32_t *env_id, int N, int H,
                              const double *cen, const double *hext, int nw,
                              const float *lo, const float *hi, double margin,
                              int32_t *ncol, uint8_t *ptcol)
{
    for (int n = 0; n < N; ++n) {
        maze_env e;
        e.cen = cen + (size_t)env_id[n] * nw * 2;
        e.hext = hext + (size_t)env_id[n] * nw * 2;
        e.nw = nw; e.margin = margin;
        e.lo[0] = lo[0]; e.lo[1] = lo[1]; e.hi[0] = hi[0]; e.hi[1] = hi[1];
        const float *t = traj + (size_t)n * H * 2;
        int cnt = 0, bad = 0;
        for (int i = 0; i < H; ++i) {
            if (!m2d_valid(&e, t + 2 * i)) bad = 1;
            int c = m2d_collides(&e, t + 2 * i);
            cnt += c;
            if (ptcol) ptcol[(size_t)n * H + i] = (uint8_t)c;
        }
        ncol[n] = bad ? -1 : cnt;
    }
}

/* ------------------------------------------------------------------ Kuka spheres */
/* joint origins of model_0.urdf: xyz | rpy, all joints revolute about local z */
static const double KUKA_XYZ[7][3] = {
    {0, 0, 0.1575}, {0, 0, 0.2025}, {0, 0.2045, 0}, {0, 0, 0.2155},
    {0, 0.1845, 0}, {0, 0, 0.2155}, {0, 0.081, 0}};
static const double KUKA_RPY[7][3] = {
    {0, 0, 0},
    {1.57079632679, 0, 3.14159265359},
    {1.57079632679, 0, 3.14159265359},
    {1.57079632679, 0, 0},
    {-1.57079632679, 3.14159265359, 0},
    {1.57079632679, 0, 0},
    {-1.57079632679, 3.14159265359, 0}};

/* fixed part of joint i as float32 3x4 [R|p]:  R = Rz(yaw) Ry(pitch) Rx(roll)  (URDF convention) */
void pmp_oracle_kuka_fixed(float out[7][12])
{
    for (int i = 0; i < 7; ++i) {
        double r = KUKA_RPY[i][0], p = KUKA_RPY[i][1], y = KUKA_RPY[i][2];
        double cr = cos(r), sr = sin(r), cp = cos(p), sp = sin(p), cy = cos(y), sy = sin(y);
        double R[9] = { cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
                        sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
                        -sp,     cp * sr,                cp * cr };
        for (int a = 0; a < 3; ++a) {
            for (int b = 0; b < 3; ++b) out[i][4 * a + b] = (float)R[3 * a + b];
            out[i][4 * a + 3] = (float)KUKA_XYZ[i][a];
        }
    }
}

/* T <- T * [F | f] * Rz(q)  in float32, fixed evaluation order, no FMA */
static void fk_step(float T[12], const float F[12], float s, float c)
{
    float M[12];
    for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 4; ++b) {
            float v = T[4 * a + 0] * F[0 + b];
            v = v + T[4 * a + 1] * F[4 + b];
            v = v + T[4 * a + 2] * F[8 + b];
            M[4 * a + b] = v;
        }
        M[4 * a + 3] = M[4 * a + 3] + T[4 * a + 3];
    }
    for (int a = 0; a < 3; ++a) {
        float m0 = M[4 * a + 0], m1 = M[4 * a + 1];
        T[4 * a + 0] = m0 * c + m1 * s;
        T[4 * a + 1] = m1 * c - m0 * s;
        T[4 * a + 2] = M[4 * a + 2];
        T[4 * a + 3] = M[4 * a + 3];
    }
}

#define PMP_MAX_SPH 64

/* sphere table rows: (link 0..7, cx, cy, cz, r) float32.  world centres for one arm */
static int arm_spheres(const float *q, const float base[3], const float fixed[7][12],
                       const float *sph, int n_sph, float *wc /*[n_sph][4]*/)
{
    float T[8][12];
    float I[12] = {1, 0, 0, base[0], 0, 1, 0, base[1], 0, 0, 1, base[2]};
    memcpy(T[0], I, sizeof I);
    for (int i = 0; i < 7; ++i) {
        memcpy(T[i + 1], T[i], sizeof I);
        float s = (float)sin((double)q[i]), c = (float)cos((double)q[i]);
        fk_step(T[i + 1], fixed[i], s, c);
    }
    for (int k = 0; k < n_sph; ++k) {
        int l = (int)sph[5 * k];
        const float *M = T[l];
        for (int a = 0; a < 3; ++a) {
            float v = M[4 * a + 0] * sph[5 * k + 1];
            v = v + M[4 * a + 1] * sph[5 * k + 2];
            v = v + M[4 * a + 2] * sph[5 * k + 3];
            wc[4 * k + a] = v + M[4 * a + 3];
        }
        wc[4 * k + 3] = sph[5 * k + 4];
    }
    return n_sph;
}

static int kuka_waypoint_collides(const float *q, int n_arms, const float *base /*[n_arms][3]*/,
                                  const float fixed[7][12], const float *sph, int n_sph,
                                  const float *boxc /*[nv][3]*/, const float *boxh /*[nv][3]*/, int nv,
                                  float table_z)
{
    float wc[2][PMP_MAX_SPH * 4];
    for (int a = 0; a < n_arms; ++a) arm_spheres(q + 7 * a, base + 3 * a, fixed, sph, n_sph, wc[a]);
    for (int a = 0; a < n_arms; ++a) {
        for (int k = 0; k < n_sph; ++k) {
            const float *c = wc[a] + 4 * k;
            float r = c[3], r2 = r * r;
            if ((int)sph[5 * k] != 0 && c[2] - r < table_z) return 1;
            for (int v = 0; v < nv; ++v) {
                float d2 = 0.f;
                for (int j = 0; j < 3; ++j) {
                    float dj = fabsf(c[j] - boxc[3 * v + j]) - boxh[3 * v + j];
                    dj = fmaxf(dj, 0.f);
                    d2 = d2 + dj * dj;
                }
                if (d2 < r2) return 1;
            }
        }
    }
    if (n_arms == 2) {
        for (int k = 0; k < n_sph; ++k)
            for (int m = 0; m < n_sph; ++m) {
                const float *c = wc[0] + 4 * k, *e = wc[1] + 4 * m;
                float d2 = 0.f;
                for (int j = 0; j < 3; ++j) { float dj = c[j] - e[j]; d2 = d2 + dj * dj; }
                float rr = c[3] + e[3];
                if (d2 < rr * rr) return 1;
            }
    }
    return 0;
}

/* per-waypoint Kuka check (kuka_colchk.py:95-104 semantics with early exit + check count,
 * and the kuka_plan_utils.py:87-103 scorer: ncol counts every colliding waypoint).
 * traj [N][H][7*n_arms] f32; boxc/boxh [E][nv][3] f32 */
void pmp_oracle_kuka_points(const float *traj, const int32_t *env_id, int N, int H, int n_arms,
                            const float *base, const float *sph, int n_sph,
                            const float *boxc, const float *boxh, int nv, float table_z,
                            uint8_t *valid, int32_t *nchk, int32_t *ncol, uint8_t *ptcol)
{
    float fixed[7][12];
    pmp_oracle_kuka_fixed(fixed);
    int D = 7 * n_arms;
    for (int n = 0; n < N; ++n) {
        const float *bc = boxc + (size_t)env_id[n] * nv * 3, *bh = boxh + (size_t)env_id[n] * nv * 3;
        int cnt = 0, first = -1;
        for (int i = 0; i < H; ++i) {
            int c = kuka_waypoint_collides(traj + ((size_t)n * H + i) * D, n_arms, base, fixed, sph, n_sph,
                                           bc, bh, nv, table_z);
            cnt += c;
            if (c && first < 0) first = i;
            if (ptcol) ptcol[(size_t)n * H + i] = (uint8_t)c;
        }
        valid[n] = (uint8_t)(cnt == 0);
        nchk[n] = first < 0 ? H : first + 1;
        ncol[n] = cnt;
    }
}
