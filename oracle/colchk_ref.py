"""Python face of the collision oracle.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

* `maze2d_swept` / `maze2d_points` / `kuka_points`: ctypes calls into oracle/colchk.c
  (the bit-exact C restatement; reference file:line citations live there).
* `kuka_sphere_table()`: the link-sphere table of the self-defined analytic Kuka model,
  derived from the link collision-mesh AABBs measured in SURVEY.md Appendix D
  (pb_diff_envs/data/robot/kuka_iiwa/meshes/link_i.stl).
* `pick_first_valid`: candidate selection of rm2d_plan_env.py:137-235 / kuka_plan_env.py:133-200.
* `pick_multi_horizon`: the composed planner's selection over several horizons (comp_plan_env.py:126-265).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle_colchk.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _boxes(cen, hext):
    cen = np.ascontiguousarray(cen, np.float64)
    if cen.ndim == 2:
        cen = cen[None]
    hext = np.broadcast_to(np.asarray(hext, np.float64), cen.shape).copy()
    return cen, hext


def maze2d_swept(traj, env_id, cen, hext, eps=0.08, margin=0.01, legacy_div=True, lo=(0, 0), hi=(5, 5)):
    """traj [N,H,2] float32 unnormalised; cen [E,nw,2]; returns (valid u8 [N], nchk i32 [N])"""
    traj = np.ascontiguousarray(traj, np.float32)
    N, H, _ = traj.shape
    env_id = np.ascontiguousarray(env_id, np.int32)
    cen, hext = _boxes(cen, hext)
    lo = np.asarray(lo, np.float32); hi = np.asarray(hi, np.float32)
    valid = np.zeros(N, np.uint8); nchk = np.zeros(N, np.int32)
    lib().pmp_oracle_maze2d_swept(_p(traj, ctypes.c_float), _p(env_id, ctypes.c_int32), N, H,
                                  _p(cen, ctypes.c_double), _p(hext, ctypes.c_double), cen.shape[1],
                                  _p(lo, ctypes.c_float), _p(hi, ctypes.c_float),
                                  ctypes.c_double(eps), ctypes.c_double(margin), int(legacy_div),
                                  _p(valid, ctypes.c_uint8), _p(nchk, ctypes.c_int32))
    return valid, nchk


def maze2d_points(traj, env_id, cen, hext, margin=0.01, lo=(0, 0), hi=(5, 5)):
    traj = np.ascontiguousarray(traj, np.float32)
    N, H, _ = traj.shape
    env_id = np.ascontiguousarray(env_id, np.int32)
    cen, hext = _boxes(cen, hext)
    lo = np.asarray(lo, np.float32); hi = np.asarray(hi, np.float32)
    ncol = np.zeros(N, np.int32); pt = np.zeros((N, H), np.uint8)
    lib().pmp_oracle_maze2d_points(_p(traj, ctypes.c_float), _p(env_id, ctypes.c_int32), N, H,
                                   _p(cen, ctypes.c_double), _p(hext, ctypes.c_double), cen.shape[1],
                                   _p(lo, ctypes.c_float), _p(hi, ctypes.c_float), ctypes.c_double(margin),
                                   _p(ncol, ctypes.c_int32), _p(pt, ctypes.c_uint8))
    return ncol, pt


# link AABBs (min, max) in the link frame, SURVEY.md Appendix D [probe of the STL meshes]
KUKA_LINK_AABB = [
    ((-0.136, -0.121, 0.000), (0.121, 0.121, 0.158)),
    ((-0.086, -0.116, -0.010), (0.086, 0.085, 0.288)),
    ((-0.085, -0.086, -0.068), (0.086, 0.205, 0.132)),
    ((-0.068, -0.068, -0.010), (0.068, 0.115, 0.283)),
    ((-0.068, -0.068, -0.068), (0.068, 0.185, 0.114)),
    ((-0.068, -0.068, -0.010), (0.068, 0.099, 0.267)),
    ((-0.066, -0.086, -0.071), (0.066, 0.081, 0.067)),
    ((-0.052, -0.052, -0.010), (0.052, 0.052, 0.045)),
]


def kuka_sphere_table():
    """[n,5] float32 rows (link, cx, cy, cz, r): spheres strung along the longest AABB axis of
    every link, radius = half the mean of the two short extents (capped at 0.10 m)."""
    rows = []
    for l, (mn, mx) in enumerate(KUKA_LINK_AABB):
        mn = np.array(mn); mx = np.array(mx)
        ext = mx - mn
        ax = int(np.argmax(ext))
        others = [i for i in range(3) if i != ax]
        r = min(0.10, 0.25 * (ext[others[0]] + ext[others[1]]))
        n = max(1, int(round(ext[ax] / (1.5 * r))))
        reff = min(r, ext[ax] / 2)
        ctr = 0.5 * (mn + mx)
        for j in range(n):
            c = ctr.copy()
            if n > 1:
                c[ax] = mn[ax] + reff + (ext[ax] - 2 * reff) * j / (n - 1)
            rows.append([l, c[0], c[1], c[2], r])
    return np.array(rows, np.float32)


KUKA_TABLE_Z = -0.001


def kuka_bases(n_arms):
    if n_arms == 1:
        return np.zeros((1, 3), np.float32)
    return np.array([[-0.5, 0, 0], [0.5, 0, 0]], np.float32)


def kuka_points(traj, env_id, boxc, hext, n_arms=1, sph=None):
    """traj [N,H,7*n_arms] float32 joint angles; boxc [E,nv,3]; returns dict(valid, nchk, ncol, ptcol)"""
    traj = np.ascontiguousarray(traj, np.float32)
    N, H, _ = traj.shape
    env_id = np.ascontiguousarray(env_id, np.int32)
    boxc = np.ascontiguousarray(boxc, np.float32)
    if boxc.ndim == 2:
        boxc = boxc[None]
    boxh = np.broadcast_to(np.asarray(hext, np.float32), boxc.shape).copy()
    sph = kuka_sphere_table() if sph is None else np.ascontiguousarray(sph, np.float32)
    base = kuka_bases(n_arms)
    valid = np.zeros(N, np.uint8); nchk = np.zeros(N, np.int32); ncol = np.zeros(N, np.int32)
    pt = np.zeros((N, H), np.uint8)
    lib().pmp_oracle_kuka_points(_p(traj, ctypes.c_float), _p(env_id, ctypes.c_int32), N, H, n_arms,
                                 _p(base, ctypes.c_float), _p(sph, ctypes.c_float), len(sph),
                                 _p(boxc, ctypes.c_float), _p(boxh, ctypes.c_float), boxc.shape[1],
                                 ctypes.c_float(KUKA_TABLE_Z),
                                 _p(valid, ctypes.c_uint8), _p(nchk, ctypes.c_int32),
                                 _p(ncol, ctypes.c_int32), _p(pt, ctypes.c_uint8))
    return dict(valid=valid, nchk=nchk, ncol=ncol, ptcol=pt)


def kuka_fixed_transforms():
    out = np.zeros((7, 12), np.float32)
    lib().pmp_oracle_kuka_fixed(_p(out, ctypes.c_float))
    return out


def pick_first_valid(valid, nchk, n_prob, n_samp):
    """first valid candidate per problem else the last; returns (chosen idx [n_prob],
    success [n_prob] bool, num_colck [n_prob]) -- rm2d_plan_env.py:184-226"""
    v = np.asarray(valid).reshape(n_prob, n_samp).astype(bool)
    c = np.asarray(nchk).reshape(n_prob, n_samp)
    chosen = np.zeros(n_prob, np.int64); suc = np.zeros(n_prob, bool); tot = np.zeros(n_prob, np.int64)
    for p in range(n_prob):
        idx = n_samp - 1
        for s in range(n_samp):
            tot[p] += c[p, s]
            if v[p, s]:
                idx = s; suc[p] = True
                break
        chosen[p] = idx
    return chosen, suc, tot


def pick_multi_horizon(valid_list, nchk_list, n_prob, n_samp):
    """Composed planner, several planning horizons (comp_plan_env.py:126-265): the candidates of problem p are
    visited sample-major, horizon-minor -- (sample 0, horizon 0), (sample 0, horizon 1), ..., (sample 1,
    horizon 0), ... (`traj_a_prob`, :205-210) -- and the scan stops at the first collision-free one (:217-238);
    when none is, the LAST candidate visited is kept (:247-249).  valid_list[h] / nchk_list[h] hold the
    checker outputs of horizon h for its n_prob*n_samp candidates (problem-major, as the policy batch is laid
    out, :143-146).  Returns (horizon index [n_prob], sample index [n_prob], success [n_prob] bool,
    num_colck [n_prob]).  Pinned: tests/golden/pick_multi_horizon.npz holds what the reference method itself chose."""
    nh = len(valid_list)
    v = [np.asarray(a).reshape(n_prob, n_samp).astype(bool) for a in valid_list]
    c = [np.asarray(a).reshape(n_prob, n_samp) for a in nchk_list]
    hsel = np.zeros(n_prob, np.int64); ssel = np.zeros(n_prob, np.int64)
    suc = np.zeros(n_prob, bool); tot = np.zeros(n_prob, np.int64)
    for p in range(n_prob):
        done = False
        for j in range(n_samp):
            for h in range(nh):
                tot[p] += c[h][p, j]
                hsel[p], ssel[p] = h, j
                if v[h][p, j]:
                    suc[p] = True
                    done = True
                    break
            if done:
                break
    return hsel, ssel, suc, tot
