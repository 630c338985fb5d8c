"""KukaDiffusionReplanner: drop-in for diffuser/guides/kuka_replan.py:10-253, batched.

The reference repairs ONE colliding trajectory at a time: find the colliding waypoints, grow every
colliding run into a section whose length suits the U-Net (`get_expanded_subtraj`), re-noise the whole
trajectory to `dn_steps` and denoise it again (`GaussianDiffusionPB.ddim_replan` / `replan`), and accept a
re-planned section when it still connects to the untouched part (`check_start_end_repl` or
`check_start_end_repl_v2`) and is collision free; up to `n_replan_trial` trials.

Here `replan_batch` runs that procedure for ALL problems of a batch together: one per-waypoint collision
launch finds the sections, every trial is ONE batched `ddim_replan` over the still-failing problems, and the
connection / section checks of all candidates of a trial are a handful of collision-kernel launches (grouped
by length).  `replan_collision_section_less_check` keeps the reference's single-trajectory signature and is the
B = 1 case.  Decision logic and check counts follow the reference line by line (restated, with its file:line,
in oracle/replan_ref.py, against which the tests compare bit for bit).

Geometry is passed as a dict:  robot 'maze2d' | 'kuka', env_id [B] int, cen [E,nw,dim] obstacle centres,
hext half extent, n_arms (kuka).  Points fed to the checker are float32 (as trajectories are): the float64
interpolation of SolutionInterp is rounded once before the box test."""
import numpy as np
import torch

import pmp_b200
from .kuka_colchk import KukaCollisionChecker, arm_bases, TABLE_Z


def split_sorted_array_into_subarrays(arr):
    """[1,2,5,6,7] -> [array([1,2]), array([5,6,7])]   (kuka_replan.py:204-223)"""
    out, cur = [], [arr[0]]
    for i in range(1, len(arr)):
        if arr[i] == arr[i - 1] + 1:
            cur.append(arr[i])
        else:
            out.append(np.array(cur))
            cur = [arr[i]]
    out.append(np.array(cur))
    return out


def get_se_thres(traj):
    """closeness threshold of a re-planned section's ends (kuka_replan.py:245-253)"""
    d = traj.shape[-1]
    if d == 2:
        return 0.15
    if d in (7, 14):
        return 0.4
    raise NotImplementedError


def expanded_subtraj(traj_len, st, end, div_factor=4):
    """grow [st, end] alternately at both ends until its length is a multiple of 4, > 4 and > the input length
    (kuka_replan.py:176-198)"""
    flag = True
    input_len = end - st + 1
    n = input_len
    while n % div_factor != 0 or n <= div_factor or n <= input_len:
        if flag:
            st = max(st - 1, 0)
            flag = False
        else:
            end = min(end + 1, traj_len - 1)
            flag = True
        n = end - st + 1
    return np.arange(st, end + 1)


def sol_interp_pair(a, b, density):
    """SolutionInterp()([a, b]) (pb_diff_envs/utils/kuka_utils_luo.py:10-55): n_insert = ceil(|b-a|_1 * density / 2)
    points between the two ends (none when |b-a|_2 <= 1e-3), float64 arithmetic, both ends included"""
    diff = b - a
    if np.linalg.norm(diff) > 1e-3:
        n_insert = int(np.ceil(np.abs(diff).sum() * density / 2).astype(np.int32))
    else:
        n_insert = 0
    x = np.linspace(0, 1, n_insert + 2).reshape(-1, 1)
    return a + x * diff.reshape(1, -1)


class _GpuGeometry:
    """the two checker primitives the replanner needs, batched on the GPU"""

    def __init__(self, geom, device, collision_eps, use_collision_eps, legacy_div):
        self.g = geom
        self.dev = device
        self.eps = collision_eps
        self.use_eps = use_collision_eps
        self.legacy_div = legacy_div
        self.sph = KukaCollisionChecker(device=device).sph if geom['robot'] != 'maze2d' else None

    def points(self, trajs, env_ids):
        """per-waypoint collision bits [N, L] (numpy u8)"""
        t = torch.as_tensor(np.ascontiguousarray(trajs, np.float32), device=self.dev)
        if self.g['robot'] == 'maze2d':
            _, pt = pmp_b200.colchk_maze2d_points(t, env_ids, self.g['cen'], self.g['hext'])
        else:
            n_arms = self.g.get('n_arms', 1)
            pt = pmp_b200.colchk_kuka(t, env_ids, self.g['cen'], self.g['hext'], n_arms, arm_bases(n_arms), self.sph,
                                      TABLE_Z)['ptcol']
        return pt.cpu().numpy()

    def check_traj(self, trajs, env_ids):
        """checker.check_single_traj on every row: (no_collision [N] bool, count [N]) -- swept with collision_eps for
        the Maze2D checker of the planners (comp_plan_env.py:170-175), per-waypoint with early exit otherwise"""
        if self.g['robot'] == 'maze2d' and self.use_eps:
            t = torch.as_tensor(np.ascontiguousarray(trajs, np.float32), device=self.dev)
            v, c = pmp_b200.colchk_maze2d_swept(t, env_ids, self.g['cen'], self.g['hext'], eps=self.eps,
                                                legacy_div=self.legacy_div)
            return v.cpu().numpy().astype(bool), c.cpu().numpy().astype(np.int64)
        pt = self.points(trajs, env_ids)
        L = pt.shape[1]
        anyc = pt.any(1)
        first = np.where(anyc, pt.argmax(1) + 1, L)
        return ~anyc, first.astype(np.int64)


class KukaDiffusionReplanner:
    """replan given trajectories (reference ctor: kuka_replan.py:14-30)"""

    def __init__(self, model, normalizer, repl_config, device):
        self.model = model
        self.gap = 3
        self.use_ddim = repl_config['use_ddim']
        self.dn_steps = torch.full((1,), fill_value=repl_config['dn_steps'], device=device, dtype=torch.long)
        self.n_replan_trial = repl_config.get('n_replan_trial', 5)
        self.normalizer = normalizer
        self.device = torch.device(device)
        if normalizer.observation_dim <= 3:   # maze2d
            self.collision_eps = 0.05
        else:
            self.collision_eps = 0.5
        self.checker = None
        self.num_colck = None
        self.interp_density = 8               # SolutionInterp(density=8) of both checkers (rm2d_colchk.py:17)

    def set_collision_checker(self, checker):
        self.checker = checker

    def get_expanded_subtraj(self, traj_len, st, end, div_factor=4):
        return expanded_subtraj(traj_len, st, end, div_factor)

    def preproc_traj(self, unnorm_traj):
        return torch.as_tensor(self.normalizer.normalize(unnorm_traj, 'observations'), dtype=torch.float32,
                               device=self.device)

    # ------------------------------------------------------------------ diffusion part
    def replan_whole_traj_batch(self, normed, walls_loc):
        """re-noise + denoise [A,H,D] normalised trajectories in one batch (kuka_replan.py:147-163)"""
        H = normed.shape[1]
        cond = {0: normed[:, 0], H - 1: normed[:, -1]}
        dn = int(self.dn_steps.item())
        if self.use_ddim:
            return self.model.ddim_replan(normed, cond, walls_loc, dn)
        return self.model.replan(normed, cond, walls_loc, dn)

    # ------------------------------------------------------------------ batched procedure
    def replan_batch(self, trajs, walls_loc, geom):
        """trajs [B,H,D] un-normalised numpy, walls_loc [B,wall_len] tensor, geom as in the module docstring.
        Returns (new trajs [B,H,D] numpy, success [B] bool, num_colck [B] int): per problem exactly what
        replan_collision_section_less_check returns (problems without any collision come back unchanged,
        success True)."""
        trajs = np.array(trajs, np.float32, copy=True)
        B, H, D = trajs.shape
        env_id = np.asarray(geom['env_id'], np.int32)
        chk = self.checker
        G = _GpuGeometry(geom, self.device, getattr(chk, 'collision_eps', 0.08), getattr(chk, 'use_collision_eps', True),
                         getattr(chk, 'legacy_div', True))
        walls_loc = torch.as_tensor(walls_loc, dtype=torch.float32).to(self.device)
        num_colck = np.full(B, H, np.int64)                       # the first full-horizon check (:47-49)
        pt = G.points(trajs, env_id)
        sections = []
        for b in range(B):
            c_list = np.nonzero(pt[b])[0]
            if len(c_list) == 0:
                sections.append([])
                continue
            runs = split_sorted_array_into_subarrays(c_list)
            sections.append([expanded_subtraj(H, int(r[0]), int(r[-1])) for r in runs])
        success = np.array([len(s) == 0 for s in sections])
        thres = get_se_thres(trajs[0])
        for _ in range(self.n_replan_trial):
            act = [b for b in range(B) if len(sections[b])]
            if not act:
                break
            normed = self.preproc_traj(trajs[act])
            new = self.replan_whole_traj_batch(normed, walls_loc[act])
            new_un = self.normalizer.unnormalize(new.detach().cpu().numpy(), 'observations').astype(np.float32)
            # Within a problem the reference handles its sections one after the other on the SAME array
            # (`new_traj = to_np(traj)` aliases it, kuka_replan.py:105), so a later section's closeness /
            # connection tests see earlier replacements: sections are therefore processed rank by rank, each
            # rank batched over all active problems.
            keep = {b: [] for b in act}
            for rank in range(max(len(sections[b]) for b in act)):
                cand = [(b, sections[b][rank], new_un[k][sections[b][rank]]) for k, b in enumerate(act)
                        if rank < len(sections[b])]
                c1 = np.zeros(len(cand), bool)
                conn = []                                        # (candidate, which end, interpolated points)
                for ci, (b, idx, sec) in enumerate(cand):
                    c1[ci] = bool((np.abs(sec[0] - trajs[b][idx[0]]) < thres).all() and
                                  (np.abs(sec[-1] - trajs[b][idx[-1]]) < thres).all())      # check_start_end_repl
                    if idx[0] > 0:
                        conn.append((ci, 0, sol_interp_pair(trajs[b][idx[0] - 1], sec[0], self.interp_density)))
                    if idx[-1] + 1 < H:
                        conn.append((ci, 1, sol_interp_pair(sec[-1], trajs[b][idx[-1] + 1], self.interp_density)))
                # connection checks (check_start_end_repl_v2 -> steerTo_repl -> check_single_traj), grouped by length
                c_ok = np.ones((len(cand), 2), bool)
                c_cnt = np.zeros((len(cand), 2), np.int64)
                for L in sorted({len(q) for _, _, q in conn}):
                    sel = [k for k, (_, _, q) in enumerate(conn) if len(q) == L]
                    ok, cnt = G.check_traj(np.stack([conn[k][2] for k in sel]), env_id[[cand[conn[k][0]][0] for k in sel]])
                    for k, o, c in zip(sel, ok, cnt):
                        c_ok[conn[k][0], conn[k][1]] = o
                        c_cnt[conn[k][0], conn[k][1]] = c
                c2 = c_ok[:, 0] & c_ok[:, 1]
                # the far end is only steered to when the near end connected (kuka_plan_utils.py:78)
                cnt2 = c_cnt[:, 0] + np.where(c_ok[:, 0], c_cnt[:, 1], 0)
                # section checks for the candidates that connect (kuka_replan.py:121-125)
                need = np.nonzero(c1 | c2)[0]
                sec_ok = np.zeros(len(cand), bool)
                sec_cnt = np.zeros(len(cand), np.int64)
                for L in sorted({len(cand[ci][2]) for ci in need}):
                    sel = [ci for ci in need if len(cand[ci][2]) == L]
                    ok, cnt = G.check_traj(np.stack([cand[ci][2] for ci in sel]), env_id[[cand[ci][0] for ci in sel]])
                    sec_ok[sel] = ok
                    sec_cnt[sel] = cnt
                # accept / keep (kuka_replan.py:127-141)
                for ci, (b, idx, sec) in enumerate(cand):
                    num_colck[b] += cnt2[ci] + sec_cnt[ci]
                    if (c1[ci] or c2[ci]) and sec_ok[ci]:
                        trajs[b][idx] = sec
                    else:
                        keep[b].append(idx)
            for b in act:
                sections[b] = keep[b]
                success[b] = len(keep[b]) == 0
        return trajs, success, num_colck

    # ------------------------------------------------------------------ reference signature (one trajectory)
    def replan_collision_section_less_check(self, env, traj, walls_loc):
        """(new_traj [H,D] numpy, is_success, num_colck) for one un-normalised trajectory (kuka_replan.py:33-75)"""
        assert traj.ndim == 2 and walls_loc.ndim == 2 and walls_loc.shape[0] == 1
        if hasattr(env, 'wall_centers'):
            cen = np.asarray(env.wall_centers, np.float64)[None]
            geom = dict(robot='maze2d', env_id=np.zeros(1, np.int32), cen=cen,
                        hext=np.broadcast_to(np.asarray(env.wall_hext, np.float64), cen.shape[1:])[None])
        else:
            geom = dict(robot='kuka', env_id=np.zeros(1, np.int32), cen=np.asarray(env.voxel_centers, np.float32)[None],
                        hext=env.voxel_hext, n_arms=env.n_arms)
        new, suc, cnt = self.replan_batch(traj[None], walls_loc, geom)
        return new[0], bool(suc[0]), int(cnt[0])


from .._overlay import reference_fallback as _reference_fallback  # noqa: E402

__getattr__ = _reference_fallback(__name__)   # overlay mode: symbols outside the hot path come from the reference's file
