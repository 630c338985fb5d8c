"""CPU restatement of the reference's trajectory replanner -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows, for ONE trajectory at a time and in the reference's own statement order:
  KukaDiffusionReplanner.replan_collision_section_less_check   diffuser/guides/kuka_replan.py:33-75
  KukaDiffusionReplanner.replan_one_time_less_check            diffuser/guides/kuka_replan.py:82-143
  get_expanded_subtraj :176-198, split_sorted_array_into_subarrays :204-223, get_se_thres :245-253
  check_start_end_repl / check_start_end_repl_v2 / check_single_traj   diffuser/guides/kuka_plan_utils.py:59-103
  KukaCollisionChecker.steerTo_repl                            diffuser/guides/kuka_colchk.py:127-143
  SolutionInterp.__call__                                      pb_diff_envs/utils/kuka_utils_luo.py:10-55
The diffusion call (model.ddim_replan, kuka_replan.py:147-163) is injected as `resample(trial, traj) -> new traj`
(un-normalised), so the deterministic part -- section finding, acceptance, check counts -- is compared bit for bit
with the batched GPU replanner on identical re-sampled trajectories.

Parity: **pinned** for Maze2D -- tests/golden/replan_maze2d.npz holds what the reference's own
KukaDiffusionReplanner + RM2DCollisionChecker + Point2DRobot (imported through oracle/refshim.load_replanner, the
diffusion call replaced by preset re-sampled trajectories) returned for 64 problems: new paths, success flags and
check counts agree bit for bit (tests/test_oracle_golden.py::test_replan_oracle_matches_golden).  The Kuka variant
differs only in the collision model, which is unpinned (pybullet, see oracle/__init__.py)."""
import numpy as np

from . import colchk_ref


def split_runs(arr):
    """kuka_replan.py:204-223"""
    out, cur = [], [arr[0]]
    for i in range(1, len(arr)):
        if arr[i] == arr[i - 1] + 1:
            cur.append(arr[i])
        else:
            out.append(np.array(cur)); cur = [arr[i]]
    out.append(np.array(cur))
    return out


def expanded_subtraj(traj_len, st, end, div_factor=4):
    """kuka_replan.py:176-198"""
    flag = True
    input_len = end - st + 1
    subtraj_len = end - st + 1
    while subtraj_len % div_factor != 0 or subtraj_len <= div_factor or subtraj_len <= input_len:
        if flag:
            st = max(st - 1, 0); flag = False
        else:
            end = min(end + 1, traj_len - 1); flag = True
        subtraj_len = end - st + 1
    return np.arange(st, end + 1)


def se_thres(traj):
    """kuka_replan.py:245-253"""
    return {2: 0.15, 7: 0.4, 14: 0.4}[traj.shape[-1]]


def sol_interp(solution, density=8):
    """kuka_utils_luo.py:10-55 for a list of waypoints"""
    sols = []
    n = len(solution)
    for i in range(n - 1):
        diff = solution[i + 1] - solution[i]
        diff_norm = np.linalg.norm(diff)
        diff_abs_sum = np.abs(diff).sum()
        n_insert = np.ceil(diff_abs_sum * density / 2).astype(np.int32) if diff_norm > 1e-3 else 0
        x = np.linspace(0, 1, n_insert + 2).reshape(-1, 1)
        if i != n - 2:
            x = x[:-1]
        sols.append(solution[i] + x * diff.reshape(1, -1))
    return np.concatenate(sols, 0)


class Maze2DChecker:
    """RM2DCollisionChecker with use_collision_eps (rm2d_colchk.py:49-93) + the robot's per-waypoint test, on the C
    restatement of the geometry (oracle/colchk.c)"""

    def __init__(self, cen, hext, eps=0.08, legacy_div=True):
        self.cen = np.asarray(cen, np.float64)[None]
        self.hext = hext
        self.eps = eps
        self.legacy_div = legacy_div

    def point_collides(self, traj):
        _, pt = colchk_ref.maze2d_points(np.asarray(traj, np.float32)[None], np.zeros(1, np.int32), self.cen, self.hext)
        return pt[0].astype(bool)

    def check_single_traj(self, traj):
        v, c = colchk_ref.maze2d_swept(np.asarray(traj, np.float32)[None], np.zeros(1, np.int32), self.cen, self.hext,
                                       eps=self.eps, legacy_div=self.legacy_div)
        return bool(v[0]), int(c[0])

    def steerTo_repl(self, start, end):
        """kuka_colchk.py:127-143"""
        return self.check_single_traj(sol_interp([start, end]))


def check_start_end_repl(traj, st, gl, thres):
    """kuka_plan_utils.py:59-61"""
    return bool((np.abs(traj[0] - st) < thres).all() and (np.abs(traj[-1] - gl) < thres).all())


def check_start_end_repl_v2(ori_traj, new_section, c_list_i, checker):
    """kuka_plan_utils.py:63-83"""
    valid, cnt = True, 0
    if c_list_i[0] > 0:
        v1, c1 = checker.steerTo_repl(ori_traj[c_list_i[0] - 1], new_section[0])
        valid &= v1; cnt += c1
    if valid and (c_list_i[-1] + 1 < len(ori_traj)):
        v2, c2 = checker.steerTo_repl(new_section[-1], ori_traj[c_list_i[-1] + 1])
        valid &= v2; cnt += c2
    return valid, cnt


def replan_one_time(traj, c_list, new_whole, checker):
    """kuka_replan.py:82-143 with the diffusion output `new_whole` (un-normalised [H,D]) given.  Modifies `traj`
    in place, as the reference does (`new_traj = to_np(traj)` is the same array)."""
    thres = se_thres(traj)
    new_traj = traj
    n_valid, c_list_after, num = 0, [], 0
    for i in range(len(c_list)):
        new_section = new_whole[c_list[i]]
        c1 = check_start_end_repl(new_section, traj[c_list[i][0]], traj[c_list[i][-1]], thres)
        c2, c_cnt2 = check_start_end_repl_v2(traj, new_section, c_list[i], checker)
        num += c_cnt2
        no_collision = False
        if c1 or c2:
            no_collision, c_cnt = checker.check_single_traj(new_section)
            num += c_cnt
        if (c1 or c2) and no_collision:
            new_traj[c_list[i]] = new_section
            n_valid += 1
        else:
            c_list_after.append(c_list[i])
    return new_traj, n_valid == len(c_list), c_list_after, num


def replan_collision_section(traj, checker, resample, n_replan_trial=5):
    """kuka_replan.py:33-75.  Returns (new_traj, is_success, num_colck); a trajectory without any colliding
    waypoint is returned unchanged (the reference asserts it never gets one)."""
    traj = np.array(traj, np.float32, copy=True)
    num_colck = len(traj)
    c_idx = np.nonzero(checker.point_collides(traj))[0]
    if len(c_idx) == 0:
        return traj, True, num_colck
    c_list = split_runs(c_idx)
    for i in range(len(c_list)):
        c_list[i] = expanded_subtraj(len(traj), int(c_list[i][0]), int(c_list[i][-1]))
    new_traj, is_success = traj, False
    for trial in range(n_replan_trial):
        new_whole = np.asarray(resample(trial, new_traj), np.float32)
        new_traj, is_success, c_list, num = replan_one_time(new_traj, c_list, new_whole, checker)
        num_colck += num
        if is_success:
            break
    return new_traj, is_success, num_colck


def replan_pred_trajs(candidates, suc, checker_of, resample, n_replan_trial=5, n_repl_max=6):
    """KukaEnvPlanner.replan_pred_trajs (diffuser/guides/kuka_plan_env.py:207-260), sequential as in the reference:
    for every problem whose candidate scan failed, up to min(6, #candidates) candidates are handed to the replanner
    one after the other until one is repaired.  candidates[p]: list of [H,D] arrays in scan order, checker_of(p):
    the problem's collision checker, resample(trial, traj) as in replan_collision_section.
    Returns ({p: repaired path}, replan_suc_list [n_p] bool, extra collision checks [n_p])."""
    suc = np.asarray(suc).astype(bool)
    repl = suc.copy()
    extra = np.zeros(len(suc), np.int64)
    fixed = {}
    for p in range(len(suc)):
        if suc[p]:
            continue
        ok = False
        for j in range(min(n_repl_max, len(candidates[p]))):
            new, ok, cnt = replan_collision_section(candidates[p][j], checker_of(p), resample, n_replan_trial)
            extra[p] += cnt
            if ok:
                break
        repl[p] = ok
        if ok:
            fixed[p] = new
    return fixed, repl, extra
