"""Batched planning + scoring of one or many environments on the GPU.

Covers, for static Maze2D and (sphere-model) Kuka environments, the per-environment pipeline of
  RandMaze2DEnvPlanner.plan_an_env / option_1_pick_1_traj   diffuser/guides/rm2d_plan_env.py:54-235
  KukaEnvPlanner.plan_an_env                                diffuser/guides/kuka_plan_env.py:54-200
  check_collision / compute_kuka_avg_result                 diffuser/guides/kuka_plan_utils.py:110-220
i.e. repeat every problem `samples_perprob` times, sample all candidates in one batch, take the
first collision-free candidate of every problem (else the last), score the chosen paths with the
per-waypoint check and aggregate the reference's statistics keys.  All environments of a call are
planned in ONE batch (the reference loops over environments)."""
import time

import numpy as np
import torch

import pmp_b200
from .kuka_colchk import KukaCollisionChecker, arm_bases, TABLE_Z


def compute_avg_result(ncol, norepl_suc, num_colck, n_env, n_prob, batch_runtime, repl_suc=None):
    """global statistics of a planning run, with the reference's keys and its arithmetic in its order: per
    environment the rates of check_collision (kuka_plan_utils.py:136-146: traj_srate = no_collision_rate =
    #(collision count == 0) / n, se_valid_rate = 1), then the sample-weighted sums and ratios of
    compute_kuka_avg_result (:175-213).  ncol [n_env*n_prob] per-waypoint collision counts of the chosen paths,
    norepl_suc / repl_suc [n_env*n_prob] success flags of the candidate scan / of replanning, num_colck
    [n_env*n_prob] collision checks spent; host arrays (a few bytes per problem).  batch_runtime is the time of the
    one batch that planned all environments (the reference sums one batch per environment)."""
    ncol = np.asarray(ncol).reshape(n_env, n_prob)
    norepl = np.asarray(norepl_suc).astype(bool).reshape(n_env, n_prob)
    repl = np.zeros((n_env, 0), bool) if repl_suc is None else np.asarray(repl_suc).astype(bool).reshape(n_env, -1)
    colck = np.asarray(num_colck).astype(np.int64).reshape(n_env, n_prob)
    total_num = 0
    w_srate = w_svr = w_ncr = 0
    n_norepl = n_repl = n_colck = 0
    for e in range(n_env):
        ok = ncol[e] == 0
        srate = ok.sum() / n_prob
        ncr = ok.sum() / n_prob
        svr = np.ones(n_prob, bool).sum() / n_prob
        total_num += n_prob
        w_srate += srate * n_prob
        w_ncr += ncr * n_prob
        w_svr += svr * n_prob
        n_norepl += norepl[e].sum()
        n_repl += repl[e].sum()
        n_colck += colck[e].sum().item()
    return dict(num_samples=total_num,
                traj_srate=float(w_srate / total_num),
                traj_norepl_srate=float(n_norepl / total_num),
                traj_repl_srate=float(n_repl / total_num),
                total_ncr=float(w_ncr / total_num),
                total_svr=float(w_svr / total_num),
                avg_batch_time=batch_runtime / n_env,
                avg_sample_time=batch_runtime / total_num,
                total_num_colck=int(n_colck),
                avg_num_colck=n_colck / total_num)


def plan_envs(policy, starts, goals, walls, env_boxes, hext, horizon, samples_perprob=10, robot='maze2d',
              n_arms=1, collision_eps=0.08, legacy_div=True):
    """starts/goals [n_env, n_prob, D] (un-normalised numpy), walls [n_env, n_prob, wall_len] model input,
    env_boxes [n_env, n_boxes, dim] obstacle centres for the checker, hext half extent.
    Returns (result dict with the reference's keys, chosen paths [n_env*n_prob, H, D] numpy)."""
    n_env, n_prob, D = starts.shape
    n_p, n_s = n_env * n_prob, samples_perprob
    rep = lambda a: np.repeat(a.reshape(n_p, -1), n_s, axis=0)
    cond = {0: rep(starts).astype(np.float32), horizon - 1: rep(goals).astype(np.float32)}
    wl = torch.as_tensor(rep(walls), dtype=torch.float32)
    dev = policy.device
    torch.cuda.synchronize(dev)
    tic = time.time()
    _, traj = policy(cond, batch_size=-1, wall_locations=wl)          # [n_p*n_s, H, D] numpy, un-normalised
    batch_runtime = time.time() - tic
    obs = torch.as_tensor(traj.observations, dtype=torch.float32, device=dev)
    env_id = np.repeat(np.arange(n_env, dtype=np.int32), n_prob * n_s)
    if robot == 'maze2d':
        valid, nchk = pmp_b200.colchk_maze2d_swept(obs, env_id, env_boxes, hext, eps=collision_eps,
                                                   legacy_div=legacy_div)
    else:
        chk = KukaCollisionChecker(device=dev)
        r = pmp_b200.colchk_kuka(obs, env_id, env_boxes, hext, n_arms, arm_bases(n_arms), chk.sph, TABLE_Z)
        valid, nchk = r['valid'], r['nchk']
    chosen, suc, total = pmp_b200.pick_first_valid(valid, nchk, n_p, n_s)
    idx = torch.arange(n_p, device=dev) * n_s + chosen.long()
    paths = obs[idx]
    env_p = np.repeat(np.arange(n_env, dtype=np.int32), n_prob)
    if robot == 'maze2d':
        ncol, _ = pmp_b200.colchk_maze2d_points(paths, env_p, env_boxes, hext)
    else:
        ncol = pmp_b200.colchk_kuka(paths, env_p, env_boxes, hext, n_arms, arm_bases(n_arms), chk.sph, TABLE_Z)['ncol']
    res = compute_avg_result(ncol.cpu().numpy(), suc.cpu().numpy(), total.cpu().numpy(), n_env, n_prob, batch_runtime)
    return res, paths.cpu().numpy()


def pick_multi_horizon(valid_list, nchk_list, n_prob, n_samp):
    """GPU-side candidate selection of the composed planner over several horizons
    (ComposedEnvPlanner.option_1_pick_1_traj, diffuser/guides/comp_plan_env.py:126-265): per problem the
    candidates are scanned sample-major / horizon-minor (:205-210) and the first collision-free one wins, else
    the last one visited.  valid_list[h], nchk_list[h]: device tensors [n_prob*n_samp] of horizon h.
    Returns device tensors (horizon index, sample index, success, num_colck), each [n_prob]."""
    nh = len(valid_list)
    v = torch.stack([a.reshape(n_prob, n_samp) for a in valid_list], dim=2).reshape(n_prob, n_samp * nh)
    c = torch.stack([a.reshape(n_prob, n_samp) for a in nchk_list], dim=2).reshape(n_prob, n_samp * nh)
    chosen, suc, total = pmp_b200.pick_first_valid(v.contiguous(), c.contiguous(), n_prob, n_samp * nh)
    chosen = chosen.long()
    return chosen % nh, chosen // nh, suc, total


def plan_envs_multi_horizon(policy, starts, goals, walls, env_boxes, hext, horizons, samples_perprob=10,
                            robot='maze2d', n_arms=1, collision_eps=0.08, legacy_div=True):
    """plan_envs for the composed planner's list of planning horizons (comp_plan_env.py:54-124): every problem is
    planned at every horizon (PolicyCompose.call_multi_horizon), all candidates are collision-checked on the
    GPU and one path per problem is selected with pick_multi_horizon.  Returns (result dict, list of chosen
    paths -- numpy [H_p, D] each, horizons differ per problem)."""
    n_env, n_prob, D = starts.shape
    n_p, n_s = n_env * n_prob, samples_perprob
    rep = lambda a: np.repeat(a.reshape(n_p, -1), n_s, axis=0)
    wl = torch.as_tensor(rep(walls), dtype=torch.float32)
    dev = policy.device
    env_id = np.repeat(np.arange(n_env, dtype=np.int32), n_prob * n_s)
    if robot != 'maze2d':
        chk = KukaCollisionChecker(device=dev)
    torch.cuda.synchronize(dev)
    tic = time.time()
    obs_h, valid_h, nchk_h = [], [], []
    for ho in horizons:
        cond = {0: rep(starts).astype(np.float32), ho - 1: rep(goals).astype(np.float32)}
        if hasattr(policy, 'call_multi_horizon'):
            traj = policy(cond, wl)[1]
        else:
            policy.diffusion_model.horizon = ho
            traj = policy(cond, batch_size=-1, wall_locations=wl)[1]
        obs = torch.as_tensor(traj.observations, dtype=torch.float32, device=dev)
        if robot == 'maze2d':
            valid, nchk = pmp_b200.colchk_maze2d_swept(obs, env_id, env_boxes, hext, eps=collision_eps,
                                                       legacy_div=legacy_div)
        else:
            r = pmp_b200.colchk_kuka(obs, env_id, env_boxes, hext, n_arms, arm_bases(n_arms), chk.sph, TABLE_Z)
            valid, nchk = r['valid'], r['nchk']
        obs_h.append(obs); valid_h.append(valid); nchk_h.append(nchk)
    hsel, ssel, suc, total = pick_multi_horizon(valid_h, nchk_h, n_p, n_s)
    torch.cuda.synchronize(dev)
    batch_runtime = time.time() - tic
    hs, ss = hsel.cpu().numpy(), ssel.cpu().numpy()
    paths = [obs_h[hs[p]][p * n_s + ss[p]].cpu().numpy() for p in range(n_p)]
    # final score: per-waypoint check of the chosen path (kuka_plan_utils.py:110-154), batched per horizon
    env_p = np.repeat(np.arange(n_env, dtype=np.int32), n_prob)
    ncol = np.zeros(n_p, np.int64)
    for i, ho in enumerate(horizons):
        sel = np.nonzero(hs == i)[0]
        if not len(sel):
            continue
        pt = obs_h[i][torch.as_tensor(sel * n_s + ss[sel], device=dev)]
        if robot == 'maze2d':
            nc, _ = pmp_b200.colchk_maze2d_points(pt, env_p[sel], env_boxes, hext)
        else:
            nc = pmp_b200.colchk_kuka(pt, env_p[sel], env_boxes, hext, n_arms, arm_bases(n_arms), chk.sph, TABLE_Z)['ncol']
        ncol[sel] = nc.cpu().numpy()
    res = compute_avg_result(ncol, suc.cpu().numpy(), total.cpu().numpy(), n_env, n_prob, batch_runtime)
    res['horizon_hist'] = [int((hs == i).sum()) for i in range(len(horizons))]
    return res, paths
