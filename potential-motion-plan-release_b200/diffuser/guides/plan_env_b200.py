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


def collision_model_tag(robot, conservative=False):
    """Kuka / Dual-Kuka validity comes from the analytic sphere-link model, not from the reference's pybullet mesh contact
    query (declared deviation): result dicts say so, so that their `traj_srate` etc. are not read as
    reference-parity numbers"""
    if robot == 'maze2d':
        return 'maze2d boxes, exact reference arithmetic'
    if conservative:
        return ('conservative sphere-link model (spheres containing the link AABBs: an accepted path is collision free for the '
                'meshes too; NOT pybullet mesh contacts, success rates are lower bounds)')
    return 'sphere-link approximation of the arm (NOT pybullet mesh contacts; success rates are sphere-model rates)'


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


def replan_failed(replanner, candidates, walls_p, env_p, geom, suc, n_repl_max=6):
    """Replanning pass of the planners (KukaEnvPlanner.replan_pred_trajs, kuka_plan_env.py:207-260), batched: the
    reference walks the failed problems one by one and feeds up to six of a problem's candidates (scan order) to the
    replanner until one is repaired; here round r repairs candidate r of ALL still-failed problems with one
    `replan_batch` per trajectory length.  candidates: {problem: list of [H_j, D] numpy candidates in scan order} for
    the failed problems, walls_p [n_p, wall_len] model input per problem, env_p [n_p] environment of each problem,
    geom: dict(robot, cen, hext[, n_arms]) as `replan_batch` takes it (env_id is filled in), suc [n_p] success flags of
    the candidate scan.  Returns ({problem: repaired path}, success-after-replan [n_p] bool -- the reference's
    replan_suc_list, which starts as a copy of suc -- and the collision checks spent per problem [n_p])."""
    suc = np.asarray(suc).astype(bool)
    repl_suc = suc.copy()
    extra = np.zeros(len(suc), np.int64)
    new_paths = {}
    walls_p = torch.as_tensor(walls_p, dtype=torch.float32)
    active = [int(p) for p in np.nonzero(~suc)[0]]
    for r in range(n_repl_max):
        by_len = {}
        for p in active:
            if r < len(candidates[p]):
                by_len.setdefault(len(candidates[p][r]), []).append(p)
        if not by_len:
            break
        for ps in by_len.values():
            trajs = np.stack([np.asarray(candidates[p][r], np.float32) for p in ps])
            g = dict(geom, env_id=np.asarray(env_p)[ps].astype(np.int32))
            new, ok, cnt = replanner.replan_batch(trajs, walls_p[ps], g)
            for i, p in enumerate(ps):
                extra[p] += int(cnt[i])
                if ok[i]:
                    repl_suc[p] = True
                    new_paths[p] = np.asarray(new[i], np.float32)
        active = [p for p in active if not repl_suc[p]]
    return new_paths, repl_suc, extra


def plan_envs(policy, starts, goals, walls, env_boxes, hext, horizon, samples_perprob=10, robot='maze2d',
              n_arms=1, collision_eps=0.08, legacy_div=True, replanner=None, kuka_conservative=False):
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
        chk = KukaCollisionChecker(device=dev, conservative=kuka_conservative)
        r = pmp_b200.colchk_kuka(obs, env_id, env_boxes, hext, n_arms, arm_bases(n_arms), chk.sph, TABLE_Z)
        valid, nchk = r['valid'], r['nchk']
    chosen, suc, total = pmp_b200.pick_first_valid(valid, nchk, n_p, n_s)
    idx = torch.arange(n_p, device=dev) * n_s + chosen.long()
    paths = obs[idx]
    env_p = np.repeat(np.arange(n_env, dtype=np.int32), n_prob)
    repl_suc = None
    if replanner is not None:      # do_replan: repair the problems whose candidates all collide (kuka_plan_env.py:207-260)
        suc_h = suc.cpu().numpy().astype(bool)
        cands = {int(p): list(obs[p * n_s:(p + 1) * n_s].cpu().numpy()) for p in np.nonzero(~suc_h)[0]}
        geom = dict(robot=robot, cen=env_boxes, hext=hext, n_arms=n_arms)
        fixed, repl_suc, extra = replan_failed(replanner, cands, walls.reshape(n_p, -1), env_p, geom, suc_h)
        for p, path in fixed.items():
            paths[p] = torch.as_tensor(path, device=dev)
        total = total + torch.as_tensor(extra, device=dev).to(total.dtype)
    if robot == 'maze2d':
        ncol, _ = pmp_b200.colchk_maze2d_points(paths, env_p, env_boxes, hext)
    else:
        ncol = pmp_b200.colchk_kuka(paths, env_p, env_boxes, hext, n_arms, arm_bases(n_arms), chk.sph, TABLE_Z)['ncol']
    res = compute_avg_result(ncol.cpu().numpy(), suc.cpu().numpy(), total.cpu().numpy(), n_env, n_prob, batch_runtime,
                             repl_suc=repl_suc)
    res['collision_model'] = collision_model_tag(robot, kuka_conservative)
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
                            robot='maze2d', n_arms=1, collision_eps=0.08, legacy_div=True, replanner=None,
                            kuka_conservative=False):
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
        chk = KukaCollisionChecker(device=dev, conservative=kuka_conservative)
    torch.cuda.synchronize(dev)
    tic = time.time()
    obs_h, valid_h, nchk_h = [], [], []
    for ho in horizons:
        cond = {0: rep(starts).astype(np.float32), ho - 1: rep(goals).astype(np.float32)}
        if hasattr(policy, 'diffusion_model_list'):     # PolicyCompose (the base Policy also has call_multi_horizon)
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
    env_p = np.repeat(np.arange(n_env, dtype=np.int32), n_prob)
    repl_suc = None
    if replanner is not None:      # do_replan (comp_plan_env.py:112-114): candidates in scan order, sample-major
        suc_h = suc.cpu().numpy().astype(bool)
        cands = {int(p): [obs_h[i][p * n_s + j].cpu().numpy() for j in range(n_s) for i in range(len(horizons))]
                 for p in np.nonzero(~suc_h)[0]}
        geom = dict(robot=robot, cen=env_boxes, hext=hext, n_arms=n_arms)
        fixed, repl_suc, extra = replan_failed(replanner, cands, walls.reshape(n_p, -1), env_p, geom, suc_h)
        total = total + torch.as_tensor(extra, device=dev).to(total.dtype)
        hsel_fixed = {p: list(horizons).index(len(path)) for p, path in fixed.items()}
    # final score: per-waypoint check of the chosen path (kuka_plan_utils.py:110-154), batched per horizon
    ncol = np.zeros(n_p, np.int64)
    if replanner is not None:
        hs = hs.copy()
        for p, path in fixed.items():
            paths[p] = path
            hs[p] = hsel_fixed[p]
    for i, ho in enumerate(horizons):
        sel = np.nonzero(hs == i)[0]
        if not len(sel):
            continue
        pt = torch.as_tensor(np.stack([paths[p] for p in sel]), dtype=torch.float32, device=dev)
        if robot == 'maze2d':
            nc, _ = pmp_b200.colchk_maze2d_points(pt, env_p[sel], env_boxes, hext)
        else:
            nc = pmp_b200.colchk_kuka(pt, env_p[sel], env_boxes, hext, n_arms, arm_bases(n_arms), chk.sph, TABLE_Z)['ncol']
        ncol[sel] = nc.cpu().numpy()
    res = compute_avg_result(ncol, suc.cpu().numpy(), total.cpu().numpy(), n_env, n_prob, batch_runtime, repl_suc=repl_suc)
    res['horizon_hist'] = [int((hs == i).sum()) for i in range(len(horizons))]
    res['collision_model'] = collision_model_tag(robot, kuka_conservative)
    return res, paths
