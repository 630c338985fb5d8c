"""Freezes golden vectors by running the *reference implementation itself* (imported from
/root/reference through oracle/refshim.py) on the deterministic synthetic inputs of
oracle/synth.py.  Only runnable in the build container; the .npz files it writes are
committed and travel to the GPU box.

    python tests/golden/gen_golden.py
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refshim, synth, unet_ref  # noqa: E402

torch.set_num_threads(8)
WEIGHT_SEED = 1


def build_ref(fam, seed=WEIGHT_SEED):
    U, G = refshim.load_models()
    a = synth.arch(fam)
    with refshim.quiet():
        m = U(**a["unet"]).eval()
        sd = unet_ref.synth_weights(a["unet"], seed)
        m.load_state_dict(sd)
        d = G(m, **a["diffusion"]).eval()
    return a, m, d, sd


def gen_unet():
    for fam in synth.FAMILIES:
        a, m, d, sd = build_ref(fam)
        B, H, D = 2, a["horizon"], a["obs_dim"]
        pr = synth.make_problems(fam, B, 3)
        g = torch.Generator().manual_seed(11)
        x = torch.randn(2 * B, H, D, generator=g)
        t = torch.tensor([37, 5, 37, 5])
        walls = torch.tensor(pr["walls"]).repeat(2, 1)
        with refshim.quiet():
            eps = m(x.clone(), None, t, walls, use_dropout=False, force_dropout=True, half_fd=True).detach()
        np.savez_compressed(os.path.join(HERE, "unet_eps_%s.npz" % fam), x=x.numpy(), t=t.numpy(), walls=walls.numpy(),
                            eps=eps.numpy(), n_cond=B, weight_seed=WEIGHT_SEED, digest=synth.state_dict_digest(sd))
        print("unet", fam, float(eps.abs().max()))


def gen_sample():
    for fam, B in (("m2d", 2), ("kuka", 1), ("dualkuka", 1)):
        a, m, d, sd = build_ref(fam)
        H, D = a["horizon"], a["obs_dim"]
        pr = synth.make_problems(fam, B, 3)
        lo, hi = pr["lo"], pr["hi"]
        start = torch.tensor(synth.normalize(pr["start"], lo, hi))
        goal = torch.tensor(synth.normalize(pr["goal"], lo, hi))
        walls = torch.tensor(pr["walls"])
        cond = {0: start, H - 1: goal}
        out = {}
        for name, ddim in (("ddpm", 0), ("ddim", 8 if fam != "dualkuka" else 10)):
            if fam != "m2d" and not ddim:
                continue   # keep generation time short: full DDPM-100 only for Maze2D
            n = ddim if ddim else 100
            torch.manual_seed(5)
            with refshim.quiet():
                d.horizon = H
                d.ddim_num_inference_steps = n
                x, tr = d.conditional_sample(cond, walls, use_ddim=bool(ddim), return_diffusion=True)
            torch.manual_seed(5)
            noise = torch.stack([torch.randn(B, H, D) for _ in range(n + 1 if not ddim else 1)])
            out[name + "_x"] = x.detach().numpy()
            out[name + "_trace"] = tr.detach().numpy()
            out[name + "_noise"] = noise.numpy()
            out[name + "_steps"] = n
        np.savez_compressed(os.path.join(HERE, "sample_%s.npz" % fam), start=start.numpy(), goal=goal.numpy(),
                            walls=walls.numpy(), weight_seed=WEIGHT_SEED, **out)
        print("sample", fam, list(out))


def gen_compose():
    """K=2 composition through the reference's own pred_unet / pred_x_tm1_* plus the six lines of
    policies_compose.py:286-297 (PolicyCompose itself hard-codes .to('cuda'))."""
    a1, m1, d1, sd1 = build_ref("m2d", 1)
    a2, m2, d2, sd2 = build_ref("m2d_nw3", 2)
    B, H, D = 2, 40, 2
    pr = synth.make_problems("m2d", B, 21)
    pr2 = synth.make_problems("m2d_nw3", B, 22)
    lo, hi = pr["lo"], pr["hi"]
    start = torch.tensor(synth.normalize(pr["start"], lo, hi))
    goal = torch.tensor(synth.normalize(pr["goal"], lo, hi))
    w1 = torch.tensor(pr["walls"]); w2 = torch.tensor(pr2["walls"])
    cg = torch.tensor([2.0, 1.5]).reshape(2, 1, 1, 1)
    out = {}
    for name, ddim, eta in (("ddim24", 24, 1.0), ("ddpm", 0, 0.0)):
        steps = [int(v) for v in d1.ddim_set_timesteps(24)] if ddim else list(range(19, -1, -1))
        torch.manual_seed(9)
        x = 1.0 * torch.randn(B, H, D)
        for dd in (d1, d2):
            dd.ddim_num_inference_steps = 24
        xs = []
        with refshim.quiet():
            for t in steps:
                tt = torch.full((B,), t, dtype=torch.long)
                ecs, eus = [], []
                for dd, wl in ((d1, w1), (d2, w2)):
                    (_, ec), (_, eu) = dd.pred_unet(x.detach(), {0: start, H - 1: goal}, tt, wl)
                    ecs.append(ec); eus.append(eu)
                ecs = torch.stack(ecs, 0); eus = torch.stack(eus, 0)
                comp = eus[0] + (cg * (ecs - eus)).sum(dim=0, keepdims=False)
                if ddim:
                    x = d1.pred_x_tm1_ddim(x, tt, comp, eta=eta, clip_denoised=True)
                else:
                    x = d1.pred_x_tm1_luo(x, tt, comp, clip_denoised=True)
                xs.append(x.clone())
        x[:, 0] = start; x[:, -1] = goal
        torch.manual_seed(9)
        noise = torch.stack([torch.randn(B, H, D) for _ in range(len(steps) + 1)])
        out[name + "_x"] = x.detach().numpy()
        out[name + "_noise"] = noise.numpy()
        out[name + "_steps"] = np.array(steps)
    np.savez_compressed(os.path.join(HERE, "compose_m2d.npz"), start=start.numpy(), goal=goal.numpy(), walls1=w1.numpy(),
                        walls2=w2.numpy(), w=np.array([2.0, 1.5], np.float32), seeds=np.array([1, 2]), **out)
    print("compose", list(out))


def gen_colchk():
    P, RW, RWG, Env = refshim.load_maze2d_geometry()
    rng = np.random.default_rng(5)
    out = {}
    for fam, hext in (("m2d", 0.5), ("m2d_nw3", 0.7), ("m2d_concave", 0.25)):
        E, n_per, H = 12, 25, 48
        pr = synth.make_problems(fam, E, 9)
        al = np.linspace(0, 1, H)[None, None, :, None].astype(np.float32)
        s = pr["start"][:, None, None, :]; g = pr["goal"][:, None, None, :]
        sig = rng.choice([0.02, 0.1, 0.3], size=(E, n_per, 1, 1))
        tr = s * (1 - al) + g * al + (rng.normal(0, 1, size=(E, n_per, H, 2)) * sig * np.sin(np.pi * al)).astype(np.float32)
        tr = np.clip(tr, -0.02, 5.02).astype(np.float32).reshape(E * n_per, H, 2)
        env_id = np.repeat(np.arange(E, dtype=np.int32), n_per)
        valid = np.zeros(len(tr), np.uint8); nchk = np.zeros(len(tr), np.int32)
        ncol = np.zeros(len(tr), np.int32)
        for i in range(len(tr)):
            e = env_id[i]
            walls = [RW(np.array(c), np.array([hext, hext])) for c in pr["boxes"][e]]
            rob = P(np.array([5., 5.]), RWG(walls), 0.01, collision_eps=0.08)
            env = object.__new__(Env); env.robot = rob
            cnt, ok = 0, True
            with contextlib.redirect_stdout(io.StringIO()):
                for j in range(H - 1):          # rm2d_colchk.py:76-93
                    rob.collision_check_count = 0
                    good = env.edge_fp(tr[i, j], tr[i, j + 1])
                    cnt += rob.collision_check_count
                    if not good:
                        ok = False
                        break
                trc = np.clip(tr[i], 0, 5)
                c2 = 0
                for j in range(H):              # kuka_plan_utils.py:87-103
                    rob.set_config(trc[j])
                    if not rob.no_collision():
                        c2 += 1
            valid[i] = ok; nchk[i] = cnt; ncol[i] = c2
        out[fam + "_traj"] = tr; out[fam + "_env"] = env_id; out[fam + "_boxes"] = pr["boxes"]
        out[fam + "_hext"] = hext; out[fam + "_valid"] = valid; out[fam + "_nchk"] = nchk; out[fam + "_ncol"] = ncol
        print("colchk", fam, "valid frac", valid.mean(), "mean checks", nchk.mean())
    out["numpy_version"] = np.__version__
    out["legacy_div"] = 0   # numpy >= 2 semantics produced these counts
    np.savez_compressed(os.path.join(HERE, "colchk_maze2d.npz"), **out)


def gen_concave():
    """layouts drawn by the reference's own concave generator (sample_one_valid_env with the registration kwargs of
    randSmaze2d-C43-ng3ks25k-ms55nw21-hExt05-v0, init_maze2d_gym_43.py:21-52) and the (cx, cy, mode) rows its
    save_center_and_mode returns for them"""
    refshim.load_maze2d_geometry()
    from pb_diff_envs.environment.rand_rec_group_43 import RandRectangleGroup_43 as G
    g = object.__new__(G)
    g.num_walls = 21; g.rand_iter_limit = 1e5; g.two = 2
    g.maze_size = np.array([5, 5]); g.half_extents = np.array([0.25, 0.25]); g.gap_betw_wall = 0.15
    g.rng = np.random.default_rng(333)
    layouts, modes0 = [], []
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(24):
            pos, _hext, env_modes = g.sample_one_valid_env()
            layouts.append(pos); modes0.append(env_modes)
    g.rec_loc_grp_list = np.array(layouts)
    tokens = g.save_center_and_mode(is_save=False)
    np.savez_compressed(os.path.join(HERE, "concave_layouts.npz"), layouts=g.rec_loc_grp_list, tokens=tokens,
                        sampled_modes=np.array(modes0), hext=0.25)
    print("concave", g.rec_loc_grp_list.shape, tokens.shape, "modes", np.bincount(tokens[..., 2].astype(int).ravel()))


def gen_resample():
    """the reference's GaussianDiffusionPB.ddim_replan (diffusion_pb.py:311-346: q_sample to the dn_steps-th last DDIM
    timestep, then dn_steps DDIM steps with eta=0) on four Maze2D trajectories, one call each as the replanner does;
    the q_sample noise is injected so that the product can be fed the same"""
    fam, B, n, dn = "m2d", 4, 8, 3
    a, m, d, sd = build_ref(fam)
    H, D = a["horizon"], a["obs_dim"]
    pr = synth.make_problems(fam, B, 17)
    lo, hi = pr["lo"], pr["hi"]
    rng = np.random.default_rng(21)
    al = np.linspace(0, 1, H)[None, :, None]
    tr = pr["start"][:, None] * (1 - al) + pr["goal"][:, None] * al + rng.normal(0, 0.1, (B, H, D)) * np.sin(np.pi * al)
    x0 = torch.tensor(synth.normalize(np.clip(tr, lo, hi).astype(np.float32), lo, hi))
    walls = torch.tensor(pr["walls"])
    noise = torch.randn(B, H, D, generator=torch.Generator().manual_seed(9))
    d.horizon = H; d.ddim_num_inference_steps = n
    q_sample = d.q_sample
    out, noisy = [], []
    for b in range(B):
        d.q_sample = lambda x_start, t, noise=None, b=b: q_sample(x_start, t, noise=noise_b[b])
        noise_b = noise[:, None]
        cond = {0: x0[b, None, 0], H - 1: x0[b, None, -1]}
        with refshim.quiet():
            out.append(d.ddim_replan(x0[b], cond, walls[b, None], torch.full((1,), dn, dtype=torch.long)).detach())
            noisy.append(d.ddim_replan(x0[b], cond, walls[b, None], torch.full((1,), dn, dtype=torch.long), ret_noisy_traj=True)[0])
    d.q_sample = q_sample
    np.savez_compressed(os.path.join(HERE, "resample_m2d.npz"), x0=x0.numpy(), walls=walls.numpy(), noise=noise.numpy(),
                        x=torch.stack(out).numpy(), noisy=torch.stack(noisy).numpy(), ddim_steps=n, dn_steps=dn,
                        weight_seed=WEIGHT_SEED)
    print("resample: max |x - x0| %.3f" % float((torch.stack(out) - x0).abs().max()))


def gen_resample_compose():
    """the reference's PolicyCompose.ddim_replan (policies_compose.py:399-447) for K=2 (6 walls + 3 walls), one
    trajectory per call; its constructor hard-codes .to('cuda'), so the object is built with object.__new__ and the
    attributes the method reads"""
    PC = refshim._import_ref("diffuser.guides.policies_compose")[0].PolicyCompose
    a1, m1, d1, sd1 = build_ref("m2d", 1)
    a2, m2, d2, sd2 = build_ref("m2d_nw3", 2)
    B, H, D, n, dn = 3, 48, 2, 8, 3
    pr = synth.make_problems("m2d", B, 31); pr2 = synth.make_problems("m2d_nw3", B, 32)
    lo, hi = pr["lo"], pr["hi"]
    rng = np.random.default_rng(23)
    al = np.linspace(0, 1, H)[None, :, None]
    tr = pr["start"][:, None] * (1 - al) + pr["goal"][:, None] * al + rng.normal(0, 0.1, (B, H, D)) * np.sin(np.pi * al)
    x0 = torch.tensor(synth.normalize(np.clip(tr, lo, hi).astype(np.float32), lo, hi))
    walls = torch.cat([torch.tensor(pr["walls"]), torch.tensor(pr2["walls"])], 1)
    noise = torch.randn(B, H, D, generator=torch.Generator().manual_seed(10))
    pc = object.__new__(PC)
    pc.diffusion_model_list = [d1, d2]; pc.diffusion_model = d1; pc.n_models = 2
    pc.cg_w = torch.tensor([2.0, 1.5]).reshape(2, 1, 1, 1)
    pc.num_walls_c = torch.tensor([6, 3]); pc.wall_dim = 2; pc.comp_type = "direct"
    pc.w_split_sizes = tuple((pc.num_walls_c * pc.wall_dim).to(torch.int32))
    pc.use_normed_wallLoc = False; pc.use_mala_sampler = False; pc.uncond_base_idx = 0; pc.ddim_eta = 0.0
    pc.is_dyn_env = False; pc.wall_is_dyn = None
    for dd in (d1, d2):
        dd.horizon = H
    pc.set_ddim_steps(n)
    q_sample = d1.q_sample
    out = []
    for b in range(B):
        d1.q_sample = lambda x_start, t, noise_=None, b=b: q_sample(x_start, t, noise=noise[b, None])
        cond = {0: x0[b, None, 0], H - 1: x0[b, None, -1]}
        with refshim.quiet():
            out.append(pc.ddim_replan(x0[b, None].clone(), cond, walls[b, None], dn).detach()[0])
    d1.q_sample = q_sample
    np.savez_compressed(os.path.join(HERE, "resample_compose_m2d.npz"), x0=x0.numpy(), walls=walls.numpy(),
                        noise=noise.numpy(), x=torch.stack(out).numpy(), ddim_steps=n, dn_steps=dn,
                        w=np.array([2.0, 1.5], np.float32), seeds=np.array([1, 2]))
    print("resample_compose: max |x - x0| %.3f" % float((torch.stack(out) - x0).abs().max()))


def gen_replan():
    """the reference's own KukaDiffusionReplanner.replan_collision_section_less_check (kuka_replan.py:33-143) with its
    RM2DCollisionChecker and Maze2D robot, on noisy straight lines; the diffusion call is replaced by preset
    re-sampled trajectories (NEW[trial]) so that section finding, acceptance and check counting are what is frozen"""
    P, RW, RWG, Env = refshim.load_maze2d_geometry()
    Replanner, Checker, Limits = refshim.load_replanner()
    B, H, n_trial, hext, eps = 64, 48, 3, 0.5, 0.08
    pr = synth.make_problems("m2d", B, 41)
    rng = np.random.default_rng(8)
    al4 = np.linspace(0, 1, H)[None, None, :, None].astype(np.float32)
    s = pr["start"][:, None, None, :]; g = pr["goal"][:, None, None, :]
    trajs = (s * (1 - al4) + g * al4 + rng.normal(0, 0.05, size=(B, 1, H, 2)) * np.sin(np.pi * al4)).astype(np.float32)
    trajs = np.clip(trajs[:, 0], 0.02, 4.98).astype(np.float32)
    al = np.linspace(0, 1, H)[None, :, None]
    NEW = []
    for t in range(n_trial):
        amp = rng.normal(0, 0.5, (B, 1, 2)); ph = rng.integers(1, 4, (B, 1, 1))
        NEW.append(np.clip(trajs + amp * np.sin(np.pi * ph * al) ** 2, 0.02, 4.98).astype(np.float32))
    NEW = np.stack(NEW)

    class Norm:                                   # DatasetNormalizer interface over the reference's LimitsNormalizer
        observation_dim = 2
        lim = Limits(np.array([[0, 0], [5, 5]], np.float32))
        def normalize(self, x, key): return self.lim.normalize(x)
        def unnormalize(self, x, key): return self.lim.unnormalize(x)
    norm = Norm()

    class Model:
        def ddim_replan(self, normed, cond, walls_loc, dn_steps):
            b = int(walls_loc[0, 0].item())
            out = torch.as_tensor(norm.normalize(NEW[self.trial, b], "observations"), dtype=torch.float32)
            self.trial += 1
            return out

    new = trajs.copy(); suc = np.ones(B, np.uint8); cnt = np.full(B, H, np.int32); had = np.zeros(B, np.uint8)
    for b in range(B):
        walls = [RW(np.array(c), np.array([hext, hext])) for c in pr["boxes"][b]]
        rob = P(np.array([5., 5.]), RWG(walls), 0.01, collision_eps=eps)
        env = object.__new__(Env); env.robot = rob
        with contextlib.redirect_stdout(io.StringIO()):
            col = [not (rob.set_config(q) or rob.no_collision()) for q in trajs[b]]
        if not any(col):
            continue                               # the reference asserts it is never given a collision-free path
        had[b] = 1
        m = Model(); m.trial = 0
        rp = Replanner(m, norm, dict(use_ddim=True, dn_steps=3, n_replan_trial=n_trial), torch.device("cpu"))
        ck = Checker(norm); ck.use_collision_eps = True; ck.collision_eps = eps
        rp.set_collision_checker(ck)
        with refshim.quiet():
            r_new, r_suc, r_cnt = rp.replan_collision_section_less_check(env, trajs[b].copy(), torch.full((1, 12), float(b)))
        new[b] = r_new; suc[b] = bool(r_suc); cnt[b] = r_cnt
    np.savez_compressed(os.path.join(HERE, "replan_maze2d.npz"), trajs=trajs, resampled=NEW, boxes=pr["boxes"], hext=hext,
                        eps=eps, new=new, success=suc, num_colck=cnt, had_collision=had, numpy_version=np.__version__,
                        legacy_div=0)
    print("replan: %d with collisions, %d repaired, mean checks %.1f" % (had.sum(), (suc & had).sum(), cnt[had > 0].mean()))


def gen_pick():
    """the reference's own ComposedEnvPlanner.option_1_pick_1_traj (comp_plan_env.py:126-265) -- candidate scan over
    several planning horizons with its RM2DCollisionChecker -- on noisy straight-line candidates, one maze at a time"""
    P, RW, RWG, Env = refshim.load_maze2d_geometry()
    Planner = refshim.load_composed_planner()
    E, n_p, n_s, horizons, hext = 6, 8, 4, (40, 48, 56), 0.5
    pr = synth.make_problems("m2d", E, 77)
    rng = np.random.default_rng(12)
    out = dict(boxes=pr["boxes"], hext=hext, horizons=np.array(horizons), n_prob=n_p, n_samp=n_s, eps=0.08,
               legacy_div=0, numpy_version=np.__version__)
    chosen_h = np.zeros((E, n_p), np.int64); suc = np.zeros((E, n_p), np.uint8); cnt = np.zeros((E, n_p), np.int64)
    paths = np.zeros((E, n_p, max(horizons), 2), np.float32)
    cands = [np.zeros((E, n_p * n_s, h, 2), np.float32) for h in horizons]

    class Obj:
        pass
    PU = refshim._import_ref("diffuser.guides.kuka_plan_utils")[0]
    per_env = []
    for e in range(E):
        walls = [RW(np.array(c), np.array([hext, hext])) for c in pr["boxes"][e]]
        rob = P(np.array([5., 5.]), RWG(walls), 0.01, collision_eps=0.08)
        env = object.__new__(Env); env.robot = rob
        st = rng.uniform(0.2, 4.8, (n_p, 1, 1, 2)); gl = rng.uniform(0.2, 4.8, (n_p, 1, 1, 2))
        near = np.clip(st + rng.uniform(-1.0, 1.0, st.shape), 0.2, 4.8)      # short hops: more collision-free candidates
        gl = np.where(np.arange(n_p)[:, None, None, None] % 2 == 0, near, gl)
        samples = []
        for i_h, h in enumerate(horizons):
            al = np.linspace(0, 1, h)[None, None, :, None]
            sig = rng.choice([0.01, 0.05, 0.3], size=(n_p, n_s, 1, 1))
            tj = st * (1 - al) + gl * al + rng.normal(0, 1, (n_p, n_s, h, 2)) * sig * np.sin(np.pi * al)
            tj = np.clip(tj, 0.02, 4.98).astype(np.float32).reshape(n_p * n_s, h, 2)
            cands[i_h][e] = tj
            o = Obj(); o.observations = tj
            samples.append(o)
        pl = object.__new__(Planner)
        pl.horizon_pl_list = list(horizons); pl.samples_perprob = n_s; pl.is_dyn_env = False; pl.is_kuka = False
        pl.horizon_wtrajs = max(horizons)
        pl.replanner = Obj(); pl.replanner.set_collision_checker = lambda c: None
        with refshim.quiet():
            d = pl.option_1_pick_1_traj(samples, env, n_p)
            # per-maze statistics of the chosen paths (plan_kuka_permaze_stat -> check_collision, kuka_plan_utils.py)
            other = dict(replan_suc_list=[], no_replan_suc_list=list(d["suc_list"]), batch_runtime=0.125 * (e + 1),
                         num_colck_list=list(d["num_colck_list"]))
            per_env.append(PU.plan_kuka_permaze_stat(list(d["modelout_path_list"]), other, env))
        for p_ in range(n_p):
            tj = d["modelout_path_list"][p_]
            chosen_h[e, p_] = horizons.index(len(tj)); paths[e, p_, :len(tj)] = tj
            suc[e, p_] = bool(d["suc_list"][p_]); cnt[e, p_] = d["num_colck_list"][p_]
    out.update(chosen_h=chosen_h, success=suc, num_colck=cnt, paths=paths)
    with refshim.quiet():
        avg = PU.compute_kuka_avg_result(per_env)           # kuka_plan_utils.py:175-213
    out["collision_cnts"] = np.array([d["collision_cnts"] for d in per_env])
    out["batch_runtime_total"] = sum(d["batch_runtime"] for d in per_env)
    out.update({"stat_" + k: np.float64(v) for k, v in avg.items()})
    print("stats:", {k: float(v) for k, v in avg.items()})
    out.update({"cand_h%d" % h: c for h, c in zip(horizons, cands)})
    np.savez_compressed(os.path.join(HERE, "pick_multi_horizon.npz"), **out)
    print("pick: success %.2f, horizon hist %s, mean checks %.1f" % (suc.mean(), np.bincount(chosen_h.ravel(), minlength=3), cnt.mean()))


def gen_logdir():
    """model_config.pkl / diffusion_config.pkl written by the REFERENCE's own Config class (diffuser/utils/config.py:22-39)
    with the reference model classes as `_class`, exactly as scripts/train.py:64-102 does: the files a trained logdir
    holds.  The drop-in must unpickle them to its own classes (tests/test_overlay_seam.py, test_gpu_parity.py)."""
    import importlib.util
    U, G = refshim.load_models()
    spec = importlib.util.spec_from_file_location("diffuser.utils.config", os.path.join(refshim.REF_ROOT, "diffuser", "utils", "config.py"))
    cfgmod = importlib.util.module_from_spec(spec)
    out = os.path.join(HERE, "logdir_m2d")
    os.makedirs(out, exist_ok=True)
    a = synth.arch("m2d")
    with refshim.quiet():
        import sys as _sys
        _sys.modules["diffuser.utils.config"] = cfgmod      # the pickle records the class by this dotted path
        spec.loader.exec_module(cfgmod)
        cfgmod.Config(U, savepath=(out, "model_config.pkl"), verbose=False, **a["unet"])
        cfgmod.Config(G, savepath=(out, "diffusion_config.pkl"), verbose=False, **a["diffusion"])
        del _sys.modules["diffuser.utils.config"]
    for f in ("model_config.pkl", "diffusion_config.pkl"):
        raw = open(os.path.join(out, f), "rb").read()
        assert b"diffuser.utils.config" in raw and b"diffuser.models." in raw
        print("logdir", f, len(raw), "bytes")


def train_inputs(fam, B, seed):
    """deterministic training batch: normalised trajectories [B,H,D] (noisy straight segments), their start / goal
    conditions and the wall vectors in the dataset's [B,1,Wd] layout (diffusion_pb.py:531-534)"""
    a = synth.arch(fam)
    H, D = a["unet"]["horizon"], a["obs_dim"]
    pr = synth.make_problems(fam, B, seed)
    lo, hi = pr["lo"], pr["hi"]
    s = torch.tensor(synth.normalize(pr["start"], lo, hi))[:, None, :]
    g = torch.tensor(synth.normalize(pr["goal"], lo, hi))[:, None, :]
    al = torch.linspace(0, 1, H)[None, :, None]
    gen = torch.Generator().manual_seed(seed)
    x = (s * (1 - al) + g * al + 0.05 * torch.randn(B, H, D, generator=gen)).clamp(-1, 1)
    return a, x.contiguous(), torch.tensor(pr["walls"])[:, None, :].contiguous()


def grad_digest(name, g):
    """compact fingerprint of one gradient tensor (the full gradients of a 15-43 M parameter model do not belong in
    the repository): L2 norm, sum, two fixed pseudo-random projections and the first 32 elements"""
    g = g.detach().double().reshape(-1)
    gen = torch.Generator().manual_seed(synth._name_seed(name, 77) % (2 ** 31))
    r1 = torch.randn(g.numel(), generator=gen, dtype=torch.float64)
    r2 = torch.randn(g.numel(), generator=gen, dtype=torch.float64)
    head = torch.zeros(32, dtype=torch.float64)
    head[:min(32, g.numel())] = g[:32]
    return np.concatenate([[g.norm().item(), g.sum().item(), (g * r1).sum().item(), (g * r2).sum().item()], head.numpy()])


def gen_train():
    """loss value and parameter gradients of the REFERENCE's training step: GaussianDiffusionPB.loss (diffusion_pb.py:
    518-545) in train mode + loss.backward() (training_pbdiff.py:123-129), with the random draws (t, noise, concept
    dropout) of that call recorded so the restatement oracle/train_ref.py and the CUDA step can repeat them"""
    for fam, B in (("m2d", 4), ("dualkuka", 2)):
        a, x, walls = train_inputs(fam, B, 21)
        _, m, d, sd = build_ref(fam)
        m.train(); d.train()
        H, D = x.shape[1], x.shape[2]
        cond = {0: x[:, 0].clone(), H - 1: x[:, -1].clone()}
        torch.manual_seed(5); np.random.seed(2)
        with refshim.quiet():
            loss, info = d.loss(x.clone(), cond, walls.clone())
            loss.backward()
        # the same draws, in the reference's order: t (diffusion_pb.py:536), noise (:481), dropout mask (temporal_cond.py:268-270)
        torch.manual_seed(5); np.random.seed(2)
        t = torch.randint(0, d.n_timesteps, (B,)).long()
        noise = torch.randn_like(x)
        keep = (np.random.rand(B) >= a["unet"]["network_config"]["concept_drop_prob"]).astype(np.float32)
        names = [k for k, _ in m.named_parameters()]
        dig = np.stack([grad_digest(k, p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in m.named_parameters()])
        np.savez_compressed(os.path.join(HERE, "train_%s.npz" % fam), x=x.numpy(), walls=walls.numpy(), t=t.numpy(),
                            noise=noise.numpy(), keep=keep, loss=float(loss), names=np.array(names), digest=dig,
                            loss_weights=d.loss_fn.weights.numpy(), weight_seed=WEIGHT_SEED)
        print("train", fam, "loss %.6f" % float(loss), "keep", keep, "t", t.tolist(), "|g| %.4e" % float(np.sqrt((dig[:, 0] ** 2).sum())))


if __name__ == "__main__":
    which = sys.argv[1:] or ["unet", "sample", "compose", "colchk", "concave", "replan", "pick", "resample", "resample_compose", "logdir", "train"]
    for w in which:
        globals()["gen_" + w]()
