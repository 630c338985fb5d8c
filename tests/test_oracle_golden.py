"""CPU: the oracle restatement against the golden vectors frozen from the live reference
(tests/golden/gen_golden.py), and -- when /root/reference is present -- against the live
reference itself."""
import contextlib
import io

import numpy as np
import pytest
import torch

from oracle import refshim, synth, unet_ref, sampler_ref, colchk_ref
from util import golden, pad_noise

torch.set_num_threads(8)


@pytest.mark.parametrize("fam", synth.FAMILIES)
def test_unet_eps_matches_golden(fam):
    g = golden("unet_eps_%s.npz" % fam)
    a = synth.arch(fam)
    sd = unet_ref.synth_weights(a["unet"], int(g["weight_seed"]))
    assert synth.state_dict_digest(sd) == str(g["digest"])
    cfg = unet_ref.UnetCfg(a["unet"])
    x, t, walls = torch.tensor(g["x"]), torch.tensor(g["t"]), torch.tensor(g["walls"])
    eps = unet_ref.eps_autograd(sd, cfg, x, t, walls, n_cond=int(g["n_cond"]))
    # same torch build, same ops -> bit-identical
    assert torch.equal(eps, torch.tensor(g["eps"]))
    eps_a, _ = unet_ref.eps_analytic(sd, cfg, x, t, walls, n_cond=int(g["n_cond"]))
    assert (eps_a - eps).abs().max() < 1e-5 * eps.abs().max() + 1e-6


def test_analytic_vjp_fp64():
    a = synth.arch("dualkuka")
    cfg = unet_ref.UnetCfg(a["unet"])
    sd = {k: v.double() for k, v in unet_ref.synth_weights(a["unet"], 3).items()}
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 40, 14, generator=g, dtype=torch.float64)
    walls = torch.rand(2, 15, generator=g, dtype=torch.float64)
    t = torch.tensor([3, 77])
    e1 = unet_ref.eps_autograd(sd, cfg, x, t, walls, n_cond=1)
    e2, _ = unet_ref.eps_analytic(sd, cfg, x, t, walls, n_cond=1)
    assert (e1 - e2).abs().max() < 1e-12


@pytest.mark.parametrize("fam,mode", [("m2d", "ddim"), ("m2d", "ddpm"), ("kuka", "ddim"), ("dualkuka", "ddim")])
def test_sampler_matches_golden(fam, mode):
    g = golden("sample_%s.npz" % fam)
    a = synth.arch(fam)
    cfg = unet_ref.UnetCfg(a["unet"])
    sd = unet_ref.synth_weights(a["unet"], int(g["weight_seed"]))
    n = int(g[mode + "_steps"])
    noise = pad_noise(g[mode + "_noise"], n)
    x = sampler_ref.sample([(sd, cfg)], torch.tensor(g["start"]), torch.tensor(g["goal"]), [torch.tensor(g["walls"])],
                           noise, T=100, ddim_steps=n if mode == "ddim" else 0)
    assert torch.equal(x, torch.tensor(g[mode + "_x"]))


def test_ddim_replan_matches_golden():
    """sampler_ref.ddim_replan against the reference's GaussianDiffusionPB.ddim_replan (gen_golden.py resample) and
    PolicyCompose.ddim_replan (gen_golden.py resample_compose); the reference ran one trajectory per call"""
    g = golden("resample_m2d.npz")
    a = synth.arch("m2d")
    sd = unet_ref.synth_weights(a["unet"], int(g["weight_seed"]))
    x, noisy = sampler_ref.ddim_replan([(sd, unet_ref.UnetCfg(a["unet"]))], torch.tensor(g["x0"]), [torch.tensor(g["walls"])],
                                       torch.tensor(g["noise"]), int(g["ddim_steps"]), int(g["dn_steps"]))
    assert torch.equal(noisy, torch.tensor(g["noisy"]))
    assert np.abs(x.numpy() - g["x"]).max() <= 1e-5, np.abs(x.numpy() - g["x"]).max()    # batch of B vs B calls of 1
    g = golden("resample_compose_m2d.npz")
    models = []
    for fam, seed in (("m2d", int(g["seeds"][0])), ("m2d_nw3", int(g["seeds"][1]))):
        a = synth.arch(fam)
        models.append((unet_ref.synth_weights(a["unet"], seed), unet_ref.UnetCfg(a["unet"])))
    wl = torch.tensor(g["walls"])
    x, _ = sampler_ref.ddim_replan(models, torch.tensor(g["x0"]), [wl[:, :12], wl[:, 12:]], torch.tensor(g["noise"]),
                                   int(g["ddim_steps"]), int(g["dn_steps"]), w=[float(v) for v in g["w"]], compose_mode=True)
    assert np.abs(x.numpy() - g["x"]).max() <= 1e-5, np.abs(x.numpy() - g["x"]).max()    # batch of B vs B calls of 1


@pytest.mark.parametrize("mode", ["ddim24", "ddpm"])
def test_compose_matches_golden(mode):
    g = golden("compose_m2d.npz")
    models = []
    for fam, seed in (("m2d", int(g["seeds"][0])), ("m2d_nw3", int(g["seeds"][1]))):
        a = synth.arch(fam)
        models.append((unet_ref.synth_weights(a["unet"], seed), unet_ref.UnetCfg(a["unet"])))
    start, goal = torch.tensor(g["start"]), torch.tensor(g["goal"])
    walls = [torch.tensor(g["walls1"]), torch.tensor(g["walls2"])]
    noise = torch.tensor(g[mode + "_noise"])
    steps = [int(v) for v in g[mode + "_steps"]]
    tb = sampler_ref.schedule(100)
    x = noise[0].clone()
    B = x.shape[0]
    for i, t in enumerate(steps):
        x = sampler_ref.apply_cond(x, start, goal)
        tt = torch.full((B,), t, dtype=torch.long)
        ecs, eus = [], []
        for (sd, cfg), wl in zip(models, walls):
            ec, eu = unet_ref.eps_pair(sd, cfg, x, tt, wl)
            ecs.append(ec); eus.append(eu)
        eps = sampler_ref.compose(ecs, eus, [float(v) for v in g["w"]], 0)
        if mode == "ddim24":
            x = sampler_ref.ddim_update(tb, x, t, eps, 24, 100, 1.0, noise[1 + i])
        else:
            x = sampler_ref.ddpm_update(tb, x, t, eps, noise[1 + i])
    x = sampler_ref.apply_cond(x, start, goal)
    assert torch.equal(x, torch.tensor(g[mode + "_x"]))


@pytest.mark.parametrize("fam", ["m2d", "m2d_nw3", "m2d_concave"])
def test_colchk_matches_golden(fam):
    g = golden("colchk_maze2d.npz")
    tr, env = g[fam + "_traj"], g[fam + "_env"]
    v, c = colchk_ref.maze2d_swept(tr, env, g[fam + "_boxes"], float(g[fam + "_hext"]), legacy_div=bool(int(g["legacy_div"])))
    assert np.array_equal(v, g[fam + "_valid"]) and np.array_equal(c, g[fam + "_nchk"])
    n, _ = colchk_ref.maze2d_points(np.clip(tr, 0, 5), env, g[fam + "_boxes"], float(g[fam + "_hext"]))
    assert np.array_equal(n, g[fam + "_ncol"])
    assert 0 < v.mean() < 1   # fixtures exercise both outcomes


def test_colchk_edge_cases():
    cen = np.array([[[2.5, 2.5]]]); hext = 0.5
    H = 8
    line = np.stack([np.linspace(0.2, 4.8, H), np.full(H, 2.5)], -1).astype(np.float32)   # straight through the box
    free = np.stack([np.linspace(0.2, 4.8, H), np.full(H, 0.5)], -1).astype(np.float32)
    oob = free.copy(); oob[3, 1] = -0.1
    tr = np.stack([line, free, oob])
    v, c = colchk_ref.maze2d_swept(tr, np.zeros(3, np.int32), cen, hext)
    assert v.tolist() == [0, 1, 0]
    # |segment| = 0.657 -> K = int(0.657/0.08) = 8 -> 2 endpoint + 7 interior checks per free segment
    K = int(np.float64(np.float32(np.sqrt(np.float32((line[1, 0] - line[0, 0]) ** 2)))) / 0.08)
    assert K == 8 and c[1] == (H - 1) * (2 + K - 1)
    # out-of-limits waypoint 3: segments 0 and 1 are checked in full, segment 2 fails the limits
    # test before any collision check (abstract_robot.py:135-136)
    assert c[2] == 2 * (2 + K - 1)
    assert c[0] < c[1]   # early exit at the first colliding point
    # no walls at all, two waypoints
    v, c = colchk_ref.maze2d_swept(free[None, :2], np.zeros(1, np.int32), np.zeros((1, 0, 2)), np.zeros((1, 0, 2)))
    assert v.tolist() == [1]
    # numpy-version hazard (Appendix C.3): the two K conventions can differ
    v1, c1 = colchk_ref.maze2d_swept(tr, np.zeros(3, np.int32), cen, hext, legacy_div=True)
    v2, c2 = colchk_ref.maze2d_swept(tr, np.zeros(3, np.int32), cen, hext, legacy_div=False)
    assert np.array_equal(v1, v2)


def test_kuka_model_sanity():
    sph = colchk_ref.kuka_sphere_table()
    assert sph.shape[1] == 5 and (np.diff(sph[:, 0]) >= 0).all()
    up = np.zeros((1, 4, 7), np.float32)            # arm straight up: free
    box_far = np.array([[[2.0, 2.0, 0.5]]], np.float32)
    r = colchk_ref.kuka_points(up, np.zeros(1, np.int32), box_far, 0.2)
    assert r["valid"].tolist() == [1] and r["nchk"].tolist() == [4]
    box_hit = np.array([[[0.0, 0.0, 1.0]]], np.float32)   # box on the arm's axis
    r = colchk_ref.kuka_points(up, np.zeros(1, np.int32), box_hit, 0.2)
    assert r["valid"].tolist() == [0] and r["nchk"].tolist() == [1] and r["ncol"].tolist() == [4]
    bent = up.copy(); bent[0, :, 1] = 2.0           # shoulder folded down: hits the table plane
    r = colchk_ref.kuka_points(bent, np.zeros(1, np.int32), box_far, 0.2)
    assert r["valid"].tolist() == [0]
    # FK check: tip of link 7 at q=0 is at z = sum of the chain offsets
    fx = colchk_ref.kuka_fixed_transforms()
    assert abs(fx[:, 11].sum() - 0) < 10  # smoke; exact geometry covered by the table/box cases above
    # dual arm: bases 1 m apart, both upright -> free; arms reaching at each other collide
    up2 = np.zeros((1, 2, 14), np.float32)
    r = colchk_ref.kuka_points(up2, np.zeros(1, np.int32), box_far, 0.22, n_arms=2)
    assert r["valid"].tolist() == [1]
    reach = up2.copy(); reach[0, :, 1] = -1.57; reach[0, :, 8] = 1.57
    r2 = colchk_ref.kuka_points(reach, np.zeros(1, np.int32), box_far, 0.22, n_arms=2)
    reach2 = up2.copy(); reach2[0, :, 1] = 1.57; reach2[0, :, 8] = -1.57
    r3 = colchk_ref.kuka_points(reach2, np.zeros(1, np.int32), box_far, 0.22, n_arms=2)
    assert r2["valid"].tolist() == [0] or r3["valid"].tolist() == [0]


def test_pick_first_valid():
    valid = np.array([0, 1, 1, 0, 0, 0, 1, 0, 0], np.uint8)
    nchk = np.arange(1, 10, dtype=np.int32)
    ch, su, tot = colchk_ref.pick_first_valid(valid, nchk, 3, 3)
    assert ch.tolist() == [1, 2, 0] and su.tolist() == [True, False, True] and tot.tolist() == [3, 15, 7]


def test_pick_multi_horizon_order():
    """comp_plan_env.py:205-238: sample-major / horizon-minor scan, stop at the first valid, else the last visited"""
    # 2 problems x 2 samples x 3 horizons
    v = [np.array([0, 0, 0, 0]), np.array([0, 1, 0, 0]), np.array([1, 1, 0, 0])]
    c = [np.array([1, 2, 3, 4]), np.array([10, 20, 30, 40]), np.array([100, 200, 300, 400])]
    h, s, suc, tot = colchk_ref.pick_multi_horizon(v, c, 2, 2)
    # problem 0: (s0,h0) no, (s0,h1) no, (s0,h2) yes -> horizon 2, sample 0, checks 1+10+100
    # problem 1: nothing valid -> last visited (s1,h2), all six counts
    assert h.tolist() == [2, 2] and s.tolist() == [0, 1] and suc.tolist() == [True, False]
    assert tot.tolist() == [111, 3 + 30 + 300 + 4 + 40 + 400]
    # a single horizon degenerates to pick_first_valid
    rng = np.random.default_rng(0)
    vv = rng.random(60) < 0.3; cc = rng.integers(1, 50, 60)
    h1, s1, su1, t1 = colchk_ref.pick_multi_horizon([vv], [cc], 12, 5)
    ch, su, tt = colchk_ref.pick_first_valid(vv, cc, 12, 5)
    assert (h1 == 0).all() and np.array_equal(s1, ch) and np.array_equal(su1, su) and np.array_equal(t1, tt)


def test_pick_multi_horizon_matches_golden():
    """colchk_ref.maze2d_swept + pick_multi_horizon against tests/golden/pick_multi_horizon.npz: what the reference's
    own ComposedEnvPlanner.option_1_pick_1_traj chose (gen_golden.py pick) -- horizon, path, success, check count"""
    z = golden("pick_multi_horizon.npz")
    hz = [int(h) for h in z["horizons"]]; n_p = int(z["n_prob"]); n_s = int(z["n_samp"])
    assert 0.15 < z["success"].mean() < 0.85 and len(np.unique(z["chosen_h"])) == len(hz)
    for e in range(len(z["boxes"])):
        V, C = [], []
        for h in hz:
            tj = z["cand_h%d" % h][e]
            v, c = colchk_ref.maze2d_swept(tj, np.zeros(len(tj), np.int32), z["boxes"][e][None], float(z["hext"]),
                                           eps=float(z["eps"]), legacy_div=bool(z["legacy_div"]))
            V.append(v); C.append(c)
        hs, ss, su, tot = colchk_ref.pick_multi_horizon(V, C, n_p, n_s)
        assert np.array_equal(hs, z["chosen_h"][e]) and np.array_equal(su, z["success"][e].astype(bool))
        assert np.array_equal(tot, z["num_colck"][e])
        for p in range(n_p):
            path = z["cand_h%d" % hz[hs[p]]][e][p * n_s + ss[p]]
            assert np.array_equal(path, z["paths"][e, p, :len(path)])


def test_replan_oracle_sections_and_acceptance():
    """oracle/replan_ref.py on a hand case: a straight line through one wall, re-sampled as a detour"""
    from oracle import replan_ref
    assert replan_ref.expanded_subtraj(48, 10, 12).tolist() == list(range(7, 15))      # 3 -> 8 waypoints, grown left first
    assert replan_ref.expanded_subtraj(48, 0, 1).tolist() == list(range(0, 8))         # clipped at the start
    assert [r.tolist() for r in replan_ref.split_runs([1, 2, 5, 6, 7])] == [[1, 2], [5, 6, 7]]
    H = 48
    cen = np.array([[2.5, 2.5]]); hext = 0.5
    line = np.stack([np.linspace(0.5, 4.5, H), np.full(H, 2.5)], -1).astype(np.float32)
    chk = replan_ref.Maze2DChecker(cen, hext)
    col = np.nonzero(chk.point_collides(line))[0]
    assert len(col) > 0 and col[0] > 0 and col[-1] < H - 1
    bump = line.copy()
    bump[:, 1] += 1.2 * np.sin(np.pi * np.linspace(0, 1, H)).astype(np.float32) ** 2          # passes above the wall
    calls = []
    def resample(trial, traj):
        calls.append(trial)
        return line if trial == 0 else bump          # first trial changes nothing -> rejected, second is the detour
    new, ok, cnt = replan_ref.replan_collision_section(line, chk, resample)
    assert calls == [0, 1] and ok and cnt > H
    assert not chk.point_collides(new).any()
    sec = replan_ref.expanded_subtraj(H, int(col[0]), int(col[-1]))
    untouched = np.setdiff1d(np.arange(H), sec)
    assert np.array_equal(new[untouched], line[untouched]) and np.array_equal(new[sec], bump[sec])
    # nothing ever fits: all trials used, failure reported, trajectory unchanged
    new2, ok2, cnt2 = replan_ref.replan_collision_section(line, chk, lambda trial, traj: line, n_replan_trial=3)
    assert not ok2 and np.array_equal(new2, line) and cnt2 > H


def test_replan_oracle_matches_golden():
    """oracle/replan_ref.py against tests/golden/replan_maze2d.npz: outputs of the reference's own
    KukaDiffusionReplanner + RM2DCollisionChecker + Maze2D robot on preset re-sampled trajectories
    (gen_golden.py replan); new paths, success flags and collision-check counts, bit for bit"""
    from oracle import replan_ref
    z = golden("replan_maze2d.npz")
    tr, NEW, boxes = z["trajs"], z["resampled"], z["boxes"]
    lo = np.zeros(2, np.float32); hi = np.full(2, 5, np.float32)

    def roundtrip(a):       # LimitsNormalizer.normalize -> float32 tensor -> unnormalize (normalization.py:142-162)
        n = ((a - lo) / (hi - lo) * 2 - 1).astype(np.float32)
        return ((n + 1) / 2. * (hi - lo) + lo).astype(np.float32)
    assert z["had_collision"].sum() > 40 and 0 < (z["success"] & z["had_collision"]).sum() < z["had_collision"].sum()
    for b in range(len(tr)):
        ck = replan_ref.Maze2DChecker(boxes[b], float(z["hext"]), eps=float(z["eps"]), legacy_div=bool(z["legacy_div"]))
        new, ok, cnt = replan_ref.replan_collision_section(tr[b], ck, lambda t, x: roundtrip(NEW[t][b]), len(NEW))
        assert np.array_equal(new, z["new"][b]), "problem %d: path differs" % b
        assert bool(ok) == bool(z["success"][b]) and cnt == int(z["num_colck"][b]), (b, ok, cnt)


# ------------------------------------------------------------------ live reference (build container only)
needs_ref = pytest.mark.skipif(not refshim.available(), reason="reference tree not present")


@needs_ref
def test_param_inventory_matches_reference():
    U, G = refshim.load_models()
    for fam in synth.FAMILIES:
        a = synth.arch(fam)
        with refshim.quiet():
            m = U(**a["unet"])
        ref = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert list(ref.items()) == list(unet_ref.param_shapes(a["unet"]).items())


@needs_ref
def test_oracle_matches_live_reference_eps():
    U, G = refshim.load_models()
    a = synth.arch("kuka")
    sd = unet_ref.synth_weights(a["unet"], 4)
    with refshim.quiet():
        m = U(**a["unet"]).eval()
        m.load_state_dict(sd)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 52, 7, generator=g); t = torch.tensor([9, 80, 9, 80])
    walls = torch.rand(4, 12, generator=g)
    with refshim.quiet():
        ref = m(x.clone(), None, t, walls, use_dropout=False, force_dropout=True, half_fd=True).detach()
    got = unet_ref.eps_autograd(sd, unet_ref.UnetCfg(a["unet"]), x, t, walls, n_cond=2)
    assert torch.equal(ref, got)


@needs_ref
def test_oracle_colchk_matches_live_reference():
    P, RW, RWG, Env = refshim.load_maze2d_geometry()
    rng = np.random.default_rng(1)
    pr = synth.make_problems("m2d", 4, 17)
    H = 48
    for e in range(4):
        rob = P(np.array([5., 5.]), RWG([RW(np.array(c), np.array([0.5, 0.5])) for c in pr["boxes"][e]]), 0.01,
                collision_eps=0.08)
        env = object.__new__(Env); env.robot = rob
        al = np.linspace(0, 1, H)[:, None].astype(np.float32)
        for sig in (0.02, 0.2):
            tr = (pr["start"][e] * (1 - al) + pr["goal"][e] * al + rng.normal(0, sig, (H, 2)) * np.sin(np.pi * al)).astype(np.float32)
            tr = np.clip(tr, 0, 5)
            cnt, ok = 0, True
            with contextlib.redirect_stdout(io.StringIO()):
                for j in range(H - 1):
                    rob.collision_check_count = 0
                    good = env.edge_fp(tr[j], tr[j + 1])
                    cnt += rob.collision_check_count
                    if not good:
                        ok = False
                        break
            v, c = colchk_ref.maze2d_swept(tr[None], np.zeros(1, np.int32), pr["boxes"][e], 0.5,
                                           legacy_div=np.__version__ < "2")
            assert bool(v[0]) == ok and int(c[0]) == cnt


@needs_ref
def test_pick_first_valid_matches_live_reference_planner():
    """colchk_ref.maze2d_swept + pick_first_valid against the reference's single-horizon candidate scan
    (RandMaze2DEnvPlanner.option_1_pick_1_traj, rm2d_plan_env.py:137-235) run here on the horizon-48 candidates of
    tests/golden/pick_multi_horizon.npz"""
    P, RW, RWG, Env = refshim.load_maze2d_geometry()
    Planner = refshim._import_ref("diffuser.guides.rm2d_plan_env")[0].RandMaze2DEnvPlanner
    z = golden("pick_multi_horizon.npz")
    n_p, n_s, hext = int(z["n_prob"]), int(z["n_samp"]), float(z["hext"])

    class Obj:
        pass
    for e in range(3):
        tj = z["cand_h48"][e]
        walls = [RW(np.array(c), np.array([hext, hext])) for c in z["boxes"][e]]
        env = object.__new__(Env); env.robot = P(np.array([5., 5.]), RWG(walls), 0.01, collision_eps=0.08)
        pl = object.__new__(Planner)
        pl.samples_perprob = n_s; pl.horizon = 48; pl.is_dyn_env = False
        pl.replanner = Obj(); pl.replanner.set_collision_checker = lambda c: None
        o = Obj(); o.observations = tj
        with refshim.quiet():
            d = pl.option_1_pick_1_traj(o, env, n_p)
        v, c = colchk_ref.maze2d_swept(tj, np.zeros(len(tj), np.int32), z["boxes"][e][None], hext, eps=float(z["eps"]),
                                       legacy_div=bool(z["legacy_div"]))
        ch, su, tot = colchk_ref.pick_first_valid(v, c, n_p, n_s)
        assert [bool(s) for s in d["suc_list"]] == su.tolist() and list(d["num_colck_list"]) == tot.tolist()
        for p in range(n_p):
            assert np.array_equal(d["modelout_path_list"][p], tj[p * n_s + ch[p]])


def test_cpu_arm_checker_sample_runs_on_both_kinds():
    """bench.py's cpu_baseline.checker leg (oracle/cpu_arm.py checker_rate): the C restatement everywhere, the
    reference's Python loop where the tree is present; both score the same straight-line candidates the same way"""
    from oracle import cpu_arm
    arm = cpu_arm.CpuArm(threads=1, prefer_reference=False)
    rate, what = arm.checker_rate(200)
    assert arm.kind == "port" and rate > 0 and "colchk.c" in what
    if refshim.available():
        arm.kind = "reference"
        arm.geom = refshim.load_maze2d_geometry()
        rate_r, what_r = arm.checker_rate(60)
        assert rate_r > 0 and "reference" in what_r
        # same candidates through both scorers: identical bits
        pr = synth.make_problems("m2d_nw3", 6, 4003)
        rng = np.random.default_rng(3)
        un = np.linspace(rng.uniform(0.2, 4.8, (60, 2)), rng.uniform(0.2, 4.8, (60, 2)), 48, axis=1).astype(np.float32)
        env = (np.arange(60) // 10).astype(np.int32)
        ok_ref = arm._score("m2d_nw3", dict(pr, boxes=pr["boxes"][env]), torch.tensor(synth.normalize(un, pr["lo"], pr["hi"])))
        un2 = synth.unnormalize(synth.normalize(un, pr["lo"], pr["hi"]), pr["lo"], pr["hi"])
        ok_c, _ = colchk_ref.maze2d_swept(un2, env, pr["boxes"], pr["hext"])
        assert np.array_equal(ok_ref, np.asarray(ok_c, bool))
