"""GPU parity tests proper: the CUDA path (through the C ABI / the drop-in classes) against the
golden vectors frozen from the reference and against the CPU oracle on the same seeded inputs."""
import os

import numpy as np
import pytest
import torch

import pmp_b200
from oracle import synth, unet_ref, sampler_ref, colchk_ref
from util import (golden, assert_eps_close, ddpm_coefs, ddim_coefs, pad_noise, TRAJ_ATOL)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_models = {}


def dropin(fam, seed=1):
    """drop-in TemporalUnet_WCond + GaussianDiffusionPB with the synthetic weights, on the GPU"""
    key = (fam, seed)
    if key not in _models:
        from diffuser.models import TemporalUnet_WCond, GaussianDiffusionPB
        a = synth.arch(fam)
        m = TemporalUnet_WCond(**a["unet"])
        m.load_state_dict(unet_ref.synth_weights(a["unet"], seed))
        _models[key] = GaussianDiffusionPB(m, **a["diffusion"]).to(DEV).eval()
    return _models[key]


# ------------------------------------------------------------------ potential gradient
@pytest.mark.parametrize("fam", synth.FAMILIES)
def test_eps_matches_reference_golden(fam):
    g = golden("unet_eps_%s.npz" % fam)
    d = dropin(fam)
    x = torch.tensor(g["x"], device=DEV); t = torch.tensor(g["t"], device=DEV); w = torch.tensor(g["walls"], device=DEV)
    eps = d.model(x, None, t, w, use_dropout=False, force_dropout=True, half_fd=True)
    assert_eps_close(eps.cpu().numpy(), g["eps"], "eps[%s]" % fam)
    # fully conditional / fully unconditional variants of the same call
    n = x.shape[0] // 2
    ec = d.model(x[:n], None, t[:n], w[:n], use_dropout=False)
    eu = d.model(x[n:], None, t[n:], w[n:], use_dropout=False, force_dropout=True)
    assert_eps_close(ec.cpu().numpy(), g["eps"][:n], "cond")
    assert_eps_close(eu.cpu().numpy(), g["eps"][n:], "uncond")


@pytest.mark.parametrize("fam,H", [("m2d", 36), ("m2d", 64), ("kuka", 44), ("dualkuka", 40)])
def test_eps_other_horizons_vs_oracle(fam, H):
    a = synth.arch(fam)
    cfg = unet_ref.UnetCfg(a["unet"]); sd = unet_ref.synth_weights(a["unet"], 1)
    d = dropin(fam)
    gen = torch.Generator().manual_seed(H)
    B = 3
    x = torch.randn(2 * B, H, a["obs_dim"], generator=gen)
    t = torch.tensor([0, 99, 50] * 2)
    walls = torch.tensor(synth.make_problems(fam, B, H)["walls"]).repeat(2, 1)
    ref, u_ref, e_ref = unet_ref.eps_autograd(sd, cfg, x, t, walls, n_cond=B, return_u=True)
    eng = d.model.engine(100)
    wemb = eng.wall_embed(walls[:B].to(DEV))
    eps, en, u = eng.eps(x.to(DEV), t.to(DEV), wemb, n_cond=B, want_energy=True, want_u=True)
    assert_eps_close(eps.cpu().numpy(), ref.numpy(), "eps")
    assert_eps_close(u.cpu().numpy(), u_ref.numpy(), "u")
    assert np.allclose(en.cpu().numpy(), e_ref.numpy(), rtol=1e-4)


def test_eps_chunking_and_determinism():
    d = dropin("m2d")
    eng = d.model.engine(100)
    gen = torch.Generator().manual_seed(4)
    B = 37
    x = torch.randn(2 * B, 48, 2, generator=gen).to(DEV)
    t = torch.randint(0, 100, (2 * B,), generator=gen).to(DEV)
    wemb = eng.wall_embed(torch.tensor(synth.make_problems("m2d", B, 8)["walls"], device=DEV))
    e1 = eng.eps(x, t, wemb, n_cond=B)
    e2 = eng.eps(x, t, wemb, n_cond=B)
    assert torch.equal(e1, e2), "eps evaluation is not run-to-run deterministic"
    eng.set_row_chunk(20)   # ragged chunks that split the cond / uncond boundary
    e3 = eng.eps(x, t, wemb, n_cond=B)
    eng.set_row_chunk(2048)
    assert torch.equal(e1, e3), "row chunking changed the result"


@pytest.mark.parametrize("fam", ["m2d", "dualkuka"])
def test_gemm_paths_agree(fam):
    """path 0 = exact fp32 CUDA-core convs, path 1 = tcgen05 BF16x3, path 2 = tcgen05 with operand reuse,
    path 3 = tcgen05 CTA pairs (cta_group::2), path 4 = path 3 with 16 epilogue warps"""
    a = synth.arch(fam)
    cfg = unet_ref.UnetCfg(a["unet"]); sd = unet_ref.synth_weights(a["unet"], 1)
    eng = dropin(fam).model.engine(100)
    gen = torch.Generator().manual_seed(9)
    B, H = 21, a["horizon"]
    x = torch.randn(2 * B, H, a["obs_dim"], generator=gen)
    t = torch.randint(0, 100, (2 * B,), generator=gen)
    walls = torch.tensor(synth.make_problems(fam, B, 12)["walls"]).repeat(2, 1)
    ref = unet_ref.eps_autograd(sd, cfg, x, t, walls, n_cond=B).numpy()
    wemb = eng.wall_embed(walls[:B].to(DEV))
    out = {}
    try:
        for path in (0, 1, 2, 3, 4):
            eng.set_gemm_path(path)
            out[path] = eng.eps(x.to(DEV), t.to(DEV), wemb, n_cond=B).cpu().numpy()
    finally:
        eng.set_gemm_path(3)
    scale = np.abs(ref).max()
    assert np.abs(out[0] - ref).max() <= 2e-5 * scale, "fp32 path"
    for path in (1, 2, 3, 4):
        assert np.abs(out[path] - ref).max() <= 3e-4 * scale, "tensor-core path %d: %.3e" % (path, np.abs(out[path] - ref).max() / scale)
        assert_eps_close(out[path], ref, "path %d" % path)


def test_wall_embed_vs_oracle_any_wall_count():
    for fam, nws in (("m2d", (1, 6, 9)), ("kuka", (4, 5)), ("m2d_concave", (7,))):
        a = synth.arch(fam)
        cfg = unet_ref.UnetCfg(a["unet"]); sd = unet_ref.synth_weights(a["unet"], 1)
        eng = dropin(fam).model.engine(100)
        for nw in nws:
            walls = torch.rand(5, nw * a["token_dim"], generator=torch.Generator().manual_seed(nw)) * 5
            got = eng.wall_embed(walls.to(DEV)).cpu()
            ref = unet_ref.wall_embed(sd, cfg, walls)
            assert (got - ref).abs().max() < 2e-5 * ref.abs().max() + 1e-6


# ------------------------------------------------------------------ sampler
def test_step_kernel_bit_exact():
    tb = sampler_ref.schedule(100)
    gen = torch.Generator().manual_seed(0)
    B, H, D = 33, 52, 7
    for K in (1, 2, 4):
        x = torch.randn(B, H, D, generator=gen)
        ec = [0.5 * torch.randn(B, H, D, generator=gen) for _ in range(K)]
        eu = [0.5 * torch.randn(B, H, D, generator=gen) for _ in range(K)]
        z = torch.randn(B, H, D, generator=gen)
        st = torch.rand(B, D, generator=gen) * 2 - 1; go = torch.rand(B, D, generator=gen) * 2 - 1
        w = [2.0, 1.5, 0.7, 3.0][:K]
        eps = sampler_ref.compose(ec, eu, w, K - 1)
        dev = lambda l: [e.to(DEV) for e in l]
        for t in (99, 1, 0):
            ref = sampler_ref.apply_cond(sampler_ref.ddpm_update(tb, x, t, eps, z), st, go)
            out = pmp_b200.step(0, x.to(DEV), dev(ec), dev(eu), w, ddpm_coefs(tb, [t])[0], noise=z.to(DEV),
                                start=st.to(DEV), goal=go.to(DEV), base=K - 1)
            assert torch.equal(out.cpu(), ref), ("ddpm", K, t)
        for eta in (0.0, 1.0):
            for t in (84, 0):
                ref = sampler_ref.apply_cond(sampler_ref.ddim_update(tb, x, t, eps, 8, 100, eta, z), st, go)
                out = pmp_b200.step(1, x.to(DEV), dev(ec), dev(eu), w, ddim_coefs(tb, [t], 8, 100, eta)[0],
                                    noise=z.to(DEV), start=st.to(DEV), goal=go.to(DEV), base=K - 1)
                assert torch.equal(out.cpu(), ref), ("ddim", K, eta, t)


@pytest.mark.parametrize("fam,mode", [("m2d", "ddpm"), ("m2d", "ddim"), ("kuka", "ddim"), ("dualkuka", "ddim")])
def test_sampler_matches_reference_golden(fam, mode):
    """final trajectories of the reference's p_sample_loop / ddim_p_sample_loop, same injected noise"""
    g = golden("sample_%s.npz" % fam)
    d = dropin(fam)
    a = synth.arch(fam)
    H, n = a["horizon"], int(g[mode + "_steps"])
    tb = sampler_ref.schedule(100)
    if mode == "ddpm":
        ts = list(range(99, -1, -1)); coefs = d.step_coefs_ddpm(ts); kind = 0
    else:
        d.ddim_num_inference_steps = n
        ts = d.ddim_set_timesteps(n); coefs = d.step_coefs_ddim(ts, 0.0, n); kind = 1
    eng = d.model.engine(100)
    wemb = eng.wall_embed(torch.tensor(g["walls"], device=DEV))
    x, tr = pmp_b200.sample([eng], [wemb], [2.0], torch.tensor(g["start"]), torch.tensor(g["goal"]), H, kind, ts, coefs,
                            noise=pad_noise(g[mode + "_noise"], n), trace=True)
    ref = g[mode + "_x"]
    err = np.abs(x.cpu().numpy() - ref).max()
    assert err <= TRAJ_ATOL, "final trajectory differs from the reference by %.3e" % err
    assert np.abs(tr.cpu().numpy() - g[mode + "_trace"]).max() <= TRAJ_ATOL
    assert np.array_equal(x[:, 0].cpu().numpy(), g["start"]) and np.array_equal(x[:, -1].cpu().numpy(), g["goal"])


@pytest.mark.parametrize("mode", ["ddim24", "ddpm"])
def test_composed_sampler_matches_reference_golden(mode):
    g = golden("compose_m2d.npz")
    d1, d2 = dropin("m2d", int(g["seeds"][0])), dropin("m2d_nw3", int(g["seeds"][1]))
    e1, e2 = d1.model.engine(100), d2.model.engine(100)
    steps = [int(v) for v in g[mode + "_steps"]]
    if mode == "ddim24":
        coefs = d1.step_coefs_ddim(steps, 1.0, 24); kind = 1
    else:
        coefs = d1.step_coefs_ddpm(steps); kind = 0
    w1 = e1.wall_embed(torch.tensor(g["walls1"], device=DEV)); w2 = e2.wall_embed(torch.tensor(g["walls2"], device=DEV))
    x = pmp_b200.sample([e1, e2], [w1, w2], [float(v) for v in g["w"]], torch.tensor(g["start"]), torch.tensor(g["goal"]),
                        40, kind, steps, coefs, noise=torch.tensor(g[mode + "_noise"]))
    err = np.abs(x.cpu().numpy() - g[mode + "_x"]).max()
    assert err <= TRAJ_ATOL, err
    # same model composed with itself (shared pmp_model, two wall subsets) must not alias its tables
    xs = pmp_b200.sample([e1, e1], [w1, w1.flip(0)], [1.0, 1.0], torch.tensor(g["start"]), torch.tensor(g["goal"]),
                         40, kind, steps, coefs, noise=torch.tensor(g[mode + "_noise"]))
    assert torch.isfinite(xs).all()


def test_dropin_forward_api_and_philox():
    d = dropin("m2d")
    B, H = 64, 48
    pr = synth.make_problems("m2d", B, 5)
    lo, hi = pr["lo"], pr["hi"]
    cond = {0: torch.tensor(synth.normalize(pr["start"], lo, hi), device=DEV),
            H - 1: torch.tensor(synth.normalize(pr["goal"], lo, hi), device=DEV)}
    walls = torch.tensor(pr["walls"], device=DEV)
    d.horizon = H; d.ddim_num_inference_steps = 8
    torch.manual_seed(0)
    x1 = d(cond, walls, use_ddim=True)
    torch.manual_seed(0)
    x2, tr = d(cond, walls, use_ddim=True, return_diffusion=True)
    assert x1.shape == (B, H, 2) and tr.shape == (B, 9, H, 2) and torch.equal(x1, x2)
    assert float(x1.abs().max()) <= 1.0 + 1e-6
    # in-kernel Philox: deterministic in the seed, and invariant to how the batch is sharded
    d.rng_mode = "philox"; d.philox_seed = 123
    try:
        ts = list(range(19, -1, -1)); coefs = d.step_coefs_ddpm(ts)
        eng = d.model.engine(100); wemb = eng.wall_embed(walls)
        full = pmp_b200.sample([eng], [wemb], [2.0], cond[0], cond[H - 1], H, 0, ts, coefs, seed=123)
        half = pmp_b200.sample([eng], [wemb[32:]], [2.0], cond[0][32:], cond[H - 1][32:], H, 0, ts, coefs, seed=123,
                               row_offset=32)
        assert torch.equal(full[32:], half)
        other = pmp_b200.sample([eng], [wemb], [2.0], cond[0], cond[H - 1], H, 0, ts, coefs, seed=124)
        assert not torch.equal(full, other)
    finally:
        d.rng_mode = "torch"

def test_sampler_graph_replay_equals_direct_launches():
    """pmp_sample replays one captured CUDA graph per step (device-resident step state, CFG halves sharing the state);
    with the graph switched off every kernel is launched from the host on the caller's buffers.  Same bits, for plain
    CFG DDPM and a K=3 same-network DDIM composition, several row chunks of unequal size, repeated calls (cached graph)"""
    d = dropin("m2d")
    B, H = 20, 48
    pr = synth.make_problems("m2d", B, 8)
    lo, hi = pr["lo"], pr["hi"]
    start = torch.tensor(synth.normalize(pr["start"], lo, hi), device=DEV)
    goal = torch.tensor(synth.normalize(pr["goal"], lo, hi), device=DEV)
    eng = d.model.engine(100)
    w1 = eng.wall_embed(torch.tensor(pr["walls"], device=DEV))
    w2 = eng.wall_embed(torch.tensor(pr["walls"][::-1].copy(), device=DEV))
    w3 = eng.wall_embed(torch.tensor(np.roll(pr["walls"], 3, 0), device=DEV))
    cases = [([eng], [w1], [2.0], 0, list(range(11, -1, -1))),
             ([eng, eng, eng], [w1, w2, w3], [2.0, 1.5, 1.0], 1, None)]
    try:
        for chunk in (8192, 16):              # one chunk / chunks of 8 + 8 + 4 plans
            eng.set_row_chunk(chunk)
            for engs, ws, gw, kind, ts in cases:
                if kind == 1:
                    ts = d.ddim_set_timesteps(8); coefs = d.step_coefs_ddim(ts, 0.0, 8)
                else:
                    coefs = d.step_coefs_ddpm(ts)
                outs = []
                for graph in (1, 1, 0):
                    pmp_b200.set_option("cuda_graph", graph)
                    n0 = pmp_b200.launch_count()
                    outs.append(pmp_b200.sample(engs, ws, gw, start, goal, H, kind, ts, coefs, noise=None, seed=77, base=0))
                    assert pmp_b200.launch_count() - n0 >= len(ts) * 60
                assert torch.isfinite(outs[0]).all()
                assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), (chunk, kind)
    finally:
        pmp_b200.set_option("cuda_graph", 1)
        eng.set_row_chunk(8192)


def test_compose_same_model_shares_unconditional_half():
    """K potentials that are one network (config/comp "gsd" compositions): the unconditional half is evaluated once;
    the sampled trajectories must equal, bit for bit, those of the K-fold evaluation"""
    import subprocess, sys, json as _json
    code = r"""
import sys, json, hashlib, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
from oracle import synth, unet_ref
import pmp_b200
from diffuser.models import TemporalUnet_WCond, GaussianDiffusionPB
DEV = "cuda:0"
a = synth.arch("m2d")
m = TemporalUnet_WCond(**a["unet"]); m.load_state_dict(unet_ref.synth_weights(a["unet"], 1))
d = GaussianDiffusionPB(m, **a["diffusion"]).to(DEV).eval(); eng = d.model.engine(100)
B, H = 24, 48
##OPT##
pr = synth.make_problems("m2d", B, 3)
lo, hi = pr["lo"], pr["hi"]
start = torch.tensor(synth.normalize(pr["start"], lo, hi), device=DEV); goal = torch.tensor(synth.normalize(pr["goal"], lo, hi), device=DEV)
w1 = eng.wall_embed(torch.tensor(pr["walls"], device=DEV)); w2 = eng.wall_embed(torch.tensor(pr["walls"][::-1].copy(), device=DEV))
w3 = eng.wall_embed(torch.tensor(np.roll(pr["walls"], 5, 0), device=DEV))
ts = d.ddim_set_timesteps(8); coefs = d.step_coefs_ddim(ts, 0.0, 8)
x = pmp_b200.sample([eng, eng, eng], [w1, w2, w3], [2.0, 1.5, 1.0], start, goal, H, 1, ts, coefs, noise=None, seed=5, base=1)
print(json.dumps({"sha": hashlib.sha256(x.cpu().numpy().tobytes()).hexdigest(), "launches": int(pmp_b200.launch_count())}))
""" % (ROOT, os.path.join(ROOT, "potential-motion-plan-release_b200"))
    outs = []
    for dedup in (1, 0):
        r = subprocess.run([sys.executable, "-c", code.replace("##OPT##", "pmp_b200.set_option('uncond_dedup', %d)" % dedup)],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(_json.loads(r.stdout.strip().splitlines()[-1]))
    assert outs[0]["sha"] == outs[1]["sha"], "sharing the unconditional half changed the result"


def test_philox_normal_statistics():
    """in-kernel Philox4x32-10 + Box-Muller (4 normals per call): N(0,1) moments, no correlation between the
    components of a call, reproducible per seed, different across seeds / draws"""
    x = torch.empty(8192, 48, 2, device=DEV)
    pmp_b200.init_noise(x, None, 1234, None, None)
    v = x.flatten().double()
    assert abs(v.mean().item()) < 5e-3 and abs(v.std().item() - 1) < 5e-3
    assert abs(((v - v.mean()) ** 4).mean().item() / v.var().item() ** 2 - 3.0) < 0.05
    q = v.view(-1, 4)
    c = torch.corrcoef(q.T)
    assert (c - torch.eye(4, device=DEV, dtype=torch.double)).abs().max().item() < 1.2e-2   # 5 sigma for 196608 groups
    assert abs((v[:-1] * v[1:]).mean().item()) < 5e-3
    y = torch.empty_like(x); pmp_b200.init_noise(y, None, 1234, None, None)
    z = torch.empty_like(x); pmp_b200.init_noise(z, None, 1235, None, None)
    assert torch.equal(x, y) and not torch.equal(x, z)
    assert (v.abs() > 4.0).float().mean().item() < 2e-4 and v.abs().max().item() < 7.0


def test_policy_and_compose_interfaces():
    from diffuser.guides import Policy, PolicyCompose
    from diffuser.datasets import DatasetNormalizer, RAND_MAZE_MIN, RAND_MAZESIZE_55
    norm = DatasetNormalizer({"observations": (RAND_MAZE_MIN, RAND_MAZESIZE_55)})
    d = dropin("m2d")
    d.ddim_num_inference_steps = 8
    B, H = 20, 48
    pr = synth.make_problems("m2d", B, 6)
    pol = Policy(d, norm, use_ddim=True)
    cond = {0: pr["start"], H - 1: pr["goal"]}
    _, traj = pol(cond, batch_size=-1, wall_locations=torch.tensor(pr["walls"]))
    obs = traj.observations
    assert obs.shape == (B, H, 2) and obs.dtype == np.float32
    assert np.allclose(obs[:, 0], pr["start"], atol=1e-5) and np.allclose(obs[:, -1], pr["goal"], atol=1e-5)
    assert obs.min() >= 0 and obs.max() <= 5
    d2 = dropin("m2d_nw3", 2)
    d2.horizon = d.horizon
    po = dict(num_walls_c=[6, 3], wall_dim=2, comp_type="direct", ddim_steps=8, wall_is_dyn=None, uncond_base_idx=0,
              ddim_eta=1.0)
    comp = PolicyCompose([d, d2], [norm, norm], False, True, po)
    walls9 = torch.cat([torch.tensor(pr["walls"]), torch.tensor(synth.make_problems("m2d_nw3", B, 7)["walls"])], 1)
    condt = {0: torch.tensor(pr["start"]), 40 - 1: torch.tensor(pr["goal"])}
    _, traj = comp(condt, walls9)
    assert traj.observations.shape == (B, 40, 2) and np.isfinite(traj.observations).all()


def test_plan_rm2d_from_problem_file():
    """scripts/plan_rm2d.py --problems <npz>: the ingest path end to end (problem dict -> plan_envs -> statistics)"""
    import subprocess, sys, json as _json
    script = os.path.join(ROOT, "potential-motion-plan-release_b200", "scripts", "plan_rm2d.py")
    prob = os.path.join(ROOT, "tests", "golden", "m2d_synth-problems.npz")
    r = subprocess.run([sys.executable, script, "--family", "m2d", "--plan_n_maze", "6", "--n_prob_env", "20",
                        "--problems", prob], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    res = _json.loads(r.stdout.strip().splitlines()[-1])
    assert res["num_samples"] == 120 and 0.0 <= res["traj_srate"] <= 1.0 and res["total_num_colck"] > 0
    # the same problems from the HDF5 file (the reference's own format, kuka_plan_utils.py:13-33): identical statistics
    r2 = subprocess.run([sys.executable, script, "--family", "m2d", "--plan_n_maze", "6", "--n_prob_env", "20",
                         "--problems", prob[:-4] + ".hdf5"], capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0, r2.stderr[-2000:]
    res2 = _json.loads(r2.stdout.strip().splitlines()[-1])
    assert res2["num_samples"] == 120 and res2["total_num_colck"] > 0


def test_plan_rm2d_concave_problem_file(tmp_path):
    """scripts/plan_rm2d.py --family m2d_concave --problems <npz>: the problem file holds the 21 small squares per
    maze; the (cx, cy, mode) model input is recomputed from them, with and without the generator's .npy / pickle"""
    import subprocess, sys, json as _json, pickle as _pickle
    from pmp_b200 import ingest
    z = golden("concave_layouts.npz")
    layouts, tokens = z["layouts"][:4], z["tokens"][:4]
    n_env, n_p = 4, 5
    rng = np.random.default_rng(2)
    probs = np.stack([np.stack([np.stack(synth.maze2d_start_goal(rng, layouts[e], 0.25)) for _ in range(n_p)])
                      for e in range(n_env)])
    prob = str(tmp_path / "c43-problems.npz")
    ingest.save_problems_npz({"problems": probs, "infos/wall_locations": np.repeat(layouts[:, None], n_p, 1)}, prob)
    npy = str(tmp_path / "c43_eval.npy"); np.save(npy, layouts)
    pkl = str(tmp_path / "c43_43_eval.pkl")
    with open(pkl, "wb") as f:
        _pickle.dump({"centers_43": tokens[..., :2].copy(), "mode_list": tokens[..., 2].astype(np.int64)}, f)
    script = os.path.join(ROOT, "potential-motion-plan-release_b200", "scripts", "plan_rm2d.py")
    out = []
    for extra in ([], ["--layouts", npy, "--concave43", pkl]):
        r = subprocess.run([sys.executable, script, "--family", "m2d_concave", "--plan_n_maze", "4", "--n_prob_env", "5",
                            "--problems", prob] + extra, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out.append(_json.loads(r.stdout.strip().splitlines()[-1]))
    assert out[0]["num_samples"] == 20 and out[0]["total_num_colck"] > 0
    for k in ("traj_srate", "traj_norepl_srate", "total_num_colck"):     # same inputs -> same plans (seeded sampler)
        assert out[0][k] == out[1][k]


def test_plan_envs_multi_horizon_composed():
    """composed planner over several horizons end to end (comp_plan_env.py:54-265): K=2 potentials, 3 horizons,
    GPU collision checks + interleaved selection; the chosen path must be the candidate the oracle's scan picks"""
    from diffuser.guides import PolicyCompose
    from diffuser.guides.plan_env_b200 import plan_envs_multi_horizon
    from diffuser.datasets import DatasetNormalizer, RAND_MAZE_MIN, RAND_MAZESIZE_55
    norm = DatasetNormalizer({"observations": (RAND_MAZE_MIN, RAND_MAZESIZE_55)})
    d = dropin("m2d"); d2 = dropin("m2d_nw3", 2)
    po = dict(num_walls_c=[6, 3], wall_dim=2, comp_type="direct", ddim_steps=8, wall_is_dyn=None, uncond_base_idx=0,
              ddim_eta=0.0)
    comp = PolicyCompose([d, d2], [norm, norm], False, True, po)
    n_env, n_prob, n_s, horizons = 3, 4, 3, (40, 48, 56)
    pr6 = synth.make_problems("m2d", n_env, 21); pr3 = synth.make_problems("m2d_nw3", n_env, 22)
    rng = np.random.default_rng(4)
    starts = rng.uniform(0.2, 4.8, (n_env, n_prob, 2)).astype(np.float32)
    goals = rng.uniform(0.2, 4.8, (n_env, n_prob, 2)).astype(np.float32)
    walls = np.repeat(np.concatenate([pr6["walls"], pr3["walls"]], 1)[:, None], n_prob, 1)
    boxes = np.concatenate([pr6["boxes"], pr3["boxes"]], 1)      # checker sees the union of both obstacle sets
    res, paths = plan_envs_multi_horizon(comp, starts, goals, walls, boxes, 0.5, horizons, samples_perprob=n_s)
    assert len(paths) == n_env * n_prob and sum(res["horizon_hist"]) == n_env * n_prob
    for p, path in enumerate(paths):
        assert path.shape[0] in horizons and path.shape[1] == 2 and np.isfinite(path).all()
        assert np.allclose(path[0], starts.reshape(-1, 2)[p], atol=1e-5) and np.allclose(path[-1], goals.reshape(-1, 2)[p], atol=1e-5)
    for k in ("traj_srate", "traj_norepl_srate", "total_num_colck", "avg_num_colck", "avg_sample_time"):
        assert k in res
    assert 0.0 <= res["traj_srate"] <= 1.0 and res["total_num_colck"] > 0


def test_plan_envs_multi_horizon_plain_policy():
    """the same multi-horizon planner with a single-model Policy (rm2d_plan_env.py:137-235 per horizon): the plain
    policy has call_multi_horizon too, so the dispatch must look at the class, not at that method; the selection is
    recomputed here from the per-horizon checks (first valid in the interleaved scan order, else the last candidate)"""
    from diffuser.guides import Policy
    from diffuser.guides.plan_env_b200 import plan_envs_multi_horizon
    from diffuser.datasets import DatasetNormalizer, RAND_MAZE_MIN, RAND_MAZESIZE_55
    norm = DatasetNormalizer({"observations": (RAND_MAZE_MIN, RAND_MAZESIZE_55)})
    d = dropin("m2d")
    d.ddim_num_inference_steps = 8
    pol = Policy(d, norm, use_ddim=True)
    n_env, n_prob, n_s, horizons = 3, 4, 2, (40, 48)
    pr = synth.make_problems("m2d", n_env, 23)
    rng = np.random.default_rng(5)
    starts = rng.uniform(0.2, 4.8, (n_env, n_prob, 2)).astype(np.float32)
    goals = rng.uniform(0.2, 4.8, (n_env, n_prob, 2)).astype(np.float32)
    walls = np.repeat(pr["walls"][:, None], n_prob, 1)
    res, paths = plan_envs_multi_horizon(pol, starts, goals, walls, pr["boxes"], float(pr["hext"]), horizons,
                                         samples_perprob=n_s)
    assert len(paths) == n_env * n_prob and sum(res["horizon_hist"]) == n_env * n_prob
    for p, path in enumerate(paths):
        assert path.shape[0] in horizons and path.shape[1] == 2 and np.isfinite(path).all()
        assert np.allclose(path[0], starts.reshape(-1, 2)[p], atol=1e-5) and np.allclose(path[-1], goals.reshape(-1, 2)[p], atol=1e-5)
    assert 0.0 <= res["traj_srate"] <= 1.0 and res["total_num_colck"] > 0
    assert d.horizon == horizons[-1]        # the policy re-targets the model's horizon per call, like policies.py:82


# ------------------------------------------------------------------ collision
@pytest.mark.parametrize("fam", ["m2d", "m2d_nw3", "m2d_concave"])
def test_maze2d_collision_matches_reference_golden(fam):
    g = golden("colchk_maze2d.npz")
    tr = torch.tensor(g[fam + "_traj"], device=DEV)
    env = g[fam + "_env"]; boxes = g[fam + "_boxes"]; hext = float(g[fam + "_hext"])
    v, c = pmp_b200.colchk_maze2d_swept(tr, env, boxes, hext, legacy_div=bool(int(g["legacy_div"])))
    assert np.array_equal(v.cpu().numpy(), g[fam + "_valid"]), "collision bits differ from the reference"
    assert np.array_equal(c.cpu().numpy(), g[fam + "_nchk"]), "check counts differ from the reference"
    n, pt = pmp_b200.colchk_maze2d_points(tr.clamp(0, 5), env, boxes, hext)
    assert np.array_equal(n.cpu().numpy(), g[fam + "_ncol"])
    assert np.array_equal(pt.cpu().numpy().sum(1), g[fam + "_ncol"])


def _noisy_lines(pr, n_per, H, rng, sig):
    E = len(pr["start"])
    al = np.linspace(0, 1, H)[None, None, :, None].astype(np.float32)
    s = pr["start"][:, None, None, :]; g = pr["goal"][:, None, None, :]
    return (s * (1 - al) + g * al + (rng.normal(0, sig, size=(E, n_per, H, s.shape[-1])) * np.sin(np.pi * al))).astype(np.float32)


@pytest.mark.parametrize("legacy", [True, False])
def test_maze2d_collision_full_size_vs_oracle(legacy):
    """BASELINE config 2 size: 10k problems (1000 envs x 10 candidates), Static2 + Concave"""
    rng = np.random.default_rng(3)
    for fam, hext in (("m2d_nw3", 0.7), ("m2d_concave", 0.25)):
        E, n_per, H = 1000, 10, 48
        pr = synth.make_problems(fam, E, 31)
        tr = np.clip(_noisy_lines(pr, n_per, H, rng, 0.15), -0.01, 5.01).reshape(E * n_per, H, 2)
        env = np.repeat(np.arange(E, dtype=np.int32), n_per)
        v_ref, c_ref = colchk_ref.maze2d_swept(tr, env, pr["boxes"], hext, legacy_div=legacy)
        v, c = pmp_b200.colchk_maze2d_swept(torch.tensor(tr, device=DEV), env, pr["boxes"], hext, legacy_div=legacy)
        assert np.array_equal(v.cpu().numpy(), v_ref) and np.array_equal(c.cpu().numpy(), c_ref)
        ch_ref, su_ref, tot_ref = colchk_ref.pick_first_valid(v_ref, c_ref, E, n_per)
        ch, su, tot = pmp_b200.pick_first_valid(v, c, E, n_per)
        assert np.array_equal(ch.cpu().numpy(), ch_ref) and np.array_equal(su.cpu().numpy().astype(bool), su_ref)
        assert np.array_equal(tot.cpu().numpy(), tot_ref)
        assert 0.02 < su_ref.mean() < 0.98

@pytest.mark.parametrize("legacy", [True, False])
def test_maze2d_swept_all_horizons_vs_oracle(legacy):
    """half-warp kernel (H <= 64: 2 / 3 / 4 waypoints per lane) and the generic one-warp kernel (H > 64): odd batch sizes,
    environments shared by runs of candidates and shuffled, long jumps with many interior points, out-of-limit points"""
    rng = np.random.default_rng(11)
    for H in (2, 5, 16, 17, 31, 32, 33, 36, 47, 48, 49, 63, 64, 65, 100, 128):
        for fam, hext in (("m2d", 0.5), ("m2d_concave", 0.25)):
            E, n_per = 37, 5
            pr = synth.make_problems(fam, E, 100 + H)
            sig = float(rng.choice([0.05, 0.3, 1.0]))
            tr = np.clip(_noisy_lines(pr, n_per, H, rng, sig), -0.05, 5.05).reshape(E * n_per, H, 2)
            env = np.repeat(np.arange(E, dtype=np.int32), n_per)
            if H % 2:                                   # shuffled order: the wall cache must follow env_id
                perm = rng.permutation(len(tr))
                tr, env = tr[perm], env[perm]
            tr, env = tr[:-1 - (H % 3)], env[:-1 - (H % 3)]
            v_ref, c_ref = colchk_ref.maze2d_swept(tr, env, pr["boxes"], hext, legacy_div=legacy)
            v, c = pmp_b200.colchk_maze2d_swept(torch.tensor(tr, device=DEV), env, pr["boxes"], hext, legacy_div=legacy)
            assert np.array_equal(v.cpu().numpy(), v_ref), (H, fam)
            assert np.array_equal(c.cpu().numpy(), c_ref), (H, fam)


def test_pick_multi_horizon_vs_oracle():
    """GPU selection over several horizons == sequential restatement of comp_plan_env.py:126-265"""
    from diffuser.guides.plan_env_b200 import pick_multi_horizon
    rng = np.random.default_rng(5)
    for n_prob, n_samp, nh, pv in ((1, 1, 1, 0.5), (257, 10, 4, 0.08), (1000, 5, 3, 0.0), (64, 20, 2, 0.6)):
        v = [(rng.random(n_prob * n_samp) < pv).astype(np.uint8) for _ in range(nh)]
        c = [rng.integers(0, 400, n_prob * n_samp).astype(np.int32) for _ in range(nh)]
        h_ref, s_ref, su_ref, tot_ref = colchk_ref.pick_multi_horizon(v, c, n_prob, n_samp)
        h, s, su, tot = pick_multi_horizon([torch.tensor(a, device=DEV) for a in v], [torch.tensor(a, device=DEV) for a in c],
                                           n_prob, n_samp)
        assert np.array_equal(h.cpu().numpy(), h_ref) and np.array_equal(s.cpu().numpy(), s_ref)
        assert np.array_equal(su.cpu().numpy().astype(bool), su_ref) and np.array_equal(tot.cpu().numpy(), tot_ref)


def test_pick_multi_horizon_matches_reference_golden():
    """GPU collision checks + selection against what the reference's own ComposedEnvPlanner.option_1_pick_1_traj chose
    for the same candidates (tests/golden/pick_multi_horizon.npz): all mazes in one batch"""
    from diffuser.guides.plan_env_b200 import pick_multi_horizon
    z = golden("pick_multi_horizon.npz")
    hz = [int(h) for h in z["horizons"]]; n_p = int(z["n_prob"]); n_s = int(z["n_samp"])
    E = len(z["boxes"])
    env = np.repeat(np.arange(E, dtype=np.int32), n_p * n_s)
    V, C = [], []
    for h in hz:
        tj = torch.tensor(z["cand_h%d" % h].reshape(E * n_p * n_s, h, 2), device=DEV)
        v, c = pmp_b200.colchk_maze2d_swept(tj, env, z["boxes"], float(z["hext"]), eps=float(z["eps"]),
                                            legacy_div=bool(z["legacy_div"]))
        V.append(v); C.append(c)
    hs, ss, su, tot = pick_multi_horizon(V, C, E * n_p, n_s)
    hs = hs.cpu().numpy(); ss = ss.cpu().numpy()
    assert np.array_equal(hs, z["chosen_h"].ravel()) and np.array_equal(su.cpu().numpy().astype(bool), z["success"].ravel().astype(bool))
    assert np.array_equal(tot.cpu().numpy().astype(np.int64), z["num_colck"].ravel())
    for q in range(E * n_p):
        path = z["cand_h%d" % hz[hs[q]]].reshape(E * n_p, n_s, hz[hs[q]], 2)[q, ss[q]]
        assert np.array_equal(path, z["paths"].reshape(E * n_p, -1, 2)[q, :len(path)])
    # final scoring of the chosen paths (check_collision, kuka_plan_utils.py:104-146) and the global statistics
    from diffuser.guides.plan_env_b200 import compute_avg_result
    env_p = np.repeat(np.arange(E, dtype=np.int32), n_p)
    ncol = np.zeros(E * n_p, np.int64)
    for i, h in enumerate(hz):
        sel = np.nonzero(hs == i)[0]
        pt = torch.tensor(z["paths"].reshape(E * n_p, -1, 2)[sel, :h], device=DEV)
        ncol[sel] = pmp_b200.colchk_maze2d_points(pt, env_p[sel], z["boxes"], float(z["hext"]))[0].cpu().numpy()
    assert np.array_equal(ncol == 0, z["collision_cnts"].ravel() == 0)      # the reference pads the shorter paths
    res = compute_avg_result(ncol, su.cpu().numpy(), tot.cpu().numpy(), E, n_p, float(z["batch_runtime_total"]))
    for k in res:
        assert float(res[k]) == float(z["stat_" + k]), k


def test_batched_replanner_vs_oracle():
    """KukaDiffusionReplanner.replan_batch (all problems together, GPU checks) == the sequential restatement of
    kuka_replan.py:33-143 on identical re-sampled trajectories: new paths, success flags and check counts"""
    from oracle import replan_ref
    from diffuser.guides.kuka_replan import KukaDiffusionReplanner
    from diffuser.guides import RM2DCollisionChecker
    from diffuser.datasets import DatasetNormalizer, RAND_MAZE_MIN, RAND_MAZESIZE_55
    norm = DatasetNormalizer({"observations": (RAND_MAZE_MIN, RAND_MAZESIZE_55)})
    B, H, n_trial = 96, 48, 3
    pr = synth.make_problems("m2d", B, 41)
    rng = np.random.default_rng(8)
    trajs = np.clip(_noisy_lines(pr, 1, H, rng, 0.05)[:, 0], 0.02, 4.98).astype(np.float32)
    al = np.linspace(0, 1, H)[None, :, None]
    NEW = []
    for t in range(n_trial):
        amp = rng.normal(0, 0.5, (B, 1, 2)); ph = rng.integers(1, 4, (B, 1, 1))
        NEW.append(np.clip(trajs + amp * np.sin(np.pi * ph * al) ** 2, 0.02, 4.98).astype(np.float32))
    roundtrip = lambda a: norm.unnormalize(torch.as_tensor(norm.normalize(a, "observations"), dtype=torch.float32).numpy(),
                                           "observations").astype(np.float32)

    class MockModel:                       # stands in for GaussianDiffusionPB.ddim_replan: row id travels in walls_loc
        def __init__(self):
            self.trial = 0
        def ddim_replan(self, normed, cond, walls_loc, dn_steps):
            ids = walls_loc[:, 0].long().cpu().numpy()
            out = torch.as_tensor(norm.normalize(NEW[self.trial][ids], "observations"), dtype=torch.float32, device=normed.device)
            self.trial += 1
            return out

    rp = KukaDiffusionReplanner(MockModel(), norm, dict(use_ddim=True, dn_steps=3, n_replan_trial=n_trial), DEV)
    chk = RM2DCollisionChecker(normalizer=None, device=DEV); chk.use_collision_eps = True; chk.collision_eps = 0.08
    rp.set_collision_checker(chk)
    walls = torch.arange(B, dtype=torch.float32)[:, None].repeat(1, 12)
    geom = dict(robot="maze2d", env_id=np.arange(B, dtype=np.int32), cen=pr["boxes"], hext=0.5)
    new, suc, cnt = rp.replan_batch(trajs, walls, geom)
    n_fix = n_fail = 0
    for b in range(B):
        ck = replan_ref.Maze2DChecker(pr["boxes"][b], 0.5)
        r_new, r_suc, r_cnt = replan_ref.replan_collision_section(trajs[b], ck, lambda t, tr: roundtrip(NEW[t][b]), n_trial)
        assert np.array_equal(new[b], r_new), "problem %d: replanned path differs" % b
        assert bool(suc[b]) == r_suc and int(cnt[b]) == r_cnt, "problem %d: %s/%d vs %s/%d" % (b, suc[b], cnt[b], r_suc, r_cnt)
        had = ck.point_collides(trajs[b]).any()
        n_fix += int(had and r_suc); n_fail += int(had and not r_suc)
    assert n_fix >= 3 and n_fail >= 3, (n_fix, n_fail)


def test_batched_replanner_matches_reference_golden():
    """replan_batch against what the reference's own replanner returned for the same inputs
    (tests/golden/replan_maze2d.npz, gen_golden.py replan): paths, success flags, check counts, bit for bit"""
    from diffuser.guides.kuka_replan import KukaDiffusionReplanner
    from diffuser.guides import RM2DCollisionChecker
    from diffuser.datasets import DatasetNormalizer, RAND_MAZE_MIN, RAND_MAZESIZE_55
    norm = DatasetNormalizer({"observations": (RAND_MAZE_MIN, RAND_MAZESIZE_55)})
    z = golden("replan_maze2d.npz")
    trajs, NEW = z["trajs"], z["resampled"]
    B = len(trajs)

    class MockModel:
        trial = 0
        def ddim_replan(self, normed, cond, walls_loc, dn_steps):
            ids = walls_loc[:, 0].long().cpu().numpy()
            out = torch.as_tensor(norm.normalize(NEW[self.trial][ids], "observations"), dtype=torch.float32, device=normed.device)
            self.trial += 1
            return out

    rp = KukaDiffusionReplanner(MockModel(), norm, dict(use_ddim=True, dn_steps=3, n_replan_trial=len(NEW)), DEV)
    chk = RM2DCollisionChecker(normalizer=None, device=DEV, legacy_div=bool(z["legacy_div"]))     # numpy >= 2 counts
    chk.use_collision_eps = True; chk.collision_eps = float(z["eps"])
    rp.set_collision_checker(chk)
    walls = torch.arange(B, dtype=torch.float32)[:, None].repeat(1, 12)
    geom = dict(robot="maze2d", env_id=np.arange(B, dtype=np.int32), cen=z["boxes"], hext=float(z["hext"]))
    new, suc, cnt = rp.replan_batch(trajs, walls, geom)
    assert np.array_equal(np.asarray(new), z["new"])
    assert np.array_equal(np.asarray(suc).astype(bool), z["success"].astype(bool))
    assert np.array_equal(np.asarray(cnt).astype(np.int64), z["num_colck"].astype(np.int64))


def test_replanner_single_trajectory_with_diffusion():
    """reference signature (one trajectory, env object) through the real batched ddim_replan"""
    from diffuser.guides.kuka_replan import KukaDiffusionReplanner
    from diffuser.guides import RM2DCollisionChecker
    from diffuser.datasets import DatasetNormalizer, RAND_MAZE_MIN, RAND_MAZESIZE_55
    norm = DatasetNormalizer({"observations": (RAND_MAZE_MIN, RAND_MAZESIZE_55)})
    d = dropin("m2d"); d.ddim_num_inference_steps = 8
    rp = KukaDiffusionReplanner(d, norm, dict(use_ddim=True, dn_steps=3, n_replan_trial=2), DEV)
    chk = RM2DCollisionChecker(normalizer=None, device=DEV); chk.use_collision_eps = True; chk.collision_eps = 0.08
    rp.set_collision_checker(chk)
    pr = synth.make_problems("m2d", 1, 5)
    class Env: pass
    env = Env(); env.wall_centers = pr["boxes"][0]; env.wall_hext = 0.5
    H = 48
    line = np.stack([np.linspace(0.3, 4.7, H), np.linspace(0.3, 4.7, H)], -1).astype(np.float32)
    new, ok, cnt = rp.replan_collision_section_less_check(env, line, torch.tensor(pr["walls"][:1]))
    assert new.shape == (H, 2) and np.isfinite(new).all() and isinstance(ok, bool) and cnt >= H


def test_maze2d_collision_edge_cases():
    cen = np.array([[[2.5, 2.5]]]); hext = 0.5
    H = 8
    line = np.stack([np.linspace(0.2, 4.8, H), np.full(H, 2.5)], -1).astype(np.float32)
    free = np.stack([np.linspace(0.2, 4.8, H), np.full(H, 0.5)], -1).astype(np.float32)
    oob = free.copy(); oob[3, 1] = -0.1
    touch = free.copy(); touch[:, 1] = np.float32(2.5 - 0.5 - 0.01)   # exactly on the inflated boundary: strict test
    tr = np.stack([line, free, oob, touch])
    env = np.zeros(4, np.int32)
    v_ref, c_ref = colchk_ref.maze2d_swept(tr, env, cen, hext)
    v, c = pmp_b200.colchk_maze2d_swept(torch.tensor(tr, device=DEV), env, cen, hext)
    assert np.array_equal(v.cpu().numpy(), v_ref) and np.array_equal(c.cpu().numpy(), c_ref)
    assert v_ref.tolist()[:3] == [0, 1, 0]
    n, pt = pmp_b200.colchk_maze2d_points(torch.tensor(tr, device=DEV), env, cen, hext)
    assert n.cpu().tolist()[2] == -1
    # empty batch, no walls, two-waypoint trajectory
    v, c = pmp_b200.colchk_maze2d_swept(torch.zeros(0, 8, 2, device=DEV), np.zeros(0, np.int32), cen, hext)
    assert v.numel() == 0
    v, c = pmp_b200.colchk_maze2d_swept(torch.tensor(free[None, :2], device=DEV), np.zeros(1, np.int32),
                                        np.zeros((1, 0, 2)), np.zeros((1, 0, 2)))
    assert v.cpu().tolist() == [1]
    from diffuser.guides import RM2DCollisionChecker

    class Env:
        wall_centers = cen[0]; wall_hext = hext
    chk = RM2DCollisionChecker(None)
    chk.use_collision_eps = True; chk.collision_eps = 0.08
    assert chk.check_single_traj(free, Env) == (True, int(c_ref[1]))
    assert chk.check_single_traj(line, Env) == (False, int(c_ref[0]))
    chk.use_collision_eps = False
    assert chk.check_single_traj(free, Env) == (True, H)
    ok, cnt = chk.check_single_traj(line, Env)
    assert not ok and cnt == int(np.argmax(pt.cpu().numpy()[0])) + 1


@pytest.mark.parametrize("fam,arms,hext", [("kuka", 1, 0.20), ("dualkuka", 2, 0.22)])
def test_kuka_collision_vs_oracle(fam, arms, hext):
    rng = np.random.default_rng(7)
    E, n_per = 300, 8
    H = synth.arch(fam)["horizon"]
    pr = synth.make_problems(fam, E, 41)
    # mixture: near-upright (mostly free) and random joint-space lines (mostly colliding)
    pr_small = dict(start=pr["start"] * 0.15, goal=pr["goal"] * 0.15)
    tr = np.concatenate([_noisy_lines(pr_small, n_per // 2, H, rng, 0.05), _noisy_lines(pr, n_per // 2, H, rng, 0.3)], 1)
    tr = tr.reshape(E * n_per, H, -1)
    env = np.repeat(np.arange(E, dtype=np.int32), n_per)
    sph = colchk_ref.kuka_sphere_table()
    ref = colchk_ref.kuka_points(tr, env, pr["boxes"], hext, n_arms=arms, sph=sph)
    got = pmp_b200.colchk_kuka(torch.tensor(tr, device=DEV), env, pr["boxes"], hext, arms, colchk_ref.kuka_bases(arms), sph)
    for k in ("valid", "nchk", "ncol", "ptcol"):
        assert np.array_equal(got[k].cpu().numpy(), ref[k]), k
    assert 0.05 < ref["valid"].mean() < 0.95
    from diffuser.guides import KukaCollisionChecker

    class Env:
        voxel_centers = pr["boxes"][0]; voxel_hext = hext; n_arms = arms
    ok, cnt = KukaCollisionChecker().check_single_traj(tr[0], Env)
    assert ok == bool(ref["valid"][0]) and cnt == int(ref["nchk"][0])
    # the conservative table (spheres that contain the link AABBs): same kernel, still bit-equal to the C restatement run on
    # that table, and never more permissive than the boxes it contains
    from diffuser.guides.kuka_colchk import sphere_table
    sphc = sphere_table(conservative=True)
    refc = colchk_ref.kuka_points(tr, env, pr["boxes"], hext, n_arms=arms, sph=sphc)
    gotc = pmp_b200.colchk_kuka(torch.tensor(tr, device=DEV), env, pr["boxes"], hext, arms, colchk_ref.kuka_bases(arms), sphc)
    for k in ("valid", "nchk", "ncol", "ptcol"):
        assert np.array_equal(gotc[k].cpu().numpy(), refc[k]), k
    assert refc["valid"].mean() <= ref["valid"].mean() + 0.05


def test_unnormalize_bit_exact():
    for fam in ("m2d", "kuka", "dualkuka"):
        lo, hi = synth.limits(fam)
        x = (torch.randn(50, 48, len(lo), generator=torch.Generator().manual_seed(1)) * 0.9)
        got = pmp_b200.unnormalize(x.to(DEV), lo, hi).cpu().numpy()
        assert np.array_equal(got, synth.unnormalize(x.numpy(), lo, hi))


def test_errors_are_reported():
    d = dropin("m2d")
    eng = d.model.engine(100)
    with pytest.raises(pmp_b200.PmpError):
        eng.eps(torch.zeros(2, 50, 2, device=DEV), torch.zeros(2, dtype=torch.long, device=DEV),
                torch.zeros(2, 64, device=DEV))       # horizon not a multiple of 4
    with pytest.raises(pmp_b200.PmpError):
        eng.wall_embed(torch.zeros(2, 7, device=DEV))  # wall vector not a multiple of token_dim
    with pytest.raises(pmp_b200.PmpError):
        pmp_b200.colchk_maze2d_swept(torch.zeros(1, 8, 2, device=DEV), np.zeros(1, np.int32), np.zeros((1, 40, 2)), 0.1)
    assert pmp_b200.launch_count() > 0


def test_plan_entry_points():
    """plan_rm2d / plan_kuka entry points: batched planning + first-valid pick + scoring, reference result keys"""
    import importlib.util
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "potential-motion-plan-release_b200", "scripts")
    spec = importlib.util.spec_from_file_location("plan_rm2d_b200", os.path.join(root, "plan_rm2d.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    keys = {"num_samples", "traj_srate", "traj_norepl_srate", "avg_batch_time", "avg_sample_time", "avg_num_colck",
            "total_num_colck", "total_ncr", "total_svr", "traj_repl_srate"}
    res = mod.main(["--plan_n_maze", "2", "--n_prob_env", "3", "--samples_perprob", "4"])
    assert keys <= set(res) and res["num_samples"] == 6 and 0.0 <= res["traj_srate"] <= 1.0
    assert res["avg_num_colck"] > 0
    res = mod.main(["--family", "dualkuka", "--plan_n_maze", "2", "--n_prob_env", "2", "--samples_perprob", "3",
                    "--ddim_steps", "10"], robot="kuka")
    assert keys <= set(res) and res["num_samples"] == 4 and res["collision_model"].startswith("sphere-link")
    res = mod.main(["--family", "kuka", "--plan_n_maze", "2", "--n_prob_env", "2", "--samples_perprob", "3",
                    "--ddim_steps", "10", "--kuka_conservative", "1"], robot="kuka")
    assert res["num_samples"] == 4 and res["collision_model"].startswith("conservative")


def test_plan_rm2d_from_a_reference_written_logdir(tmp_path):
    """scripts/plan_rm2d.py --logdir: model_config.pkl / diffusion_config.pkl written by the REFERENCE's own Config class
    (tests/golden/logdir_m2d, gen_golden.py:gen_logdir) + state_<epoch>.pt = {'step','model','ema'} in the reference
    trainer's format (training_pbdiff.py:177-191).  The highest epoch's EMA weights must be the ones that plan:
    same statistics as a model built directly with those weights, different from the other checkpoint's."""
    import importlib.util
    import os
    import shutil
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for f in ("model_config.pkl", "diffusion_config.pkl"):
        shutil.copy(os.path.join(root, "tests", "golden", "logdir_m2d", f), tmp_path / f)
    d_model, d_ema, d_old = dropin("m2d", 1), dropin("m2d", 2), dropin("m2d", 3)
    cpu = lambda sd: {k: v.detach().cpu() for k, v in sd.items()}
    torch.save({"step": 9, "model": cpu(d_model.state_dict()), "ema": cpu(d_ema.state_dict())}, tmp_path / "state_10.pt")
    torch.save({"step": 1, "model": cpu(d_old.state_dict()), "ema": cpu(d_old.state_dict())}, tmp_path / "state_2.pt")
    spec = importlib.util.spec_from_file_location("plan_rm2d_b200", os.path.join(root, "potential-motion-plan-release_b200", "scripts", "plan_rm2d.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert os.path.basename(mod.latest_checkpoint(str(tmp_path))) == "state_10.pt"
    a, loaded = mod.load_diffusion("m2d", str(tmp_path), DEV)
    assert type(loaded).__module__ == "diffuser.models.diffusion_pb" and type(loaded.model).__module__ == "diffuser.models.temporal_cond"
    for (k, v), (k2, v2) in zip(loaded.state_dict().items(), d_ema.state_dict().items()):
        assert k == k2 and torch.equal(v.cpu(), v2.cpu()), k
    argv = ["--plan_n_maze", "2", "--n_prob_env", "3", "--samples_perprob", "4"]
    torch.manual_seed(123)
    res = mod.main(argv + ["--logdir", str(tmp_path)])
    assert res["num_samples"] == 6 and res["avg_num_colck"] > 0
    # the same plan through the public classes with the EMA weights directly
    pr_x = torch.randn(3, 48, 2, device=DEV)
    t = torch.full((3,), 40, device=DEV, dtype=torch.long)
    walls = torch.tensor(synth.make_problems("m2d", 3, 9)["walls"], device=DEV)
    e1 = loaded.model(pr_x, None, t, walls, use_dropout=False, force_dropout=False)
    e2 = d_ema.model(pr_x, None, t, walls, use_dropout=False, force_dropout=False)
    e3 = d_model.model(pr_x, None, t, walls, use_dropout=False, force_dropout=False)
    assert torch.equal(e1, e2) and not torch.equal(e1, e3)


def test_ddim_replan_matches_reference_golden():
    """GaussianDiffusionPB.ddim_replan (partial re-noise + the last dn_steps DDIM steps, diffusion_pb.py:311-346) on a
    batch of four against the reference's outputs for the same trajectories and q_sample noise
    (tests/golden/resample_m2d.npz)"""
    g = golden("resample_m2d.npz")
    d = dropin("m2d", int(g["weight_seed"]))
    H = g["x0"].shape[1]
    d.horizon = H; d.ddim_num_inference_steps = int(g["ddim_steps"])
    x0 = torch.tensor(g["x0"], device=DEV); noise = torch.tensor(g["noise"], device=DEV)
    walls = torch.tensor(g["walls"], device=DEV)
    cond = {0: x0[:, 0].clone(), H - 1: x0[:, -1].clone()}
    dn = torch.full((1,), int(g["dn_steps"]), dtype=torch.long, device=DEV)
    q_sample = d.q_sample
    d.q_sample = lambda x_start, t, noise_=None: q_sample(x_start, t, noise=noise)
    try:
        noisy = d.ddim_replan(x0, cond, walls, dn, ret_noisy_traj=True)
        x = d.ddim_replan(x0, cond, walls, dn)
    finally:
        del d.q_sample
    assert np.abs(noisy.cpu().numpy() - g["noisy"]).max() <= 1e-6
    err = np.abs(x.cpu().numpy() - g["x"]).max()
    assert err <= TRAJ_ATOL, "re-sampled trajectories differ from the reference by %.3e" % err
    assert torch.equal(x[:, 0], x0[:, 0]) and torch.equal(x[:, -1], x0[:, -1])


def test_composed_ddim_replan_matches_reference_golden():
    """PolicyCompose.ddim_replan (K=2, policies_compose.py:399-447) on a batch of three against the reference
    method's outputs for the same trajectories and q_sample noise (tests/golden/resample_compose_m2d.npz)"""
    from diffuser.guides import PolicyCompose
    from diffuser.datasets import DatasetNormalizer, RAND_MAZE_MIN, RAND_MAZESIZE_55
    g = golden("resample_compose_m2d.npz")
    norm = DatasetNormalizer({"observations": (RAND_MAZE_MIN, RAND_MAZESIZE_55)})
    d1, d2 = dropin("m2d", int(g["seeds"][0])), dropin("m2d_nw3", int(g["seeds"][1]))
    H = g["x0"].shape[1]
    w_old = (d1.condition_guidance_w, d2.condition_guidance_w)
    d1.horizon = d2.horizon = H
    d1.condition_guidance_w, d2.condition_guidance_w = float(g["w"][0]), float(g["w"][1])
    po = dict(num_walls_c=[6, 3], wall_dim=2, comp_type="direct", ddim_steps=int(g["ddim_steps"]), wall_is_dyn=None,
              uncond_base_idx=0, ddim_eta=0.0)
    comp = PolicyCompose([d1, d2], [norm, norm], False, True, po)
    comp.set_ddim_steps(int(g["ddim_steps"]))
    x0 = torch.tensor(g["x0"], device=DEV); noise = torch.tensor(g["noise"], device=DEV)
    cond = {0: x0[:, 0].clone(), H - 1: x0[:, -1].clone()}
    q_sample = d1.q_sample
    d1.q_sample = lambda x_start, t, noise_=None: q_sample(x_start, t, noise=noise)
    try:
        x = comp.ddim_replan(x0.clone(), cond, torch.tensor(g["walls"]), int(g["dn_steps"]))
    finally:
        del d1.q_sample
        d1.condition_guidance_w, d2.condition_guidance_w = w_old
    err = np.abs(x.cpu().numpy() - g["x"]).max()
    assert err <= TRAJ_ATOL, "composed re-sampling differs from the reference by %.3e" % err


def test_plan_envs_with_replanning():
    """plan_envs(..., replanner=...): problems whose candidates all collide go through the batched replanner
    (kuka_plan_env.py:207-260); statistics keep the reference's meaning (replan_suc_list starts from the scan's
    successes) and the run is the plain one when nothing fails"""
    from diffuser.guides import Policy, RM2DCollisionChecker
    from diffuser.guides.kuka_replan import KukaDiffusionReplanner
    from diffuser.guides.plan_env_b200 import plan_envs
    from diffuser.datasets import DatasetNormalizer, RAND_MAZE_MIN, RAND_MAZESIZE_55
    norm = DatasetNormalizer({"observations": (RAND_MAZE_MIN, RAND_MAZESIZE_55)})
    d = dropin("m2d"); d.horizon = 48; d.ddim_num_inference_steps = 8
    policy = Policy(d, norm, use_ddim=True)
    rp = KukaDiffusionReplanner(d, norm, dict(use_ddim=True, dn_steps=3, n_replan_trial=2), DEV)
    chk = RM2DCollisionChecker(normalizer=None, device=DEV); chk.use_collision_eps = True; chk.collision_eps = 0.08
    rp.set_collision_checker(chk)
    n_env, n_prob, n_s = 3, 4, 2
    pr = synth.make_problems("m2d", n_env, 61)
    rng = np.random.default_rng(9)
    starts = rng.uniform(0.2, 4.8, (n_env, n_prob, 2)).astype(np.float32)
    goals = rng.uniform(0.2, 4.8, (n_env, n_prob, 2)).astype(np.float32)
    walls = np.repeat(pr["walls"][:, None], n_prob, 1)
    torch.manual_seed(3)
    res0, paths0 = plan_envs(policy, starts, goals, walls, pr["boxes"], 0.5, 48, samples_perprob=n_s)
    torch.manual_seed(3)
    res1, paths1 = plan_envs(policy, starts, goals, walls, pr["boxes"], 0.5, 48, samples_perprob=n_s, replanner=rp)
    assert res0["traj_repl_srate"] == 0.0 and res1["traj_norepl_srate"] == res0["traj_norepl_srate"]
    assert res1["traj_repl_srate"] >= res1["traj_norepl_srate"] and res1["traj_srate"] >= res0["traj_srate"]
    if res0["traj_norepl_srate"] < 1.0:          # random weights: most candidates collide, so the replanner ran
        assert res1["total_num_colck"] > res0["total_num_colck"]
    assert paths1.shape == paths0.shape and np.isfinite(paths1).all()
    same = [np.array_equal(a, b) for a, b in zip(paths0, paths1)]
    assert sum(same) >= round(res0["traj_norepl_srate"] * n_env * n_prob)      # scan successes are left alone
