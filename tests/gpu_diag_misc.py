"""GPU diagnostic (not a pytest file): sampler-step / collision kernels vs the CPU oracle."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "potential-motion-plan-release_b200"))

from oracle import synth, unet_ref, sampler_ref, colchk_ref  # noqa: E402
import pmp_b200  # noqa: E402

dev = torch.device("cuda:0")


def ddpm_coefs(tb, t):
    sig = (0.5 * tb["posterior_log_variance_clipped"][t]).exp()
    sig = sig * (0.0 if t == 0 else 1.0)
    return [tb["sqrt_recip_alphas_cumprod"][t], tb["sqrt_recipm1_alphas_cumprod"][t],
            tb["posterior_mean_coef1"][t], tb["posterior_mean_coef2"][t], sig, 0, 0, 0]


def ddim_coefs(tb, t, n, T, eta):
    c = sampler_ref.ddim_coefs(tb, t, n, T, eta)
    return [c["recip"], c["recipm1"], c["sqrt_a"], c["sqrt_b"], c["sqrt_a_prev"], c["dirc"], c["std"],
            1.0 if eta > 0 else 0.0]


def test_step():
    tb = sampler_ref.schedule(100)
    g = torch.Generator().manual_seed(0)
    B, H, D = 37, 48, 2
    for K in (1, 2, 3):
        x = torch.randn(B, H, D, generator=g)
        ec = [0.5 * torch.randn(B, H, D, generator=g) for _ in range(K)]
        eu = [0.5 * torch.randn(B, H, D, generator=g) for _ in range(K)]
        z = torch.randn(B, H, D, generator=g)
        st = torch.rand(B, D, generator=g) * 2 - 1
        go = torch.rand(B, D, generator=g) * 2 - 1
        w = [2.0, 1.5, 0.7][:K]
        for t in (99, 50, 1, 0):
            eps = sampler_ref.compose(ec, eu, w, 0)
            ref = sampler_ref.apply_cond(sampler_ref.ddpm_update(tb, x, t, eps, z), st, go)
            out = pmp_b200.step(0, x.to(dev), [e.to(dev) for e in ec], [e.to(dev) for e in eu], w,
                                [float(v) for v in ddpm_coefs(tb, t)], noise=z.to(dev), start=st.to(dev), goal=go.to(dev))
            d = (out.cpu() - ref).abs().max().item()
            print("  ddpm K=%d t=%2d max|diff| %.3e bitexact=%s" % (K, t, d, torch.equal(out.cpu(), ref)))
        for eta in (0.0, 1.0):
            for t in (84, 36, 0):
                eps = sampler_ref.compose(ec, eu, w, 0)
                ref = sampler_ref.apply_cond(sampler_ref.ddim_update(tb, x, t, eps, 8, 100, eta, z), st, go)
                out = pmp_b200.step(1, x.to(dev), [e.to(dev) for e in ec], [e.to(dev) for e in eu], w,
                                    [float(v) for v in ddim_coefs(tb, t, 8, 100, eta)], noise=z.to(dev), start=st.to(dev),
                                    goal=go.to(dev))
                d = (out.cpu() - ref).abs().max().item()
                print("  ddim K=%d eta=%.0f t=%2d max|diff| %.3e bitexact=%s" % (K, eta, t, d, torch.equal(out.cpu(), ref)))
    # philox noise statistics
    x = torch.empty(4096, 48, 2, device=dev)
    pmp_b200.init_noise(x, None, 1234, None, None)
    print("  philox normal: mean %.4f std %.4f  kurt %.3f" % (x.mean().item(), x.std().item(),
                                                             ((x - x.mean()) ** 4).mean().item() / x.var().item() ** 2))
    lo, hi = synth.limits("kuka")
    xx = torch.randn(100, 52, 7) * 0.8
    ref = synth.unnormalize(xx.numpy(), lo, hi)
    got = pmp_b200.unnormalize(xx.to(dev), lo, hi).cpu().numpy()
    print("  unnormalize bitexact=%s" % np.array_equal(ref, got))


def make_trajs(pr, n_per, H, rng, sigma=0.25):
    E = len(pr["start"])
    al = np.linspace(0, 1, H)[None, None, :, None].astype(np.float32)
    s = pr["start"][:, None, None, :]; g = pr["goal"][:, None, None, :]
    tr = s * (1 - al) + g * al + (rng.normal(0, sigma, size=(E, n_per, H, s.shape[-1])) * np.sin(np.pi * al)).astype(np.float32)
    return tr.astype(np.float32)


def test_maze():
    rng = np.random.default_rng(5)
    for fam, hext in (("m2d", 0.5), ("m2d_nw3", 0.7), ("m2d_concave", 0.25)):
        E, n_per, H = 500, 20, 48
        pr = synth.make_problems(fam, E, 9)
        tr = make_trajs(pr, n_per, H, rng)
        tr = np.clip(tr, -0.03, 5.03).reshape(E * n_per, H, 2)
        env_id = np.repeat(np.arange(E, dtype=np.int32), n_per)
        for legacy in (True, False):
            t0 = time.time()
            v, c = colchk_ref.maze2d_swept(tr, env_id, pr["boxes"], hext, legacy_div=legacy)
            t_cpu = time.time() - t0
            vg, cg = pmp_b200.colchk_maze2d_swept(torch.tensor(tr, device=dev), env_id, pr["boxes"], hext, legacy_div=legacy)
            torch.cuda.synchronize()
            print("  %s swept legacy=%d: valid mismatches %d, count mismatches %d of %d (valid frac %.3f, cpu %.1f ms)" % (
                fam, legacy, int((vg.cpu().numpy() != v).sum()), int((cg.cpu().numpy() != c).sum()), len(v), v.mean(), 1e3 * t_cpu))
        trc = np.clip(tr, 0, 5)
        n, pt = colchk_ref.maze2d_points(trc, env_id, pr["boxes"], hext)
        ng, ptg = pmp_b200.colchk_maze2d_points(torch.tensor(trc, device=dev), env_id, pr["boxes"], hext)
        print("  %s points: ncol mismatches %d, ptcol mismatches %d" % (fam, int((ng.cpu().numpy() != n).sum()),
                                                                       int((ptg.cpu().numpy() != pt).sum())))
        ch, su, to = colchk_ref.pick_first_valid(v, c, E, n_per)
        chg, sug, tog = pmp_b200.pick_first_valid(vg, cg, E, n_per)
        print("  %s pick: mismatches %d %d %d  success rate %.3f" % (fam, int((chg.cpu().numpy() != ch).sum()),
              int((sug.cpu().numpy().astype(bool) != su).sum()), int((tog.cpu().numpy() != to).sum()), su.mean()))


def test_kuka():
    rng = np.random.default_rng(6)
    sph = colchk_ref.kuka_sphere_table()
    print("  sphere table rows:", len(sph))
    for fam, arms, hext in (("kuka", 1, 0.20), ("dualkuka", 2, 0.22)):
        E, n_per = 200, 10
        a = synth.arch(fam)
        H = a["horizon"]
        pr = synth.make_problems(fam, E, 4)
        tr = make_trajs(pr, n_per, H, rng, sigma=0.5).reshape(E * n_per, H, -1)
        env_id = np.repeat(np.arange(E, dtype=np.int32), n_per)
        t0 = time.time()
        ref = colchk_ref.kuka_points(tr, env_id, pr["boxes"], hext, n_arms=arms, sph=sph)
        t_cpu = time.time() - t0
        got = pmp_b200.colchk_kuka(torch.tensor(tr, device=dev), env_id, pr["boxes"], hext, arms,
                                   colchk_ref.kuka_bases(arms), sph)
        torch.cuda.synchronize()
        print("  %s: valid mism %d nchk mism %d ncol mism %d ptcol mism %d of %d trajs (valid frac %.3f, collide frac pts %.3f, cpu %.1f ms)" % (
            fam, *(int((got[k].cpu().numpy() != ref[k]).sum()) for k in ("valid", "nchk", "ncol", "ptcol")), len(tr),
            ref["valid"].mean(), ref["ptcol"].mean(), 1e3 * t_cpu))


def test_sample():
    fam = "m2d"
    a = synth.arch(fam)
    cfg = unet_ref.UnetCfg(a["unet"])
    sd = unet_ref.synth_weights(a["unet"], 1)
    B, H, D = 3, 48, 2
    pr = synth.make_problems(fam, B, 3)
    lo, hi = pr["lo"], pr["hi"]
    start = torch.tensor(synth.normalize(pr["start"], lo, hi)); goal = torch.tensor(synth.normalize(pr["goal"], lo, hi))
    walls = torch.tensor(pr["walls"])
    tb = sampler_ref.schedule(100)
    eng = pmp_b200.UnetEngine(a["unet"], sd, dev)
    wemb = eng.wall_embed(walls)
    for name, ddim in (("ddim8", 8), ("ddpm20", 0)):
        if ddim:
            ts = [int(v) for v in sampler_ref.ddim_timesteps(ddim, 100)]
            coefs = [[float(v) for v in ddim_coefs(tb, t, ddim, 100, 0.0)] for t in ts]
            noise = synth.noise(len(ts), B, H, D, 7)
            ref = sampler_ref.sample([(sd, cfg)], start, goal, [walls], noise, T=100, ddim_steps=ddim)
        else:
            # last 20 DDPM steps only (oracle is slow): start from t=19
            ts = list(range(19, -1, -1))
            coefs = [[float(v) for v in ddpm_coefs(tb, t)] for t in ts]
            noise = synth.noise(len(ts), B, H, D, 8)
            x = sampler_ref.apply_cond(noise[0].clone(), start, goal)
            for i, t in enumerate(ts):
                ec, eu = unet_ref.eps_pair(sd, cfg, x, torch.full((B,), t), walls)
                x = sampler_ref.apply_cond(sampler_ref.ddpm_update(tb, x, t, sampler_ref.compose([ec], [eu], [2.0]), noise[1 + i]), start, goal)
            ref = x
        got = pmp_b200.sample([eng], [wemb], [2.0], start, goal, H, 1 if ddim else 0, ts, coefs, noise=noise)
        print("  sample %s: max|diff| %.3e (range %.2f..%.2f)" % (name, (got.cpu() - ref).abs().max().item(), ref.min(), ref.max()))


if __name__ == "__main__":
    which = sys.argv[1:] or ["step", "maze", "kuka", "sample"]
    for w in which:
        print("[%s]" % w)
        globals()["test_" + w]()
