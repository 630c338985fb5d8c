"""Training-step oracle (oracle/train_ref.py) pinned against the reference's own `GaussianDiffusionPB.loss` +
`loss.backward()` (tests/golden/train_*.npz, written by tests/golden/gen_golden.py:gen_train from the live reference):
loss value and a fingerprint of every parameter gradient.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import sampler_ref, synth, train_ref, unet_ref
from util import golden


def grad_digest(name, g):
    """same fingerprint as tests/golden/gen_golden.py:grad_digest"""
    g = g.detach().double().reshape(-1)
    gen = torch.Generator().manual_seed(synth._name_seed(name, 77) % (2 ** 31))
    r1 = torch.randn(g.numel(), generator=gen, dtype=torch.float64)
    r2 = torch.randn(g.numel(), generator=gen, dtype=torch.float64)
    head = torch.zeros(32, dtype=torch.float64)
    head[:min(32, g.numel())] = g[:32]
    return np.concatenate([[g.norm().item(), g.sum().item(), (g * r1).sum().item(), (g * r2).sum().item()], head.numpy()])


def golden_batch(fam):
    """(cfg, sd, x_noisy, t, walls, noise, keep, loss_weights, golden) of the reference's training call"""
    z = golden("train_%s.npz" % fam)
    a = synth.arch(fam)
    cfg = unet_ref.UnetCfg(a["unet"])
    sd = unet_ref.synth_weights(a["unet"], int(z["weight_seed"]))
    x0, t, noise = torch.tensor(z["x"]), torch.tensor(z["t"]), torch.tensor(z["noise"])
    tb = sampler_ref.schedule(a["diffusion"]["n_timesteps"])
    xn = train_ref.q_sample(tb["sqrt_alphas_cumprod"], tb["sqrt_one_minus_alphas_cumprod"], x0, t, noise)
    xn[:, 0], xn[:, -1] = x0[:, 0], x0[:, -1]            # apply_conditioning (diffusion_pb.py:484-485)
    walls = torch.tensor(z["walls"])[:, 0, :]            # diffusion_pb.py:534
    return cfg, sd, xn, t, walls, noise, torch.tensor(z["keep"]), torch.tensor(z["loss_weights"]), z


@pytest.mark.parametrize("fam", ["m2d", "dualkuka"])
def test_training_oracle_matches_the_reference_loss_and_gradients(fam):
    cfg, sd, xn, t, walls, noise, keep, lw, z = golden_batch(fam)
    loss, grads, _ = train_ref.loss_and_grads(sd, cfg, xn, t, walls, noise, keep, lw)
    assert abs(float(loss) - float(z["loss"])) <= 1e-6 * abs(float(z["loss"]))
    names = [str(n) for n in z["names"]]
    assert set(names) == set(sd.keys())
    gn = float(np.sqrt((z["digest"][:, 0] ** 2).sum()))
    for i, k in enumerate(names):
        d = grad_digest(k, grads[k])
        ref = z["digest"][i]
        scale = max(ref[0], 1e-6 * gn)
        # same fp32 torch ops in the same order as the reference: equal up to reduction-order noise
        assert abs(d[0] - ref[0]) <= 2e-4 * scale, (k, d[0], ref[0])
        assert np.abs(d[2:4] - ref[2:4]).max() <= 2e-3 * scale + 6e-3 * scale, (k, d[:4], ref[:4])
        assert np.abs(d[4:] - ref[4:]).max() <= 5e-4 * max(np.abs(ref[4:]).max(), 1e-5 * gn), k


def test_adam_and_ema_restatement_match_torch():
    torch.manual_seed(0)
    p = torch.randn(1000); q = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([q], lr=2e-4)
    m, v, ema = torch.zeros(1000), torch.zeros(1000), p.clone()
    for step in range(1, 6):
        g = torch.randn(1000)
        q.grad = g.clone()
        opt.step()
        train_ref.adam_step(p, g, m, v, step, 2e-4)
        train_ref.ema_step(ema, p, 0.995)
        assert torch.allclose(p, q.detach(), rtol=1e-6, atol=1e-8)
    assert torch.isfinite(ema).all()
