"""CPU restatement of the training step of the potential diffusion model: the energy loss
`0.5 * mean(w * (dE/dx - noise)^2)` (double backward through the U-Net), Adam and the EMA copy.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows (reference file:line):
  * GaussianDiffusionPB.q_sample / p_losses / loss     diffuser/models/diffusion_pb.py:468-545
  * TemporalUnet_WCond.forward, training branch         diffuser/models/temporal_cond.py:245-329
    (concept dropout :266-270, `torch.autograd.grad(..., create_graph=True)` :323-326)
  * WeightedStateLoss / WeightedStateL2                 diffuser/models/helpers.py:201-214, 226-254
  * apply_conditioning / set_loss_noise                 diffuser/models/helpers.py:168-178
  * Trainer.train (Adam step, zero_grad) and step_ema   diffuser/utils/training_pbdiff.py:108-143
  * EMA.update_model_average                            diffuser/utils/train_utils.py:15-31

Pinned against the live reference (`GaussianDiffusionPB.loss` + `loss.backward()` run through
oracle/refshim.py, frozen as tests/golden/train_*.npz by tests/golden/gen_golden.py).
The random draws of the reference's `loss` (t, noise, the concept-dropout mask) are inputs here.
"""
import torch

from . import unet_ref


def q_sample(sqrt_acp, sqrt_1m_acp, x0, t, noise):
    """diffusion_pb.py:468-478"""
    return sqrt_acp[t][:, None, None] * x0 + sqrt_1m_acp[t][:, None, None] * noise


def loss_and_grads(sd, cfg, x_noisy, t, walls, noise, keep_mask, loss_weights, dtype=torch.float32):
    """loss (the value `loss` returns: 0.5 * weighted MSE) and d loss / d parameter for every key of `sd`.

    x_noisy [B,H,D] (already conditioned), t [B] int64, walls [B,Wd], noise [B,H,D] (target; rows of conditioned
    waypoints already zeroed when set_cond_noise_to_0), keep_mask [B] (0 where the concept was dropped),
    loss_weights [H,D]."""
    P = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    x = x_noisy.detach().to(dtype).clone().requires_grad_(True)
    w = unet_ref.wall_embed(P, cfg, walls.to(dtype))
    w = w * keep_mask.to(dtype)[:, None]
    u = unet_ref.unet_u(P, cfg, x, t, w)
    energy = (0.5 * u * u).sum()
    eps = torch.autograd.grad([energy], [x], create_graph=True)[0]
    loss = 0.5 * ((eps - noise.to(dtype)) ** 2 * loss_weights.to(dtype)).mean()
    keys = list(P.keys())
    grads = torch.autograd.grad(loss, [P[k] for k in keys], allow_unused=True)
    g = {k: (torch.zeros_like(P[k]) if gi is None else gi.detach()) for k, gi in zip(keys, grads)}
    return loss.detach(), g, eps.detach()


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (no weight decay, no amsgrad) on flat tensors, in place; `step` counts from 1"""
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


def ema_step(ema, p, beta):
    """train_utils.py:27-31: old * beta + (1 - beta) * new"""
    ema.mul_(beta).add_(p, alpha=1 - beta)
