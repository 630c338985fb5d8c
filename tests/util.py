import os

import numpy as np
import torch

from oracle import sampler_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Tolerances (stated once, used by every parity test).
# eps = dE/dx: the north star asks for rtol 1e-3 vs the reference run in fp32.  Strict
# element-wise rtol is ill-posed for near-zero entries (fp32-vs-fp64 of the reference itself
# violates it), so the acceptance test is |got-ref| <= EPS_RTOL*|ref| + EPS_ATOL_FRAC*max|ref|.
EPS_RTOL = 1e-3
EPS_ATOL_FRAC = 1e-4
# final trajectories in normalised coordinates ([-1,1]); 100 DDPM steps amplify eps error a little
TRAJ_ATOL = 2e-4


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def assert_eps_close(got, ref, what=""):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    tol = EPS_RTOL * np.abs(ref) + EPS_ATOL_FRAC * np.abs(ref).max()
    bad = np.abs(got - ref) > tol
    assert not bad.any(), "%s: %d/%d elements out of tolerance, max err %.3e (max|ref| %.3e)" % (
        what, bad.sum(), bad.size, np.abs(got - ref).max(), np.abs(ref).max())


def ddpm_coefs(tb, ts):
    out = np.zeros((len(ts), 8), np.float32)
    for i, t in enumerate(ts):
        sig = (0.5 * tb["posterior_log_variance_clipped"][t]).exp() * (0.0 if t == 0 else 1.0)
        out[i, :5] = [tb["sqrt_recip_alphas_cumprod"][t], tb["sqrt_recipm1_alphas_cumprod"][t],
                      tb["posterior_mean_coef1"][t], tb["posterior_mean_coef2"][t], sig]
    return out


def ddim_coefs(tb, ts, n, T, eta):
    out = np.zeros((len(ts), 8), np.float32)
    for i, t in enumerate(ts):
        c = sampler_ref.ddim_coefs(tb, int(t), n, T, eta)
        out[i] = [c["recip"], c["recipm1"], c["sqrt_a"], c["sqrt_b"], c["sqrt_a_prev"], c["dirc"], c["std"],
                  1.0 if eta > 0 else 0.0]
    return out


def pad_noise(noise, n_steps):
    """golden DDIM (eta=0) fixtures hold only the x_T draw; pad zeros for the unused per-step draws"""
    noise = torch.as_tensor(noise)
    if noise.shape[0] == n_steps + 1:
        return noise
    z = torch.zeros((n_steps + 1 - noise.shape[0],) + tuple(noise.shape[1:]))
    return torch.cat([noise, z], 0)
