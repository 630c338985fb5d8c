"""PolicyCompose: drop-in for diffuser/guides/policies_compose.py:11-459 -- composes K learned
potentials: eps = eps_u[base] + sum_k w_k (eps_c,k - eps_u,k), each evaluated on its own
subset of the obstacles.  The whole loop (all K U-Nets per step + the fused composition /
DDPM / DDIM update) runs inside one pmp_sample call."""
import numpy as np
import torch

import diffuser.utils as utils
import pmp_b200
from diffuser.models import GaussianDiffusionPB
from diffuser.models.helpers import apply_conditioning
from .policies import Policy, Trajectories  # noqa: F401


class PolicyCompose(Policy):

    def __init__(self, diffusion_model_list, normalizer_list, use_normed_wallLoc, use_ddim, po_config):
        self.diffusion_model_list = diffusion_model_list
        self.normalizer_list = normalizer_list
        self.action_dim = normalizer_list[0].action_dim
        self.n_models = len(diffusion_model_list)
        assert self.n_models >= 2, 'composition needs at least two potentials'
        self.num_walls_c = torch.tensor(po_config['num_walls_c'])
        self.wall_dim = po_config['wall_dim']
        self.comp_type = po_config['comp_type']
        assert self.comp_type in ['direct', 'share', 'share-model2fix']
        self.use_ddim = use_ddim
        self.ddim_steps = po_config['ddim_steps']
        self.wall_is_dyn = po_config.get('wall_is_dyn', None)
        self.is_dyn_env = self.wall_is_dyn is not None
        if self.is_dyn_env:
            raise NotImplementedError('dynamic-wall composition is outside the built hot path')
        self.uncond_base_idx = po_config['uncond_base_idx']
        self.ddim_eta = po_config['ddim_eta']
        self.use_normed_wallLoc = use_normed_wallLoc
        self.w_split_sizes = tuple(int(v) for v in (self.num_walls_c * self.wall_dim))
        self.diffusion_model = diffusion_model_list[0]
        for dm in diffusion_model_list:
            assert dm.horizon == self.diffusion_model.horizon
            dm.eval()
        self.is_same_model = all(dm is diffusion_model_list[0] for dm in diffusion_model_list)
        self.normalizer = normalizer_list[0]
        self.n_timesteps = self.diffusion_model.n_timesteps
        self.horizon = self.diffusion_model.horizon
        self.transition_dim = self.diffusion_model.transition_dim
        self.observation_dim = self.diffusion_model.observation_dim
        self.use_mala_sampler = False
        for n in normalizer_list:
            assert (n.normalizers['observations'].mins == normalizer_list[0].normalizers['observations'].mins).all()
            assert (n.normalizers['observations'].maxs == normalizer_list[0].normalizers['observations'].maxs).all()
        self.return_diffusion = False

    def _format_conditions(self, conditions, batch_size, norm_idx):
        conditions = utils.apply_dict(self.normalizer_list[norm_idx].normalize, conditions, 'observations')
        if batch_size > 0:
            raise RuntimeError
        return utils.to_torch(conditions, dtype=torch.float32, device=self.device)

    def wallLoc_preproc(self, wall_c):
        """[B, nw*dim] -> list (K) of per-model wall vectors (policies_compose.py:95-181).
        'direct': disjoint split by num_walls_c; a second model given one (two) obstacles gets them padded to three
        with jittered copies, as the reference does for its 3-obstacle model.  'share' / 'share-model2fix' (two
        models): the first num_walls_c[0] obstacles, and the last num_walls_c[0] (num_walls_c[1]) obstacles."""
        if self.comp_type == 'direct':
            wall_c = list(torch.split(wall_c, self.w_split_sizes, dim=1))
            if self.w_split_sizes[1] == 2:
                tmp = torch.repeat_interleave(wall_c[1], repeats=2, dim=1)
                tmp = tmp + torch.randn_like(tmp) * 0.03
                wall_c[1] = torch.cat([wall_c[1], tmp], dim=1)
            elif self.w_split_sizes[1] == 4:
                tmp = torch.clone(wall_c[1][:, :2])
                tmp = tmp + torch.randn_like(tmp) * 0.03
                wall_c[1] = torch.cat([wall_c[1], tmp], dim=1)
            return wall_c
        assert self.n_models == 2
        B, nwd = wall_c.shape
        nw = round(nwd / self.wall_dim)
        wall_c = wall_c.reshape(B, -1, self.wall_dim)
        first = torch.clone(wall_c[:, :self.num_walls_c[0], :]).flatten(1, 2)
        if self.comp_type == 'share':
            from_idx = nw - int(self.num_walls_c[0])
        else:
            from_idx = max(0, nw - int(self.num_walls_c[1]))
        second = torch.clone(wall_c[:, from_idx:, :]).flatten(1, 2)
        if self.comp_type == 'share':
            assert first.shape == second.shape
        return [first, second]

    def call_multi_horizon(self, cond_mulho, wall_c):
        self.multi_hor_x0_perstep = []
        out = []
        for cond in cond_mulho:
            if torch.is_tensor(wall_c):
                wall_c = wall_c.detach()
            out.append(self(cond, wall_c)[1])
            self.multi_hor_x0_perstep.append(getattr(self, 'x0_perstep', None))
        return out

    def __call__(self, cond, wall_c):
        assert len(cond[0].shape) == 2
        cond_list = [self._format_conditions(cond, -1, i) for i in range(self.n_models)]
        B = cond_list[0][0].shape[0]
        wall_c = self.wallLoc_preproc(torch.as_tensor(wall_c, dtype=torch.float32))
        horizon = sorted(cond_list[0].keys())[-1] + 1
        assert horizon >= 20, 'prevent bugs'
        dm0 = self.diffusion_model_list[0]
        dev = self.device
        shape = (B, horizon, self.observation_dim)
        engines, wembs, w = [], [], []
        for i, dm in enumerate(self.diffusion_model_list):
            wl = wall_c[i]
            if self.use_normed_wallLoc:
                wl = self.normalizer_list[i].normalize(wl, 'infos/wall_locations')
            eng = dm.model.engine(dm.n_timesteps)
            engines.append(eng)
            wembs.append(eng.wall_embed(wl.to(dev)))
            w.append(float(dm.condition_guidance_w))
        if self.use_ddim:
            ts = dm0.ddim_set_timesteps(self.ddim_steps)
            coefs = dm0.step_coefs_ddim(ts, self.ddim_eta, self.ddim_steps)
            kind, noisy = 1, self.ddim_eta > 0
        else:
            ts = np.arange(self.n_timesteps - 1, -1, -1)
            coefs = dm0.step_coefs_ddpm(ts)
            kind, noisy = 0, True
        draws = [torch.randn(shape, device=dev)]
        for _ in range(len(ts)):
            draws.append(torch.randn(shape, device=dev) if noisy else torch.zeros(1, device=dev).expand(shape))
        start, goal = dm0._start_goal(cond_list[0], horizon)
        out = pmp_b200.sample(engines, wembs, w, start, goal, horizon, kind, ts, coefs, noise=torch.stack(draws, 0),
                              base=self.uncond_base_idx, var_temp=dm0.var_temp, trace=self.return_diffusion)
        if self.return_diffusion:
            latents, tr = out
            self.x0_perstep = [self.output_traj_postproc(tr[:, i])[1] for i in range(1, tr.shape[1])]
        else:
            latents = out
        latents = apply_conditioning(latents, cond_list[0], 0)
        return self.output_traj_postproc(latents)

    def models_pred(self, x_t, cond_list, t, wall_c):
        """per-model (eps_cond, eps_uncond) stacks [K,B,H,D] (policies_compose.py:353-391)"""
        ecs, eus = [], []
        for i, dm in enumerate(self.diffusion_model_list):
            wl = wall_c[i]
            if self.use_normed_wallLoc:
                wl = self.normalizer_list[i].normalize(wl, 'infos/wall_locations')
            (_, ec), (_, eu) = dm.pred_unet(x_t.detach(), cond_list[i], t, wl)
            ecs.append(ec)
            eus.append(eu)
        return (None, torch.stack(ecs, 0)), (None, torch.stack(eus, 0))

    def ddim_replan(self, traj, cond, wall_c, dn_steps):
        """composed re-sampling of given trajectories (policies_compose.py:399-447): q_sample to the dn_steps-th
        last DDIM timestep with the first model's schedule, then the last dn_steps composed DDIM updates; x_t is
        conditioned before every evaluation and not after the last update, as in the reference.  traj [H,D] or
        [B,H,D] normalised (the reference takes one trajectory; a batch is re-sampled together here), cond
        {0: [B,D], H-1: [B,D]} normalised, wall_c [B, sum_k nw_k*dim]."""
        dm0 = self.diffusion_model_list[0]
        dev = self.device
        input2d = traj.ndim == 2
        x = dm0.ddim_get_noisy_traj(traj.to(dev), self.ddim_steps, dn_steps)
        cond_list = [cond for _ in range(self.n_models)]
        wall_c = self.wallLoc_preproc(torch.as_tensor(wall_c, dtype=torch.float32).to(dev))
        w = [float(dm.condition_guidance_w) for dm in self.diffusion_model_list]
        ts = dm0.ddim_set_timesteps(self.ddim_steps)
        for t in ts[-dn_steps:]:
            tt = torch.full((x.shape[0],), int(t), device=dev, dtype=torch.long)
            (_, ec), (_, eu) = self.models_pred(x, cond_list, tt, wall_c)          # conditions x in place
            coef = dm0.step_coefs_ddim([int(t)], self.ddim_eta, self.ddim_steps)[0]
            noise = torch.randn_like(x) if self.ddim_eta > 0 else None
            x = pmp_b200.step(1, x.contiguous(), list(ec), list(eu), w, coef, noise=noise, base=self.uncond_base_idx)
        return x[0] if input2d else x

    def set_ddim_steps(self, ddim_steps):
        self.ddim_steps = ddim_steps
        for dm in self.diffusion_model_list:
            dm.ddim_num_inference_steps = ddim_steps


from .._overlay import reference_fallback as _reference_fallback  # noqa: E402

__getattr__ = _reference_fallback(__name__)   # overlay mode: symbols outside the hot path come from the reference's file
