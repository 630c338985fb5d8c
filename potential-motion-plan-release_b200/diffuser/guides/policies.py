"""Policy: drop-in for diffuser/guides/policies.py:11-162 -- normalise the start/goal
conditions, run the reverse diffusion on the GPU, un-normalise to numpy trajectories."""
from collections import namedtuple

import einops
import numpy as np
import torch

import diffuser.utils as utils
from diffuser.models import GaussianDiffusionPB

Trajectories = namedtuple('Trajectories', 'actions observations')


class Policy:

    def __init__(self, diffusion_model, normalizer, use_ddim=False):
        self.diffusion_model = diffusion_model
        self.diffusion_model.eval()
        self.use_ddim = use_ddim
        self.normalizer = normalizer
        self.action_dim = normalizer.action_dim

    @property
    def device(self):
        return next(self.diffusion_model.parameters()).device

    def _format_conditions(self, conditions, batch_size):
        conditions = utils.apply_dict(self.normalizer.normalize, conditions, 'observations')
        conditions = utils.to_torch(conditions, dtype=torch.float32, device=self.device)
        if batch_size > 0:
            conditions = utils.apply_dict(einops.repeat, conditions, 'd -> repeat d', repeat=batch_size)
        return conditions

    def call_multi_horizon(self, cond_mulho, wall_locations, use_normed_wallLoc):
        out = []
        for cond in cond_mulho:
            if torch.is_tensor(wall_locations):
                wall_locations = wall_locations.detach()
            out.append(self(cond, batch_size=-1, wall_locations=wall_locations,
                            use_normed_wallLoc=use_normed_wallLoc)[1])
        return out

    def __call__(self, conditions, debug=False, batch_size=1, wall_locations=None, use_normed_wallLoc=False,
                 return_diffusion=False):
        assert not self.diffusion_model.training
        conditions = self._format_conditions(conditions, batch_size)
        self.diffusion_model.horizon = sorted(conditions.keys())[-1] + 1
        if type(self.diffusion_model) != GaussianDiffusionPB:
            raise NotImplementedError('type: %s' % type(self.diffusion_model))
        assert wall_locations is not None
        if use_normed_wallLoc:
            wall_locations = self.normalizer.normalize(wall_locations, 'infos/wall_locations')
        wall_locations = torch.as_tensor(wall_locations, dtype=torch.float32)
        sample = self.diffusion_model(conditions, wall_locations, return_diffusion=return_diffusion,
                                      use_ddim=self.use_ddim)
        if return_diffusion:
            sample, diffusion = sample
        action, trajectories = self.output_traj_postproc(sample)
        if return_diffusion:
            dfu_obs = np.stack([self.output_traj_postproc(diffusion[:, i])[1].observations
                                for i in range(diffusion.shape[1])], axis=1)
            return action, trajectories, [], dfu_obs
        return action, trajectories

    def output_traj_postproc(self, sample):
        """normalised cuda tensor [B,H,D] -> (first action, Trajectories(actions, observations))"""
        assert isinstance(self.diffusion_model, GaussianDiffusionPB)
        bsize, horizon = sample.shape[0:2]
        actions = np.zeros(shape=(bsize, horizon, self.diffusion_model.action_dim))
        observations = self.normalizer.unnormalize(utils.to_np(sample), 'observations')
        return actions[0, 0], Trajectories(actions, observations)


from .._overlay import reference_fallback as _reference_fallback  # noqa: E402

__getattr__ = _reference_fallback(__name__)   # overlay mode: symbols outside the hot path come from the reference's file
