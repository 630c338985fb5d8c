"""RM2DCollisionChecker: interface of diffuser/guides/rm2d_colchk.py:12-94 on the GPU.
`env` is any object with `wall_centers` [nw,2] float64, `wall_hext` (scalar or [nw,2]) and
optionally `maze_size` (default 5x5) -- e.g. a pb_diff_envs maze env adapter."""
import numpy as np
import torch

import pmp_b200
from .kuka_colchk import KukaCollisionChecker


class RM2DCollisionChecker(KukaCollisionChecker):
    """static Maze2D: swept-segment check (eps 0.08 in the planners) or per-waypoint check"""

    def __init__(self, normalizer=None, device='cuda:0', legacy_div=True):
        self.normalizer = normalizer
        self.device = torch.device(device)
        self.num_collision_check = 0
        self.use_collision_eps = False
        self.collision_eps = 0.03
        self.min_to_wall_dist = 0.01
        # numpy 1.23.4 (the reference's pin) evaluates K=int(d/eps) in float64; numpy>=2 in float32
        self.legacy_div = legacy_div

    @staticmethod
    def _env_arrays(env):
        cen = np.asarray(env.wall_centers, np.float64)
        hext = np.broadcast_to(np.asarray(env.wall_hext, np.float64), cen.shape)
        size = np.asarray(getattr(env, 'maze_size', (5., 5.)), np.float32)
        return cen, hext, size

    def check_batch(self, trajs, env_ids, cen, hext, maze_size=(5., 5.)):
        """trajs [N,H,2] un-normalised -> (valid u8 [N], count i32 [N]) cuda tensors; semantics of
        check_single_traj applied to every trajectory"""
        t = torch.as_tensor(trajs, dtype=torch.float32).to(self.device)
        if self.use_collision_eps:
            return pmp_b200.colchk_maze2d_swept(t, env_ids, cen, hext, eps=self.collision_eps,
                                                margin=self.min_to_wall_dist, legacy_div=self.legacy_div,
                                                hi=(float(maze_size[0]), float(maze_size[1])))
        ncol, pt = pmp_b200.colchk_maze2d_points(t, env_ids, cen, hext, margin=self.min_to_wall_dist,
                                                 hi=(float(maze_size[0]), float(maze_size[1])))
        if bool((ncol < 0).any()):
            raise AssertionError('waypoint outside the maze limits (point2d_robot.py:43-44)')
        # per-waypoint early exit: count = index of the first colliding waypoint + 1
        first = torch.where(pt.bool().any(1), pt.float().argmax(1) + 1, torch.full_like(ncol, t.shape[1]))
        return (ncol == 0).to(torch.uint8), first.to(torch.int32)

    def check_single_traj(self, traj, env):
        assert traj.ndim == 2 and type(traj) == np.ndarray
        cen, hext, size = self._env_arrays(env)
        valid, cnt = self.check_batch(traj[None], np.zeros(1, np.int32), cen[None], hext[None], size)
        return bool(valid.item()), int(cnt.item())

    def check_1_traj_eps(self, traj, env):
        keep = self.use_collision_eps
        self.use_collision_eps = True
        try:
            return self.check_single_traj(traj, env)
        finally:
            self.use_collision_eps = keep


from .._overlay import reference_fallback as _reference_fallback  # noqa: E402

__getattr__ = _reference_fallback(__name__)   # overlay mode: symbols outside the hot path come from the reference's file
