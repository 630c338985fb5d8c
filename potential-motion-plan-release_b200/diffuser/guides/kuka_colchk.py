"""KukaCollisionChecker: interface of diffuser/guides/kuka_colchk.py:81-104 on the GPU.

DECLARED DEVIATION: the reference asks pybullet 3.2.5 for mesh contacts
(pb_diff_envs/robot/individual_robot.py:48-95); this build uses the analytic sphere-link
model defined in oracle/colchk.c (link spheres from the iiwa collision-mesh AABBs, boxes,
table plane, arm-vs-arm), so success statistics are "sphere-model" numbers.
`env` is any object with: voxel_centers [nv,3], voxel_hext (scalar or [nv,3]), n_arms."""
import numpy as np
import torch

import pmp_b200

# rows (link, cx, cy, cz, r): see oracle/colchk_ref.py:kuka_sphere_table (generated, frozen here)
KUKA_LINK_AABB = [
    ((-0.136, -0.121, 0.000), (0.121, 0.121, 0.158)), ((-0.086, -0.116, -0.010), (0.086, 0.085, 0.288)),
    ((-0.085, -0.086, -0.068), (0.086, 0.205, 0.132)), ((-0.068, -0.068, -0.010), (0.068, 0.115, 0.283)),
    ((-0.068, -0.068, -0.068), (0.068, 0.185, 0.114)), ((-0.068, -0.068, -0.010), (0.068, 0.099, 0.267)),
    ((-0.066, -0.086, -0.071), (0.066, 0.081, 0.067)), ((-0.052, -0.052, -0.010), (0.052, 0.052, 0.045))]
TABLE_Z = -0.001


def sphere_table(conservative=False):
    """rows (link, cx, cy, cz, r) of the sphere-link model, at most 32.

    default: the model the oracle defines (oracle/colchk_ref.py:kuka_sphere_table): spheres of the mean cross-section half
    extent strung along the long axis of every link's collision-mesh AABB -- tight, NOT conservative (box corners and the
    gaps between spheres are not covered), so a path it accepts can graze an obstacle the reference's mesh query reports.
    conservative=True: up to 4 spheres per link whose union CONTAINS the link's AABB (radius^2 = half-diagonal^2 of the
    cross-section + (segment / 2)^2), hence the link mesh: a path this table accepts is collision free for the real
    geometry as well (voxels and table are exact boxes / a plane); the price is false alarms near obstacles."""
    rows = []
    for l, (mn, mx) in enumerate(KUKA_LINK_AABB):
        mn = np.array(mn); mx = np.array(mx)
        ext = mx - mn
        ax = int(np.argmax(ext))
        o = [i for i in range(3) if i != ax]
        ctr = 0.5 * (mn + mx)
        if conservative:
            hd2 = 0.25 * (ext[o[0]] ** 2 + ext[o[1]] ** 2)
            n = int(min(4, max(1, np.ceil(ext[ax] / np.sqrt(hd2)))))
            seg = ext[ax] / n
            r = float(np.sqrt(hd2 + 0.25 * seg * seg)) * (1 + 1e-6)
            for j in range(n):
                c = ctr.copy()
                c[ax] = mn[ax] + (j + 0.5) * seg
                rows.append([l, c[0], c[1], c[2], r])
            continue
        r = min(0.10, 0.25 * (ext[o[0]] + ext[o[1]]))
        n = max(1, int(round(ext[ax] / (1.5 * r))))
        reff = min(r, ext[ax] / 2)
        for j in range(n):
            c = ctr.copy()
            if n > 1:
                c[ax] = mn[ax] + reff + (ext[ax] - 2 * reff) * j / (n - 1)
            rows.append([l, c[0], c[1], c[2], r])
    return np.array(rows, np.float32)


def arm_bases(n_arms):
    return np.zeros((1, 3), np.float32) if n_arms == 1 else np.array([[-0.5, 0, 0], [0.5, 0, 0]], np.float32)


class KukaCollisionChecker:
    def __init__(self, normalizer=None, device='cuda:0', conservative=False):
        self.normalizer = normalizer
        self.device = torch.device(device)
        self.use_collision_eps = False
        self.collision_eps = 0.03
        self.conservative = conservative      # True: spheres that contain the link AABBs (see sphere_table)
        self.sph = sphere_table(conservative)

    def check_batch(self, trajs, env_ids, boxc, hext, n_arms):
        """trajs [N,H,7*n_arms] (numpy or cuda tensor, un-normalised) -> dict of cuda tensors
        valid u8 [N], nchk i32 [N], ncol i32 [N], ptcol u8 [N,H]"""
        t = torch.as_tensor(trajs, dtype=torch.float32).to(self.device)
        return pmp_b200.colchk_kuka(t, env_ids, boxc, hext, n_arms, arm_bases(n_arms), self.sph, TABLE_Z)

    def check_single_traj(self, traj, env):
        """(no_collision: bool, num_collision_check: int) for one [H,D] numpy trajectory"""
        assert traj.ndim == 2 and type(traj) == np.ndarray
        out = self.check_batch(traj[None], np.zeros(1, np.int32), np.asarray(env.voxel_centers, np.float32)[None],
                               env.voxel_hext, env.n_arms)
        return bool(out['valid'].item()), int(out['nchk'].item())


from .._overlay import reference_fallback as _reference_fallback  # noqa: E402

__getattr__ = _reference_fallback(__name__)   # overlay mode: symbols outside the hot path come from the reference's file
