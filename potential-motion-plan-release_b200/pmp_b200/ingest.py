"""Problem-file ingest for the batched planners.

The reference evaluates on `<dataset>-problems.hdf5` files read with h5py into a flat dict
(`OfflineEnv.get_dataset(h5path, no_check=True)`, pb_diff_envs/environment/offline_env.py:83-103), of which the
planners use
  problems                 [n_env, n_prob, 2, D]        start / goal configurations   (kuka_plan_utils.py:36-57)
  infos/wall_locations     [n_env, n_prob, nw*wd]       flattened model input, already pre-processed for the unseen
                                                        mazes (kuka_plan_env.py:285-313), or [n_env, n_prob, nw, >=wd]
  maze_idx / is_sol_kp ... not needed by the sampler
`load_problems` returns that dict from an HDF5 file (needs h5py, which this image does not ship: the call then
fails loudly) or from an `.npz` archive with the same keys (what `save_problems_npz` writes; '/' in key names is
kept).  `problem_arrays` slices it into the arrays `diffuser.guides.plan_env_b200.plan_envs` takes.
"""
import os

import numpy as np


def _walk_h5(group, prefix, out):
    for k in group.keys():
        item = group[k]
        name = prefix + k
        if hasattr(item, "keys"):
            _walk_h5(item, name + "/", out)
        else:
            try:
                out[name] = item[:]
            except ValueError:      # scalar dataset
                out[name] = item[()]


def load_problems(path):
    """dict of numpy arrays keyed like the reference's `get_dataset(no_check=True)` result"""
    ext = os.path.splitext(path)[1].lower()
    if ext in (".hdf5", ".h5"):
        try:
            import h5py
        except ImportError as e:
            raise ImportError("reading %s needs h5py (not installed here); convert the file once with "
                              "pmp_b200.ingest.save_problems_npz on a machine that has it" % path) from e
        out = {}
        with h5py.File(path, "r") as f:
            _walk_h5(f, "", out)
        return out
    if ext == ".npz":
        with np.load(path, allow_pickle=False) as z:
            return {k: z[k] for k in z.files}
    raise ValueError("unknown problem-file type %r (expected .hdf5 or .npz)" % ext)


def save_problems_npz(problems_dict, path):
    """writes the dict `load_problems` returns as an .npz archive (same keys)"""
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in problems_dict.items()})


def problem_arrays(problems_dict, maze_idxs, prob_permaze, wall_dim, n_walls=None, prob_start_idx=0):
    """(starts [n_env,n_p,D], goals [n_env,n_p,D], walls [n_env,n_p,nw*wall_dim] float32 model input,
    wall_centers [n_env,nw,wall_dim] float64 for the collision checker) for the chosen environments.

    Follows kuka_obs_target_list (kuka_plan_utils.py:36-57: problems[maze_idx, :prob_permaze, 0 / 1]) and
    get_wallLoc_tensor_single (kuka_plan_env.py:285-313: wall_locations[maze_idx, start:start+prob_permaze],
    rows of >= 6 numbers kept as they are).  A 4-D wall array [.., nw, >=wall_dim] (training-set layout) is cut to
    the first `wall_dim` numbers of every wall and flattened.  The obstacle layout is constant within an
    environment, so the checker centres are taken from its first problem (first `n_walls` walls if given)."""
    P = np.asarray(problems_dict["problems"])
    W = np.asarray(problems_dict["infos/wall_locations"])
    maze_idxs = np.atleast_1d(np.asarray(maze_idxs, np.int64))
    if P.ndim != 4 or P.shape[2] != 2:
        raise ValueError("problems must be [n_env, n_prob, 2, D], got %s" % (P.shape,))
    if prob_start_idx + prob_permaze > P.shape[1]:
        raise ValueError("prob_permaze %d exceeds the %d problems per environment" % (prob_permaze, P.shape[1]))
    sl = slice(prob_start_idx, prob_start_idx + prob_permaze)
    starts = P[maze_idxs][:, sl, 0, :].astype(np.float32)
    goals = P[maze_idxs][:, sl, 1, :].astype(np.float32)
    Wm = W[maze_idxs][:, sl]
    if Wm.ndim == 4:
        Wm = Wm[..., :wall_dim].reshape(Wm.shape[0], Wm.shape[1], -1)
    if Wm.ndim != 3 or Wm.shape[-1] % wall_dim != 0:
        raise ValueError("wall_locations rows of %d numbers are not a multiple of wall_dim %d" % (Wm.shape[-1], wall_dim))
    walls = Wm.astype(np.float32)
    nw = Wm.shape[-1] // wall_dim if n_walls is None else n_walls
    centers = Wm[:, 0, :nw * wall_dim].reshape(len(maze_idxs), nw, wall_dim).astype(np.float64)
    return starts, goals, walls, centers
