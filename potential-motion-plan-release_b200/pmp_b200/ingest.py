"""Problem-file ingest for the batched planners.

The reference evaluates on `<dataset>-problems.hdf5` files read with h5py into a flat dict
(`OfflineEnv.get_dataset(h5path, no_check=True)`, pb_diff_envs/environment/offline_env.py:83-103), of which the
planners use
  problems                 [n_env, n_prob, 2, D]        start / goal configurations   (kuka_plan_utils.py:36-57)
  infos/wall_locations     [n_env, n_prob, nw*wd]       flattened model input, already pre-processed for the unseen
                                                        mazes (kuka_plan_env.py:285-313), or [n_env, n_prob, nw, >=wd]
  maze_idx / is_sol_kp ... not needed by the sampler
and, for the concave mazes, the obstacle layouts saved by the environment generator
  rand_rec2dgrp/<env_list_name>[_eval].npy      [n_env, 3*nw, 2] centres of the small squares   (rand_rec_group.py:163-187)
  rand_rec2dgrp/<env_list_name>_43[_eval].pkl   {'centers_43' [n_env,nw,2], 'mode_list' [n_env,nw] in 1..4}
                                                                                  (rand_rec_group_43.py:154-188)
from which the evaluation scripts build the (cx, cy, mode) model input (preprocessing.py:220-234).
`load_problems` returns that dict from an HDF5 file (through h5py when it is installed, else through the built-in
minimal reader pmp_b200/h5min.py, which covers the layouts h5py's defaults write) or from an `.npz` archive with the same keys (what `save_problems_npz` writes; '/' in key names is
kept).  `problem_arrays` slices it into the arrays `diffuser.guides.plan_env_b200.plan_envs` takes.
"""
import io
import os
import pickle

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
        except ImportError:
            # no h5py in this image: the built-in reader covers what h5py's defaults write (pmp_b200/h5min.py) and
            # raises H5FormatError for anything outside that subset
            from . import h5min
            return h5min.read_flat(path)
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


# ---------------------------------------------------------------------------------------------------------------
# concave mazes: every obstacle is 3 of the 4 quadrant squares around a centre; the model sees (cx, cy, mode)
# ---------------------------------------------------------------------------------------------------------------
# quadrant order of rand_rec_group_43.py:44-61 (1:(-,+) 2:(+,+) 4:(-,-) 3:(+,-) drawn as "1 2 / 4 3"); `mode` in
# 1..4 is the quadrant that is left out
_QUADS = np.array([[-1.0, 1.0], [1.0, 1.0], [1.0, -1.0], [-1.0, -1.0]])
_KEPT = np.array([[q for q in range(4) if q != m] for m in range(4)])      # [mode-1] -> the 3 kept quadrants, in order


def load_wall_layouts(path):
    """obstacle centres of every environment of a group, as saved by `save_rlg_list` (rand_rec_group.py:175-187):
    float64 [n_env, n_boxes, 2]"""
    a = np.load(path, allow_pickle=False)
    if a.ndim != 3:
        raise ValueError("wall layout file %s holds an array of shape %s, expected [n_env, n_boxes, dim]" % (path, a.shape))
    return a


def concave_tokens_from_boxes(boxes):
    """[n_env, 3*nw, 2] small-square centres -> [n_env, nw, 3] (cx, cy, mode) rows, mode in 1..4.

    Same arithmetic as save_center_and_mode / compute_mode_list (rand_rec_group_43.py:154-211), vectorised: with
    b0, b1, b2 the three squares of an obstacle, cx = (min(b0x, b2x) + max(b0x, b1x)) / 2, cy = (b0y + b2y) / 2, and
    the mode follows from which coordinates coincide."""
    b = np.asarray(boxes, np.float64)
    if b.ndim != 3 or b.shape[1] % 3 or b.shape[2] != 2:
        raise ValueError("concave layouts must be [n_env, 3*nw, 2], got %s" % (b.shape,))
    b = b.reshape(b.shape[0], -1, 3, 2)
    b0, b1, b2 = b[:, :, 0], b[:, :, 1], b[:, :, 2]
    cx = (np.minimum(b0[..., 0], b2[..., 0]) + np.maximum(b0[..., 0], b1[..., 0])) / 2
    cy = (b0[..., 1] + b2[..., 1]) / 2
    same_y01 = b0[..., 1] == b1[..., 1]
    same_x02 = b0[..., 0] == b2[..., 0]
    ok = np.where(same_y01, same_x02 | (b1[..., 0] == b2[..., 0]), same_x02 | (b0[..., 0] == b1[..., 0]))
    if not ok.all():
        raise ValueError("layout is not made of 3-square concave obstacles (%d bad)" % int((~ok).sum()))
    mode = np.where(same_y01, np.where(same_x02, 2, 3), np.where(same_x02, 1, 0)) + 1
    return np.stack([cx, cy, mode.astype(np.float64)], -1)


def concave_boxes(centers, modes, hext):
    """(centres [n_env,nw,2], modes [n_env,nw] in 1..4) -> small-square centres [n_env, 3*nw, 2] in the generator's
    order (sample_43square, rand_rec_group_43.py:44-61).  centre +- hext in float64: equal to the saved .npy layout
    up to the rounding of the generator's own (c - h + c + h) / 2, so prefer `load_wall_layouts` for the checker
    when the .npy is at hand."""
    c = np.asarray(centers, np.float64)
    m = np.asarray(modes).astype(np.int64)
    if m.min() < 1 or m.max() > 4:
        raise ValueError("concave modes must be in 1..4")
    off = _QUADS[_KEPT[m - 1]] * float(hext)                 # [n_env, nw, 3, 2]
    return (c[:, :, None, :] + off).reshape(c.shape[0], -1, 2)


class _NumpyOnlyUnpickler(pickle.Unpickler):
    """the layout pickles hold a dict of numpy arrays; nothing else is allowed to be constructed"""
    _OK = {("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
           ("numpy", "ndarray"), ("numpy", "dtype"),
           ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar")}

    def find_class(self, module, name):
        if (module, name) in self._OK:
            return super().find_class(module, name)
        raise pickle.UnpicklingError("refusing to load %s.%s from a layout pickle" % (module, name))


def load_concave_43(path):
    """(centres [n_env,nw,2] float64, modes [n_env,nw] int64 in 1..4) from `<env_list_name>_43[_eval].pkl`
    (written by save_center_and_mode, rand_rec_group_43.py:176-186)"""
    with open(path, "rb") as f:
        d = _NumpyOnlyUnpickler(io.BytesIO(f.read())).load()
    c = np.asarray(d["centers_43"], np.float64)
    m = np.asarray(d["mode_list"]).astype(np.int64)
    if c.ndim != 3 or c.shape[2] != 2 or m.shape != c.shape[:2]:
        raise ValueError("unexpected shapes in %s: centers_43 %s, mode_list %s" % (path, c.shape, m.shape))
    return c, m


def concave_wall_locations(centers, modes, n_probs):
    """model input of the concave evaluation problems, [n_env, n_probs, nw, 3] rows (cx, cy, mode): what
    preproc_val_srm2dwloc_43 (preprocessing.py:220-234) puts into problems_dict['infos/wall_locations'];
    `problem_arrays(..., wall_dim=3)` flattens it."""
    w = np.concatenate([np.asarray(centers, np.float64), np.asarray(modes, np.float64)[..., None]], -1)
    return np.repeat(w[:, None], n_probs, axis=1)


def concave_problem_arrays(problems_dict, maze_idxs, prob_permaze, centers_modes=None, layouts=None, hext=0.25,
                           prob_start_idx=0):
    """`problem_arrays` for the concave mazes: (starts, goals, walls [n_env,n_p,nw*3] float32 rows (cx, cy, mode),
    boxes [n_env, 3*nw, 2] float64 small-square centres for the collision checker).

    The reference overwrites the problem file's wall array with the (cx, cy, mode) rows of the `_43_eval.pkl`
    before flattening (plan_pb_diff_helper.py:128-134, preprocessing.py:220-234).  Here the rows come from
    `centers_modes` = `load_concave_43(...)` when given, else they are recomputed from the layout -- `layouts`
    (`load_wall_layouts(...)`, indexed by environment like the problem file) or the problem file's own
    [n_env, n_p, 3*nw, 2] wall array -- with the generator's arithmetic (`concave_tokens_from_boxes`).  The checker
    squares are the layout when there is one, else centre +- hext."""
    maze_idxs = np.atleast_1d(np.asarray(maze_idxs, np.int64))
    W = np.asarray(problems_dict["infos/wall_locations"])
    raw_layout = W.ndim == 4 and W.shape[-1] == 2 and W.shape[2] % 3 == 0       # [n_env, n_p, 3*nw, 2] squares
    if layouts is not None:
        boxes = np.asarray(layouts, np.float64)[maze_idxs]
    elif raw_layout:
        boxes = W[maze_idxs][:, prob_start_idx].astype(np.float64)
    else:
        boxes = None
    if centers_modes is not None:
        cen, modes = centers_modes
        tokens = np.concatenate([np.asarray(cen, np.float64), np.asarray(modes, np.float64)[..., None]], -1)[maze_idxs]
    elif boxes is not None:
        tokens = concave_tokens_from_boxes(boxes)
    elif W.ndim == 4 and W.shape[-1] == 3:                                       # already (cx, cy, mode) rows
        tokens = W[maze_idxs][:, prob_start_idx].astype(np.float64)
    else:
        raise ValueError("no concave layout: give centers_modes (the _43 pickle), layouts (the .npy) or a problem file "
                         "whose infos/wall_locations is [n_env, n_prob, 3*nw, 2] or [n_env, n_prob, nw, 3]")
    if boxes is None:
        boxes = concave_boxes(tokens[..., :2], tokens[..., 2], hext)
    d = {"problems": np.asarray(problems_dict["problems"])[maze_idxs],
         "infos/wall_locations": np.repeat(tokens[:, None], np.asarray(problems_dict["problems"]).shape[1], axis=1)}
    starts, goals, walls, _ = problem_arrays(d, np.arange(len(maze_idxs)), prob_permaze, 3, prob_start_idx=prob_start_idx)
    return starts, goals, walls, boxes
