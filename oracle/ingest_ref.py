"""CPU restatement of the reference's concave-maze layout handling -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows, loop for loop:
  RandRectangleGroup_43.sample_43square          pb_diff_envs/environment/rand_rec_group_43.py:44-61
  RandRectangleGroup_43.save_center_and_mode     rand_rec_group_43.py:154-188
  RandRectangleGroup_43.compute_mode_list        rand_rec_group_43.py:190-211
  preproc_val_srm2dwloc_43                       diffuser/datasets/preprocessing.py:220-234

Parity: pinned against the live reference class -- tests/golden/concave_layouts.npz holds layouts drawn by the
reference's own sample_one_valid_env and the (cx, cy, mode) rows its save_center_and_mode returned for them
(tests/golden/gen_golden.py concave)."""
import numpy as np


def squares_of(center, hext, mode0):
    """rand_rec_group_43.py:44-61; mode0 in 0..3 is the index deleted from the quadrant list"""
    cl = [[center[0] - hext, center[1] + hext],
          [center[0] + hext, center[1] + hext],
          [center[0] + hext, center[1] - hext],
          [center[0] - hext, center[1] - hext]]
    del cl[mode0]
    return np.array(cl)


def centers_and_modes(layouts):
    """rand_rec_group_43.py:161-176, 190-211 -> [n_env, nw, 3] rows (cx, cy, mode in 1..4)"""
    out = []
    for wlocs in layouts:
        rows = []
        for i in range(0, len(wlocs), 3):
            x_max = max(wlocs[i][0], wlocs[i + 1][0])
            x_min = min(wlocs[i][0], wlocs[i + 2][0])
            x = (x_min + x_max) / 2
            y = (wlocs[i][1] + wlocs[i + 2][1]) / 2
            if wlocs[i][1] == wlocs[i + 1][1]:
                if wlocs[i][0] == wlocs[i + 2][0]:
                    mode = 2
                else:
                    assert wlocs[i + 1][0] == wlocs[i + 2][0]
                    mode = 3
            else:
                if wlocs[i][0] == wlocs[i + 2][0]:
                    mode = 1
                else:
                    assert wlocs[i][0] == wlocs[i + 1][0]
                    mode = 0
            rows.append([x, y, mode + 1])
        out.append(rows)
    return np.array(out, np.float64)


def val_wall_locations(centers, modes, n_probs):
    """preprocessing.py:220-234"""
    new_wloc = np.concatenate([centers, modes[..., None]], axis=-1)
    return new_wloc[:, None, :, :].repeat(n_probs, axis=1)
