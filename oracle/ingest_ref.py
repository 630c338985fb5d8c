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


# compute_mode_list's decision tree as a table: (first two squares share y, first and third share x) -> mode0
_MODE0 = {(True, True): 2, (True, False): 3, (False, True): 1, (False, False): 0}


def centers_and_modes(layouts):
    """rand_rec_group_43.py:161-176, 190-211 -> [n_env, nw, 3] rows (cx, cy, mode in 1..4), one obstacle (three
    consecutive squares a, b, c) at a time with scalar float64 arithmetic"""
    out = []
    for squares in layouts:
        rows = []
        for a, b, c in zip(squares[0::3], squares[1::3], squares[2::3]):
            cx = (min(a[0], c[0]) + max(a[0], b[0])) / 2
            cy = (a[1] + c[1]) / 2
            same_y, same_x = bool(a[1] == b[1]), bool(a[0] == c[0])
            if not same_x:          # the reference asserts the remaining coincidence
                assert (b[0] == c[0]) if same_y else (a[0] == b[0])
            rows.append([cx, cy, _MODE0[(same_y, same_x)] + 1])
        out.append(rows)
    return np.array(out, np.float64)


def val_wall_locations(centers, modes, n_probs):
    """preprocessing.py:220-234"""
    new_wloc = np.concatenate([centers, modes[..., None]], axis=-1)
    return new_wloc[:, None, :, :].repeat(n_probs, axis=1)
