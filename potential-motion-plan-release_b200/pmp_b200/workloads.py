"""Synthetic planning workloads for benchmarks and examples (product side).

Restates the reference's configuration files and layout samplers (citations inline) so that
`bench.py` needs nothing from the test-only `oracle/` package on its measured path.  A CPU
test (tests/test_abi_host.py::test_workloads_match_oracle) pins these generators to the
oracle's copies.
"""
import hashlib
import math

import numpy as np
import torch

FAMILIES = ("m2d", "m2d_nw3", "m2d_concave", "kuka", "dualkuka")


def _vit_cfg(n_tokens, token_dim, posemb):
    tf = {"mish": False}
    if posemb:
        tf["wall_sinPosEmb"] = {"num_octaves": 4, "norm_factor": 30}
    return dict(input_size=n_tokens * token_dim, patch_size=token_dim, num_classes=0,
                embed_dim=64, depth=3, mlp_ratio=4, pool="cls", attn_dim=64, num_heads=1,
                dim_head=64, channels=1, dropout=0.0, emb_dropout=0.0, tf_config=tf)


def arch(family):
    """dict(unet=kwargs for TemporalUnet_WCond, diffusion=kwargs for GaussianDiffusionPB,
    horizon=planning horizon, wall_dim=len of the wall vector fed to the model)"""
    if family == "m2d":
        nw, td, D, H, mults, extra, pe = 6, 2, 2, 48, (1, 4, 8), {}, True
    elif family == "m2d_nw3":
        nw, td, D, H, mults, extra, pe = 3, 2, 2, 48, (1, 4, 8), {}, True
    elif family == "m2d_concave":
        nw, td, D, H, mults, extra, pe = 7, 3, 2, 48, (1, 4, 8), {}, True
    elif family == "kuka":
        # config says num_walls=5 (input_size 15) but envs have 4 voxels; the ViT is
        # token-count agnostic (vit_vanilla.py:202-230), the planner feeds 4*3=12 values
        nw, td, D, H, mults, extra, pe = 5, 3, 7, 52, (1, 4, 8), {}, False
    elif family == "dualkuka":
        nw, td, D, H, mults, pe = 5, 3, 14, 56, (1, 4, 8, 12), False
        extra = dict(time_mlp_config=3, down_times=2)
    else:
        raise KeyError(family)
    net_cfg = dict(wallLoc_encoder="vit1d", vit1d_config=_vit_cfg(nw, td, pe),
                   concept_drop_prob=0.2, cat_t_w=True, energy_mode=True,
                   energy_param_type="L2", conv_zero_init=False, last_conv_ksize=1, **extra)
    unet = dict(horizon=48, transition_dim=D, cond_dim=D, dim=64, dim_mults=mults,
                num_walls=nw, wall_embed_dim=64, network_config=net_cfg)
    diffusion = dict(horizon=48, observation_dim=D, action_dim=0, n_timesteps=100,
                     loss_type="l2", clip_denoised=True, predict_epsilon=True,
                     loss_discount=1, loss_weights=None, condition_guidance_w=2.0,
                     diff_config=dict(set_cond_noise_to_0=False,
                                      manual_loss_weights={0: 0.0, -1: 0.0}))
    n_env_walls = {"m2d": 6, "m2d_nw3": 3, "m2d_concave": 7, "kuka": 4, "dualkuka": 5}[family]
    return dict(unet=unet, diffusion=diffusion, horizon=H, obs_dim=D,
                wall_dim=n_env_walls * td, token_dim=td, n_env_walls=n_env_walls)


# ----------------------------------------------------------------------------- weights

def _name_seed(name, seed):
    h = hashlib.sha256(("%d|%s" % (seed, name)).encode()).digest()
    return int.from_bytes(h[:8], "little") & ((1 << 62) - 1)


def synth_tensor(name, shape, seed, kind):
    """kind: 'w' fan-in-scaled uniform, 'b' small uniform, 'gain' 1+0.2*N, 'shift' 0.1*N, 'n' N(0,1)"""
    g = torch.Generator().manual_seed(_name_seed(name, seed))
    shape = tuple(shape)
    if kind == "w":
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
        bound = 1.0 / math.sqrt(max(fan_in, 1))
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
    if kind == "b":
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * 0.1
    if kind == "gain":
        return 1.0 + 0.2 * torch.randn(shape, generator=g, dtype=torch.float32)
    if kind == "shift":
        return 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
    if kind == "n":
        return torch.randn(shape, generator=g, dtype=torch.float32)
    raise KeyError(kind)


def _kind_of(name, shape):
    last = name.rsplit(".", 1)[-1]
    if "cls_token" in name:
        return "n"
    if len(shape) == 1:
        # 1-D tensors: GroupNorm / LayerNorm affine or a bias
        is_norm = (".block.2." in name) or (".norm." in name) or _is_patch_ln(name)
        if is_norm:
            return "gain" if last == "weight" else "shift"
        return "b"
    return "w"


_PATCH_LN = {}


def _is_patch_ln(name):
    return _PATCH_LN.get(name, False)


def synth_state_dict(template, seed=0):
    """template: mapping name -> tensor (only shapes are used; e.g. model.state_dict()).
    Returns a new state dict with name-keyed deterministic values.  LayerNorms inside
    `wallLoc_encoder.to_patch_embedding` are recognised by their position: a 1-D
    `weight` whose sibling `bias` has the same shape and for which no 2-D weight exists."""
    names = list(template.keys())
    two_d = {n.rsplit(".", 1)[0] for n in names if template[n].dim() >= 2}
    _PATCH_LN.clear()
    for n in names:
        mod = n.rsplit(".", 1)[0]
        if "to_patch_embedding" in n and template[n].dim() == 1 and mod not in two_d:
            _PATCH_LN[n] = True
    out = {}
    for n in names:
        t = template[n]
        if not torch.is_floating_point(t):
            out[n] = t.clone()
            continue
        out[n] = synth_tensor(n, t.shape, seed, _kind_of(n, tuple(t.shape))).to(t.dtype)
    return out


def state_dict_digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


# ----------------------------------------------------------------------------- geometry

def maze2d_walls(rng, nw, hext, maze=5.0, gap=0.15):
    """random non-overlapping square walls (rand_rec_group.py:73-147, 242-251); returns
    centres [nw,2] float64 sorted lexicographically.  Greedy placement restarts when stuck."""
    while True:
        acc, tries = [], 0
        while len(acc) < nw and tries < 200:
            tries += 1
            c = rng.uniform(hext + gap, maze - hext - gap, size=2)
            if any(_box_overlap(a, c, hext, gap) for a in acc):
                continue
            acc.append(c)
        if len(acc) == nw:
            break
    acc = np.array(acc, dtype=np.float64)
    order = np.lexsort((acc[:, 1], acc[:, 0]))
    return acc[order]


_CONCAVE_QUADS = np.array([[-1, 1], [1, 1], [1, -1], [-1, -1]], dtype=np.float64)


def _box_overlap(a, b, hext, gap):
    return not (a[0] + hext + gap < b[0] - hext or a[0] - hext - gap > b[0] + hext
                or a[1] + hext + gap < b[1] - hext or a[1] - hext - gap > b[1] + hext)


def maze2d_concave(rng, n_obj=7, hext=0.25, maze=5.0, gap=0.15):
    """concave obstacles (rand_rec_group_43.py:44-61,134-142): each object = 3 of the 4
    quadrant squares (order 1:(-,+) 2:(+,+) 3:(+,-) 4:(-,-)) around a centre, the deleted
    quadrant index is `mode`; overlap is tested per small square.  Restarts when stuck.
    Returns (tokens [n_obj,3] (cx,cy,mode+1) float64, boxes [3*n_obj,2] centres float64)."""
    while True:
        toks, boxes, tries = [], [], 0
        while len(toks) < n_obj and tries < 400:
            tries += 1
            c = rng.uniform(2 * hext + gap, maze - 2 * hext - gap, size=2)
            mode = int(rng.integers(0, 4))
            new = [c + hext * _CONCAVE_QUADS[q] for q in range(4) if q != mode]
            if any(_box_overlap(b, nb, hext, gap) for b in boxes for nb in new):
                continue
            toks.append(np.array([c[0], c[1], mode + 1.0]))
            boxes.extend(new)
        if len(toks) == n_obj:
            return np.array(toks), np.array(boxes)


def point_in_boxes(p, centers, hext, margin=0.01):
    lo = centers - hext - margin
    hi = centers + hext + margin
    return bool(np.any(np.all(p > lo, axis=1) & np.all(p < hi, axis=1)))


def maze2d_start_goal(rng, centers, hext, maze=5.0, min_l1=3.5):
    while True:
        s = rng.uniform(0, maze, size=2)
        g = rng.uniform(0, maze, size=2)
        if np.abs(s - g).sum() < min_l1:
            continue
        if point_in_boxes(s, centers, hext) or point_in_boxes(g, centers, hext):
            continue
        return s.astype(np.float32), g.astype(np.float32)


def kuka_voxels(rng, n=4, dual=False):
    """voxel centres [n,3] float64, sorted (rand_cuboid_group.py:153-202 restated with an
    AABB void around the robot base)"""
    acc = []
    hext = 0.22 if dual else 0.20
    while len(acc) < n:
        if dual:
            c = np.array([rng.uniform(-1, 1), rng.uniform(-0.5, 0.5), rng.uniform(0.3, 0.9)])
            if abs(abs(c[0]) - 0.5) < 0.42 and abs(c[1]) < 0.42:
                continue
        else:
            c = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), rng.uniform(0.3, 0.9)])
            if abs(c[0]) < 0.4 and abs(c[1]) < 0.4:
                continue
        if any(np.all(np.abs(c - a) < 2 * hext) for a in acc):
            continue
        acc.append(c)
    acc = np.array(acc)
    order = np.lexsort((acc[:, 2], acc[:, 1], acc[:, 0]))
    return acc[order]


KUKA_LIMIT = np.array([2.96705972839, 2.09439510239, 2.96705972839, 2.09439510239,
                       2.96705972839, 2.09439510239, 3.05432619099], dtype=np.float32)


def limits(family):
    if family.startswith("m2d"):
        return np.zeros(2, np.float32), np.full(2, 5.0, np.float32)
    if family == "kuka":
        return -KUKA_LIMIT, KUKA_LIMIT.copy()
    lim = np.concatenate([KUKA_LIMIT, KUKA_LIMIT])
    return -lim, lim


def make_problems(family, n, seed):
    """n independent planning problems.  Returns dict with float32 arrays:
    start [n,D], goal [n,D] (unnormalised), walls [n,wall_dim] (model input, metres),
    boxes [n,nb,dim] float64 box centres for the collision checker, hext float."""
    rng = np.random.default_rng(seed)
    a = arch(family)
    D = a["obs_dim"]
    lo, hi = limits(family)
    start = np.zeros((n, D), np.float32)
    goal = np.zeros((n, D), np.float32)
    walls = np.zeros((n, a["wall_dim"]), np.float32)
    boxes = []
    if family in ("m2d", "m2d_nw3"):
        nw, hext = (6, 0.5) if family == "m2d" else (3, 0.7)
        for i in range(n):
            c = maze2d_walls(rng, nw, hext)
            start[i], goal[i] = maze2d_start_goal(rng, c, hext)
            walls[i] = c.reshape(-1).astype(np.float32)
            boxes.append(c)
    elif family == "m2d_concave":
        hext = 0.25
        for i in range(n):
            toks, bx = maze2d_concave(rng)
            start[i], goal[i] = maze2d_start_goal(rng, bx, hext, min_l1=3.5)
            walls[i] = toks.reshape(-1).astype(np.float32)
            boxes.append(bx)
    else:
        dual = family == "dualkuka"
        hext = 0.22 if dual else 0.20
        nv = a["n_env_walls"]
        for i in range(n):
            c = kuka_voxels(rng, nv, dual)
            walls[i] = c.reshape(-1).astype(np.float32)
            boxes.append(c)
            start[i] = rng.uniform(lo, hi).astype(np.float32)
            goal[i] = rng.uniform(lo, hi).astype(np.float32)
    return dict(start=start, goal=goal, walls=walls, boxes=np.array(boxes, np.float64), hext=hext,
                lo=lo, hi=hi)


def normalize(x, lo, hi):
    """LimitsNormalizer.normalize (normalization.py:142-149), float32"""
    x = (x - lo) / (hi - lo)
    return (2 * x - 1).astype(np.float32)


def unnormalize(x, lo, hi):
    """LimitsNormalizer.unnormalize (normalization.py:151-162): clip, (x+1)/2, scale; three
    separately rounded float32 ops"""
    x = np.asarray(x, np.float32)
    if x.max() > 1 or x.min() < -1:
        x = np.clip(x, -1, 1)
    x = (x + np.float32(1)) / np.float32(2.)
    return x * (hi - lo) + lo


def noise(T, B, H, D, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(T + 1, B, H, D, generator=g, dtype=torch.float32)
