"""Functional CPU restatement of the energy U-Net and of its analytic dE/dx.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows (reference file:line):
  * TemporalUnet_WCond.forward            diffuser/models/temporal_cond.py:245-334
  * ResidualTemporalBlock_dd.forward      diffuser/models/temporal_dd.py:56-74
  * Conv1dBlock_dd                        diffuser/models/helpers.py:74-97
  * Downsample1d / Upsample1d             diffuser/models/helpers.py:38-52
  * SinusoidalPosEmb                      diffuser/models/helpers.py:16-36
  * ViT1D / Transformer / Attention       diffuser/models/vit_vanilla.py:12-230
  * PositionalEncoding2D                  diffuser/models/helpers.py:276-328

It takes a plain state dict (reference key names) instead of nn.Modules, so it can
travel to the GPU box where /root/reference does not exist.  `eps_autograd` uses
torch.autograd exactly like the reference; `eps_analytic` is the hand-derived VJP
 that the CUDA backward implements, kept here so the formulas
are checked against autograd on CPU (in fp64 too).
"""
import math

import torch
import torch.nn.functional as F


class UnetCfg:
    def __init__(self, unet_kwargs):
        k = unet_kwargs
        nc = k["network_config"]
        self.D = k["transition_dim"]
        self.dim = k.get("dim", 32)
        self.mults = tuple(k.get("dim_mults", (1, 2, 4, 8)))
        self.dims = [self.D] + [self.dim * m for m in self.mults]
        self.down_times = nc.get("down_times", 10 ** 5)
        self.time_mlp_config = nc.get("time_mlp_config", False)
        self.energy_mode = nc.get("energy_mode", False)
        self.mish = not self.energy_mode
        self.wemb = k.get("wall_embed_dim", 32)
        v = nc["vit1d_config"]
        self.vit_dim = v["embed_dim"]
        self.vit_depth = v["depth"]
        self.token_dim = v["patch_size"]
        pe = v["tf_config"].get("wall_sinPosEmb", False)
        self.vit_octaves = pe["num_octaves"] if pe else 0
        self.vit_norm = pe["norm_factor"] if pe else 30
        self.vit_mish = v["tf_config"].get("mish", True)
        self.n_levels = len(self.mults)

    def level_has_down(self, ind):
        return not (ind >= self.n_levels - 1 or ind >= self.down_times)

    def up_has_up(self, ind):
        return not (ind >= self.n_levels - 1 or ind < (self.n_levels - self.down_times - 1))


def _act(x, mish):
    return F.mish(x) if mish else F.silu(x)


def _dact(z, mish):
    if mish:
        sp = F.softplus(z)
        th = torch.tanh(sp)
        return th + z * (1 - th * th) * torch.sigmoid(z)
    s = torch.sigmoid(z)
    return s * (1 + z * (1 - s))


def sinusoidal(t, dim):
    half = dim // 2
    f = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
    e = t[:, None].to(f.dtype) * f[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def wall_embed(sd, cfg, walls, prefix="wallLoc_encoder."):
    """ViT1D cls vector [B, 64]"""
    dt = walls.dtype
    B = walls.shape[0]
    x = walls.reshape(B, -1, cfg.token_dim)
    i = 1
    if cfg.vit_octaves:
        c = (x / cfg.vit_norm).unsqueeze(-1)
        mult = (2 ** torch.arange(cfg.vit_octaves).float() * math.pi).to(dt)
        sc = c * mult
        n = x.shape[1]
        x = torch.cat((torch.sin(sc).reshape(B, n, -1), torch.cos(sc).reshape(B, n, -1)), -1)
        x = F.layer_norm(x, (x.shape[-1],), sd[prefix + "to_patch_embedding.2.weight"],
                         sd[prefix + "to_patch_embedding.2.bias"])
        i = 3
    p = prefix + "to_patch_embedding."
    x = F.linear(x, sd[p + "%d.weight" % i], sd[p + "%d.bias" % i])
    x = F.linear(x, sd[p + "%d.weight" % (i + 1)], sd[p + "%d.bias" % (i + 1)])
    x = F.mish(x) if cfg.vit_mish else F.silu(x)
    x = F.linear(x, sd[p + "%d.weight" % (i + 3)], sd[p + "%d.bias" % (i + 3)])
    x = F.layer_norm(x, (cfg.vit_dim,), sd[p + "%d.weight" % (i + 4)], sd[p + "%d.bias" % (i + 4)])
    cls = sd[prefix + "cls_token"].expand(B, 1, cfg.vit_dim)
    x = torch.cat((cls, x), dim=1)
    scale = cfg.vit_dim ** -0.5
    for l in range(cfg.vit_depth):
        q = prefix + "transformer.layers.%d." % l
        y = F.layer_norm(x, (cfg.vit_dim,), sd[q + "0.norm.weight"], sd[q + "0.norm.bias"])
        qkv = F.linear(y, sd[q + "0.fn.to_qkv.weight"])
        qq, kk, vv = qkv.chunk(3, dim=-1)
        att = torch.softmax(torch.matmul(qq, kk.transpose(-1, -2)) * scale, dim=-1)
        x = torch.matmul(att, vv) + x
        y = F.layer_norm(x, (cfg.vit_dim,), sd[q + "1.norm.weight"], sd[q + "1.norm.bias"])
        y = F.linear(y, sd[q + "1.fn.net.0.weight"], sd[q + "1.fn.net.0.bias"])
        y = F.gelu(y) if cfg.vit_mish else F.silu(y)
        y = F.linear(y, sd[q + "1.fn.net.3.weight"], sd[q + "1.fn.net.3.bias"])
        x = y + x
    return x[:, 0, :]


def time_embed(sd, cfg, t):
    e = sinusoidal(t, cfg.dim).to(sd["time_mlp.1.weight"].dtype)
    e = F.linear(e, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"])
    e = _act(e, cfg.mish)
    return F.linear(e, sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])


def block_names(cfg):
    """residual blocks in forward order with (prefix, cin, cout)"""
    io = list(zip(cfg.dims[:-1], cfg.dims[1:]))
    out = []
    for ind, (ci, co) in enumerate(io):
        out.append(("downs.%d.0." % ind, ci, co))
        out.append(("downs.%d.1." % ind, co, co))
    mid = cfg.dims[-1]
    out.append(("mid_block1.", mid, mid))
    out.append(("mid_block2.", mid, mid))
    for ind, (ci, co) in enumerate(reversed(io[1:])):
        out.append(("ups.%d.0." % ind, co * 2, ci))
        out.append(("ups.%d.1." % ind, ci, ci))
    return out


def block_embed(sd, cfg, pfx, tcat):
    """ResidualTemporalBlock_dd.time_mlp (temporal_dd.py:31-49) -> [R, Cout]"""
    if cfg.time_mlp_config == 3:
        h = F.linear(tcat, sd[pfx + "time_mlp.0.weight"], sd[pfx + "time_mlp.0.bias"])
        h = _act(h, cfg.mish)
        return F.linear(h, sd[pfx + "time_mlp.2.weight"], sd[pfx + "time_mlp.2.bias"])
    if cfg.time_mlp_config == 2:
        h = F.linear(_act(tcat, cfg.mish), sd[pfx + "time_mlp.1.weight"], sd[pfx + "time_mlp.1.bias"])
        return F.linear(_act(h, cfg.mish), sd[pfx + "time_mlp.3.weight"], sd[pfx + "time_mlp.3.bias"])
    return F.linear(_act(tcat, cfg.mish), sd[pfx + "time_mlp.1.weight"], sd[pfx + "time_mlp.1.bias"])


def _cb(sd, cfg, pfx, x):
    y = F.conv1d(x, sd[pfx + "block.0.weight"], sd[pfx + "block.0.bias"], padding=2)
    z = F.group_norm(y, 8, sd[pfx + "block.2.weight"], sd[pfx + "block.2.bias"], eps=1e-5)
    return _act(z, cfg.mish)


def _res(sd, cfg, pfx, x, tcat):
    out = _cb(sd, cfg, pfx + "blocks.0.", x) + block_embed(sd, cfg, pfx, tcat)[:, :, None]
    out = _cb(sd, cfg, pfx + "blocks.1.", out)
    if (pfx + "residual_conv.weight") in sd:
        return out + F.conv1d(x, sd[pfx + "residual_conv.weight"], sd[pfx + "residual_conv.bias"])
    return out + x


def unet_u(sd, cfg, x_bhd, t, wemb):
    """U-Net output u [R,H,D] given the (already masked) wall embedding wemb [R,64]"""
    x = x_bhd.transpose(1, 2)
    tcat = torch.cat([time_embed(sd, cfg, t), wemb], dim=-1)
    h = []
    for ind in range(cfg.n_levels):
        x = _res(sd, cfg, "downs.%d.0." % ind, x, tcat)
        x = _res(sd, cfg, "downs.%d.1." % ind, x, tcat)
        h.append(x)
        if cfg.level_has_down(ind):
            x = F.conv1d(x, sd["downs.%d.2.conv.weight" % ind], sd["downs.%d.2.conv.bias" % ind],
                         stride=2, padding=1)
    x = _res(sd, cfg, "mid_block1.", x, tcat)
    x = _res(sd, cfg, "mid_block2.", x, tcat)
    for ind in range(cfg.n_levels - 1):
        x = torch.cat((x, h.pop()), dim=1)
        x = _res(sd, cfg, "ups.%d.0." % ind, x, tcat)
        x = _res(sd, cfg, "ups.%d.1." % ind, x, tcat)
        if cfg.up_has_up(ind):
            x = F.conv_transpose1d(x, sd["ups.%d.2.conv.weight" % ind], sd["ups.%d.2.conv.bias" % ind],
                                   stride=2, padding=1)
    x = _cb(sd, cfg, "final_conv.0.", x)
    x = F.conv1d(x, sd["final_conv.1.weight"], sd["final_conv.1.bias"])
    return x.transpose(1, 2)


def eps_autograd(sd, cfg, x, t, walls, n_cond=None, return_u=False):
    """eps = d(0.5*sum u^2)/dx with rows >= n_cond unconditional (wall embedding zeroed,
    temporal_cond.py:272-279).  x [R,H,D], t [R] int64, walls [R,Wd]."""
    R = x.shape[0]
    n_cond = R if n_cond is None else n_cond
    with torch.enable_grad():
        xi = x.detach().clone().requires_grad_(True)
        w = wall_embed(sd, cfg, walls)
        mask = (torch.arange(R) < n_cond).to(w.dtype)[:, None]
        w = w * mask
        u = unet_u(sd, cfg, xi, t, w)
        e = (0.5 * u * u).sum(dim=(1, 2))
        g = torch.autograd.grad([e.sum()], [xi])[0]
    if return_u:
        return g.detach(), u.detach(), e.detach()
    return g.detach()


def eps_pair(sd, cfg, x, t, walls):
    """the classifier-free-guidance pair of diffusion_pb.py:168-175 -> (eps_cond, eps_uncond)"""
    B = x.shape[0]
    g = eps_autograd(sd, cfg, x.repeat(2, 1, 1), t.repeat(2), walls.repeat(2, 1), n_cond=B)
    return g[:B], g[B:]


# ----------------------------------------------------------------------------- analytic VJP

def _gn_fwd(y, gamma, beta):
    R, C, H = y.shape
    yg = y.reshape(R, 8, -1)
    mu = yg.mean(-1, keepdim=True)
    var = yg.var(-1, unbiased=False, keepdim=True)
    r = torch.rsqrt(var + 1e-5)
    yh = ((yg - mu) * r).reshape(R, C, H)
    return yh, r


def _cb_fwd(sd, cfg, pfx, x, tape):
    y = F.conv1d(x, sd[pfx + "block.0.weight"], sd[pfx + "block.0.bias"], padding=2)
    yh, r = _gn_fwd(y, None, None)
    z = yh * sd[pfx + "block.2.weight"][None, :, None] + sd[pfx + "block.2.bias"][None, :, None]
    tape[pfx] = (yh, r)
    return _act(z, cfg.mish)


def _cb_bwd(sd, cfg, pfx, g_a, tape):
    """Appendix G item 3: gradient w.r.t. the conv *input*"""
    yh, r = tape[pfx]
    gam = sd[pfx + "block.2.weight"][None, :, None]
    z = yh * gam + sd[pfx + "block.2.bias"][None, :, None]
    g_yh = g_a * _dact(z, cfg.mish) * gam
    R, C, H = g_yh.shape
    a = g_yh.reshape(R, 8, -1)
    b = yh.reshape(R, 8, -1)
    g_y = (r * (a - a.mean(-1, keepdim=True) - b * (a * b).mean(-1, keepdim=True))).reshape(R, C, H)
    W = sd[pfx + "block.0.weight"]
    return F.conv_transpose1d(g_y, W, padding=2)


def _res_fwd(sd, cfg, pfx, x, tcat, tape):
    m = _cb_fwd(sd, cfg, pfx + "blocks.0.", x, tape) + block_embed(sd, cfg, pfx, tcat)[:, :, None]
    o = _cb_fwd(sd, cfg, pfx + "blocks.1.", m, tape)
    if (pfx + "residual_conv.weight") in sd:
        return o + F.conv1d(x, sd[pfx + "residual_conv.weight"], sd[pfx + "residual_conv.bias"])
    return o + x


def _res_bwd(sd, cfg, pfx, g_out, tape):
    g_m = _cb_bwd(sd, cfg, pfx + "blocks.1.", g_out, tape)
    g_x = _cb_bwd(sd, cfg, pfx + "blocks.0.", g_m, tape)
    if (pfx + "residual_conv.weight") in sd:
        return g_x + F.conv_transpose1d(g_out, sd[pfx + "residual_conv.weight"])
    return g_x + g_out


def eps_analytic(sd, cfg, x, t, walls, n_cond=None, return_tape=False):
    """same result as eps_autograd, via the hand-written VJP (Appendix G items 1-8)"""
    R = x.shape[0]
    n_cond = R if n_cond is None else n_cond
    with torch.no_grad():
        w = wall_embed(sd, cfg, walls)
        w = w * (torch.arange(R) < n_cond).to(w.dtype)[:, None]
        tcat = torch.cat([time_embed(sd, cfg, t), w], dim=-1)
        tape = {}
        a = x.transpose(1, 2)
        L = cfg.n_levels
        for ind in range(L):
            a = _res_fwd(sd, cfg, "downs.%d.0." % ind, a, tcat, tape)
            a = _res_fwd(sd, cfg, "downs.%d.1." % ind, a, tcat, tape)
            tape["skip%d" % ind] = a
            if cfg.level_has_down(ind):
                a = F.conv1d(a, sd["downs.%d.2.conv.weight" % ind], sd["downs.%d.2.conv.bias" % ind],
                             stride=2, padding=1)
        a = _res_fwd(sd, cfg, "mid_block1.", a, tcat, tape)
        a = _res_fwd(sd, cfg, "mid_block2.", a, tcat, tape)
        for ind in range(L - 1):
            a = torch.cat((a, tape["skip%d" % (L - 1 - ind)]), dim=1)
            a = _res_fwd(sd, cfg, "ups.%d.0." % ind, a, tcat, tape)
            a = _res_fwd(sd, cfg, "ups.%d.1." % ind, a, tcat, tape)
            if cfg.up_has_up(ind):
                a = F.conv_transpose1d(a, sd["ups.%d.2.conv.weight" % ind], sd["ups.%d.2.conv.bias" % ind],
                                       stride=2, padding=1)
        f = _cb_fwd(sd, cfg, "final_conv.0.", a, tape)
        u = F.conv1d(f, sd["final_conv.1.weight"], sd["final_conv.1.bias"])
        # ---- backward, seed g_u = u
        g = F.conv_transpose1d(u, sd["final_conv.1.weight"])
        g = _cb_bwd(sd, cfg, "final_conv.0.", g, tape)
        skip_g = {}
        for ind in reversed(range(L - 1)):
            if cfg.up_has_up(ind):
                # ConvTranspose1d backward = strided conv with the same weight
                g = F.conv1d(g, sd["ups.%d.2.conv.weight" % ind], stride=2, padding=1)
            g = _res_bwd(sd, cfg, "ups.%d.1." % ind, g, tape)
            g = _res_bwd(sd, cfg, "ups.%d.0." % ind, g, tape)
            C = g.shape[1] // 2
            skip_g[L - 1 - ind] = g[:, C:]
            g = g[:, :C]
        g = _res_bwd(sd, cfg, "mid_block2.", g, tape)
        g = _res_bwd(sd, cfg, "mid_block1.", g, tape)
        for ind in reversed(range(L)):
            if cfg.level_has_down(ind):
                g = F.conv_transpose1d(g, sd["downs.%d.2.conv.weight" % ind], stride=2, padding=1,
                                       output_padding=1)
            if ind in skip_g:
                g = g + skip_g[ind]
            g = _res_bwd(sd, cfg, "downs.%d.1." % ind, g, tape)
            g = _res_bwd(sd, cfg, "downs.%d.0." % ind, g, tape)
        if return_tape:
            return g.transpose(1, 2).contiguous(), u.transpose(1, 2).contiguous(), tape
        return g.transpose(1, 2).contiguous(), u.transpose(1, 2).contiguous()


# ----------------------------------------------------------------------------- parameter inventory

def param_shapes(unet_kwargs):
    """name -> shape of every parameter of TemporalUnet_WCond(**unet_kwargs), in the reference's
    state_dict order (checked against the live reference in tests)."""
    cfg = UnetCfg(unet_kwargs)
    nc = unet_kwargs["network_config"]
    dim, we = cfg.dim, cfg.wemb
    tdim = dim + we
    out = {}

    def lin(name, o, i, bias=True):
        out[name + ".weight"] = (o, i)
        if bias:
            out[name + ".bias"] = (o,)

    def norm(name, n):
        out[name + ".weight"] = (n,)
        out[name + ".bias"] = (n,)

    lin("time_mlp.1", dim * 4, dim)
    lin("time_mlp.3", dim, dim * 4)
    v = nc["vit1d_config"]
    vd, td = v["embed_dim"], v["patch_size"] * v.get("channels", 1)
    p = "wallLoc_encoder."
    out[p + "cls_token"] = (1, 1, vd)
    i = 1
    fdim = td
    if cfg.vit_octaves:
        fdim = td * cfg.vit_octaves * 2
        norm(p + "to_patch_embedding.2", fdim)
        i = 3
    lin(p + "to_patch_embedding.%d" % i, vd, fdim)
    lin(p + "to_patch_embedding.%d" % (i + 1), vd * 4, vd)
    lin(p + "to_patch_embedding.%d" % (i + 3), vd, vd * 4)
    norm(p + "to_patch_embedding.%d" % (i + 4), vd)
    mlp = vd * v["mlp_ratio"]
    for l in range(v["depth"]):
        q = p + "transformer.layers.%d." % l
        norm(q + "0.norm", vd)
        lin(q + "0.fn.to_qkv", 3 * v["num_heads"] * v["dim_head"], vd, bias=False)
        norm(q + "1.norm", vd)
        lin(q + "1.fn.net.0", mlp, vd)
        lin(q + "1.fn.net.3", vd, mlp)

    def cb(name, ci, co):
        out[name + "block.0.weight"] = (co, ci, 5)
        out[name + "block.0.bias"] = (co,)
        norm(name + "block.2", co)

    def res(name, ci, co):
        cb(name + "blocks.0.", ci, co)
        cb(name + "blocks.1.", co, co)
        if cfg.time_mlp_config == 3:
            lin(name + "time_mlp.0", tdim * 2, tdim)
            lin(name + "time_mlp.2", co, tdim * 2)
        elif cfg.time_mlp_config == 2:
            lin(name + "time_mlp.1", co * 2, tdim)
            lin(name + "time_mlp.3", co, co * 2)
        else:
            lin(name + "time_mlp.1", co, tdim)
        if ci != co:
            out[name + "residual_conv.weight"] = (co, ci, 1)
            out[name + "residual_conv.bias"] = (co,)

    io = list(zip(cfg.dims[:-1], cfg.dims[1:]))
    for ind, (ci, co) in enumerate(io):
        res("downs.%d.0." % ind, ci, co)
        res("downs.%d.1." % ind, co, co)
        if cfg.level_has_down(ind):
            out["downs.%d.2.conv.weight" % ind] = (co, co, 3)
            out["downs.%d.2.conv.bias" % ind] = (co,)
    ups = {}
    for ind, (ci, co) in enumerate(reversed(io[1:])):
        res("ups.%d.0." % ind, co * 2, ci)
        res("ups.%d.1." % ind, ci, ci)
        if cfg.up_has_up(ind):
            out["ups.%d.2.conv.weight" % ind] = (ci, ci, 4)
            out["ups.%d.2.conv.bias" % ind] = (ci,)
    mid = cfg.dims[-1]
    res("mid_block1.", mid, mid)
    res("mid_block2.", mid, mid)
    cb("final_conv.0.", dim, dim)
    out["final_conv.1.weight"] = (cfg.D, dim, 1)
    out["final_conv.1.bias"] = (cfg.D,)
    return out


def synth_weights(unet_kwargs, seed):
    """deterministic name-keyed weights for an architecture (no reference needed)"""
    from . import synth
    shapes = param_shapes(unet_kwargs)
    tmpl = {k: torch.empty(s) for k, s in shapes.items()}
    return synth.synth_state_dict(tmpl, seed)
