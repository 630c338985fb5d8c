"""ctypes binding of include/pmp_b200.h.  There is no CPU fallback: every compute entry point
raises PmpError when the shared library is missing or no sm_100 GPU is usable."""
import ctypes as C
import os

import numpy as np
import torch

LIB_PATH = os.environ.get("PMP_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpmp_b200.so")   # override: A/B timing of two builds
MAX_LEVELS = 8
MAX_COMPOSE = 8


class PmpError(RuntimeError):
    pass


class Arch(C.Structure):
    _fields_ = [("transition_dim", C.c_int32), ("dim", C.c_int32), ("n_mults", C.c_int32),
                ("dim_mults", C.c_int32 * MAX_LEVELS), ("down_times", C.c_int32),
                ("time_mlp_config", C.c_int32), ("act", C.c_int32), ("wall_embed_dim", C.c_int32),
                ("vit_token_dim", C.c_int32), ("vit_octaves", C.c_int32), ("vit_norm_factor", C.c_float),
                ("vit_depth", C.c_int32), ("vit_mlp_ratio", C.c_int32), ("vit_act", C.c_int32),
                ("n_timesteps", C.c_int32), ("reserved", C.c_int32 * 8)]


class SampleArgs(C.Structure):
    _fields_ = [("K", C.c_int32), ("models", C.c_void_p * MAX_COMPOSE), ("wemb_dev", C.c_void_p * MAX_COMPOSE),
                ("w", C.c_float * MAX_COMPOSE), ("base", C.c_int32), ("B", C.c_int32), ("H", C.c_int32),
                ("D", C.c_int32), ("kind", C.c_int32), ("n_steps", C.c_int32), ("t_host", C.c_void_p),
                ("coef_host", C.c_void_p), ("noise_dev", C.c_void_p), ("seed", C.c_uint64),
                ("row_offset", C.c_uint64), ("var_temp", C.c_float), ("start_dev", C.c_void_p),
                ("goal_dev", C.c_void_p), ("x_dev", C.c_void_p), ("trace_dev", C.c_void_p)]


_lib = None

_PROTOS = {
    "pmp_last_error": (C.c_char_p, []),
    "pmp_abi_version": (C.c_int, []),
    "pmp_device_count": (C.c_int, []),
    "pmp_model_create": (C.c_int, [C.POINTER(Arch), C.c_int, C.POINTER(C.c_void_p)]),
    "pmp_model_destroy": (None, [C.c_void_p]),
    "pmp_model_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_int]),
    "pmp_model_finalize": (C.c_int, [C.c_void_p]),
    "pmp_model_set_row_chunk": (C.c_int, [C.c_void_p, C.c_int]),
    "pmp_model_set_gemm_path": (C.c_int, [C.c_void_p, C.c_int]),
    "pmp_model_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "pmp_wall_embed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "pmp_unet_eps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pmp_step": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_float),
                           C.c_int, C.c_int, C.POINTER(C.c_float), C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p,
                           C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "pmp_init_noise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_float, C.c_void_p, C.c_void_p,
                                 C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "pmp_sample": (C.c_int, [C.POINTER(SampleArgs), C.c_void_p]),
    "pmp_unnormalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "pmp_colchk_maze2d_swept": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_float, C.c_float, C.c_float, C.c_float, C.c_double, C.c_double, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "pmp_colchk_maze2d_points": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_float, C.c_float, C.c_float, C.c_float, C.c_double, C.c_void_p,
                                           C.c_void_p, C.c_void_p]),
    "pmp_colchk_kuka": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                  C.POINTER(C.c_float), C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pmp_pick_first_valid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "pmp_train_bind": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.c_void_p, C.c_void_p,
                                 C.c_int64]),
    "pmp_train_sync_weights": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "pmp_train_loss_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pmp_train_set_buckets": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]),
    "pmp_train_wait_bucket": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "pmp_q_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_void_p, C.c_void_p]),
    "pmp_adam_ema_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float,
                                    C.c_float, C.c_float, C.c_int, C.c_float, C.c_int, C.c_float, C.c_void_p]),
    "pmp_launch_count": (C.c_uint64, []),
    "pmp_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "pmp_profile_enable": (C.c_int, [C.c_int]),
    "pmp_profile_read": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "pmp_debug_counters": (C.c_int, [C.POINTER(C.c_double), C.c_int]),
    "pmp_debug_wgrad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "pmp_debug_wgrad_scratch_floats": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "pmp_debug_fetch": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
}

EXPORTS = tuple(sorted(_PROTOS))


def lib():
    """loads libpmp_b200.so (built by __graft_entry__.build() / csrc/Makefile); raises if absent"""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PmpError("libpmp_b200.so is not built (run `python __graft_entry__.py build`); "
                           "there is no CPU fallback for the planning hot path")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            f = getattr(l, name)
            f.restype = res
            f.argtypes = args
        if l.pmp_abi_version() != 1:
            raise PmpError("libpmp_b200.so ABI version mismatch")
        _lib = l
    return _lib


def _check(rc):
    if rc != 0:
        raise PmpError("pmp_b200 error %d: %s" % (rc, lib().pmp_last_error().decode(errors="replace")))


def _stream(dev=None):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def device_count():
    return lib().pmp_device_count()


def launch_count():
    return int(lib().pmp_launch_count())


def set_option(name, value):
    """library-wide options: 'uncond_dedup' (default 1) -- same-network compositions share the unconditional half"""
    _check(lib().pmp_set_option(name.encode(), int(value)))


def profile_enable(on=True):
    _check(lib().pmp_profile_enable(1 if on else 0))


def profile_read(n_classes=2):
    """[(ms, algorithmic flops, launches)] per conv-kernel class since the last read"""
    ms = (C.c_double * n_classes)()
    fl = (C.c_double * n_classes)()
    ln = (C.c_int64 * n_classes)()
    _check(lib().pmp_profile_read(n_classes, ms, fl, ln))
    return [(ms[i], fl[i], int(ln[i])) for i in range(n_classes)]


class UnetEngine:
    """Owns one pmp_model: the energy U-Net (+ViT wall encoder) of TemporalUnet_WCond with a
    hand-written backward for eps = dE/dx (replaces temporal_cond.py:245-334)."""

    def __init__(self, arch_kwargs, state_dict, device, n_timesteps=100, prefix=""):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise PmpError("the planning hot path runs on a B200 only (got device %s); no CPU fallback" % dev)
        if dev.index is None:    # torch.device('cuda') means the CURRENT device, not device 0
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.arch = make_arch(arch_kwargs, n_timesteps)
        self.h = C.c_void_p()
        _check(lib().pmp_model_create(C.byref(self.arch), dev.index, C.byref(self.h)))
        self.load(state_dict, prefix)

    def load(self, state_dict, prefix=""):
        for k, v in state_dict.items():
            if prefix and not k.startswith(prefix):
                continue
            if not torch.is_floating_point(v):
                continue
            t = v.detach().to("cpu", torch.float32).contiguous()
            shape = (C.c_int64 * max(t.dim(), 1))(*(t.shape if t.dim() else (1,)))
            _check(lib().pmp_model_set_weight(self.h, k[len(prefix):].encode(), C.c_void_p(t.data_ptr()), shape,
                                              max(t.dim(), 1), 0))
        _check(lib().pmp_model_finalize(self.h))

    def set_row_chunk(self, rows):
        _check(lib().pmp_model_set_row_chunk(self.h, int(rows)))

    def set_gemm_path(self, path):
        _check(lib().pmp_model_set_gemm_path(self.h, int(path)))

    def wall_embed(self, walls):
        walls = _f32(walls, self.device)
        out = torch.empty(walls.shape[0], self.arch.wall_embed_dim, device=self.device, dtype=torch.float32)
        _check(lib().pmp_wall_embed(self.h, _ptr(walls), walls.shape[0], walls.shape[1], _ptr(out), _stream(self.device)))
        return out

    def eps(self, x, t, wemb, n_cond=None, want_energy=False, want_u=False):
        """x [R,H,D], t [R] long, wemb [n_cond,64] -> eps [R,H,D]; rows >= n_cond unconditional"""
        x = _f32(x, self.device)
        R, H, D = x.shape
        t = t.detach().to(torch.int64)
        # the time-embedding table has arch.n_timesteps rows: an index outside it is an error, not something to clamp
        # (one small host read; the whole-loop sampler pmp_sample validates its host-side timestep list instead)
        if t.numel() and (int(t.min()) < 0 or int(t.max()) >= self.arch.n_timesteps):
            raise PmpError("timestep out of range: the engine was built for n_timesteps=%d, got t in [%d, %d]"
                           % (self.arch.n_timesteps, int(t.min()), int(t.max())))
        t = t.to(self.device).contiguous()
        n_cond = R if n_cond is None else int(n_cond)
        wemb = _f32(wemb, self.device) if wemb is not None else None
        eps = torch.empty_like(x)
        en = torch.empty(R, device=self.device, dtype=torch.float32) if want_energy else None
        u = torch.empty_like(x) if want_u else None
        _check(lib().pmp_unet_eps(self.h, _ptr(x), _ptr(t), _ptr(wemb), n_cond, R, H, _ptr(eps), _ptr(en), _ptr(u),
                                  _stream(self.device)))
        out = (eps,)
        if want_energy:
            out += (en,)
        if want_u:
            out += (u,)
        return out[0] if len(out) == 1 else out

    def debug_fetch(self, name):
        n = lib().pmp_debug_fetch(self.h, name.encode(), None, 0)
        if n < 0:
            return None
        buf = np.empty(n, np.float32)
        got = lib().pmp_debug_fetch(self.h, name.encode(), buf.ctypes.data_as(C.c_void_p), n)
        if got != n:
            raise PmpError("debug fetch failed (%d)" % got)
        return buf

    def __del__(self):
        try:
            if self.h:
                lib().pmp_model_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass


def make_arch(k, n_timesteps=100):
    """constructor kwargs of TemporalUnet_WCond (temporal_cond.py:67-78) -> pmp_arch"""
    nc = k.get("network_config", {})
    if nc.get("wallLoc_encoder", "mlp").lower() != "vit1d":
        raise PmpError("only the ViT1D wall encoder is built (wallLoc_encoder='vit1d')")
    if not nc.get("cat_t_w", False):
        raise PmpError("only cat_t_w=True (ResidualTemporalBlock_dd) is built")
    if not nc.get("energy_mode", False):
        raise PmpError("only energy_mode=True (potential-based) models are built")
    a = Arch()
    a.transition_dim = int(k["transition_dim"])
    a.dim = int(k.get("dim", 32))
    mults = tuple(k.get("dim_mults", (1, 2, 4, 8)))
    a.n_mults = len(mults)
    for i, m in enumerate(mults):
        a.dim_mults[i] = int(m)
    dt = nc.get("down_times", 1 << 20)
    a.down_times = int(min(dt, 1 << 20))
    a.time_mlp_config = int(nc.get("time_mlp_config", 0) or 0)
    a.act = 0
    a.wall_embed_dim = int(k.get("wall_embed_dim", 32))
    v = nc["vit1d_config"]
    tf = v.get("tf_config", {})
    pe = tf.get("wall_sinPosEmb", False)
    a.vit_token_dim = int(v["patch_size"]) * int(v.get("channels", 1))
    a.vit_octaves = int(pe["num_octaves"]) if pe else 0
    a.vit_norm_factor = float(pe["norm_factor"]) if pe else 30.0
    a.vit_depth = int(v["depth"])
    a.vit_mlp_ratio = int(v["mlp_ratio"])
    a.vit_act = 1 if tf.get("mish", True) else 0
    a.n_timesteps = int(n_timesteps)
    if int(v["embed_dim"]) != 64 or int(v["num_heads"]) != 1 or int(v["dim_head"]) != 64 or tf.get("dyn_posemb"):
        raise PmpError("wall encoder: only embed_dim=64, one head of 64, no dyn_posemb is built")
    return a


def init_noise(x, noise, seed, start, goal, var_temp=1.0, offset=0):
    B, H, D = x.shape
    _check(lib().pmp_init_noise(_ptr(x), _ptr(noise), int(seed), int(offset), float(var_temp), _ptr(start), _ptr(goal),
                                B, H, D, _stream(x.device)))
    return x


def step(kind, x, eps_c, eps_u, w, coef, noise=None, seed=0, offset=0, start=None, goal=None, base=0, out=None):
    """one fused DDPM (kind 0) / DDIM (kind 1) update; eps_c/eps_u lists of [B,H,D] tensors"""
    K = len(eps_c)
    B, H, D = x.shape
    out = torch.empty_like(x) if out is None else out
    ec = (C.c_void_p * K)(*[e.data_ptr() for e in eps_c])
    eu = (C.c_void_p * K)(*[e.data_ptr() for e in eps_u])
    wv = (C.c_float * K)(*[float(v) for v in w])
    cf = (C.c_float * 8)(*[float(v) for v in coef])
    _check(lib().pmp_step(int(kind), _ptr(x), ec, eu, wv, K, int(base), cf, _ptr(noise), int(seed), int(offset),
                          _ptr(start), _ptr(goal), _ptr(out), B, H, D, _stream(x.device)))
    return out


def sample(engines, wembs, w, start, goal, H, kind, t_steps, coefs, noise=None, seed=0, row_offset=0, base=0,
           var_temp=1.0, trace=False):
    """whole reverse-diffusion loop on the GPU (pmp_sample).  engines: list of UnetEngine (K composed
    potentials); wembs: list of [B,64]; t_steps: int array [n]; coefs: float32 array [n,8]"""
    K = len(engines)
    dev = engines[0].device
    start = _f32(start, dev)
    goal = _f32(goal, dev)
    B, D = start.shape
    n = len(t_steps)
    a = SampleArgs()
    a.K = K
    keep = []
    for k in range(K):
        a.models[k] = engines[k].h.value
        wk = _f32(wembs[k], dev)
        keep.append(wk)
        a.wemb_dev[k] = wk.data_ptr()
        a.w[k] = float(w[k])
    a.base = int(base)
    a.B, a.H, a.D = B, int(H), D
    a.kind = int(kind)
    a.n_steps = n
    th = np.ascontiguousarray(t_steps, np.int32)
    ch = np.ascontiguousarray(coefs, np.float32).reshape(n, 8)
    a.t_host = th.ctypes.data
    a.coef_host = ch.ctypes.data
    if noise is not None:
        noise = _f32(noise, dev)
        assert tuple(noise.shape) == (n + 1, B, H, D), noise.shape
        a.noise_dev = noise.data_ptr()
    a.seed = int(seed)
    a.row_offset = int(row_offset)
    a.var_temp = float(var_temp)
    a.start_dev = start.data_ptr()
    a.goal_dev = goal.data_ptr()
    x = torch.empty(B, H, D, device=dev, dtype=torch.float32)
    a.x_dev = x.data_ptr()
    tr = None
    if trace:
        tr = torch.empty(B, n + 1, H, D, device=dev, dtype=torch.float32)
        a.trace_dev = tr.data_ptr()
    _check(lib().pmp_sample(C.byref(a), _stream(dev)))
    return (x, tr) if trace else x


def unnormalize(x, lo, hi):
    D = x.shape[-1]
    x = x.contiguous()
    lo = _f32(torch.as_tensor(lo), x.device)
    hi = _f32(torch.as_tensor(hi), x.device)
    out = torch.empty_like(x)
    _check(lib().pmp_unnormalize(_ptr(x), _ptr(lo), _ptr(hi), _ptr(out), x.numel() // D, D, _stream(x.device)))
    return out


def _walls64(cen, hext, dev):
    cen = torch.as_tensor(np.ascontiguousarray(cen, np.float64))
    if cen.dim() == 2:
        cen = cen[None]
    he = torch.as_tensor(np.broadcast_to(np.asarray(hext, np.float64), tuple(cen.shape)).copy())
    return cen.to(dev).contiguous(), he.to(dev).contiguous()


def colchk_maze2d_swept(traj, env_id, cen, hext, eps=0.08, margin=0.01, legacy_div=True, lo=(0., 0.), hi=(5., 5.)):
    """traj [N,H,2] cuda float32 (un-normalised); returns (valid uint8 [N], nchk int32 [N]) on the GPU"""
    dev = traj.device
    traj = traj.contiguous()
    N, H, _ = traj.shape
    env_id = torch.as_tensor(env_id).to(dev, torch.int32).contiguous()
    cen, he = _walls64(cen, hext, dev)
    valid = torch.empty(N, dtype=torch.uint8, device=dev)
    nchk = torch.empty(N, dtype=torch.int32, device=dev)
    _check(lib().pmp_colchk_maze2d_swept(_ptr(traj), _ptr(env_id), N, H, _ptr(cen), _ptr(he), cen.shape[1],
                                         lo[0], lo[1], hi[0], hi[1], float(eps), float(margin), int(legacy_div),
                                         _ptr(valid), _ptr(nchk), _stream(dev)))
    return valid, nchk


def colchk_maze2d_points(traj, env_id, cen, hext, margin=0.01, lo=(0., 0.), hi=(5., 5.)):
    dev = traj.device
    traj = traj.contiguous()
    N, H, _ = traj.shape
    env_id = torch.as_tensor(env_id).to(dev, torch.int32).contiguous()
    cen, he = _walls64(cen, hext, dev)
    ncol = torch.empty(N, dtype=torch.int32, device=dev)
    pt = torch.empty(N, H, dtype=torch.uint8, device=dev)
    _check(lib().pmp_colchk_maze2d_points(_ptr(traj), _ptr(env_id), N, H, _ptr(cen), _ptr(he), cen.shape[1],
                                          lo[0], lo[1], hi[0], hi[1], float(margin), _ptr(ncol), _ptr(pt),
                                          _stream(dev)))
    return ncol, pt


def colchk_kuka(traj, env_id, boxc, hext, n_arms, base, sph, table_z=-0.001):
    dev = traj.device
    traj = traj.contiguous()
    N, H, _ = traj.shape
    env_id = torch.as_tensor(env_id).to(dev, torch.int32).contiguous()
    bc = torch.as_tensor(np.ascontiguousarray(boxc, np.float32))
    if bc.dim() == 2:
        bc = bc[None]
    bh = torch.as_tensor(np.broadcast_to(np.asarray(hext, np.float32), tuple(bc.shape)).copy())
    bc = bc.to(dev).contiguous()
    bh = bh.to(dev).contiguous()
    base = np.ascontiguousarray(base, np.float32)
    sph = np.ascontiguousarray(sph, np.float32)
    valid = torch.empty(N, dtype=torch.uint8, device=dev)
    nchk = torch.empty(N, dtype=torch.int32, device=dev)
    ncol = torch.empty(N, dtype=torch.int32, device=dev)
    pt = torch.empty(N, H, dtype=torch.uint8, device=dev)
    _check(lib().pmp_colchk_kuka(_ptr(traj), _ptr(env_id), N, H, int(n_arms),
                                 base.ctypes.data_as(C.POINTER(C.c_float)), sph.ctypes.data_as(C.POINTER(C.c_float)),
                                 len(sph), _ptr(bc), _ptr(bh), bc.shape[1], float(table_z), _ptr(valid), _ptr(nchk),
                                 _ptr(ncol), _ptr(pt), _stream(dev)))
    return dict(valid=valid, nchk=nchk, ncol=ncol, ptcol=pt)


def pick_first_valid(valid, nchk, n_prob, n_samp):
    dev = valid.device
    chosen = torch.empty(n_prob, dtype=torch.int32, device=dev)
    suc = torch.empty(n_prob, dtype=torch.uint8, device=dev)
    tot = torch.empty(n_prob, dtype=torch.int32, device=dev)
    _check(lib().pmp_pick_first_valid(_ptr(valid.contiguous()), _ptr(nchk.contiguous()), n_prob, n_samp, _ptr(chosen),
                                      _ptr(suc), _ptr(tot), _stream(dev)))
    return chosen, suc, tot
