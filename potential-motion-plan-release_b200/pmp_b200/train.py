"""Training step of the potential diffusion model on the B200.

Replaces `GaussianDiffusionPB.loss` + `loss.backward()` + `Adam.step` + `EMA.update_model_average`
(diffuser/models/diffusion_pb.py:480-545, diffuser/utils/training_pbdiff.py:118-140, utils/train_utils.py:15-31):
the energy loss 0.5*mean(w*(dE/dx - noise)^2) and its gradient with respect to every parameter are computed by
hand-written kernels (pmp_train_loss_grad: forward, backward-data, tangent forward, reverse pass over the pair), the
optimizer and the EMA copy by one fused kernel over flat arenas, and -- with more than one GPU -- the gradient arena is
all-reduced over NCCL in buckets that overlap the reverse pass (GradAllReduce).  PyTorch supplies device memory, streams and
torch.distributed; no torch autograd, cuDNN or cuBLAS call is on this path."""
import ctypes as C

import numpy as np
import torch

from ._lib import lib, _check, _ptr, _stream, _f32, PmpError, UnetEngine


class TrainArgs(C.Structure):
    _fields_ = [("x_noisy_dev", C.c_void_p), ("t_dev", C.c_void_p), ("walls_dev", C.c_void_p), ("wall_len", C.c_int32),
                ("noise_dev", C.c_void_p), ("keep_dev", C.c_void_p), ("loss_weights_dev", C.c_void_p),
                ("B", C.c_int32), ("H", C.c_int32), ("loss_dev", C.c_void_p), ("eps_dev", C.c_void_p),
                ("energy_dev", C.c_void_p)]


class FlatParams:
    """every parameter of `module` becomes a view into ONE flat float32 arena `P` (reference layouts, named_parameters
    order, 16-byte aligned), so that the optimizer, the EMA copy and the gradient all-reduce are passes over flat memory
    and state_dict() / load_state_dict() / checkpoints keep working on the module unchanged."""

    def __init__(self, module, device):
        self.names, self.offsets, self.shapes = [], [], []
        off = 0
        params = sorted(module.named_parameters(), key=lambda kp: self._rank(kp[0]))     # stable
        for k, p in params:
            self.names.append(k)
            self.offsets.append(off)
            self.shapes.append(tuple(p.shape))
            off += (p.numel() + 3) // 4 * 4
        self.total = off
        self.P = torch.zeros(off, device=device, dtype=torch.float32)
        for (k, p), o in zip(params, self.offsets):
            v = self.P[o:o + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v

    @staticmethod
    def _rank(name):
        """forward module order of the U-Net (embeddings / wall encoder, down path, middle, up path, final conv): the reverse
        pass then finishes the gradient arena from its end towards its start (bucketed all-reduce, GradAllReduce)"""
        for i, pfx in enumerate(("downs.", "mid_block1.", "mid_block2.", "ups.", "final_conv.")):
            if name.startswith(pfx):
                return i + 1
        return 0

    def views(self, flat):
        return [flat[o:o + int(np.prod(s))].view(s) for o, s in zip(self.offsets, self.shapes)]


class TrainEngine:
    """owns the flat arenas of one TemporalUnet_WCond mirror module and the pmp_model bound to them"""

    def __init__(self, unet, device, n_timesteps):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise PmpError("the training step runs on a B200 only (got device %s); no CPU fallback" % dev)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.unet = unet
        self.flat = FlatParams(unet, dev)
        self.G = torch.zeros_like(self.flat.P)          # gradient of the current micro-batch (written by the kernels)
        self.engine = UnetEngine(unet._ctor, unet.state_dict(), dev, n_timesteps=n_timesteps)
        n = len(self.flat.names)
        keys = (C.c_char_p * n)(*[k.encode() for k in self.flat.names])
        offs = (C.c_int64 * n)(*self.flat.offsets)
        _check(lib().pmp_train_bind(self.engine.h, n, keys, offs, _ptr(self.flat.P), _ptr(self.G), C.c_int64(self.flat.total)))
        self.loss_buf = torch.zeros(1, device=dev, dtype=torch.float32)
        self.energy_buf = torch.zeros(1, device=dev, dtype=torch.float32)   # sum of the last batch's energies (`engy_batch`)
        self.dirty = False      # parameters changed since the kernels' weight layouts were refreshed

    # -------------------------------------------------------------------------------------------------------------
    def sync_weights(self, rebuild_tables=False):
        """re-layout of the flat parameters for the conv kernels (after an optimizer step / load_state_dict)"""
        _check(lib().pmp_train_sync_weights(self.engine.h, 1 if rebuild_tables else 0, _stream(self.device)))
        self.dirty = False

    def loss_grad(self, x_noisy, t, walls, noise, keep, loss_weights, want_eps=False, zero=True, check_t=True):
        """loss [1] (device) of the micro-batch; its gradient is written (zero=True) or added to the flat arena `G`"""
        dev = self.device
        if self.dirty:
            self.sync_weights()
        x_noisy = _f32(x_noisy, dev)
        B, H, D = x_noisy.shape
        t = t.detach().to(dev, torch.int64).contiguous()
        if check_t and t.numel() and (int(t.min()) < 0 or int(t.max()) >= self.engine.arch.n_timesteps):   # one host read
            raise PmpError("timestep out of range for n_timesteps=%d" % self.engine.arch.n_timesteps)
        walls = _f32(walls, dev)
        noise = _f32(noise, dev)
        keep = _f32(torch.as_tensor(keep), dev)
        w = _f32(loss_weights, dev)
        assert tuple(noise.shape) == (B, H, D) and tuple(w.shape) == (H, D) and keep.numel() == B and walls.shape[0] == B
        if zero:
            self.G.zero_()
        eps = torch.empty_like(x_noisy) if want_eps else None
        a = TrainArgs()
        a.x_noisy_dev, a.t_dev, a.walls_dev, a.wall_len = x_noisy.data_ptr(), t.data_ptr(), walls.data_ptr(), walls.shape[1]
        a.noise_dev, a.keep_dev, a.loss_weights_dev = noise.data_ptr(), keep.data_ptr(), w.data_ptr()
        a.B, a.H = B, H
        a.loss_dev = self.loss_buf.data_ptr()
        a.eps_dev = eps.data_ptr() if want_eps else None
        a.energy_dev = self.energy_buf.data_ptr()
        _check(lib().pmp_train_loss_grad(self.engine.h, C.byref(a), _stream(dev)))
        loss = self.loss_buf.clone()
        return (loss, eps) if want_eps else loss

    def grad_views(self):
        return dict(zip(self.flat.names, self.flat.views(self.G)))


def q_sample(x0, noise, t, sqrt_acp, sqrt_1m_acp, zero_cond_noise=False):
    """q_sample + start/goal conditioning from x0 (diffusion_pb.py:468-487); noise is modified in place when
    zero_cond_noise (set_loss_noise)"""
    dev = x0.device
    x0 = x0.contiguous()
    B, H, D = x0.shape
    out = torch.empty_like(x0)
    t = t.to(dev, torch.int64).contiguous()
    _check(lib().pmp_q_sample(_ptr(x0), _ptr(noise), _ptr(t), _ptr(_f32(sqrt_acp, dev)), _ptr(_f32(sqrt_1m_acp, dev)), B, H, D,
                              1 if zero_cond_noise else 0, _ptr(out), _stream(dev)))
    return out


def adam_ema_step(P, G, M, V, ema, step, lr, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0, ema_mode=0, ema_beta=0.995):
    """fused torch.optim.Adam.step (no weight decay / amsgrad) + EMA over flat float32 arenas (in place)"""
    _check(lib().pmp_adam_ema_step(_ptr(P), _ptr(G), _ptr(M), _ptr(V), _ptr(ema), C.c_int64(P.numel()), C.c_float(lr),
                                   C.c_float(betas[0]), C.c_float(betas[1]), C.c_float(eps), int(step), C.c_float(grad_scale),
                                   int(ema_mode), C.c_float(ema_beta), _stream(P.device)))


class FlatAdamEMA:
    """optimizer + EMA state of a TrainEngine: Adam moments and the EMA parameters as flat arenas
    (training_pbdiff.py:57-113: Adam(lr), EMA(0.995), step_start_ema, update_ema_every)"""

    def __init__(self, te, lr=2e-5, ema_decay=0.995, step_start_ema=2000, update_ema_every=10, betas=(0.9, 0.999), eps=1e-8):
        self.te = te
        self.lr, self.betas, self.eps = lr, betas, eps
        self.ema_decay, self.step_start_ema, self.update_ema_every = ema_decay, step_start_ema, update_ema_every
        self.M = torch.zeros_like(te.flat.P)
        self.V = torch.zeros_like(te.flat.P)
        self.ema = te.flat.P.clone()          # reset_parameters(): ema := model
        self.t = 0                            # optimizer steps taken
        self.step = 0                         # trainer step counter (training_pbdiff.py:175)

    def update(self, grad_scale=1.0):
        """optimizer.step(); optimizer.zero_grad(); if step % update_ema_every == 0: step_ema(); step += 1"""
        self.t += 1
        mode = 0
        if self.step % self.update_ema_every == 0:
            mode = 2 if self.step < self.step_start_ema else 1
        adam_ema_step(self.te.flat.P, self.te.G, self.M, self.V, self.ema, self.t, self.lr, self.betas, self.eps, grad_scale,
                      mode, self.ema_decay)
        self.te.dirty = True
        self.te.unet._weights_epoch = getattr(self.te.unet, '_weights_epoch', 0) + 1     # the sampling engine re-reads the weights
        self.step += 1

    def ema_state_dict(self, like):
        """the EMA parameters under the module's parameter names (checkpoint key 'ema')"""
        sd = {k: v.clone() for k, v in like.state_dict().items()}
        for k, v in zip(self.te.flat.names, self.te.flat.views(self.ema)):
            sd[k] = v.clone()
        return sd


class GradAllReduce:
    """bucketed all-reduce (sum) of a flat gradient arena.  On the GPU (NCCL) the buckets are reduced on a side stream from
    the END of the arena towards its start, bucket k as soon as the reverse pass has finished every gradient at or above
    its lower bound (pmp_train_set_buckets / pmp_train_wait_bucket: one CUDA event per bucket, recorded inside
    pmp_train_loss_grad), so the reduction of the up path / middle / down path overlaps the rest of the reverse pass; the
    optimizer kernel is ordered behind the last bucket.  On CPU tensors (gloo, tests) the buckets are reduced in place."""

    def __init__(self, flat_grad, bucket_mb=32, group=None, engine=None):
        import torch.distributed as dist
        self.dist, self.G, self.group, self.engine = dist, flat_grad, group, engine
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        n = flat_grad.numel()
        per = max(4, int(bucket_mb * (1 << 20) // 4) // 4 * 4)
        self.bounds = [(max(0, e - per), e) for e in range(n, 0, -per)]      # back to front
        self.stream = None
        if self.world > 1 and flat_grad.is_cuda:
            self.stream = torch.cuda.Stream(device=flat_grad.device)
            if engine is not None:
                lo = (C.c_int64 * len(self.bounds))(*[b[0] for b in self.bounds])
                _check(lib().pmp_train_set_buckets(engine.h, len(self.bounds), lo))

    def run(self):
        """call right after loss_grad has been enqueued; returns when the reduction is ENQUEUED (stream-ordered)"""
        if self.world == 1:
            return
        if self.stream is None:                       # CPU tensors (gloo)
            for lo, hi in self.bounds:
                self.dist.all_reduce(self.G[lo:hi], op=self.dist.ReduceOp.SUM, group=self.group)
            return
        cur = torch.cuda.current_stream(self.G.device)
        if self.engine is None:
            self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            for k, (lo, hi) in enumerate(self.bounds):
                if self.engine is not None:
                    _check(lib().pmp_train_wait_bucket(self.engine.h, k, C.c_void_p(self.stream.cuda_stream)))
                self.dist.all_reduce(self.G[lo:hi], op=self.dist.ReduceOp.SUM, group=self.group)
        cur.wait_stream(self.stream)
