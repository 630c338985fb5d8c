"""TrainerPBDiff: the training loop of the reference (diffuser/utils/training_pbdiff.py:19-205) on the B200 training
step -- same constructor arguments, step order (gradient accumulation, Adam, EMA every `update_ema_every` steps,
`reset_parameters` before `step_start_ema`) and checkpoint format `state_<epoch>.pt = {'step','model','ema'}`.
Rendering / wandb logging of the reference are not part of it.  With torch.distributed initialised every rank trains
on its own batches and the flat gradient arena is all-reduced (NCCL) before the fused Adam + EMA kernel."""
import math
import os

import torch

from pmp_b200 import train as ptrain


def cycle(dl):
    while True:
        for data in dl:
            yield data


def warmup_cosine_lr(step, first_cycle_steps, max_lr, min_lr, warmup_steps):
    """CosineAnnealingWarmupRestarts (train_utils.py:34-119) with one cycle, gamma 1"""
    step = min(step, first_cycle_steps - 1) if step >= first_cycle_steps else step
    if step < warmup_steps:
        return (max_lr - min_lr) * step / max(warmup_steps, 1) + min_lr
    return min_lr + (max_lr - min_lr) * (1 + math.cos(math.pi * (step - warmup_steps) / (first_cycle_steps - warmup_steps))) / 2


class TrainerPBDiff(object):

    def __init__(self, diffusion_model, dataset, renderer=None, ema_decay=0.995, train_batch_size=32, train_lr=2e-5,
                 gradient_accumulate_every=2, step_start_ema=2000, update_ema_every=10, log_freq=100, sample_freq=1000,
                 save_freq=1000, label_freq=100000, save_parallel=False, n_reference=8, bucket=None, train_device='cuda',
                 save_checkpoints=False, results_folder='./results', n_samples=2, num_render_samples=2, clip_grad=False,
                 clip_grad_max=False, n_train_steps=None, trainer_dict={}, num_workers=0, shuffle=True):
        self.model = diffusion_model
        self.dataset = dataset
        self.batch_size = train_batch_size
        self.gradient_accumulate_every = gradient_accumulate_every
        self.update_ema_every, self.step_start_ema = update_ema_every, step_start_ema
        self.log_freq, self.save_freq, self.label_freq = log_freq, save_freq, label_freq
        self.logdir = results_folder
        self.clip_grad, self.clip_grad_max = clip_grad, clip_grad_max
        self.device = train_device
        self.lr_warmupDecay = trainer_dict.get('lr_warmupDecay', False)
        self.train_lr, self.n_train_steps = train_lr, n_train_steps
        self.warmup_steps = trainer_dict.get('warmup_steps_pct', 0.0) * (n_train_steps or 0)
        if dataset is not None:
            it = isinstance(dataset, torch.utils.data.IterableDataset)
            self.dataloader = cycle(torch.utils.data.DataLoader(dataset, batch_size=train_batch_size, num_workers=num_workers,
                                                                shuffle=shuffle and not it, pin_memory=True))
        self.model.train()
        self.te = self.model.train_engine()
        self.opt = ptrain.FlatAdamEMA(self.te, lr=train_lr, ema_decay=ema_decay, step_start_ema=step_start_ema,
                                      update_ema_every=update_ema_every)
        self.reducer = ptrain.GradAllReduce(self.te.G, bucket_mb=trainer_dict.get('bucket_mb', 32), engine=self.te.engine)
        self.step = 0

    # ------------------------------------------------------------------------------------------------------------
    def micro_batch(self, batch, first):
        """loss of one micro-batch; its gradient is accumulated in the flat arena (`first` starts a new sum)"""
        x, cond, walls = batch[0], batch[1], batch[2]
        dev = self.te.device
        x = x.to(dev, non_blocking=True).float()
        cond = {k: v.to(dev, non_blocking=True).float() for k, v in cond.items()}
        walls = walls.to(dev, non_blocking=True).float()
        m = self.model
        walls = torch.flatten(walls, 1, 2) if m.is_dyn_env else walls[:, 0, :]
        B = x.shape[0]
        import numpy as np
        from diffuser.models.helpers import apply_conditioning
        t = torch.randint(0, m.n_timesteps, (B,), device=dev).long()
        xs = x[:, :, m.action_dim:].contiguous()
        noise = torch.randn_like(xs)
        x_noisy = apply_conditioning(ptrain.q_sample(xs, noise, t, m.sqrt_alphas_cumprod, m.sqrt_one_minus_alphas_cumprod), cond, 0)
        if m.set_cond_noise_to_0:
            for k in cond.keys():
                noise[:, k, m.action_dim:] = 0.0
        keep = torch.as_tensor(np.random.rand(B) >= m.model.concept_drop_prob, dtype=torch.float32)
        return self.te.loss_grad(x_noisy, t, walls, noise, keep, m.loss_fn.weights, zero=first, check_t=False)   # t drawn in range above

    def train_step(self, batches):
        """one optimizer step over `gradient_accumulate_every` micro-batches (training_pbdiff.py:121-143)"""
        loss = None
        for i, batch in enumerate(batches):
            loss = self.micro_batch(batch, first=(i == 0))
        self.reducer.run()                               # sum over the ranks (overlaps the reverse pass of the last micro-batch)
        scale = 1.0 / (len(batches) * self.reducer.world)
        if self.clip_grad_max or self.clip_grad:
            # clipping acts on the averaged gradient (clip_grad_value_ / clip_grad_norm_, training_pbdiff.py:125-129)
            G = self.te.G
            G.mul_(scale)
            scale = 1.0
            if self.clip_grad_max:
                G.clamp_(-self.clip_grad_max, self.clip_grad_max)
            if self.clip_grad:
                G.mul_(torch.clamp(self.clip_grad / (G.norm() + 1e-6), max=1.0))
        if self.lr_warmupDecay:
            self.opt.lr = warmup_cosine_lr(self.step, self.n_train_steps, self.train_lr, self.train_lr / 100., self.warmup_steps)
        self.opt.update(grad_scale=scale)
        self.step += 1
        return loss

    def train(self, n_train_steps):
        for _ in range(n_train_steps):
            loss = self.train_step([next(self.dataloader) for _ in range(self.gradient_accumulate_every)])
            if self.save_freq and (self.step - 1) % self.save_freq == 0:
                self.save((self.step - 1) // self.label_freq * self.label_freq)
            if self.log_freq and (self.step - 1) % self.log_freq == 0:
                print('%d: %8.4f' % (self.step - 1, float(loss) / 1.0), flush=True)

    # ------------------------------------------------------------------------------------------------------------
    def save(self, epoch):
        """state_<epoch>.pt = {'step', 'model', 'ema'} (training_pbdiff.py:177-191)"""
        ema_sd = {k: v.clone() for k, v in self.model.state_dict().items()}
        for k, v in zip(self.te.flat.names, self.te.flat.views(self.opt.ema)):
            ema_sd['model.' + k] = v.clone()
        os.makedirs(self.logdir, exist_ok=True)
        path = os.path.join(self.logdir, 'state_%s.pt' % epoch)
        torch.save({'step': self.step, 'model': self.model.state_dict(), 'ema': ema_sd}, path)
        return path

    def load(self, epoch):
        data = torch.load(os.path.join(self.logdir, 'state_%s.pt' % epoch), map_location=self.te.device, weights_only=True)
        self.step = self.opt.step = data['step']
        self.model.load_state_dict(data['model'])      # copies into the flat arena (parameters are views of it)
        for k, v in zip(self.te.flat.names, self.te.flat.views(self.opt.ema)):
            v.copy_(data['ema']['model.' + k])
        self.te.dirty = True
