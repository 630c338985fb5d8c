"""train entry point (reference: scripts/train.py:49-197 + train_pb_diff.sh) on the B200 training step.

The reference builds its model / diffusion / trainer from `Config` objects, pickles them into the log directory and runs
`trainer.train(n_train_steps)` over a `SequenceDataset` of planner demonstrations (HDF5 files that do not exist here).  This
entry point keeps that flow -- `model_config.pkl`, `diffusion_config.pkl`, `state_<epoch>.pt = {'step','model','ema'}` in
`--logdir`, loadable by `scripts/plan_rm2d.py --logdir` and by the reference's own `utils.load_potential_diffusion_model` --
and trains on SYNTHETIC demonstrations (collision-unaware straight segments between free start / goal points with a little
noise, one obstacle layout per sample): enough to drive the loss / gradient / Adam / EMA kernels end to end, not a planner
worth evaluating.  Under `torchrun --nproc-per-node N` it trains data parallel (NCCL all-reduce of the gradient arena).

    python scripts/train.py --family m2d --n_train_steps 200 --batch_size 128 --logdir /tmp/pmp_logdir
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from pmp_b200 import workloads
from diffuser.utils import Config
from diffuser.utils.training_pbdiff import TrainerPBDiff


class SyntheticDemos(torch.utils.data.IterableDataset):
    """endless stream of (trajectories [H,D], conditions {0, H-1}, wall_locations [1,Wd]) samples, the batch layout of the
    reference's dataset (datasets/sequence.py: Batch(trajectories, conditions, wall_locations))"""

    def __init__(self, family, horizon, seed):
        self.family, self.H, self.seed = family, horizon, seed

    def __iter__(self):
        rng_seed = self.seed
        al = np.linspace(0, 1, self.H, dtype=np.float32)[:, None]
        while True:
            pr = workloads.make_problems(self.family, 64, rng_seed)
            rng = np.random.default_rng(rng_seed)
            rng_seed += 1
            s = workloads.normalize(pr["start"], pr["lo"], pr["hi"])
            g = workloads.normalize(pr["goal"], pr["lo"], pr["hi"])
            for i in range(len(s)):
                x = np.clip(s[i] * (1 - al) + g[i] * al + 0.03 * rng.standard_normal((self.H, s.shape[1])).astype(np.float32), -1, 1)
                x[0], x[-1] = s[i], g[i]
                yield (torch.from_numpy(x.astype(np.float32)), {0: torch.from_numpy(x[0].copy()), self.H - 1: torch.from_numpy(x[-1].copy())},
                       torch.from_numpy(pr["walls"][i][None].astype(np.float32)))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--family', default='m2d', choices=workloads.FAMILIES)
    ap.add_argument('--logdir', required=True)
    ap.add_argument('--n_train_steps', type=int, default=100)
    ap.add_argument('--batch_size', type=int, default=64)
    ap.add_argument('--learning_rate', type=float, default=2e-4)
    ap.add_argument('--gradient_accumulate_every', type=int, default=1)
    ap.add_argument('--ema_decay', type=float, default=0.995)
    ap.add_argument('--step_start_ema', type=int, default=20)
    ap.add_argument('--update_ema_every', type=int, default=10)
    ap.add_argument('--save_freq', type=int, default=0, help='0: only at the end')
    ap.add_argument('--log_freq', type=int, default=20)
    ap.add_argument('--seed', type=int, default=0)
    args = ap.parse_args(argv)

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    a = workloads.arch(args.family)
    H = a['unet']['horizon']
    os.makedirs(args.logdir, exist_ok=True)
    torch.manual_seed(args.seed)                 # same initial weights on every rank
    np.random.seed(args.seed + rank)
    # scripts/train.py:64-102: the Configs are pickled into the log directory, then instantiated
    save = (lambda name: (args.logdir, name)) if rank == 0 else (lambda name: None)
    model_config = Config('models.TemporalUnet_WCond', savepath=save('model_config.pkl'), **a['unet'])
    diffusion_config = Config('models.GaussianDiffusionPB', savepath=save('diffusion_config.pkl'), **a['diffusion'])
    model = model_config()
    diffusion = diffusion_config(model).to(dev)
    torch.manual_seed(args.seed + 1000 * rank)   # per-rank draws of t / noise from here on
    data = SyntheticDemos(args.family, H, seed=10_000 * (rank + 1) + args.seed)
    trainer = TrainerPBDiff(diffusion, data, train_batch_size=args.batch_size, train_lr=args.learning_rate,
                            gradient_accumulate_every=args.gradient_accumulate_every, ema_decay=args.ema_decay,
                            step_start_ema=args.step_start_ema, update_ema_every=args.update_ema_every,
                            save_freq=args.save_freq, log_freq=args.log_freq if rank == 0 else 0, label_freq=1,
                            results_folder=args.logdir, n_train_steps=args.n_train_steps, shuffle=False)
    trainer.train(args.n_train_steps)
    torch.cuda.synchronize()
    if rank == 0:
        path = trainer.save(trainer.step)
        print('saved', path, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return trainer


if __name__ == '__main__':
    main()
