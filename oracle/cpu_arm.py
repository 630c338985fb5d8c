"""The reference's CPU path for the bench workload, timed on the host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/__init__.py): used by `bench.py --impl reference`
and by the `cpu_baseline` leg of the GPU arm, never by the product.

One *step* = the bench's workload (BASELINE.json configs[1]) on a bounded sample: n/2 Maze2D Static2 and
n/2 Concave problems, each planned with the FULL reverse diffusion (DDPM, T = 100, CFG pair, energy U-Net +
autograd dE/dx per step, diffusion_pb.py:249-281) in fp32 on all host threads, then un-normalised and
collision-scored with the swept check (rm2d_colchk.py:76-93).  Nothing is extrapolated.

kind "reference": the reference's own classes (TemporalUnet_WCond, GaussianDiffusionPB.conditional_sample and
                  the pb_diff_envs point robot / wall group), imported through oracle/refshim.py -- only where
                  the reference tree is present (the build container);
kind "port"     : the oracle restatement (bit-identical to the reference on CPU, tests/test_oracle_golden.py)
                  -- what runs on the GPU box, where /root/reference does not exist.
"""
import contextlib
import io
import os
import time

import numpy as np
import torch

from . import refshim, synth, unet_ref, sampler_ref, colchk_ref

FAMILIES = ("m2d_nw3", "m2d_concave")     # Static2 (3 walls, hExt 0.7) + Concave (7 obstacles)
T_DIFF = 100
COND_W = 2.0


class CpuArm:
    def __init__(self, threads=None, prefer_reference=True):
        self.threads = threads or (os.cpu_count() or 1)
        torch.set_num_threads(self.threads)
        self.kind = "reference" if (prefer_reference and refshim.available()) else "port"
        self.models = {}
        for pi, fam in enumerate(FAMILIES):
            a = synth.arch(fam)
            sd = unet_ref.synth_weights(a["unet"], 1 + pi)      # the GPU arm's weights (bench.py: seed 1 + family index)
            ent = dict(a=a, sd=sd, cfg=unet_ref.UnetCfg(a["unet"]))
            if self.kind == "reference":
                U, G = refshim.load_models()
                with refshim.quiet():
                    m = U(**a["unet"]).eval()
                    m.load_state_dict(sd)
                    d = G(m, **a["diffusion"]).eval()
                    d.horizon = a["horizon"]
                    d.condition_guidance_w = COND_W
                ent["diffusion"] = d
            self.models[fam] = ent
        if self.kind == "reference":
            self.geom = refshim.load_maze2d_geometry()

    # ------------------------------------------------------------------ one family, n plans
    def _sample(self, fam, pr, seed):
        ent = self.models[fam]
        a = ent["a"]
        H, D = a["horizon"], a["obs_dim"]
        lo, hi = pr["lo"], pr["hi"]
        n = len(pr["start"])
        start = torch.tensor(synth.normalize(pr["start"], lo, hi))
        goal = torch.tensor(synth.normalize(pr["goal"], lo, hi))
        walls = torch.tensor(pr["walls"])
        torch.manual_seed(seed)
        if self.kind == "reference":
            with refshim.quiet():
                x = ent["diffusion"].conditional_sample({0: start, H - 1: goal}, walls, use_ddim=False)
            return x.detach()
        tb = sampler_ref.schedule(T_DIFF)
        x = sampler_ref.apply_cond(torch.randn(n, H, D), start, goal)
        for t in range(T_DIFF - 1, -1, -1):
            ec, eu = unet_ref.eps_pair(ent["sd"], ent["cfg"], x, torch.full((n,), t), walls)
            x = sampler_ref.ddpm_update(tb, x, t, sampler_ref.compose([ec], [eu], [COND_W]), torch.randn(n, H, D))
            x = sampler_ref.apply_cond(x, start, goal)
        return x

    def _score(self, fam, pr, x):
        un = synth.unnormalize(x.numpy(), pr["lo"], pr["hi"])
        n = len(un)
        if self.kind != "reference":
            valid, _ = colchk_ref.maze2d_swept(un, np.arange(n, dtype=np.int32), pr["boxes"], pr["hext"])
            return np.asarray(valid, bool)
        P, RW, RWG, Env = self.geom
        hext = float(pr["hext"])
        ok = np.zeros(n, bool)
        with contextlib.redirect_stdout(io.StringIO()):
            for i in range(n):
                rob = P(np.array([5., 5.]), RWG([RW(np.array(c), np.array([hext, hext])) for c in pr["boxes"][i]]), 0.01,
                        collision_eps=0.08)
                env = object.__new__(Env)
                env.robot = rob
                good = True
                for j in range(un.shape[1] - 1):          # rm2d_colchk.py:76-93
                    if not env.edge_fp(un[i, j], un[i, j + 1]):
                        good = False
                        break
                ok[i] = good
        return ok

    # ------------------------------------------------------------------ one bench step
    def step(self, n_plans, seed):
        """plans n_plans problems (half per family); returns (seconds, success bits)"""
        ok = []
        t0 = time.perf_counter()
        for pi, fam in enumerate(FAMILIES):
            n = n_plans - n_plans // 2 if pi == 0 else n_plans // 2
            if n == 0:
                continue
            pr = synth.make_problems(fam, n, 1000 + pi + 17 * seed)
            x = self._sample(fam, pr, seed)
            ok.append(self._score(fam, pr, x))
        return time.perf_counter() - t0, np.concatenate(ok)

    # ------------------------------------------------------------------ the collision checker alone
    def checker_rate(self, n_traj=2000, seed=3):
        """Swept Maze2D check (rm2d_colchk.py:76-93) of n_traj straight start-goal candidates (Static2, 10 candidates per
        environment -- the GPU arm's aux workload), ONE host thread like the reference.  Returns (trajectories/s, what ran):
        the reference's own Python loop where the tree is present, the C restatement oracle/colchk.c otherwise."""
        fam, H = "m2d_nw3", 48
        n_env = max(1, n_traj // 10)
        pr = synth.make_problems(fam, n_env, 4000 + seed)
        rng = np.random.default_rng(seed)
        s_ = rng.uniform(0.2, 4.8, (n_traj, 1, 2)).astype(np.float32)
        g_ = rng.uniform(0.2, 4.8, (n_traj, 1, 2)).astype(np.float32)
        al = np.linspace(0, 1, H, dtype=np.float32)[None, :, None]
        un = (s_ * (1 - al) + g_ * al).astype(np.float32)
        env = (np.arange(n_traj) // 10 % n_env).astype(np.int32)
        sub = dict(pr, boxes=pr["boxes"][env])
        if self.kind == "reference":
            t0 = time.perf_counter()
            self._score(fam, sub, torch.tensor(synth.normalize(un, pr["lo"], pr["hi"])))
            return n_traj / (time.perf_counter() - t0), "reference Python loop (edge_fp per segment), 1 thread"
        colchk_ref.maze2d_swept(un[:8], env[:8], pr["boxes"], pr["hext"])      # loads the library
        t0 = time.perf_counter()
        colchk_ref.maze2d_swept(un, env, pr["boxes"], pr["hext"])
        return n_traj / (time.perf_counter() - t0), "C restatement oracle/colchk.c, 1 thread"

    def describe(self, n_plans):
        return ("%d + %d plans per step (Static2 nw3 + Concave nw7), full DDPM-%d (CFG pair, autograd dE/dx) + swept collision "
                "check eps 0.08, fp32 torch CPU, %s" % (n_plans - n_plans // 2, n_plans // 2, T_DIFF,
                                                         "reference classes via oracle/refshim.py" if self.kind == "reference"
                                                         else "oracle port (bit-identical to the reference on CPU)"))
