"""GPU profiling driver (not a pytest file): a few launches of the HBM-bound kernels of the path on inputs larger than
the L2 -- fused sampler update, Maze2D swept check, Kuka sphere check -- for `ncu --set full` (profiles/README.md)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "potential-motion-plan-release_b200"))

import pmp_b200  # noqa: E402
from pmp_b200 import workloads as synth  # noqa: E402
from diffuser.guides.kuka_colchk import sphere_table, arm_bases, TABLE_Z  # noqa: E402

dev = torch.device("cuda:0")
H, D = 48, 2
nb = 400000
xs = torch.randn(nb, H, D, device=dev); ec = torch.randn_like(xs); eu = torch.randn_like(xs); xo = torch.empty_like(xs)
sg = torch.zeros(nb, D, device=dev)
coef = np.array([1.01, 0.1, 0.3, 0.7, 0.05, 0, 0, 0], np.float32)
for _ in range(4):
    pmp_b200.step(0, xs, [ec], [eu], [2.0], coef, noise=None, seed=7, offset=0, start=sg, goal=sg, out=xo)
del xs, ec, eu, xo

ntr, E = 1000000, 5000
pr = synth.make_problems("m2d_nw3", E, 5)
al = torch.linspace(0, 1, H, device=dev)[None, :, None]
s_ = torch.rand(ntr, 1, 2, device=dev) * 4.6 + 0.2; g_ = torch.rand(ntr, 1, 2, device=dev) * 4.6 + 0.2
tr = (s_ * (1 - al) + g_ * al).contiguous()
eid = ((torch.arange(ntr, device=dev) // 10) % E).to(torch.int32)
for _ in range(4):
    pmp_b200.colchk_maze2d_swept(tr, eid, pr["boxes"], pr["hext"])
del tr

nk, Ek, Hk = 200000, 2000, 52
prk = synth.make_problems("kuka", Ek, 6)
lo, hi = prk["lo"], prk["hi"]
q = (torch.rand(nk, Hk, 7, device=dev) * torch.tensor(hi - lo, device=dev) + torch.tensor(lo, device=dev)).contiguous()
ek = (torch.arange(nk, device=dev) % Ek).to(torch.int32)
sph = sphere_table()
for _ in range(4):
    pmp_b200.colchk_kuka(q, ek, np.asarray(prk["boxes"], np.float32), float(prk["hext"]), 1, arm_bases(1), sph, TABLE_Z)
torch.cuda.synchronize()
print("ok")
