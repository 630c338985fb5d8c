"""GPU diagnostic (not a pytest file): run-to-run differences of the training gradient on identical inputs (first call on a
fresh workspace vs later calls), per gemm path."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "potential-motion-plan-release_b200"))
from pmp_b200 import train as ptrain, workloads as synth  # noqa: E402
from diffuser.models import TemporalUnet_WCond  # noqa: E402

dev = torch.device("cuda:0")
fam, B, H = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
path = int(sys.argv[4]) if len(sys.argv) > 4 else 3
a = synth.arch(fam)
D = a["obs_dim"]
model = TemporalUnet_WCond(**a["unet"])
model.load_state_dict(synth.synth_state_dict(model.state_dict(), 1))
model = model.to(dev)
te = ptrain.TrainEngine(model, dev, 100)
te.engine.set_gemm_path(path)
g = torch.Generator().manual_seed(3)
xn = torch.randn(B, H, D, generator=g).clamp(-1, 1)
noise = torch.randn(B, H, D, generator=g)
t = torch.randint(0, 100, (B,), generator=g)
walls = torch.tensor(synth.make_problems(fam, B, 5)["walls"])
keep = (torch.rand(B, generator=g) > 0.2).float()
lw = torch.ones(H, D); lw[0] = 0; lw[-1] = 0
G = []
for i in range(4):
    if i == 2:   # dirty the workspace with garbage between runs
        import ctypes as C
        import pmp_b200
        n = pmp_b200.lib().pmp_model_workspace_bytes(te.engine.h)
        print("workspace %.2f GB" % (n / 1e9))
    loss = te.loss_grad(xn, t, walls, noise, keep, lw)
    torch.cuda.synchronize()
    G.append(te.G.clone())
    print("run %d loss %.7f |G| %.6e" % (i, float(loss), float(te.G.norm())), flush=True)
for i in range(4):
    for j in range(i + 1, 4):
        print("run %d vs %d: rel diff %.3e" % (i, j, float((G[i] - G[j]).norm() / G[j].norm())))
names = te.flat.names
d01 = sorted(((float((x - y).norm() / (y.norm() + 1e-30)), k) for k, x, y in zip(names, te.flat.views(G[0]), te.flat.views(G[1]))), reverse=True)[:6]
print("worst 0 vs 1:", d01)
