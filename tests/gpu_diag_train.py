"""GPU diagnostic (not a pytest file): per-parameter error of the CUDA training step against the CPU oracle."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "potential-motion-plan-release_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import train_ref  # noqa: E402
from test_train_oracle import golden_batch  # noqa: E402
from pmp_b200 import train as ptrain  # noqa: E402
from oracle import synth  # noqa: E402
from diffuser.models import TemporalUnet_WCond  # noqa: E402

dev = torch.device("cuda:0")
fams = sys.argv[1].split(",") if len(sys.argv) > 1 else ["m2d", "dualkuka"]
paths = [int(p) for p in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 3]
for fam in fams:
    cfg, sd, xn, t, walls, noise, keep, lw, z = golden_batch(fam)
    t0 = time.time()
    loss_ref, g_ref, eps_ref = train_ref.loss_and_grads(sd, cfg, xn, t, walls, noise, keep, lw)
    print(fam, "oracle loss %.6f (golden %.6f) in %.1fs" % (float(loss_ref), float(z["loss"]), time.time() - t0), flush=True)
    a = synth.arch(fam)
    for path in paths:
        model = TemporalUnet_WCond(**a["unet"])
        model.load_state_dict(sd)
        model = model.to(dev)
        te = ptrain.TrainEngine(model, dev, a["diffusion"]["n_timesteps"])
        te.engine.set_gemm_path(path)
        try:
            loss, eps = te.loss_grad(xn, t, walls, noise, keep, lw, want_eps=True)
            torch.cuda.synchronize()
        except Exception as e:
            print(fam, "path", path, "FAILED:", e, flush=True)
            continue
        e_err = float((eps.cpu() - eps_ref).abs().max() / eps_ref.abs().max())
        print(fam, "path", path, "loss %.6f  rel err %.2e  eps err %.2e" % (float(loss), abs(float(loss) - float(loss_ref)) / float(loss_ref), e_err), flush=True)
        gv = te.grad_views()
        gn = float(torch.sqrt(sum((g.double() ** 2).sum() for g in g_ref.values())))
        worst = []
        for k in g_ref:
            d = (gv[k].cpu().double() - g_ref[k].double()).norm().item()
            n = g_ref[k].double().norm().item()
            worst.append((d / max(n, 1e-6 * gn), d / gn, k, n))
        tot = float(np.sqrt(sum((w[1] * gn) ** 2 for w in worst)) / gn)
        print("  total normwise grad err %.3e" % tot)
        for r, ra, k, n in worst:
            if r > 2e-3 or os.environ.get("VERBOSE"):
                print("   %-52s rel %.3e  (of total %.3e) |g| %.3e" % (k, r, ra, n))
        # timing of the step
        for _ in range(2):
            te.loss_grad(xn, t, walls, noise, keep, lw)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(5):
            te.loss_grad(xn, t, walls, noise, keep, lw)
        torch.cuda.synchronize()
        print("  %.2f ms per loss_grad (B=%d)" % ((time.time() - t0) / 5 * 1e3, xn.shape[0]), flush=True)
        del te, model
print("done")
