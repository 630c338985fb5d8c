"""GPU diagnostic (not a pytest file): layer-by-layer comparison of the CUDA U-Net forward /
analytic backward with the CPU oracle.  Usage: python tests/gpu_diag_unet.py [family ...]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "potential-motion-plan-release_b200"))

from oracle import synth, unet_ref  # noqa: E402
import pmp_b200  # noqa: E402


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def main(fams):
    dev = torch.device("cuda:0")
    for fam in fams:
        a = synth.arch(fam)
        cfg = unet_ref.UnetCfg(a["unet"])
        sd = unet_ref.synth_weights(a["unet"], 1)
        B, H, D = 3, a["horizon"], a["obs_dim"]
        pr = synth.make_problems(fam, B, 3)
        g = torch.Generator().manual_seed(11)
        x = torch.randn(2 * B, H, D, generator=g)
        t = torch.tensor([37] * (2 * B))
        walls = torch.tensor(pr["walls"]).repeat(2, 1)
        t0 = time.time()
        eps_ref, u_ref, tape = unet_ref.eps_analytic(sd, cfg, x, t, walls, n_cond=B, return_tape=True)
        eps_ag = unet_ref.eps_autograd(sd, cfg, x, t, walls, n_cond=B)
        print("[%s] oracle %.2fs  analytic-vs-autograd %.2e" % (fam, time.time() - t0, rel(eps_ref, eps_ag)))
        eng = pmp_b200.UnetEngine(a["unet"], sd, dev)
        eng.set_gemm_path(int(os.environ.get("PMP_GEMM_PATH", "3")))
        print("  gemm path", os.environ.get("PMP_GEMM_PATH", "3"))
        if os.environ.get("PMP_ROW_CHUNK"):
            eng.set_row_chunk(int(os.environ["PMP_ROW_CHUNK"]))
        wemb = eng.wall_embed(torch.tensor(pr["walls"]))
        wemb_ref = unet_ref.wall_embed(sd, cfg, torch.tensor(pr["walls"]))
        print("  wall_embed rel err %.3e" % rel(wemb.cpu().numpy(), wemb_ref.numpy()))
        eps, en, u = eng.eps(x, t, wemb, n_cond=B, want_energy=True, want_u=True)
        torch.cuda.synchronize()
        for pfx, val in tape.items():
            if not isinstance(val, tuple):
                continue
            yh, r = val
            got = eng.debug_fetch("yhat:" + pfx)
            if got is None:
                continue
            R_, C_, H_ = yh.shape
            got = got.reshape(R_, H_, C_).transpose(0, 2, 1)
            print("  yhat %-28s rel err %.3e" % (pfx, rel(got, yh.numpy())))
        print("  u    rel err %.3e" % rel(u.cpu().numpy(), u_ref.numpy()))
        print("  E    rel err %.3e" % rel(en.cpu().numpy(), (0.5 * u_ref * u_ref).sum(dim=(1, 2)).numpy()))
        print("  eps  rel err %.3e   (max|eps| %.3f)" % (rel(eps.cpu().numpy(), eps_ag.numpy()), eps_ag.abs().max()))
        for name in ("gx:ups.1.1.", "gx:ups.1.0.", "gx:ups.0.1.", "gx:ups.0.0.", "gx:mid_block2.", "gx:mid_block1.",
                     "gx:downs.2.1.", "gx:downs.2.0.", "gx:downs.1.1.", "gx:downs.1.0.", "gx:downs.0.1."):
            got = eng.debug_fetch(name)
            if got is not None:
                print("  %-20s finite=%s absmax=%.3e" % (name, bool(np.isfinite(got).all()), float(np.abs(got).max())))
        # timing at a moderate batch
        Bt = int(os.environ.get("PMP_DIAG_BT", "512"))
        xt = torch.randn(2 * Bt, H, D, device=dev)
        tt = torch.full((2 * Bt,), 50, device=dev, dtype=torch.long)
        wt = wemb[:1].repeat(Bt, 1)
        eng.eps(xt, tt, wt, n_cond=Bt)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        pmp_b200.profile_enable(True)
        e0.record()
        for _ in range(3):
            eng.eps(xt, tt, wt, n_cond=Bt)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        prof = pmp_b200.profile_read(2)
        pmp_b200.profile_enable(False)
        print("  eps-eval %d rows: %.2f ms  -> %.1f us/row" % (2 * Bt, ms, 1e3 * ms / (2 * Bt)))
        import ctypes as _C
        ctr = (_C.c_double * 8)()
        if pmp_b200.lib().pmp_debug_counters(ctr, 1) == 0 and ctr[6] > 0:
            n = ctr[6]
            if os.environ.get("PMP_TC_TIMERS") == "3":
                print("    staging waits (cycles, group 0 leader) [%s]: store-drain %.0f  acquire(total) %.0f per acquire (%d) | load-wait %.0f per load (%d) | tiles %d" % (
                    os.environ.get("PMP_TC_TIMER_MATCH", "all"), ctr[0] / max(ctr[3], 1), ctr[1] / max(ctr[3], 1), ctr[3],
                    ctr[2] / max(ctr[4], 1), ctr[4], n))
            elif os.environ.get("PMP_TC_TIMERS") == "2":
                print("    epilogue laps per tile (cycles) [%s]: wait_acc %.0f  loop1 %.0f  reduce %.0f  loop2 %.0f  release %.0f  preamble %.0f  (tiles %d)" % (
                    os.environ.get("PMP_TC_TIMER_MATCH", "all"), ctr[0] / n, ctr[1] / n, ctr[2] / n, ctr[3] / n, ctr[4] / n, ctr[5] / n, n))
            else:
              print("    [afull wait %.0f]" % (ctr[7] / n))
              print("    tc role timers per tile (cycles): mma_wait_tma %.0f  mma_wait_acc %.0f  mma_total %.0f | epi_wait %.0f  epi_busy %.0f | tma_wait_stage %.0f  (tiles %d)" % (
                ctr[0] / n, ctr[1] / n, ctr[2] / n, ctr[3] / n, ctr[4] / n, ctr[5] / n, n))
        for cls, (pms, pfl, pn) in zip(("simt", "tc"), prof):
            if pn:
                print("    conv_%s: %d launches, %.2f ms/eval, %.1f algorithmic TFLOP/s" % (cls, pn // 3, pms / 3, pfl / pms / 1e9))
        del eng


if __name__ == "__main__":
    main(sys.argv[1:] or ["m2d", "kuka", "dualkuka"])
