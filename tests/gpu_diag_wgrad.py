"""GPU diagnostic (not a pytest file): tcgen05 weight-gradient GEMM against the fp32 CUDA-core one and torch."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "potential-motion-plan-release_b200"))
import pmp_b200  # noqa: E402

L = pmp_b200.lib()
dev = torch.device("cuda:0")
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
shapes = [(4, 48, 64, 64, 5), (3, 24, 256, 64, 5), (2, 12, 512, 1024, 5), (5, 14, 768, 768, 1), (256, 14, 768, 768, 5)]
if len(sys.argv) > 2:
    shapes = shapes[:int(sys.argv[2])]
for R, H, Co, Ci, KS in shapes:
    g = torch.Generator().manual_seed(R + H)
    o1, o2 = torch.randn(R, H, Co, generator=g).to(dev), torch.randn(R, H, Co, generator=g).to(dev)
    i1, i2 = torch.randn(R, H, Ci, generator=g).to(dev), torch.randn(R, H, Ci, generator=g).to(dev)
    n = L.pmp_debug_wgrad_scratch_floats(R, H, Co, Ci, KS)
    scratch = torch.empty(n, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    out = {}
    for tc in (0, 1):
        dw = torch.zeros(Co, Ci, KS, device=dev)
        rc = L.pmp_debug_wgrad(p(o1), p(i1), p(o2), p(i2), H, Co, Ci, KS, R, p(dw), p(scratch), tc, flags, st)
        assert rc == 0, L.pmp_last_error()
        torch.cuda.synchronize()
        out[tc] = dw
    # torch reference in fp64 on the GPU
    pad = KS // 2
    ref = torch.zeros(Co, Ci, KS, device=dev, dtype=torch.float64)
    for o, i in ((o1, i1), (o2, i2)):
        ip = torch.nn.functional.pad(i.double(), (0, 0, pad, pad))
        for k in range(KS):
            ref[:, :, k] += torch.einsum("rhc,rhd->cd", o.double(), ip[:, k:k + H])
    e0 = float((out[0].double() - ref).norm() / ref.norm())
    e1 = float((out[1].double() - ref).norm() / ref.norm())
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for tc in (0, 1):
        dw = torch.zeros(Co, Ci, KS, device=dev)
        for _ in range(2):
            L.pmp_debug_wgrad(p(o1), p(i1), p(o2), p(i2), H, Co, Ci, KS, R, p(dw), p(scratch), tc, flags, st)
        t0.record()
        for _ in range(5):
            L.pmp_debug_wgrad(p(o1), p(i1), p(o2), p(i2), H, Co, Ci, KS, R, p(dw), p(scratch), tc, flags, st)
        t1.record(); torch.cuda.synchronize()
        ms.append(t0.elapsed_time(t1) / 5)
    fl = 2.0 * 2 * R * H * Co * Ci * KS
    print("R=%d H=%d Co=%d Ci=%d KS=%d: simt err %.2e (%.3f ms, %.1f TF/s)  tc err %.2e (%.3f ms, %.1f TF/s)" % (
        R, H, Co, Ci, KS, e0, ms[0], fl / ms[0] / 1e9, e1, ms[1], fl / ms[1] / 1e9), flush=True)
print("done flags", flags)
