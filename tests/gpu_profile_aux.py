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
HBM = 6549.1      # GB/s, MEASURED_PEAKS.json (copy bandwidth)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


H, D = 48, 2
nb = 400000
xs = torch.randn(nb, H, D, device=dev); ec = torch.randn_like(xs); eu = torch.randn_like(xs); xo = torch.empty_like(xs)
sg = torch.zeros(nb, D, device=dev)
coef = np.array([1.01, 0.1, 0.3, 0.7, 0.05, 0, 0, 0], np.float32)
t = timed(lambda: pmp_b200.step(0, xs, [ec], [eu], [2.0], coef, noise=None, seed=7, offset=0, start=sg, goal=sg, out=xo))
print("step_kernel4: %.4f ms, %.0f GB/s = %.3f of HBM" % (t * 1e3, 16.0 * H * D * nb / t / 1e9, 16.0 * H * D * nb / t / 1e9 / HBM))
del xs, ec, eu, xo

ntr, E = 1000000, 5000
pr = synth.make_problems("m2d_nw3", E, 5)
al = torch.linspace(0, 1, H, device=dev)[None, :, None]
s_ = torch.rand(ntr, 1, 2, device=dev) * 4.6 + 0.2; g_ = torch.rand(ntr, 1, 2, device=dev) * 4.6 + 0.2
tr = (s_ * (1 - al) + g_ * al).contiguous()
eid = ((torch.arange(ntr, device=dev) // 10) % E).to(torch.int32)
cen, he = pmp_b200._lib._walls64(pr["boxes"], pr["hext"], dev)
vb = torch.empty(ntr, dtype=torch.uint8, device=dev); nc_ = torch.empty(ntr, dtype=torch.int32, device=dev)
import ctypes as C
L = pmp_b200.lib()
def chk():
    rc = L.pmp_colchk_maze2d_swept(C.c_void_p(tr.data_ptr()), C.c_void_p(eid.data_ptr()), ntr, H, C.c_void_p(cen.data_ptr()),
                                   C.c_void_p(he.data_ptr()), cen.shape[1], 0.0, 0.0, 5.0, 5.0, 0.08, 0.01, 1,
                                   C.c_void_p(vb.data_ptr()), C.c_void_p(nc_.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
t = timed(chk)
by = (4.0 * H * D + 5) * ntr
print("maze2d swept (nw=3): %.4f ms, %.0f GB/s = %.3f of HBM, valid %.3f" % (t * 1e3, by / t / 1e9, by / t / 1e9 / HBM, vb.float().mean().item()))
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
