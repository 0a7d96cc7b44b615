"""One-box A/B of step-kernel builds through the stable part of the C ABI only (pinn_create / pinn_loss_fwd_bwd), so that
libraries of other rounds or variant builds (tools/dbg/*.so) can be timed side by side:

    python tools/ab_time.py [--steps 300] [--rounds 3] lib_a.so lib_b.so ...

Each round times every library once (interleaved, so that clock / thermal drift hits all alike): 2^18 points per step, 40
rotating batches (> L2), loss weights given, CUDA events around `steps` steps after 50 warm-up steps.  Prints ms/step per
library and round, and the relative difference of Ltot and of the gradient to the first library."""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("libs", nargs="+")
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--rounds", type=int, default=3)
ap.add_argument("--points", type=int, default=1 << 18)
a = ap.parse_args()
dev = torch.device("cuda:0")
n = a.points
nb = max(4, -(-160_000_000 // (16 * n)))
batches = [bench.synth_batch(n, 100 + b).to(dev) for b in range(nb)]
theta = torch.from_numpy(bench.load_theta().astype(np.float32)).to(dev)
wts = []
for b in batches:
    r1 = torch.sqrt((b[0] - b[3]) ** 2 + b[1] ** 2 + b[2] ** 2)
    r2 = torch.sqrt((b[0] + b[3]) ** 2 + b[1] ** 2 + b[2] ** 2)
    wts.append(torch.tensor([1.0 / n, 1.0 / float((r1 >= 17.5).sum()), 1.0 / float((r2 >= 17.5).sum())], dtype=torch.float64, device=dev))
vp, i32, i64, u32, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint32, ctypes.c_float
libs = []
for path in a.libs:
    L = ctypes.CDLL(os.path.abspath(path))
    L.pinn_create.argtypes = [i32, ctypes.POINTER(vp)]
    L.pinn_loss_fwd_bwd.argtypes = [vp, i32, i64, vp, vp, vp, vp, i32, vp, vp, vp, u32, f32, vp, vp, vp, vp]
    h = vp()
    assert L.pinn_create(0, ctypes.byref(h)) == 0, path
    libs.append((os.path.basename(path), L, h))
out = torch.zeros(8 + 1521, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream


def step(L, h, i):
    b = batches[i % nb]
    rc = L.pinn_loss_fwd_bwd(h, 0, n, b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr(), b[3].data_ptr(), 0, None, theta.data_ptr(),
                             wts[i % nb].data_ptr(), 0xFFFF, 17.5, out.data_ptr(), out.data_ptr() + 64, None, st)
    assert rc == 0, rc


ref = None
for name, L, h in libs:
    step(L, h, 0)
    torch.cuda.synchronize()
    r = out.clone()
    if ref is None:
        ref = r
    print("%-28s Ltot %.10e  rel dLtot %.2e  rel dgrad %.2e" % (name, float(r[0]), abs(float(r[0] - ref[0])) / abs(float(ref[0])),
                                                                float((r[8:] - ref[8:]).abs().max() / ref[8:].abs().max())))
res = {name: [] for name, _, _ in libs}
for rnd in range(a.rounds):
    for name, L, h in libs:
        for i in range(50):
            step(L, h, i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            step(L, h, 50 + i)
        e1.record()
        torch.cuda.synchronize()
        res[name].append(e0.elapsed_time(e1) / a.steps)
for name in res:
    v = res[name]
    print("%-28s ms/step %s  best %.5f  -> %.4e points/s" % (name, " ".join("%.5f" % x for x in v), min(v), n / min(v) * 1e3))
