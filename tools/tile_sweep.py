"""Fixed cost vs per-super-tile cost of the fused step kernel: time(n) for n = 148*128*k points (k super-tiles per CTA).
   python tools/tile_sweep.py [variant]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pinn_for_quantum_wavefunction_surfaces_b200 as pk
from oracle import ref_autograd as ra
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
dev = torch.device("cuda:0")
h = pk.Handle.get(0)
th = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "checkpoints.npz"))["ionHsym"].astype(np.float32)).to(dev)
g = torch.Generator().manual_seed(5)
nmax = 148 * 128 * 32
x, y, z, R, i1, i2 = ra.sample_box(nmax, "poc" if variant == 0 else "trainpy", g)
xs = [t.ravel().float().to(dev) for t in (x, y, z, R)]
rows = []
for k in (1, 2, 3, 4, 6, 8, 12, 14, 16, 24, 32):
    n = 148 * 128 * k
    w = torch.tensor([1.0 / n, 2.0 / n, 2.0 / n], dtype=torch.float64, device=dev)
    a = [t[:n] for t in xs]
    for _ in range(5):
        pk.loss_and_grad_raw(variant, *a, th, None, w)
    torch.cuda.synchronize()
    h.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        pk.loss_and_grad_raw(variant, *a, th, None, w)
    e1.record()
    torch.cuda.synchronize()
    ms, cnt = h.profile_collect()
    rows.append((k, n, ms / cnt * 1e3, e0.elapsed_time(e1) / 50 * 1e3))
    print("k=%2d n=%7d  step kernel %7.2f us   whole step %7.2f us" % rows[-1])
ks = np.array([r[0] for r in rows], float); ts = np.array([r[2] for r in rows])
b, a = np.polyfit(ks, ts, 1)
print("fit: kernel = %.2f us + %.3f us per super-tile per CTA (%.0f cycles at 1965 MHz)" % (a, b, b * 1965))
