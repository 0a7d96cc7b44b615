"""Small training / inference / host-entry / trainer calls for compute-sanitizer (one tool per GPU-box visit):
   compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pinn_for_quantum_wavefunction_surfaces_b200 as pk
from oracle import ref_autograd as ra
dev = torch.device("cuda:0")
th = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "checkpoints.npz"))["ionHsym"]
for variant, n in ((0, 5000), (1, 3001), (0, 129)):
    g = torch.Generator().manual_seed(7 + n)
    x, y, z, R, i1, i2 = ra.sample_box(n, "poc" if variant == 0 else "trainpy", g)
    cols = [t.ravel().float() for t in (x, y, z, R)]
    m = torch.zeros(n, dtype=torch.uint8); m[i1] |= 1; m[i2] |= 2
    w = torch.tensor([1.0 / n, 1.0 / max(len(i1), 1), 1.0 / max(len(i2), 1)], dtype=torch.float64)
    d = [c.to(dev) for c in cols]
    t32 = torch.from_numpy(th.astype(np.float32)).to(dev)
    s, gth, _ = pk.loss_and_grad_raw(variant, *d, t32, m.to(dev), w.to(dev), want_E=True)
    s2, _, _ = pk.loss_and_grad_raw(variant, *d, t32, None, None)
    f = pk.fields("poc" if variant == 0 else "trainpy", *d, t32)
    hs = pk.HostStep("poc" if variant == 0 else "trainpy", *[c.pin_memory() for c in cols], mask=m.pin_memory())
    hsum, _ = hs(np.ascontiguousarray(th), w.numpy())
    hs2 = pk.HostStep("poc" if variant == 0 else "trainpy", *cols)
    hsum2, _ = hs2(np.ascontiguousarray(th), w.numpy())
    torch.cuda.synchronize()
    print(variant, n, float(s[0]), float(s2[0]), float(hsum[0]), float(hsum2[0]), float(f["psi"].abs().max()))
tr = pk.Trainer("trainpy", 4096, th, seed=3, history_capacity=4)
tr.run(4)
print("trainer", tr.read()["history"][:, 0])
tr.close()
print(pk.analysis.grid_sums(th, 2.0, n=24)["psi2"])
print("SANITIZE_CASE_OK")
