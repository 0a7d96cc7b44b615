"""Python-side overhead of the prepared host call (HostStep) on top of pinn_loss_fwd_bwd_host, and the spread of the
per-call time over many page-locked batch buffers (bench.py's e2e leg uses 40 of them)."""
import sys, time, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pinn_for_quantum_wavefunction_surfaces_b200 as pk
import bench
n = 1 << 18
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 40
bs = [bench.synth_batch(n, s).pin_memory() for s in range(nb)]
th = np.ascontiguousarray(bench.load_theta())
w = np.array([1.0 / n, 2.0 / n, 2.0 / n])
steps = [pk.HostStep("poc", b[0], b[1], b[2], b[3]) for b in bs]
h = pk.Handle.get(0)
for i in range(3): steps[i % nb](th, w)
tot = []; inner = []; wait = []
t_all = time.perf_counter_ns()
for i in range(200):
    t0 = time.perf_counter_ns()
    steps[(3 + i) % nb](th, w)
    t1 = time.perf_counter_ns()
    ht = h.host_timing()
    tot.append((t1 - t0) / 1e3); inner.append(ht["total"]); wait.append(ht["wait"])
t_all = (time.perf_counter_ns() - t_all) / 200e3
tot, inner, wait = np.array(tot), np.array(inner), np.array(wait)
print("%d buffers: python per call median %.1f us (mean %.1f, loop incl. bookkeeping %.1f), inside C median %.1f (mean %.1f), wait median %.1f max %.1f"
      % (nb, np.median(tot), tot.mean(), t_all, np.median(inner), inner.mean(), np.median(wait), wait.max()))
print("first 10 calls (python us):", np.round(tot[:10], 1), " calls 40..49:", np.round(tot[40:50], 1))
