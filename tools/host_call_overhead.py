"""Python-side overhead of the prepared host call (HostStep) on top of pinn_loss_fwd_bwd_host."""
import sys, time, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pinn_for_quantum_wavefunction_surfaces_b200 as pk
import bench
n = 1 << 18
bs = [bench.synth_batch(n, s).pin_memory() for s in range(8)]
th = np.ascontiguousarray(bench.load_theta())
w = np.array([1.0 / n, 2.0 / n, 2.0 / n])
steps = [pk.HostStep("poc", b[0], b[1], b[2], b[3]) for b in bs]
h = pk.Handle.get(0)
for i in range(20): steps[i % 8](th, w)
tot = []; inner = []
for i in range(200):
    t0 = time.perf_counter_ns()
    steps[i % 8](th, w)
    t1 = time.perf_counter_ns()
    tot.append((t1 - t0) / 1e3); inner.append(h.host_timing()["total"])
print("python per call %.1f us, inside C %.1f us, difference %.1f us" % (np.median(tot), np.median(inner), np.median(np.array(tot) - np.array(inner))))
t0 = time.perf_counter_ns()
for i in range(200): steps[i % 8](th, w)
t1 = time.perf_counter_ns()
print("tight loop %.1f us per call" % ((t1 - t0) / 200e3))
fn = h.L.pinn_version
t0 = time.perf_counter_ns()
for i in range(20000): fn()
print("empty ctypes call %.2f us" % ((time.perf_counter_ns() - t0) / 20000e3))
