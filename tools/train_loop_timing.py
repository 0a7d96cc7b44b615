"""Device-resident training loop (pinn_trainer): time per optimizer step, CUDA-graph replay vs plain launches."""
import sys, time, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pinn_for_quantum_wavefunction_surfaces_b200 as pk
import bench
n = 1 << 18
dev = torch.device("cuda:0")
for graph in (True, False, True, False):
    tr = pk.Trainer("poc", n, bench.load_theta(), seed=1, lr=1e-6, device=0)
    tr.run(20, use_graph=graph); tr.read()
    ts = torch.cuda.ExternalStream(tr.h.L.pinn_trainer_stream(tr.t), device=dev)
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record(ts); t0 = time.time()
    tr.run(500, use_graph=graph)
    t1 = time.time()
    l1.record(ts); tr.read()
    ms = l0.elapsed_time(l1)
    print("graph=%s: %.2f us/step -> %.3e points/s (host enqueue %.1f us/step)" % (graph, ms / 500 * 1e3, n * 500 / (ms * 1e-3), (t1 - t0) / 500 * 1e6))
    tr.close()
