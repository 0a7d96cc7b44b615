#!/usr/bin/env python
"""Benchmark of the PINN residual-and-gradient hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--points P]

A step = one training evaluation (forward + Laplacian + loss + parameter gradients; for N>1 including the
sum of the 8 + 1521 float64 over the ranks, fused into the reduction kernel or, --allreduce nccl, as a
separate all-reduce) of the poc-form ionHsym model on one batch of P synthetic collocation points per GPU
(default 2^18 = BASELINE config 3; weak scaling: every GPU gets its own P points).
Prints ONE JSON line (rank 0): value (device-resident inputs), roofline (kernel events, ncu figures of the
committed capture), e2e (page-locked host inputs through HostStep / pinn_loss_fwd_bwd_host), and at N=1
cpu_baseline, reference_autograd_on_gpu, dense_grid_inference, device_train_loop.  See DESIGN.md section 5.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_POINT = 27044.0          # canonical algorithmic FLOPs of one poc training step
FP32_PEAK_MEASURED = 72.0e12      # tools/microbench/pipes.cu on this pool's B200: 3.60e13 FFMA/s (profiles/r01_microbench_pipes.jsonl)
FP32_PEAK_NOMINAL = 148 * 128 * 2 * 1.965e9
N_BATCHES = 40                    # rotating input batches: 40 x 4 MiB = 168 MB > 126 MB L2


def synth_batch(n, seed, variant="poc"):
    """x,y,z ~ U(-18,18), R ~ U(0.2,4), clamp at 0.005 from the nuclei (poc/main.py:124-156)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(4, n, generator=g, dtype=torch.float32)
    x, y, z = (2 * u[0] - 1) * 18, (2 * u[1] - 1) * 18, (2 * u[2] - 1) * 18
    R = 0.2 + (4.0 - 0.2) * u[3]
    r1 = torch.sqrt((x - R) ** 2 + y ** 2 + z ** 2)
    r2 = torch.sqrt((x + R) ** 2 + y ** 2 + z ** 2)
    x = torch.where((r1 < 0.005) | (r2 < 0.005), torch.full_like(x, 0.005), x)
    return torch.stack([x, y, z, R]).contiguous()


def load_theta():
    return np.load(os.path.join(ROOT, "tests", "golden", "checkpoints.npz"))["ionHsym"]


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.05 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[3 + i] == "Active" for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(float(r[2]) for r in rows)}


def cpu_reference_points_per_s(theta, min_seconds, max_points=1 << 16, threads=None):
    """The reference's own way of computing the step (nested autograd, float64, all host threads):
    oracle/ref_autograd.py.  Bounded sample: `max_points` points per evaluation, repeated >= min_seconds."""
    import torch
    from oracle import ref_autograd as ra
    if threads:
        torch.set_num_threads(threads)
    cores = torch.get_num_threads()
    g = torch.Generator().manual_seed(1234)
    x, y, z, R, i1, i2 = ra.sample_box(max_points, "poc", g)
    th = torch.tensor(theta, dtype=torch.float64)
    ra.loss_and_grad("poc", th, x, y, z, R, i1, i2)  # warm-up
    times = []
    t_all = time.time()
    while time.time() - t_all < min_seconds or len(times) < 2:
        t0 = time.time()
        ra.loss_and_grad("poc", th, x, y, z, R, i1, i2)
        times.append(time.time() - t0)
    best = min(times)
    return max_points / best, cores, "%d points x %d evaluations (best of), float64 nested autograd" % (max_points, len(times)), best


def ref_autograd_on_gpu(theta, dev, n=1 << 18):
    """BASELINE config 2: the reference's own algorithm (nested torch autograd, oracle/ref_autograd.py) executed by
    PyTorch on the same B200, float64 as shipped and float32, on one 2^18-point batch."""
    import torch
    from oracle import ref_autograd as ra
    g = torch.Generator().manual_seed(4321)
    x, y, z, R, i1, i2 = ra.sample_box(n, "poc", g)
    out = {}
    for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
        a = [t.to(device=dev, dtype=dt) for t in (x, y, z, R)]
        j1, j2 = i1.to(dev), i2.to(dev)
        th = torch.tensor(theta, dtype=dt, device=dev)
        for _ in range(2):
            ra.loss_and_grad("poc", th, *a, j1, j2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            ra.loss_and_grad("poc", th, *a, j1, j2)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[name] = {"value": n / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms}
    out["what"] = "torch eager nested autograd of the reference model on cuda:0, %d points, %d evaluations each" % (n, reps)
    return out


def dense_grid_inference(dev, n_axis=464):
    """BASELINE config 5: psi, H psi and the Simpson-weighted energy sums on a 464^3 = 1.0e8-point grid at one R
    (models/ionHsym_fineTune.pt weights), points generated in-kernel, only 5 sums leave each SM."""
    import torch
    import pinn_for_quantum_wavefunction_surfaces_b200 as pk
    th = np.load(os.path.join(ROOT, "tests", "golden", "checkpoints.npz"))["ionHsym_fineTune"]
    pk.analysis.grid_sums(th, 2.0, n=80, device=dev.index)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = pk.analysis.grid_sums(th, 2.0, n=n_axis, device=dev.index)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    pts = float(n_axis) ** 3
    # fixed-R inference: 6 656 FLOP/point
    return {"value": pts / (ms * 1e-3), "unit": "points/s", "ms": ms, "grid": "%d^3" % n_axis,
            "E_int": r["psiHpsi"] / r["psi2"], "frac_of_fp32_peak": 6656.0 * pts / (ms * 1e-3) / FP32_PEAK_MEASURED}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    theta = load_theta()
    import torch
    from oracle import ref_autograd as ra
    cores = torch.get_num_threads()
    ns = 1 << 16
    g = torch.Generator().manual_seed(1234)
    x, y, z, R, i1, i2 = ra.sample_box(ns, "poc", g)
    th = torch.tensor(theta, dtype=torch.float64)
    for _ in range(max(1, min(args.warmup, 3))):
        ra.loss_and_grad("poc", th, x, y, z, R, i1, i2)
    steps = max(1, min(args.steps, 40))
    t0 = time.time()
    for _ in range(steps):
        ra.loss_and_grad("poc", th, x, y, z, R, i1, i2)
    dt = (time.time() - t0) / steps
    v = ns / dt
    sample = "each step = %d-point sample of the 2^18-point batch; float64 nested autograd (oracle/ref_autograd.py)" % ns
    emit({
        "impl": "reference", "metric": "collocation points/sec per training step (fwd+lap+bwd)", "value": v,
        "unit": "points/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "ionHsym poc-form training step, 2^18 collocation points/step/GPU (BASELINE config 3)",
                   "sample_points": ns},
        "cpu_baseline": {"value": v, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def hbm_ceiling(bytes_per_launch, kernel_ms):
    """The contract's other denominator: the kernel against the measured HBM copy rate (MEASURED_PEAKS.json, burst figure for
    a kernel timed alone; the profiling recipe's 6650 GB/s "of fallback" when the file is absent).  27 kFLOP per 16 bytes puts this path
    three orders of magnitude on the compute side of the ridge, which is why `bound` is the FP32 pipe."""
    peak, src = 6650.0, "of fallback (B200_PROFILING.md)"
    pj = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pj):
        try:
            peak, src = float(json.load(open(pj))["hbm_gbs"]), "of measured (MEASURED_PEAKS.json hbm_gbs)"
        except (KeyError, ValueError):
            pass
    ach = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    return {"achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": src}


_RESULT_OUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries below us write there too (NCCL prints its version banner on
    communicator creation when NCCL_DEBUG is set in the environment), so file descriptor 1 is pointed at stderr for the
    whole run and the result line goes to a private duplicate of the original stdout."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    _RESULT_OUT.write(json.dumps(line) + "\n")
    _RESULT_OUT.flush()


def tensor_ceiling(flop_per_launch, kernel_ms):
    """SURVEY section 8(d)'s denominator for contractions moved to the tensor pipe: the measured dense tensor rate in TF32
    (= half the bf16 figure of MEASURED_PEAKS.json, burst, for a kernel timed alone; 1590 TFLOP/s bf16 "of fallback"
    without the file) divided by 3 for 3xTF32.  The numerator is the canonical FLOP count of the WHOLE step, of which about
    half (layer-2 mat-vecs of the MLP and the E-net, forward and reverse) runs on tcgen05; ncu's tensor-pipe utilisation
    (roofline.ncu_pipe_utilisation_pct.tensor) is the direct measurement."""
    bf16, src = 1590.0, "of fallback (B200_PROFILING.md)"
    pj = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pj):
        try:
            bf16, src = float(json.load(open(pj))["bf16_tflops"]), "of measured (MEASURED_PEAKS.json bf16_tflops / 2 / 3)"
        except (KeyError, ValueError):
            pass
    peak = bf16 / 2.0 / 3.0
    ach = flop_per_launch / (kernel_ms * 1e-3) / 1e12
    return {"achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "peak_source": src}


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=1 << 18, help="collocation points per GPU per step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--engine", default="tcgen05", choices=["tcgen05", "ffma"],
                    help="implementation of the fused step kernel (include/pinn_b200.h: pinn_set_engine)")
    ap.add_argument("--allreduce", default="fused", choices=["fused", "nccl"],
                    help="N>1: sum over ranks fused into the reduction kernel over NVLink peer memory (pinn_dp_*), or a "
                         "separate NCCL all-reduce per step (the baseline it replaces)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import pinn_for_quantum_wavefunction_surfaces_b200 as pk
    from pinn_for_quantum_wavefunction_surfaces_b200 import dp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.points
    W = max(args.warmup, 3)
    K = args.steps
    h = pk.Handle.get(local)
    h.set_engine(args.engine)
    fused = world > 1 and args.allreduce == "fused"
    if fused and not dp.attach_fused(h, strict=False):
        fused = False   # all ranks agreed: no peer mapping on this box -> the NCCL all-reduce path (reported in config)
        if rank == 0:
            print("bench: fused exchange unavailable, falling back to --allreduce nccl", file=sys.stderr)

    theta = torch.from_numpy(load_theta().astype(np.float32)).to(dev)
    host_batches = [synth_batch(n, 1000 * rank + b).pin_memory() for b in range(N_BATCHES)]
    dev_batches = [b.to(dev) for b in host_batches]
    out = torch.zeros(dp.N_OUT, dtype=torch.float64, device=dev)
    # boundary sets are derived in-kernel (r >= 17.5); their global sizes for the 1/count weights come from a
    # count pass per batch done up front (the sampler knows them in a real run)
    wts = []
    for b in host_batches:   # on the host, so that the only kernels this process launches are the library's own
        r1 = torch.sqrt((b[0] - b[3]) ** 2 + b[1] ** 2 + b[2] ** 2)
        r2 = torch.sqrt((b[0] + b[3]) ** 2 + b[1] ** 2 + b[2] ** 2)
        wts.append(dp.global_weights(n, int((r1 >= 17.5).sum()), int((r2 >= 17.5).sum()), device=dev))

    def step(i):
        b = dev_batches[i % N_BATCHES]
        pk.loss_and_grad_raw(0, b[0], b[1], b[2], b[3], theta, None, wts[i % N_BATCHES], sums=out[:8], dtheta=out[8:])
        if world > 1 and not fused:
            dist.all_reduce(out)

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step(i)
    fence()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    # Two back-to-back timed regions of the same K steps (clocks sampled over both): the first without any extra
    # stream operation gives `value`; the second brackets every step-kernel launch with CUDA events on the launching
    # stream (pinn_profile_begin/collect) and gives the kernel's average duration for the roofline.  (The event
    # records sit between the step kernel and its programmatic-dependent reduction kernel and cost a few
    # microseconds per step, which is why they are kept out of the first region.)
    launches0 = h.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    t0 = time.time()
    e0.record()
    for i in range(K):
        step(W + i)
    e1.record()
    fence()
    ms = e0.elapsed_time(e1)
    launches = h.launch_count() - launches0
    h.profile_begin()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(K):
        step(W + K + i)
    p1.record()
    fence()
    t1 = time.time()
    ms_profiled = p0.elapsed_time(p1)
    kern_ms, kern_n = h.profile_collect()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    loss = float(out[0].item())

    # ---- end to end: host buffers through the C ABI (H2D of the batch, kernel, D2H of loss+gradient), every step
    th64 = np.ascontiguousarray(load_theta())
    wts_h = [np.ascontiguousarray(w.cpu().numpy()) for w in wts]   # the caller's sampler knows the set sizes
    # the public call of a host-resident caller: one prepared HostStep per (reused, pinned) batch buffer
    host_steps = [pk.HostStep("poc", b[0], b[1], b[2], b[3], device=local) for b in host_batches]

    def e2e_step(i):
        sums_h, dth_h = host_steps[i % N_BATCHES](th64, wts_h[i % N_BATCHES])
        if world > 1 and not fused:
            out[:8] = torch.from_numpy(sums_h).to(dev)
            out[8:] = torch.from_numpy(dth_h).to(dev)
            dist.all_reduce(out)

    Ke = max(3, min(K, 200))
    for i in range(N_BATCHES + 3):   # one pass over every page-locked batch buffer first (the GPU's first touch of a
        e2e_step(i)                  # host page is slower), then the timed calls
    fence()
    te0 = time.time()
    for i in range(Ke):
        e2e_step(N_BATCHES + 3 + i)
    fence()
    te = (time.time() - te0) / Ke
    e2e_split = h.host_timing()   # of the last call
    h.profile_begin()             # the kernel's share, measured on a few more calls outside the timed loop
    for i in range(10):
        e2e_step(N_BATCHES + 3 + Ke + i)
    fence()
    e2e_kern_ms, e2e_kern_n = h.profile_collect()
    tte = torch.tensor([te], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tte, op=dist.ReduceOp.MAX)
    te = float(tte.item())

    # ---- the whole training loop on the device (sampler -> loss/gradient -> Adam, CUDA-graph replay): N=1 only
    loop = None
    if world == 1 or fused:
        tr = pk.Trainer("poc", n, load_theta(), seed=1, lr=1e-6, device=local)
        Kl = max(10, min(K, 300))
        ts = torch.cuda.ExternalStream(tr.h.L.pinn_trainer_stream(tr.t), device=dev)
        res = {}
        for graph in (False, True):
            tr.run(20, use_graph=graph)
            tr.read()
            l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0.record(ts)
            tr.run(Kl, use_graph=graph)
            l1.record(ts)
            tr.read()
            lms = torch.tensor([l0.elapsed_time(l1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(lms, op=dist.ReduceOp.MAX)
            res[graph] = n * world * Kl / (float(lms.item()) * 1e-3)
        loop = {"value": res[False], "unit": "points/s", "steps": Kl, "value_with_cuda_graph_replay": res[True],
                "what": "pinn_trainer: per step the fused loss/gradient kernel and one kernel with reduction%s + float64 Adam + "
                        "Philox sampler of the next batch, launched as programmatic dependents, no host sync"
                        % (" + set-size and gradient exchange over NVLink" if world > 1 else "")}
        tr.close()

    # ---- two more reference points for N = 1 (BASELINE configs 2 and 5) ----
    extra = {}
    if world == 1 and not args.no_cpu_baseline:
        extra["reference_autograd_on_gpu"] = ref_autograd_on_gpu(load_theta(), dev)
        extra["dense_grid_inference"] = dense_grid_inference(dev)

    if rank == 0:
        total_points = float(n) * world
        value = total_points * K / (ms * 1e-3)
        kern_avg_ms = kern_ms / max(kern_n, 1)
        achieved = FLOP_PER_POINT * n / (kern_avg_ms * 1e-3)
        traffic, ncu_pipes = None, None
        tf = os.path.join(ROOT, "profiles", "traffic.json")   # numbers of the committed ncu --set full capture of this kernel
        if os.path.exists(tf) and args.engine == "tcgen05":
            tj = json.load(open(tf))
            traffic, ncu_pipes = tj.get("dram_bytes_per_launch"), tj.get("ncu_pipe_utilisation_pct")
        line = {
            "metric": "collocation points/sec per training step (fwd+lap+bwd)",
            "value": value, "unit": "points/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "ms_per_step_with_kernel_events": ms_profiled / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ionHsym poc-form training step, 2^18 collocation points/step/GPU (BASELINE config 3)",
                       "points_per_gpu": n, "global_points": int(total_points), "weights": "models/ionHsym.pt (tests/golden/checkpoints.npz)",
                       "inputs": "%d rotating batches resident in HBM, %.0f MB > 126 MB L2" % (N_BATCHES, N_BATCHES * n * 16 / 1e6),
                       "parallelism": ("dp%d point sharding; sum of 1529 f64 over ranks %s" % (
                           world, "fused into the reduction kernel (NVLink peer stores + flags, pinn_dp_*)" if fused
                           else "by one NCCL all-reduce per step")) if world > 1 else "single GPU"},
            "roofline": {"bound": "fp32_ffma", "achieved": achieved / 1e12, "peak": FP32_PEAK_MEASURED / 1e12,
                         "unit": "TFLOP/s", "frac": achieved / FP32_PEAK_MEASURED,
                         "frac_of_nominal_74.4": achieved / FP32_PEAK_NOMINAL,
                         "peak_source": "measured FFMA rate on this pool's B200 (tools/microbench/pipes.cu); MEASURED_PEAKS.json has no FP32 entry",
                         "flop_per_point": FLOP_PER_POINT, "engine": args.engine,
                         "kernel": "pinn_step_tc_kernel<2,true>" if args.engine == "tcgen05" else "pinn_step_kernel<2,4,true>",
                         "kernel_ms": kern_avg_ms, "kernel_launches_timed": kern_n, "traffic": traffic,
                         "algorithmic_bytes_per_launch": 16 * n, "ncu_pipe_utilisation_pct": ncu_pipes,
                         "vs_hbm_ceiling": hbm_ceiling(16 * n, kern_avg_ms),
                         "vs_3xtf32_tensor_ceiling": tensor_ceiling(FLOP_PER_POINT * n, kern_avg_ms)},
            "e2e": {"value": total_points / te, "unit": "points/s", "h2d_bytes_per_step": int(16 * n + 1521 * 4),
                    "d2h_bytes_per_step": int((8 + 1521) * 8), "ms_per_step": te * 1e3, "steps": Ke, "kernel_ms": e2e_kern_ms / max(e2e_kern_n, 1),
                    "last_call_us": {k: round(v, 1) for k, v in e2e_split.items()},
                    "api": "HostStep -> pinn_loss_fwd_bwd_host (pinned float32 host batches read in place by the kernel: cp.async of the next super-tile over PCIe while the current one is computed; pageable inputs are staged in up to 4 chunks)"},
            "gpu_launches": int(launches), "clocks": clocks, "loss": loss,
        }
        if loop:
            line["device_train_loop"] = loop
        line.update(extra)
        if not args.no_cpu_baseline and world == 1:
            v, cores, sample, _ = cpu_reference_points_per_s(th64, args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample}
        emit(line)
    if fused:
        h.dp_status()
        dp.detach_fused(h)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
