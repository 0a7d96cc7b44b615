"""Device-resident training loop: the reference's two `train()` drivers (train.py:21-72; poc/main.py:359-430) with the
sampler, the fused loss/gradient kernel and Adam all on the GPU, one CUDA-graph replay per step.

The reference loops stay usable unchanged through `patch_nn_ion` / `run_train_py` (they keep torch's CPU RNG stream,
which parity runs need); this module is the fast path for runs that do not need that stream:

    sample(...)              <->  train.py:26-39; sampling()+radial()+torch.where  poc/main.py:124-156, 390-393
    adam_step(...)           <->  optimizer.step() + best/history bookkeeping      train.py:58-72; poc/main.py:403-417
    Trainer                  <->  train()                                          train.py:21-72; poc/main.py:359-430
    train_trainpy / train_poc   convenience drivers with the reference's defaults and return values
"""
import ctypes

import numpy as np
import torch

from . import params as P
from ._lib import Handle, PinnError, TrainConfig

INT64_MAX = (1 << 63) - 1
BOX_POC = (-18.0, 18.0, -18.0, 18.0, -18.0, 18.0, 0.2, 4.0)      # set_params() poc/main.py:17-27
BOX_TRAINPY = (-18.0, 18.0, -18.0, 18.0, -18.0, 18.0, 0.2, 3.0)  # train.py:80-83


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _handle(device=None):
    if not torch.cuda.is_available():
        raise PinnError("no CUDA device: the PINN hot path has no CPU fallback")
    return Handle.get(torch.cuda.current_device() if device is None else device)


def sample(n, seed, batch, box=BOX_POC, cutoff=0.005, bcutoff=17.5, device=None):
    """One batch of the device sampler -> dict(x, y, z, R float32 CUDA (n,), mask uint8, counts int64[2], weights f64[3])."""
    h = _handle(device)
    dev = torch.device("cuda", h.device)
    out = {k: torch.empty(n, dtype=torch.float32, device=dev) for k in "xyzR"}
    out["mask"] = torch.empty(n, dtype=torch.uint8, device=dev)
    out["counts"] = torch.zeros(2, dtype=torch.int64, device=dev)
    out["weights"] = torch.empty(3, dtype=torch.float64, device=dev)
    b = (ctypes.c_float * 8)(*box)
    stream = torch.cuda.current_stream(dev).cuda_stream
    rc = h.L.pinn_sample(h.h, n, seed, batch, b, cutoff, bcutoff, _ptr(out["x"]), _ptr(out["y"]), _ptr(out["z"]),
                         _ptr(out["R"]), _ptr(out["mask"]), _ptr(out["counts"]), _ptr(out["weights"]),
                         ctypes.c_void_p(stream))
    h.check(rc, "pinn_sample")
    return out


class AdamState:
    """Device state of the fused Adam step (float64, like the reference's parameters)."""

    def __init__(self, theta0, device=None, history_capacity=0, best_mode=0):
        h = _handle(device)
        self.h = h
        dev = torch.device("cuda", h.device)
        t64 = lambda a: torch.as_tensor(np.asarray(a, np.float64)).to(dev)
        self.theta = t64(theta0).clone()
        self.m = torch.zeros(P.N_THETA, dtype=torch.float64, device=dev)
        self.v = torch.zeros(P.N_THETA, dtype=torch.float64, device=dev)
        self.theta32 = self.theta.float()
        self.step = torch.zeros(1, dtype=torch.int64, device=dev)
        self.best_loss = torch.full((1,), 10.0, dtype=torch.float64, device=dev)  # Llim = 10, poc/main.py:370
        self.best_theta = self.theta.clone()
        self.best_step = torch.full((1,), -1, dtype=torch.int64, device=dev)
        self.hist = torch.zeros(max(history_capacity, 1), 4, dtype=torch.float64, device=dev)
        self.hist_cap = history_capacity
        self.best_mode = best_mode


def adam_step(state, grad, sums, n, lr=8e-3, betas=(0.9, 0.999), eps=1e-8, grad_mask=0xFFFF, best_after=0,
              history_mean_E=True):
    """One fused optimizer step on `state` (AdamState) from the device outputs of loss_and_grad_raw."""
    h = state.h
    stream = torch.cuda.current_stream(state.theta.device).cuda_stream
    rc = h.L.pinn_adam_step(h.h, _ptr(state.theta), _ptr(state.m), _ptr(state.v), _ptr(grad), _ptr(sums), _ptr(state.theta32),
                            _ptr(state.step), _ptr(state.best_loss), _ptr(state.best_theta), _ptr(state.best_step),
                            _ptr(state.hist) if state.hist_cap else None, state.hist_cap, n, lr, betas[0], betas[1], eps,
                            grad_mask, state.best_mode, best_after, int(history_mean_E), ctypes.c_void_p(stream))
    h.check(rc, "pinn_adam_step")


class Trainer:
    """pinn_trainer: everything of a training run lives on the device until read()."""

    def __init__(self, variant, n, theta0, seed=12345, lr=8e-3, betas=(0.9, 0.999), eps=1e-8, box=None, cutoff=0.005,
                 bcutoff=17.5, grad_mask=0xFFFF, best_mode=None, best_after=0, sc_sampling=1, freeze_after=INT64_MAX,
                 history_capacity=0, history_mean_E=None, device=None, handle=None):
        variant = {"poc": 0, "trainpy": 1}.get(variant, variant)
        self.h = handle if handle is not None else _handle(device)   # default: the cached per-device Handle
        if box is None:
            box = BOX_POC if variant == 0 else BOX_TRAINPY
        c = TrainConfig()
        c.variant, c.n, c.seed = variant, n, seed
        c.best_mode = (1 if variant == 0 else 0) if best_mode is None else best_mode
        c.history_mean_E = (0 if variant == 0 else 1) if history_mean_E is None else int(history_mean_E)
        c.sc_sampling, c.freeze_after, c.best_after, c.history_capacity = sc_sampling, freeze_after, best_after, history_capacity
        c.xL, c.xR, c.yL, c.yR, c.zL, c.zR, c.RL, c.RR = box
        c.cutoff, c.bcutoff, c.grad_mask = cutoff, bcutoff, grad_mask
        c.lr, c.beta1, c.beta2, c.eps = lr, betas[0], betas[1], eps
        self.cfg = c
        th = np.ascontiguousarray(np.asarray(theta0, np.float64).ravel())
        if th.size != P.N_THETA:
            raise ValueError("theta0 must have 1521 entries")
        tp = ctypes.c_void_p()
        rc = self.h.L.pinn_trainer_create(self.h.h, ctypes.byref(c), th.ctypes.data_as(ctypes.c_void_p), ctypes.byref(tp))
        self.h.check(rc, "pinn_trainer_create")
        self.t = tp
        self._step_base = 0

    def load_state(self, theta, m=None, v=None, step=0):
        """Resume: parameters, Adam moments (None = zeros) and the number of optimizer steps already done.  The resumed run
        starts its own best-model record (best_theta = the loaded parameters until a step qualifies) and its own history
        (row 0 = the first step after the resume); `freeze_after`, `best_after` and `sc_sampling` keep counting absolute steps."""
        a = lambda x: None if x is None else np.ascontiguousarray(np.asarray(x, np.float64).ravel())
        th, mm, vv = a(theta), a(m), a(v)
        p = lambda x: None if x is None else x.ctypes.data_as(ctypes.c_void_p)
        self.h.check(self.h.L.pinn_trainer_load_state(self.t, p(th), p(mm), p(vv), step), "pinn_trainer_load_state")
        self._step_base = int(step)

    def set_batch(self, x, y, z, R, mask, weights):
        """Use caller-provided points (CUDA or CPU float32 tensors) until the next resampling step."""
        cols = [t.detach().reshape(-1).to(torch.float32).contiguous() for t in (x, y, z, R)]
        mask = mask.detach().reshape(-1).to(torch.uint8).contiguous()
        w = np.ascontiguousarray(np.asarray(weights, np.float64))
        rc = self.h.L.pinn_trainer_set_batch(self.t, *[_ptr(c) for c in cols], _ptr(mask), w.ctypes.data_as(ctypes.c_void_p))
        self.h.check(rc, "pinn_trainer_set_batch")

    def run(self, steps, resample=True, use_graph=False):
        """Enqueue `steps` optimizer steps (asynchronous; read() synchronises).  A step is two launches (step kernel,
        reduction + Adam + next batch), chained as programmatic dependents; use_graph=True replays captured CUDA graphs
        instead, which costs the host less (1.5 vs 7 microseconds per step) but breaks the dependent-launch chain between
        replays (measured 147.6 vs 145.5 microseconds per step at 2^18 points)."""
        self.h.check(self.h.L.pinn_trainer_run(self.t, steps, int(resample), int(use_graph)), "pinn_trainer_run")

    def read(self, history_rows=None):
        f = lambda: np.empty(P.N_THETA, np.float64)
        theta, m, v, best = f(), f(), f(), f()
        sc = np.zeros(4, np.float64)
        rows = self.cfg.history_capacity if history_rows is None else min(history_rows, self.cfg.history_capacity)
        hist = np.zeros((max(rows, 0), 4), np.float64)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        rc = self.h.L.pinn_trainer_read(self.t, p(theta), p(m), p(v), p(best), p(sc), p(hist) if rows > 0 else None, rows)
        self.h.check(rc, "pinn_trainer_read")
        steps = int(sc[0])
        return {"theta": theta, "m": m, "v": v, "best_theta": best, "steps": steps, "best_loss": float(sc[1]),
                "best_step": int(sc[2]), "batches": int(sc[3]), "history": hist[:max(0, min(rows, steps - self._step_base))]}

    def close(self):
        if self.t:
            self.h.L.pinn_trainer_destroy(self.t)
            self.t = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def init_trainpy(seed=12345):
    """train.py's initialisation (ini(), train.py:13-18, 74, 88-103) -> packed theta (float64, canonical layout)."""
    g = torch.Generator().manual_seed(seed)
    nh, ne, nl = 16, 32, 10
    shapes = [(2, nh), (nh,), (nh, nh), (nh,), (nh, 1), (1,), (1, nl), (nl,), (nl, 1), (1,),
              (1, ne), (ne,), (ne, ne), (ne,), (ne, 1), (1,)]
    ts = []
    for s in shapes:
        t = torch.empty(s, dtype=torch.float64)
        lim = 1 / s[0] ** 0.5
        t.uniform_(-lim, lim, generator=g)
        ts.append(t)
    return P.pack_trainpy(ts, dtype=torch.float64).numpy()


def init_poc(seed=0):
    """A freshly constructed NN_ion (poc/main.py:233-245): `nn.Linear` default initialisation in float64, layers created in
    the reference's order, `Lin_Eout.bias = -1` -> packed theta (float64).  Equals the real class under
    `torch.manual_seed(seed)` (tests/golden/checkpoints.npz: init_seed0)."""
    nh, ne, nl = 16, 32, 10
    state = torch.get_rng_state()
    try:
        torch.manual_seed(seed)
        dims = [(2, nh), (nh, nh), (nh, 1), (1, ne), (ne, ne), (ne, 1), (1, nl), (nl, 1)]
        layers = [torch.nn.Linear(i, o, dtype=torch.float64) for i, o in dims]
    finally:
        torch.set_rng_state(state)
    with torch.no_grad():
        layers[5].bias.fill_(-1.0)
    ts = []
    for L in layers:
        ts += [L.weight, L.bias]
    return P.pack_poc(ts, dtype=torch.float64).numpy()


def train_trainpy(theta0=None, n=10000, epochs=1000, lr=8e-3, seed=12345, fine_tune=False, use_graph=False, log_every=0):
    """train(params, lr, epochs) of train.py on the device.  Like the reference it evaluates epochs+1 losses, takes `epochs`
    optimizer steps and returns the parameters of the best loss seen (train.py:58-69).  Returns (theta_best, info)."""
    if theta0 is None:
        theta0 = init_trainpy(seed)
    mask = P.FINE_TUNE_GRAD_MASK if fine_tune else 0xFFFF  # train.py:111 (commented) trains the E-net tensors only
    tr = Trainer("trainpy", n, theta0, seed=seed, lr=lr, grad_mask=mask, history_capacity=epochs + 1)
    done = 0
    while done < epochs:
        k = min(log_every or epochs, epochs - done)
        tr.run(k, use_graph=use_graph)
        done += k
        if log_every:
            r = tr.read()
            print("%8d: %.2e %.2e %.2e (%.2e) [%.5e]" % (done - 1, *r["history"][done - 1], r["best_loss"]))
    # the reference's loop evaluates the loss once more at tt == epochs before it stops (train.py:41-69): one more step
    # whose parameter update is discarded
    before = tr.read()
    tr.run(1, use_graph=use_graph)
    after = tr.read()
    best = after["best_theta"]
    info = {"history": after["history"], "best_loss": after["best_loss"], "best_step": after["best_step"],
            "theta_last": before["theta"], "batches": after["batches"]}
    tr.close()
    return best, info


def train_poc(theta0, params=None, freezeUnits=False, seed=0, use_graph=False):
    """train(params, loadWeights, freezeUnits) of poc/main.py:359-430 on the device.  `params` uses the reference's keys
    (set_params(), poc/main.py:14-45).  Returns (theta_last, theta_saved or None, lossDictionary like poc/main.py:422-427)."""
    pr = {"xL": -18, "xR": 18, "yL": -18, "yR": 18, "zL": -18, "zR": 18, "RxL": 0.2, "RxR": 4, "cutOff": 0.005,
          "BCcutoff": 17.5, "sc_sampling": 1, "n_train": 100000, "epochs": 5000, "lr": 8e-3}
    pr.update(params or {})
    P.check_supported_model(pr)
    epochs = int(pr["epochs"])
    box = (pr["xL"], pr["xR"], pr["yL"], pr["yR"], pr["zL"], pr["zR"], pr["RxL"], pr["RxR"])
    # resample while tt < 0.9*epochs (main.py:396); save when tt > 0.5*epochs and Ltot < Llim (main.py:414)
    freeze_after = int(np.ceil(0.9 * epochs))
    best_after = int(np.floor(0.5 * epochs))
    tr = Trainer("poc", int(pr["n_train"]), theta0, seed=seed, lr=pr["lr"], box=box, cutoff=pr["cutOff"],
                 bcutoff=pr["BCcutoff"], grad_mask=P.FINE_TUNE_GRAD_MASK if freezeUnits else 0xFFFF, best_mode=1,
                 best_after=best_after, sc_sampling=int(pr["sc_sampling"]), freeze_after=freeze_after,
                 history_capacity=epochs, history_mean_E=False)
    tr.run(epochs, use_graph=use_graph)
    r = tr.read()
    tr.close()
    hcol = lambda k: r["history"][:, k:k + 1].copy()
    loss = {"Ltot": hcol(0), "Lpde": hcol(1), "Lbc": hcol(2), "Energy": hcol(3)}
    saved = r["best_theta"] if r["best_step"] >= 0 else None
    return r["theta"], saved, loss
