"""The reference's training drivers restated around a pluggable loss (TEST INFRASTRUCTURE ONLY).

    trainpy_run   train.py:13-18 (ini), 21-72 (train), 74-110 (script body)    - pinned: with the float64 autograd
                  oracle as the loss it reproduces the golden trace and model.bin of the real script
                  (tests/golden/trainpy_trace_n4096_e40.json, made by tests/golden/make_golden.py)
    adam_reference  torch.optim.Adam itself (the reference's optimizer) driven on a flat float64 vector

`loss_op(x, y, z, R, i1, i2, *params) -> (Ltot, Lpde, Lbc, e)` has the signature of the fused op
(PinnLossTrainPy.apply); the CPU oracle version is `oracle_trainpy_op`.
"""
import numpy as np
import torch

from . import closed_form as cf
from . import layout


def oracle_trainpy_op(x, y, z, R, i1, i2, *params):
    """train.py:41-57 through the float64 closed-form oracle (== the nested-autograd restatement to 1e-15,
    tests/test_oracle.py), as an autograd.Function of the 16 tensors"""
    class F(torch.autograd.Function):
        @staticmethod
        def forward(ctx, *ps):
            theta = layout.from_trainpy([p.detach().numpy() for p in ps])
            n = x.shape[0]
            m1 = np.zeros(n); m1[i1.numpy()] = 1
            m2 = np.zeros(n); m2[i2.numpy()] = 1
            o = cf.loss_and_grad("trainpy", theta, *[v.detach().numpy().ravel() for v in (x, y, z, R)], m1, m2)
            ctx.g = [torch.tensor(a) for a in layout.to_trainpy(o["grad"])]
            t = lambda v: torch.tensor(v, dtype=torch.float64)
            outs = (t(o["Ltot"]), t(o["Lpde"]), t(o["Lbc"]), t(o["E"]).reshape(-1, 1))
            ctx.mark_non_differentiable(*outs[1:])
            return outs

        @staticmethod
        def backward(ctx, gL, *_):
            return tuple(gL * a for a in ctx.g)
    return F.apply(*params)


def trainpy_run(loss_op, n=10000, epochs=1000, lr=8e-3, seed=12345, log=None, snapshots=None):
    """Returns (params tuple in train.py layout with the best parameters restored, trace lines, history array).
    snapshots: optional dict, filled with {tt: canonical packed parameters the loss of step tt was evaluated with} for
    every tt it already has as a key (drift curves of a float32 loss against the float64 one)."""
    dtype = torch.double
    torch.manual_seed(seed)                                   # train.py:74
    bcutoff, cutoff, L, Rlo, Rhi, sc_sampling = 17.5, 0.005, 18, 0.2, 3, 1   # train.py:78-84
    nh, ne, nl = 16, 32, 10

    def ini(*shape):                                           # train.py:13-18
        t = torch.empty(shape, dtype=dtype, requires_grad=True)
        lim = 1 / shape[0] ** 0.5
        with torch.no_grad():
            t.uniform_(-lim, lim)
        return t
    shapes = [(2, nh), (nh,), (nh, nh), (nh,), (nh, 1), (1,), (1, nl), (nl,), (nl, 1), (1,),
              (1, ne), (ne,), (ne, ne), (ne,), (ne, 1), (1,)]    # train.py:88-103, same creation order
    params = tuple(ini(*s) for s in shapes)
    x = torch.empty(n, 1, dtype=dtype)
    y = torch.empty(n, 1, dtype=dtype)
    z = torch.empty(n, 1, dtype=dtype)
    R = torch.empty(n, 1, dtype=dtype)
    opt = torch.optim.Adam(params, lr=lr)                      # train.py:23
    trace, hist = [], []
    tt, Lbest, best = 0, None, None
    while True:                                                # train.py:24-72
        opt.zero_grad()
        if tt % sc_sampling == 0:
            with torch.no_grad():
                x.uniform_(-L, L); y.uniform_(-L, L); z.uniform_(-L, L); R.uniform_(Rlo, Rhi)
                r1sq = (x - R) ** 2 + y ** 2 + z ** 2
                r2sq = (x + R) ** 2 + y ** 2 + z ** 2
                x[r1sq < cutoff ** 2] = cutoff
                x[r2sq < cutoff ** 2] = cutoff
                r1sq = (x - R) ** 2 + y ** 2 + z ** 2
                r2sq = (x + R) ** 2 + y ** 2 + z ** 2
                i1, = torch.where(r1sq[:, 0] >= bcutoff ** 2)
                i2, = torch.where(r2sq[:, 0] >= bcutoff ** 2)
        if snapshots is not None and tt in snapshots:
            snapshots[tt] = layout.from_trainpy([p.detach().numpy().copy() for p in params])
        Ltot, Lpde, Lbc, e = loss_op(x, y, z, R, i1, i2, *params)
        lt = float(Ltot.detach())
        if tt == 0 or lt < Lbest:
            Lbest = lt
            best = [p.clone().detach() for p in params]
        hist.append([lt, float(Lpde.detach()), float(Lbc.detach()), float(torch.mean(e.detach()))])
        if tt % 10 == 0:
            trace.append("%d: %.2e %.2e %.2e (%.2e) [%.5e]" % (tt, *hist[-1], Lbest))
            if log:
                log(trace[-1])
        if tt == epochs:
            with torch.no_grad():
                for a, b in zip(params, best):
                    a.copy_(b)
            break
        tt += 1
        Ltot.backward()
        opt.step()
    return params, trace, np.array(hist)


def model_bin_bytes(params):
    """train.py:112-119"""
    out = b""
    for p in params:
        a = p.detach().numpy()
        out += a.ndim.to_bytes(4, "little")
        for d in a.shape:
            out += int(d).to_bytes(4, "little")
        out += a.tobytes()
    return out


def adam_reference(theta0, grads, lr=8e-3, betas=(0.9, 0.999), eps=1e-8, frozen=None):
    """torch.optim.Adam on a flat float64 vector for a list of gradient vectors; `frozen` = boolean mask of entries
    whose gradient is None in the reference (frozen tensors are skipped by the optimizer)."""
    th = torch.tensor(np.asarray(theta0, np.float64))
    if frozen is None:
        p = [th.clone().requires_grad_(True)]
        opt = torch.optim.Adam(p, lr=lr, betas=betas, eps=eps)
        out = []
        for g in grads:
            p[0].grad = torch.tensor(np.asarray(g, np.float64))
            opt.step()
            out.append(p[0].detach().numpy().copy())
        return out
    frozen = np.asarray(frozen, bool)
    act = torch.tensor(np.asarray(theta0, np.float64)[~frozen]).requires_grad_(True)
    opt = torch.optim.Adam([act], lr=lr, betas=betas, eps=eps)
    out = []
    for g in grads:
        act.grad = torch.tensor(np.asarray(g, np.float64)[~frozen])
        opt.step()
        full = np.asarray(theta0, np.float64).copy()
        full[~frozen] = act.detach().numpy()
        out.append(full)
    return out
