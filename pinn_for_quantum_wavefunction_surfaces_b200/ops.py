"""Python mirror of the reference's hot-path call sites on top of the C ABI.

    loss_poc(...)      <->  NN_ion.LossFunctions(x,y,z,R,params,bIndex1,bIndex2)   poc/main.py:341-355
    loss_trainpy(...)  <->  the inline block train.py:41-57
    fields(...)        <->  parametricPsi + hamiltonian                             poc/main.py:321, 118, 451-454

The parameter gradient is produced by the same fused kernel launch as the loss (the upstream
gradient of a scalar loss is a scalar), stashed on the autograd context, and ``backward`` only
scales it - no second kernel, no saved activations.
"""
import ctypes

import torch

from . import params as P
from ._lib import Handle, PinnError

VARIANT_POC, VARIANT_TRAINPY = 0, 1
_F32, _F64 = 0, 1
BC_CUTOFF = 17.5  # params['BCcutoff'] poc/main.py:25 ; bcutoff train.py:78


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _col(t, name):
    """(n,1) or (n,) float32/float64 tensor -> contiguous 1-D view + dtype code."""
    if t.dim() == 2 and t.shape[1] == 1:
        t = t[:, 0]
    if t.dim() != 1:
        raise ValueError("%s must have shape (n,) or (n,1), got %s" % (name, tuple(t.shape)))
    if t.dtype not in (torch.float32, torch.float64):
        raise TypeError("%s must be float32 or float64" % name)
    return t.detach().contiguous()


def indices_to_mask(n, idx1, idx2, device):
    """Boundary index sets (torch.where tuples as in poc/main.py:392-393, or 1-D row indices as in
    train.py:38-39) -> per-point uint8 mask (bit0: set 1, bit1: set 2) and the two set sizes."""
    def rows(ix):
        if isinstance(ix, (tuple, list)):
            ix = ix[0]
        return ix.to(device=device, dtype=torch.long).reshape(-1)
    r1, r2 = rows(idx1), rows(idx2)
    mask = torch.zeros(n, dtype=torch.uint8, device=device)
    m2 = torch.zeros(n, dtype=torch.uint8, device=device)
    mask[r1] = 1
    m2[r2] = 2
    mask |= m2
    return mask, int(r1.numel()), int(r2.numel())


def loss_and_grad_raw(variant, x, y, z, R, theta, mask=None, weights=None, grad_mask=0xFFFF, want_E=False,
                      bcutoff=BC_CUTOFF, sums=None, dtheta=None, E_out=None):
    """Thin wrapper over pinn_loss_fwd_bwd for DEVICE tensors (no allocation when the outputs are given).

    x,y,z,R: CUDA float32/float64 (n,) ; theta: CUDA float32 (1521,) ; mask: CUDA uint8 (n,) or None ;
    weights: CUDA float64 (3,) or None.  Returns (sums[8] f64, dtheta[1521] f64, E_out or None), all on
    the device, enqueued on torch's current stream."""
    if not x.is_cuda:
        raise PinnError("loss_and_grad_raw needs CUDA tensors (there is no CPU path)")
    dev = x.device
    h = Handle.get(dev.index if dev.index is not None else torch.cuda.current_device())
    n = x.numel()
    code = _F64 if x.dtype == torch.float64 else _F32
    for t in (y, z, R):
        if t.dtype != x.dtype or t.numel() != n or not t.is_contiguous():
            raise ValueError("x,y,z,R must share dtype/length and be contiguous")
    if theta.dtype != torch.float32 or theta.numel() != P.N_THETA or not theta.is_contiguous():
        raise ValueError("theta must be a contiguous float32 vector of 1521 elements")
    if sums is None:
        sums = torch.empty(8, dtype=torch.float64, device=dev)
    if dtheta is None:
        dtheta = torch.empty(P.N_THETA, dtype=torch.float64, device=dev)
    if want_E and E_out is None:
        E_out = torch.empty(n, dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    rc = h.L.pinn_loss_fwd_bwd(h.h, int(variant), n, _ptr(x), _ptr(y), _ptr(z), _ptr(R), code, _ptr(mask),
                               _ptr(theta), _ptr(weights), int(grad_mask), float(bcutoff), _ptr(sums),
                               _ptr(dtheta), _ptr(E_out), ctypes.c_void_p(stream))
    h.check(rc, "pinn_loss_fwd_bwd")
    return sums, dtheta, E_out


def _host_step(variant, x, y, z, R, theta64, mask, weights, grad_mask, want_E, bcutoff):
    """CPU tensors in, CPU float64 results out, through pinn_loss_fwd_bwd_host on cuda:current."""
    if not torch.cuda.is_available():
        raise PinnError("no CUDA device: the PINN hot path has no CPU fallback")
    h = Handle.get(torch.cuda.current_device())
    n = x.numel()
    code = _F64 if x.dtype == torch.float64 else _F32
    sums = torch.empty(8, dtype=torch.float64)
    dtheta = torch.empty(P.N_THETA, dtype=torch.float64)
    E_out = torch.empty(n, dtype=torch.float32) if want_E else None
    w = (ctypes.c_double * 3)(*weights) if weights is not None else None
    rc = h.L.pinn_loss_fwd_bwd_host(h.h, int(variant), n, _ptr(x), _ptr(y), _ptr(z), _ptr(R), code, _ptr(mask),
                                    _ptr(theta64), ctypes.cast(w, ctypes.c_void_p) if w is not None else None,
                                    int(grad_mask), float(bcutoff), _ptr(sums), _ptr(dtheta), _ptr(E_out))
    h.check(rc, "pinn_loss_fwd_bwd_host")
    return sums, dtheta, E_out


class HostStep:
    """A training evaluation on one fixed set of HOST buffers, prepared once and called many times
    (``pinn_loss_fwd_bwd_host``).  Binding the 16 ctypes arguments costs ~15 microseconds in Python, an order of
    magnitude more than the library needs to enqueue the work; a loop that reuses its batch buffers (the reference's
    training loops do) prepares the call once per buffer set.  Pin the tensors (``tensor.pin_memory()``) to let the
    kernel read them in place.

        step = HostStep("poc", x, y, z, R)            # CPU float32/float64 tensors, (n,) or (n,1); optional uint8 mask
        sums, dtheta = step(theta64, weights)         # numpy float64 (1521,), (3,) -> views of reused float64 outputs
    """

    def __init__(self, variant, x, y, z, R, mask=None, grad_mask=0xFFFF, bcutoff=BC_CUTOFF, device=None):
        if not torch.cuda.is_available():
            raise PinnError("no CUDA device: the PINN hot path has no CPU fallback")
        import numpy as np
        self.h = Handle.get(torch.cuda.current_device() if device is None else device)
        self._keep = [_col(t, nm) for t, nm in ((x, "x"), (y, "y"), (z, "z"), (R, "R"))]
        if any(t.is_cuda for t in self._keep):
            raise ValueError("HostStep is for host tensors; use loss_and_grad_raw for CUDA tensors")
        n = self._keep[0].numel()
        for t in self._keep[1:]:
            if t.dtype != self._keep[0].dtype or t.numel() != n:
                raise ValueError("x,y,z,R must share dtype and length")
        self._mask = None if mask is None else mask.detach().reshape(-1).to(torch.uint8).contiguous()
        self.sums = np.zeros(8)
        self.dtheta = np.zeros(P.N_THETA)
        vp = ctypes.c_void_p
        self._fn = self.h.L.pinn_loss_fwd_bwd_host
        self._head = (self.h.h, int({"poc": 0, "trainpy": 1}.get(variant, variant)), n) + tuple(vp(t.data_ptr()) for t in self._keep) + (
            _F64 if self._keep[0].dtype == torch.float64 else _F32, None if self._mask is None else vp(self._mask.data_ptr()))
        self._th = self._w = self._thp = self._wp = None
        self._tail = (int(grad_mask), float(bcutoff), vp(self.sums.ctypes.data), vp(self.dtheta.ctypes.data), None)

    def __call__(self, theta64, weights=None):
        """theta64: contiguous numpy float64 (1521,); weights: contiguous numpy float64 (3,) {1/n, 1/|set1|, 1/|set2|} or
        None (sets counted on the device).  Returns (sums[8], dtheta[1521]); the arrays are overwritten by the next call."""
        if theta64 is not self._th:       # `.ctypes.data` is slow; the usual caller passes the same arrays every step
            self._th, self._thp = theta64, theta64.ctypes.data
        if weights is not self._w:
            self._w, self._wp = weights, (None if weights is None else weights.ctypes.data)
        rc = self._fn(*self._head, self._thp, self._wp, *self._tail)
        if rc:
            self.h.check(rc, "pinn_loss_fwd_bwd_host")
        return self.sums, self.dtheta


class _PinnLoss(torch.autograd.Function):
    """forward(variant, order, x, y, z, R, idx1, idx2, *16 params) -> (Ltot, Lpde, Lbc, E)."""

    @staticmethod
    def forward(ctx, variant, order, x, y, z, R, idx1, idx2, *params):
        if len(params) != 16:
            raise ValueError("expected the model's 16 parameter tensors")
        xs, ys, zs, Rs = _col(x, "x"), _col(y, "y"), _col(z, "z"), _col(R, "R")
        n = xs.numel()
        dev = xs.device
        pdt = params[0].dtype
        grad_mask = P.grad_mask_from_requires_grad(params, order)
        ctx.order, ctx.shapes, ctx.needs = order, [tuple(p.shape) for p in params], [p.requires_grad for p in params]
        mask, c1, c2 = indices_to_mask(n, idx1, idx2, dev)
        if c1 == 0 or c2 == 0:
            # the reference takes the mean of an empty selection -> NaN loss (hard part 6)
            weights = [1.0 / n, float("inf") if c1 == 0 else 1.0 / c1, float("inf") if c2 == 0 else 1.0 / c2]
        else:
            weights = [1.0 / n, 1.0 / c1, 1.0 / c2]
        pack = P.pack_poc if order == "poc" else P.pack_trainpy
        if dev.type == "cuda":
            theta = pack(params, dtype=torch.float32, device=dev)
            wdev = torch.tensor(weights, dtype=torch.float64, device=dev)
            sums, dtheta, E = loss_and_grad_raw(variant, xs, ys, zs, Rs, theta, mask, wdev, grad_mask, True)
        else:
            theta64 = pack(params, dtype=torch.float64).contiguous()
            sums, dtheta, E = _host_step(variant, xs, ys, zs, Rs, theta64, mask, weights, grad_mask, True, BC_CUTOFF)
        ctx.save_for_backward(dtheta)
        Ltot, Lpde, Lbc = sums[0].to(pdt), sums[1].to(pdt), sums[2].to(pdt)
        E = E.to(pdt).reshape(n, 1)
        ctx.mark_non_differentiable(Lpde, Lbc, E)
        return Ltot, Lpde, Lbc, E

    @staticmethod
    def backward(ctx, gLtot, gLpde, gLbc, gE):
        (dtheta,) = ctx.saved_tensors
        g = dtheta * gLtot.to(dtheta.dtype)
        parts = P.unpack_poc(g) if ctx.order == "poc" else P.unpack_trainpy(g)
        outs = []
        for part, shp, need in zip(parts, ctx.shapes, ctx.needs):
            outs.append(part.reshape(shp) if need else None)
        return (None,) * 8 + tuple(outs)


class PinnLossPoc:
    """Fused replacement of NN_ion.LossFunctions (poc/main.py:341-355); params in state_dict order."""

    @staticmethod
    def apply(x, y, z, R, bIndex1, bIndex2, *params):
        return _PinnLoss.apply(VARIANT_POC, "poc", x, y, z, R, bIndex1, bIndex2, *params)


class PinnLossTrainPy:
    """Fused replacement of train.py:41-57; params in train.py's tuple order/(in,out) layout (train.py:108-109)."""

    @staticmethod
    def apply(x, y, z, R, i1, i2, *params):
        return _PinnLoss.apply(VARIANT_TRAINPY, "trainpy", x, y, z, R, i1, i2, *params)


def loss_poc(model, x, y, z, R, bIndex1, bIndex2):
    """Ltot, LossPDE, Lbc, E for an ``NN_ion``-like module (16 parameters in state_dict order)."""
    return PinnLossPoc.apply(x, y, z, R, bIndex1, bIndex2, *list(model.parameters()))


def loss_trainpy(x, y, z, R, i1, i2, *params):
    """Ltot, Lpde, Lbc, e for train.py's 16 parameter tensors."""
    return PinnLossTrainPy.apply(x, y, z, R, i1, i2, *params)


def fields(variant, x, y, z, R, theta, want=("psi", "lap", "hpsi", "res", "E")):
    """Fused parametricPsi + hamiltonian (poc/main.py:321, 118): per-point psi, laplacian, H psi,
    residual and E as float32 CUDA tensors.  x,y,z,R: CUDA float32/float64 (n,) or (n,1); theta: packed
    parameters (any float dtype/device, converted to CUDA float32)."""
    variant = {"poc": 0, "trainpy": 1}.get(variant, variant)
    xs, ys, zs, Rs = _col(x, "x"), _col(y, "y"), _col(z, "z"), _col(R, "R")
    if not xs.is_cuda:
        if not torch.cuda.is_available():
            raise PinnError("no CUDA device: the PINN hot path has no CPU fallback")
        xs, ys, zs, Rs = (t.cuda() for t in (xs, ys, zs, Rs))
    dev = xs.device
    h = Handle.get(dev.index if dev.index is not None else torch.cuda.current_device())
    n = xs.numel()
    theta = theta.detach().to(device=dev, dtype=torch.float32).contiguous()
    out = {k: torch.empty(n, dtype=torch.float32, device=dev) for k in want}
    code = _F64 if xs.dtype == torch.float64 else _F32
    stream = torch.cuda.current_stream(dev).cuda_stream
    rc = h.L.pinn_fields(h.h, int(variant), n, _ptr(xs), _ptr(ys), _ptr(zs), _ptr(Rs), code, _ptr(theta),
                         _ptr(out.get("psi")), _ptr(out.get("lap")), _ptr(out.get("hpsi")), _ptr(out.get("res")),
                         _ptr(out.get("E")), ctypes.c_void_p(stream))
    h.check(rc, "pinn_fields")
    return out
