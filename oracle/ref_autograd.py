"""Reference-style autograd oracle (TEST INFRASTRUCTURE ONLY; parity pinned, see tests/test_oracle.py).

Restates the way the reference evaluates the PINN loss: a plain PyTorch forward of
the model followed by a Laplacian assembled from nested ``torch.autograd.grad``
calls with ``create_graph=True``, then ``backward()`` for the parameter gradients.

    poc form      NN_ion.forward / atomicUnit / base / lcao_solution  poc/main.py:247-303
                  dfx, d2fx, lapl                                     poc/main.py:82-97
                  radial, V, hamiltonian                              poc/main.py:101-120
                  NN_ion.LossFunctions                                poc/main.py:341-355
    train.py form inline forward / d() / residual / loss             train.py:8-10, 41-57

Parameters are one packed float64 vector ``theta`` (layout.py) so the gradient is one
vector.  Everything is float64 on the CPU, like the reference.
"""
import torch

from . import layout

_OFF = layout.offsets()


def split(theta):
    """flat (1521,) tensor -> 16 views in canonical (out,in) layout."""
    out = []
    for off, (_, shp) in zip(_OFF, layout.POC_TENSORS):
        n = 1
        for s in shp:
            n *= s
        out.append(theta[off:off + n].reshape(shp))
    return out


def _second(psi, c):
    # d2fx: differentiate twice w.r.t. one coordinate (poc/main.py:82-91; train.py:8-10)
    ones = torch.ones_like(c)
    g1 = torch.autograd.grad([psi], [c], grad_outputs=ones, create_graph=True)[0]
    return torch.autograd.grad([g1], [c], grad_outputs=ones, create_graph=True)[0]


def laplacian(psi, x, y, z):
    # lapl (poc/main.py:94-97)
    return _second(psi, x) + _second(psi, y) + _second(psi, z)


def _radii(x, y, z, R):
    # radial (poc/main.py:101-108) with Ry = Rz = 0 (poc/main.py:28-29)
    r1 = torch.sqrt((x - R) ** 2 + y ** 2 + z ** 2)
    r2 = torch.sqrt((x + R) ** 2 + y ** 2 + z ** 2)
    return r1, r2


def _mlp2(a, b, W1, b1, W2, b2):
    # NN_ion.base (poc/main.py:295-303)
    h = torch.sigmoid(torch.cat((a, b), 1) @ W1.T + b1)
    return torch.sigmoid(h @ W2.T + b2)


def _enet(R, WE1, bE1, WE2, bE2, wE, bE):
    # E(R) branch (poc/main.py:249-253; train.py:50-52)
    e = torch.sigmoid(R @ WE1.T + bE1)
    e = torch.sigmoid(e @ WE2.T + bE2)
    return e @ wE.T + bE


def _gate(R, WgL, bgL, wg, bg):
    # netDecayL -> sig -> netDecay (poc/main.py:262-264; train.py:48-49)
    return torch.sigmoid(R @ WgL.T + bgL) @ wg.T + bg


def poc_psi_E(theta, x, y, z, R):
    """NN_ion.forward (poc/main.py:247-267): returns (psi, E), each (n,1)."""
    W1, b1, W2, b2, wo, bo, WE1, bE1, WE2, bE2, wE, bE, WgL, bgL, wg, bg = split(theta)
    E = _enet(R, WE1, bE1, WE2, bE2, wE, bE)
    r1, r2 = _radii(x, y, z, R)
    f1, f2 = torch.exp(-r1), torch.exp(-r2)
    # atomicUnit(-x, ...) swaps the two orbitals because Ry=Rz=0 (poc/main.py:255-256)
    B = _mlp2(f1, f2, W1, b1, W2, b2) + _mlp2(f2, f1, W1, b1, W2, b2)
    N = B @ wo.T + bo
    psi = N * _gate(R, WgL, bgL, wg, bg) + (f1 + f2)
    return psi, E


def poc_fields(theta, x, y, z, R):
    """psi, laplacian, H psi, residual, E per point (poc/main.py:118-120, 343-345)."""
    x = x.detach().clone().requires_grad_(True)
    y = y.detach().clone().requires_grad_(True)
    z = z.detach().clone().requires_grad_(True)
    psi, E = poc_psi_E(theta, x, y, z, R)
    lap = laplacian(psi, x, y, z)
    r1, r2 = _radii(x, y, z, R)
    hpsi = -0.5 * lap + (-1 / r1 - 1 / r2) * psi
    res = hpsi - E * psi
    return psi, lap, hpsi, res, E


def poc_loss(theta, x, y, z, R, idx1, idx2):
    """NN_ion.LossFunctions (poc/main.py:341-355): (Ltot, Lpde, Lbc, E).

    idx1/idx2 are 1-D row-index tensors (the reference passes torch.where tuples on
    an (n,1) tensor, which select the same rows)."""
    psi, _, _, res, E = poc_fields(theta, x, y, z, R)
    Lpde = (res ** 2).mean()
    Lbc = (psi[idx1, 0] ** 2).mean() + (psi[idx2, 0] ** 2).mean()
    return Lpde + Lbc, Lpde, Lbc, E


def trainpy_psi_e(theta, x, y, z, R):
    """train.py:41-53 with the canonical (out,in) weights: returns (psi, e)."""
    W1, b1, W2, b2, wo, bo, WE1, bE1, WE2, bE2, wE, bE, WgL, bgL, wg, bg = split(theta)
    r1, r2 = _radii(x, y, z, R)
    f1, f2 = torch.exp(-r1), torch.exp(-r2)
    h = _mlp2(f1, f2, W1, b1, W2, b2)
    N = (2 * h) @ wo.T + bo
    psi = f1 + f2 + N * _gate(R, WgL, bgL, wg, bg)
    return psi, _enet(R, WE1, bE1, WE2, bE2, wE, bE)


def trainpy_fields(theta, x, y, z, R):
    """psi, laplacian, residual (train.py:54), e per point."""
    x = x.detach().clone().requires_grad_(True)
    y = y.detach().clone().requires_grad_(True)
    z = z.detach().clone().requires_grad_(True)
    psi, e = trainpy_psi_e(theta, x, y, z, R)
    lap = laplacian(psi, x, y, z)
    r1, r2 = _radii(x, y, z, R)
    res = lap + (e + 1 / r1 + 1 / r2) * psi
    return psi, lap, res, e


def trainpy_loss(theta, x, y, z, R, i1, i2):
    """train.py:55-57: (Ltot, Lpde, Lbc, e)."""
    psi, _, res, e = trainpy_fields(theta, x, y, z, R)
    Lpde = torch.mean(res ** 2)
    Lbc = torch.mean(psi[i1, 0] ** 2) + torch.mean(psi[i2, 0] ** 2)
    return Lpde + Lbc, Lpde, Lbc, e


def loss_and_grad(variant, theta, x, y, z, R, i1, i2):
    """One reference-style training evaluation: loss terms + dLtot/dtheta (float64)."""
    theta = theta.detach().clone().requires_grad_(True)
    fn = poc_loss if variant == "poc" else trainpy_loss
    Ltot, Lpde, Lbc, E = fn(theta, x, y, z, R, i1, i2)
    Ltot.backward()
    return Ltot.detach(), Lpde.detach(), Lbc.detach(), E.detach(), theta.grad.detach()


def sample_box(n, variant="poc", generator=None, L=18.0, cutoff=0.005, bcutoff=17.5):
    """Collocation sampler + clamp + boundary index sets (poc/main.py:124-156, 390-393;
    train.py:26-39).  Returns x,y,z,R (n,1) float64 and the two row-index tensors."""
    Rlo, Rhi = (0.2, 4.0) if variant == "poc" else (0.2, 3.0)
    u = torch.rand(n, 4, dtype=torch.float64, generator=generator)
    x = (2 * u[:, 0:1] - 1) * L
    y = (2 * u[:, 1:2] - 1) * L
    z = (2 * u[:, 2:3] - 1) * L
    R = Rlo + (Rhi - Rlo) * u[:, 3:4]
    r1, r2 = _radii(x, y, z, R)
    x[r1 < cutoff] = cutoff
    x[r2 < cutoff] = cutoff
    r1, r2 = _radii(x, y, z, R)
    i1 = torch.where(r1[:, 0] >= bcutoff)[0]
    i2 = torch.where(r2[:, 0] >= bcutoff)[0]
    return x, y, z, R, i1, i2
