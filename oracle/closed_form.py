"""Closed-form oracle: the specification of the CUDA kernels (TEST INFRASTRUCTURE ONLY).

numpy float64 restatement of SURVEY.md Appendix A.3.  The network sees (x,y,z) only
through a = exp(-r1), b = exp(-r2), so the Laplacian of psi follows from a *single*
second-order operator carried forward through the MLP together with the value and
the two first derivatives (4 channels: v, d/da, d/db, D):

    D = al1 d/da + al2 d/db + al11 d2/da2 + al12 d2/dadb + al22 d2/db2
    al1 = lap f1 = f1 (1 - 2/r1)      al11 = |grad f1|^2 = f1^2
    al2 = lap f2 = f2 (1 - 2/r2)      al22 = f2^2        al12 = 2 f1 f2 cos(r1,r2)
    lap psi = al1 + al2 + g(R) * D[N]                 (poc/main.py:94-97, 258-266)

and for a composition  D[sig(u)] = sig'(u) D[u] + sig''(u) Q[u],
Q[u] = al11 u_a^2 + al12 u_a u_b + al22 u_b^2.

The residual is written in the cusp-cancelled form, i.e. the LCAO part of
-1/2 lap psi + V psi is simplified analytically:

    res = cL (f1+f2) + (cV - 2 cL)(f1/r1 + f2/r2) + cV (f1/r2 + f2/r1)
          + g (cL D[N] + cV q N) + cE E psi,          q = 1/r1 + 1/r2
    poc      (cL,cV,cE) = (-1/2,-1,-1)   poc/main.py:114,120,345   two MLP evaluations summed
    train.py (cL,cV,cE) = ( 1 , 1 , 1)   train.py:54               one evaluation, doubled (train.py:47)

``loss_and_grad`` also contains the hand-written reverse sweep that the CUDA kernel
implements; tests/test_oracle.py checks it against ref_autograd (nested autograd) and
against golden outputs of the real reference.
"""
import numpy as np

from . import layout

VARIANTS = {
    # name: (number of MLP evaluations, output scale, cL, cV, cE)
    "poc": (2, 1.0, -0.5, -1.0, -1.0),
    "trainpy": (1, 2.0, 1.0, 1.0, 1.0),
}


def _sig(u):
    return 1.0 / (1.0 + np.exp(-u))


def geometry(x, y, z, R):
    """Per-point quantities shared by every evaluation (1-D float64 arrays)."""
    dx1, dx2 = x - R, x + R
    yz = y * y + z * z
    r1 = np.sqrt(dx1 * dx1 + yz)
    r2 = np.sqrt(dx2 * dx2 + yz)
    f1, f2 = np.exp(-r1), np.exp(-r2)
    ir1, ir2 = 1.0 / r1, 1.0 / r2
    c12 = (dx1 * dx2 + yz) * ir1 * ir2
    return dict(r1=r1, r2=r2, f1=f1, f2=f2, ir1=ir1, ir2=ir2,
                al1=f1 * (1 - 2 * ir1), al2=f2 * (1 - 2 * ir2),
                al11=f1 * f1, al22=f2 * f2, al12=2 * f1 * f2 * c12)


def mlp_fwd(P, a, b, al1, al2, al11, al12, al22):
    """4-channel Taylor forward of wo . sig(W2 sig(W1 [a,b] + b1) + b2)."""
    W1, b1, W2, b2, wo = P[0], P[1], P[2], P[3], P[4].reshape(-1)
    w0, w1 = W1[:, 0], W1[:, 1]
    c = lambda v: v[:, None]
    u = c(a) * w0 + c(b) * w1 + b1
    s = _sig(u)
    sp = s * (1 - s)
    spp = sp * (1 - 2 * s)
    sppp = sp * (1 - 6 * sp)
    d1 = c(al1) * w0 + c(al2) * w1
    q = c(al11) * w0 * w0 + c(al12) * w0 * w1 + c(al22) * w1 * w1
    h, ha, hb, hD = s, sp * w0, sp * w1, sp * d1 + spp * q
    v, va, vb, vD = h @ W2.T + b2, ha @ W2.T, hb @ W2.T, hD @ W2.T
    t = _sig(v)
    tp = t * (1 - t)
    tpp = tp * (1 - 2 * t)
    tppp = tp * (1 - 6 * tp)
    Q = c(al11) * va * va + c(al12) * va * vb + c(al22) * vb * vb
    gD = tp * vD + tpp * Q
    st = dict(a=a, b=b, al=(al1, al2, al11, al12, al22), sp=sp, spp=spp, sppp=sppp, d1=d1, q=q,
              h=h, ha=ha, hb=hb, hD=hD, va=va, vb=vb, vD=vD, t=t, tp=tp, tpp=tpp, tppp=tppp, Q=Q, gD=gD)
    return t @ wo, gD @ wo, st


def mlp_bwd(P, st, lamN, lamD):
    """Reverse sweep of mlp_fwd for the seeds lamN = dL/dNv, lamD = dL/dDv (per point).
    Returns gradients of (W1, b1, W2, b2, wo)."""
    W1, W2, wo = P[0], P[2], P[4].reshape(-1)
    w0, w1 = W1[:, 0], W1[:, 1]
    c = lambda v: v[:, None]
    al1, al2, al11, al12, al22 = st["al"]
    tbar, gDbar = c(lamN) * wo, c(lamD) * wo
    dwo = (c(lamN) * st["t"] + c(lamD) * st["gD"]).sum(0)
    vDbar = gDbar * st["tp"]
    vabar = gDbar * st["tpp"] * (2 * c(al11) * st["va"] + c(al12) * st["vb"])
    vbbar = gDbar * st["tpp"] * (c(al12) * st["va"] + 2 * c(al22) * st["vb"])
    vbar = tbar * st["tp"] + gDbar * (st["tpp"] * st["vD"] + st["tppp"] * st["Q"])
    dW2 = vbar.T @ st["h"] + vabar.T @ st["ha"] + vbbar.T @ st["hb"] + vDbar.T @ st["hD"]
    db2 = vbar.sum(0)
    hbar, habar, hbbar, hDbar = vbar @ W2, vabar @ W2, vbbar @ W2, vDbar @ W2
    sp, spp, sppp = st["sp"], st["spp"], st["sppp"]
    ubar = hbar * sp + (habar * w0 + hbbar * w1 + hDbar * st["d1"]) * spp + hDbar * st["q"] * sppp
    dw0 = (habar * sp + hDbar * (sp * c(al1) + spp * (2 * c(al11) * w0 + c(al12) * w1)) + ubar * c(st["a"])).sum(0)
    dw1 = (hbbar * sp + hDbar * (sp * c(al2) + spp * (c(al12) * w0 + 2 * c(al22) * w1)) + ubar * c(st["b"])).sum(0)
    db1 = ubar.sum(0)
    return [np.stack([dw0, dw1], 1), db1, dW2, db2, dwo.reshape(1, -1)]


def enet_fwd(P, R):
    WE1, bE1, WE2, bE2, wE, bE = P[6].reshape(-1), P[7], P[8], P[9], P[10].reshape(-1), P[11]
    e1 = _sig(R[:, None] * WE1 + bE1)
    e2 = _sig(e1 @ WE2.T + bE2)
    return e2 @ wE + bE[0], (e1, e2)


def enet_bwd(P, R, st, Ebar):
    WE2, wE = P[8], P[10].reshape(-1)
    e1, e2 = st
    c = lambda v: v[:, None]
    dwE = (c(Ebar) * e2).sum(0)
    vbar = c(Ebar) * wE * e2 * (1 - e2)
    dWE2 = vbar.T @ e1
    ubar = (vbar @ WE2) * e1 * (1 - e1)
    return [(ubar * c(R)).sum(0).reshape(-1, 1), ubar.sum(0), dWE2, vbar.sum(0), dwE.reshape(1, -1),
            np.array([Ebar.sum()])]


def gate_fwd(P, R):
    WgL, bgL, wg, bg = P[12].reshape(-1), P[13], P[14].reshape(-1), P[15]
    s = _sig(R[:, None] * WgL + bgL)
    return s @ wg + bg[0], s


def gate_bwd(P, R, s, gbar):
    wg = P[14].reshape(-1)
    c = lambda v: v[:, None]
    ubar = c(gbar) * wg * s * (1 - s)
    return [(ubar * c(R)).sum(0).reshape(-1, 1), ubar.sum(0), (c(gbar) * s).sum(0).reshape(1, -1),
            np.array([gbar.sum()])]


def _forward(variant, P, x, y, z, R):
    nev, sN, cL, cV, cE = VARIANTS[variant]
    g = geometry(x, y, z, R)
    NA, DA, stA = mlp_fwd(P, g["f1"], g["f2"], g["al1"], g["al2"], g["al11"], g["al12"], g["al22"])
    Nsum, Dsum, stB = NA, DA, None
    if nev == 2:  # second evaluation with the orbitals swapped (poc/main.py:255-260)
        NB, DB, stB = mlp_fwd(P, g["f2"], g["f1"], g["al2"], g["al1"], g["al22"], g["al12"], g["al11"])
        Nsum, Dsum = NA + NB, DA + DB
    N = sN * Nsum + P[5][0]
    DN = sN * Dsum
    E, stE = enet_fwd(P, R)
    gate, stG = gate_fwd(P, R)
    q = g["ir1"] + g["ir2"]
    psi = g["f1"] + g["f2"] + gate * N
    lap = g["al1"] + g["al2"] + gate * DN
    res = (cL * (g["f1"] + g["f2"]) + (cV - 2 * cL) * (g["f1"] * g["ir1"] + g["f2"] * g["ir2"])
           + cV * (g["f1"] * g["ir2"] + g["f2"] * g["ir1"]) + gate * (cL * DN + cV * q * N) + cE * E * psi)
    return dict(g=g, stA=stA, stB=stB, stE=stE, stG=stG, N=N, DN=DN, E=E, gate=gate, q=q,
                psi=psi, lap=lap, res=res)


def fields(variant, theta, x, y, z, R):
    """Per-point psi, laplacian, residual, E (and H psi for the poc form)."""
    P = layout.unpack_poc(theta)
    f = _forward(variant, P, *[np.asarray(v, np.float64).reshape(-1) for v in (x, y, z, R)])
    g = f["g"]
    hpsi = (-0.5 * (g["f1"] + g["f2"]) - g["f1"] * g["ir2"] - g["f2"] * g["ir1"]
            + f["gate"] * (-0.5 * f["DN"] - f["q"] * f["N"]))
    return dict(psi=f["psi"], lap=f["lap"], res=f["res"], E=f["E"], hpsi=hpsi)


def loss_and_grad(variant, theta, x, y, z, R, m1, m2, w_pde=None, w_bc1=None, w_bc2=None):
    """Loss terms and dLtot/dtheta by the hand-written reverse sweep.

    m1/m2 are boolean (or 0/1) per-point membership of the two boundary sets; the
    weights default to the reference's means: 1/n, 1/|set1|, 1/|set2|."""
    P = layout.unpack_poc(theta)
    x, y, z, R = [np.asarray(v, np.float64).reshape(-1) for v in (x, y, z, R)]
    m1 = np.asarray(m1, np.float64).reshape(-1)
    m2 = np.asarray(m2, np.float64).reshape(-1)
    n = x.size
    w_pde = 1.0 / n if w_pde is None else w_pde
    w_bc1 = 1.0 / m1.sum() if w_bc1 is None else w_bc1
    w_bc2 = 1.0 / m2.sum() if w_bc2 is None else w_bc2
    nev, sN, cL, cV, cE = VARIANTS[variant]
    f = _forward(variant, P, x, y, z, R)
    res, psi, gate, E, N, DN, q = f["res"], f["psi"], f["gate"], f["E"], f["N"], f["DN"], f["q"]
    sums = dict(res2=(res * res).sum(), psi2_1=(m1 * psi * psi).sum(), psi2_2=(m2 * psi * psi).sum(),
                E=E.sum())
    Lpde = w_pde * sums["res2"]
    Lbc = w_bc1 * sums["psi2_1"] + w_bc2 * sums["psi2_2"]
    # seeds
    rbar = 2 * w_pde * res
    pbar = 2 * (w_bc1 * m1 + w_bc2 * m2) * psi
    lamN = rbar * gate * (cV * q + cE * E) + pbar * gate
    lamD = rbar * cL * gate
    gbar = rbar * (cL * DN + cV * q * N + cE * E * N) + pbar * N
    Ebar = rbar * cE * psi
    gA = mlp_bwd(P, f["stA"], sN * lamN, sN * lamD)
    if nev == 2:
        gB = mlp_bwd(P, f["stB"], sN * lamN, sN * lamD)
        gA = [a + b for a, b in zip(gA, gB)]
    grads = gA + [np.array([lamN.sum()])] + enet_bwd(P, R, f["stE"], Ebar) + gate_bwd(P, R, f["stG"], gbar)
    return dict(Ltot=Lpde + Lbc, Lpde=Lpde, Lbc=Lbc, E=E, grad=layout.pack_poc(grads), sums=sums,
                psi=psi, res=res, lap=f["lap"])
