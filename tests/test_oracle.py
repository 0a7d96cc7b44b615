"""Pin the CPU oracle against outputs of the REAL reference (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import torch

from oracle import closed_form as cf
from oracle import layout
from oracle import ref_autograd as ra


def _masks(n, i1, i2):
    m1 = np.zeros(n)
    m2 = np.zeros(n)
    m1[i1] = 1
    m2[i2] = 1
    return m1, m2


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max()


def test_layout_roundtrip():
    rng = np.random.default_rng(0)
    theta = rng.standard_normal(layout.N_THETA)
    assert np.array_equal(layout.pack_poc(layout.unpack_poc(theta)), theta)
    assert np.array_equal(layout.from_trainpy(layout.to_trainpy(theta)), theta)
    tp = layout.to_trainpy(theta)
    assert [t.shape for t in tp] == [(2, 16), (16,), (16, 16), (16,), (16, 1), (1,), (1, 10), (10,), (10, 1), (1,),
                                     (1, 32), (32,), (32, 32), (32,), (32, 1), (1,)]


def test_enet_matches_reference_pkl(golden_dir):
    # poc/energy_R_ion.pkl['E_net'] is the E-net of models/ionHsym_fineTune.pt at R=0.2..4.0
    ck = np.load(os.path.join(golden_dir, "checkpoints.npz"))
    en = np.load(os.path.join(golden_dir, "energy_R_ion.npz"))
    E, _ = cf.enet_fwd(layout.unpack_poc(ck["ionHsym_fineTune"]), en["R"])
    assert np.abs(E - en["E_net"]).max() < 1e-14
    assert abs(E[0] - (-1.62905384)) < 1e-8


def test_enet_against_exact_h2plus_energies(golden_dir):
    """Physics pin (exactE() poc/main.py:48-61): the E-net of the shipped fine-tuned model against the tabulated
    exact electronic energies of H2+ (Hartree) at internuclear distance D = 2R: within 6.4e-3 for R >= 1, 3.1e-3 for R >= 2."""
    exact = {2.0: -1.1026342, 3.0: -0.9108962, 4.0: -0.7960849, 6.0: -0.6786357, 8.0: -0.6275704}
    ck = np.load(os.path.join(golden_dir, "checkpoints.npz"))
    R = np.array([d / 2 for d in exact])
    E, _ = cf.enet_fwd(layout.unpack_poc(ck["ionHsym_fineTune"]), R)
    err = np.abs(E - np.array(list(exact.values())))
    assert err.max() < 6.5e-3 and err[R >= 2.0].max() < 3.2e-3


def test_poc_oracles_match_reference_outputs(golden_dir):
    ck = np.load(os.path.join(golden_dir, "checkpoints.npz"))
    g = np.load(os.path.join(golden_dir, "poc_seed0_n4096.npz"))
    n = g["x"].size
    assert (len(g["i1"]), len(g["i2"])) == (2151, 2134)
    m1, m2 = _masks(n, g["i1"], g["i2"])
    for tag in ("ionHsym", "ionHsym_fineTune"):
        theta = ck[tag]
        o = cf.loss_and_grad("poc", theta, g["x"], g["y"], g["z"], g["R"], m1, m2)
        ref = g[tag + "_loss"]
        assert abs(o["Ltot"] - ref[0]) / ref[0] < 1e-11
        assert abs(o["Lpde"] - ref[1]) / ref[1] < 1e-11
        assert abs(o["Lbc"] - ref[2]) / ref[2] < 1e-11
        assert _rel(o["grad"], g[tag + "_grad"]) < 1e-11
        f = cf.fields("poc", theta, g["x"], g["y"], g["z"], g["R"])
        assert _rel(f["psi"], g[tag + "_psi"]) < 1e-13
        assert _rel(f["lap"], g[tag + "_lap"]) < 1e-12
        assert _rel(f["hpsi"], g[tag + "_hpsi"]) < 1e-11
        assert _rel(f["E"], g[tag + "_E"]) < 1e-14
        # the autograd restatement (the CPU-baseline "port")
        t = lambda a: torch.tensor(a).reshape(-1, 1)
        Lt, Lp, Lb, E, gr = ra.loss_and_grad("poc", torch.tensor(theta), t(g["x"]), t(g["y"]), t(g["z"]), t(g["R"]),
                                             torch.tensor(g["i1"]), torch.tensor(g["i2"]))
        assert abs(Lt.item() - ref[0]) / ref[0] < 1e-12
        assert _rel(gr.numpy(), g[tag + "_grad"]) < 1e-12
    assert abs(g["ionHsym_loss"][0] - 8.2198452149e-07) < 1e-16


def test_trainpy_oracles_match_reference_hot_lines(golden_dir):
    g = np.load(os.path.join(golden_dir, "trainpy_n2048.npz"))
    n = g["x"].size
    m1, m2 = _masks(n, g["i1"], g["i2"])
    o = cf.loss_and_grad("trainpy", g["theta"], g["x"], g["y"], g["z"], g["R"], m1, m2)
    for k, v in zip(("Ltot", "Lpde", "Lbc"), g["loss"]):
        assert abs(o[k] - v) / v < 1e-12
    assert _rel(o["grad"], g["grad"]) < 1e-11
    assert _rel(o["psi"], g["psi"]) < 1e-13
    assert _rel(o["res"], g["res"]) < 1e-11
    assert _rel(o["E"], g["e"]) < 1e-14
    t = lambda a: torch.tensor(a).reshape(-1, 1)
    Lt, Lp, Lb, e, gr = ra.loss_and_grad("trainpy", torch.tensor(g["theta"]), t(g["x"]), t(g["y"]), t(g["z"]),
                                         t(g["R"]), torch.tensor(g["i1"]), torch.tensor(g["i2"]))
    assert abs(Lt.item() - g["loss"][0]) / g["loss"][0] < 1e-12
    assert _rel(gr.numpy(), g["grad"]) < 1e-12


def test_shard_sums_add_up(golden_dir):
    # data-parallel algebra: per-shard weighted sums/gradients add to the full result
    ck = np.load(os.path.join(golden_dir, "checkpoints.npz"))
    g = np.load(os.path.join(golden_dir, "poc_seed0_n4096.npz"))
    n = g["x"].size
    m1, m2 = _masks(n, g["i1"], g["i2"])
    w = (1.0 / n, 1.0 / m1.sum(), 1.0 / m2.sum())
    full = cf.loss_and_grad("poc", ck["ionHsym"], g["x"], g["y"], g["z"], g["R"], m1, m2)
    tot, grad = 0.0, 0.0
    for s in np.array_split(np.arange(n), 3):
        o = cf.loss_and_grad("poc", ck["ionHsym"], g["x"][s], g["y"][s], g["z"][s], g["R"][s], m1[s], m2[s], *w)
        tot += o["Ltot"]
        grad = grad + o["grad"]
    assert abs(tot - full["Ltot"]) / full["Ltot"] < 1e-12
    assert _rel(grad, full["grad"]) < 1e-11


def test_sampler_statistics():
    g = torch.Generator().manual_seed(0)
    x, y, z, R, i1, i2 = ra.sample_box(20000, "poc", g)
    assert x.abs().max() <= 18 and 0.2 <= R.min() and R.max() <= 4.0
    assert 0.45 < len(i1) / 20000 < 0.60 and 0.45 < len(i2) / 20000 < 0.60  # ~52 %


def test_trainpy_trace_fixture_is_the_survey_trace(golden_dir):
    tr = json.load(open(os.path.join(golden_dir, "trainpy_trace_n4096_e40.json")))
    assert tr["trace"][0] == "0: 1.47e-02 3.74e-03 1.09e-02 (7.01e-01) [1.46833e-02]"
    assert tr["trace"][2] == "20: 4.43e-04 7.47e-05 3.69e-04 (-6.26e-01) [4.61586e-05]"
    assert tr["model_bin_size"] == 12328
