"""CPU tests of the oracle pieces next to the hot path (sampler stream, Simpson rule, training-loop restatement,
checkpoint containers).  The oracle is test infrastructure; it is pinned here before the GPU tests trust it."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import pinn_for_quantum_wavefunction_surfaces_b200 as pk
from oracle import philox as ph
from oracle import simpson as sp
from oracle import train_loop as tl


def test_philox_known_answer_vectors():
    """Random123 kat_vectors for philox4x32-10."""
    def kat(c, k):
        r = ph.philox4x32_10(*[np.array([v], np.uint32) for v in c], k[0], k[1])
        return ["%08x" % int(v[0]) for v in r]
    assert kat((0, 0, 0, 0), (0, 0)) == ["6627e8d5", "e169c58d", "bc57ac4c", "9b00dbd8"]
    assert kat((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == ["408f276d", "41c83b0e", "a20bc7c6", "6d5451fd"]
    assert kat((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        ["d16cfe09", "94fdcceb", "5001e420", "24126ea1"]


def test_sampler_stream_rules():
    x, y, z, R, mask, (c1, c2) = ph.sample_batch(200000, seed=7, batch=3)
    for v, lo, hi in ((x, -18, 18), (y, -18, 18), (z, -18, 18), (R, 0.2, 4.0)):
        assert v.dtype == np.float32 and v.min() >= lo and v.max() <= hi
        assert abs(v.mean() - 0.5 * (lo + hi)) < 0.1 and abs(v.std() - (hi - lo) / 12 ** 0.5) < 0.05
    r1 = np.sqrt((x.astype(np.float64) - R) ** 2 + y.astype(np.float64) ** 2 + z.astype(np.float64) ** 2)
    far = np.abs(r1 - 17.5) > 1e-3
    assert np.array_equal((mask & 1).astype(bool)[far], (r1 >= 17.5)[far])
    assert c1 == int((mask & 1).sum()) and c2 == int((mask >> 1).sum())
    assert 0.45 < c1 / x.size < 0.60   # about 52 % of the box is farther than 17.5 from a nucleus
    # a box that is entirely inside the clamp radius of nucleus 1: every x becomes the VALUE cutoff (train.py:34)
    xs, _, _, Rs, _, _ = ph.sample_batch(64, 1, 0, box=(0.999, 1.001, -1e-3, 1e-3, -1e-3, 1e-3, 1.0, 1.0 + 1e-7))
    assert np.all(xs == np.float32(0.005))
    # different batches / seeds give different points; same arguments the same
    a = ph.sample_batch(100, 5, 0)[0]
    assert np.array_equal(a, ph.sample_batch(100, 5, 0)[0])
    assert not np.array_equal(a, ph.sample_batch(100, 5, 1)[0]) and not np.array_equal(a, ph.sample_batch(100, 6, 0)[0])


def test_simpson_rules():
    for n in (3, 9, 10, 79, 80):
        x = np.linspace(-1.0, 2.5, n)
        h = x[1] - x[0]
        for f in (np.exp(-x * x), x ** 3 - x, np.cos(3 * x)):
            assert abs(sp.simps_avg(f, x) - sp.weights_avg(n, h) @ f) < 1e-13
            assert abs(pk.analysis.simpson_weights(n, h, "avg") @ f - sp.simps_avg(f, x)) < 1e-13
    x = np.linspace(-1.0, 2.0, 9)   # odd N: Simpson is exact for cubics
    assert abs(sp.simps_avg(x ** 3, x) - (2 ** 4 - 1) / 4) < 1e-13
    from scipy.integrate import simpson
    for n in (9, 10, 80, 81):       # today's scipy rule == the 'simpson' weights; == 'avg' for odd N
        x = np.linspace(0.0, 3.0, n)
        f = np.exp(-x) * np.sin(2 * x)
        assert abs(simpson(f, x=x) - sp.weights_simpson(n, x[1] - x[0]) @ f) < 1e-13
        assert abs(pk.analysis.simpson_weights(n, x[1] - x[0], "simpson") @ f - simpson(f, x=x)) < 1e-13
        if n % 2:
            assert abs(simpson(f, x=x) - sp.simps_avg(f, x)) < 1e-13
    # nested integra3d == product weights
    rng = np.random.default_rng(0)
    ax = np.linspace(-2, 2, 8)
    F = rng.standard_normal((8, 8, 8))
    w = sp.weights_avg(8, ax[1] - ax[0])
    assert abs(sp.integra3d(ax, ax, ax, F) - np.einsum("i,j,k,ijk->", w, w, w, F)) < 1e-12


@pytest.mark.parametrize("epochs", [40, 200])
def test_training_loop_restatement_reproduces_the_reference_run(golden_dir, epochs):
    """oracle.train_loop.trainpy_run (train.py:13-110 restated) + the float64 oracle loss == the real script's trace and
    model.bin (fixtures written by tests/golden/make_golden.py from the unmodified reference).  200 epochs at n = 4096 is
    BASELINE config 1 as written (train.py:81, 110 with n, epochs replaced)."""
    tr = json.load(open(os.path.join(golden_dir, "trainpy_trace_n4096_e%d.json" % epochs)))
    params, trace, hist = tl.trainpy_run(tl.oracle_trainpy_op, n=4096, epochs=epochs)
    assert trace == tr["trace"]
    ref = pk.convert.read_model_bin(os.path.join(golden_dir, "trainpy_model_n4096_e%d.bin" % epochs))
    for a, b in zip(params, ref):
        # (the real script's own result moves by 4e-15 with the thread count of torch.mean over 200 steps)
        assert a.shape == b.shape and np.abs(a.detach().numpy() - b).max() < 1e-12
    assert len(tl.model_bin_bytes(params)) == tr["model_bin_size"]


def test_model_bin_container_roundtrip(golden_dir, tmp_path):
    src = os.path.join(golden_dir, "trainpy_model_n4096_e40.bin")
    theta = pk.convert.theta_from_model_bin(src)
    assert theta.shape == (1521,)
    out = tmp_path / "model.bin"
    pk.convert.theta_to_model_bin(theta, str(out))
    assert hashlib.md5(out.read_bytes()).hexdigest() == hashlib.md5(open(src, "rb").read()).hexdigest()
    # energy.py's reader logic (energy.py:8-19) sees the same 16 tensors
    ts = pk.convert.read_model_bin(str(out))
    assert [t.shape for t in ts][:4] == [(2, 16), (16,), (16, 16), (16,)] and len(ts) == 16
    sd = pk.convert.state_dict_from_theta(theta)
    assert list(sd.keys()) == pk.POC_TENSOR_NAMES and sd["Lin_H2.weight"].shape == (16, 16)
    back = pk.pack_poc([sd[k] for k in pk.POC_TENSOR_NAMES], dtype=torch.float64).numpy()
    assert np.array_equal(back, theta)


def test_pt_container(golden_dir, tmp_path):
    theta = np.load(os.path.join(golden_dir, "checkpoints.npz"))["ionHsym"]
    sd = pk.convert.state_dict_from_theta(theta)
    ps = [torch.nn.Parameter(v.clone()) for v in sd.values()]
    opt = torch.optim.Adam(ps, lr=1e-3)
    for p in ps:
        p.grad = torch.ones_like(p)
    opt.step()
    path = tmp_path / "m.pt"
    torch.save({"model_state_dict": sd, "optimizer_state_dict": opt.state_dict()}, str(path))  # poc/main.py:332-339
    th, ost = pk.convert.theta_from_pt(str(path))
    assert np.array_equal(th, theta)
    m, v, step = pk.convert.adam_state_from_pt(ost)
    assert step == 1 and np.allclose(m, 0.1) and np.allclose(v, 1e-3)


def test_init_trainpy_is_the_scripts_init(golden_dir):
    assert np.array_equal(pk.init_trainpy(12345), np.load(os.path.join(golden_dir, "trainpy_n2048.npz"))["theta"])
