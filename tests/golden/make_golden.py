"""Generate the golden fixtures in this directory FROM THE REAL REFERENCE CODE.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

Nothing here is imported by the product.  The reference cannot travel to the GPU box,
so its outputs are committed as small .npz/.json fixtures next to this script.

How the reference is executed (no source is copied into the repo):
* poc/main.py cannot be imported (it needs matplotlib, scipy.integrate.simps and
  data files at import time), so its FunctionDef/ClassDef nodes are compiled from the
  file where it lies and executed in a scratch namespace (NN_ion, sampling, radial,
  lapl, hamiltonian, LossFunctions are then the reference's own objects).
* train.py is a script without a function seam: (a) the statements of its loop body
  that form the hot path (train.py:41-57) are compiled from the file by line range and
  run on fixed inputs; (b) the whole script is run with only `n` and `epochs` replaced.
"""
import ast
import hashlib
import io
import json
import os
import pickle
import contextlib
import sys
import warnings

import numpy as np
import torch

REF = os.environ.get("PINN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import layout  # noqa: E402

warnings.filterwarnings("ignore")


class _NumpyOnlyUnpickler(pickle.Unpickler):
    """poc/energy_R_ion.pkl is a dict of numpy arrays; refuse every other global (untrusted content)."""
    _ALLOWED = {("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
                ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy.core.multiarray", "scalar"),
                ("numpy._core.multiarray", "scalar")}

    def find_class(self, module, name):
        if (module, name) not in self._ALLOWED:
            raise pickle.UnpicklingError("refusing to load %s.%s from the reference pickle" % (module, name))
        return super().find_class(module, name)


def load_poc_namespace():
    import torch.nn as nn
    import torch.optim as optim
    from torch.autograd import grad
    src = open(os.path.join(REF, "poc", "main.py")).read()
    tree = ast.parse(src)
    # keep the FIRST definition of each name: the legacy block (main.py:680-910) re-defines
    # energy_from_psi etc. with dead single-R code
    seen, body = set(), []
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name not in seen:
            seen.add(node.name)
            body.append(node)
    ns = dict(torch=torch, np=np, nn=nn, optim=optim, grad=grad, dtype=torch.double, pickle=pickle)
    exec(compile(ast.Module(body=body, type_ignores=[]), "poc/main.py", "exec"), ns)
    ns["params"] = ns["set_params"]()
    return ns


def golden_poc():
    torch.set_default_dtype(torch.double)
    ns = load_poc_namespace()
    params = ns["params"]
    out = {}
    thetas = {}
    for tag in ("ionHsym", "ionHsym_fineTune"):
        # the reference tree is untrusted content: tensors only, no arbitrary pickles
        ck = torch.load(os.path.join(REF, "models", tag + ".pt"), map_location="cpu", weights_only=True)
        thetas[tag] = layout.pack_poc([v.numpy() for v in ck["model_state_dict"].values()])
    # a freshly constructed NN_ion (nn.Linear default init + Lin_Eout.bias = -1, poc/main.py:233-245) under seed 0:
    # pins init_poc(), the starting point of the device-resident paper-schedule run
    torch.manual_seed(0)
    m0 = ns["NN_ion"](params)
    thetas["init_seed0"] = layout.pack_poc([v.detach().numpy() for v in m0.state_dict().values()])
    assert [k for k in m0.state_dict()] == [k for k, _ in m0.named_parameters()]   # parameters() order == state_dict order
    np.savez(os.path.join(HERE, "checkpoints.npz"), **thetas)

    torch.manual_seed(0)
    n = 4096
    x, y, z, R = ns["sampling"](params, n)
    r1, r2 = ns["radial"](x, y, z, R, params)
    b1 = torch.where(r1 >= params["BCcutoff"])
    b2 = torch.where(r2 >= params["BCcutoff"])
    out.update(x=x.detach().numpy().ravel(), y=y.detach().numpy().ravel(), z=z.detach().numpy().ravel(),
               R=R.detach().numpy().ravel(), i1=b1[0].numpy(), i2=b2[0].numpy())
    for tag in ("ionHsym", "ionHsym_fineTune"):
        params["loadModelPath"] = os.path.join(REF, "models", tag + ".pt")
        m = ns["NN_ion"](params)
        m.loadModel(params)
        Ltot, Lpde, Lbc, E = m.LossFunctions(x, y, z, R, params, b1, b2)
        Ltot.backward()
        grad = layout.pack_poc([p.grad.numpy() for p in m.parameters()])
        out[tag + "_loss"] = np.array([Ltot.item(), Lpde.item(), Lbc.item()])
        out[tag + "_grad"] = grad
        psi, E = m.parametricPsi(x, y, z, R)
        lap = ns["lapl"](x, y, z, psi)
        hpsi = ns["hamiltonian"](x, y, z, R, psi, params)
        out[tag + "_psi"] = psi.detach().numpy().ravel()
        out[tag + "_lap"] = lap.detach().numpy().ravel()
        out[tag + "_hpsi"] = hpsi.detach().numpy().ravel()
        out[tag + "_E"] = E.detach().numpy().ravel()
    np.savez_compressed(os.path.join(HERE, "poc_seed0_n4096.npz"), **out)

    with open(os.path.join(REF, "poc", "energy_R_ion.pkl"), "rb") as f:
        d = _NumpyOnlyUnpickler(f).load()
    Rex, Eex = ns["exactE"]()
    np.savez(os.path.join(HERE, "energy_R_ion.npz"), R=np.asarray(d["R"]), E_net=np.asarray(d["E_net"]),
             E_int=np.asarray(d["E_int"]), Elcao=np.asarray(d["Elcao"]), R_exact=np.asarray(Rex),
             E_exact=np.asarray(Eex))
    print("poc goldens:", {k: out[k] for k in out if k.endswith("_loss")})


def trainpy_hot_lines():
    """Compile train.py lines 41-57 (the inline hot path) as a function of its free names."""
    lines = open(os.path.join(REF, "train.py")).read().split("\n")
    block = lines[40:57]
    assert block[0].strip().startswith("r1 = torch.sqrt") and block[-1].strip() == "Ltot = Lpde + Lbc", block
    ind = len(block[0]) - len(block[0].lstrip())
    body = "\n".join("    " + ln[ind:] for ln in block)
    names = "x, y, z, R, i1, i2, H1a, H1b, H2a, H2b, H3a, H3b, L1a, L1b, L2a, L2b, E1a, E1b, E2a, E2b, E3a, E3b"
    src = "def hot(%s):\n%s\n    return Ltot, Lpde, Lbc, e, psi, res\n" % (names, body)
    lin = "\n".join(lines[3:11])  # linear() and d() (train.py:4-10)
    ns = dict(torch=torch)
    exec(compile(lin + "\n" + src, "train.py[4-10,41-57]", "exec"), ns)
    return ns["hot"]


def golden_trainpy():
    hot = trainpy_hot_lines()
    torch.manual_seed(12345)
    shapes = [(2, 16), (16,), (16, 16), (16,), (16, 1), (1,), (1, 10), (10,), (10, 1), (1,),
              (1, 32), (32,), (32, 32), (32,), (32, 1), (1,)]
    ps = []
    for s in shapes:  # same rule as train.py:13-18
        t = torch.empty(s, dtype=torch.double)
        t.uniform_(-1 / s[0] ** 0.5, 1 / s[0] ** 0.5)
        ps.append(t.requires_grad_(True))
    n = 2048
    g = torch.Generator().manual_seed(7)
    x = ((2 * torch.rand(n, 1, generator=g, dtype=torch.double) - 1) * 18).requires_grad_(True)
    y = ((2 * torch.rand(n, 1, generator=g, dtype=torch.double) - 1) * 18).requires_grad_(True)
    z = ((2 * torch.rand(n, 1, generator=g, dtype=torch.double) - 1) * 18).requires_grad_(True)
    R = 0.2 + 2.8 * torch.rand(n, 1, generator=g, dtype=torch.double)
    r1sq = (x - R) ** 2 + y ** 2 + z ** 2
    r2sq = (x + R) ** 2 + y ** 2 + z ** 2
    i1, = torch.where(r1sq[:, 0] >= 17.5 ** 2)
    i2, = torch.where(r2sq[:, 0] >= 17.5 ** 2)
    Ltot, Lpde, Lbc, e, psi, res = hot(x, y, z, R, i1, i2, *ps)
    Ltot.backward()
    theta = layout.from_trainpy([p.detach().numpy() for p in ps])
    grad = layout.from_trainpy([p.grad.numpy() for p in ps])
    np.savez_compressed(os.path.join(HERE, "trainpy_n2048.npz"), theta=theta, grad=grad,
                        x=x.detach().numpy().ravel(), y=y.detach().numpy().ravel(), z=z.detach().numpy().ravel(),
                        R=R.numpy().ravel(), i1=i1.numpy(), i2=i2.numpy(),
                        loss=np.array([Ltot.item(), Lpde.item(), Lbc.item()]),
                        psi=psi.detach().numpy().ravel(), res=res.detach().numpy().ravel(),
                        e=e.detach().numpy().ravel())
    print("train.py hot-line goldens:", Ltot.item(), Lpde.item(), Lbc.item())

    # whole-script traces: BASELINE config 1 (n=4096, epochs=200) and its first 40 epochs, otherwise unmodified
    for epochs in (40, 200):
        whole_script_trace(4096, epochs)


def whole_script_trace(n, epochs):
    src = open(os.path.join(REF, "train.py")).read()
    src = src.replace("n = 10000", "n = %d" % n).replace("epochs=1000)", "epochs=%d)" % epochs)
    assert "n = %d" % n in src and "epochs=%d)" % epochs in src
    cwd = os.getcwd()
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        buf = io.StringIO()
        try:
            with contextlib.redirect_stdout(buf):
                exec(compile(src, "train.py", "exec"), {"__name__": "__main__"})
            md5 = hashlib.md5(open("model.bin", "rb").read()).hexdigest()
            blob = open("model.bin", "rb").read()
        finally:
            os.chdir(cwd)
    torch.set_default_dtype(torch.double)
    trace = [ln.strip() for ln in buf.getvalue().strip().split("\n")]
    tag = "n%d_e%d" % (n, epochs)
    with open(os.path.join(HERE, "trainpy_trace_%s.json" % tag), "w") as f:
        json.dump({"n": n, "epochs": epochs, "seed": 12345, "trace": trace, "model_bin_md5": md5,
                   "model_bin_size": len(blob), "torch": torch.__version__,
                   # the md5 is bit-level and follows the summation order of torch.mean: 1 / 8 / 16 threads give three
                   # different files whose parameters agree to 4e-15 after 200 steps (tests compare values, not the md5)
                   "threads": torch.get_num_threads()}, f, indent=1)
    with open(os.path.join(HERE, "trainpy_model_%s.bin" % tag), "wb") as f:
        f.write(blob)
    print("\n".join(trace))
    print("md5(model.bin) =", md5)


if __name__ == "__main__":
    golden_poc()
    golden_trainpy()
