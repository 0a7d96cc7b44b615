"""Drop-in seams: route the reference's own training code through the fused kernels.

* ``patch_nn_ion(NN_ion)`` replaces ``NN_ion.LossFunctions`` (poc/main.py:341-355); the training loop
  poc/main.py:359-430, its optimizer, freeze flags, history arrays and .pt checkpoints run unchanged.
* ``run_train_py(path)`` executes the reference ``train.py`` with its inline hot block (train.py:41-57,
  which has no function seam) replaced by one call; sampler, Adam loop, best-parameter tracking, prints
  and the ``model.bin`` writer are the reference's own lines.  The reference file is read where it lies
  and is never modified or copied.
"""
import contextlib
import io
import os

from . import ops
from .params import check_supported_model

_HOT_FIRST = "r1 = torch.sqrt((x - R)**2 + y**2 + z**2)"
_HOT_LAST = "Ltot = Lpde + Lbc"


def patch_nn_ion(nn_ion_cls, loss_fn=None):
    """Monkey-patch an ``NN_ion`` class (the reference's, poc/main.py:223).  Returns the original method."""
    original = getattr(nn_ion_cls, "LossFunctions", None)
    fn = loss_fn or ops.loss_poc

    def LossFunctions(self, x, y, z, R, params, bIndex1, bIndex2):  # same signature as poc/main.py:341
        # the kernel hard-codes what the shipped model uses: P = +1, nuclei on the x axis; refuse anything else
        check_supported_model(params if isinstance(params, dict) else None, self)
        return fn(self, x, y, z, R, bIndex1, bIndex2)

    nn_ion_cls.LossFunctions = LossFunctions
    return original


def trainpy_patched_source(src, n=None, epochs=None, op_name="__pinn_loss__"):
    """Return train.py's source with the hot block (train.py:41-57) replaced by a call to `op_name`."""
    lines = src.split("\n")
    first = [i for i, ln in enumerate(lines) if ln.strip() == _HOT_FIRST]
    last = [i for i, ln in enumerate(lines) if ln.strip() == _HOT_LAST]
    if len(first) != 1 or len(last) != 1 or last[0] <= first[0]:
        raise ValueError("train.py does not contain the expected hot block (train.py:41-57)")
    i0, i1 = first[0], last[0]
    indent = lines[i0][:len(lines[i0]) - len(lines[i0].lstrip())]
    call = indent + "Ltot, Lpde, Lbc, e = %s(x, y, z, R, i1, i2, *params)" % op_name
    out = lines[:i0] + [call] + lines[i1 + 1:]
    text = "\n".join(out)
    if n is not None:
        if "\nn = 10000\n" not in text:
            raise ValueError("train.py: `n = 10000` not found")
        text = text.replace("\nn = 10000\n", "\nn = %d\n" % n)
    if epochs is not None:
        if "epochs=1000)" not in text:
            raise ValueError("train.py: `epochs=1000)` not found")
        text = text.replace("train(params, lr=8e-3, epochs=1000)", "train(params, lr=8e-3, epochs=%d)" % epochs)
    return text


def run_train_py(path, n=None, epochs=None, workdir=None, loss_op=None, capture=True):
    """Run the reference train.py with its hot block routed through the fused kernel.

    path: the reference's train.py; workdir: where model.bin is written (default: cwd).
    loss_op: callable with PinnLossTrainPy.apply's signature (tests inject the CPU oracle here to check
    the patching itself).  Returns (namespace, stdout text)."""
    with open(path) as f:
        src = f.read()
    text = trainpy_patched_source(src, n=n, epochs=epochs)
    ns = {"__name__": "__pinn_train_py__", "__pinn_loss__": loss_op or ops.loss_trainpy}
    cwd = os.getcwd()
    buf = io.StringIO()
    try:
        if workdir:
            os.chdir(workdir)
        code = compile(text, os.path.basename(path) + "[patched 41-57]", "exec")
        if capture:
            with contextlib.redirect_stdout(buf):
                exec(code, ns)
        else:
            exec(code, ns)
    finally:
        os.chdir(cwd)
    return ns, buf.getvalue()
