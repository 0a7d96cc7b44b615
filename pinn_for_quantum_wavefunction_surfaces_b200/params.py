"""Packed parameter vector `theta` (1521 scalars) <-> the reference's 16 tensors.

Canonical order is ``NN_ion.state_dict()`` order with ``nn.Linear`` (out,in) layout
(poc/main.py:233-245); ``train.py`` keeps the same tensors in (in,out) layout with the gate
before the E-net (train.py:88-109).
"""
import torch

N_THETA = 1521

POC_TENSOR_NAMES = [
    "Lin_H1.weight", "Lin_H1.bias", "Lin_H2.weight", "Lin_H2.bias", "Lin_out.weight", "Lin_out.bias",
    "Lin_E1.weight", "Lin_E1.bias", "Lin_E2.weight", "Lin_E2.bias", "Lin_Eout.weight", "Lin_Eout.bias",
    "netDecayL.weight", "netDecayL.bias", "netDecay.weight", "netDecay.bias",
]
POC_SHAPES = [(16, 2), (16,), (16, 16), (16,), (1, 16), (1,), (32, 1), (32,), (32, 32), (32,), (1, 32), (1,),
              (10, 1), (10,), (1, 10), (1,)]
TRAINPY_TENSOR_NAMES = ["H1a", "H1b", "H2a", "H2b", "H3a", "H3b", "L1a", "L1b", "L2a", "L2b",
                        "E1a", "E1b", "E2a", "E2b", "E3a", "E3b"]
# position of each train.py tensor in the canonical list
TRAINPY_TO_POC = [0, 1, 2, 3, 4, 5, 12, 13, 14, 15, 6, 7, 8, 9, 10, 11]

# grad_mask of the reference's fine-tune mode: freezeBase + freezeDecayUnit (poc/main.py:305-319)
FINE_TUNE_GRAD_MASK = 0x0FC0


def pack_poc(tensors, dtype=torch.float32, device=None):
    """16 tensors in state_dict order -> flat vector."""
    flat = [t.detach().reshape(-1) for t in tensors]
    v = torch.cat(flat).to(dtype=dtype, device=device if device is not None else flat[0].device)
    if v.numel() != N_THETA:
        raise ValueError("expected 1521 parameters, got %d" % v.numel())
    return v


def unpack_poc(theta):
    out, off = [], 0
    for shp in POC_SHAPES:
        n = 1
        for s in shp:
            n *= s
        out.append(theta[off:off + n].reshape(shp))
        off += n
    return out


def pack_trainpy(tensors, dtype=torch.float32, device=None):
    """16 tensors in train.py order and (in,out) layout -> canonical flat vector."""
    poc = [None] * 16
    for t, dst in zip(tensors, TRAINPY_TO_POC):
        t = t.detach()
        poc[dst] = t.t().contiguous() if t.dim() == 2 else t
    return pack_poc(poc, dtype=dtype, device=device)


def unpack_trainpy(theta):
    """canonical flat vector -> 16 tensors shaped like train.py's."""
    poc = unpack_poc(theta)
    out = []
    for dst in TRAINPY_TO_POC:
        t = poc[dst]
        out.append(t.t().contiguous() if t.dim() == 2 else t.clone())
    return out


def grad_mask_from_requires_grad(tensors, order="poc"):
    """bit i set iff canonical tensor i wants a gradient (honours freezeBase/freezeDecayUnit)."""
    mask = 0
    for i, t in enumerate(tensors):
        if getattr(t, "requires_grad", False):
            mask |= 1 << (i if order == "poc" else TRAINPY_TO_POC[i])
    return mask


def check_supported_model(params=None, model=None):
    """The kernels implement the model the reference ships and trains: inversion symmetry P = +1 (base(f1,f2) + base(f2,f1)
    and LCAO f1 + f2, poc/main.py:255-260, 289) with the nuclei on the x axis (Ry = Rz = 0, poc/main.py:28-29, 44).
    Anything else - `params['inversion_symmetry'] = -1`, displaced nuclei - would silently train on the wrong loss, so
    it is refused.  `params`: the reference's parameter dict; `model`: an NN_ion instance (its P, Ry, Rz attributes)."""
    from ._lib import PinnError
    vals = {}
    if params is not None:
        vals.update({"inversion_symmetry": params.get("inversion_symmetry", 1), "Ry": params.get("Ry", 0),
                     "Rz": params.get("Rz", 0)})
    if model is not None:
        vals.update({"inversion_symmetry": getattr(model, "P", vals.get("inversion_symmetry", 1)),
                     "Ry": getattr(model, "Ry", vals.get("Ry", 0)), "Rz": getattr(model, "Rz", vals.get("Rz", 0))})
    if vals.get("inversion_symmetry", 1) != 1 or vals.get("Ry", 0) != 0 or vals.get("Rz", 0) != 0:
        raise PinnError("unsupported model configuration %r: the fused kernels implement inversion_symmetry = +1 with "
                        "Ry = Rz = 0 (poc/main.py:28-29, 44, 255-260) only" % (vals,))
