"""Packed parameter layout used by the oracle (TEST INFRASTRUCTURE ONLY).

Canonical order = ``NN_ion.state_dict()`` order of the reference (poc/main.py:233-245),
every matrix in ``nn.Linear`` (out, in) row-major layout; 1521 scalars:

    W1(16,2) b1(16) W2(16,16) b2(16) wo(16) bo(1)           base MLP   337
    WE1(32)  bE1(32) WE2(32,32) bE2(32) wE(32) bE(1)        E-net     1153
    WgL(10)  bgL(10) wg(10) bg(1)                           gate        31

``train.py`` keeps its 16 tensors in another order (gate before E-net) and in
(in, out) layout (train.py:4-5, 108-109); ``from_trainpy`` / ``to_trainpy`` convert.
"""
import numpy as np

NH, NE, NL = 16, 32, 10
N_THETA = 1521

# (name, shape) in canonical order; names are the reference's state_dict keys
POC_TENSORS = [
    ("Lin_H1.weight", (NH, 2)), ("Lin_H1.bias", (NH,)),
    ("Lin_H2.weight", (NH, NH)), ("Lin_H2.bias", (NH,)),
    ("Lin_out.weight", (1, NH)), ("Lin_out.bias", (1,)),
    ("Lin_E1.weight", (NE, 1)), ("Lin_E1.bias", (NE,)),
    ("Lin_E2.weight", (NE, NE)), ("Lin_E2.bias", (NE,)),
    ("Lin_Eout.weight", (1, NE)), ("Lin_Eout.bias", (1,)),
    ("netDecayL.weight", (NL, 1)), ("netDecayL.bias", (NL,)),
    ("netDecay.weight", (1, NL)), ("netDecay.bias", (1,)),
]

# train.py tuple order (train.py:108-109) -> index of the matching canonical tensor
TRAINPY_NAMES = ["H1a", "H1b", "H2a", "H2b", "H3a", "H3b", "L1a", "L1b", "L2a", "L2b",
                 "E1a", "E1b", "E2a", "E2b", "E3a", "E3b"]
TRAINPY_TO_POC = [0, 1, 2, 3, 4, 5, 12, 13, 14, 15, 6, 7, 8, 9, 10, 11]


def offsets():
    off, out = 0, []
    for _, shp in POC_TENSORS:
        out.append(off)
        off += int(np.prod(shp))
    assert off == N_THETA
    return out


def pack_poc(tensors):
    """16 arrays in state_dict order -> flat float64 vector."""
    flat = [np.asarray(t, dtype=np.float64).reshape(-1) for t in tensors]
    v = np.concatenate(flat)
    assert v.size == N_THETA
    return v


def unpack_poc(theta):
    theta = np.asarray(theta)
    out = []
    for off, (_, shp) in zip(offsets(), POC_TENSORS):
        out.append(theta[off:off + int(np.prod(shp))].reshape(shp))
    return out


def from_trainpy(tensors):
    """16 arrays in train.py order/(in,out) layout -> canonical flat vector."""
    poc = [None] * 16
    for t, dst in zip(tensors, TRAINPY_TO_POC):
        a = np.asarray(t, dtype=np.float64)
        poc[dst] = a.T if a.ndim == 2 else a
    return pack_poc([np.ascontiguousarray(p) for p in poc])


def to_trainpy(theta):
    """canonical flat vector -> 16 arrays in train.py order/(in,out) layout."""
    poc = unpack_poc(theta)
    out = []
    for dst in TRAINPY_TO_POC:
        a = poc[dst]
        out.append(np.ascontiguousarray(a.T) if a.ndim == 2 else a.copy())
    return out
