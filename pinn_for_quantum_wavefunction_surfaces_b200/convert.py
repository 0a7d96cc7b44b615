"""Checkpoint formats either side of the hot path; host-only, no device needed.

    read_model_bin / write_model_bin   the little-endian format train.py:112-119 writes and energy.py:8-19 / plot.py:6-17 read
    theta_from_model_bin / theta_to_model_bin   (in,out) tuple order of train.py  <->  packed canonical theta
    theta_from_pt / state_dict_from_theta       `torch.save({'model_state_dict', 'optimizer_state_dict'})` of poc/main.py:325-339

NB the two files hold the same 16 tensors but belong to two FORMULATIONS of the model: a `.pt`
trained with poc/main.py evaluates the base network twice and sums, `model.bin` trained with train.py once and doubles.
Converting moves numbers between containers; it does not translate one formulation into the other.
"""
import numpy as np
import torch

from . import params as P


def read_model_bin(path):
    """-> list of float64 arrays in file order (train.py's tuple: H1a,H1b,...,E3a,E3b)"""
    out = []
    with open(path, "rb") as f:
        while True:
            head = f.read(4)
            if len(head) < 4:
                break
            ndim = int.from_bytes(head, "little")
            if ndim == 0:
                break
            shape = [int.from_bytes(f.read(4), "little") for _ in range(ndim)]
            size = int(np.prod(shape)) * 8
            out.append(np.frombuffer(f.read(size), dtype="<f8").reshape(shape).copy())
    return out


def write_model_bin(path, tensors):
    """train.py:112-119 byte for byte"""
    with open(path, "wb") as f:
        for x in tensors:
            x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
            f.write(x.ndim.to_bytes(4, "little"))
            for d in x.shape:
                f.write(int(d).to_bytes(4, "little"))
            f.write(x.tobytes())


def theta_from_model_bin(path):
    ts = read_model_bin(path)
    if len(ts) != 16:
        raise ValueError("model.bin should hold 16 tensors, found %d" % len(ts))
    return P.pack_trainpy([torch.from_numpy(t) for t in ts], dtype=torch.float64).numpy()


def theta_to_model_bin(theta, path):
    write_model_bin(path, [t.numpy() for t in P.unpack_trainpy(torch.as_tensor(np.asarray(theta, np.float64)))])


def theta_from_pt(path):
    """packed theta (float64) and the raw optimizer_state_dict (or None) of a poc/main.py checkpoint"""
    ck = torch.load(path, map_location="cpu", weights_only=True)
    sd = ck["model_state_dict"]
    theta = P.pack_poc([sd[k] for k in P.POC_TENSOR_NAMES], dtype=torch.float64).numpy()
    return theta, ck.get("optimizer_state_dict")


def state_dict_from_theta(theta):
    """-> OrderedDict loadable by NN_ion.load_state_dict (float64, nn.Linear layout)"""
    from collections import OrderedDict
    parts = P.unpack_poc(torch.as_tensor(np.asarray(theta, np.float64)))
    return OrderedDict((k, t.clone()) for k, t in zip(P.POC_TENSOR_NAMES, parts))


def adam_state_from_pt(opt_state):
    """(m, v, step) packed like theta from a torch.optim.Adam state_dict whose param order is model.parameters()"""
    st = opt_state["state"]
    m, v, step = [], [], 0
    for i in range(16):
        e = st.get(i)
        shape_n = int(np.prod(P.POC_SHAPES[i]))
        if e is None:
            m.append(np.zeros(shape_n)); v.append(np.zeros(shape_n))
            continue
        m.append(e["exp_avg"].double().numpy().ravel()); v.append(e["exp_avg_sq"].double().numpy().ravel())
        step = max(step, int(e["step"]))
    return np.concatenate(m), np.concatenate(v), step
