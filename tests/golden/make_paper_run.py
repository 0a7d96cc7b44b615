"""Run the REAL reference's paper schedule (poc/main.py:919-942) on the CPU of the build container and keep its outcome as a
fixture: `train(params)` with n_train = 100 000, 5000 epochs @ 8e-3, then `train(params, loadWeights=True, freezeUnits=True)`
for 2000 epochs @ 5e-4 from the saved best model - the reference's own functions (AST-loaded from /root/reference, see
make_golden.py), torch CPU float64, `torch.manual_seed(seed)`.

    python tests/golden/make_paper_run.py [--seed 0] [--epochs1 5000] [--epochs2 2000] [--n 100000] [--threads 4]

Takes ~1 h on a few cores (0.35 s per step at 1e5 points).  Output: tests/golden/poc_paper_run_seed<seed>.npz with the
loss histories, the E-net table on exactE()'s R grid after each stage, and the packed parameters (initial, saved best of
stage 1, saved best of stage 2).  The device-resident run of the same schedule (tools/acceptance.py C) is compared with
it and with the authors' own table (poc/energy_R_ion.pkl).
"""
import argparse
import os
import pickle
import sys
import tempfile
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import make_golden as mg  # noqa: E402
from oracle import layout  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--epochs1", type=int, default=5000)
    ap.add_argument("--epochs2", type=int, default=2000)
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    torch.set_num_threads(a.threads)
    torch.set_default_dtype(torch.double)
    ns = mg.load_poc_namespace()
    from os import path
    ns.update(path=path, time=time, pickle=pickle)
    params = ns["set_params"]()
    Rt, Eex = ns["exactE"]()
    Rt = np.asarray(Rt, np.float64)

    def enet(model):
        with torch.no_grad():
            R = torch.tensor(Rt).reshape(-1, 1)
            z = torch.zeros_like(R)
            return model.parametricPsi(z + 1.0, z, z, R)[1].numpy().ravel()

    def theta_of(ptfile):
        ck = torch.load(ptfile, map_location="cpu", weights_only=True)
        return layout.pack_poc([v.numpy() for v in ck["model_state_dict"].values()])

    cwd = os.getcwd()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        os.makedirs("models"); os.makedirs("data")
        try:
            torch.manual_seed(a.seed)
            m0 = ns["NN_ion"](params)        # the same construction train() performs first under this seed
            out["theta_init"] = layout.pack_poc([v.detach().numpy() for v in m0.state_dict().values()])
            torch.manual_seed(a.seed)
            params.update(epochs=a.epochs1, n_train=a.n, lr=8e-3, saveModelPath="models/ionHsym.pt",
                          loadModelPath="models/ionHsym.pt", lossPath="data/loss_ionH.pkl")
            t0 = time.time()
            ns["train"](params, loadWeights=False)
            t1 = time.time()
            out["theta_stage1_saved"] = theta_of("models/ionHsym.pt")
            with open("data/loss_ionH.pkl", "rb") as f:
                l1 = pickle.load(f)                     # written a moment ago by the reference's saveLoss()
            params.update(epochs=a.epochs2, lr=5e-4, saveModelPath="models/ionHsym_fineTune.pt",
                          lossPath="data/loss_ionH_fineTune.pkl")
            ns["train"](params, loadWeights=True, freezeUnits=True)
            t2 = time.time()
            out["theta_stage2_saved"] = theta_of("models/ionHsym_fineTune.pt")
            with open("data/loss_ionH_fineTune.pkl", "rb") as f:
                l2 = pickle.load(f)
            for tag, pt in (("stage1", "models/ionHsym.pt"), ("stage2", "models/ionHsym_fineTune.pt")):
                params["loadModelPath"] = pt
                m = ns["NN_ion"](params)
                m.loadModel(params)
                out["E_net_" + tag] = enet(m)
        finally:
            os.chdir(cwd)
    for k in ("Ltot", "Lpde", "Lbc", "Energy"):
        out["loss1_" + k] = np.asarray(l1[k]).ravel()
        out["loss2_" + k] = np.asarray(l2[k]).ravel()
    out.update(R=Rt, E_exact=np.asarray(Eex, np.float64), seed=a.seed, n=a.n, epochs=np.array([a.epochs1, a.epochs2]),
               seconds=np.array([t1 - t0, t2 - t1]), threads=a.threads)
    dst = a.out or os.path.join(HERE, "poc_paper_run_seed%d.npz" % a.seed)
    np.savez_compressed(dst, **out)
    for tag in ("stage1", "stage2"):
        err = np.abs(out["E_net_" + tag] - out["E_exact"])
        print(tag, "max|E_net-exact| R>=1: %.3e  R>=2: %.3e" % (err[Rt >= 1 - 1e-9].max(), err[Rt >= 2 - 1e-9].max()))
    print("stage1 Ltot min %.3e tail %.3e ; stage2 Ltot min %.3e tail %.3e ; %.0f s + %.0f s"
          % (out["loss1_Ltot"].min(), out["loss1_Ltot"][-100:].mean(), out["loss2_Ltot"].min(), out["loss2_Ltot"][-100:].mean(),
             t1 - t0, t2 - t1))


if __name__ == "__main__":
    main()
