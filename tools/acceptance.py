"""Acceptance evidence of the north star's training clauses, measured on a B200 (run under gpurun; results -> gpurun_out/,
copied by hand into profiles/):

  A. BASELINE config 1 as written: train.py, n = 4096, 200 epochs (train.py:81, 110), the reference loop (torch RNG seed
     12345, float64 Adam on the host) driven by the fused CUDA op instead of lines 41-57, against the float64 oracle run
     of the same loop (pinned bit-for-bit to the real script, tests/test_oracle_train.py): drift of E(R) and of the
     parameters at steps 10, 40, 100, 200, final best-model E(R) against the real script's model.bin.
  B. The same run through the device-resident trainer fed with the reference's batches: per-step loss history against the
     oracle's.
  C. The paper schedule of poc/main.py:919-942 on the device: 100 000 points x 5000 Adam steps @ 8e-3, then 2000 fine-tune
     steps @ 5e-4 on the E-net only from the saved best model; E(R) table next to exactE() (poc/main.py:48-61) and
     poc/energy_R_ion.pkl (the authors' table).

    python tools/acceptance.py [A] [B] [C] [--seeds 0,1,2]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pinn_for_quantum_wavefunction_surfaces_b200 as pk  # noqa: E402
from oracle import layout  # noqa: E402
from oracle import train_loop as tl  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
OUT = os.path.join(ROOT, "gpurun_out")
SNAP = (10, 40, 100, 200)


def enet(theta, R):
    P = layout.unpack_poc(np.asarray(theta, np.float64))
    sig = lambda u: 1.0 / (1.0 + np.exp(-u))
    e = sig(R[:, None] * P[6][:, 0][None, :] + P[7][None, :])
    e = sig(e @ P[8].T + P[9][None, :])
    return e @ P[10][0] + P[11][0]


def part_a():
    Rg = np.linspace(0.2, 3.0, 57)
    s_ref = {k: None for k in SNAP}
    t0 = time.time()
    p_ref, trace_ref, hist_ref = tl.trainpy_run(tl.oracle_trainpy_op, n=4096, epochs=200, snapshots=s_ref)
    t_ref = time.time() - t0
    s_gpu = {k: None for k in SNAP}
    t0 = time.time()
    p_gpu, trace_gpu, hist_gpu = tl.trainpy_run(pk.loss_trainpy, n=4096, epochs=200, snapshots=s_gpu)
    t_gpu = time.time() - t0
    gold = json.load(open(os.path.join(GOLD, "trainpy_trace_n4096_e200.json")))
    out = {"oracle_trace_equals_real_script": trace_ref == gold["trace"], "seconds_oracle_loop": t_ref,
           "seconds_fused_loop": t_gpu, "drift": []}
    for k in SNAP:
        dE = float(np.abs(enet(s_gpu[k], Rg) - enet(s_ref[k], Rg)).max())
        dth = float(np.abs(s_gpu[k] - s_ref[k]).max())
        out["drift"].append({"step": k, "max_abs_dE_R": dE, "max_abs_dtheta": dth})
    best_gpu = pk.pack_trainpy(p_gpu, dtype=torch.float64).numpy()
    best_ref = pk.convert.theta_from_model_bin(os.path.join(GOLD, "trainpy_model_n4096_e200.bin"))
    out["best_model_max_abs_dE_R"] = float(np.abs(enet(best_gpu, Rg) - enet(best_ref, Rg)).max())
    out["best_model_max_abs_dtheta"] = float(np.abs(best_gpu - best_ref).max())
    rel = np.abs(hist_gpu - hist_ref) / np.maximum(np.abs(hist_ref), 1e-300)
    out["loss_history_rel_diff"] = {"Ltot_max": float(rel[:, 0].max()), "Ltot_median": float(np.median(rel[:, 0])),
                                    "Ltot_at": {str(k): float(rel[k, 0]) for k in SNAP},
                                    "meanE_abs_max": float(np.abs(hist_gpu[:, 3] - hist_ref[:, 3]).max())}
    out["trace_fused"] = trace_gpu
    out["trace_reference"] = gold["trace"]
    return out


def part_b():
    n, epochs = 4096, 200
    hist_ref = tl.trainpy_run(tl.oracle_trainpy_op, n=n, epochs=epochs)[2]
    torch.manual_seed(12345)
    theta0 = pk.init_trainpy(12345)
    # consume the generator exactly like the script's parameter draw (train.py:88-103)
    shapes = [(2, 16), (16,), (16, 16), (16,), (16, 1), (1,), (1, 10), (10,), (10, 1), (1,), (1, 32), (32,), (32, 32), (32,),
              (32, 1), (1,)]
    for s in shapes:
        torch.empty(s, dtype=torch.double).uniform_(-1 / s[0] ** 0.5, 1 / s[0] ** 0.5)
    x, y, z, R = [torch.empty(n, 1, dtype=torch.double) for _ in range(4)]
    tr = pk.Trainer("trainpy", n, theta0, lr=8e-3, history_capacity=epochs + 1)
    for tt in range(epochs + 1):
        x.uniform_(-18, 18); y.uniform_(-18, 18); z.uniform_(-18, 18); R.uniform_(0.2, 3)
        r1sq = (x - R) ** 2 + y ** 2 + z ** 2
        r2sq = (x + R) ** 2 + y ** 2 + z ** 2
        x[r1sq < 0.005 ** 2] = 0.005
        x[r2sq < 0.005 ** 2] = 0.005
        m1 = ((x - R) ** 2 + y ** 2 + z ** 2 >= 17.5 ** 2)[:, 0]
        m2 = ((x + R) ** 2 + y ** 2 + z ** 2 >= 17.5 ** 2)[:, 0]
        mask = (m1.to(torch.uint8) + 2 * m2.to(torch.uint8))
        tr.set_batch(x.float(), y.float(), z.float(), R.float(), mask, [1.0 / n, 1.0 / int(m1.sum()), 1.0 / int(m2.sum())])
        tr.run(1, resample=False, use_graph=(tt > 0))
    r = tr.read()
    tr.close()
    rel = np.abs(r["history"] - hist_ref) / np.maximum(np.abs(hist_ref), 1e-300)
    best_ref = pk.convert.theta_from_model_bin(os.path.join(GOLD, "trainpy_model_n4096_e200.bin"))
    Rg = np.linspace(0.2, 3.0, 57)
    return {"Ltot_rel_diff_max": float(rel[:, 0].max()), "Ltot_rel_diff_median": float(np.median(rel[:, 0])),
            "Ltot_rel_diff_at": {str(k): float(rel[k, 0]) for k in SNAP}, "best_loss": r["best_loss"],
            "best_step": r["best_step"], "best_loss_reference": float(hist_ref[:, 0].min()),
            "best_model_max_abs_dE_R": float(np.abs(enet(r["best_theta"], Rg) - enet(best_ref, Rg)).max()),
            "note": "points travel as float32 here (set_batch), the host loop of part A hands the op float64 points"}


def part_c(seeds):
    en = np.load(os.path.join(GOLD, "energy_R_ion.npz"))
    Rt, Eex, Eref = en["R"], np.asarray(en["E_exact"]).ravel(), en["E_net"].ravel()
    runs = []
    for seed in seeds:
        theta0 = pk.init_poc(seed)
        torch.cuda.synchronize()
        t0 = time.time()
        last1, saved1, loss1 = pk.train_poc(theta0, {"n_train": 100000, "epochs": 5000, "lr": 8e-3}, seed=seed)
        t1 = time.time()
        start2 = saved1 if saved1 is not None else last1       # loadModel reads the saved best model (main.py:325-329)
        last2, saved2, loss2 = pk.train_poc(start2, {"n_train": 100000, "epochs": 2000, "lr": 5e-4}, freezeUnits=True,
                                            seed=seed + 1000)
        t2 = time.time()
        final = saved2 if saved2 is not None else last2
        row = {"seed": seed, "seconds_stage1": t1 - t0, "seconds_stage2": t2 - t1,
               "points_per_s_stage1": 1e5 * 5000 / (t1 - t0), "points_per_s_stage2": 1e5 * 2000 / (t2 - t1)}
        for tag, th, loss in (("main", start2, loss1), ("fine_tune", final, loss2)):
            E = enet(th, Rt)
            L = loss["Ltot"][:, 0]
            row[tag] = {"E_net": E.tolist(),
                        "max_abs_err_vs_exact_R_ge_1": float(np.abs(E - Eex)[Rt >= 1.0 - 1e-9].max()),
                        "max_abs_err_vs_exact_R_ge_2": float(np.abs(E - Eex)[Rt >= 2.0 - 1e-9].max()),
                        "max_abs_diff_vs_authors_table_R_ge_1": float(np.abs(E - Eref)[Rt >= 1.0 - 1e-9].max()),
                        "Ltot_min": float(L.min()), "Ltot_argmin": int(L.argmin()),
                        "Ltot_tail_mean_last_100": float(L[-100:].mean()), "Ltot_last": float(L[-1]),
                        "all_finite": bool(np.all(np.isfinite(L)))}
        # quadrature energy of the trained wavefunction on the reference's grid (n_test = 80) at a few R
        row["E_int_n80"] = {str(Rv): pk.analysis.energy_from_psi(final, float(Rv))[0] for Rv in (1.0, 2.0, 3.0)}
        # Where the fine-tune stage is heading: with the wavefunction frozen, the E(R) that minimises mean(res^2) is the
        # Rayleigh quotient <psi|H|psi>/<psi|psi> of that wavefunction.  2000 steps @ 5e-4 do not get there (neither in the
        # authors' run: their E_net moves from -0.8165 to -0.7930 at R = 2 while their E_int is -0.7887); 40 000 steps do.
        if seed == seeds[0]:
            t3 = time.time()
            _, saved3, loss3 = pk.train_poc(start2, {"n_train": 100000, "epochs": 40000, "lr": 5e-4}, freezeUnits=True,
                                            seed=seed + 2000)
            th3 = saved3 if saved3 is not None else _
            Rs = (1.0, 1.5, 2.0, 2.5, 3.0, 3.5)
            Ebig = [pk.analysis.grid_sums(th3, Rv, n=400) for Rv in Rs]
            row["long_fine_tune_40000"] = {
                "seconds": time.time() - t3, "R": list(Rs), "E_net": enet(th3, np.array(Rs)).tolist(),
                "E_int_400cube": [b["psiHpsi"] / b["psi2"] for b in Ebig],
                "E_net_after_2000": enet(final, np.array(Rs)).tolist(), "E_exact": [float(Eex[np.argmin(np.abs(Rt - Rv))]) for Rv in Rs],
                "Ltot_tail_mean_last_100": float(loss3["Ltot"][-100:, 0].mean())}
        runs.append(row)
    rr = np.load(os.path.join(GOLD, "poc_paper_run_seed0.npz"))   # the real reference, re-run on the CPU (make_paper_run.py)
    rerun = {}
    for tag in ("stage1", "stage2"):
        e = np.abs(rr["E_net_" + tag] - Eex)
        L = rr["loss1_Ltot" if tag == "stage1" else "loss2_Ltot"]
        rerun[tag] = {"max_abs_err_vs_exact_R_ge_1": float(e[Rt >= 1.0 - 1e-9].max()), "max_abs_err_vs_exact_R_ge_2": float(e[Rt >= 2.0 - 1e-9].max()),
                      "Ltot_min": float(L.min()), "Ltot_tail_mean_last_100": float(L[-100:].mean()), "E_net": rr["E_net_" + tag].tolist()}
    rerun["seconds_on_cpu"] = rr["seconds"].tolist()
    ref = {"reference_rerun_seed0": rerun, "authors_table_max_abs_err_vs_exact_R_ge_1": float(np.abs(Eref - Eex)[Rt >= 1.0 - 1e-9].max()),
           "authors_table_max_abs_err_vs_exact_R_ge_2": float(np.abs(Eref - Eex)[Rt >= 2.0 - 1e-9].max()),
           "authors_loss_tail_main": 7.42e-07, "authors_loss_tail_fine_tune": 4.69e-07, "R": Rt.tolist(),
           "E_exact": Eex.tolist(), "E_net_authors": Eref.tolist()}
    return {"reference": ref, "runs": runs}


def main():
    parts = [a for a in sys.argv[1:] if a in ("A", "B", "C")] or ["A", "B", "C"]
    seeds = [0, 1, 2]
    for i, a in enumerate(sys.argv):
        if a == "--seeds":
            seeds = [int(v) for v in sys.argv[i + 1].split(",")]
    os.makedirs(OUT, exist_ok=True)
    res = {"device": torch.cuda.get_device_name(0), "torch": torch.__version__}
    if "A" in parts:
        res["A_config1_host_loop_fused_op"] = part_a()
    if "B" in parts:
        res["B_config1_device_trainer"] = part_b()
    if "C" in parts:
        res["C_paper_schedule_device"] = part_c(seeds)
    with open(os.path.join(OUT, "acceptance.json"), "w") as f:
        json.dump(res, f, indent=1)
    # human-readable summary
    lines = []
    a = res.get("A_config1_host_loop_fused_op")
    if a:
        lines.append("A. train.py n=4096, 200 epochs, fused op in the reference loop vs float64 oracle loop")
        for d in a["drift"]:
            lines.append("   step %3d: max|dE(R)| %.3e Ha   max|dtheta| %.3e" % (d["step"], d["max_abs_dE_R"], d["max_abs_dtheta"]))
        lines.append("   best model vs the real script's model.bin: max|dE(R)| %.3e Ha, max|dtheta| %.3e"
                     % (a["best_model_max_abs_dE_R"], a["best_model_max_abs_dtheta"]))
        lines.append("   Ltot history rel diff: max %.3e median %.3e" % (a["loss_history_rel_diff"]["Ltot_max"],
                                                                          a["loss_history_rel_diff"]["Ltot_median"]))
        for x, y in zip(a["trace_fused"], a["trace_reference"]):
            lines.append("   fused %-60s | reference %s" % (x, y))
    b = res.get("B_config1_device_trainer")
    if b:
        lines.append("B. same run through the device-resident trainer: Ltot rel diff max %.3e median %.3e; best loss %.5e (ref %.5e); "
                     "best-model max|dE(R)| %.3e Ha" % (b["Ltot_rel_diff_max"], b["Ltot_rel_diff_median"], b["best_loss"],
                                                          b["best_loss_reference"], b["best_model_max_abs_dE_R"]))
    c = res.get("C_paper_schedule_device")
    if c:
        r = c["reference"]
        lines.append("C. paper schedule on the device (1e5 points x 5000 @ 8e-3, fine-tune 2000 @ 5e-4)")
        lines.append("   authors' table: max|E_net - exact| %.2e (R>=1) %.2e (R>=2); loss tails %.2e / %.2e"
                     % (r["authors_table_max_abs_err_vs_exact_R_ge_1"], r["authors_table_max_abs_err_vs_exact_R_ge_2"],
                        r["authors_loss_tail_main"], r["authors_loss_tail_fine_tune"]))
        for tag, nm in (("stage1", "main"), ("stage2", "fine_tune")):
            t = r["reference_rerun_seed0"][tag]
            lines.append("   REAL reference re-run on the CPU, seed 0, %-9s: max|E_net - exact| %.2e (R>=1) %.2e (R>=2); Ltot min %.2e, tail(100) %.2e; %.0f s"
                         % (nm, t["max_abs_err_vs_exact_R_ge_1"], t["max_abs_err_vs_exact_R_ge_2"], t["Ltot_min"], t["Ltot_tail_mean_last_100"],
                            r["reference_rerun_seed0"]["seconds_on_cpu"][0 if tag == "stage1" else 1]))
        for run in c["runs"]:
            for tag in ("main", "fine_tune"):
                t = run[tag]
                lines.append("   seed %d %-9s: max|E_net - exact| %.2e (R>=1) %.2e (R>=2); Ltot min %.2e @%d, tail(100) %.2e; "
                             "%.2f s" % (run["seed"], tag, t["max_abs_err_vs_exact_R_ge_1"], t["max_abs_err_vs_exact_R_ge_2"],
                                         t["Ltot_min"], t["Ltot_argmin"], t["Ltot_tail_mean_last_100"],
                                         run["seconds_stage1" if tag == "main" else "seconds_stage2"]))
            lines.append("   seed %d E_int(n_test=80): %s" % (run["seed"], {k: round(float(v), 5) for k, v in run["E_int_n80"].items()}))
            lf = run.get("long_fine_tune_40000")
            if lf:
                lines.append("   seed %d, fine-tune continued to 40 000 steps (%.1f s): E_net converges to the Rayleigh quotient of the frozen psi"
                             % (run["seed"], lf["seconds"]))
                lines.append("      R     exact    E_net(2000)  E_net(40000)  E_int(400^3 grid)")
                for i, Rv in enumerate(lf["R"]):
                    lines.append("      %.1f  %.4f  %.5f     %.5f      %.5f" % (Rv, lf["E_exact"][i], lf["E_net_after_2000"][i],
                                                                               lf["E_net"][i], lf["E_int_400cube"][i]))
        run = c["runs"][0]
        lines.append("   R      exact     authors   ref. re-run  this(seed %d)" % run["seed"])
        for Rv, ex, au, rrun, me in zip(r["R"], r["E_exact"], r["E_net_authors"], r["reference_rerun_seed0"]["stage2"]["E_net"], run["fine_tune"]["E_net"]):
            lines.append("   %.1f  %.4f  %.5f  %.5f     %.5f" % (Rv, ex, au, rrun, me))
    txt = "\n".join(lines)
    open(os.path.join(OUT, "acceptance.txt"), "w").write(txt + "\n")
    print(txt)


if __name__ == "__main__":
    main()
