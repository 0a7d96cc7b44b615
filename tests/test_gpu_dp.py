"""Data-parallel exchange fused into the reduction kernel (include/pinn_b200.h: pinn_dp_*).

Needs two B200s in the box; on a one-GPU box only the single-rank and argument-error cases run.  Through the C ABI,
compared with the single-GPU evaluation of the whole batch and with the float64 oracle.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import pinn_for_quantum_wavefunction_surfaces_b200 as pk
from oracle import closed_form as cf
from oracle import ref_autograd as ra

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def two_gpus():
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


def sample(n, seed):
    g = torch.Generator().manual_seed(seed)
    x, y, z, R, i1, i2 = ra.sample_box(n, "poc", g)
    a32 = [v.numpy().ravel().astype(np.float32) for v in (x, y, z, R)]
    m = np.zeros(n, np.uint8)
    m[i1.numpy()] |= 1
    m[i2.numpy()] |= 2
    return a32, m


def test_world_of_one_is_the_plain_evaluation(golden_dir):
    """dp with world = 1 must not change anything (and must not wait for anybody)."""
    th = np.load(os.path.join(golden_dir, "trainpy_n2048.npz"))["theta"].astype(np.float32)
    a32, m = sample(5000, 3)
    d = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(d)
    w = torch.tensor([1.0 / 5000, 1.0 / max((m & 1).sum(), 1), 1.0 / max((m >> 1).sum(), 1)], dtype=torch.float64, device=d)
    h = pk.Handle.get(0)
    s0, g0, _ = pk.loss_and_grad_raw(0, *[t(a) for a in a32], t(th), t(m), w)
    s0, g0 = s0.cpu().numpy(), g0.cpu().numpy()
    h.dp_init(0, 1)
    h.dp_connect([b"\0" * 64])
    try:
        s1, g1, _ = pk.loss_and_grad_raw(0, *[t(a) for a in a32], t(th), t(m), w)
        assert np.array_equal(s1.cpu().numpy(), s0) and np.array_equal(g1.cpu().numpy(), g0)
        assert h.dp_status() == 0   # no exchange happened
    finally:
        h.dp_shutdown()


def test_dp_argument_errors():
    h = pk.Handle.get(0)
    with pytest.raises(pk.PinnError):
        h.dp_init(3, 2)
    with pytest.raises(pk.PinnError):
        h.dp_init(0, 9)
    with pytest.raises(pk.PinnError):
        h.dp_enable(True)   # not initialised


def test_a_peer_that_never_delivers_is_reported_and_stops_the_optimizer(golden_dir):
    """World of 2 whose second rank never runs (two handles on ONE device suffice: rank 1 stays idle).  Rank 0 polls for
    the configured time-out, gives up instead of hanging the GPU, and from then on: pinn_dp_status and Trainer.read()
    report PINN_ETIMEDOUT, later steps return at once (no second wait), and the fused Adam leaves the replica untouched."""
    import time
    th = np.load(os.path.join(golden_dir, "trainpy_n2048.npz"))["theta"]
    h0, h1 = pk.Handle(0), pk.Handle(0)
    try:
        h0.dp_init(0, 2, want_ipc=False)
        h1.dp_init(1, 2, want_ipc=False)
        h0.dp_connect_local([h0, h1])
        h1.dp_connect_local([h0, h1])
        h0.dp_set_timeout(0.25)
        import ctypes
        th64 = np.ascontiguousarray(th, np.float64)
        tr0 = pk.Trainer("poc", 4096, th, seed=3, lr=8e-3, history_capacity=4, handle=h0)
        t0 = time.time()
        tr0.run(1)
        with pytest.raises(pk.PinnError, match="did not deliver"):
            tr0.read()
        first = time.time() - t0
        assert 0.1 < first < 5.0                     # waited about the time-out (the sampler's set-size exchange, then the sums)
        t0 = time.time()
        tr0.run(3)                                   # poisoned: no further waiting
        with pytest.raises(pk.PinnError, match="did not deliver"):
            tr0.read()
        assert time.time() - t0 < 0.2
        with pytest.raises(pk.PinnError):
            h0.dp_status()
        theta_now = np.empty(1521)
        h0.L.pinn_trainer_read(tr0.t, theta_now.ctypes.data_as(ctypes.c_void_p), None, None, None, None, None, 0)
        assert np.array_equal(theta_now, th64)       # no optimizer step was applied with incomplete sums
        tr0.close()
    finally:
        h0.dp_shutdown(); h1.dp_shutdown()
        h0.close(); h1.close()


@pytest.mark.skipif(not two_gpus(), reason="needs two GPUs")
@pytest.mark.parametrize("n", [4096, 100003])
def test_two_ranks_in_one_process_equal_the_whole_batch(golden_dir, n):
    """Two handles on two devices, peers connected directly: each evaluates one shard with the GLOBAL weights; both
    must end up with bit-identical results that equal the one-GPU evaluation of the whole batch (to the rounding of
    a different partition) and the float64 oracle."""
    th = np.load(os.path.join(golden_dir, "trainpy_n2048.npz"))["theta"].astype(np.float32)
    a32, m = sample(n, 11)
    c1, c2 = float((m & 1).sum()), float((m >> 1).sum())
    wv = [1.0 / n, 1.0 / c1, 1.0 / c2]
    devs = [torch.device("cuda:0"), torch.device("cuda:1")]
    hs = [pk.Handle.get(0), pk.Handle.get(1)]
    # whole batch on GPU 0, exchange off
    t0 = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(devs[0])
    sw, gw, _ = pk.loss_and_grad_raw(0, *[t0(a) for a in a32], t0(th), t0(m), torch.tensor(wv, dtype=torch.float64, device=devs[0]))
    sw, gw = sw.cpu().numpy(), gw.cpu().numpy()
    for r in range(2):
        hs[r].dp_init(r, 2, want_ipc=False)
    try:
        for r in range(2):
            hs[r].dp_connect_local(hs)
        cut = n // 2 + 17
        shards = [slice(0, cut), slice(cut, n)]
        outs = []
        for rep in range(3):   # several exchanges: slots alternate, the step counter advances
            res = []
            for r in range(2):
                with torch.cuda.device(devs[r]):
                    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(devs[r])
                    sl = shards[r]
                    res.append(pk.loss_and_grad_raw(0, *[t(a[sl]) for a in a32], t(th), t(m[sl]),
                                                    torch.tensor(wv, dtype=torch.float64, device=devs[r])))
            for r in range(2):
                torch.cuda.synchronize(devs[r])
            outs.append([(s.cpu().numpy(), g.cpu().numpy()) for s, g, _ in res])
        assert hs[0].dp_status() == 3 and hs[1].dp_status() == 3
        for rep in range(3):
            (s_a, g_a), (s_b, g_b) = outs[rep]
            assert np.array_equal(s_a[:7], s_b[:7]) and np.array_equal(g_a, g_b)       # same bits on both ranks
            assert np.array_equal(s_a[:7], outs[0][0][0][:7]) and np.array_equal(g_a, outs[0][0][1])  # and every time
        s_a, g_a = outs[0][0]
        assert np.allclose(s_a[:7], sw[:7], rtol=2e-6)
        assert np.abs(g_a - gw).max() / np.abs(gw).max() < 1e-5
        if n <= 5000:
            m1, m2 = (m & 1).astype(np.float64), (m >> 1).astype(np.float64)
            ref = cf.loss_and_grad("poc", th.astype(np.float64), *[a.astype(np.float64) for a in a32], m1, m2)
            assert abs(s_a[0] - ref["Ltot"]) / ref["Ltot"] < 1e-5
            assert np.abs(g_a - ref["grad"]).max() / np.abs(ref["grad"]).max() < 1e-5
        # switching the exchange off gives the local shard sums again
        hs[0].dp_enable(False)
        with torch.cuda.device(devs[0]):
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(devs[0])
            sl = shards[0]
            s_l, g_l, _ = pk.loss_and_grad_raw(0, *[t(a[sl]) for a in a32], t(th), t(m[sl]),
                                               torch.tensor(wv, dtype=torch.float64, device=devs[0]))
        assert s_l.cpu().numpy()[0] < s_a[0]
    finally:
        for r in range(2):
            torch.cuda.synchronize(devs[r])
        for r in range(2):
            hs[r].dp_shutdown()


_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["PINN_ROOT"])
import pinn_for_quantum_wavefunction_surfaces_b200 as pk
from pinn_for_quantum_wavefunction_surfaces_b200 import dp
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
d = np.load(os.environ["PINN_CASE"])
n = d["x"].size
lo, hi = rank * n // world, (rank + 1) * n // world
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
h = pk.Handle.get(rank)
w = t(d["w"])
args = [t(d[k][lo:hi]) for k in "xyzR"]
# baseline: local evaluation + NCCL all-reduce
out = torch.empty(dp.N_OUT, dtype=torch.float64, device=dev)
pk.loss_and_grad_raw(0, *args, t(d["theta"]), t(d["mask"][lo:hi]), w, sums=out[:8], dtheta=out[8:])
dist.all_reduce(out)
ref = out.cpu().numpy().copy()
dp.attach_fused(h)
for rep in range(4):
    fo = torch.empty(dp.N_OUT, dtype=torch.float64, device=dev)
    pk.loss_and_grad_raw(0, *args, t(d["theta"]), t(d["mask"][lo:hi]), w, sums=fo[:8], dtheta=fo[8:])
torch.cuda.synchronize()
got = fo.cpu().numpy()
assert h.dp_status() == 4
gather = [torch.empty_like(fo) for _ in range(world)]
dist.all_gather(gather, fo)
for g in gather:
    assert torch.equal(g[:7], fo[:7]) and torch.equal(g[8:], fo[8:]), "ranks disagree"
err_s = np.abs(got[:7] - ref[:7]).max() / np.abs(ref[:7]).max()
err_g = np.abs(got[8:] - ref[8:]).max() / np.abs(ref[8:]).max()
assert err_s < 1e-12 and err_g < 1e-12, (err_s, err_g)   # same shard rows, both sums in float64
dp.detach_fused(h)
dist.destroy_process_group()
if rank == 0:
    print("DP_OK", err_s, err_g)
"""


@pytest.mark.skipif(not two_gpus(), reason="needs two GPUs")
def test_two_processes_over_ipc_match_nccl_allreduce(tmp_path, golden_dir):
    """One process per GPU (torchrun), exchange buffers shared through CUDA IPC: the fused exchange must give what
    local evaluation + NCCL all-reduce gives, identically on every rank."""
    th = np.load(os.path.join(golden_dir, "trainpy_n2048.npz"))["theta"].astype(np.float32)
    n = 50000
    a32, m = sample(n, 23)
    w = np.array([1.0 / n, 1.0 / (m & 1).sum(), 1.0 / (m >> 1).sum()])
    case = tmp_path / "case.npz"
    np.savez(case, x=a32[0], y=a32[1], z=a32[2], R=a32[3], mask=m, theta=th, w=w)
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, PINN_ROOT=ROOT, PINN_CASE=str(case))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "DP_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.skipif(not two_gpus(), reason="needs two GPUs")
@pytest.mark.parametrize("use_graph", [True, False])
def test_device_resident_training_loop_data_parallel(golden_dir, use_graph):
    """Two device-resident trainers (one per GPU, exchange connected) with n/2 points each draw exactly the batch a single
    trainer with n points draws (Philox counter = global point index), see the global set sizes (count exchange in the
    sampler) and the global gradient (exchange in the reduction kernel): identical replicas on both ranks, and the same
    trajectory as the one-GPU run up to the rounding of a different summation order (which it could not be if the
    shards were not the two halves of the single-GPU batch: the loss of a fresh batch differs by O(10 %))."""
    th0 = np.load(os.path.join(golden_dir, "trainpy_n2048.npz"))["theta"]
    n, steps, seed = 8192, 8, 777
    one = pk.Trainer("trainpy", n, th0, seed=seed, lr=8e-3, history_capacity=steps, device=0)
    one.run(steps, use_graph=use_graph)
    ref = one.read()
    one.close()
    hs = [pk.Handle.get(0), pk.Handle.get(1)]
    for r in range(2):
        hs[r].dp_init(r, 2, want_ipc=False)
    trs = []
    try:
        for r in range(2):
            hs[r].dp_connect_local(hs)
        # one host thread per rank, as separate processes would be (the first graph step synchronises its stream, which
        # needs the peer to be running: ctypes releases the GIL during the calls)
        import threading
        trs, out, errs = [None, None], [None, None], []

        def rank_main(r):
            try:
                trs[r] = pk.Trainer("trainpy", n // 2, th0, seed=seed, lr=8e-3, history_capacity=steps, device=r)
                trs[r].run(steps, use_graph=use_graph)
                out[r] = trs[r].read()
            except Exception as e:  # noqa: BLE001 - reported below
                errs.append((r, e))

        th = [threading.Thread(target=rank_main, args=(r,)) for r in range(2)]
        for t in th:
            t.start()
        for t in th:
            t.join(120)
        assert not errs, errs
        assert hs[0].dp_status() == steps and hs[1].dp_status() == steps
        for k in ("theta", "m", "v", "best_theta"):
            assert np.array_equal(out[0][k], out[1][k]), k          # the replicas stay bit-identical
        assert np.array_equal(out[0]["history"][:, :3], out[1]["history"][:, :3])
        assert np.allclose(out[0]["history"][:, :3], ref["history"][:, :3], rtol=2e-5)
        assert np.allclose(out[0]["history"][:, 3], ref["history"][:, 3], rtol=1e-5)   # mean E over the GLOBAL batch
        assert np.abs(out[0]["theta"] - ref["theta"]).max() < 2e-4 * np.abs(ref["theta"]).max()
    finally:
        for t in trs:
            if t is not None:
                t.close()
        for r in range(2):
            torch.cuda.synchronize(r)
        for r in range(2):
            hs[r].dp_shutdown()
