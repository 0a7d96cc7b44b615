"""Data-parallel step: collocation points shard across ranks, one small all-reduce.

Two ways to do the sum over ranks:
  * ``attach_fused(handle, group)``: the library's own exchange, FUSED into the reduction kernel of the training evaluation
    (peer stores over NVLink + release/acquire flags, include/pinn_b200.h pinn_dp_*).  After it every
    ``ops.loss_and_grad_raw`` / ``pinn_loss_fwd_bwd[_host]`` on that handle returns the GLOBAL sums; no separate collective
    is launched.  torch.distributed is used once, to all-gather the 64-byte IPC handles.
  * ``dp_loss_and_grad``: local evaluation followed by ``torch.distributed.all_reduce`` (NCCL on GPUs, gloo in the CPU
    tests) - the baseline the fused exchange is measured against (bench.py --allreduce nccl).

Each rank evaluates its shard with the GLOBAL weights {1/n, 1/|set1|, 1/|set2|}, so the shard
results (loss terms, sums, dLtot/dtheta) simply add; one ``all_reduce(SUM)`` over a fused
float64 buffer of 8 + 1521 scalars finishes the step.  The global set sizes come from one tiny
all-reduce of the local counts (or from the caller, who usually knows them from the sampler).
No other collective exists on this path: the model has 1521 parameters, so replicas + point sharding
is the only parallelism.
"""
import torch
import torch.distributed as dist

from . import params as P

N_OUT = 8 + P.N_THETA


def global_weights(n_local, c1_local, c2_local, group=None, device="cpu"):
    """All-reduce the local point/set counts -> float64 tensor {1/n, 1/c1, 1/c2} (global)."""
    c = torch.tensor([float(n_local), float(c1_local), float(c2_local)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / c


def dp_loss_and_grad(local_step, weights, group=None, out=None):
    """Run `local_step(weights, out)` on this rank's shard and all-reduce the fused result.

    local_step must fill `out` (float64, 8+1521: sums then dLtot/dtheta) for the shard using the global
    `weights`.  On the GPU path it is a closure over ``ops.loss_and_grad_raw`` writing into views of `out`;
    CPU tests inject the oracle.  Returns (sums[8], dtheta[1521]) views of the reduced buffer; sums[7]
    (E of the last point) is rank-local information and is meaningless after the reduction."""
    if out is None:
        out = torch.empty(N_OUT, dtype=torch.float64, device=weights.device)
    local_step(weights, out)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out[:8], out[8:]


def make_gpu_local_step(variant, x, y, z, R, theta, mask=None, grad_mask=0xFFFF):
    """Closure for dp_loss_and_grad on CUDA shards (tensors as in ops.loss_and_grad_raw)."""
    from . import ops

    def step(weights, out):
        ops.loss_and_grad_raw(variant, x, y, z, R, theta, mask, weights, grad_mask, sums=out[:8], dtheta=out[8:])

    return step


def attach_fused(handle, group=None, strict=True):
    """Connect `handle` (one per rank/GPU) to the other ranks of `group` for the fused exchange and enable it.

    Collective: every rank of the group must call it.  The ranks agree on the outcome: if mapping the peers' buffers
    fails on ANY rank (no CUDA IPC between the processes, no peer access), every rank tears its side down again and the
    call raises (strict) or returns 0, so that the caller can fall back to ``dp_loss_and_grad`` on all ranks together.
    Returns the world size on success."""
    import torch
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    err = None
    try:
        mine = handle.dp_init(rank, world)
    except Exception as e:  # noqa: BLE001 - reported to all ranks below
        mine, err = b"\0" * handle.DP_HANDLE_BYTES, e
    allh = [None] * world
    dist.all_gather_object(allh, mine, group=group)
    if err is None:
        try:
            handle.dp_connect(allh)
        except Exception as e:  # noqa: BLE001
            err = e
    dev = torch.device("cuda", handle.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)   # also the barrier: nobody exchanges before all have mapped
    if int(ok.item()) == 1:
        return world
    try:
        handle.dp_shutdown()
    except Exception:  # noqa: BLE001 - nothing was set up on this rank
        pass
    if strict:
        raise RuntimeError("fused data-parallel exchange could not be set up on every rank" + (": %s" % err if err else ""))
    return 0


def detach_fused(handle, group=None):
    """Collective teardown of the fused exchange."""
    import torch
    torch.cuda.synchronize()
    dist.barrier(group)
    handle.dp_shutdown()
