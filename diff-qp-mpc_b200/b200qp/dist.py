"""Batch sharding across the GPUs of one box (one process per GPU, torch.distributed).

Every QP / MPC problem of a batch is independent in its arithmetic, so the
data path has NO collective: rank r solves the contiguous slice `shard_slice(nb, r, world)`.
Two places exchange data:

  * parameters SHARED across the batch get the mean of the per-problem gradients over the GLOBAL
    batch (qpth/qp.py:160-178 uses .mean(0)); with a sharded batch that is one SUM all-reduce of
    the per-rank sums, `allreduce_shared_grad`;
  * the reference's termination test and get_step fill are reductions over the whole batch
    (qpth/solvers/pdipm/batch.py:127-131,141,213).  Sharded runs therefore equal "the reference
    invoked per shard" (the default, parity is defined per shard).

The backend is whatever the process group was created with (NCCL on the B200 box; gloo in the
CPU tests of this host logic).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_slice(nb: int, rank: int, world: int) -> slice:
    """Contiguous, balanced slice of a batch of nb problems owned by `rank` (first nb % world
    ranks get one extra problem)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(nb, world)
    lo = rank * base + min(rank, extra)
    return slice(lo, lo + base + (1 if rank < extra else 0))


def shard_params(params, nb: int, rank: int, world: int, batched_ndim=(3, 2, 3, 2, 3, 2)):
    """Slice the batched members of (Q, p, G, h, A, b); un-batched (shared) members pass through."""
    sl = shard_slice(nb, rank, world)
    out = []
    for t, nd in zip(params, batched_ndim):
        out.append(t[sl] if t.dim() == nd and t.size(0) == nb else t)
    return tuple(out)


def allreduce_shared_grad(local_mean: torch.Tensor, local_nb: int, group=None) -> torch.Tensor:
    """Mean over the GLOBAL batch of a shared parameter's gradient, given this rank's mean over
    its own `local_nb` problems: sum_r(local_mean_r * nb_r) / sum_r nb_r, one all-reduce."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_mean
    buf = torch.cat([(local_mean * float(local_nb)).reshape(-1),
                     torch.tensor([float(local_nb)], dtype=local_mean.dtype, device=local_mean.device)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return (buf[:-1] / buf[-1]).reshape(local_mean.shape)


def gather_batch(local: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenate per-rank result slices along the batch dimension (ragged shards allowed)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    mx = int(max(int(s) for s in sizes))
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: int(s)] for p, s in zip(parts, sizes)], dim=0)


def allreduce_gradients(params, group=None) -> None:
    """Data-parallel training of the policy network (deqmpc/train.py:165-175: loss.backward() ... optimizer.step()): average
    the gradients of `params` over the ranks with ONE all-reduce of a flat bucket (the DEQLayer has ~50 k parameters: one
    NCCL launch over NVLink instead of one per tensor), written back in place."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(float(dist.get_world_size(group)))
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
