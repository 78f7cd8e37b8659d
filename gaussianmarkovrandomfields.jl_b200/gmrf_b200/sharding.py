"""One-process-per-GPU sharding of the two workloads of the path that shard naturally:

* a batch of independent hyperparameter evaluations (theta -> Q(theta) -> refactorize -> logdet, quadratic form): the
  reference runs them through a `WorkspacePool` of replicas that share ONE resolved ordering
  (src/workspace/workspace_pool.jl:42-67, "resolve the permutation ONCE" :55-58). Here every rank owns one workspace
  on its GPU and a contiguous block of the batch; the only exchange is an all-gather of the per-point scalars.
* blocks of right-hand sides / samples against ONE factorization (`backend_solve(b, RHS::Matrix)` backend.jl:207-209,
  column-at-a-time `_rand!` workspace_gmrf.jl:275-286): rank `src` factorizes, the numeric factor (supernodal panels +
  inverted diagonal blocks) is broadcast to the peers over NCCL/NVLink, every rank solves its own block of columns.

A single factorization is never split across GPUs (north star). The collectives go through `torch.distributed`
(`nccl` on the GPU box, `gloo` in the CPU tests); tensors live on the device the process group needs.
"""
from __future__ import annotations

import numpy as np

__all__ = ["shard_range", "shard_counts", "all_gather_blocks", "sharded_map", "sharded_columns", "broadcast_factor"]


def shard_counts(n_items: int, world: int) -> list:
    """Sizes of the contiguous blocks: the first n_items % world ranks take one extra item."""
    base, extra = divmod(int(n_items), int(world))
    return [base + (1 if r < extra else 0) for r in range(world)]


def shard_range(n_items: int, world: int, rank: int):
    """[lo, hi) of the block of `rank` (contiguous blocks in rank order, so a gather is a concatenation)."""
    counts = shard_counts(n_items, world)
    lo = sum(counts[:rank])
    return lo, lo + counts[rank]


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def _group_device(group=None):
    import torch
    import torch.distributed as dist
    backend = dist.get_backend(group)
    return torch.device("cuda", torch.cuda.current_device()) if "nccl" in str(backend) else torch.device("cpu")


def all_gather_blocks(local: np.ndarray, n_items: int, group=None) -> np.ndarray:
    """Concatenate per-rank blocks (leading axis = items of that rank) in rank order; every rank gets the full array.
    Ragged blocks are padded to the largest one for the collective and trimmed afterwards."""
    import torch
    local = np.ascontiguousarray(local, dtype=np.float64)
    dist = _dist()
    if dist is None or dist.get_world_size(group) == 1:
        if local.shape[0] != n_items:
            raise ValueError("single-rank gather: the local block must hold every item")
        return local
    world = dist.get_world_size(group)
    counts = shard_counts(n_items, world)
    rank = dist.get_rank(group)
    if local.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} items, expected {counts[rank]}")
    tail = local.shape[1:]
    width = int(np.prod(tail)) if tail else 1
    cmax = max(counts)
    dev = _group_device(group)
    send = torch.zeros((cmax, width), dtype=torch.float64, device=dev)
    if counts[rank]:
        send[:counts[rank]] = torch.from_numpy(local.reshape(counts[rank], width)).to(dev)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    parts = [recv[r][:counts[r]].cpu().numpy() for r in range(world)]
    return np.concatenate(parts, axis=0).reshape((n_items,) + tail)


def sharded_map(evaluate, items, group=None) -> np.ndarray:
    """Evaluate `evaluate(item) -> float | 1-D array` on this rank's contiguous block of `items` and all-gather the
    results in item order. `evaluate` typically closes over this rank's workspace:
        lambda th: (ws.update_precision_values(model.values(*th)), ws.logdet())[1]"""
    dist = _dist()
    world = dist.get_world_size(group) if dist else 1
    rank = dist.get_rank(group) if dist else 0
    lo, hi = shard_range(len(items), world, rank)
    vals = [np.atleast_1d(np.asarray(evaluate(items[i]), dtype=np.float64)) for i in range(lo, hi)]
    width = None
    if vals:
        width = vals[0].size
    if dist and world > 1:       # ranks with an empty block still need the item width
        import torch
        w = torch.tensor([width or 0], dtype=torch.int64, device=_group_device(group))
        dist.all_reduce(w, op=dist.ReduceOp.MAX, group=group)
        width = int(w.item())
    local = np.stack(vals) if vals else np.zeros((0, width or 1))
    out = all_gather_blocks(local, len(items), group)
    return out[:, 0] if out.shape[1] == 1 else out


def sharded_columns(solve, B: np.ndarray, group=None) -> np.ndarray:
    """X = solve(B) with the COLUMNS of B (right-hand sides / white-noise draws) sharded over the ranks: each rank
    calls `solve` on its contiguous column block (one blocked multi-RHS call) and the blocks are all-gathered."""
    B = np.asarray(B, dtype=np.float64)
    n, m = B.shape
    dist = _dist()
    world = dist.get_world_size(group) if dist else 1
    rank = dist.get_rank(group) if dist else 0
    lo, hi = shard_range(m, world, rank)
    Xl = solve(np.asfortranarray(B[:, lo:hi])) if hi > lo else np.zeros((n, 0))
    out = all_gather_blocks(np.ascontiguousarray(np.asarray(Xl).T), m, group)     # items = columns
    return np.asfortranarray(out.T)


class _DeviceArray:
    """Minimal __cuda_array_interface__ carrier so torch can wrap a library-owned device buffer without a copy."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def broadcast_factor(backend, src: int = 0, group=None, with_selinv: bool = False) -> None:
    """Make the numeric factor of rank `src` the factor of every rank's backend (same pattern, same ordering):
    NCCL broadcast of the supernodal panels and the inverted diagonal blocks straight out of / into the handles' HBM,
    then `adopt_factor` on the receivers. Needs the nccl backend (device buffers)."""
    import torch
    dist = _dist()
    if dist is None or dist.get_world_size(group) == 1:
        return
    dev = torch.device("cuda", backend.device)
    is_src = dist.get_rank(group) == src
    # header first: the sender's analysis fingerprint (split into two exactly representable doubles), its pivot status
    # and log-determinant -- a peer created with another ordering / other options must refuse the panels, not solve
    # with them, and a failed factorization must not look healthy on the receivers
    fp = backend.analysis_fingerprint()
    head = torch.tensor([float(fp >> 32), float(fp & 0xFFFFFFFF), float(backend.status), backend.compute_logdet()] if is_src
                        else [0.0, 0.0, 0.0, 0.0], dtype=torch.float64, device=dev)
    dist.broadcast(head, src=src, group=group)
    hi, lo, status, logdet = head.tolist()
    sender_fp = (int(hi) << 32) | int(lo)
    if not is_src and sender_fp != fp:
        raise ValueError("broadcast_factor: this rank's symbolic analysis (pattern / ordering / supernodes / options) differs "
                         "from the sender's; create every handle from the same ordering or analysis blob")
    which = [0, 1] + ([2] if with_selinv else [])
    for w in which:
        ptr, n = backend.device_array(w)
        t = torch.as_tensor(_DeviceArray(ptr, n), device=dev)
        dist.broadcast(t, src=src, group=group)
    torch.cuda.synchronize(dev)
    if not is_src:
        backend.adopt_factor(float(logdet), with_selinv, fingerprint=sender_fp, status=int(status))
