"""Sharding of the two data-parallel axes of the hot path over the GPUs of one node.

The reference evaluates a GA/PSO population in one Python loop (krige.py:436-451) and a swarm /
plot grid of query points in another (krige.py:243-257, 531-537).  Both are independent units given
replicated ``(X, y)`` / a replicated fitted state, so with one process per GPU (``torchrun``):

  * population: rank r evaluates candidates ``[r*ceil(C/W), (r+1)*ceil(C/W))`` on its GPU and ONE
    ``all_gather`` of the per-candidate scalars (NegLnLike, mu, SigmaSqr, LnDetPsi, info -- 40 bytes per
    candidate) gives every rank the full fitness vector.  GA/PSO state is evolved redundantly and
    deterministically on every rank from the same seed, so nothing else is exchanged;
  * query grid: rank r predicts rows ``[lo, hi)`` of the query array; the arg-max of the expected
    improvement (what ``infill`` needs) is one ``all_gather`` of a (value, index) pair per rank.

There is no data-path collective: only scalars travel (NCCL over NVLink on GPUs, gloo in the CPU
tests).  This module contains host logic only; the per-shard work is whatever callable the model
passes in (the C-ABI calls of ``_lib.Handle``).
"""
import numpy as np


def shard_bounds(count, rank, world):
    """Contiguous block partition: returns (lo, hi) of rank's slice of ``count`` units."""
    per = -(-int(count) // int(world)) if count > 0 else 0
    lo = min(rank * per, count)
    hi = min(lo + per, count)
    return lo, hi


def _dist():
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return None
    return dist


def _comm_device(dist, group):
    import torch

    backend = dist.get_backend(group)
    if str(backend).lower() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def all_gather_rows(local, count, group=None):
    """``local``: float64 array with this rank's rows (its ``shard_bounds`` slice of ``count`` rows,
    any number of columns).  Returns the (count, ncols) array of all ranks, identical on every rank."""
    import torch

    dist = _dist()
    local = np.ascontiguousarray(np.asarray(local, dtype=np.float64))
    if local.ndim == 1:
        local = local[:, None]
    if dist is None or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    per = -(-int(count) // world)
    ncols = local.shape[1]
    dev = _comm_device(dist, group)
    buf = torch.zeros(per, ncols, dtype=torch.float64, device=dev)
    if local.shape[0]:
        buf[: local.shape[0]] = torch.from_numpy(local).to(dev)
    out = torch.empty(world * per, ncols, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, buf, group=group)
    return out[:count].cpu().numpy()


def sharded_rows(count, eval_slice, group=None):
    """Evaluate ``eval_slice(lo, hi) -> (hi-lo, ncols) float64`` on this rank's slice and gather."""
    dist = _dist()
    if dist is None or dist.get_world_size(group) == 1:
        return np.asarray(eval_slice(0, count), dtype=np.float64).reshape(count, -1)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(count, rank, world)
    local = eval_slice(lo, hi) if hi > lo else np.zeros((0, 1))
    local = np.asarray(local, dtype=np.float64)
    if hi > lo:
        local = local.reshape(hi - lo, -1)
    ncols = _agree_ncols(local.shape[1] if hi > lo else 0, group)
    if hi <= lo:
        local = np.zeros((0, ncols))
    return all_gather_rows(local, count, group)


def _agree_ncols(ncols, group):
    import torch

    dist = _dist()
    t = torch.tensor([ncols], dtype=torch.int64, device=_comm_device(dist, group))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def global_argmax(local_values, lo, group=None):
    """Arg-max over a sharded 1-D array: ``local_values`` are this rank's entries, starting at global
    index ``lo``.  Returns (value, global_index), identical on every rank; ties resolve to the lowest
    global index, as ``max()`` over the reference's in-order Python list would."""
    import torch

    local_values = np.asarray(local_values, dtype=np.float64)
    if local_values.size:
        i = int(np.argmax(local_values))
        pair = np.array([[local_values[i], float(lo + i)]])
    else:
        pair = np.array([[-np.inf, np.inf]])
    dist = _dist()
    if dist is None or dist.get_world_size(group) == 1:
        return float(pair[0, 0]), int(pair[0, 1])
    world = dist.get_world_size(group)
    dev = _comm_device(dist, group)
    out = torch.empty(world, 2, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, torch.from_numpy(pair).to(dev), group=group)
    pairs = out.cpu().numpy()
    best = np.max(pairs[:, 0])
    idx = int(np.min(pairs[pairs[:, 0] == best, 1]))
    return float(best), idx


def broadcast_candidate(values, src=0, group=None):
    """Safety resync of an elite (theta, p[, lambda]) vector (2k[+1] doubles) from ``src``."""
    import torch

    dist = _dist()
    values = np.ascontiguousarray(np.asarray(values, dtype=np.float64))
    if dist is None or dist.get_world_size(group) == 1:
        return values
    t = torch.from_numpy(values.copy()).to(_comm_device(dist, group))
    dist.broadcast(t, src=src, group=group)
    return t.cpu().numpy()
