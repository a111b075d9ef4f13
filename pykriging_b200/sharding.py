"""Sharding of the two data-parallel axes of the hot path over the GPUs of one node.

The reference evaluates a GA/PSO population in one Python loop (krige.py:436-451) and a swarm /
plot grid of query points in another (krige.py:243-257, 531-537).  Both are independent units given
replicated ``(X, y)`` / a replicated fitted state, so with one process per GPU (``torchrun``):

  * population: rank r evaluates candidates ``[r*ceil(C/W), (r+1)*ceil(C/W))`` on its GPU and ONE
    ``all_gather`` of the per-candidate scalars (NegLnLike, mu, SigmaSqr, LnDetPsi, info -- 40 bytes per
    candidate, plus one status word per rank so that a failure reaches every rank inside the same
    collective) gives every rank the full fitness vector.  On GPUs the kernels write straight into the
    send buffer and the collective runs on the handle's stream (``sharded_nll_device``).  GA/PSO state is evolved redundantly and
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


class ShardError(RuntimeError):
    """Raised on EVERY rank when the evaluation of some rank's slice failed (the failing rank re-raises its own
    exception instead): nobody is left waiting in the next collective."""


def all_gather_rows(local, count, group=None, ncols=None, status=0.0):
    """``local``: float64 array with this rank's rows (its ``shard_bounds`` slice of ``count`` rows,
    ``ncols`` columns).  ONE all_gather; returns ((count, ncols) array of all ranks, per-rank status vector),
    identical on every rank.  ``status`` travels with the rows (one extra row per rank), so that a rank whose
    evaluation failed can say so inside the same collective."""
    import torch

    dist = _dist()
    local = np.ascontiguousarray(np.asarray(local, dtype=np.float64))
    if local.ndim == 1:
        local = local[:, None]
    if ncols is None:
        ncols = local.shape[1]
    if dist is None or dist.get_world_size(group) == 1:
        return local.reshape(-1, ncols), np.array([float(status)])
    world = dist.get_world_size(group)
    per = -(-int(count) // world)
    dev = _comm_device(dist, group)
    buf = torch.zeros(per + 1, ncols, dtype=torch.float64)
    if local.shape[0]:
        buf[: local.shape[0]] = torch.from_numpy(local.reshape(-1, ncols))
    buf[per, 0] = float(status)
    buf = buf.to(dev)
    out = torch.empty(world * (per + 1), ncols, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.cpu().numpy().reshape(world, per + 1, ncols)
    rows = out[:, :per, :].reshape(world * per, ncols)[:count]
    return np.ascontiguousarray(rows), out[:, per, 0].copy()


def sharded_rows(count, eval_slice, group=None, ncols=None):
    """Evaluate ``eval_slice(lo, hi) -> (hi-lo, ncols) float64`` on this rank's slice and gather: ONE collective per
    call.  ``ncols`` must be given when a rank's slice can be empty (it cannot learn the width from its own result).
    If ``eval_slice`` raises on any rank, every rank raises (the failing one its own exception, the others
    ``ShardError``) AFTER the collective, so the process group stays usable."""
    dist = _dist()
    if dist is None or dist.get_world_size(group) == 1:
        return np.asarray(eval_slice(0, count), dtype=np.float64).reshape(count, -1)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(count, rank, world)
    err = None
    local = None
    if hi > lo:
        try:
            local = np.asarray(eval_slice(lo, hi), dtype=np.float64).reshape(hi - lo, -1)
        except Exception as e:  # noqa: BLE001 -- reported to every rank below
            err = e
    if ncols is None:
        # width unknown to the caller: every rank must have a non-empty, successfully evaluated slice
        if local is None:
            raise err if err is not None else ValueError("sharded_rows: pass ncols when a slice can be empty")
        ncols = local.shape[1]
    if local is None:
        local = np.full((hi - lo, ncols), np.nan)
    rows, status = all_gather_rows(local, count, group, ncols=ncols, status=1.0 if err is not None else 0.0)
    if err is not None:
        raise err
    if np.any(status != 0.0):
        raise ShardError("evaluation of the slice failed on rank(s) %s" % np.nonzero(status)[0].tolist())
    return rows


_NLL_BUFFERS = {}


def release_buffers():
    """Frees the cached send / receive / pinned buffers (also registered with atexit: torch's pinned-memory allocator
    records a CUDA event when such a tensor dies, which must happen while the context is still alive)."""
    _NLL_BUFFERS.clear()


import atexit  # noqa: E402

atexit.register(release_buffers)


def _nll_buffers(handle, per, world):
    """Send / receive / pinned host buffers of the sharded population path, cached per (handle, slice size, world).
    One rank's record: 4 x per doubles (NegLnLike | mu | SigmaSqr | LnDetPsi), per int32 (info), one int32 status word,
    padded to 8 bytes -- gathered as raw bytes, so the kernels' outputs go out exactly as pk_nll_batch wrote them."""
    import torch

    key = (id(handle), handle.stream, handle.device, per, world)
    buf = _NLL_BUFFERS.get(key)
    if buf is None:
        _NLL_BUFFERS.clear()  # one live configuration is all a training run needs
        dev = torch.device("cuda", handle.device)
        nbytes = (4 * per * 8 + per * 4 + 4 + 7) // 8 * 8
        loc = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        f64 = loc[: 4 * per * 8].view(torch.float64)
        buf = {"nbytes": nbytes, "loc": loc, "rows": [f64[i * per:(i + 1) * per] for i in range(4)],
               "info": loc[4 * per * 8: 4 * per * 8 + per * 4].view(torch.int32),
               "status": loc[4 * per * 8 + per * 4: 4 * per * 8 + per * 4 + 4].view(torch.int32),
               "out": torch.empty(world * nbytes, dtype=torch.uint8, device=dev),
               "host": torch.empty(world * nbytes, dtype=torch.uint8).pin_memory(),
               "stream": torch.cuda.ExternalStream(handle.stream, device=dev), "dev": dev}
        _NLL_BUFFERS[key] = buf
    return buf


def sharded_nll_device(handle, theta, p, lam, mode, group=None):
    """The population path on GPUs (NCCL): this rank's slice goes through ``pk_nll_batch`` with DEVICE outputs that
    ARE the send buffer, the per-candidate numbers (NegLnLike, mu, SigmaSqr, LnDetPsi, info) plus a status word per
    rank are gathered by ONE ``all_gather`` on the handle's stream, and one copy into pinned host memory returns the
    (C, 5) table -- no host round trip and no extra kernel between the likelihood epilogue and the collective.
    Same contract as ``sharded_rows``."""
    import torch

    from . import _lib

    dist = _dist()
    C = int(theta.shape[0])
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(C, rank, world)
    per = -(-C // world)
    buf = _nll_buffers(handle, per, world)
    err = None
    with torch.cuda.device(buf["dev"]), torch.cuda.stream(buf["stream"]):
        if hi > lo:
            try:
                r = buf["rows"]
                handle.set_shard_offset(lo)   # this slice starts at candidate lo of the hinted population
                handle.nll_batch_raw(hi - lo, np.ascontiguousarray(theta[lo:hi]), np.ascontiguousarray(p[lo:hi]),
                                     None if lam is None else np.ascontiguousarray(lam[lo:hi]), mode,
                                     r[0], r[1], r[2], r[3], buf["info"])
            except _lib.PkError as e:
                err = e
                buf["status"].fill_(1)
        dist.all_gather_into_tensor(buf["out"], buf["loc"], group=group)
        buf["host"].copy_(buf["out"], non_blocking=True)
        if err is not None:
            buf["status"].fill_(0)
        buf["stream"].synchronize()
    raw = buf["host"].numpy().reshape(world, buf["nbytes"])
    status = raw[:, 4 * per * 8 + per * 4: 4 * per * 8 + per * 4 + 4].copy().view(np.int32)[:, 0]
    if err is not None:
        raise err
    if np.any(status != 0):
        raise ShardError("pk_nll_batch failed on rank(s) %s" % np.nonzero(status)[0].tolist())
    rows = np.empty((world * per, 5))
    rows[:, :4] = raw[:, : 4 * per * 8].copy().view(np.float64).reshape(world, 4, per).transpose(0, 2, 1).reshape(world * per, 4)
    rows[:, 4] = raw[:, 4 * per * 8: 4 * per * 8 + per * 4].copy().view(np.int32).reshape(world * per)
    rows = rows[:C]
    if np.any(rows[:, 4] < 0):
        # info < 0: a spin-wait of the persistent Cholesky timed out on some rank; every rank sees it in the gathered table
        raise ShardError("a Cholesky spin-wait timed out on the rank that owns candidate(s) %s"
                         % np.nonzero(rows[:, 4] < 0)[0][:8].tolist())
    return rows


def rank_and_world(group=None):
    dist = _dist()
    if dist is None:
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def uses_nccl(group=None):
    dist = _dist()
    return dist is not None and dist.get_world_size(group) > 1 and str(dist.get_backend(group)).lower() == "nccl"


def global_argmax(local_values, lo, group=None, failed=False):
    """Arg-max over a sharded 1-D array: ``local_values`` are this rank's entries, starting at global
    index ``lo``.  Returns (value, global_index), identical on every rank; ties resolve to the lowest
    global index, as ``max()`` over the reference's in-order Python list would.  ONE all_gather of a
    (value, index, status) triple per rank; ``failed=True`` on any rank makes every rank raise ShardError."""
    import torch

    local_values = np.asarray(local_values, dtype=np.float64)
    if local_values.size and not failed:
        i = int(np.argmax(local_values))
        pair = np.array([[local_values[i], float(lo + i), 0.0]])
    else:
        pair = np.array([[-np.inf, np.inf, 1.0 if failed else 0.0]])
    dist = _dist()
    if dist is None or dist.get_world_size(group) == 1:
        if failed:
            raise ShardError("evaluation of the slice failed")
        return float(pair[0, 0]), int(pair[0, 1])
    world = dist.get_world_size(group)
    dev = _comm_device(dist, group)
    out = torch.empty(world, 3, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, torch.from_numpy(pair).to(dev), group=group)
    pairs = out.cpu().numpy()
    if np.any(pairs[:, 2] != 0.0):
        raise ShardError("evaluation of the slice failed on rank(s) %s" % np.nonzero(pairs[:, 2])[0].tolist())
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
