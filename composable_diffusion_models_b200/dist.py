"""Data-parallel plumbing for the sampler: every sample's chain is independent, so the batch
is split contiguously over ranks with NO collective inside the step loop.  The only collectives on the path are

  * ``gather_samples``            -- all_gather (or gather to rank 0) of the final x_0 / log-q shards, and
  * ``normalize_log_weights``     -- batch-level log-weight normalisation: log w_i - logsumexp_j(log w_j) over the
                                     WHOLE batch, as all_reduce(MAX) + all_reduce(SUM) of one scalar per expert.

One process per GPU (``torchrun``); backend NCCL on GPUs, gloo in the CPU tests.  Works unchanged when
``torch.distributed`` is not initialised (world size 1).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(total, rank=None, world_size=None):
    """[lo, hi) of this rank's contiguous slice; the first ``total % world`` ranks get one extra sample."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    base, extra = divmod(total, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(tensor, dim=0):
    """This rank's slice of a batch-major tensor that every rank holds in full (e.g. injected x_T / noise)."""
    lo, hi = shard_bounds(tensor.shape[dim])
    return tensor.narrow(dim, lo, hi - lo)


def gather_samples(local, total=None, dst=None, out=None):
    """Concatenate the per-rank shards along dim 0 in rank order.  dst=None: every rank gets the result
    (all_gather); dst=k: only rank k does (others get None).  Shards may be ragged (``shard_bounds``).
    ``out`` (optional, equal shards only): a preallocated [total, ...] tensor the shards land in -- a caller that gathers
    repeatedly (or wants the collective without a fresh 100 MB allocation in front of it) passes the same buffer every time.

    With ``total`` given the shard sizes follow from ``shard_bounds`` (no size exchange, no pickling): the shards land
    directly in ONE preallocated output -- ``all_gather_into_tensor`` (dst=None) or ``gather`` into row views of it -- so
    the equal-shard case is a single collective with no copy before or after it."""
    rank, w = world()
    if w == 1:
        return local
    local = local.contiguous()
    if total is not None:
        sizes = [hi - lo for lo, hi in (shard_bounds(total, r, w) for r in range(w))]
        if sizes[rank] != local.shape[0]:
            raise RuntimeError(f"rank {rank} holds {local.shape[0]} samples, shard_bounds({total}) expects {sizes[rank]}")
    else:
        mine = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        allsz = torch.empty(w, dtype=torch.int64, device=local.device)
        dist.all_gather_into_tensor(allsz, mine)
        sizes = [int(v) for v in allsz.tolist()]
    mx, tail = max(sizes), tuple(local.shape[1:])
    ragged = min(sizes) != mx
    send = local
    if local.shape[0] < mx:   # the collectives need equal shapes
        send = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tail)])
    if out is not None and (ragged or tuple(out.shape) != (w * mx,) + tail or not out.is_contiguous()):
        raise RuntimeError("gather_samples: `out` must be a contiguous [total, ...] tensor and the shards equal")
    if dst is None:
        out = local.new_empty((w * mx,) + tail) if out is None else out
        dist.all_gather_into_tensor(out, send)
    else:
        if rank == dst and out is None:
            out = local.new_empty((w * mx,) + tail)
        dist.gather(send, list(out.view((w, mx) + tail).unbind(0)) if rank == dst else None, dst=dst)
        if rank != dst:
            return None
    if ragged:
        out = torch.cat([out[r * mx:r * mx + n] for r, n in enumerate(sizes)])
    if total is not None and out.shape[0] != total:
        raise RuntimeError(f"gathered {out.shape[0]} samples, expected {total}")
    return out


def normalize_log_weights(log_w):
    """log_w [B_local, K] -> log_w - logsumexp over the global batch (per expert).  Two tiny all-reduces."""
    rank, w = world()
    m = log_w.max(dim=0).values if log_w.shape[0] else torch.full(log_w.shape[1:], -float("inf"), device=log_w.device)
    if w > 1:
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
    s = torch.exp(log_w - m).sum(dim=0)
    if w > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return log_w - (m + torch.log(s))
