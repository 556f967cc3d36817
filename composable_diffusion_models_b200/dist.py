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


def gather_samples(local, total=None, dst=None):
    """Concatenate the per-rank shards along dim 0 in rank order.  dst=None: every rank gets the result
    (all_gather); dst=k: only rank k does (others get None).  Shards may be ragged (``shard_bounds``)."""
    rank, w = world()
    if w == 1:
        return local
    total = total if total is not None else None
    sizes = [None] * w
    dist.all_gather_object(sizes, int(local.shape[0]))
    mx = max(sizes)
    pad = local
    if local.shape[0] < mx:   # all_gather needs equal shapes
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))])
    pad = pad.contiguous()
    if dst is None:
        bufs = [torch.empty_like(pad) for _ in range(w)]
        dist.all_gather(bufs, pad)
    else:
        bufs = [torch.empty_like(pad) for _ in range(w)] if rank == dst else None
        dist.gather(pad, bufs, dst=dst)
        if rank != dst:
            return None
    out = torch.cat([b[:n] for b, n in zip(bufs, sizes)])
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
