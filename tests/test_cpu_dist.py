"""CPU: the multi-GPU host logic (sharding, final gather, batch-level log-weight normalisation) with world_size 2
over gloo.  The data path itself has no collective (each rank's chains are independent)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from composable_diffusion_models_b200 import dist as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        full = torch.randn(total, 3, 4, 4, generator=g)          # every rank can regenerate the whole batch
        logq = torch.randn(total, 2, generator=g) * 5
        lo, hi = D.shard_bounds(total)
        mine = D.shard(full)
        assert mine.shape[0] == hi - lo and torch.equal(mine, full[lo:hi])
        local = mine * 2 + 1                                      # stands in for this rank's independent chains
        allg = D.gather_samples(local, total=total)
        assert torch.equal(allg, full * 2 + 1)
        only0 = D.gather_samples(local, total=total, dst=0)
        assert (only0 is None) == (rank != 0)
        if rank == 0:
            assert torch.equal(only0, full * 2 + 1)
        norm = D.normalize_log_weights(D.shard(logq))
        want = (logq - torch.logsumexp(logq, dim=0))[lo:hi]
        assert torch.allclose(norm, want, atol=1e-5)
        q.put((rank, "ok"))
    except Exception as e:   # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def _run(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_two_rank_shard_gather_normalise_even():
    _run(8)


def test_two_rank_shard_gather_normalise_ragged():
    _run(7)


def test_single_process_is_a_noop():
    x = torch.randn(5, 2)
    assert D.shard_bounds(5) == (0, 5)
    assert torch.equal(D.gather_samples(x), x)
    assert torch.allclose(D.normalize_log_weights(x), x - torch.logsumexp(x, dim=0))
