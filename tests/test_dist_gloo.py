"""CPU, world_size 2, gloo: the N>1 path's host logic (batch sharding + output gather)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fusionocc_b200.dist import gather_outputs, shard_batch, shard_bounds


def test_shard_bounds_cover_the_batch():
    for batch in (0, 1, 7, 8, 64):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        full = torch.arange(batch * 2 * 3 * 4 * 5, dtype=torch.float32).view(batch, 2, 3, 4, 5)   # (B,C,Z,Y,X)
        other = torch.arange(batch * 6, dtype=torch.float32).view(batch, 6)
        mine, o, none = shard_batch([full, other, None], world, rank)
        lo, hi = shard_bounds(batch, world, rank)
        assert none is None and mine.shape[0] == hi - lo and torch.equal(o, other[lo:hi])
        local = mine * 2.0                                    # stand-in for this rank's view transform
        everyone = gather_outputs(local, batch)
        assert torch.equal(everyone, full * 2.0)
        only0 = gather_outputs(local, batch, dst=0)
        if rank == 0:
            assert torch.equal(only0, full * 2.0)
        else:
            assert only0 is None
        q.put((rank, 'ok'))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f'FAIL {type(e).__name__}: {e}'))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('batch', [8, 5])
def test_two_rank_shard_and_gather(batch):
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, 'ok'), (1, 'ok')], res
