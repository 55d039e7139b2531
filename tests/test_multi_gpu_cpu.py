"""World-size-2 gloo test of the N>1 host logic (sharding, key broadcast, gather) on CPU tensors."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nufhe_b200.sharding import shard_bounds, broadcast_tensors, gather_shards


def test_shard_bounds_cover_batch_exactly():
    for batch in (0, 1, 7, 8, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            covered = []
            for r in range(world):
                s, e = shard_bounds(batch, world, r)
                assert 0 <= s <= e <= batch
                covered += list(range(s, e))
            assert covered == list(range(batch))
            sizes = [shard_bounds(batch, world, r)[1] - shard_bounds(batch, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, results):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # "cloud key": rank 0 owns the real values, the others start from garbage
        g = torch.Generator().manual_seed(5)
        bk = torch.randint(-2**62, 2**62, (6, 2, 2, 2, 1024), generator=g, dtype=torch.int64)
        ks_a = torch.randint(-2**31, 2**31, (16, 8, 4, 50), generator=g, dtype=torch.int64).to(torch.int32)
        want = [bk.clone(), ks_a.clone()]
        if rank != 0:
            bk.zero_()
            ks_a.fill_(-1)
        broadcast_tensors([bk, ks_a], src=0)
        ok = all(torch.equal(a, b) for a, b in zip([bk, ks_a], want))
        # shard a batch of "ciphertexts", process locally (a stand-in op), gather, compare with unsharded
        batch = 11
        full = torch.arange(batch * 5, dtype=torch.int32).reshape(batch, 5)
        s, e = shard_bounds(batch, world, rank)
        local = full[s:e] * 3 + 1
        gathered = gather_shards(local, batch, world, rank)
        ok = ok and torch.equal(gathered, full * 3 + 1)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_broadcast_and_gather_world_size_2():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    with ctx.Manager() as manager:
        results = manager.dict()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, results)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=120)
        assert all(p.exitcode == 0 for p in procs)
        assert dict(results) == {0: True, 1: True}
