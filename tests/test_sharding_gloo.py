"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU batch split (shard bounds, cost
balancing, variable-length result gather).  The data path itself has no collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from megatts2_b200 import sharding


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 64, 512, 513):
        for w in (1, 2, 4, 8):
            b = sharding.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.shard_bounds(512, 8) == [(64 * r, 64 * (r + 1)) for r in range(8)]   # C5: 64 per GPU


def test_balance_by_cost():
    lens = [512, 64, 64, 64, 300, 300, 128, 128]
    bins = sharding.balance_by_cost([l * l for l in lens], 2)
    assert sorted(i for b in bins for i in b) == list(range(8))
    loads = [sum(lens[i] ** 2 for i in b) for b in bins]
    assert max(loads) / min(loads) < 1.35


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.my_shard(n_items, rank, world)
        # stand-in "synthesis": utterance i yields i+3 samples, all equal to i (no CUDA in this test)
        lens = torch.tensor([i + 3 for i in range(lo, hi)])
        rows = torch.zeros(hi - lo, int(lens.max()) if hi > lo else 0)
        for j, i in enumerate(range(lo, hi)):
            rows[j, : i + 3] = float(i)
        gathered = sharding.gather_variable(rows, lens)
        seen = []
        for r, (t, l) in enumerate(gathered):
            rlo, rhi = sharding.my_shard(n_items, r, world)
            assert t.shape[0] == rhi - rlo
            for j, i in enumerate(range(rlo, rhi)):
                assert int(l[j]) == i + 3
                assert torch.all(t[j, : i + 3] == float(i)) and torch.all(t[j, i + 3:] == 0)
                seen.append(i)
        assert seen == list(range(n_items))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [5, 8])
def test_gather_world2_gloo(n_items):
    mp.spawn(_worker, args=(2, _free_port(), n_items), nprocs=2, join=True)


def _bcast_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from megatts2_b200 import vq_train as VT
        # the codebook buffers after a k-means initialisation differ per worker until rank 0's are broadcast
        # (EuclideanCodebook.init_embed_ -> distrib.broadcast_tensors, core_vq.py:141-149)
        bufs = [torch.full((1,), float(rank)), torch.full((4,), 10.0 + rank), torch.full((4, 3), 20.0 + rank),
                torch.tensor([rank], dtype=torch.int64)]
        VT.broadcast_buffers(bufs)
        assert torch.all(bufs[0] == 0) and torch.all(bufs[1] == 10) and torch.all(bufs[2] == 20)
        assert int(bufs[3]) == rank            # integer buffers are left alone, like the reference
        if rank == 1:
            with pytest.raises(RuntimeError):
                VT.broadcast_buffers(bufs[:2])
        else:
            with pytest.raises(RuntimeError):
                VT.broadcast_buffers(bufs[:3])
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_codebook_buffer_broadcast_world2_gloo():
    mp.spawn(_bcast_worker, args=(2, _free_port()), nprocs=2, join=True)
