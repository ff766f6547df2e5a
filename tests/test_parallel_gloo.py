"""CPU, world_size 2, gloo: the multi-process host logic (pair sharding, bucketed gradient all-reduce, row-tile halos)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dualpixelface_b200 import parallel


def test_shard_pairs_and_row_tiles():
    assert [len(parallel.shard_pairs(10, r, 4)) for r in range(4)] == [3, 3, 2, 2]
    assert sorted(i for r in range(4) for i in parallel.shard_pairs(10, r, 4)) == list(range(10))
    tiles = parallel.row_tiles(2240, 8)
    assert [(e - s) // 16 for s, e in tiles] == [18, 18, 18, 18, 17, 17, 17, 17]
    assert tiles[0][0] == 0 and tiles[-1][1] == 2240 and all(a[1] == b[0] for a, b in zip(tiles, tiles[1:]))
    with pytest.raises(ValueError):
        parallel.row_tiles(2250, 8)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Conv3d(4, 8, 3), torch.nn.BatchNorm3d(8), torch.nn.Conv3d(8, 1, 1))
        sync = parallel.make_grad_sync(model, bucket_bytes=100)
        x = torch.randn(2, 4, 5, 6, 7, generator=torch.Generator().manual_seed(10 + rank))
        model(x).square().mean().backward()
        local = [p.grad.clone() for p in model.parameters()]
        sync(model)
        gathered = [[torch.zeros_like(g) for _ in range(world)] for g in local]
        for g, lst in zip(local, gathered):
            dist.all_gather(lst, g)
        ok_grad = all(torch.allclose(p.grad, sum(lst) / world, atol=1e-6) for p, lst in zip(model.parameters(), gathered))
        # the collectives of the filled buckets were issued from the hooks, i.e. during backward (not from sync())
        ok_hooks = sync.launched_in_backward == len(sync.buckets) >= 2
        # second step: state was reset; a parameter without a gradient contributes zeros; set_to_none grads are re-created
        for p in model.parameters():
            p.grad = None
        head = torch.nn.Conv3d(8, 1, 1)
        head.weight.data.copy_(model[2].weight.data); head.bias.data.copy_(model[2].bias.data)
        y = model[1](model[0](x))
        (y.detach() * 0 + head(y)).mean().backward()            # model[2] gets no gradient in this step
        local2 = [None if p.grad is None else p.grad.clone() for p in model.parameters()]
        sync(model)
        for p, g in zip(model.parameters(), local2):
            g = torch.zeros_like(p) if g is None else g
            lst = [torch.zeros_like(g) for _ in range(world)]
            dist.all_gather(lst, g)
            ok_grad = ok_grad and torch.allclose(p.grad, sum(lst) / world, atol=1e-6)
        # row-tile halo exchange: global image rows 0..7 split in two, halo 1
        full = torch.arange(8.0).view(1, 8, 1).repeat(1, 1, 3)
        mine = full[:, rank * 4:(rank + 1) * 4]
        got = parallel.exchange_row_halo(mine.contiguous(), 1, 1)
        want = torch.cat([full[:, rank * 4 - 1: rank * 4] if rank else torch.zeros(1, 1, 3), mine,
                          full[:, (rank + 1) * 4:(rank + 1) * 4 + 1] if rank + 1 < world else torch.zeros(1, 1, 3)], 1)
        got_w = parallel.exchange_row_halo(mine.contiguous(), 1, 1, wrap=True)
        want_w = torch.cat([full[:, (rank * 4 - 1) % 8].unsqueeze(1), mine, full[:, ((rank + 1) * 4) % 8].unsqueeze(1)], 1)
        ret[rank] = (bool(ok_grad), bool(ok_hooks), bool(torch.equal(got, want) and torch.equal(got_w, want_w)))
    finally:
        dist.destroy_process_group()


def test_grad_sync_and_halo_exchange_world2():
    world, port = 2, 29500 + os.getpid() % 2000
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: (True, True, True), 1: (True, True, True)}      # (gradients, overlap, halos)
