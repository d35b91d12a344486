"""CPU, world_size 2, gloo: the host-side sharding logic (row partition, differentiable all-gather with
reduce-scatter backward, flat gradient all-reduce) against single-process results."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, n, f, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dgg_b200 import sharding as S

        gen = torch.Generator().manual_seed(0)
        x = torch.randn(n, f, generator=gen)
        w = torch.randn(n, f, generator=gen)
        rb, cnt, per = S.row_block(n, world, rank)
        xl = x[rb:rb + cnt].clone().requires_grad_(True)
        xa = S.all_gather_rows(xl, n)
        assert torch.equal(xa, x)
        # a function of ALL rows evaluated on every rank with a rank-dependent weight: each rank's
        # gradient w.r.t. its own rows must be the SUM of every rank's contribution to those rows
        ((xa * w).sum() * (rank + 1)).backward()
        want = w[rb:rb + cnt] * sum(r + 1 for r in range(world))
        torch.testing.assert_close(xl.grad, want)
        p = torch.nn.Parameter(torch.ones(3))
        p.grad = torch.full((3,), float(rank + 1))
        q = torch.nn.Parameter(torch.ones(2, 2))
        q.grad = torch.full((2, 2), 10.0 * (rank + 1))
        S.all_reduce_grads([p, q])
        assert torch.equal(p.grad, torch.full((3,), 3.0)) and torch.equal(q.grad, torch.full((2, 2), 30.0))
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 7])
def test_all_gather_rows_world2(n):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 1000) + n
    mp.spawn(_worker, args=(world, port, n, 3, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))


def test_row_block_partition_covers_everything():
    from dgg_b200.sharding import row_block

    for n in (1, 7, 128, 19717, 232965):
        for world in (1, 2, 4, 8):
            covered, prev_end = 0, 0
            for r in range(world):
                b, c, per = row_block(n, world, r)
                assert b == min(prev_end, n) or c == 0
                assert c <= per
                covered += c
                prev_end = b + c
            assert covered == n
