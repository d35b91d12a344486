"""Row-block sharding of the DGG path across GPUs.

Rows of the score / top-K / output CSR / node features are split into contiguous blocks, one per rank;
every output row needs its own embedding plus ALL column embeddings, so there is exactly one exchange
per stage: ``all_gather`` of the [N, d] embeddings before scoring (and of [N, F] features before an
aggregation layer), and in the backward the matching ``reduce_scatter`` of the column-side gradients.
Weight gradients are all-reduced by the caller.  Collectives go through ``torch.distributed`` (NCCL over
NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def row_block(n: int, world: int, rank: int):
    """-> (row_begin, row_count, rows_per_rank) of the contiguous block owned by ``rank``."""
    per = (n + world - 1) // world
    b = min(rank * per, n)
    e = min(b + per, n)
    return b, e - b, per


class _AllGatherRows(torch.autograd.Function):
    """x_local [cnt, F] on every rank -> x_all [n, F] (rank order); backward = reduce-scatter(sum)."""

    @staticmethod
    def forward(ctx, x_local, n, group):
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        _, cnt, per = row_block(n, world, rank)
        assert x_local.shape[0] == cnt, (x_local.shape, cnt)
        f = x_local.shape[1:]
        send = x_local.contiguous()
        if cnt < per:
            send = torch.cat([send, send.new_zeros((per - cnt,) + tuple(f))])
        out = send.new_empty((per * world,) + tuple(f))
        dist.all_gather_into_tensor(out, send, group=group)
        ctx.meta = (n, group, cnt, per, world)
        return out[:n]

    @staticmethod
    def backward(ctx, g_all):
        n, group, cnt, per, world = ctx.meta
        f = g_all.shape[1:]
        g = g_all.contiguous()
        if per * world > n:
            g = torch.cat([g, g.new_zeros((per * world - n,) + tuple(f))])
        if dist.get_backend(group) == "gloo":          # gloo has no reduce_scatter: all-reduce + slice
            dist.all_reduce(g, group=group)
            r = dist.get_rank(group)
            out = g[r * per:(r + 1) * per]
        else:
            out = g.new_empty((per,) + tuple(f))
            dist.reduce_scatter_tensor(out, g, group=group)
        return out[:cnt].contiguous(), None, None


def all_gather_rows(x_local, n, group=None):
    if group is None:
        group = dist.group.WORLD
    return _AllGatherRows.apply(x_local, n, group)


def all_reduce_grads(params, group=None):
    """Sum the gradients of replicated parameters across ranks with ONE flat all-reduce."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def flatten_grads(params):
    """Back every ``p.grad`` by a view into ONE flat buffer, so a single all-reduce of that buffer (no
    concatenation, capturable in a CUDA graph) sums all replicated-parameter gradients.  Returns the buffer."""
    params = [p for p in params if p.requires_grad]
    flat = torch.zeros(sum(p.numel() for p in params), dtype=params[0].dtype, device=params[0].device)
    off = 0
    for p in params:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    return flat


def sharded_allpairs_topk(z_local, t, n, kc, group=None, precision=3, seed=0, noise_scale=0.0, noise_local=None):
    """Row-sharded all-pairs top-K: all-gather the embeddings, score this rank's row block against all
    columns.  Returns (idx [cnt,kc] global column ids, y [cnt,kc]); gradients flow back to ``z_local``
    on every rank through the reduce-scatter of the column-side contributions."""
    from . import functional as K

    if group is None:
        group = dist.group.WORLD
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    rb, cnt, _ = row_block(n, world, rank)
    z_all = all_gather_rows(z_local, n, group)
    return K.allpairs_topk(z_all, t, noise_local, kc, precision, rb, cnt, seed, noise_scale)


def sharded_spmm(vals_local, x_local, graph_local, n, group=None, row_scale=None):
    """Y_local = A_local X for a row-sharded CSR (local rows, GLOBAL column ids) and row-sharded X."""
    from . import functional as K

    x_all = all_gather_rows(x_local, n, group)
    return K.spmm(vals_local, x_all, graph_local, row_scale)
