"""Row-block sharding of the DGG path across GPUs.

Rows of the score / top-K / output CSR / node features are split into contiguous blocks, one per rank;
every output row needs its own embedding plus ALL column embeddings, so there is exactly one exchange
per stage: ``all_gather`` of the [N, d] embeddings before scoring (and of [N, F] features before an
aggregation layer), and in the backward the matching ``reduce_scatter`` of the column-side gradients.
Weight gradients are all-reduced by the caller.  Collectives go through ``torch.distributed`` (NCCL over
NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import os

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


def flatten_grads(params, flat=None):
    """Back every ``p.grad`` by a view into ONE flat buffer, so a single all-reduce of that buffer (no
    concatenation, capturable in a CUDA graph) sums all replicated-parameter gradients.  Returns the buffer
    (``flat``: use this preallocated buffer, e.g. ``PeerAllReduce.buffer``, instead of a new one)."""
    params = [p for p in params if p.requires_grad]
    total = sum(p.numel() for p in params)
    if flat is None:
        flat = torch.zeros(total, dtype=params[0].dtype, device=params[0].device)
    assert flat.numel() >= total
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


class PeerAllReduce:
    """Sum of a flat fp32 buffer over the ranks by ONE kernel over NVLink peer memory (``dggb_allreduce_oneshot``),
    launched on the caller's stream: it can be captured into the CUDA graph of the training step, unlike a
    library collective issued after the replay.

        ar = PeerAllReduce(numel)            # collective: every rank constructs it
        flatten_grads(params, ar.buffer)     # gradients accumulate straight into symmetric memory
        ... backward ...
        total = ar()                         # -> ar.out [numel]: the sum over ranks, identical bits on every rank

    The buffer lives in ``torch.distributed._symmetric_memory`` (peer-mapped; bound to an NVSwitch multicast
    object when the fabric supports it, in which case the switch performs the reduction: multimem.ld_reduce).
    ``end_barrier=False``: for callers that alternate between TWO instances from step to step (the closing "done
    reading" round trip is then implied by the next step's opening barrier; see csrc/peer.cu).
    ``PeerAllReduce.available()`` is False where symmetric memory cannot be set up (then use ``dist.all_reduce``)."""

    PAD_SLOT_BASE = 1024      # uint32 slots of the signal pad this kernel owns (torch's own barriers use low slots)

    def __init__(self, numel, group=None, use_multicast=True, blocks=None, end_barrier=True):
        import torch.distributed._symmetric_memory as symm_mem

        from ._lib import lib

        self.group = dist.group.WORLD if group is None else group
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.numel = int(numel)
        padded = (self.numel + 3) // 4 * 4
        dev = torch.device("cuda", torch.cuda.current_device())
        self._sym = symm_mem.empty(padded, dtype=torch.float32, device=dev)
        self._sym.zero_()
        self._hdl = symm_mem.rendezvous(self._sym, self.group.group_name)
        assert self._hdl.signal_pad_size >= 4 * (self.PAD_SLOT_BASE + 2 * self.world)
        self.buffer = self._sym[:self.numel]
        self._out = torch.zeros(padded, dtype=torch.float32, device=dev)
        self.out = self._out[:self.numel]
        self._state = torch.zeros(2, dtype=torch.int32, device=dev)
        if os.environ.get("DGGB_PEER_NO_MC"):
            use_multicast = False
        mc = int(getattr(self._hdl, "multicast_ptr", 0) or 0) if use_multicast else 0
        self.multicast = mc != 0
        self._mc = mc
        # one 16-byte element per thread where possible: every load of the sum is in flight at once
        self._blocks = int(blocks) if blocks else max(1, min(148, (padded // 4 + 255) // 256))
        self._lib = lib()
        self._end_barrier = 1 if end_barrier else 0
        torch.cuda.synchronize()
        dist.barrier(self.group)            # every rank's pad and state are zeroed before the first call

    @staticmethod
    def available():
        try:
            import torch.distributed._symmetric_memory as symm_mem  # noqa: F401

            return torch.cuda.is_available() and dist.is_initialized() and dist.get_backend() == "nccl"
        except Exception:
            return False

    def __call__(self):
        from ._lib import check, stream

        check(self._lib.dggb_allreduce_oneshot(int(self._hdl.buffer_ptrs_dev), int(self._hdl.signal_pad_ptrs_dev),
                                               self.rank, self.world, self._sym.numel(), self._out.data_ptr(),
                                               self._state.data_ptr(), self._mc or None, self.PAD_SLOT_BASE,
                                               self._blocks, self._end_barrier, stream()), "allreduce_oneshot")
        return self.out


class RowShardedSAGE_DGG(torch.nn.Module):
    """``SAGE_DGG``-style model (reference model.py:122-193) for graphs too large for one GPU, with the all-pairs DGG
    of the north star instead of an input edge list (BASELINE.json configs[3]; the reference's own answer,
    train_reddit.py:340-369, feeds the 233 k-node graph to a dense N x N pipeline and cannot run).

    Rank r owns the contiguous node block [r ceil(N/R), ...): rows of the scores, of the learned adjacency, of X.
      z = softmax(LeakyReLU(x Wz + bz))                      local rows        (dgm.py:217-221)
      all_gather(z) -> y_ij = -t |z_i - z_j| + G_ij, top-Kc of every local row (tcgen05 score GEMM + streaming
                        top-K, Philox Gumbel noise keyed on (row, col): no exchange)            (dgm.py:275-301)
      a_ij = exp(y_ij) * (1 - 0.5 (1 + tanh(rank_ij - k_i)))   (the live pipeline of SURVEY A.2 on P = exp(-t D))
      s_i = sum_j a_ij;  all_gather(s^-1/2)  ->  ahat_ij = a_ij s_i^-1/2 s_j^-1/2            (model.py:146-149)
      per conv layer (PyG DenseGraphConv, aggr = mean; model.py:128-129):
          p = h W_rel^T (local, narrow);  all_gather(p);  agg = (ahat p_all) / clamp(rowsum(ahat), 1)
          h' = agg + b_rel + h W_root^T
    Backward: the autograd of ``all_gather_rows`` is the matching reduce-scatter (column-side gradients of p, z and
    s^-1/2 travel back to the owners); replicated-parameter gradients are summed by the caller."""

    def __init__(self, nfeat, nhidden, nclass, d=64, kc=32, k_init=9.0, t_init=4.0):
        super().__init__()
        nn = torch.nn
        self.input_project = nn.Sequential(nn.Linear(nfeat, d), nn.LeakyReLU(), nn.Softmax(dim=-1))
        self.t = nn.Parameter(torch.full((1,), float(t_init)))
        self.k_net = nn.Linear(nfeat, 1)
        self.lin_rel1, self.lin_root1 = nn.Linear(nfeat, nhidden), nn.Linear(nfeat, nhidden, bias=False)
        self.lin_rel2, self.lin_root2 = nn.Linear(nhidden, nclass), nn.Linear(nhidden, nclass, bias=False)
        with torch.no_grad():
            self.k_net.weight.mul_(0.1)
            self.k_net.bias.fill_(k_init)
        self.kc = int(kc)

    def adjacency(self, x_local, n, group, seed):
        """-> (idx [cnt, kc] global columns, ahat [cnt, kc], row scale [cnt] = 1 / clamp(rowsum(ahat), 1))"""
        from . import functional as K

        z = self.input_project(x_local)
        k = torch.relu(self.k_net(x_local)) + 1.0                                  # [cnt, 1], >= 1 (dgm.py:1580-1584)
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            idx, y = sharded_allpairs_topk(z, self.t, n, self.kc, group, seed=seed, noise_scale=1.0)
        else:
            idx, y = K.allpairs_topk(z, self.t, None, self.kc, 3, seed=seed, noise_scale=1.0)
        r = torch.arange(self.kc, device=x_local.device, dtype=torch.float32).reshape(1, -1)
        a = torch.exp(y) * (1 - 0.5 * (1 + torch.tanh(r - k)))                     # dgm.py:1213-1229, 1410-1417
        dinv = a.sum(-1).clamp_min(1e-30) ** -0.5                                  # [cnt]
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dinv_all = all_gather_rows(dinv.unsqueeze(-1), n, group).squeeze(-1)
        else:
            dinv_all = dinv
        ahat = a * dinv.unsqueeze(-1) * dinv_all[idx.long()]
        scale = 1.0 / ahat.sum(-1).clamp(min=1)
        return idx, ahat, scale

    def forward(self, x_local, n, group=None, seed=0):
        from . import functional as K
        from .graph import CSRGraph

        group = dist.group.WORLD if (group is None and dist.is_initialized()) else group
        sharded = dist.is_initialized() and dist.get_world_size(group) > 1
        cnt = x_local.shape[0]
        idx, ahat, scale = self.adjacency(x_local, n, group, seed)
        rowptr = torch.arange(0, (cnt + 1) * self.kc, self.kc, dtype=torch.int32, device=x_local.device)
        g = CSRGraph(cnt, rowptr, idx.reshape(-1).contiguous())          # local rows, GLOBAL column ids
        g._max_row_nnz = self.kc
        vals = ahat.reshape(-1)
        h = x_local
        for lin_rel, lin_root, last in ((self.lin_rel1, self.lin_root1, False), (self.lin_rel2, self.lin_root2, True)):
            p = torch.nn.functional.linear(h, lin_rel.weight)            # aggregate in the narrow space
            fo = p.shape[1]
            if fo % 4:                                                   # 128-bit gathers: pad 41 classes to 44
                p = torch.nn.functional.pad(p, (0, 4 - fo % 4))
            # (scale carries gradient: applied outside, not as the kernel's constant row_scale)
            agg = (sharded_spmm(vals, p, g, n, group) if sharded else K.spmm(vals, p, g))[:, :fo] * scale.unsqueeze(-1)
            h_new = agg + lin_rel.bias + lin_root(h)
            h = h_new if last else torch.nn.functional.dropout(torch.relu(h_new), 0.5, self.training)
        return torch.log_softmax(h, dim=-1), idx, ahat
