"""GPU: the device-side GraphSAINT random-walk sampler (csrc/sampler.cu, dgg_b200.samplers) against a Python
restatement of torch_sparse ``random_walk`` + ``saint_subgraph`` driven by the same Philox stream (tests/philox_ref.py).
The library the reference calls (torch_geometric.loader, train_large_graphs.py:402-413) is not in the reference tree
and not installed: walks / node sets / induced sub-graphs are bit-exact against THIS restatement only."""
import types

import numpy as np
import pytest
import torch

from tests.philox_ref import philox4x32_7

pytestmark = pytest.mark.gpu


def _graph(n, avg_deg, seed, isolated=3):
    gen = torch.Generator().manual_seed(seed)
    m = n * avg_deg
    src = torch.randint(0, n - isolated, (m,), generator=gen)       # the last rows have no out-edges
    dst = torch.randint(0, n, (m,), generator=gen)
    a = torch.sparse_coo_tensor(torch.stack([src, dst]), torch.ones(m), (n, n)).coalesce()
    return a.indices()


def _csr(idx, n):
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, idx[0].numpy() + 1, 1)
    return np.cumsum(rowptr), idx[1].numpy()


def walk_ref(rowptr, col, start, length, seed, offset=0):
    b = len(start)
    cur = start.astype(np.int64).copy()
    walk = np.zeros((b, length + 1), dtype=np.int64)
    walk[:, 0] = cur
    outs = None
    for s in range(length):
        if s % 4 == 0:
            outs = philox4x32_7(torch.arange(b, dtype=torch.int64) + offset, torch.full((b,), s >> 2, dtype=torch.int64),
                                seed)
        x = outs[s % 4].numpy()
        deg = rowptr[cur + 1] - rowptr[cur]
        j = ((x >> 8) * deg) >> 24
        nxt = col[np.minimum(rowptr[cur] + j, len(col) - 1)]
        cur = np.where(deg > 0, nxt, cur)
        walk[:, s + 1] = cur
    return walk


def induced_ref(rowptr, col, nodes):
    pos = -np.ones(len(rowptr) - 1, dtype=np.int64)
    pos[nodes] = np.arange(len(nodes))
    sub_rowptr, sub_col, eid = [0], [], []
    for u in nodes:
        for e in range(rowptr[u], rowptr[u + 1]):
            if pos[col[e]] >= 0:
                sub_col.append(pos[col[e]])
                eid.append(e)
        sub_rowptr.append(len(sub_col))
    return np.array(sub_rowptr), np.array(sub_col, dtype=np.int64), np.array(eid, dtype=np.int64)


@pytest.mark.parametrize("n,avg_deg,walkers,length", [(200, 4, 50, 4), (3000, 8, 700, 9), (64, 1, 64, 2)])
def test_random_walk_and_induced_subgraph_match_restatement(n, avg_deg, walkers, length):
    from dgg_b200 import CSRGraph, samplers

    idx = _graph(n, avg_deg, n + length)
    rowptr, col = _csr(idx, n)
    g = CSRGraph.from_indices(idx.cuda(), n)
    start = torch.randint(0, n, (walkers,), generator=torch.Generator().manual_seed(5))
    seed = 0x1234_5678_9ABC_DEF1
    walk = samplers.random_walk(g, start.cuda().to(torch.int32), length, seed, walker_offset=11)
    want = walk_ref(rowptr, col, start.numpy(), length, seed, offset=11)
    assert np.array_equal(walk.cpu().numpy().astype(np.int64), want)
    # every step is an edge of the graph (or a stay on a node without out-edges)
    a = set(zip(idx[0].tolist(), idx[1].tolist()))
    for w in want[:20]:
        for s in range(length):
            assert (w[s], w[s + 1]) in a or (rowptr[w[s] + 1] == rowptr[w[s]] and w[s] == w[s + 1])

    nodes = np.unique(want.reshape(-1))
    sub, eid = samplers.induced_subgraph(g, torch.from_numpy(nodes).cuda().to(torch.int32))
    r_rowptr, r_col, r_eid = induced_ref(rowptr, col, nodes)
    assert np.array_equal(sub.rowptr.cpu().numpy(), r_rowptr)
    assert np.array_equal(sub.col.cpu().numpy(), r_col)
    assert np.array_equal(eid.cpu().numpy(), r_eid)


def test_graphsaint_sampler_batches_are_induced_subgraphs_with_sliced_attributes():
    from dgg_b200 import samplers

    n, f = 2500, 12
    idx = _graph(n, 6, 3)
    perm = torch.randperm(idx.shape[1], generator=torch.Generator().manual_seed(1))   # un-sorted edge list, as loaded
    ei = idx[:, perm]
    gen = torch.Generator().manual_seed(2)
    data = types.SimpleNamespace(
        x=torch.randn(n, f, generator=gen), y=torch.randint(0, 5, (n,), generator=gen), edge_index=ei, num_nodes=n,
        train_mask=torch.rand(n, generator=gen) < 0.5, edge_weight=torch.rand(ei.shape[1], generator=gen))
    torch.manual_seed(0)
    loader = samplers.GraphSAINTRandomWalkSampler(data, batch_size=120, walk_length=3, num_steps=4, sample_coverage=5,
                                                  save_dir=None, num_workers=4)
    assert len(loader) == 4
    dense = torch.zeros(n, n)
    dense[ei[0], ei[1]] = data.edge_weight
    batches = list(loader)
    assert len(batches) == 4
    for b in batches:
        b = b.to("cpu")
        m = b.num_nodes
        assert b.x.shape == (m, f) and b.y.shape == (m,) and b.train_mask.shape == (m,)
        # recover the node ids from the features (rows of x are distinct)
        ids = torch.cdist(b.x, data.x).argmin(1)
        assert torch.equal(b.x, data.x[ids]) and torch.equal(b.y, data.y[ids]) and torch.equal(b.train_mask, data.train_mask[ids])
        assert bool((ids[1:] > ids[:-1]).all())                          # sorted unique node list
        sub = torch.zeros(m, m)
        sub[b.edge_index[0], b.edge_index[1]] = b.edge_weight
        assert torch.equal(sub, dense[ids][:, ids])                       # induced: every parent edge between them
        assert b.node_norm.shape == (m,) and b.edge_norm.shape == (b.num_edges,)
        assert bool(torch.isfinite(b.node_norm).all()) and bool(torch.isfinite(b.edge_norm).all())
        assert 100 < m <= 120 * 4
    # the scripts' seed determines the batches
    torch.manual_seed(0)
    again = list(samplers.GraphSAINTRandomWalkSampler(data, batch_size=120, walk_length=3, num_steps=4,
                                                      sample_coverage=5))
    for a, b in zip(batches, again):
        assert torch.equal(a.edge_index.cpu(), b.edge_index.cpu()) and torch.equal(a.x.cpu(), b.x.cpu())
