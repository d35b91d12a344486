"""Device-side sub-graph sampler for the large-graph drivers.

``train_large_graphs.py:402-413`` / ``train_reddit.py:400-411`` build their mini-batches with
``torch_geometric.loader.GraphSAINTRandomWalkSampler(data, batch_size=, walk_length=, num_steps=, sample_coverage=,
save_dir=, num_workers=)`` -- third-party, CPU, worker processes; every batch is then moved to the device and converted
edge list -> scipy -> torch sparse (train_large_graphs.py:226-233).  ``GraphSAINTRandomWalkSampler`` here keeps the
constructor and the batch attribute surface (``x``, ``y``, ``edge_index``, the ``*_mask`` / edge attributes of the
parent, ``node_norm`` / ``edge_norm`` when ``sample_coverage > 0``, ``num_nodes``, ``.to(device)``) but samples on the
GPU from the int32 CSR the rest of the path uses: ``dggb_random_walk`` (one thread per walker, Philox uniforms) and
``dggb_induced_subgraph_count/fill``.  A maintainer swaps ``import torch_geometric.loader as dataloaders`` for
``import dgg_b200.samplers as dataloaders``.

Semantics restated from PyG 2.1 ``GraphSAINTSampler`` / torch_sparse (neither is in the reference tree: parity is
pinned only against the Python restatement in tests/test_gpu_sampler.py, i.e. **unpinned** against the library):
start nodes ``torch.randint(0, N, (batch_size,))`` from the global CPU generator (what ``torch.manual_seed(args.seed)``
of the scripts seeds), uniform neighbour choice per step, walkers on a node without out-edges stay, the batch is the
sub-graph induced by the sorted unique visited nodes with all parent edges between them in row-major order, and the
normalisation statistics are visit counts over pre-sampled batches (``edge_norm = node_count[row] / edge_count``
clamped to 1e4, ``node_norm = num_samples / node_count / N``).  The Cluster / Neighbor loaders the scripts can also
select are not provided.
"""
from __future__ import annotations

import torch

from ._lib import check, i32, i64, lib, p, stream
from .graph import CSRGraph


class SubgraphBatch:
    """Attribute bag standing in for ``torch_geometric.data.Data`` (``.to`` / ``.num_nodes`` / ``.keys()``)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def to(self, device, non_blocking=False):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device, non_blocking=non_blocking))
        return self

    @property
    def num_edges(self):
        return int(self.edge_index.shape[1])


def _data_items(data):
    keys = data.keys() if callable(getattr(data, "keys", None)) else getattr(data, "keys", None)
    if keys is None:
        keys = [k for k in vars(data) if not k.startswith("_")]
    return [(k, getattr(data, k)) for k in keys]


def induced_subgraph(graph: CSRGraph, nodes: torch.Tensor):
    """Sub-graph induced by the sorted unique int32 node list -> (CSRGraph over len(nodes) nodes, eid int32 [E'] =
    positions of the kept entries in ``graph``).  One device->host read (the entry count sizes the outputs)."""
    assert nodes.is_cuda and nodes.dtype == torch.int32 and nodes.dim() == 1
    m = int(nodes.numel())
    dev = nodes.device
    relabel = torch.empty(graph.n, dtype=torch.int32, device=dev)
    counts = torch.empty(m + 1, dtype=torch.int32, device=dev)
    L = lib()
    check(L.dggb_induced_subgraph_count(p(graph.rowptr), p(graph.col), i32(graph.n), p(nodes), i32(m), p(relabel),
                                        p(counts), stream()), "induced_subgraph_count")
    sub_rowptr = torch.zeros(m + 1, dtype=torch.int32, device=dev)
    torch.cumsum(counts[:m], 0, out=sub_rowptr[1:])
    nnz = int(sub_rowptr[-1].item())
    sub_col = torch.empty(nnz, dtype=torch.int32, device=dev)
    sub_eid = torch.empty(nnz, dtype=torch.int32, device=dev)
    check(L.dggb_induced_subgraph_fill(p(graph.rowptr), p(graph.col), p(nodes), i32(m), p(relabel), p(sub_rowptr),
                                       p(sub_col), p(sub_eid), stream()), "induced_subgraph_fill")
    return CSRGraph(m, sub_rowptr, sub_col), sub_eid


def random_walk(graph: CSRGraph, start: torch.Tensor, walk_length: int, seed: int, walker_offset: int = 0):
    """-> int32 [len(start), walk_length + 1] (column 0 = start)."""
    assert start.is_cuda and start.dtype == torch.int32
    b = int(start.numel())
    walk = torch.empty(b, walk_length + 1, dtype=torch.int32, device=start.device)
    check(lib().dggb_random_walk(p(graph.rowptr), p(graph.col), i32(graph.n), p(start), i32(b), i32(walk_length),
                                 int(seed) & 0xFFFFFFFFFFFFFFFF, i64(walker_offset), p(walk), stream()), "random_walk")
    return walk


class GraphSAINTRandomWalkSampler:
    """``for batch in loader`` yields ``num_steps`` sub-graphs per epoch (``__len__ == num_steps``)."""

    def __init__(self, data, batch_size, walk_length, num_steps=1, sample_coverage=0, save_dir=None, log=True,
                 device=None, **kwargs):
        assert data.edge_index is not None
        self.data = data
        self.walk_length = int(walk_length)
        self.num_steps = int(num_steps)
        self.batch_size = int(batch_size)
        self.sample_coverage = sample_coverage
        dev = torch.device(device) if device is not None else (
            data.edge_index.device if data.edge_index.is_cuda else torch.device("cuda"))
        self.device = dev
        self.N = int(getattr(data, "num_nodes", None) or data.x.shape[0])
        ei = data.edge_index.to(dev)
        self.E = int(ei.shape[1])
        # parent adjacency in CSR order; perm maps a CSR position to the caller's edge id (edge attributes follow it)
        key = ei[0] * self.N + ei[1]
        self.perm = torch.argsort(key, stable=True)
        self.graph = CSRGraph.from_indices(ei[:, self.perm].contiguous(), self.N)
        self._node_attrs, self._edge_attrs, self._other = {}, {}, {}
        for k, v in _data_items(data):
            if k in ("edge_index", "num_nodes"):
                continue
            if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == self.N:
                self._node_attrs[k] = v.to(dev)
            elif torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == self.E:
                self._edge_attrs[k] = v.to(dev)
            else:
                self._other[k] = v
        self.node_norm = self.edge_norm = None
        if self.sample_coverage > 0:
            self.node_norm, self.edge_norm = self._compute_norm()

    def __len__(self):
        return self.num_steps

    # -------------------------------------------------------------------------------------------- sampling
    def _sample(self):
        """-> (node_idx int32 sorted unique, sub CSRGraph, eid int32 positions in the parent CSR)."""
        start = torch.randint(0, self.N, (self.batch_size,), dtype=torch.long)       # global CPU generator, as PyG
        seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.long))
        walk = random_walk(self.graph, start.to(self.device, dtype=torch.int32), self.walk_length, seed)
        node_idx = torch.unique(walk.reshape(-1))                                    # sorted ascending
        sub, eid = induced_subgraph(self.graph, node_idx.to(torch.int32))
        return node_idx, sub, eid

    def _compute_norm(self):
        node_count = torch.zeros(self.N, dtype=torch.float32, device=self.device)
        edge_count = torch.zeros(self.E, dtype=torch.float32, device=self.device)
        num_samples = total = 0
        while total < self.N * self.sample_coverage:
            node_idx, _, eid = self._sample()
            node_count[node_idx.long()] += 1
            edge_count[eid.long()] += 1
            total += int(node_idx.numel())
            num_samples += 1
        row = self.graph.erow.long()
        edge_norm = (node_count[row] / edge_count).clamp_(0, 1e4)
        edge_norm[torch.isnan(edge_norm)] = 0.1
        node_count[node_count == 0] = 0.1
        node_norm = num_samples / node_count / self.N
        return node_norm, edge_norm        # edge_norm is indexed by CSR position

    def _collate(self, node_idx, sub, eid):
        nl, el = node_idx.long(), eid.long()
        batch = SubgraphBatch(num_nodes=int(node_idx.numel()), edge_index=sub.coo_indices())
        for k, v in self._node_attrs.items():
            setattr(batch, k, v[nl])
        orig = self.perm[el]
        for k, v in self._edge_attrs.items():
            setattr(batch, k, v[orig])
        for k, v in self._other.items():
            setattr(batch, k, v)
        if self.node_norm is not None:
            batch.node_norm = self.node_norm[nl]
            batch.edge_norm = self.edge_norm[el]
        batch._csr = sub           # consumers inside this package skip the edge list -> CSR conversion
        return batch

    def __iter__(self):
        for _ in range(self.num_steps):
            yield self._collate(*self._sample())
