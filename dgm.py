"""Drop-in replacement for the reference ``dgm.py`` (Differentiable Graph Generator), B200-native.

Same class names, constructor/forward signatures, parameter names (``state_dict`` compatible) and
``args`` attribute surface as the reference, but nothing here ever builds
a dense N x N matrix: adjacencies stay CSR-resident and every hot op is a hand-written sm_100a kernel
behind the C-ABI in include/dggb.h (bound in dgg_b200/functional.py).  CUDA tensors only.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

import dgg_b200
from dgg_b200 import CSRGraph
from dgg_b200 import functional as K


def sample_gumbel_from_uniform(shape, eps=1e-20):
    """Gumbel(0,1) from uniform draws; same recipe as reference dgm.py:6-11."""
    u = torch.rand(shape, device="cuda" if torch.cuda.is_available() else "cpu")
    return -torch.log(-torch.log(u + eps) + eps)


def gumbel_sample(logits, noise_sample):
    """Reference dgm.py:14-29: the self-loop mask is computed there but never applied, so the
    contract is plain ``logits + noise_sample``."""
    assert logits.shape == noise_sample.shape
    return logits + noise_sample


class _EdgeRankerBase(nn.Module):
    """Shared body of ``DGG`` (dgm.py:1730-1815) and ``DGG_Ablations`` (dgm.py:1876-1968)."""

    def __init__(self, in_dim=32, latent_dim=64, args=None):
        super().__init__()
        self.args = args
        self.node_encoder = nn.Sequential(nn.Linear(in_dim, latent_dim), nn.LeakyReLU())
        self.edge_encoder = nn.Sequential(
            nn.Linear(latent_dim + self.args.extra_edge_dim, latent_dim), nn.LeakyReLU())
        self.degree_decoder = nn.Sequential(nn.Linear(1, 1, bias=True), nn.LeakyReLU())
        self.var_grads = {"edge_p": [], "first_k": [], "out_adj": []}

    def _rank_edges(self, x, adj, noise=None, hard_k=-1):
        assert x.ndim == 2
        assert len(adj.shape) == 2
        graph, _ = CSRGraph.from_coo(adj)
        x_enc = self.node_encoder(x)                                   # dgm.py:1778
        lin = self.edge_encoder[0]
        # Linear is linear: We (x_u - x_v) + be == y_u - y_v + be with y = x_enc We^T, so the per-edge
        # E x h x h GEMM of dgm.py:1783-1784 becomes one N x h x h GEMM plus a gather.
        y = F.linear(x_enc, lin.weight)
        dd = self.degree_decoder[0]
        out_vals, k, R, rank = K.dgg_edge(y, lin.bias, dd.weight, dd.bias, graph, noise, hard_k)
        self.last_k = k
        return graph, out_vals, x_enc


class DGG(_EdgeRankerBase):
    """Differentiable graph generator, edge-restricted ranker (reference dgm.py:1730-1815).

    forward(x [N,F], adj sparse COO [N,N]) -> (sparse COO [N,N] with the support of ``adj``, x_enc [N,h]).
    ``noise`` is accepted and ignored exactly like the reference (dgm.py:1758)."""

    def forward(self, x, adj, noise=True, writer=None, epoch=None):
        graph, out_vals, x_enc = self._rank_edges(x, adj)
        return graph.to_coo(out_vals), x_enc


class DGG_Ablations(_EdgeRankerBase):
    """Reference dgm.py:1876-1968: uniform(-1,1) score noise + optional hard integer k."""

    def forward(self, x, adj, k=None, writer=None, epoch=None):
        graph, _ = CSRGraph.from_coo(adj)
        noise = torch.rand(graph.nnz, device=x.device) * 2 - 1          # dgm.py:1933
        if k is None:
            graph, out_vals, x_enc = self._rank_edges(x, adj, noise=noise)
            return graph.to_coo(out_vals), x_enc
        graph, out_vals, x_enc = self._rank_edges(x, adj, noise=noise, hard_k=int(k))
        # srt[:, k:] = 0 then to_sparse(): entries ranked >= k are dropped from the support (1943-1945)
        keep = (out_vals != 0).nonzero().flatten()
        idx = graph.coo_indices()[:, keep]
        return torch.sparse_coo_tensor(idx, out_vals[keep], (graph.n, graph.n), is_coalesced=True), x_enc
