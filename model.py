"""Drop-in replacement for the DGG model families of the reference ``model.py``, B200-native.

Class names, constructor arguments, forward signatures/return arities and parameter names follow the
reference; aggregation runs on CSR through dgg_b200 kernels instead of dense
N x N ``torch.mm``.  No torch_geometric dependency.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from dgg_b200 import CSRGraph
from dgg_b200 import functional as K
from dgm import DGG, DGG_Ablations


# --------------------------------------------------------------------------- helpers
def add_self_loops_coo(in_adj):
    """(in_adj.to_dense() + eye).to_sparse().coalesce() without the dense round trip
    (reference model.py:1381-1392 and twins)."""
    g0, v0 = CSRGraph.from_coo(in_adj)
    g, v = g0.with_self_loops(v0)
    return g.to_coo(v)


def normalize_adj_sparse(adj):
    """D^-1/2 A D^-1/2 with row sums on both sides (reference model.py:1215-1218), O(nnz)."""
    g, v = CSRGraph.from_coo(adj)
    return g.to_coo(K.sym_normalize(v, g))


class _NormalizeMixin:
    def normalize_adj(self, A_hat):
        if A_hat.is_sparse:
            return normalize_adj_sparse(A_hat)
        row_sum = A_hat.sum(-1) ** -0.5          # dense input kept for API compatibility
        return row_sum.unsqueeze(-1) * A_hat * row_sum.unsqueeze(0)

    def normalize_adj_gcn(self, A_hat):
        return self.normalize_adj(A_hat)


def _aggregate(adj, x, row_scale=None):
    """adj @ x for a sparse (CSR-backed) or dense adjacency."""
    if adj.is_sparse:
        g, v = CSRGraph.from_coo(adj)
        return K.spmm(v, x, g, row_scale)
    out = torch.mm(adj, x)
    return out if row_scale is None else out * row_scale.unsqueeze(-1)


# --------------------------------------------------------------------------- conv layers
class GCNConv(nn.Module):
    """relu((A x) W), W ~ U[0,1) (reference model.py:580-599)."""

    def __init__(self, in_channels, out_channels, A=None, cached=False):
        super().__init__()
        self.W = nn.Parameter(torch.rand(in_channels, out_channels, requires_grad=True))

    def forward(self, x, adj):
        return torch.relu(torch.mm(_aggregate(adj, x), self.W))


# --------------------------------------------------------------------------- GCN + DGG
class GCN_DGG_00(torch.nn.Module, _NormalizeMixin):
    """Reference model.py:1314-1433."""

    _dgg_cls = DGG

    def __init__(self, nfeat=32, nlayers=None, nhidden=32, nclass=10, args=None, **kwargs):
        super().__init__()
        self.convs = nn.ModuleList()
        self.conv1 = GCNConv(nhidden, nhidden)
        self.conv2 = GCNConv(nhidden, nclass)
        self.convs.append(self.conv1)
        self.convs.append(self.conv2)
        self.dgg_adj_input = args.dgg_adj_input
        self.dggs = nn.ModuleList()
        self.dggs.append(self._dgg_cls(in_dim=nfeat, latent_dim=nhidden, args=args))
        self.params1 = list(self.conv1.parameters())
        self.params2 = list(self.conv2.parameters())
        self.params2.extend(list(self.dggs.parameters()))

    def forward(self, x, in_adj, noise=True, epoch=None, writer=None, **kwargs):
        in_adj = add_self_loops_coo(in_adj)
        unnorm_adj = in_adj
        for i, conv in enumerate(self.convs):
            if i < len(self.dggs):
                src = in_adj if self.dgg_adj_input == "input_adj" else unnorm_adj
                unnorm_adj, x_dgg = self.dgg_net(x, i, src, writer, epoch)
                norm_adj = self.normalize_adj(unnorm_adj)
                x = x_dgg
            x = conv(x + x_dgg, norm_adj)
            if i < len(self.convs) - 1:
                x = F.dropout(x, training=self.training)
            if writer is not None:
                writer.add_histogram("gcn_conv{}_dist".format(i + 1), x, epoch)
        out = F.log_softmax(x, dim=-1)
        return out, unnorm_adj, x_dgg

    def dgg_net(self, x, i, unnorm_adj, writer, epoch):
        return self.dggs[i](x=x, adj=unnorm_adj, noise=False, writer=writer, epoch=epoch)


class GCN_DGG_Ablations(GCN_DGG_00):
    """Reference model.py:1436-1559 (same forward, DGG_Ablations inside)."""

    _dgg_cls = DGG_Ablations

    def dgg_net(self, x, i, unnorm_adj, writer, epoch):
        return self.dggs[i](x=x, adj=unnorm_adj, writer=writer, epoch=epoch)
