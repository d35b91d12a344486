"""Auxiliary adjacency losses of the large-graph drivers without the dense N x N round trip.

``train_large_graphs.train_gcn_dgg`` / ``train_reddit`` add ``1e4 * F.mse_loss(out_adj.to_dense(), gt_adj.to_dense())``
to the classification loss (train_large_graphs.py:247-251), ``gt_adj`` being the input graph with its inter-class
edges removed (``utils.remove_interclass_edges``, utils.py:1310-1326 -- already O(E), re-exported unchanged).  Two
[N, N] fp32 tensors per step: 217 GB each at Reddit scale.  On the stored entries only,

    mse = ( sum_o o^2  -  2 sum_{o and g} o g  +  sum_g g^2 ) / N^2

which is what ``sparse_adj_mse`` evaluates (same value, same gradient w.r.t. the DGG output's values).  A maintainer
swaps the one call; the scripts' flags and everything else stay as they are."""
from __future__ import annotations

import torch


def _coalesced(a):
    return a if a.is_coalesced() else a.coalesce()


def sparse_adj_mse(out_adj: torch.Tensor, gt_adj: torch.Tensor) -> torch.Tensor:
    """== F.mse_loss(out_adj.to_dense(), gt_adj.to_dense()) for sparse COO [N, N] inputs; differentiable in the values
    of ``out_adj`` (and of ``gt_adj``).  O((E_o + E_g) log E) time and memory."""
    assert out_adj.shape == gt_adj.shape and out_adj.dim() == 2
    n_rows, n_cols = out_adj.shape
    o, g = _coalesced(out_adj), _coalesced(gt_adj)
    ov = getattr(out_adj, "_dgg_vals", None)          # the autograd-tracked values DGG attached (same order)
    ov = o.values() if ov is None else ov
    gv = g.values().to(ov.dtype)
    okey = o.indices()[0] * n_cols + o.indices()[1]
    gkey = g.indices()[0] * n_cols + g.indices()[1]
    cross = ov.new_zeros(())
    if okey.numel() and gkey.numel():
        pos = torch.searchsorted(okey, gkey).clamp(max=okey.numel() - 1)
        hit = okey[pos] == gkey
        cross = (ov[pos[hit]] * gv[hit]).sum()
    return ((ov * ov).sum() - 2.0 * cross + (gv * gv).sum()) / float(n_rows * n_cols)
