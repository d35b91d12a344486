"""CPU: the sparse adjacency MSE equals ``F.mse_loss(out_adj.to_dense(), gt_adj.to_dense())`` of
train_large_graphs.py:247-251 in value and gradient, for overlapping / disjoint / empty supports."""
import pytest
import torch
import torch.nn.functional as F

from dgg_b200.losses import sparse_adj_mse


def _rand_coo(n, nnz, seed, requires_grad=False):
    g = torch.Generator().manual_seed(seed)
    idx = torch.stack([torch.randint(0, n, (nnz,), generator=g), torch.randint(0, n, (nnz,), generator=g)])
    a = torch.sparse_coo_tensor(idx, torch.rand(nnz, generator=g) + 0.1, (n, n)).coalesce()
    vals = a.values().clone().requires_grad_(requires_grad)
    return torch.sparse_coo_tensor(a.indices(), vals, (n, n), is_coalesced=True), vals


@pytest.mark.parametrize("n,nnz_o,nnz_g", [(50, 300, 200), (40, 100, 0), (64, 0, 80), (30, 500, 500)])
def test_sparse_adj_mse_matches_dense(n, nnz_o, nnz_g):
    out, ov = _rand_coo(n, nnz_o, 1, requires_grad=True)
    gt, _ = _rand_coo(n, nnz_g, 2)
    if nnz_o and nnz_g:                      # force some overlap: the ground truth keeps part of the output's support
        k = min(nnz_g, out._nnz()) // 2
        gi = torch.cat([gt.indices(), out.indices()[:, :k]], 1)
        gt = torch.sparse_coo_tensor(gi, torch.ones(gi.shape[1]), (n, n)).coalesce()
    got = sparse_adj_mse(out, gt)
    want = F.mse_loss(out.to_dense(), gt.to_dense())
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-8)
    if nnz_o:
        (g1,) = torch.autograd.grad(got, ov, retain_graph=True)
        (g2,) = torch.autograd.grad(want, ov)
        torch.testing.assert_close(g1, g2, rtol=1e-5, atol=1e-9)
