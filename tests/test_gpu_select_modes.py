"""GPU: select modes and hub rows of DGG_LearnableK_debug against the dense CPU oracle.

* ``k_only`` (dgm.py:1423-1435) incl. the spill of a row's window into its first non-edge columns;
* a 6 000-entry hub row ranked by the grid-wide long-row kernel (first-k on CSR rows);
* off-edge TensorBoard scalars of get_adj_diff_stats (dgm.py:1313-1350)."""
import argparse

import pytest
import torch

from oracle import dgg_oracle as O
from tests.helpers import assert_grad_close, coo, near_tie_entries, random_graph, sparse_ranks

pytestmark = pytest.mark.gpu


def _args(**kw):
    d = dict(extra_edge_dim=0, extra_k_dim=1, dgg_hard=False, deg_mean=3.899, deg_std=5.288,
             dgg_mode_edge_net="u-v-dist", dgg_mode_k_net="x", dgg_mode_k_select="k_times_edge_prob",
             debug_step=3, perturb_edge_prob=False, symmetric_noise=True, stochastic_k=False,
             dgg_adj_input="input_adj", n_dgg_layers=2)
    d.update(kw)
    return argparse.Namespace(**d)


class _Writer:
    def __init__(self):
        self.scalars = {}

    def add_scalar(self, tag, value, step=None):
        self.scalars[tag] = float(value)

    def add_histogram(self, *a, **k):
        pass


@pytest.mark.parametrize("edge_mode,extra", [("u-v-dist", 0), ("u-v-deg", 2)])
def test_k_only_matches_oracle_including_spill(edge_mode, extra):
    import dgm

    n, f, h = 300, 24, 16
    idx, val = random_graph(n, 5, seed=2)
    gen = torch.Generator().manual_seed(3)
    x = torch.rand(n, f, generator=gen)
    x = x / x.sum(-1, keepdim=True)
    args = _args(dgg_mode_k_select="k_only", dgg_mode_edge_net=edge_mode, extra_edge_dim=extra)
    torch.manual_seed(4)
    m = dgm.DGG_LearnableK_debug(in_dim=f, latent_dim=h, args=args)
    with torch.no_grad():
        m.node_encode_for_edges[0].weight.mul_(6.0)
    state = {k: v.clone() for k, v in m.state_dict().items()}
    p = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    r = O.learnable_k_forward(x, idx, val, n, p, edge_mode=edge_mode, k_mode="x", select_mode="k_only")
    wt = torch.randn(n, n, generator=gen)
    (r["out"] * wt).sum().backward()
    m = m.cuda()
    w = _Writer()
    m.train()
    out = m(x.cuda(), coo(idx, val, n).cuda(), writer=w, epoch=0)
    dense = out.to_dense()
    # rows with near-tied probabilities may order two edges differently; everything else must agree everywhere
    near_rows = torch.unique(idx[0][near_tie_entries(idx, r["P"].detach(), n)])
    keep = torch.ones(n, dtype=torch.bool)
    keep[near_rows] = False
    assert keep.float().mean() > 0.9
    torch.testing.assert_close(dense.cpu()[keep], r["out"].detach()[keep], rtol=1e-5, atol=2e-6)
    assert int((r["out"].detach()[keep] != 0).sum()) > int(keep.sum()) * 8          # windows longer than the degrees
    spilled = (r["out"].detach() != 0) & (O.dense_from_edges(idx, val, n) == 0)
    assert int(spilled[keep].sum()) > 0                                            # ... so the spill is exercised
    (dense * (wt * keep.unsqueeze(-1)).cuda()).sum().backward()
    p2 = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    r2 = O.learnable_k_forward(x, idx, val, n, p2, edge_mode=edge_mode, k_mode="x", select_mode="k_only")
    (r2["out"] * wt * keep.unsqueeze(-1)).sum().backward()
    for name, q in m.named_parameters():
        want = p2[name].grad
        if want is None or float(want.abs().max()) == 0.0:
            assert q.grad is None or float(q.grad.abs().max()) == 0.0, name        # k_only: no gradient to the edge net
        else:
            assert_grad_close(q.grad.cpu(), want, what=name)
    # off-edge statistics (the spilled entries) against the dense definition of dgm.py:1321-1330
    ind = O.dense_from_edges(idx, val, n)
    off = ((ind - dense.detach().cpu()) * (ind == 0).float())
    off = off[off != 0]
    assert w.scalars["train_stats/off_edge_mean"] == pytest.approx(float(off.mean()), rel=1e-4)
    assert w.scalars["train_stats/off_edge_std"] == pytest.approx(float(off.std()), rel=1e-3)


def test_first_k_hub_row_uses_grid_wide_ranking():
    from dgg_b200 import CSRGraph
    from dgg_b200 import functional as K

    n, hub_deg = 9000, 6000
    gen = torch.Generator().manual_seed(5)
    hub_cols = torch.randperm(n, generator=gen)[:hub_deg]
    src = torch.cat([torch.full((hub_deg,), 17), torch.randint(0, n, (4 * n,), generator=gen)])
    dst = torch.cat([hub_cols, torch.randint(0, n, (4 * n,), generator=gen)])
    a = torch.sparse_coo_tensor(torch.stack([src, dst]), torch.ones(src.numel()), (n, n)).coalesce()
    idx = a.indices()
    g = CSRGraph.from_indices(idx.cuda(), n)
    assert g.max_row_nnz > K._LONG_ROW
    score = torch.rand(idx.shape[1], generator=gen)
    k = 1.0 + 40.0 * torch.rand(n, generator=gen)
    sc, kc = score.cuda().requires_grad_(True), k.cuda().requires_grad_(True)
    out, rank = K.row_firstk(sc, kc, g, return_rank=True)
    want_rank = sparse_ranks(idx, score, n)
    assert torch.equal(rank.cpu().long(), want_rank)                               # exact (ties: lower column first)
    fk = 1 - 0.5 * (1 + torch.tanh(want_rank.float() - k[idx[0]]))
    torch.testing.assert_close(out.detach().cpu(), score * fk, rtol=1e-5, atol=1e-7)
    wt = torch.randn(idx.shape[1], generator=gen)
    (out * wt.cuda()).sum().backward()
    so, ko = score.clone().requires_grad_(True), k.clone().requires_grad_(True)
    ((so * (1 - 0.5 * (1 + torch.tanh(want_rank.float() - ko[idx[0]])))) * wt).sum().backward()
    assert_grad_close(sc.grad.cpu(), so.grad, what="score")
    assert_grad_close(kc.grad.cpu(), ko.grad, what="k")


def test_dgg_hub_row_two_launch_path_with_long_row_kernel():
    """class DGG on a graph whose hub row (3 000 entries) exceeds both the fused kernels' 512 and the per-warp
    ranking's 1 024: scores / ranks / values vs the sparse restatement of the oracle's formulas."""
    from dgg_b200 import CSRGraph
    from dgg_b200 import functional as K

    n, h, hub_deg = 5000, 32, 3000
    gen = torch.Generator().manual_seed(6)
    hub_cols = torch.randperm(n, generator=gen)[:hub_deg]
    src = torch.cat([torch.full((hub_deg,), 3), torch.randint(0, n, (3 * n,), generator=gen)])
    dst = torch.cat([hub_cols, torch.randint(0, n, (3 * n,), generator=gen)])
    a = torch.sparse_coo_tensor(torch.stack([src, dst]), torch.ones(src.numel()), (n, n)).coalesce()
    idx = a.indices()
    g = CSRGraph.from_indices(idx.cuda(), n)
    y = torch.randn(n, h, generator=gen)
    be = 0.1 * torch.randn(h, generator=gen)
    dw, db = torch.tensor([[0.05]]), torch.tensor([0.3])
    out, k, R, rank = K.dgg_edge(y.cuda(), be.cuda(), dw.cuda(), db.cuda(), g)
    pre = torch.nn.functional.leaky_relu(y[idx[0]] - y[idx[1]] + be, 0.01).sum(-1)
    R_ref = torch.sigmoid(pre)
    torch.testing.assert_close(R.cpu(), R_ref, rtol=1e-5, atol=1e-6)
    keep = ~near_tie_entries(idx, R_ref, n)
    assert torch.equal(rank.cpu().long()[keep], sparse_ranks(idx, R_ref, n)[keep])
    s = torch.zeros(n).index_add(0, idx[0], R_ref)
    k_ref = torch.nn.functional.leaky_relu(0.05 * s + 0.3, 0.01)
    torch.testing.assert_close(k.cpu(), k_ref, rtol=1e-5, atol=1e-5)
    want = R_ref * (2 - 0.5 * (1 + torch.tanh(sparse_ranks(idx, R_ref, n).float() - k_ref[idx[0]])))
    torch.testing.assert_close(out.cpu()[keep], want[keep], rtol=1e-5, atol=2e-6)
