"""GPU: the fused edge-probability kernels (csrc/edge_mlp.cu) against the reference's formulas written with plain
tensor ops on the CPU (dgm.py:1596-1727), forward and every gradient, all five node-feature modes; plus the module
path (``DGG_LearnableK_debug.edge_prob_net``) against the oracle with features whose width is not a multiple of 4
(Cora: 1433) so the padded tensor-core encoder is exercised."""
import argparse

import pytest
import torch
import torch.nn.functional as F

from oracle import dgg_oracle as O
from tests.helpers import assert_grad_close, coo, random_graph

pytestmark = pytest.mark.gpu


def _ref(mode, xe, idx, val, deg, w1, b1, w2, b2, h):
    u, v = xe[idx[0]], xe[idx[1]]
    if mode == "u-v-dist":
        return torch.exp(-0.05 * torch.linalg.vector_norm(u - v, dim=-1, ord=2))
    if mode == "u-v-A_uv":
        feat = torch.cat([u, v, val.unsqueeze(-1)], -1)
    elif mode == "u-v-deg":
        feat = torch.cat([u, v, deg[idx[0]].unsqueeze(-1), deg[idx[1]].unsqueeze(-1)], -1)
    else:
        d = torch.exp(-1.0 * torch.linalg.vector_norm(u - v, dim=-1, ord=2))
        feat = torch.cat([u, v, deg[idx[0]].unsqueeze(-1), deg[idx[1]].unsqueeze(-1), d.unsqueeze(-1)], -1)
    return torch.sigmoid(F.linear(F.leaky_relu(F.linear(feat, w1, b1), 0.01), w2, b2).flatten())


@pytest.mark.parametrize("mode,m", [("u-v-dist", 0), ("u-v-A_uv", 1), ("u-v-deg", 2), ("u-v-deg-dist", 3)])
@pytest.mark.parametrize("n,h,avg_deg", [(900, 64, 7), (300, 16, 4), (500, 128, 5)])
def test_edge_mlp_modes_vs_tensor_ops(mode, m, n, h, avg_deg):
    from dgg_b200 import CSRGraph
    from dgg_b200 import functional as K

    idx, _ = random_graph(n, avg_deg, seed=n + m, hubs=1, hub_deg=200)
    E = idx.shape[1]
    gen = torch.Generator().manual_seed(h + m)
    val = 0.5 + torch.rand(E, generator=gen)
    xe = torch.randn(n, h, generator=gen) * 0.5
    deg = torch.zeros(n).index_add(0, idx[0], val)
    w1 = torch.randn(h, 2 * h + m, generator=gen) / (2 * h) ** 0.5
    b1, w2, b2 = torch.randn(h, generator=gen) * 0.1, torch.randn(1, h, generator=gen) / h ** 0.5, torch.randn(1, generator=gen)
    wt = torch.randn(E, generator=gen)
    leaves = [t.clone().requires_grad_(True) for t in (xe, w1, b1, w2, b2)]
    want = _ref(mode, leaves[0], idx, val, deg, *leaves[1:], h)
    (want * wt).sum().backward()

    g = CSRGraph.from_indices(idx.cuda(), n)
    dl = [t.clone().cuda().requires_grad_(True) for t in (xe, w1, b1, w2, b2)]
    xc, w1c, b1c, w2c, b2c = dl
    if mode == "u-v-dist":
        got = K.edge_dist_score(xc, g, 0.05)
    else:
        p_uv = xc @ torch.cat([w1c[:, :h], w1c[:, h:2 * h]], 0).t()          # the tall GEMM (any implementation)
        kw = dict(edge_val=val.cuda(), flags=K.EX_VAL) if mode == "u-v-A_uv" else (
            dict(deg=deg.cuda(), flags=K.EX_DEG) if mode == "u-v-deg" else
            dict(deg=deg.cuda(), xe=xc, flags=K.EX_DEG | K.EX_DIST, dist_scale=1.0))
        got = K.edge_mlp(p_uv, w1c[:, 2 * h:], b1c, w2c, b2c, g, slope=0.01, **kw)
    torch.testing.assert_close(got.detach().cpu(), want.detach(), rtol=1e-5, atol=2e-6)
    (got * wt.cuda()).sum().backward()
    names = ["xe", "w1", "b1", "w2", "b2"]
    for nm, a, b in zip(names, dl, leaves):
        if b.grad is None:
            assert a.grad is None or float(a.grad.abs().max()) == 0.0, nm
        else:
            assert_grad_close(a.grad.cpu(), b.grad, what=f"{mode}:{nm}")


def test_edge_conv_vs_tensor_ops():
    from dgg_b200 import CSRGraph
    from dgg_b200 import functional as K

    n, h = 700, 64
    idx, _ = random_graph(n, 6, seed=3)
    gen = torch.Generator().manual_seed(4)
    xe = torch.randn(n, h, generator=gen) * 0.5
    th_w, th_b = torch.randn(h // 2, h, generator=gen) / h ** 0.5, torch.randn(h // 2, generator=gen) * 0.1
    ph_w, ph_b = torch.randn(h // 2, h, generator=gen) / h ** 0.5, torch.randn(h // 2, generator=gen) * 0.1
    en_w, en_b = torch.randn(1, h // 2, generator=gen), torch.randn(1, generator=gen)
    wt = torch.randn(idx.shape[1], generator=gen)
    leaves = [t.clone().requires_grad_(True) for t in (xe, th_w, th_b, ph_w, ph_b, en_w, en_b)]
    x_, tw, tb, pw, pb, ew, eb = leaves
    u, v = x_[idx[0]], x_[idx[1]]
    want = torch.sigmoid(F.linear(F.linear(v - u, tw, tb) + F.linear(u, pw, pb), ew, eb).flatten())   # dgm.py:1710-1715
    (want * wt).sum().backward()
    g = CSRGraph.from_indices(idx.cuda(), n)
    dl = [t.clone().cuda().requires_grad_(True) for t in (xe, th_w, th_b, ph_w, ph_b, en_w, en_b)]
    x_, tw, tb, pw, pb, ew, eb = dl
    p_uv = x_ @ torch.cat([pw - tw, tw], 0).t()
    got = K.edge_mlp(p_uv, None, tb + pb, ew, eb, g, flags=0, slope=1.0)
    torch.testing.assert_close(got.detach().cpu(), want.detach(), rtol=1e-5, atol=2e-6)
    (got * wt.cuda()).sum().backward()
    for nm, a, b in zip(["xe", "th_w", "th_b", "ph_w", "ph_b", "en_w", "en_b"], dl, leaves):
        assert_grad_close(a.grad.cpu(), b.grad, what=nm)


@pytest.mark.parametrize("mode,extra", [("u-v-deg", 2), ("u-v-deg-dist", 3), ("edge_conv", 0), ("u-v-dist", 0)])
def test_module_edge_prob_net_with_unaligned_feature_width(mode, extra):
    import dgm

    n, f, h = 1000, 1433, 64                       # Cora's feature width: not a multiple of 4 -> padded once
    idx, val = random_graph(n, 5, seed=9)
    gen = torch.Generator().manual_seed(10)
    x = (torch.rand(n, f, generator=gen) < 0.02).float()
    x = x / x.sum(-1, keepdim=True).clamp_min(1)
    args = argparse.Namespace(extra_edge_dim=extra, extra_k_dim=1, dgg_hard=False, deg_mean=3.899, deg_std=5.288,
                              dgg_mode_edge_net=mode, dgg_mode_k_net="x", dgg_mode_k_select="k_times_edge_prob",
                              debug_step=0, perturb_edge_prob=False, symmetric_noise=True, stochastic_k=False)
    torch.manual_seed(11)
    m = dgm.DGG_LearnableK_debug(in_dim=f, latent_dim=h, args=args)
    state = {k: v.clone() for k, v in m.state_dict().items()}
    p = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    want = O.edge_prob_net(x, idx, val, n, p, mode)
    wt = torch.randn(idx.shape[1], generator=gen)
    (want * wt).sum().backward()
    m = m.cuda()
    out = m(x.cuda(), coo(idx, val, n).cuda())             # debug_step == 0: returns the edge probabilities
    got = out.coalesce().values()
    torch.testing.assert_close(got.detach().cpu(), want.detach(), rtol=1e-5, atol=2e-6)
    (got * wt.cuda()).sum().backward()
    for name, q in m.named_parameters():
        ref = p[name].grad
        if ref is None or float(ref.abs().max()) == 0.0:
            assert q.grad is None or float(q.grad.abs().max()) == 0.0, name
        else:
            assert_grad_close(q.grad.cpu(), ref, what=name)
