"""GPU: the models under the training scripts' exact call conventions, one train step each.

"small"  = keyword style of train_small_graphs.train_debug / train_large_graphs.train_gcn:
           ``output, out_adj, x_dgg = model(features, adj, edge_index=..., epoch=..., writer=...)``  (268-270)
"pubmed" = positional style of train_pubmed.train: ``output, out_adj, x_dgg = model(x, adj, epoch, writer)`` (242)
Optimisers are chosen from the model NAME exactly like train_small_graphs.py:399-418.  The argument namespaces
restate the scripts' parser defaults (train_small_graphs.py:20-207; train_pubmed.py:29-203 lacks debug_step /
perturb_edge_prob / symmetric_noise / stochastic_k)."""
import argparse

import pytest
import torch
import torch.nn.functional as F

from tests.helpers import coo, random_graph

pytestmark = pytest.mark.gpu

SMALL = dict(layer=16, hidden=64, dropout=0.6, lamda=0.5, alpha=0.1, variant=False, lr=0.01, wd1=0.01, wd2=5e-4,
             extra_edge_dim=0, extra_k_dim=1, dgg_hard=False, deg_mean=3.899, deg_std=5.288, debug_step=3,
             dgm_dim=128, dgm_temp=10, n_dgg_layers=2, symmetric_noise=True, perturb_edge_prob=False,
             stochastic_k=False, pre_normalize_adj=False, dgg_adj_input="input_adj", dgg_mode_edge_net="u-v-deg",
             dgg_mode_k_net="x", dgg_mode_k_select="k_times_edge_prob")
PUBMED = {k: v for k, v in SMALL.items() if k not in ("debug_step", "perturb_edge_prob", "symmetric_noise",
                                                      "stochastic_k")}
PUBMED.update(extra_edge_dim=2, n_dgg_layers=1, dgg_mode_k_select="edge_p-cdf", dgg_mode_k_net="pass")


def _optimizer(name, model, args):
    if "GCN" in name and "II" in name:
        return torch.optim.Adam([{"params": model.params1, "weight_decay": args.wd1},
                                 {"params": model.params2, "weight_decay": args.wd2}], lr=args.lr)
    if "GCN" in name:
        return torch.optim.Adam([dict(params=model.params1, weight_decay=5e-4),
                                 dict(params=model.params2, weight_decay=0)], lr=args.lr)
    if "SAGE" in name:
        return torch.optim.Adam(model.parameters(), lr=args.lr)
    return torch.optim.Adam(model.parameters(), lr=0.005, weight_decay=5e-4)


def _data(n=400, f=48, c=5):
    idx, val = random_graph(n, 6, seed=11, self_loops=False)
    gen = torch.Generator().manual_seed(12)
    x = torch.rand(n, f, generator=gen)
    x = x / x.sum(-1, keepdim=True)
    y = torch.randint(0, c, (n,), generator=gen)
    mask = torch.zeros(n, dtype=torch.bool)
    mask[:60] = True
    return coo(idx, val, n).cuda(), idx.cuda(), x.cuda(), y.cuda(), mask.cuda(), f, c


def _build(name, args, f, c):
    import model as models

    torch.manual_seed(0)
    return models.__dict__[name](nfeat=f, nlayers=args.layer, nhidden=args.hidden, nclass=c, dropout=args.dropout,
                                 lamda=args.lamda, alpha=args.alpha, variant=args.variant, args=args).cuda()


def _step(model, opt, call, y, mask, expect_tuple):
    model.train()
    opt.zero_grad()
    res = call()
    if expect_tuple:
        output, out_adj, x_dgg = res                                   # the scripts' 3-tuple unpack
    else:
        assert torch.is_tensor(res)                                    # SAGE_DGG / GCNII_DGG return the tensor only
        output = res
    loss = F.nll_loss(output[mask], y[mask])
    loss.backward()
    before = [p.detach().clone() for p in model.parameters()]
    opt.step()
    assert torch.isfinite(loss)
    assert any(not torch.equal(a, b) for a, b in zip(before, model.parameters()))
    return output


SMALL_OK = ["GCN_DGG_00", "GCN_DGG_Ablations", "GCN_DGG_00_LargeGraphs", "SAGE_DGG_00", "GAT_DGG_00",
            "GAT_DGG_Ablations", "GCN", "SAGE", "GAT"]


@pytest.mark.parametrize("name", SMALL_OK)
def test_small_keyword_style(name):
    args = argparse.Namespace(**SMALL)
    adj, ei, x, y, mask, f, c = _data()
    m = _build(name, args, f, c)
    opt = _optimizer(name, m, args)
    if name == "GCN_DGG_00_LargeGraphs":     # sigmoid head: trained with BCE in train_large_graphs_multiclass.py
        m.train()
        out, out_adj, x_dgg = m(x, adj, edge_index=ei, epoch=3, writer=None)
        assert out.shape == (x.shape[0], c) and x_dgg is None and float(out.min()) >= 0 and float(out.max()) <= 1
        out.sum().backward()
        return
    out = _step(m, opt, lambda: m(x, adj, edge_index=ei, epoch=3, writer=None), y, mask, True)
    assert out.shape == (x.shape[0], c)
    m.eval()
    with torch.no_grad():
        output, out_adj, _ = m(x, adj, edge_index=ei,)                 # validate(): train_small_graphs.py:296
    torch.testing.assert_close(output.exp().sum(-1), torch.ones(x.shape[0], device="cuda"), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", ["GCN_DGG", "GCNII_DGG"])
def test_small_keyword_style_type_errors_like_the_reference(name):
    """model.py:1236 / 693 take no ``edge_index`` keyword: the README command raises TypeError in the reference,
    and must here."""
    args = argparse.Namespace(**dict(SMALL, extra_edge_dim=2))
    adj, ei, x, y, mask, f, c = _data()
    m = _build(name, args, f, c)
    with pytest.raises(TypeError):
        m(x, adj, edge_index=ei, epoch=0, writer=None)


@pytest.mark.parametrize("name,tuple_out", [("GCN_DGG_00", True), ("GCN_DGG_Ablations", True), ("SAGE_DGG_00", True),
                                            ("GCN_DGG", True), ("GCNII_DGG", False), ("SAGE_DGG", False)])
def test_pubmed_positional_style(name, tuple_out):
    """train_pubmed passes (x, adj, epoch, writer) positionally; its parser has no debug_step & co, so the
    DGG_LearnableK_debug-based models need those attributes from the small-graphs parser (the reference raises
    AttributeError without them) -- they are exercised with the merged namespace."""
    base = PUBMED if name.endswith("_00") or "Ablations" in name else dict(SMALL, extra_edge_dim=2)
    if name.endswith("_00") or "Ablations" in name:
        base = dict(base, extra_edge_dim=0)                            # class DGG feeds no extra edge features
    args = argparse.Namespace(**base)
    adj, ei, x, y, mask, f, c = _data()
    m = _build(name, args, f, c)
    opt = _optimizer(name, m, args)
    out = _step(m, opt, lambda: m(x, adj, 5, None), y, mask, tuple_out)
    assert out.shape == (x.shape[0], c)


def test_pubmed_parser_without_debug_step_raises_attribute_error_like_the_reference():
    args = argparse.Namespace(**dict(PUBMED, dgg_mode_k_net="x", dgg_mode_k_select="k_times_edge_prob"))
    adj, ei, x, y, mask, f, c = _data()
    m = _build("GCN_DGG", args, f, c)
    with pytest.raises(AttributeError):
        m(x, adj, 5, None)


def test_gat_pubmed_positional_style_type_error_like_the_reference():
    """GAT_DGG_00.forward(x, in_adj, edge_index, ...): the positional epoch lands in ``edge_index`` (an int) and the
    reference fails inside remove_self_loops."""
    args = argparse.Namespace(**dict(PUBMED, extra_edge_dim=0))
    adj, ei, x, y, mask, f, c = _data()
    m = _build("GAT_DGG_00", args, f, c)
    with pytest.raises(TypeError):
        m(x, adj, 5, None)
