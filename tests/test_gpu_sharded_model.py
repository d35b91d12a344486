"""GPU (1 device): RowShardedSAGE_DGG -- the Reddit-shape model of bench.py -- against a dense tensor-op restatement
of the same formulas (all-pairs distances, injected-equivalent Philox Gumbel noise materialised on the host, dense
sort, tanh first-k, D^-1/2 A D^-1/2, DenseGraphConv mean aggregation), forward and every gradient; and the row-block
decomposition the multi-GPU run uses (rank r = rows [r cnt, ...)) reproduced sequentially on one device."""
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import assert_grad_close
from tests.philox_ref import gumbel_matrix

pytestmark = pytest.mark.gpu


def _dense_forward(m, x, seed, kc):
    n = x.shape[0]
    z = m.input_project(x)
    k = torch.relu(m.k_net(x)) + 1.0
    d = torch.cdist(z.double(), z.double()).float()
    y = -m.t * d + gumbel_matrix(n, n, seed, 1.0)
    srt, order = torch.sort(y, dim=-1, descending=True, stable=True)
    r = torch.arange(n, dtype=torch.float32).reshape(1, -1)
    fk = 1 - 0.5 * (1 + torch.tanh(r - k))
    fk = fk * (r < kc)                                   # the selector keeps kc entries per row (fk is 0 beyond k + 8.47)
    a = torch.zeros_like(y).scatter(-1, order, torch.exp(srt) * fk)
    dinv = a.sum(-1) ** -0.5
    ahat = dinv.unsqueeze(-1) * a * dinv.unsqueeze(0)
    scale = 1.0 / ahat.sum(-1, keepdim=True).clamp(min=1)
    h = x
    for lr, lo, last in ((m.lin_rel1, m.lin_root1, False), (m.lin_rel2, m.lin_root2, True)):
        hn = (ahat @ h) * scale @ lr.weight.t() + lr.bias + lo(h)
        h = hn if last else torch.relu(hn)
    return torch.log_softmax(h, -1), y, d


def test_row_sharded_sage_dgg_matches_dense_restatement():
    import copy

    from dgg_b200 import sharding as S

    n, f, hdim, c, kc = 300, 40, 16, 5, 16
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(n, f, generator=gen)
    labels = torch.randint(0, c, (n,), generator=gen)
    torch.manual_seed(1)
    m = S.RowShardedSAGE_DGG(f, hdim, c, d=32, kc=kc, k_init=4.0, t_init=3.0).eval()
    # a swap of two near-tied scores inside a row's window moves weight between two columns, i.e. it changes the
    # output and every gradient: pick (deterministically) a noise seed whose selected scores are all > 5e-5 apart,
    # an order of magnitude above the 3xTF32-vs-fp64 distance error, so that EVERYTHING can be asserted tightly
    for seed in range(1, 200):
        with torch.no_grad():
            _, y, _ = _dense_forward(m, x, seed, kc)
            srt = torch.sort(y, -1, descending=True).values[:, :kc + 1]
        if float((srt[:, :-1] - srt[:, 1:]).min()) > 5e-5:
            break
    else:
        pytest.fail("no tie-free noise seed found")
    ref = copy.deepcopy(m)
    want, y, dmat = _dense_forward(ref, x, seed, kc)
    y.retain_grad()
    F.nll_loss(want, labels).backward()
    # dL/dt = -sum_ij (dL/dy_ij) D_ij cancels to ~1e-4 of the sum of its terms' magnitudes here (weights exp(y) reach
    # several hundred): its error floor is the 3xTF32 score error (~4e-6 absolute on y, i.e. relative on exp(y)) times THAT sum
    t_scale = float((y.grad.abs() * dmat).sum())
    m = m.cuda()
    got, idx, ahat = m(x.cuda(), n, seed=seed)
    torch.testing.assert_close(got.detach().cpu(), want.detach(), rtol=1e-4, atol=1e-4)
    F.nll_loss(got, labels.cuda()).backward()
    for (name, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        if name == "t":
            assert abs(float(p.grad) - float(q.grad)) <= 2e-4 * t_scale, (float(p.grad), float(q.grad), t_scale)
        else:
            assert_grad_close(p.grad.cpu(), q.grad, rtol=2e-3, atol_rel=2e-4, what=name)


def test_row_blocks_reproduce_the_unsharded_adjacency():
    """What each rank computes in the multi-GPU run (its row block against all columns) equals the corresponding
    rows of the single-device result bit for bit: Philox noise is keyed on the GLOBAL (row, col)."""
    from dgg_b200 import functional as K
    from dgg_b200 import sharding as S

    n, d, kc = 5000, 64, 32
    gen = torch.Generator().manual_seed(2)
    z = torch.softmax(torch.randn(n, d, generator=gen), -1).cuda()
    t = torch.tensor([4.0], device="cuda")
    full_i, full_v = K.allpairs_topk(z, t, None, kc, 3, seed=5, noise_scale=1.0)
    for world in (2, 3, 8):
        for rank in range(world):
            rb, cnt, _ = S.row_block(n, world, rank)
            i, v = K.allpairs_topk(z, t, None, kc, 3, rb, cnt, seed=5, noise_scale=1.0)
            assert torch.equal(i, full_i[rb:rb + cnt]) and torch.equal(v, full_v[rb:rb + cnt])
