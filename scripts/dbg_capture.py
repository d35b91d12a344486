"""Which piece of the Cora-shape GCN_DGG train step breaks CUDA-graph capture?"""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import bench, dgg_b200, model as models, dgm
dev = torch.device("cuda")
lk = dict(extra_edge_dim=2, extra_k_dim=1, dgg_hard=False, deg_mean=3.899, deg_std=5.288, dgg_mode_edge_net="u-v-deg",
          dgg_mode_k_net="x", dgg_mode_k_select="k_times_edge_prob", debug_step=3, perturb_edge_prob=False,
          symmetric_noise=True, stochastic_k=False, dgg_adj_input="input_adj", n_dgg_layers=2)
n, f, c = 2708, 1433, 7
idx, val = bench.chung_lu_graph(n, 3.9, 60, 7); keep = idx[0] != idx[1]; idx, val = idx[:, keep].contiguous(), val[keep]
x = bench._sparse_features(n, f, 0.013, 8).to(dev)
adj = torch.sparse_coo_tensor(idx.to(dev), val.to(dev), (n, n)).coalesce(); dgg_b200.CSRGraph.from_coo(adj)
labels = torch.randint(0, c, (n,), device=dev); ti = torch.arange(140, device=dev)
torch.manual_seed(0)
net = models.GCN_DGG(nfeat=f, nlayers=2, nhidden=64, nclass=c, dropout=0.6, lamda=0.5, alpha=0.1, variant=False, args=argparse.Namespace(**lk)).to(dev)
dgg = net.dggs[0]
from model import add_self_loops_coo
def try_capture(name, fn):
    try:
        gs = dgg_b200.GraphedStep(fn); gs(); torch.cuda.synchronize(); print("OK  ", name, flush=True)
    except Exception as e:
        print("FAIL", name, repr(e)[:160], flush=True); raise SystemExit
adjl = add_self_loops_coo(adj)
try_capture("self loops", lambda: add_self_loops_coo(adj)._dgg_vals)
try_capture("edge_prob_net", lambda: dgg.edge_prob_net(*dgg_b200.CSRGraph.from_coo(adjl), x, mode="u-v-deg"))
try_capture("k_net", lambda: dgg.k_estimate_net(n, *dgg_b200.CSRGraph.from_coo(adjl), x, None, mode="x"))
try_capture("dgg fwd", lambda: dgg(x, adjl)._dgg_vals)
def fb():
    for p in net.parameters(): p.grad = None
    out = dgg(x, adjl)._dgg_vals; out.sum().backward(); return out
try_capture("dgg fwd+bwd", fb)
net.eval()
try_capture("model fwd eval", lambda: net(x, adj)[0])
net.train()
try_capture("model fwd train", lambda: net(x, adj)[0])
def step():
    for p in net.parameters(): p.grad = None
    loss = F.nll_loss(net(x, adj)[0][ti], labels[ti]); loss.backward(); return loss
try_capture("model fwd+bwd", step)

groups = [dict(params=net.params1, weight_decay=5e-4), dict(params=net.params2, weight_decay=0)]
opt = torch.optim.Adam(groups, lr=0.01, capturable=True, fused=True)
def full():
    net.train(); opt.zero_grad(set_to_none=True)
    res = net(x, adj); loss = F.nll_loss(res[0][ti], labels[ti]); loss.backward(); opt.step(); return loss
try_capture("full step with fused capturable Adam", full)
opt2 = torch.optim.Adam(groups, lr=0.01, fused=True)
def eager():
    net.train(); opt2.zero_grad(set_to_none=True)
    res = net(x, adj); loss = F.nll_loss(res[0][ti], labels[ti]); loss.backward(); opt2.step(); return loss
for _ in range(3): eager()
opt3 = torch.optim.Adam(groups, lr=0.01, capturable=True, fused=True)
def full3():
    net.train(); opt3.zero_grad(set_to_none=True)
    res = net(x, adj); loss = F.nll_loss(res[0][ti], labels[ti]); loss.backward(); opt3.step(); return loss
try_capture("same after an eager non-capturable Adam ran on the same params", full3)
