"""Warm per-kernel durations inside ONE CUDA-graph replay of the Cora-shape GCN_DGG training step."""
import os, sys, argparse, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, dgg_b200, model as models
import torch.nn.functional as F
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda")
lk = dict(extra_edge_dim=2, extra_k_dim=1, dgg_hard=False, deg_mean=3.899, deg_std=5.288, dgg_mode_edge_net="u-v-deg",
          dgg_mode_k_net="x", dgg_mode_k_select="k_times_edge_prob", debug_step=3, perturb_edge_prob=False,
          symmetric_noise=True, stochastic_k=False, dgg_adj_input="input_adj", n_dgg_layers=2)
n, f, c = 2708, 1433, 7
idx, val = bench.chung_lu_graph(n, 3.9, 60, 7); keep = idx[0] != idx[1]; idx, val = idx[:, keep].contiguous(), val[keep]
x = bench._sparse_features(n, f, 0.0127, 8).to(dev)
adj = torch.sparse_coo_tensor(idx.to(dev), val.to(dev), (n, n)).coalesce(); dgg_b200.CSRGraph.from_coo(adj)
labels = torch.randint(0, c, (n,), device=dev); ti = torch.arange(140, device=dev)
torch.manual_seed(0)
net = models.GCN_DGG(nfeat=f, nlayers=2, nhidden=64, nclass=c, dropout=0.6, lamda=0.5, alpha=0.1, variant=False, args=argparse.Namespace(**lk)).to(dev)
opt = torch.optim.Adam([dict(params=net.params1, weight_decay=5e-4), dict(params=net.params2, weight_decay=0)], lr=0.01, capturable=True, fused=True)
def body():
    net.train(); opt.zero_grad(set_to_none=True)
    res = net(x, adj); logp = res[0] if isinstance(res, tuple) else res; loss = F.nll_loss(logp[ti], labels[ti]); loss.backward(); opt.step(); return loss
g = dgg_b200.GraphedStep(body)
for i in range(10): g()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(3): g()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if "cuda" in str(e.device_type).lower()]
per = collections.OrderedDict()
for e in evs:
    d = per.setdefault(e.name[:90], [0, 0.0]); d[0] += 1; d[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in per.values())
print(f"kernels per replay {sum(v[0] for v in per.values())/3:.1f}, kernel time per replay {tot/3:.1f} us")
for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{v[1]/3:8.1f} us  x{v[0]/3:6.1f}  {k}")
