"""Warm per-kernel durations inside ONE CUDA-graph replay of the Pubmed-shape GAT_DGG_00 training step (config 3)."""
import os, sys, argparse, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, dgg_b200, model as models
import torch.nn.functional as F
from torch.profiler import profile, ProfilerActivity
shape = bench.PUBMED; dev = torch.device("cuda"); n = shape["n"]
hs = bench.make_set(shape, 0)
ii = hs["idx"].to(dev); nl = ii[:, ii[0] != ii[1]].contiguous()
adj = torch.sparse_coo_tensor(nl, torch.ones(nl.shape[1], device=dev), (n, n)).coalesce()
x = hs["x"].to(dev)
torch.manual_seed(0)
net = models.GAT_DGG_00(nfeat=shape["f"], nlayers=2, nhidden=shape["h"], nclass=3,
                        args=argparse.Namespace(extra_edge_dim=0, dgg_adj_input="input_adj")).to(dev)
opt = torch.optim.Adam(net.parameters(), lr=0.005, weight_decay=5e-4, capturable=True, fused=True)
labels = torch.randint(0, 3, (n,), device=dev); ti = torch.arange(60, device=dev)
def body():
    net.train(); opt.zero_grad(set_to_none=True)
    logp, _, _ = net(x, adj, edge_index=nl)
    loss = F.nll_loss(logp[ti], labels[ti]); loss.backward(); opt.step(); return loss
g = dgg_b200.GraphedStep(body)
for i in range(10): g()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(3): g()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if "cuda" in str(e.device_type).lower()]
per = collections.OrderedDict()
for e in evs:
    d = per.setdefault(e.name[:100], [0, 0.0]); d[0] += 1; d[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in per.values())
print(f"kernels per replay {sum(v[0] for v in per.values())/3:.1f}, kernel time per replay {tot/3:.1f} us")
for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{v[1]/3:8.1f} us  x{v[0]/3:6.1f}  {k}")
