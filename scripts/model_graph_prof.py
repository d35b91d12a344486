"""Warm per-kernel durations inside ONE CUDA-graph replay of the GCN_DGG_00 training step (CUPTI via torch.profiler;
rotating 6 input sets like bench.py, so x is not L2-resident across replays)."""
import os, sys, argparse, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, dgg_b200, model as models
import torch.nn.functional as F
from torch.profiler import profile, ProfilerActivity
shape = bench.PUBMED; dev = torch.device("cuda")
dsets = []
for i in range(bench.N_SETS):
    hs = bench.make_set(shape, i)
    adj = torch.sparse_coo_tensor(hs["idx"].to(dev), hs["val"].to(dev), (shape["n"],) * 2, is_coalesced=True)
    dsets.append(dict(x=hs["x"].to(dev), adj=adj))
args = argparse.Namespace(extra_edge_dim=0, dgg_adj_input="input_adj")
torch.manual_seed(0)
net = models.GCN_DGG_00(nfeat=shape["f"], nlayers=2, nhidden=shape["h"], nclass=3, dropout=0.5, lamda=0.5, alpha=0.1,
                        variant=False, args=args).to(dev)
opt = torch.optim.Adam([dict(params=net.params1, weight_decay=5e-4), dict(params=net.params2, weight_decay=0.0)],
                       lr=0.01, capturable=True, fused=True)
n = shape["n"]; labels = torch.randint(0, 3, (n,), device=dev); idx_train = torch.arange(60, device=dev); y_train = labels[idx_train]
def body(s):
    net.train(); opt.zero_grad(set_to_none=True)
    out, _, _ = net(s["x"], s["adj"])
    loss = F.nll_loss(out[idx_train], y_train); loss.backward(); opt.step(); return loss
for s in dsets[:2]: body(s)
graphs = [dgg_b200.GraphedStep(lambda s=s: body(s)) for s in dsets]
for i in range(12): graphs[i % 6]()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(6): graphs[i]()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type.name == "CUDA" or "cuda" in str(e.device_type).lower()]
per = collections.OrderedDict()
for e in evs:
    d = per.setdefault(e.name[:70], [0, 0.0]); d[0] += 1; d[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in per.values())
print(f"kernels per replay {sum(v[0] for v in per.values())/6:.1f}, kernel time per replay {tot/6:.1f} us")
for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v[1]/6:7.1f} us  x{v[0]/6:4.1f}  {k}")
