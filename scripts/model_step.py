"""A few eager GCN_DGG_00 training steps at Pubmed shape (for ncu launch lists)."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, model as models
import torch.nn.functional as F
from dgg_b200 import CSRGraph
shape = bench.PUBMED; dev = torch.device("cuda")
hs = bench.make_set(shape, 0)
adj = torch.sparse_coo_tensor(hs["idx"].to(dev), hs["val"].to(dev), (shape["n"],)*2, is_coalesced=True)
x = hs["x"].to(dev)
args = argparse.Namespace(extra_edge_dim=0, dgg_adj_input="input_adj")
net = models.GCN_DGG_00(nfeat=500, nlayers=2, nhidden=64, nclass=3, dropout=0.5, lamda=0.5, alpha=0.1, variant=False, args=args).to(dev)
opt = torch.optim.Adam([dict(params=net.params1, weight_decay=5e-4), dict(params=net.params2, weight_decay=0.0)], lr=0.01, fused=True)
labels = torch.randint(0, 3, (shape["n"],), device=dev); idx = torch.arange(60, device=dev)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    net.train(); opt.zero_grad()
    out, _, _ = net(x, adj)
    F.nll_loss(out[idx], labels[idx]).backward()
    opt.step()
torch.cuda.synchronize()
