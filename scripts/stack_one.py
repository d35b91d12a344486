"""A few forward launches of the GCNII stack kernel at Citeseer shape (for an ncu capture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from dgg_b200 import CSRGraph, functional as K
n, f, nl = 3327, 64, 63
idx, val = bench.chung_lu_graph(n, 2.8, 60, 7)
g = CSRGraph.from_indices(idx.cuda(), n)
v = torch.rand(g.nnz, device="cuda") * 0.2
x = torch.randn(n, f, device="cuda"); h0 = torch.randn(n, f, device="cuda")
ws = [torch.randn(f, f, device="cuda") / 8 for _ in range(nl)]
keep = (torch.rand(nl, n, f, device="cuda") > 0.4).float() / 0.6
with torch.no_grad():
    for _ in range(4):
        y = K.gcnii_stack(v, x, h0, ws, g, 0.9, 0.1, [0.3] * nl, keep=keep)
torch.cuda.synchronize()
