import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from dgg_b200 import CSRGraph, functional as K
shape = bench.PUBMED; n = shape["n"]
s = bench.make_set(shape, 0)
g = CSRGraph.from_indices(s["idx"].cuda(), n); g.erow
heads, f = 8, 64
hd = torch.randn(n, heads * f, device="cuda").requires_grad_(True)
pq = torch.randn(n, heads, 2, device="cuda").requires_grad_(True)
av = (torch.rand(g.nnz, device="cuda") + 0.5).requires_grad_(True)
bias = torch.zeros(heads * f, device="cuda")
gy = torch.randn(n, heads * f, device="cuda")
def run():
    out = K.gat_aggregate(hd, pq, g, heads, f, avals=av, htot=hd.sum(0), bias=bias, alpha=0.2, bg=float(n))
    return out
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it * 1e3
with torch.no_grad():
    print("fwd  %.1f us" % t(run))
def fb():
    hd.grad = pq.grad = av.grad = None
    run().backward(gy)
print("fwd+bwd  %.1f us" % t(fb))
