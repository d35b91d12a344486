"""Pubmed-shape conv-layer kernels timed alone: SpMM fwd/bwd vs the fused SpMM+W layer, and the GCN_DGG_00 step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dgg_b200 import CSRGraph, functional as K
shape = bench.PUBMED
sets = [bench.make_set(shape, s) for s in range(3)]
n = shape["n"]
gs, xs, vs = [], [], []
for s in sets:
    g = CSRGraph.from_indices(s["idx"].cuda(), n); g.erow; gs.append(g)
    xs.append(torch.randn(n, 64, device="cuda")); vs.append(torch.rand(g.nnz, device="cuda") + 0.1)
w = torch.randn(64, 64, device="cuda") / 8
gy = torch.randn(n, 64, device="cuda")
def timed(fn, it=40):
    for i in range(6): fn(i)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(it): fn(i)
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it * 1e3
E = gs[0].nnz
print("E", E)
with torch.no_grad():
    print("spmm fwd F=64          : %.1f us" % timed(lambda i: K.spmm(vs[i % 3], xs[i % 3], gs[i % 3])))
    print("spmm_gemm fwd 64->64   : %.1f us" % timed(lambda i: K.spmm_gemm(vs[i % 3], xs[i % 3], w, gs[i % 3], relu=True)))
    print("mm N x 64 x 64         : %.1f us" % timed(lambda i: torch.mm(xs[i % 3], w)))
from dgg_b200._lib import lib, check, p as P, stream
L = lib()
ye = torch.zeros(n, 64, device="cuda"); dv = torch.empty(E, device="cuda"); dxe = torch.zeros(n, 64, device="cuda")
def efwd(i):
    g = gs[i % 3]; ye.zero_()
    check(L.dggb_spmm_edge_fwd(P(g.rowptr), P(g.erow), P(g.col), P(vs[i % 3]), n, g.nnz, P(xs[i % 3]), 64, None, P(ye), stream()), "e")
def ebwd(i):
    g = gs[i % 3]
    check(L.dggb_spmm_edge_bwd(P(g.erow), P(g.col), P(vs[i % 3]), g.nnz, P(xs[i % 3]), 64, None, P(gy), P(dv), P(dxe), stream()), "e")
def rbwd(i):
    g = gs[i % 3]
    check(L.dggb_spmm_csr_bwd(P(g.rowptr), P(g.col), P(vs[i % 3]), n, P(xs[i % 3]), 64, None, P(gy), P(dv), P(dxe), stream()), "e")
print("spmm EDGE fwd (+zero)  : %.1f us" % timed(efwd))
print("spmm EDGE bwd          : %.1f us" % timed(ebwd))
print("spmm row  bwd          : %.1f us" % timed(rbwd))
def fb(fused):
    def f(i):
        v = vs[i % 3].clone().requires_grad_(True); x = xs[i % 3].clone().requires_grad_(True); ww = w.clone().requires_grad_(True)
        if fused: y = K.spmm_gemm(v, x, ww, gs[i % 3], relu=True)
        else: y = torch.relu(torch.mm(K.spmm(v, x, gs[i % 3]), ww))
        y.backward(gy)
    return f
print("layer fwd+bwd fused    : %.1f us" % timed(fb(True)))
print("layer fwd+bwd unfused  : %.1f us" % timed(fb(False)))
dsets = []
for s in sets * 2:
    adj = torch.sparse_coo_tensor(s["idx"].cuda(), s["val"].cuda(), (n, n), is_coalesced=True)
    CSRGraph.from_coo(adj)
    dsets.append(dict(adj=adj, x=s["x"].cuda()))
print(bench.full_model_epoch(dsets, shape, torch.device("cuda")))
