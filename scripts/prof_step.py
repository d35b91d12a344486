"""Where does the Pubmed-shape DGG step spend its time? (CPU issue time vs GPU time)"""
import cProfile, pstats, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, dgm
from dgg_b200 import CSRGraph, functional as K
from dgg_b200._lib import check, i32, lib, p, stream
import torch.nn.functional as F

shape = bench.PUBMED
dev = torch.device("cuda")
hs = [bench.make_set(shape, s) for s in range(3)]
m = dgm.DGG(in_dim=shape["f"], latent_dim=shape["h"], args=bench.dgg_args()); m.load_state_dict(bench.ref_state(shape)); m = m.to(dev)
ds = []
for s in hs:
    adj = torch.sparse_coo_tensor(s["idx"].to(dev), s["val"].to(dev), (shape["n"],)*2, is_coalesced=True)
    CSRGraph.from_coo(adj)
    ds.append(dict(adj=adj, x=s["x"].to(dev), g_vals=s["g_vals"].to(dev), g_xenc=s["g_xenc"].to(dev)))
params = list(m.parameters())
def step(i):
    s = ds[i % 3]
    for q in params: q.grad = None
    out, x_enc = m(s["x"], s["adj"])
    torch.autograd.backward([out._dgg_vals, x_enc], [s["g_vals"], s["g_xenc"]])
def wall(fn, it=200):
    for i in range(10): fn(i)
    torch.cuda.synchronize(); t=time.perf_counter()
    for i in range(it): fn(i)
    t_issue = time.perf_counter()-t
    torch.cuda.synchronize(); return t_issue/it*1e6, (time.perf_counter()-t)/it*1e6
print("full step: issue %.1f us, total %.1f us" % wall(step))
lin, dd = m.edge_encoder[0], m.degree_decoder[0]
g, _ = CSRGraph.from_coo(ds[0]["adj"])
with torch.no_grad():
    y = F.linear(m.node_encoder(ds[0]["x"]), lin.weight)
def fwd_apply(i):
    with torch.no_grad():
        return K._DGGEdge.apply(y, lin.bias, dd.weight, dd.bias, g, None, -1)
print("edge fwd via Function.apply: issue %.1f us, total %.1f us" % wall(fwd_apply))
n, h, E = shape["n"], shape["h"], g.nnz
R = torch.empty(E, device=dev); rank = torch.empty(E, dtype=torch.int32, device=dev); s_ = torch.empty(n, device=dev); k_ = torch.empty(n, device=dev); out = torch.empty(E, device=dev)
be = lin.bias.detach(); dw = dd.weight.detach().reshape(-1); db = dd.bias.detach().reshape(-1)
L = lib()
def fwd_raw(i):
    check(L.dggb_dgg_edge_fwd(p(g.rowptr), p(g.erow), p(g.col), i32(n), i32(E), i32(h), p(y), p(be), p(dw), p(db), p(None), i32(-1), p(R), p(rank), p(s_), p(k_), p(out), stream()), "f")
print("edge fwd raw ctypes: issue %.1f us, total %.1f us" % wall(fwd_raw))
def allocs(i):
    a = torch.empty(E, dtype=torch.float32, device=dev); b = torch.empty(E, dtype=torch.int32, device=dev); c = torch.empty(n, device=dev); d = torch.empty(n, device=dev); e = torch.empty(E, device=dev)
print("5x torch.empty: issue %.1f us, total %.1f us" % wall(allocs))
def nodeenc(i):
    with torch.no_grad():
        return F.linear(m.node_encoder(ds[i%3]["x"]), lin.weight)
print("node encoder + y (torch): issue %.1f us, total %.1f us" % wall(nodeenc))
pr = cProfile.Profile(); pr.enable()
for i in range(200): step(i)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
