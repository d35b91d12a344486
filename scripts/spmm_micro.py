"""CSR SpMM fwd/bwd micro-benchmark: achieved GB/s by algorithmic bytes vs the measured HBM peak."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from dgg_b200 import CSRGraph, functional as K
peak = 6461.2
def graph(n, deg, seed):
    g = torch.Generator().manual_seed(seed)
    m = n * deg
    src = torch.randint(0, n, (m,), generator=g); dst = torch.randint(0, n, (m,), generator=g)
    a = torch.sparse_coo_tensor(torch.stack([src, dst]), torch.ones(m), (n, n)).coalesce()
    return a.indices().cuda(), a._nnz()
for (n, deg, f) in [(19717, 5, 64), (232965, 50, 64), (232965, 50, 128), (232965, 16, 602), (1000000, 16, 64)]:
    idx, nnz = graph(n, deg, n)
    G = CSRGraph.from_indices(idx, n)
    vals = torch.rand(nnz, device="cuda"); x = torch.randn(n, f, device="cuda"); dy = torch.randn(n, f, device="cuda")
    def timed(fn, it=10):
        for _ in range(3): fn()
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it): fn()
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/it*1e-3
    with torch.no_grad():
        t_f = timed(lambda: K.spmm(vals, x, G))
    vg = vals.clone().requires_grad_(True); xg = x.clone().requires_grad_(True)
    y = K.spmm(vg, xg, G)
    def bwd():
        vg.grad = None; xg.grad = None
        y.backward(dy, retain_graph=True)
    t_b = timed(bwd)
    b_f = nnz * (8 + f * 4) + 4 * (n + 1) + n * f * 4
    b_b = b_f + nnz * (f * 4 + 4) + nnz * f * 4
    print(json.dumps(dict(n=n, nnz=nnz, f=f, fwd_us=t_f*1e6, fwd_gbps=b_f/t_f/1e9, fwd_frac=b_f/t_f/1e9/peak,
                          bwd_us=t_b*1e6, bwd_gbps=b_b/t_b/1e9, bwd_frac=b_b/t_b/1e9/peak)), flush=True)
