"""Citeseer-shape layer kernels inside a CUDA graph (20 back-to-back launches per replay)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from dgg_b200 import CSRGraph
from dgg_b200._lib import lib, check, p as P, stream
n = int(os.environ.get("NN", 3327)); h = 64
idx, val = bench.chung_lu_graph(n, 2.8, 60, 7)
g = CSRGraph.from_indices(idx.cuda(), n); g.erow
E = g.nnz
x = torch.randn(n, h, device="cuda"); v = torch.rand(E, device="cuda") + .1; w = torch.randn(h, h, device="cuda") / 8
y = torch.empty(n, h, device="cuda"); s = torch.empty(n, h, device="cuda"); gy = torch.randn(n, h, device="cuda")
dv = torch.empty(E, device="cuda"); dx = torch.zeros(n, h, device="cuda"); ds = torch.empty(n, h, device="cuda")
L = lib()
def gt(fn, k=20):
    st = torch.cuda.Stream(); st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        fn()
    torch.cuda.current_stream().wait_stream(st)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(k): fn()
    t_end = time.perf_counter() + 0.2
    while time.perf_counter() < t_end: gr.replay()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): gr.replay()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / (50 * k) * 1e3
print("N", n, "E", E)
print("spmm_csr_fwd        %.2f us" % gt(lambda: check(L.dggb_spmm_csr_fwd(P(g.rowptr), P(g.col), P(v), n, P(x), h, None, P(y), stream()), "a")))
print("spmm_gemm_fwd       %.2f us" % gt(lambda: check(L.dggb_spmm_gemm_fwd(P(g.rowptr), P(g.col), P(v), n, P(x), h, None, P(x), 0.9, 0.1, P(w), h, 0.4, 0.6, None, 1, None, P(y), P(s), stream()), "b")))
def efwd():
    y.zero_()
    check(L.dggb_spmm_edge_fwd(P(g.rowptr), P(g.erow), P(g.col), P(v), n, E, P(x), h, None, P(y), stream()), "e")
print("spmm_edge_fwd+zero  %.2f us" % gt(efwd))
print("spmm_csr_bwd        %.2f us" % gt(lambda: check(L.dggb_spmm_csr_bwd(P(g.rowptr), P(g.col), P(v), n, P(x), h, None, P(gy), P(dv), P(dx), stream()), "c")))
print("spmm_edge_bwd       %.2f us" % gt(lambda: check(L.dggb_spmm_edge_bwd(P(g.erow), P(g.col), P(v), E, P(x), h, None, P(gy), P(dv), P(dx), stream()), "c")))
print("spmm_gemm_bwd       %.2f us" % gt(lambda: check(L.dggb_spmm_gemm_bwd(P(g.rowptr), P(g.col), P(v), n, P(x), h, None, 0.9, P(w), h, 0.4, 0.6, P(gy), P(dv), P(dx), P(ds), 0.1, None, 0, None, None, None, 0, stream()), "d")))
print("torch mm            %.2f us" % gt(lambda: torch.mm(x, w, out=y)))
print("torch add           %.2f us" % gt(lambda: torch.add(x, gy, out=y)))
