"""Encoder GEMM kernels timed through the C-ABI inside a CUDA graph (no Python / launch overhead in the number):
6 rotating Pubmed-shape inputs (> L2) per replay.  Prints us per call and GB/s by algorithmic bytes."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgg_b200 import functional as K
from dgg_b200._lib import lib, check, p, stream
n, f, h = int(os.environ.get("LIN_N", 19717)), 500, 64
xs = [torch.rand(n, f, device="cuda") for _ in range(6)]
w = torch.randn(h, f, device="cuda") / 20; b = torch.randn(h, device="cuda"); we = torch.randn(h, h, device="cuda") / 8
out = torch.empty(n, h, device="cuda"); out2 = torch.empty(n, h, device="cuda")
ws = torch.empty(2 * h * f + 2 * h * h, device="cuda")
dpre = torch.randn(n, h, device="cuda")
L = lib()
ws_tn_bytes = int(L.dggb_gemm_tn_tc_workspace_bytes(n, h)); ws_tn = torch.empty(ws_tn_bytes // 4, device="cuda")
dw = torch.zeros(h * f + h, device="cuda")
def lin(i, fuse):
    check(L.dggb_linear_fused(p(xs[i % 6]), p(w), 0, p(b), None, None, 0.01, n, f, h, p(out), p(we) if fuse else None,
                              p(out2) if fuse else None, p(ws), ws.numel() * 4, None, 0, None, 0, stream()), "lin")
def tn(i):
    check(L.dggb_gemm_tn_tc(p(dpre), p(xs[i % 6]), n, h, f, p(dw[:h * f]), p(dw[h * f:]), p(ws_tn), ws_tn_bytes, stream()), "tn")
def graph_time(fn, reps=200):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(6): fn(i)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(6): fn(i)
    import time
    t_end = time.perf_counter() + 0.2          # let the SM clocks ramp up before timing
    while time.perf_counter() < t_end: g.replay()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / (reps * 6) * 1e3
bytes_ = n * f * 4 + n * h * 4 + h * f * 4
for name, fn in (("linear (plain)", lambda i: lin(i, False)), ("linear (FUSE2)", lambda i: lin(i, True)), ("gemm_tn_tc", tn),
                 ("copy 39 MB", lambda i: xs[(i + 1) % 6].copy_(xs[i % 6]))):
    t = graph_time(fn)
    print("%-16s %6.2f us   %7.0f GB/s  (%.3f of 6461)" % (name, t, bytes_ / t / 1e3, bytes_ / t / 1e3 / 6461.2))
