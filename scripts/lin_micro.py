import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgg_b200 import functional as K
n, f, h = 19717, 500, 64
xs = [torch.rand(n, f, device="cuda") for _ in range(6)]
w = torch.randn(h, f, device="cuda") / 20; b = torch.randn(h, device="cuda")
we = torch.randn(h, h, device="cuda") / 8
d = [torch.randn(n, h, device="cuda") for _ in range(6)]
def timed(fn, it=60):
    for i in range(6): fn(i)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(it): fn(i)
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/it*1e3
print("linear_tc  N x 500 -> 64 : %.1f us" % timed(lambda i: K._linear_act_tc(xs[i%6], w, b, 0.01)))
print("linear_tc  N x 64 -> 64  : %.1f us" % timed(lambda i: K._linear_act_tc(d[i%6], we, None, 1.0)))
print("cublas     N x 500 -> 64 : %.1f us" % timed(lambda i: torch.nn.functional.linear(xs[i%6], w, b)))
print("gemm_tn    dpre^T x       : %.1f us" % timed(lambda i: K.gemm_tn(d[i%6], xs[i%6], True)))
print("gemm_tn    dy^T x_enc     : %.1f us" % timed(lambda i: K.gemm_tn(d[i%6], d[(i+1)%6], False)))
print("copy 39MB (bw ref)        : %.1f us" % timed(lambda i: xs[(i+1)%6].copy_(xs[i%6])))
