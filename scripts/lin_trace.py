"""Per-phase clock64 trace of linear_tf32x3_kernel<64,3,FUSE2> (build with DGGB_NVCC_EXTRA=-DDGGB_LIN_TRACE)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgg_b200 import functional as K
from dgg_b200 import _lib as _L
if os.environ.get('DGGB_TRACE_LIB'): _L.LIB_PATH = os.environ['DGGB_TRACE_LIB']
from dgg_b200._lib import lib
n, f, h = 19717, 500, 64
xs = [torch.rand(n, f, device="cuda") for _ in range(6)]
wn = torch.randn(h, f, device="cuda") / 20; b = torch.randn(h, device="cuda"); we = torch.randn(h, h, device="cuda") / 8
for fuse in (False, True):
    for i in range(6):
        if fuse: K._linear_act_tc(xs[i], wn, b, 0.01, w2=we)
        else: K._linear_act_tc(xs[i], wn, b, 0.01)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 320)()
    assert lib().dggb_debug_lin_trace(buf) == 0
    for c in range(2):
        t = list(buf[c * 160:(c + 1) * 160]); t0 = t[0]
        r = lambda i: t[i] - t0
        print(f"fuse={fuse} cta{c}: setup {r(1)} acc_full {r(2)} drained {r(6)} batch0 {r(7)} out_written {r(3)} acc2_full {r(4)} end {r(5)}")
        print("  producer  ", [r(8 + k) for k in range(16)])
        print("  conv land ", [r(72 + k) for k in range(16)])
        print("  conv done ", [r(104 + k) for k in range(16)])
        print("  mma A rdy ", [r(136 + k) for k in range(16)])
        print("  mma W rdy ", [r(24 + k) for k in range(16)])
        print("  mma issued", [r(40 + k) for k in range(16)])

sm = (ctypes.c_int * 1024)()
assert lib().dggb_debug_lin_smid(sm) == 0
ids = list(sm[:155])
import collections
cnt = collections.Counter(ids)
print("distinct SMs", len(cnt), "max CTAs per SM", max(cnt.values()), "first 12 smids", ids[:12])
