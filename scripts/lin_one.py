import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgg_b200 import functional as K
n, f, h = 19717, 500, 64
xs = [torch.rand(n, f, device="cuda") for _ in range(6)]
w = torch.randn(h, f, device="cuda") / 20; b = torch.randn(h, device="cuda"); we = torch.randn(h, h, device="cuda") / 8
d = [torch.randn(n, h, device="cuda") for _ in range(6)]
for i in range(6):
    K._linear_act_tc(xs[i], w, b, 0.01, w2=we)
    K._linear_act_tc(xs[i], w, b, 0.01)
    K.gemm_tn(d[i], xs[i], True)
torch.cuda.synchronize()
