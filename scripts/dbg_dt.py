import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgg_b200 import functional as K
from tests.philox_ref import gumbel_matrix
n, d, kc, seed = 700, 32, 32, 11
gen = torch.Generator().manual_seed(0)
z = torch.softmax(torch.randn(n, d, generator=gen), -1)
G = gumbel_matrix(n, n, seed, 1.0)
gy = torch.randn(n, kc, generator=gen) * 1e-3
for mode in ("philox", "injected"):
    zc = z.cuda().requires_grad_(True); tc = torch.tensor([3.0], device="cuda", requires_grad=True)
    if mode == "philox":
        idx, y = K.allpairs_topk(zc, tc, None, kc, 3, seed=seed, noise_scale=1.0)
    else:
        idx, y = K.allpairs_topk(zc, tc, G.cuda(), kc, 3)
    (y * gy.cuda()).sum().backward()
    # reference through torch ops on the SAME selected pairs
    z2 = z.clone().double().requires_grad_(True); t2 = torch.tensor([3.0], dtype=torch.double, requires_grad=True)
    ii = idx.cpu().long()
    rows = torch.arange(n).reshape(-1, 1).expand(n, kc)
    D = (z2[rows] - z2[ii]).norm(dim=-1)
    y2 = -t2 * D + G.double()[rows, ii]
    (y2 * gy.double()).sum().backward()
    print(mode, "y max abs", float((y.detach().cpu().double() - y2.detach()).abs().max()),
          "dt kernel", float(tc.grad), "dt ref", float(t2.grad),
          "dz max abs", float((zc.grad.cpu().double() - z2.grad).abs().max()), "dz scale", float(z2.grad.abs().max()))
