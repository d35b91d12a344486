"""One all-pairs launch at the Reddit column count (232 965 columns, one wave of 148 row blocks) for an ncu capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgg_b200 import functional as K
n, rows = 232965, 18944
g = torch.Generator(device="cuda").manual_seed(0)
z = torch.randn(n, 64, device="cuda", generator=g) * 0.577
t = torch.tensor([1.0], device="cuda")
for _ in range(3):
    K.allpairs_topk(z, t, None, 32, 3, 0, rows, seed=1, noise_scale=1.0)
torch.cuda.synchronize()
