import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgg_b200 import functional as K
n = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
mode = sys.argv[2] if len(sys.argv) > 2 else "none"
z = torch.softmax(torch.randn(n, 64, device="cuda"), -1)
t = torch.tensor([4.0], device="cuda")
kw = dict(noise=None, kc=32, precision=3)
if mode == "philox":
    kw.update(seed=1, noise_scale=1.0)
for _ in range(3):
    K.allpairs_topk(z, t, **kw)
torch.cuda.synchronize()
