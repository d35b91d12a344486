"""Micro-benchmark of the all-pairs kernel: which part bounds it? (noise mode x precision x kc)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgg_b200 import functional as K

n = int(sys.argv[1]) if len(sys.argv) > 1 else 37888   # 296 row blocks = 2 waves of 148
d = 64
dev = "cuda"
z = torch.softmax(torch.randn(n, d, device=dev), -1)
t = torch.tensor([4.0], device=dev)
G = None
def run(label, **kw):
    for _ in range(2): K.allpairs_topk(z, t, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): K.allpairs_topk(z, t, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{label:45s} {ms:8.2f} ms  {n*n/ms/1e6:8.1f} Gpairs/s", flush=True)
run("no noise, 3xTF32, kc=32", noise=None, kc=32, precision=3)
run("no noise, 1xTF32, kc=32", noise=None, kc=32, precision=1)
run("philox,   3xTF32, kc=32", noise=None, kc=32, precision=3, seed=1, noise_scale=1.0)
run("philox,   1xTF32, kc=32", noise=None, kc=32, precision=1, seed=1, noise_scale=1.0)
run("philox,   3xTF32, kc=8", noise=None, kc=8, precision=3, seed=1, noise_scale=1.0)
run("philox,   3xTF32, kc=64", noise=None, kc=64, precision=3, seed=1, noise_scale=1.0)
run("philox small scale, 3xTF32, kc=32", noise=None, kc=32, precision=3, seed=1, noise_scale=0.01)
