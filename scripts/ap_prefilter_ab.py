"""A/B of the Philox candidate pre-filter of the all-pairs kernel (DGGB_AP_NO_PREFILTER=1: one-step scoring)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgg_b200 import functional as K

dev = "cuda"
def bench(label, z, t, n_rows, **kw):
    n = z.shape[0]
    res = {}
    for mode in ("prefilter", "one-step"):
        if mode == "one-step": os.environ["DGGB_AP_NO_PREFILTER"] = "1"
        else: os.environ.pop("DGGB_AP_NO_PREFILTER", None)
        for _ in range(2): out = K.allpairs_topk(z, t, None, 32, 3, 0, n_rows, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): out = K.allpairs_topk(z, t, None, 32, 3, 0, n_rows, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        res[mode] = (ms, out)
        print(f"{label:40s} {mode:10s} {ms:8.2f} ms  {n_rows*n/ms/1e6:8.1f} Gpairs/s", flush=True)
    a, b = res["prefilter"][1], res["one-step"][1]
    print("   identical:", bool(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])))

g = torch.Generator(device=dev).manual_seed(0)
n = 37888
z = torch.softmax(torch.randn(n, 64, device=dev, generator=g), -1)
bench("softmax z, t=4 (noise-dominated) N=37888", z, torch.tensor([4.0], device=dev), n, seed=1, noise_scale=1.0)
n = 232965
z = torch.randn(n, 64, device=dev, generator=g) * 0.577
bench("randn z (D~6.5), t=1, N=232965, 37888 rows", z, torch.tensor([1.0], device=dev), 37888, seed=1, noise_scale=1.0)
bench("same, t=30 (distance-dominated)", z, torch.tensor([30.0], device=dev), 37888, seed=1, noise_scale=1.0)

# column parts at the row count of one of 8 / 4 ranks (Reddit shape)
def parts_bench(rows):
    for parts in ("1", "auto"):
        if parts == "auto": os.environ.pop("DGGB_AP_PARTS", None)
        else: os.environ["DGGB_AP_PARTS"] = parts
        for _ in range(2): out = K.allpairs_topk(z, torch.tensor([1.0], device=dev), None, 32, 3, 1000, rows, seed=1, noise_scale=1.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): out = K.allpairs_topk(z, torch.tensor([1.0], device=dev), None, 32, 3, 1000, rows, seed=1, noise_scale=1.0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"rows {rows:7d}  parts {parts:5s} {ms:8.2f} ms  {rows*n/ms/1e6:8.1f} Gpairs/s", flush=True)
parts_bench(29121); parts_bench(9472); parts_bench(4000)
