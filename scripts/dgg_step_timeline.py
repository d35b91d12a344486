"""Warm timeline (CUPTI start offsets + durations) of the kernels inside one CUDA-graph replay of the Pubmed-shape
DGG fwd+bwd step -- the same capture bench.py times (6 rotating input sets, static flat gradient buffer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, dgm, dgg_b200
from dgg_b200 import CSRGraph
from dgg_b200.sharding import flatten_grads
from torch.profiler import profile, ProfilerActivity
shape = bench.PUBMED; dev = torch.device("cuda")
m = dgm.DGG(in_dim=shape["f"], latent_dim=shape["h"], args=bench.dgg_args()); m.load_state_dict(bench.ref_state(shape)); m = m.to(dev)
params = [p for p in m.parameters()]
sets = []
for s in range(bench.N_SETS):
    hs = bench.make_set(shape, s)
    adj = torch.sparse_coo_tensor(hs["idx"].to(dev), hs["val"].to(dev), (shape["n"],) * 2, is_coalesced=True)
    CSRGraph.from_coo(adj)
    sets.append((adj, hs["x"].to(dev), hs["g_vals"].to(dev), hs["g_xenc"].to(dev)))
def body(s):
    adj, x, gv, gx = s
    for p in params: p.grad = None
    out, xe = m(x, adj)
    torch.autograd.backward([out._dgg_vals, xe], [gv, gx])
    return out._dgg_vals, [p.grad for p in params]
for s in sets[:2]: body(s)
graphs = [dgg_b200.GraphedStep(lambda s=s: body(s)) for s in sets]
for i in range(60): graphs[i % 6]()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(60): graphs[i % 6]()
e1.record(); torch.cuda.synchronize()
print("ms/step %.4f" % (e0.elapsed_time(e1) / 60))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(12): graphs[i % 6]()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if "cuda" in str(e.device_type).lower()], key=lambda e: e.time_range.start)
per = len(evs) // 12
print("kernels per replay", per)
for rep in (6, 7):
    seg = evs[rep * per:(rep + 1) * per]
    t0 = seg[0].time_range.start
    prev_end = t0
    print("replay", rep, "span %.1f us" % (seg[-1].time_range.end - t0))
    for e in seg:
        print("  start %7.1f  gap %6.1f  dur %6.1f  %s" % (e.time_range.start - t0, e.time_range.start - prev_end,
                                                        e.time_range.end - e.time_range.start, e.name[:60]))
        prev_end = max(prev_end, e.time_range.end)
