"""Which ops of the Citeseer-shape GCNII_DGG-64 step still land on library GEMMs (eager step, shapes + stacks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.argv = sys.argv[:1]
import torch
from torch.profiler import profile, ProfilerActivity
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gcnii_graph_prof.py")).read().split("g = dgg_b200.GraphedStep(body)")[0]
exec(src)
for _ in range(3): body()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    body(); torch.cuda.synchronize()
for e in prof.events():
    if e.name in ("aten::mm", "aten::addmm", "aten::bmm") and e.input_shapes:
        st = [s for s in (e.stack or []) if "/repo/" in s][:3]
        print(e.name, e.input_shapes, "%.1f us" % (e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total), st)
