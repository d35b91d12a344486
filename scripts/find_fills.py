"""Which Python lines launch the small fill kernels inside one DGG fwd+bwd step (torch.profiler with stacks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, dgm
from torch.profiler import profile, ProfilerActivity
from dgg_b200 import CSRGraph
shape = bench.PUBMED; dev = torch.device("cuda")
m = dgm.DGG(in_dim=shape["f"], latent_dim=shape["h"], args=bench.dgg_args()); m.load_state_dict(bench.ref_state(shape)); m = m.to(dev)
hs = bench.make_set(shape, 0)
adj = torch.sparse_coo_tensor(hs["idx"].to(dev), hs["val"].to(dev), (shape["n"],)*2, is_coalesced=True)
x, gv, gx = hs["x"].to(dev), hs["g_vals"].to(dev), hs["g_xenc"].to(dev)
def step():
    for p in m.parameters(): p.grad = None
    out, xe = m(x, adj)
    torch.autograd.backward([out._dgg_vals, xe], [gv, gx])
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages(group_by_stack_n=8).table(sort_by="self_cuda_time_total", row_limit=40, max_name_column_width=40, max_src_column_width=90))
