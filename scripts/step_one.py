"""A few eager Pubmed-shape DGG fwd+bwd steps (for ncu captures of the individual kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, dgm
from dgg_b200 import CSRGraph
shape = bench.PUBMED; dev = torch.device("cuda")
m = dgm.DGG(in_dim=shape["f"], latent_dim=shape["h"], args=bench.dgg_args()); m.load_state_dict(bench.ref_state(shape)); m = m.to(dev)
sets = []
for s in range(3):
    hs = bench.make_set(shape, s)
    adj = torch.sparse_coo_tensor(hs["idx"].to(dev), hs["val"].to(dev), (shape["n"],)*2, is_coalesced=True)
    CSRGraph.from_coo(adj)
    sets.append((adj, hs["x"].to(dev), hs["g_vals"].to(dev), hs["g_xenc"].to(dev)))
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    adj, x, gv, gx = sets[i % 3]
    for p in m.parameters(): p.grad = None
    out, xe = m(x, adj)
    torch.autograd.backward([out._dgg_vals, xe], [gv, gx])
torch.cuda.synchronize()
