import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, dgm
import torch.nn.functional as F
from dgg_b200 import CSRGraph, functional as K
shape = bench.PUBMED; dev = torch.device("cuda")
m = dgm.DGG(in_dim=shape["f"], latent_dim=shape["h"], args=bench.dgg_args()); m.load_state_dict(bench.ref_state(shape)); m = m.to(dev)
lin, dd = m.edge_encoder[0], m.degree_decoder[0]
for s in range(6):
    hs = bench.make_set(shape, s)
    adj = torch.sparse_coo_tensor(hs["idx"].to(dev), hs["val"].to(dev), (shape["n"],)*2, is_coalesced=True)
    g, _ = CSRGraph.from_coo(adj)
    deg = (g.rowptr[1:] - g.rowptr[:-1])
    with torch.no_grad():
        y = F.linear(m.node_encoder(hs["x"].to(dev)), lin.weight)
        for _ in range(3): out = K._DGGEdge.apply(y, lin.bias, dd.weight, dd.bias, g, None, -1)
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): out = K._DGGEdge.apply(y, lin.bias, dd.weight, dd.bias, g, None, -1)
        e1.record(); torch.cuda.synchronize()
    print(s, "E", g.nnz, "max deg", int(deg.max()), "fwd us", e0.elapsed_time(e1)/10*1e3, flush=True)
