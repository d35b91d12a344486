"""A/B of the conv-layer routing: train-step graph replay of GCN_DGG_00 (Pubmed), GCN_DGG (Cora), GCNII_DGG-64 (Citeseer)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, dgg_b200
dev = torch.device("cuda")
shape = bench.PUBMED
dsets = []
for s in [bench.make_set(shape, i) for i in range(6)]:
    adj = torch.sparse_coo_tensor(s["idx"].to(dev), s["val"].to(dev), (shape["n"], shape["n"]), is_coalesced=True)
    dgg_b200.CSRGraph.from_coo(adj)
    dsets.append(dict(adj=adj, x=s["x"].to(dev)))
e = bench.full_model_epoch(dsets, shape, dev)
print("GCN_DGG_00 pubmed  graph %.4f ms  eager %.3f" % (e["train_step_graph_ms"], e["train_step_ms"]))
for r in bench.config_epochs(dev, dsets):
    print(r.get("model"), r.get("shape"), "graph", r.get("train_step_graph_ms"), "eager", r.get("train_step_ms"), r.get("error"))
