"""One eager training step of each model family at its BASELINE shape (for ncu captures of the r02 kernels:
spmm / spmm_gemm (GCN_DGG_00, Pubmed), gat (GAT_DGG_00, Pubmed), edge_mlp / row_firstk (GCN_DGG, Cora))."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import bench, dgg_b200, model as models
dev = torch.device("cuda")
shape = bench.PUBMED; n = shape["n"]
hs = bench.make_set(shape, 0)
adj = torch.sparse_coo_tensor(hs["idx"].to(dev), hs["val"].to(dev), (n, n), is_coalesced=True); dgg_b200.CSRGraph.from_coo(adj)
x = hs["x"].to(dev); labels = torch.randint(0, 3, (n,), device=dev); ti = torch.arange(60, device=dev)
a0 = argparse.Namespace(extra_edge_dim=0, dgg_adj_input="input_adj")
torch.manual_seed(0)
gcn = models.GCN_DGG_00(nfeat=shape["f"], nlayers=2, nhidden=64, nclass=3, args=a0).to(dev)
gat = models.GAT_DGG_00(nfeat=shape["f"], nlayers=2, nhidden=64, nclass=3, args=a0).to(dev)
ii = adj.indices(); nl = ii[:, ii[0] != ii[1]].contiguous()
adj_nl = torch.sparse_coo_tensor(nl, torch.ones(nl.shape[1], device=dev), (n, n)).coalesce()
lk = dict(extra_edge_dim=2, extra_k_dim=1, dgg_hard=False, deg_mean=3.899, deg_std=5.288, dgg_mode_edge_net="u-v-deg",
          dgg_mode_k_net="x", dgg_mode_k_select="k_times_edge_prob", debug_step=3, perturb_edge_prob=False,
          symmetric_noise=True, stochastic_k=False, dgg_adj_input="input_adj", n_dgg_layers=2)
nc, fc = 2708, 1433
ic, vc = bench.chung_lu_graph(nc, 3.9, 60, 7); keep = ic[0] != ic[1]
adjc = torch.sparse_coo_tensor(ic[:, keep].to(dev), vc[keep].to(dev), (nc, nc)).coalesce()
xc = bench._sparse_features(nc, fc, 0.013, 8).to(dev); lc = torch.randint(0, 7, (nc,), device=dev)
cora = models.GCN_DGG(nfeat=fc, nlayers=2, nhidden=64, nclass=7, args=argparse.Namespace(**lk)).to(dev)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    for net, call in ((gcn, lambda: gcn(x, adj)[0]), (gat, lambda: gat(x, adj_nl, edge_index=nl)[0])):
        net.train(); net.zero_grad()
        F.nll_loss(call()[ti], labels[ti]).backward()
    cora.train(); cora.zero_grad()
    F.nll_loss(cora(xc, adjc)[0][ti], lc[ti]).backward()
torch.cuda.synchronize()
