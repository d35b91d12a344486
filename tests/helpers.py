import torch


def random_graph(n, avg_deg, seed, self_loops=True, hubs=0, hub_deg=0, device="cpu"):
    """Symmetric random graph (+ optional high-degree hub rows) as coalesced COO indices/values."""
    g = torch.Generator().manual_seed(seed)
    m = n * avg_deg // 2
    src = torch.randint(0, n, (m,), generator=g)
    dst = torch.randint(0, n, (m,), generator=g)
    if hubs:
        hs = torch.randint(0, n, (hubs,), generator=g).repeat_interleave(hub_deg)
        hd = torch.randint(0, n, (hubs * hub_deg,), generator=g)
        src, dst = torch.cat([src, hs]), torch.cat([dst, hd])
    keep = src != dst
    src, dst = src[keep], dst[keep]
    i, j = torch.cat([src, dst]), torch.cat([dst, src])
    if self_loops:
        loops = torch.arange(n)
        i, j = torch.cat([i, loops]), torch.cat([j, loops])
    a = torch.sparse_coo_tensor(torch.stack([i, j]), torch.ones(i.numel()), (n, n)).coalesce()
    idx = a.indices().clone()
    val = torch.ones(idx.shape[1])
    return idx.to(device), val.to(device)


def coo(idx, val, n):
    return torch.sparse_coo_tensor(idx, val, (n, n)).coalesce()


def tie_free_rows(dense_scores, idx, rel_gap=1e-5):
    """Rows whose on-edge scores are pairwise separated by > rel_gap (rank parity is asserted there)."""
    n = dense_scores.shape[0]
    ok = torch.ones(n, dtype=torch.bool)
    srt = torch.sort(dense_scores, dim=-1, descending=True).values
    gap = (srt[:, :-1] - srt[:, 1:]).abs()
    both_pos = (srt[:, :-1] > 0) & (srt[:, 1:] > 0)
    bad = (both_pos & (gap <= rel_gap * srt[:, :-1].abs())).any(-1)
    return ~bad
