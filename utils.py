"""Host-side helpers on the path's input side: the per-step pipeline of the reference
training loops is  to_scipy_sparse_matrix -> add_noisy_edges -> sparse_mx_to_torch_sparse_tensor -> .to(device)
(train_small_graphs.py:251-255, 283-289, 305-311), three times per epoch, with a dense N x N numpy RNG draw
inside ``add_noisy_edges`` that is re-seeded to 0 on every call (reference utils.py:92-110) -- i.e. it returns
the SAME noisy graph every time at O(N^2) host cost.  These drop-ins keep the reference's behaviour (same numpy
random stream, same outputs) but memoise the result and never hold more than a row chunk of the random
matrix; ``cached_device_adj`` additionally keeps the device-side sparse tensor (with its CSR handle) alive
across epochs.  Only the helpers that touch the DGG path's inputs live here (utils.py:31-35, 81-110, 1260-1268);
dataset loaders and the other experiment utilities of the reference are out of scope.
"""
from __future__ import annotations

import argparse

import numpy as np
import scipy.sparse as sp
import torch

_NOISY_CACHE = {}
_DEVICE_CACHE = {}


def accuracy(output, labels):
    """Fraction of correct argmax predictions as a double tensor (reference utils.py:31-35)."""
    preds = output.max(1)[1].type_as(labels)
    correct = preds.eq(labels).double().sum()
    return correct / len(labels)


def str2bool(v):
    """argparse boolean (reference utils.py:1260-1268)."""
    if isinstance(v, bool):
        return v
    if v.lower() in ("yes", "true", "t", "y", "1"):
        return True
    if v.lower() in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError("Boolean value expected.")


def sparse_mx_to_torch_sparse_tensor(sparse_mx):
    """scipy sparse -> torch sparse COO float32 (reference utils.py:81-89)."""
    sparse_mx = sparse_mx.tocoo().astype(np.float32)
    indices = torch.from_numpy(np.vstack((sparse_mx.row, sparse_mx.col)).astype(np.int64))
    values = torch.from_numpy(sparse_mx.data)
    return torch.sparse_coo_tensor(indices, values, torch.Size(sparse_mx.shape))


def _graph_key(adj):
    adj = sp.coo_matrix(adj)
    h = hash((adj.shape, adj.nnz, adj.row[:64].tobytes(), adj.col[:64].tobytes(), adj.row[-64:].tobytes(),
              adj.col[-64:].tobytes(), float(adj.data.sum())))
    return h, adj


def add_noisy_edges(adj, noise_level=0.1, chunk_rows=1024):
    """Reference utils.py:92-110: add an edge of weight 1 wherever ``np.random.rand(N, N) < 10 * noise_level``
    (seed 0), except on existing edges and on the diagonal.  Same random stream (numpy fills row-major, so
    drawing the matrix in row chunks after one ``seed(0)`` is bit-identical), O(chunk * N) memory instead of
    three dense N x N arrays, and memoised: the reference recomputes the identical result on every call."""
    key, adj = _graph_key(adj)
    key = (key, float(noise_level))
    hit = _NOISY_CACHE.get(key)
    if hit is not None:
        return hit.copy()
    level = noise_level * 10
    n_rows, n_cols = adj.shape
    state = np.random.get_state()
    np.random.seed(0)
    existing = sp.csr_matrix((np.ones(adj.nnz, dtype=bool), (adj.row, adj.col)), shape=adj.shape)
    rows, cols = [], []
    for r0 in range(0, n_rows, chunk_rows):
        r1 = min(n_rows, r0 + chunk_rows)
        hitmask = np.random.rand(r1 - r0, n_cols) < level
        rr, cc = np.nonzero(hitmask)
        rr = rr + r0
        keep = rr != cc
        if keep.any():
            rr, cc = rr[keep], cc[keep]
            on_edge = np.asarray(existing[rr, cc]).ravel()
            rr, cc = rr[~on_edge], cc[~on_edge]
            rows.append(rr)
            cols.append(cc)
    # the reference leaves the global numpy RNG re-seeded to 0 and advanced by N*N draws; callers that
    # depended on that side effect do not exist on this path, so the caller's stream is restored instead
    np.random.set_state(state)
    if rows:
        rows, cols = np.concatenate(rows), np.concatenate(cols)
    else:
        rows = cols = np.zeros(0, dtype=np.int64)
    noise = sp.coo_matrix((np.ones(len(rows), dtype=np.float64), (rows, cols)), shape=adj.shape)
    noisy = sp.csr_matrix(adj.astype(np.float64) + noise)
    _NOISY_CACHE[key] = noisy
    return noisy.copy()


def cached_device_adj(edge_index, num_nodes, noise_level=0.0, device="cuda"):
    """The whole per-step host pipeline, once: edge_index [2,E] -> (optionally noisy) coalesced sparse COO on
    ``device`` carrying the CSR handle the DGG modules reuse.  Later calls with the same graph return the same
    device tensor (no host work, no H2D copy)."""
    ei = edge_index.detach().cpu().numpy()
    key = (int(num_nodes), ei.shape[1], hash(ei[:, :64].tobytes()), hash(ei[:, -64:].tobytes()), float(noise_level),
           str(device))
    hit = _DEVICE_CACHE.get(key)
    if hit is not None:
        return hit
    adj = sp.coo_matrix((np.ones(ei.shape[1], dtype=np.float32), (ei[0], ei[1])), shape=(num_nodes, num_nodes))
    if noise_level > 0.0:
        adj = add_noisy_edges(adj, noise_level=noise_level)
    t = sparse_mx_to_torch_sparse_tensor(adj).coalesce().to(device)
    if t.is_cuda:
        from dgg_b200 import CSRGraph

        CSRGraph.from_coo(t)       # build + attach the int32 CSR once
    _DEVICE_CACHE[key] = t
    return t
