"""torch.autograd.Function wrappers over the C-ABI (include/dggb.h).  CUDA tensors only; there is no
CPU path here by design (the CPU oracle lives under oracle/ and is test infrastructure)."""
from __future__ import annotations

import torch

from ._lib import DggbError, check, i32, i64, lib, p, stream
from .graph import CSRGraph


def _f32c(t):
    return t.contiguous() if t.dtype == torch.float32 else t.to(torch.float32).contiguous()


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("dgg_b200 ops need CUDA tensors (no CPU fallback)")


import os as _os

_NO_FUSED_CONV = bool(_os.environ.get("DGGB_NO_FUSED_CONV"))
_NO_EDGE_SPMM = bool(_os.environ.get("DGGB_NO_EDGE_SPMM"))
# The one-launch conv layer (SpMM + W in shared memory + epilogue) pays off where the step is launch-bound: Cora /
# Citeseer sized graphs (GCNII-64 Citeseer train step 5.24 vs 5.38 ms, Cora GCN_DGG 0.551 vs 0.553).  At Pubmed size
# its SIMT dense part (N x 64 x 64 FMAs next to the gathers) loses to the entry-parallel SpMM + a library GEMM:
# GCN_DGG_00 train step 0.337 vs 0.308 ms (scripts/model_ab.py, DGGB_NO_FUSED_CONV A/B).
_FUSED_CONV_MAX_N = int(_os.environ.get("DGGB_FUSED_CONV_MAX_N", "8192"))
_NO_STACK = bool(_os.environ.get("DGGB_NO_STACK"))     # A/B: GCNII layers one launch each instead of the stack kernel
_NO_GRAD_SHARE = bool(_os.environ.get("DGGB_NO_GRAD_SHARE"))   # A/B: autograd sums the per-layer h0 / value gradients
# rows from which the weight-gradient GEMM of a wide encoder (Q >= 256) runs on the tensor-core kernel
_TN_TC_MIN_N = int(_os.environ.get("DGGB_TN_TC_MIN_N", "2048"))
_FUSED_MAX_ROW = 512   # kFusedMaxDeg of csrc/dgg_edge.cu
_LONG_ROW = 1024       # kRankCap of csrc/dgg_edge.cu: longer rows are ranked by a grid-wide launch


def _long_ws(graph: CSRGraph):
    """Zero-headed scratch list for hub rows (None when the graph has none: the row kernels then take no detour)."""
    if graph.max_row_nnz <= _LONG_ROW and graph.max_row_nnz >= 0:
        return None
    return torch.zeros(graph.n + 1, dtype=torch.int32, device=graph.rowptr.device)


class _DGGEdge(torch.autograd.Function):
    """dgm.py:1781-1810 on CSR: scores, degree estimate, in-row ranks, soft first-k."""

    @staticmethod
    def forward(ctx, y, be, deg_w, deg_b, graph: CSRGraph, noise, hard_k: int):
        _require_cuda(y, be, deg_w, deg_b)
        y, be = _f32c(y), _f32c(be)
        deg_w, deg_b = _f32c(deg_w).reshape(-1), _f32c(deg_b).reshape(-1)
        n, h = y.shape
        E = graph.nnz
        dev = y.device
        R = torch.empty(E, dtype=torch.float32, device=dev)
        rank = torch.empty(E, dtype=torch.int32, device=dev)
        s = torch.empty(n, dtype=torch.float32, device=dev)
        k = torch.empty(n, dtype=torch.float32, device=dev)
        out = torch.empty(E, dtype=torch.float32, device=dev)
        mx = graph.max_row_nnz
        fused = 0 < mx <= _FUSED_MAX_ROW and E > 0   # no hub rows: one launch per direction
        zbuf = None
        if fused:
            if any(ctx.needs_input_grad[:4]):
                # the backward's accumulation buffers (dy | dbe | ddeg (+pad) | ds) are cleared by the forward
                # launch: no separate fill kernel between the two fused launches of a training step
                zbuf = torch.empty(n * h + h + 4 + n, dtype=torch.float32, device=dev)
            check(lib().dggb_dgg_edge_fwd_fused(p(graph.rowptr), p(graph.erow), p(graph.col), i32(n), i32(E), i32(mx),
                                                i32(h), p(y), p(be), p(deg_w), p(deg_b), p(noise), i32(hard_k), p(R),
                                                p(rank), p(s), p(k), p(out), p(zbuf),
                                                i64(0 if zbuf is None else zbuf.numel()), stream()),
                  "dgg_edge_fwd_fused")
        else:
            check(lib().dggb_dgg_edge_fwd(p(graph.rowptr), p(graph.erow), p(graph.col), i32(n), i32(E), i32(h), p(y),
                                          p(be), p(deg_w), p(deg_b), p(noise), i32(hard_k), p(R), p(rank), p(s), p(k),
                                          p(out), p(_long_ws(graph)), stream()), "dgg_edge_fwd")
        ctx.graph, ctx.hard_k, ctx.fused_mx = graph, hard_k, (mx if fused else -1)
        ctx.zbuf = zbuf
        ctx.set_materialize_grads(False)   # k / R / rank never carry gradients: no zero tensors for them per step
        ctx.save_for_backward(y, be, deg_w, deg_b, noise, R, rank, s, k)
        ctx.mark_non_differentiable(k, R, rank)
        return out, k, R, rank

    @staticmethod
    def backward(ctx, g_out, _gk, _gR, _grank):
        if g_out is None:
            return (None,) * 7
        y, be, deg_w, deg_b, noise, R, rank, s, k = ctx.saved_tensors
        g = ctx.graph
        n, h = y.shape
        # one zero-filled buffer: dy | dbe | ddeg (+pad) | ds scratch (already cleared by the fused forward launch;
        # a second backward through the same node gets a fresh one)
        zbuf, ctx.zbuf = getattr(ctx, "zbuf", None), None
        if zbuf is None:
            zbuf = torch.zeros(n * h + h + 4 + n, dtype=torch.float32, device=y.device)
        dy = zbuf[:n * h].view(n, h)
        small = zbuf[n * h:]
        dbe, ddeg, ds_ws = small[:h], small[h:h + 2], small[h + 4:]
        if ctx.fused_mx > 0:
            check(lib().dggb_dgg_edge_bwd_fused(p(g.rowptr), p(g.erow), p(g.col), i32(n), i32(g.nnz), i32(ctx.fused_mx),
                                                i32(h), p(y), p(be), p(deg_w), p(deg_b), p(noise), i32(ctx.hard_k),
                                                p(R), p(rank), p(s), p(k), p(_f32c(g_out)), p(ds_ws), p(dy), p(dbe),
                                                p(ddeg), stream()), "dgg_edge_bwd_fused")
        else:
            check(lib().dggb_dgg_edge_bwd(p(g.rowptr), p(g.erow), p(g.col), i32(n), i32(g.nnz), i32(h), p(y), p(be),
                                          p(deg_w), p(deg_b), p(noise), i32(ctx.hard_k), p(R), p(rank), p(s), p(k),
                                          p(_f32c(g_out)), p(ds_ws), p(dy), p(dbe), p(ddeg), stream()), "dgg_edge_bwd")
        return dy, dbe, ddeg[0:1].reshape(1, 1), ddeg[1:2], None, None, None


def dgg_edge(y, be, deg_w, deg_b, graph, noise=None, hard_k=-1):
    """-> (out_vals [E], k [N], R [E], rank [E] int32)"""
    return _DGGEdge.apply(y, be, deg_w, deg_b, graph, noise, hard_k)


class _SymNormalize(torch.autograd.Function):
    """normalize_adj (model.py:1215-1218) on CSR values."""

    @staticmethod
    def forward(ctx, vals, graph: CSRGraph):
        _require_cuda(vals)
        vals = _f32c(vals)
        dinv = torch.empty(graph.n, dtype=torch.float32, device=vals.device)
        out = torch.empty_like(vals)
        check(lib().dggb_sym_normalize_fwd(p(graph.rowptr), p(graph.col), p(vals), i32(graph.n), p(dinv), p(out),
                                           stream()), "sym_normalize_fwd")
        ctx.graph = graph
        ctx.save_for_backward(vals, dinv)
        return out

    @staticmethod
    def backward(ctx, g):
        vals, dinv = ctx.saved_tensors
        gr = ctx.graph
        t_ws = torch.zeros(gr.n, dtype=torch.float32, device=vals.device)
        dval = torch.empty_like(vals)
        check(lib().dggb_sym_normalize_bwd(p(gr.rowptr), p(gr.col), p(vals), i32(gr.n), p(dinv), p(_f32c(g)),
                                           p(t_ws), p(dval), stream()), "sym_normalize_bwd")
        return dval, None


def sym_normalize(vals, graph):
    return _SymNormalize.apply(vals, graph)


class _Spmm(torch.autograd.Function):
    """Y = A X (optionally Y_i *= row_scale_i) on CSR (model.py:594, 67)."""

    @staticmethod
    def forward(ctx, vals, x, graph: CSRGraph, row_scale):
        _require_cuda(vals, x)
        vals, x = _f32c(vals), _f32c(x)
        n, f = graph.n, x.shape[1]
        # short rows (citation graphs, learned top-K adjacencies): entry-parallel kernels in both directions
        # (Pubmed shape, F = 64, inside a CUDA graph: forward 9.7 us incl. the zero fill vs 15.2 warp-per-row,
        # backward 7.7 vs 21.7 us)
        ctx.edge = (not _NO_EDGE_SPMM) and f % 4 == 0 and f <= 512 and graph.nnz > 0 and graph.nnz <= 64 * n
        if ctx.edge:
            y = torch.zeros(n, f, dtype=torch.float32, device=x.device)
            check(lib().dggb_spmm_edge_fwd(p(graph.rowptr), p(graph.erow), p(graph.col), p(vals), i32(n),
                                           i64(graph.nnz), p(x), i32(f), p(row_scale), p(y), stream()), "spmm_edge_fwd")
        else:
            y = torch.empty(n, f, dtype=torch.float32, device=x.device)
            check(lib().dggb_spmm_csr_fwd(p(graph.rowptr), p(graph.col), p(vals), i32(n), p(x), i32(f), p(row_scale),
                                          p(y), stream()), "spmm_csr_fwd")
        ctx.graph = graph
        ctx.save_for_backward(vals, x, row_scale)
        return y

    @staticmethod
    def backward(ctx, gy):
        vals, x, row_scale = ctx.saved_tensors
        g = ctx.graph
        need_v, need_x = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dval = torch.empty_like(vals) if need_v else None
        dx = torch.zeros_like(x) if need_x else None
        if ctx.edge:
            check(lib().dggb_spmm_edge_bwd(p(g.erow), p(g.col), p(vals), i64(g.nnz), p(x), i32(x.shape[1]),
                                           p(row_scale), p(_f32c(gy)), p(dval), p(dx), stream()), "spmm_edge_bwd")
        else:
            check(lib().dggb_spmm_csr_bwd(p(g.rowptr), p(g.col), p(vals), i32(g.n), p(x), i32(x.shape[1]),
                                          p(row_scale), p(_f32c(gy)), p(dval), p(dx), stream()), "spmm_csr_bwd")
        return dval, dx, None, None


def spmm(vals, x, graph, row_scale=None):
    return _Spmm.apply(vals, x, graph, row_scale)


class GradShare:
    """One gradient accumulator shared by the layers of a stack that consume the SAME tensor (GCNII: h0 in all 64
    layers, the normalised adjacency values in all layers behind the last DGG layer).  Autograd would add the per-layer
    gradients with one launch per layer; here the layer that runs its backward FIRST (the last one of the group in
    forward order) allocates and overwrites the buffer, the others accumulate in place inside their backward launch,
    and only the group's first layer -- whose backward runs last -- hands the buffer to autograd.  Valid because the
    layers form a chain (each consumes its predecessor's output), so their backwards run in reverse forward order."""

    def __init__(self):
        self.buf = None


class _SpmmGemm(torch.autograd.Function):
    """y = act(theta * (s W) + beta * s + resid), s = c1 * rs * (A x) + c2 * h0 in one launch per direction
    (GCNConv model.py:594-598; GraphConvolution model.py:32-44, 65-77); see include/dggb.h."""

    @staticmethod
    def forward(ctx, vals, x, w, h0, resid, graph: CSRGraph, row_scale, c1, c2, theta, beta, relu, out_keep,
                h0_share, val_share):
        _require_cuda(vals, x, w, h0, resid, out_keep)
        vals, x, w = _f32c(vals), _f32c(x), _f32c(w)
        out_keep = None if out_keep is None else _f32c(out_keep)
        h0 = None if h0 is None else _f32c(h0)
        resid = None if resid is None else _f32c(resid)
        n, fin, fout = graph.n, x.shape[1], w.shape[1]
        y = torch.empty(n, fout, dtype=torch.float32, device=x.device)
        need_s = ctx.needs_input_grad[2]
        s = torch.empty(n, fin, dtype=torch.float32, device=x.device) if need_s else None
        check(lib().dggb_spmm_gemm_fwd(p(graph.rowptr), p(graph.col), p(vals), i32(n), p(x), i32(fin), p(row_scale),
                                       p(h0), float(c1), float(c2 if h0 is not None else 0.0), p(w), i32(fout),
                                       float(theta), float(beta), p(resid), i32(1 if relu else 0), p(out_keep), p(y),
                                       p(s), stream()), "spmm_gemm_fwd")
        ctx.graph, ctx.meta = graph, (c1, c2, theta, beta, relu, h0 is not None, resid is not None)
        ctx.h0_share, ctx.val_share = h0_share, val_share      # (GradShare, first in group, last in group) or None
        ctx.save_for_backward(vals, x, w, row_scale, s, y if relu else None, out_keep)
        return y

    @staticmethod
    def backward(ctx, gy):
        vals, x, w, row_scale, s, y, out_keep = ctx.saved_tensors
        c1, c2, theta, beta, relu, has_h0, has_resid = ctx.meta
        g = ctx.graph
        gy = _f32c(gy)
        need_v, need_x, need_w, need_h0 = ctx.needs_input_grad[:4]
        accumulate = 0

        def shared(share, like, bit):
            # the group's LAST layer runs its backward first: fresh buffer, overwritten; the others add in place
            nonlocal accumulate
            sh, _first, last = share
            if last:
                sh.buf = torch.empty_like(like)
            else:
                accumulate |= bit
            return sh.buf

        dval = ds = None
        if need_v:
            dval = shared(ctx.val_share, vals, 1) if ctx.val_share is not None else torch.empty_like(vals)
        dx = torch.zeros_like(x) if need_x else None
        if need_h0 and has_h0:                                               # receives c2 * ds = d h0
            ds = shared(ctx.h0_share, x, 2) if ctx.h0_share is not None else torch.empty_like(x)
        launch = need_v or need_x or ds is not None
        # the weight gradient's split-K accumulator is cleared by the layer's own backward launch (no fill kernel)
        dwbuf = (torch.empty(w.numel(), dtype=torch.float32, device=x.device)
                 if (need_w and launch and w.shape[1] % 4 == 0) else None)
        if launch:
            # the ReLU backward runs inside the layer's launch (relu_y = forward output); the masked gradient comes
            # back for dW and for a residual branch
            fold = relu or out_keep is not None
            gm = torch.empty_like(gy) if (fold and (need_w or has_resid)) else None
            check(lib().dggb_spmm_gemm_bwd(p(g.rowptr), p(g.col), p(vals), i32(g.n), p(x), i32(x.shape[1]),
                                           p(row_scale), float(c1), p(w), i32(w.shape[1]), float(theta), float(beta),
                                           p(gy), p(dval), p(dx), p(ds), float(c2), p(dwbuf),
                                           i64(0 if dwbuf is None else dwbuf.numel()), p(y if relu else None), p(gm),
                                           p(out_keep), i32(accumulate), stream()), "spmm_gemm_bwd")
            if gm is not None:
                gy = gm
        else:
            if out_keep is not None:
                gy = gy * out_keep
            if relu:
                gy = torch.ops.aten.threshold_backward(gy, y, 0.0)
        dw = None
        if need_w:
            dw = gemm_tn(s, gy, False, zeroed=dwbuf)[0]       # s was saved as theta * s
        # a shared accumulator reaches autograd once: through the group's first layer, whose backward runs last
        if ctx.val_share is not None and dval is not None and not ctx.val_share[1]:
            dval = None
        dh0 = ds
        if ctx.h0_share is not None and ds is not None and not ctx.h0_share[1]:
            dh0 = None
        dres = gy if (has_resid and ctx.needs_input_grad[4]) else None
        return dval, dx, dw, dh0, dres, None, None, None, None, None, None, None, None, None, None


class _StackUnsupported(Exception):
    pass


class _GcniiStack(torch.autograd.Function):
    """A run of GCNII layers with one adjacency and one h0 (model.py:722-729): forward = ONE cooperative launch
    (dggb_gcnii_stack_fwd), backward = the per-layer launches of ``_SpmmGemm`` in a loop, with the h0 / adjacency-value
    gradients accumulated in place."""

    @staticmethod
    def forward(ctx, vals, x0, h0, keep, graph: CSRGraph, c1, c2, thetas, *ws):
        import ctypes

        _require_cuda(vals, x0, h0, keep, *ws)
        vals, x0, h0 = _f32c(vals), _f32c(x0), _f32c(h0)
        keep = None if keep is None else _f32c(keep)
        ws = [_f32c(w) for w in ws]
        n, f = x0.shape
        nl = len(ws)
        y = torch.empty(nl, n, f, dtype=torch.float32, device=x0.device)
        s = torch.empty(nl, n, f, dtype=torch.float32, device=x0.device)
        bar = torch.zeros(1, dtype=torch.int32, device=x0.device)
        wp = (ctypes.c_void_p * nl)(*[w.data_ptr() for w in ws])
        th = (ctypes.c_float * nl)(*[float(t) for t in thetas])
        rc = lib().dggb_gcnii_stack_fwd(p(graph.rowptr), p(graph.col), p(vals), i32(n), p(x0), p(h0), i32(f), i32(nl),
                                        ctypes.cast(wp, ctypes.c_void_p), ctypes.cast(th, ctypes.c_void_p), float(c1),
                                        float(c2), p(keep), p(y), p(s), p(bar), stream())
        if rc == -3:            # DGGB_ERR_UNSUPPORTED: the grid cannot be resident at once / shape outside the range
            raise _StackUnsupported()
        check(rc, "gcnii_stack_fwd")
        ctx.graph, ctx.meta = graph, (float(c1), float(c2), [float(t) for t in thetas])
        ctx.save_for_backward(vals, x0, keep, y, s, *ws)
        return y[nl - 1]

    @staticmethod
    def backward(ctx, g_out):
        vals, x0, keep, y, s, *ws = ctx.saved_tensors
        c1, c2, thetas = ctx.meta
        g = ctx.graph
        nl, n, f = y.shape
        need_v, need_x, need_h0 = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        dval = torch.empty_like(vals) if need_v else None
        dh0 = torch.empty_like(x0) if need_h0 else None
        dws = [None] * nl
        gy = _f32c(g_out)
        for k in range(nl - 1, -1, -1):
            xk = x0 if k == 0 else y[k - 1]
            need_dx = need_x or k > 0
            dx = torch.zeros_like(x0) if need_dx else None
            need_w = ctx.needs_input_grad[8 + k]
            dwbuf = torch.empty(f * f, dtype=torch.float32, device=x0.device) if need_w else None
            gm = torch.empty_like(gy) if need_w else None
            check(lib().dggb_spmm_gemm_bwd(p(g.rowptr), p(g.col), p(vals), i32(n), p(xk), i32(f), None, float(c1),
                                           p(ws[k]), i32(f), float(thetas[k]), float(1.0 - thetas[k]), p(gy), p(dval),
                                           p(dx), p(dh0), float(c2), p(dwbuf), i64(0 if dwbuf is None else f * f),
                                           p(y[k]), p(gm), p(None if keep is None else keep[k]),
                                           i32(0 if k == nl - 1 else 3), stream()), "spmm_gemm_bwd")
            if need_w:
                dws[k] = gemm_tn(s[k], gm, False, zeroed=dwbuf)[0]
            gy = dx
        return (dval, gy if need_x else None, dh0, None, None, None, None, None, *dws)


_STACK_MAX_ROWS = {}


def gcnii_stack_applies(x0, ws):
    """Whether ``gcnii_stack`` takes this run of layers: CUDA, square [f, f] weights with f in {32, 64, 128}, 2..96
    layers, and a graph small enough for a fully resident grid (one warp per row)."""
    n, f = x0.shape
    if not (x0.is_cuda and f in (32, 64, 128) and 2 <= len(ws) <= 96 and not _NO_FUSED_CONV and not _NO_STACK
            and all(tuple(w.shape) == (f, f) for w in ws)):
        return False
    if f not in _STACK_MAX_ROWS:
        _STACK_MAX_ROWS[f] = int(lib().dggb_gcnii_stack_max_rows(i32(f)))
    return n <= _STACK_MAX_ROWS[f]


def gcnii_stack(vals, x0, h0, ws, graph, c1, c2, thetas, keep=None):
    """-> the LAST layer's output of a run of GCNII layers sharing (graph, vals) and h0 (``gcnii_stack_applies`` says
    beforehand whether the cooperative one-launch forward takes it)."""
    if not gcnii_stack_applies(x0, ws):
        raise DggbError("gcnii_stack: shape / residency outside the stack kernel's range (see gcnii_stack_applies)")
    try:
        return _GcniiStack.apply(vals, x0, h0, keep, graph, float(c1), float(c2), tuple(float(t) for t in thetas), *ws)
    except _StackUnsupported:
        raise DggbError("gcnii_stack: the grid is not resident on this device") from None


def spmm_gemm_applies(x, w, n, beta=0.0):
    """Whether ``spmm_gemm`` takes this layer shape (callers that set up gradient sharing over a stack need to know
    beforehand: every layer of a group has to go the same way)."""
    fin, fout = x.shape[1], w.shape[1]
    if not (x.is_cuda and fin % 4 == 0 and fin <= 128 and fout <= 128 and (beta == 0.0 or fin == fout)):
        return False
    if _NO_FUSED_CONV:      # A/B measurements: DGGB_NO_FUSED_CONV=1 selects SpMM + library GEMM + elementwise ops
        return False
    return n < _FUSED_CONV_MAX_N


def spmm_gemm(vals, x, w, graph, h0=None, resid=None, row_scale=None, c1=1.0, c2=0.0, theta=1.0, beta=0.0, relu=False,
              out_keep=None, h0_share=None, val_share=None):
    """act(theta * (s W) + beta * s + resid) * out_keep with s = c1 * rs * (A x) + c2 * h0; None if the shape is outside
    the fused kernel's range (Fin % 4 != 0, Fin or Fout > 128).  out_keep: optional dropout multipliers
    (0 or 1 / (1 - p)) of the layer OUTPUT (the dropout in front of the next layer), applied in the epilogue.
    h0_share / val_share: (GradShare, first_in_group, last_in_group) when this call is one of a chain of layers that
    share h0 resp. vals (see GradShare); every layer of the group must pass the same object."""
    if not spmm_gemm_applies(x, w, graph.n, beta):
        return None
    return _SpmmGemm.apply(vals, x, w, h0, resid, graph, row_scale, float(c1), float(c2), float(theta), float(beta),
                           bool(relu), out_keep, h0_share, val_share)


class _AllPairsTopK(torch.autograd.Function):
    """y_ij = -t |z_i - z_j| [+ noise], top-Kc per row, sorted descending (dgm.py:275-301 without N x N)."""

    @staticmethod
    def forward(ctx, z, t, noise, kc: int, precision: int, row_begin: int, row_count: int, seed: int,
                noise_scale: float, inv_temp: float = 0.0, after_val=None, after_idx=None):
        _require_cuda(z, t, noise)
        z, t = _f32c(z), _f32c(t).reshape(-1)
        n, d = z.shape
        L = lib()
        ws_bytes = int(L.dggb_allpairs_workspace_bytes_rows(i32(n), i32(d), i32(row_count), i32(kc)))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=z.device)
        idx = torch.empty(row_count, kc, dtype=torch.int32, device=z.device)
        val = torch.empty(row_count, kc, dtype=torch.float32, device=z.device)
        rowsum = torch.empty(row_count, dtype=torch.float32, device=z.device) if inv_temp != 0.0 else None
        if noise is not None:
            noise = _f32c(noise)
            assert noise.dim() == 2 and noise.shape[0] == row_count and noise.shape[1] >= n
        if after_val is not None:
            after_val, after_idx = _f32c(after_val), after_idx.to(torch.int32).contiguous()
            assert after_val.numel() == row_count and after_idx.numel() == row_count
        check(L.dggb_allpairs_topk_after_fwd(p(z), i32(n), i32(d), i32(row_begin), i32(row_count), p(t), p(noise),
                                             i64(0 if noise is None else noise.stride(0)), int(seed),
                                             float(noise_scale), i32(kc), i32(precision), p(ws), i64(ws_bytes),
                                             p(after_val), p(after_idx), p(idx), p(val), float(inv_temp), p(rowsum),
                                             stream()), "allpairs_topk_fwd")
        ctx.meta = (kc, row_begin, row_count)
        ctx.save_for_backward(z, t, idx)
        ctx.mark_non_differentiable(idx)
        if rowsum is not None:
            ctx.mark_non_differentiable(rowsum)
            return idx, val, rowsum
        return idx, val

    @staticmethod
    def backward(ctx, _gidx, gy, *_unused):
        z, t, idx = ctx.saved_tensors
        kc, row_begin, row_count = ctx.meta
        n, d = z.shape
        dz = torch.zeros_like(z)
        dt = torch.zeros(1, dtype=torch.float32, device=z.device)
        check(lib().dggb_allpairs_pair_bwd(p(z), i32(n), i32(d), i32(row_begin), i32(row_count), p(idx),
                                           p(_f32c(gy)), i32(kc), p(t), p(dz), p(dt), stream()), "allpairs_pair_bwd")
        return dz, dt, None, None, None, None, None, None, None, None, None, None


def allpairs_topk(z, t, noise=None, kc=32, precision=3, row_begin=0, row_count=None, seed=0, noise_scale=0.0,
                  inv_temp=0.0, after=None):
    """-> (idx int32 [rows,kc], y fp32 [rows,kc]) sorted descending per row; differentiable in z and t.
    noise: injected [rows, n] tensor, or None with noise_scale != 0 for in-kernel Philox Gumbel noise.
    inv_temp != 0 additionally returns rowsum [rows] = sum_j exp(y_ij * inv_temp) over all columns.
    after = (val [rows], idx [rows]): continuation pass, only entries sorting strictly after that one per row."""
    if row_count is None:
        row_count = z.shape[0] - row_begin
    av, ai = (None, None) if after is None else after
    return _AllPairsTopK.apply(z, t, noise, int(kc), int(precision), int(row_begin), int(row_count), int(seed),
                               float(noise_scale), float(inv_temp), av, ai)


class _RowFirstK(torch.autograd.Function):
    """select_top_k(mode="k_times_edge_prob") on CSR rows (dgm.py:1402-1421)."""

    @staticmethod
    def forward(ctx, score, k, graph: CSRGraph, mode: int):
        _require_cuda(score, k)
        score, k = _f32c(score), _f32c(k).reshape(-1)
        rank = torch.empty(graph.nnz, dtype=torch.int32, device=score.device)
        out = torch.empty_like(score)
        check(lib().dggb_row_firstk_fwd(p(graph.rowptr), i32(graph.n), p(score), p(k), i32(mode), p(rank), p(out),
                                        p(_long_ws(graph)), stream()), "row_firstk_fwd")
        ctx.graph, ctx.mode = graph, mode
        ctx.save_for_backward(score, k, rank)
        ctx.mark_non_differentiable(rank)
        return out, rank

    @staticmethod
    def backward(ctx, g, _grank):
        score, k, rank = ctx.saved_tensors
        gr = ctx.graph
        dscore = torch.empty_like(score)
        dk = torch.empty_like(k)
        check(lib().dggb_row_firstk_bwd(p(gr.rowptr), i32(gr.n), p(score), p(k), i32(ctx.mode), p(rank), p(_f32c(g)),
                                        p(dscore), p(dk), stream()), "row_firstk_bwd")
        return (dscore if ctx.mode == 0 else None), dk, None, None


def row_firstk(score, k, graph, k_only=False, return_rank=False):
    """out_e = score_e * fk_e, fk_e = 1 - 0.5 (1 + tanh(rank_e - k_row));  k_only: out_e = fk_e (dgm.py:1423-1435)."""
    out, rank = _RowFirstK.apply(score, k, graph, 1 if k_only else 0)
    return (out, rank) if return_rank else out


EX_VAL, EX_DEG, EX_DIST, DIST_ONLY = 1, 2, 4, 8


class _EdgeMLP(torch.autograd.Function):
    """edge_prob_net of DGG_LearnableK_debug on CSR entries (dgm.py:1596-1727); see include/dggb.h."""

    @staticmethod
    def forward(ctx, p_uv, xe, wx, b1, w2, b2, graph: CSRGraph, edge_val, deg, flags: int, slope: float,
                dist_scale: float):
        dist_only = bool(flags & DIST_ONLY)
        _require_cuda(p_uv, xe, edge_val, deg)
        ctx.shapes = tuple(None if t is None else t.shape for t in (wx, b1, w2, b2))
        if dist_only:
            xe = _f32c(xe)
            w, ldp = 0, 0
        else:
            p_uv, b1, w2, b2 = _f32c(p_uv), _f32c(b1).reshape(-1), _f32c(w2).reshape(-1), _f32c(b2).reshape(-1)
            w, ldp = b1.numel(), p_uv.shape[1]
            wx = None if wx is None or wx.numel() == 0 else _f32c(wx)
            xe = None if xe is None else _f32c(xe)
        edge_val = None if edge_val is None else _f32c(edge_val)
        deg = None if deg is None else _f32c(deg).reshape(-1)
        hx = 0 if xe is None else xe.shape[1]
        dev = (xe if dist_only else p_uv).device
        score = torch.empty(graph.nnz, dtype=torch.float32, device=dev)
        check(lib().dggb_edge_mlp_fwd(p(graph.erow), p(graph.col), i32(graph.nnz), i32(w), i32(ldp), p(p_uv), p(xe),
                                      i32(hx), p(edge_val), p(deg), p(wx), p(b1), p(w2), p(b2), float(slope),
                                      float(dist_scale), i32(flags), p(score), stream()), "edge_mlp_fwd")
        ctx.graph, ctx.meta = graph, (flags, slope, dist_scale, w, ldp, hx)
        ctx.save_for_backward(p_uv, xe, wx, b1, w2, b2, edge_val, deg, score)
        return score

    @staticmethod
    def backward(ctx, g):
        p_uv, xe, wx, b1, w2, b2, edge_val, deg, score = ctx.saved_tensors
        flags, slope, dist_scale, w, ldp, hx = ctx.meta
        gr = ctx.graph
        dist_only = bool(flags & DIST_ONLY)
        dev = score.device
        need_xe = xe is not None and ctx.needs_input_grad[1]
        d_xe = torch.zeros_like(xe) if need_xe else None
        if dist_only:
            d_p = d_wx = d_b1 = d_w2 = d_b2 = None
        else:
            m = 0 if wx is None else wx.shape[1]
            d_p = torch.zeros_like(p_uv)
            small = torch.zeros(w * (2 + m) + 1, dtype=torch.float32, device=dev)
            d_b1, d_w2, d_b2 = small[:w], small[w:2 * w], small[2 * w + w * m:]
            d_wx = small[2 * w:2 * w + w * m].view(w, m) if m else None
        check(lib().dggb_edge_mlp_bwd(p(gr.erow), p(gr.col), i32(gr.nnz), i32(w), i32(ldp), p(p_uv), p(xe), i32(hx),
                                      p(edge_val), p(deg), p(wx), p(b1), p(w2), p(b2), float(slope),
                                      float(dist_scale), i32(flags), p(score), p(_f32c(g)), p(d_p), p(d_xe), p(d_wx),
                                      p(d_b1), p(d_w2), p(d_b2), stream()), "edge_mlp_bwd")
        d_wx, d_b1, d_w2, d_b2 = (None if (g_ is None or shp is None) else g_.reshape(shp)
                                  for g_, shp in zip((d_wx, d_b1, d_w2, d_b2), ctx.shapes))
        return d_p, d_xe, d_wx, d_b1, d_w2, d_b2, None, None, None, None, None, None


def edge_mlp(p_uv, wx, b1, w2, b2, graph, edge_val=None, deg=None, xe=None, flags=0, slope=0.01, dist_scale=1.0):
    """score_e = sigmoid(b2 + w2 . act(p_uv[u, :w] + p_uv[v, w:] + b1 + wx extra_e)); extras per ``flags``
    (EX_VAL: edge_val, EX_DEG: deg[u], deg[v], EX_DIST: exp(-dist_scale |xe_u - xe_v|)).  Differentiable in p_uv, xe
    and the weights; edge_val / deg are constants (the input graph)."""
    return _EdgeMLP.apply(p_uv, xe, wx, b1, w2, b2, graph, edge_val, deg, int(flags), float(slope), float(dist_scale))


def edge_dist_score(xe, graph, dist_scale):
    """score_e = exp(-dist_scale |xe_u - xe_v|_2)  ("u-v-dist", dgm.py:1618-1623)."""
    return _EdgeMLP.apply(None, xe, None, None, None, None, graph, None, None, DIST_ONLY, 1.0, float(dist_scale))


_PAD_CACHE = {}


def pad_features(x):
    """[N, F] -> [N, ceil4(F)] zero-padded copy, made ONCE per feature tensor (node features are static across
    epochs): the TMA-fed tensor-core encoder needs 16-byte row pitches (Cora F = 1433, Citeseer F = 3703).  The
    source tensor is kept alive by the cache so its address cannot be recycled; 4 slots."""
    f = x.shape[1]
    if f % 4 == 0 or x.requires_grad:
        return x
    key = (x.data_ptr(), tuple(x.shape), x._version, x.dtype)
    hit = _PAD_CACHE.get(key)
    if hit is not None:
        return hit[0]
    xp = torch.nn.functional.pad(x, (0, 4 - f % 4)).contiguous()
    if len(_PAD_CACHE) >= 4:
        _PAD_CACHE.pop(next(iter(_PAD_CACHE)))
    _PAD_CACHE[key] = (xp, x)
    return xp


def encoder_linear(x, w, b, slope, dropout=0.0, training=False):
    """LeakyReLU_slope(dropout(x) W^T + b) for the tall node encoders (slope = 0: ReLU, 1: plain Linear): on tcgen05
    also when F % 4 != 0 (features padded ONCE -- before the dropout, so the cached padded copy is reused every
    epoch -- and the [h, F] weight padded per call, it is tiny)."""
    f = x.shape[1]
    if f % 4 != 0 and x.is_cuda and not x.requires_grad:
        x = pad_features(x)
    if x.shape[1] > w.shape[1]:       # x padded here or by the caller (``pad_features``): zero weight columns for it
        w = torch.nn.functional.pad(w, (0, x.shape[1] - w.shape[1]))
    if dropout > 0.0 and training:
        x = torch.nn.functional.dropout(x, dropout, training=True)
    return tall_linear(x, w, b, slope)


class _GatAggregate(torch.autograd.Function):
    """GATConv_DGG / GATConv attention + aggregation for all heads (model.py:556-577); see include/dggb.h."""

    @staticmethod
    def forward(ctx, hd, pq, avals, htot, bias, graph: CSRGraph, heads: int, f: int, alpha: float, bg: float, keep):
        _require_cuda(hd, pq, avals, htot, bias, keep)
        hd, pq = _f32c(hd), _f32c(pq)
        avals = None if avals is None else _f32c(avals)
        htot = None if htot is None else _f32c(htot).reshape(-1)
        bias = None if bias is None else _f32c(bias).reshape(-1)
        keep = None if keep is None else _f32c(keep)
        n = graph.n
        out = torch.empty(n, heads * f, dtype=torch.float32, device=hd.device)
        mz = torch.empty(2, n, heads, dtype=torch.float32, device=hd.device)
        check(lib().dggb_gat_aggregate_fwd(p(graph.rowptr), p(graph.col), i32(n), i64(graph.nnz), i32(heads), i32(f),
                                           p(hd), i32(hd.shape[1]), p(pq), p(avals), p(keep), p(htot), p(bias),
                                           float(alpha), float(bg), p(out), i32(out.shape[1]), p(mz[0]), p(mz[1]),
                                           stream()), "gat_aggregate_fwd")
        ctx.graph, ctx.meta = graph, (heads, f, alpha, bg)
        ctx.save_for_backward(hd, pq, avals, htot, bias, keep, out, mz)
        return out

    @staticmethod
    def backward(ctx, g):
        hd, pq, avals, htot, bias, keep, out, mz = ctx.saved_tensors
        heads, f, alpha, bg = ctx.meta
        gr = ctx.graph
        g = _f32c(g)
        d_hd = torch.zeros_like(hd)
        d_pq = torch.zeros_like(pq)
        d_av = torch.zeros_like(avals) if (avals is not None and ctx.needs_input_grad[2]) else None
        d_ht = torch.zeros_like(htot) if (htot is not None and ctx.needs_input_grad[3]) else None
        check(lib().dggb_gat_aggregate_bwd(p(gr.rowptr), p(gr.col), i32(gr.n), i64(gr.nnz), i32(heads), i32(f), p(hd),
                                           i32(hd.shape[1]), p(pq), p(avals), p(keep), p(bias), float(alpha),
                                           float(bg), p(out), p(g), i32(out.shape[1]), p(mz[0]), p(mz[1]), p(d_hd),
                                           p(d_pq), p(d_av), p(d_ht), stream()), "gat_aggregate_bwd")
        d_bias = g.sum(0) if (bias is not None and ctx.needs_input_grad[4]) else None
        return d_hd, d_pq, d_av, d_ht, d_bias, None, None, None, None, None, None


def gat_aggregate(hd, pq, graph, heads, f, avals=None, htot=None, bias=None, alpha=0.2, bg=0.0, keep=None):
    """-> out [N, heads * f].  hd [N, heads * f] (head k in columns [k f, (k + 1) f)), pq [N, heads, 2]."""
    return _GatAggregate.apply(hd, pq, avals, htot, bias, graph, int(heads), int(f), float(alpha), float(bg), keep)


def row_sum(vals, graph):
    """Row sums of a CSR matrix (in_adj.to_dense().sum(-1), dgm.py:1568) without densifying."""
    ones = torch.ones(graph.n, 1, dtype=torch.float32, device=vals.device)
    # every column index is < n, so A @ 1 with the SpMM kernel is the row sum
    return spmm(vals, ones, graph).reshape(-1)


class _TallLinear(torch.autograd.Function):
    """y = act(x W^T + b) for TALL x ([N, F] with N >> F, out): the weight gradient dW = dpre^T x is a GEMM whose
    reduction runs over the N nodes and whose output is tiny, which cuBLAS runs on a handful of CTAs; it is
    evaluated split-K (batched over row chunks, then summed) so all SMs take part.  act = LeakyReLU(slope)
    (slope = 1: identity)."""

    @staticmethod
    def forward(ctx, x, w, b, slope: float):
        out = _linear_act_tc(x, w, b, slope)
        if out is None:      # shapes the tensor-core kernel does not take (e.g. F % 4 != 0): library GEMM
            pre = torch.nn.functional.linear(x, w, b)
            out = pre if slope == 1.0 else torch.nn.functional.leaky_relu(pre, slope)
        ctx.slope = slope
        ctx.save_for_backward(x, w, out)
        ctx.has_bias = b is not None
        return out

    @staticmethod
    def backward(ctx, g):
        x, w, out = ctx.saved_tensors
        dpre = g if ctx.slope == 1.0 else torch.ops.aten.leaky_relu_backward(g, out, ctx.slope, True)
        dx = dpre @ w if ctx.needs_input_grad[0] else None
        dw, db = gemm_tn(dpre, x, ctx.has_bias)
        return dx, dw, db, None


def _linear_act_tc(x, w, b, slope, w_transposed=False, addend=None, act_src=None, w2=None, zero=None):
    """out = epi(x W_eff^T + b + addend) through dggb_linear_fused (tcgen05, 3xTF32); None if the shape is
    not supported.  epi = LeakyReLU(slope), or * LeakyReLU'(act_src) when act_src is given (backward form)."""
    h = w.shape[1] if w_transposed else w.shape[0]
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[1] % 4 == 0
            and h in (16, 32, 64, 128) and x.shape[0] >= 512):
        return None
    x, w = x.contiguous(), w.contiguous()
    n, f_in = x.shape
    out = torch.empty(n, h, dtype=torch.float32, device=x.device)
    bb = None if b is None else b.contiguous()
    ad = None if addend is None else _f32c(addend)
    ac = None if act_src is None else _f32c(act_src)
    ws = torch.empty(2 * h * f_in + 2 * h * h, dtype=torch.float32, device=x.device)
    out2 = torch.empty(n, h, dtype=torch.float32, device=x.device) if w2 is not None else None
    w2c = None if w2 is None else w2.contiguous()
    # few row tiles and wide features (Cora / Citeseer raw inputs): partial-tile scratch for the split-K form
    sk_bytes = int(lib().dggb_linear_splitk_workspace_bytes(i32(n), i32(f_in), i32(h))) if w2 is None else 0
    sk = torch.empty(sk_bytes // 4, dtype=torch.float32, device=x.device) if sk_bytes > 0 else None
    rc = lib().dggb_linear_fused(p(x), p(w), i32(1 if w_transposed else 0), p(bb), p(ad), p(ac),
                                 float(slope), i32(n), i32(f_in), i32(h), p(out), p(w2c), p(out2), p(ws),
                                 i64(ws.numel() * 4), p(zero), i64(0 if zero is None else zero.numel()), p(sk),
                                 i64(sk_bytes), stream())
    if rc == -2:
        return None
    check(rc, "linear_fused")
    return out if w2 is None else (out, out2)


class _EncodeProject(torch.autograd.Function):
    """(x_enc, y) = (LeakyReLU(x Wn^T + bn), x_enc We^T): the node encoder and the edge-encoder projection of
    ``DGG`` (dgm.py:1778, 1784) as ONE autograd node.  Forward: two launches (weight split, GEMM with the chained
    second GEMM); the split launch also leaves the pre-split We^T and the cleared accumulators for the backward, which
    is then TWO launches and no fills:
        dpre = LeakyReLU'(x_enc) * (g_y We + g_xenc), leaving as a transposed TF32 split + column sums (= dbn), next
               to the same split of its input g_y
        dWn  = dpre^T x  and  dWe = g_y^T x_enc  in one tensor-core launch (dggb_gemm_tn_tc_presplit)"""

    @staticmethod
    def forward(ctx, x, wn, bn, we, slope: float):
        h, f_in = wn.shape
        n = x.shape[0]
        ctx.fast = h in (32, 64) and n >= 4096 and f_in >= 256      # the tensor-core weight-gradient GEMM's domain
        wet = zbuf = None
        if ctx.fast:
            x, wn, bn, we = x.contiguous(), wn.contiguous(), bn.contiguous(), we.contiguous()
            dev = x.device
            x_enc = torch.empty(n, h, dtype=torch.float32, device=dev)
            y = torch.empty(n, h, dtype=torch.float32, device=dev)
            ws = torch.empty(2 * h * f_in + 2 * h * h, dtype=torch.float32, device=dev)
            if any(ctx.needs_input_grad[:4]):
                wet = torch.empty(2 * h * h, dtype=torch.float32, device=dev)
                zbuf = torch.empty(h * h + h * f_in + h, dtype=torch.float32, device=dev)   # dWe | dWn | dbn
            check(lib().dggb_encoder_fwd(p(x), p(wn), p(bn), float(slope), i32(n), i32(f_in), i32(h), p(x_enc), p(we),
                                         p(y), p(ws), i64(ws.numel() * 4), p(wet), p(zbuf),
                                         i64(0 if zbuf is None else zbuf.numel()), stream()), "encoder_fwd")
        elif h in (32, 64):     # one launch: y chained in the encoder kernel's epilogue
            res = _linear_act_tc(x, wn, bn, slope, w2=we)
            x_enc, y = res if res is not None else (None, None)
        else:
            x_enc = _linear_act_tc(x, wn, bn, slope)
            y = _linear_act_tc(x_enc, we, None, 1.0) if x_enc is not None else None
        if y is None:
            raise RuntimeError("encode_project: shape not supported by the tensor-core kernels")
        ctx.slope = slope
        ctx.wet, ctx.zbuf = wet, zbuf
        ctx.save_for_backward(x, wn, we, x_enc)
        ctx.set_materialize_grads(False)   # an unused output must not cost an [N, h] zero fill
        return x_enc, y

    @staticmethod
    def backward(ctx, g_xenc, g_y):
        x, wn, we, x_enc = ctx.saved_tensors
        if g_xenc is None and g_y is None:
            return (None,) * 5
        g_y = torch.zeros_like(x_enc) if g_y is None else _f32c(g_y)
        g_xenc = None if g_xenc is None else _f32c(g_xenc)
        h, f_in = wn.shape
        n = x.shape[0]
        if ctx.fast and ctx.wet is not None:
            dev = x.device
            zbuf, ctx.zbuf = ctx.zbuf, None       # cleared by the forward's split launch; a second backward: fresh zeros
            if zbuf is None:
                zbuf = torch.zeros(h * h + h * f_in + h, dtype=torch.float32, device=dev)
            dwe, dwn, dbn = zbuf[:h * h].view(h, h), zbuf[h * h:h * h + h * f_in].view(h, f_in), zbuf[h * h + h * f_in:]
            npad = (n + 31) // 32 * 32
            want_dx = ctx.needs_input_grad[0]
            dpre = torch.empty(n, h, dtype=torch.float32, device=dev) if want_dx else None
            tsp = torch.empty(4, h, npad, dtype=torch.float32, device=dev)   # dpre^T hi | lo | g_y^T hi | lo
            check(lib().dggb_encoder_bwd_dpre(p(g_y), p(ctx.wet), p(g_xenc), p(x_enc), float(ctx.slope), i32(n), i32(h),
                                              p(dpre), p(tsp[0]), p(tsp[1]), i32(npad), p(dbn), p(tsp[2]), p(tsp[3]),
                                              stream()), "encoder_bwd_dpre")
            # dWn = dpre^T x and, from extra CTAs of the same launch, dWe = g_y^T x_enc
            check(lib().dggb_gemm_tn_tc_presplit(p(tsp[0]), p(tsp[1]), i32(npad), p(x), i32(n), i32(h), i32(f_in),
                                                 p(dwn), p(tsp[2]), p(tsp[3]), p(x_enc), i32(h), p(dwe), stream()),
                  "gemm_tn_tc_presplit")
            dx = dpre @ wn if want_dx else None
            return dx, dwn, dbn, dwe, None
        # split-K accumulators of both weight-gradient GEMMs: cleared by the dpre launch (its weight-split kernel)
        zbuf = torch.empty(h * h + h * f_in + h, dtype=torch.float32, device=x.device)
        dpre = _linear_act_tc(g_y, we, None, ctx.slope, w_transposed=True, addend=g_xenc, act_src=x_enc, zero=zbuf)
        if dpre is None:
            raise RuntimeError("encode_project backward: shape not supported by the tensor-core kernels")
        dwe, _ = gemm_tn(g_y, x_enc, False, zeroed=zbuf[:h * h])
        dwn, dbn = gemm_tn(dpre, x, True, zeroed=zbuf[h * h:])
        dx = dpre @ wn if ctx.needs_input_grad[0] else None
        return dx, dwn, dbn, dwe, None


def encode_project(x, wn, bn, we, slope):
    """-> (x_enc, y); falls back to two tall_linear nodes when the fused tensor-core path does not apply."""
    n, f_in = x.shape
    h = wn.shape[0]
    if (x.is_cuda and x.dtype == torch.float32 and f_in % 4 == 0 and h in (16, 32, 64, 128) and n >= 512
            and tuple(we.shape) == (h, h) and bn is not None):
        return _EncodeProject.apply(x, wn, bn, we, float(slope))
    x_enc = tall_linear(x, wn, bn, slope)
    return x_enc, tall_linear(x_enc, we)


def gemm_tn(a, b, want_colsum=False, use_tc=None, zeroed=None):
    """(a^T b [P,Q], column sums of a [P] or None) for tall a [N,P], b [N,Q] via dggb_gemm_tn_splitk / _tc.
    ``zeroed``: optional pre-zeroed flat fp32 buffer of pp*q (+pp) elements to accumulate into."""
    a, b = _f32c(a), _f32c(b)
    n, pp = a.shape
    q = b.shape[1]
    if q % 4 != 0 and q < 64 and pp % 4 == 0 and zeroed is None:
        # narrow right operand (e.g. the 3 class logits): the kernel's 16-byte staging path needs q % 4 == 0;
        # padding N x q to N x 4 costs one small copy, the scalar staging path costs 3x the whole GEMM
        out, cs = gemm_tn(a, torch.nn.functional.pad(b, (0, 4 - q % 4)), want_colsum, use_tc)
        return out[:, :q].contiguous(), cs
    if pp % 4 != 0 and pp < 64 and q % 4 == 0 and zeroed is None:
        # narrow LEFT operand (dW of a narrowing layer, e.g. 64 -> 3 class logits: a = d logits [N, 3]): the transposed
        # product b^T a has the narrow operand on the right, where padding it to 4 columns keeps the 16-byte staging
        # path (measured at Pubmed shape: 20.7 us on the scalar path)
        out_t, _ = gemm_tn(b, torch.nn.functional.pad(a, (0, 4 - pp % 4)), False, use_tc)
        return out_t[:, :pp].t().contiguous(), (a.sum(0) if want_colsum else None)
    need = pp * q + (pp if want_colsum else 0)
    buf = zeroed if zeroed is not None else torch.zeros(need, dtype=torch.float32, device=a.device)
    assert buf.numel() == need
    out = buf[:pp * q].view(pp, q)
    cs = buf[pp * q:] if want_colsum else None
    if use_tc is None:
        # wide outputs amortise the transpose pre-pass; narrow ones (dWe, Q = h) stay on the SIMT split-K kernel
        use_tc = n >= _TN_TC_MIN_N and q % 4 == 0 and q >= 256 and pp in (16, 32, 64, 128)
    if use_tc:
        L = lib()
        ws_bytes = int(L.dggb_gemm_tn_tc_workspace_bytes(i32(n), i32(pp)))
        ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=a.device)
        check(L.dggb_gemm_tn_tc(p(a), p(b), i32(n), i32(pp), i32(q), p(out), p(cs), p(ws), i64(ws_bytes),
                                stream()), "gemm_tn_tc")
    else:
        check(lib().dggb_gemm_tn_splitk(p(a), p(b), i32(n), i32(pp), i32(q), p(out), p(cs), stream()),
              "gemm_tn_splitk")
    return out, cs


def tall_linear(x, w, b=None, slope=1.0):
    return _TallLinear.apply(x, w, b, float(slope))


class _TallMatmul(torch.autograd.Function):
    """act(x [N, P] @ w [P, Q]) for tall x (the (A x) W products of the conv layers, model.py:596, 74; act = ReLU or
    identity).  Forward and d x = dpre w^T run on the tcgen05 3xTF32 kernel (dggb_linear_fused; the ReLU is its
    epilogue) when its shapes allow -- Q resp. P in {16, 32, 64, 128}, the other one a multiple of 4 -- and on the
    library GEMM otherwise (e.g. the 64 -> 3 class logits).  The weight gradient x^T dpre reduces over the N nodes into
    a tiny [P, Q] output, which cuBLAS leaves on a handful of CTAs (81 us at Pubmed shape): split-K kernels instead."""

    @staticmethod
    def forward(ctx, x, w, relu: bool):
        out = _linear_act_tc(x, w, None, 0.0 if relu else 1.0, w_transposed=True)
        if out is None:
            out = torch.mm(x, w)
            if relu:
                out = torch.relu_(out)
        ctx.relu = relu
        ctx.save_for_backward(x, w, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, g):
        x, w, out = ctx.saved_tensors
        g = _f32c(g)
        dpre = torch.ops.aten.threshold_backward(g, out, 0.0) if ctx.relu else g
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _linear_act_tc(dpre, w, None, 1.0)           # dpre [N, Q] w[P, Q]^T: w is the [h = P, F = Q] weight
            if dx is None:
                dx = torch.mm(dpre, w.t())
        dw = gemm_tn(x, dpre, False)[0] if ctx.needs_input_grad[1] else None
        return dx, dw, None


class _HeadDots(torch.autograd.Function):
    """pq[n, k, c] = sum_f h[n, k, f] a[k, f, c] -- the per-node halves p_i = a1^T h_i, q_j = a2^T h_j of the GAT logits
    e_ij = LeakyReLU(a^T [h_i || h_j]) (model.py:561-563) for all heads.  The forward is a skinny batched product; its
    weight gradient da[k] = h_k^T g_k reduces over the N nodes into a [F, 2] output per head, which the batched
    library GEMM runs on a handful of CTAs (455 us at Pubmed shape, 8 heads): it goes through the split-K kernel as
    ONE [heads F, N] x [N, 2 heads] product whose diagonal blocks are the answer (~15 us)."""

    @staticmethod
    def forward(ctx, h, a):
        ctx.save_for_backward(h, a)
        return torch.einsum("nkf,kfc->nkc", h, a)

    @staticmethod
    def backward(ctx, g):
        h, a = ctx.saved_tensors
        n, heads, f = h.shape
        g = _f32c(g)
        dh = torch.einsum("nkc,kfc->nkf", g, a) if ctx.needs_input_grad[0] else None
        da = None
        if ctx.needs_input_grad[1]:
            if h.is_cuda and n >= 2048 and (heads * f) % 4 == 0:
                full = gemm_tn(h.reshape(n, heads * f), g.reshape(n, heads * 2), False)[0]      # [heads f, heads 2]
                full = full.view(heads, f, heads, 2)
                idx = torch.arange(heads, device=h.device)
                da = full[idx, :, idx, :]                                                        # diagonal blocks
            else:
                da = torch.einsum("nkf,nkc->kfc", h, g)
        return dh, da


def head_dots(h, a):
    return _HeadDots.apply(h, a)


def tall_matmul(x, w, relu=False):
    """relu?(x @ w) for a tall x [N, P] and a small w [P, Q]."""
    if x.is_cuda and x.dtype == torch.float32 and x.shape[0] >= 2048 and x.shape[1] <= 512:
        return _TallMatmul.apply(x, w, bool(relu))
    out = torch.mm(x, w)
    return torch.relu(out) if relu else out
