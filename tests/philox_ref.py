"""Host restatement (torch int64 ops) of the counter-based Gumbel noise generated inside the all-pairs
kernel (csrc/allpairs.cu: philox4x32_7 + gumbel_from_bits), so tests can materialise the same N x N
matrix and feed it to the dense oracle."""
import torch

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_7(c0, c1, seed):
    """c0, c1: int64 tensors holding uint32 counters -> 4 int64 tensors of uint32 outputs."""
    k0, k1 = seed & MASK, (seed >> 32) & MASK
    x0, x1 = c0.clone(), c1.clone()
    x2, x3 = torch.zeros_like(c0), torch.zeros_like(c0)

    def mulhilo(a, b):
        # 32x32 -> 64-bit product without overflowing int64: split b into 16-bit halves
        bl, bh = b & 0xFFFF, b >> 16
        lo_part = a * bl                      # < 2^48
        hi_part = a * bh                      # < 2^48
        total_lo = (lo_part & MASK) + ((hi_part & 0xFFFF) << 16)
        lo = total_lo & MASK
        hi = (lo_part >> 32) + (hi_part >> 16) + (total_lo >> 32)
        return hi & MASK, lo

    for _ in range(7):
        hi0, lo0 = mulhilo(torch.full_like(x0, M0), x0)
        hi1, lo1 = mulhilo(torch.full_like(x2, M1), x2)
        y0 = hi1 ^ x1 ^ k0
        y2 = hi0 ^ x3 ^ k1
        x0, x1, x2, x3 = y0, lo1, y2, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return x0, x1, x2, x3


def gumbel_matrix(n_rows, n_cols, seed, scale, row_begin=0):
    rows = torch.arange(row_begin, row_begin + n_rows, dtype=torch.int64).reshape(-1, 1)
    cols = torch.arange(n_cols, dtype=torch.int64).reshape(1, -1)
    c0 = rows.expand(n_rows, n_cols).contiguous()
    c1 = (cols >> 2).expand(n_rows, n_cols).contiguous()
    outs = philox4x32_7(c0, c1, seed)
    lane = (cols & 3).expand(n_rows, n_cols)
    bits = torch.where(lane == 0, outs[0], torch.where(lane == 1, outs[1], torch.where(lane == 2, outs[2], outs[3])))
    v = ((bits >> 8).to(torch.float64) + 0.5) * (2.0 ** -24)
    return (-scale * torch.log(-torch.log1p(-v))).to(torch.float32)
