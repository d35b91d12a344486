// GATConv_DGG / GATConv aggregation (model.py:556-577, 510-531), all heads in ONE launch per direction.
//
// The reference builds a dense [N, N] attention per head: -1e20 everywhere, the listed edges' logits scattered in,
// multiplied by the dense DGG adjacency, softmax over all N columns, dropout, matmul.  Because the mask is applied by
// MULTIPLICATION, a non-listed pair has logit -1e20 * 0 = -0.0: every row's softmax runs over all N columns ("dense
// background").  With s_e = LeakyReLU_alpha(p_i + q_j) * A_ij on the stored entries, m_i = max(0, max_e s_e),
// x_e = exp(s_e - m_i), em_i = exp(-m_i):
//     Z_i   = sum_e (x_e - em_i) + N em_i
//     out_i = [ sum_e (kappa_e x_e - em_i) hd_j + em_i sum_all hd ] / Z_i + bias        (kappa: attention dropout)
// i.e. an edge softmax + SpMM plus one rank-1 background term; O(E F) instead of 5 dense N x N temporaries per head.
// bg == 0 selects the plain GATConv (logits -1e20 off the edge list: a true masked softmax; em := 0, m = max_e s_e).
// p_i = h_i . a[:F], q_j = h_j . a[F:] (e_ij = LeakyReLU(a^T [h_i || h_j]) split into two N-vectors, computed by the
// caller).  Heads are concatenated along the feature axis: head k owns columns [k F, (k+1) F) of hd / out.
// HBM-bound: E * heads * (F * 4 gathered + 8) + N * heads * F * 4.
#include "common.cuh"
#include <cstdlib>

namespace dggb {

constexpr int kGatWarps = 8;

__device__ __forceinline__ float4 ld4g(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

struct GatArgs {
  const int32_t* rowptr;
  const int32_t* col;
  int n, heads, f, L;
  long long nnz;
  const float* hd;     // [N, ldh]
  int ldh;
  const float* pq;     // [N, heads, 2]
  const float* aval;   // [E] or NULL
  const float* keep;   // [heads, E] attention-dropout multipliers (0 or 1/(1-p)) or NULL
  const float* htot;   // [heads * F] column sums of hd (background term) or NULL
  const float* bias;   // [heads * F] or NULL
  float alpha;
  float bg;            // number of background columns per row (N), 0: plain masked softmax
};

__device__ __forceinline__ float gat_logit(const GatArgs& A, float p_i, int c, int k, long long e) {
  const float pre = p_i + __ldg(A.pq + ((size_t)c * A.heads + k) * 2 + 1);
  const float lr = pre > 0.f ? pre : A.alpha * pre;
  return A.aval ? lr * __ldg(A.aval + e) : lr;
}

template <int T>
__global__ void __launch_bounds__(kGatWarps* kWarp)
    gat_fwd_kernel(GatArgs A, float* __restrict__ out, int ldo, float* __restrict__ m_out, float* __restrict__ z_out) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int L = A.L, G = kWarp / L, lg = lane % L, grp = lane / L;
  const long long items = (long long)A.n * A.heads;
  for (long long item = (long long)blockIdx.x * kGatWarps + (threadIdx.x >> 5); item < items;
       item += (long long)gridDim.x * kGatWarps) {
    const int i = (int)(item / A.heads), k = (int)(item % A.heads);
    const int beg = __ldg(A.rowptr + i), end = __ldg(A.rowptr + i + 1);
    const float p_i = __ldg(A.pq + (size_t)item * 2);
    // pass 1: row maximum of the logits (floored at 0 = the background's logit)
    float mx = A.bg > 0.f ? 0.f : -INFINITY;
    for (int e = beg + lane; e < end; e += kWarp) mx = fmaxf(mx, gat_logit(A, p_i, __ldg(A.col + e), k, e));
    mx = warp_max(mx);
    const float em = A.bg > 0.f ? __expf(-mx) : 0.f;
    // pass 2: weights + weighted sum of the neighbour rows
    float4 acc[T];
#pragma unroll
    for (int t = 0; t < T; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    float zsum = 0.f;
    for (int w0 = beg; w0 < end; w0 += kWarp) {
      const int e_l = w0 + lane;
      const bool ok = e_l < end;
      const int c_l = ok ? __ldg(A.col + e_l) : 0;
      float a_l = 0.f;
      if (ok) {
        const float x = __expf(gat_logit(A, p_i, c_l, k, e_l) - mx);
        zsum += x - em;
        a_l = (A.keep ? __ldg(A.keep + (size_t)k * A.nnz + e_l) : 1.f) * x - em;
      }
      const int cnt = min(kWarp, end - w0);
#pragma unroll 4
      for (int j0 = 0; j0 < cnt; j0 += G) {
        const int j = j0 + grp;
        const int v = __shfl_sync(0xffffffffu, c_l, j & 31);
        float a = __shfl_sync(0xffffffffu, a_l, j & 31);
        if (j >= cnt) a = 0.f;
        const float* xr = A.hd + (size_t)v * A.ldh + (size_t)k * A.f;
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const int c = 4 * (lg + L * t);
          if (c < A.f) {
            const float4 xv = ld4g(xr + c);
            acc[t].x = fmaf(a, xv.x, acc[t].x); acc[t].y = fmaf(a, xv.y, acc[t].y);
            acc[t].z = fmaf(a, xv.z, acc[t].z); acc[t].w = fmaf(a, xv.w, acc[t].w);
          }
        }
      }
    }
    const float Z = warp_sum(zsum) + A.bg * em;
    const float rz = 1.f / Z;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      for (int o = L; o < kWarp; o <<= 1) {
        acc[t].x += __shfl_xor_sync(0xffffffffu, acc[t].x, o); acc[t].y += __shfl_xor_sync(0xffffffffu, acc[t].y, o);
        acc[t].z += __shfl_xor_sync(0xffffffffu, acc[t].z, o); acc[t].w += __shfl_xor_sync(0xffffffffu, acc[t].w, o);
      }
      const int c = 4 * (lg + L * t);
      if (grp == 0 && c < A.f) {
        const int cc = k * A.f + c;
        float4 v = acc[t];
        if (A.htot && A.bg > 0.f) {
          const float4 ht = ld4g(A.htot + cc);
          v.x = fmaf(em, ht.x, v.x); v.y = fmaf(em, ht.y, v.y); v.z = fmaf(em, ht.z, v.z); v.w = fmaf(em, ht.w, v.w);
        }
        v.x *= rz; v.y *= rz; v.z *= rz; v.w *= rz;
        if (A.bias) {
          const float4 b = ld4g(A.bias + cc);
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        st4(out + (size_t)i * ldo + cc, v);
      }
    }
    if (lane == 0) {
      m_out[item] = mx;
      z_out[item] = Z;
    }
  }
}

// Backward.  With dnum_i = g_i / Z_i and c_i = <g_i, out_i - bias>:
//   d x_e = (kappa_e <g_i, hd_j> - c_i) / Z_i ;  d s_e = d x_e x_e   (m_i is a constant: the softmax is shift invariant)
//   d hd_j += (kappa_e x_e - em_i) dnum_i ;  d htot += em_i dnum_i ;  d A_e += d s_e e_e ;  d pre_e = d s_e A_e LReLU'(pre_e)
//   d p_i += d pre_e ;  d q_j += d pre_e
template <int T>
__global__ void __launch_bounds__(kGatWarps* kWarp)
    gat_bwd_kernel(GatArgs A, const float* __restrict__ g_out, const float* __restrict__ out, int ldo,
                   const float* __restrict__ m_in, const float* __restrict__ z_in, float* __restrict__ d_hd,
                   float* __restrict__ d_pq, float* __restrict__ d_aval, float* __restrict__ d_htot) {
  pdl_trigger();
  extern __shared__ float ht_s[];    // [heads * F] block-level accumulator of d htot
  const int HF = A.heads * A.f;
  if (d_htot != nullptr)
    for (int c = threadIdx.x; c < HF; c += blockDim.x) ht_s[c] = 0.f;
  pdl_wait();
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int L = A.L, G = kWarp / L, lg = lane % L, grp = lane / L;
  const long long items = (long long)A.n * A.heads;
  for (long long item = (long long)blockIdx.x * kGatWarps + (threadIdx.x >> 5); item < items;
       item += (long long)gridDim.x * kGatWarps) {
    const int i = (int)(item / A.heads), k = (int)(item % A.heads);
    const int beg = __ldg(A.rowptr + i), end = __ldg(A.rowptr + i + 1);
    const float p_i = __ldg(A.pq + (size_t)item * 2);
    const float mx = __ldg(m_in + item), Z = __ldg(z_in + item);
    const float em = A.bg > 0.f ? __expf(-mx) : 0.f;
    const float rz = 1.f / Z;
    float4 g[T];
    float ci = 0.f;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int c = 4 * (lg + L * t);
      g[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < A.f) {
        const int cc = k * A.f + c;
        g[t] = ld4g(g_out + (size_t)i * ldo + cc);
        float4 o = ld4g(out + (size_t)i * ldo + cc);
        if (A.bias) {
          const float4 b = ld4g(A.bias + cc);
          o.x -= b.x; o.y -= b.y; o.z -= b.z; o.w -= b.w;
        }
        ci += g[t].x * o.x + g[t].y * o.y + g[t].z * o.z + g[t].w * o.w;
        g[t].x *= rz; g[t].y *= rz; g[t].z *= rz; g[t].w *= rz;      // g now holds dnum_i
        if (d_htot != nullptr && grp == 0 && em != 0.f) {
          atomicAdd(&ht_s[cc + 0], em * g[t].x); atomicAdd(&ht_s[cc + 1], em * g[t].y);
          atomicAdd(&ht_s[cc + 2], em * g[t].z); atomicAdd(&ht_s[cc + 3], em * g[t].w);
        }
      }
    }
    ci = group_sum(ci, L);             // every group holds the full <g_i, o_i>
    float dp = 0.f;
    for (int w0 = beg; w0 < end; w0 += kWarp) {
      const int e_l = w0 + lane;
      const bool ok = e_l < end;
      const int c_l = ok ? __ldg(A.col + e_l) : 0;
      float pre_l = 0.f, av_l = 1.f, x_l = 0.f, kp_l = 1.f;
      if (ok) {
        pre_l = p_i + __ldg(A.pq + ((size_t)c_l * A.heads + k) * 2 + 1);
        av_l = A.aval ? __ldg(A.aval + e_l) : 1.f;
        x_l = __expf((pre_l > 0.f ? pre_l : A.alpha * pre_l) * av_l - mx);
        kp_l = A.keep ? __ldg(A.keep + (size_t)k * A.nnz + e_l) : 1.f;
      }
      const int cnt = min(kWarp, end - w0);
#pragma unroll 2
      for (int j0 = 0; j0 < cnt; j0 += G) {
        const int j = j0 + grp;
        const bool valid = j < cnt;
        const int v = __shfl_sync(0xffffffffu, c_l, j & 31);
        const float x = __shfl_sync(0xffffffffu, x_l, j & 31);
        const float kp = __shfl_sync(0xffffffffu, kp_l, j & 31);
        const float pre = __shfl_sync(0xffffffffu, pre_l, j & 31);
        const float av = __shfl_sync(0xffffffffu, av_l, j & 31);
        const float wgt = kp * x - em;
        float dot = 0.f;
        const size_t roff = (size_t)v * A.ldh + (size_t)k * A.f;
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const int c = 4 * (lg + L * t);
          if (c < A.f && valid) {
            const float4 xv = ld4g(A.hd + roff + c);
            dot += g[t].x * xv.x + g[t].y * xv.y + g[t].z * xv.z + g[t].w * xv.w;
            red_add4(d_hd + roff + c, make_float4(wgt * g[t].x, wgt * g[t].y, wgt * g[t].z, wgt * g[t].w));
          }
        }
        dot = group_sum(dot, L) * Z;                  // <g_i, hd_j> (g holds g_i / Z)
        if (lg == 0 && valid) {
          const float dx = (kp * dot - ci) * rz;
          const float ds = dx * x;
          const float lr = pre > 0.f ? pre : A.alpha * pre;
          if (d_aval != nullptr) atomicAdd(d_aval + w0 + j, ds * lr);
          const float dpre = ds * av * (pre > 0.f ? 1.f : A.alpha);
          atomicAdd(d_pq + ((size_t)v * A.heads + k) * 2 + 1, dpre);
          dp += dpre;
        }
      }
    }
    dp = warp_sum(dp);
    if (lane == 0) atomicAdd(d_pq + (size_t)item * 2, dp);
  }
  if (d_htot != nullptr) {
    __syncthreads();
    for (int c = threadIdx.x; c < HF; c += blockDim.x)
      if (ht_s[c] != 0.f) atomicAdd(d_htot + c, ht_s[c]);
  }
}

// ------------------------------------------------------------------------------------------------
// Row-warp variants (heads * F <= 1024, F a power of two <= 128): ONE warp owns a row for ALL heads.  The 32 lanes
// cover the whole concatenated feature row (heads * F floats = one contiguous 2 KB line at 8 x 64): lane l owns the
// 16-byte chunks c = l + 32 t, chunk c belongs to head 4c / F.  The row's entries are walked sequentially, every
// neighbour row is read with fully coalesced 128-bit loads and accumulated with that chunk's head weight -- no
// cross-lane reduction of the aggregate at all.  The per-(entry, head) weights are computed by lane = (entry slot,
// head) in batches of 32 / heads entries and handed over by shuffles.  (The kernels above give a warp to every
// (row, head) pair: at 8 heads x 64 that is 111 M warp instructions per forward, ncu r02: 275 us; this one: 150 us
// incl. the column sum of hd, against ~40 us for the 221 MB it gathers from L2.)
// ------------------------------------------------------------------------------------------------
template <int CPL>
__global__ void __launch_bounds__(kGatWarps* kWarp)
    gat_row_fwd_kernel(GatArgs A, int HP /* pow2 >= heads */, float* __restrict__ out, int ldo,
                       float* __restrict__ m_out, float* __restrict__ z_out) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int C = A.heads * A.f / 4, cph = A.f / 4;           // chunks per row / per head
  const int EB = kWarp / HP, es = lane / HP, hh = lane % HP; // weight phase: lane = (entry slot, head)
  const bool head_ok = hh < A.heads;
  int ht[CPL];                                               // aggregation phase: head of each owned chunk
#pragma unroll
  for (int t = 0; t < CPL; ++t) ht[t] = min((lane + 32 * t) / cph, A.heads - 1);
  for (int i = blockIdx.x * kGatWarps + (threadIdx.x >> 5); i < A.n; i += gridDim.x * kGatWarps) {
    const int beg = __ldg(A.rowptr + i), end = __ldg(A.rowptr + i + 1);
    const float p_i = head_ok ? __ldg(A.pq + ((size_t)i * A.heads + hh) * 2) : 0.f;
    // pass 1: per-head maximum of the logits (floored at 0 = the background's logit)
    float mx = A.bg > 0.f ? 0.f : -INFINITY;
    for (int e0 = beg; e0 < end; e0 += EB) {
      const int e = e0 + es;
      if (e < end && head_ok) mx = fmaxf(mx, gat_logit(A, p_i, __ldg(A.col + e), hh, e));
    }
    for (int o = HP; o < kWarp; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float em = A.bg > 0.f ? __expf(-mx) : 0.f;
    // pass 2: weights, then the weighted sum of the neighbour rows
    float4 acc[CPL];
#pragma unroll
    for (int t = 0; t < CPL; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    float zsum = 0.f;
    for (int e0 = beg; e0 < end; e0 += EB) {
      const int e = e0 + es;
      const bool ok = e < end && head_ok;
      const int c_l = (e < end) ? __ldg(A.col + e) : 0;
      float w_l = 0.f;
      if (ok) {
        const float x = __expf(gat_logit(A, p_i, c_l, hh, e) - mx);
        zsum += x - em;
        w_l = (A.keep ? __ldg(A.keep + (size_t)hh * A.nnz + e) : 1.f) * x - em;
      }
      const int cnt = min(EB, end - e0);
      for (int j = 0; j < cnt; ++j) {
        const int v = __shfl_sync(0xffffffffu, c_l, j * HP);
        const float* xr = A.hd + (size_t)v * A.ldh;
#pragma unroll
        for (int t = 0; t < CPL; ++t) {
          const float w = __shfl_sync(0xffffffffu, w_l, j * HP + ht[t]);
          const int c = lane + 32 * t;
          if (c < C) {
            const float4 xv = ld4g(xr + 4 * c);
            acc[t].x = fmaf(w, xv.x, acc[t].x); acc[t].y = fmaf(w, xv.y, acc[t].y);
            acc[t].z = fmaf(w, xv.z, acc[t].z); acc[t].w = fmaf(w, xv.w, acc[t].w);
          }
        }
      }
    }
    for (int o = HP; o < kWarp; o <<= 1) zsum += __shfl_xor_sync(0xffffffffu, zsum, o);
    const float Z = zsum + A.bg * em;                        // per head, held by every lane with that hh
#pragma unroll
    for (int t = 0; t < CPL; ++t) {
      const float em_t = __shfl_sync(0xffffffffu, em, ht[t]);
      const float rz = 1.f / __shfl_sync(0xffffffffu, Z, ht[t]);
      const int c = lane + 32 * t;
      if (c < C) {
        float4 v = acc[t];
        if (A.htot && A.bg > 0.f) {
          const float4 hv = ld4g(A.htot + 4 * c);
          v.x = fmaf(em_t, hv.x, v.x); v.y = fmaf(em_t, hv.y, v.y); v.z = fmaf(em_t, hv.z, v.z); v.w = fmaf(em_t, hv.w, v.w);
        }
        v.x *= rz; v.y *= rz; v.z *= rz; v.w *= rz;
        if (A.bias) {
          const float4 b = ld4g(A.bias + 4 * c);
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        st4(out + (size_t)i * ldo + 4 * c, v);
      }
    }
    if (es == 0 && head_ok) {
      m_out[(size_t)i * A.heads + hh] = mx;
      z_out[(size_t)i * A.heads + hh] = Z;
    }
  }
}

// (A row-warp BACKWARD was measured as well and rejected: 605 us vs 374 us for the per-(row, head) kernel at Pubmed
// shape, 8 x 64 -- every lane recomputes the entry's logit / exp for each of its chunks' heads and the per-head dot
// needs four shuffle stages per chunk; the backward's cost is the 221 MB of vector reductions into d_hd either way.)

// row-warp kernels apply when a head's chunks are a power of two <= 32 and the row fits 8 chunks per lane
static bool gat_row_ok(int heads, int f, int* cpl_out, int* hp_out) {
  const int cph = f / 4;
  if (f % 4 != 0 || cph > 32 || (cph & (cph - 1)) != 0 || heads > 32) return false;
  const int C = heads * cph;
  const int cpl = (C + 31) / 32;
  if (cpl > 8) return false;
  int hp = 1;
  while (hp < heads) hp *= 2;
  *cpl_out = cpl <= 1 ? 1 : (cpl <= 2 ? 2 : (cpl <= 4 ? 4 : 8));
  *hp_out = hp;
  return true;
}

static int gat_lanes(int f, int* t_out) {
  const int chunks = f / 4;
  int L = 1;
  while (L < 32 && 4 * L < chunks) L *= 2;
  int T = (chunks + L - 1) / L;
  *t_out = T > 2 ? 4 : (T >= 2 ? 2 : 1);
  return L;
}

static int gat_check(const GatArgs& A) {
  if (!A.rowptr || !A.col || !A.hd || !A.pq || A.n < 0 || A.heads <= 0 || A.f <= 0) return DGGB_ERR_BAD_ARG;
  if (A.f % 4 != 0 || A.f > 512 || A.ldh % 4 != 0 || A.ldh < A.heads * A.f || ((uintptr_t)A.hd % 16) ||
      (A.htot && ((uintptr_t)A.htot % 16)) || (A.bias && ((uintptr_t)A.bias % 16)))
    return DGGB_ERR_BAD_SHAPE;
  if (A.bg < 0.f) return DGGB_ERR_BAD_ARG;
  return DGGB_OK;
}

}  // namespace dggb
using namespace dggb;

extern "C" int dggb_gat_aggregate_fwd(const int32_t* rowptr, const int32_t* col, int32_t n, int64_t nnz,
                                      int32_t heads, int32_t f, const float* hd, int32_t ldh, const float* pq,
                                      const float* adj_val, const float* keep, const float* htot,
                                      const float* bias, float alpha, float bg_count, float* out, int32_t ldo,
                                      float* m_out, float* z_out, void* stream) {
  int T = 1;
  GatArgs A{rowptr, col, n, heads, f, 1, nnz, hd, ldh, pq, adj_val, keep, htot, bias, alpha, bg_count};
  int rc = gat_check(A);
  if (rc != DGGB_OK) return rc;
  if (!out || !m_out || !z_out || ldo % 4 != 0 || ldo < heads * f || ((uintptr_t)out % 16)) return DGGB_ERR_BAD_ARG;
  if (n == 0) return DGGB_OK;
  int cpl = 1, hp = 1;
  static const bool no_row = getenv("DGGB_GAT_NO_ROWWARP") != nullptr;
  if (!no_row && heads > 1 && gat_row_ok(heads, f, &cpl, &hp)) {
    auto gor = [&](auto kern) {
      const int grid = rows_grid(n, kGatWarps, resident_blocks(kern, kGatWarps * kWarp));
      launch_pdl(kern, dim3(grid), dim3(kGatWarps * kWarp), 0, as_stream(stream), A, hp, out, (int)ldo, m_out, z_out);
      return launch_status();
    };
    return cpl == 1 ? gor(gat_row_fwd_kernel<1>) : (cpl == 2 ? gor(gat_row_fwd_kernel<2>)
                    : (cpl == 4 ? gor(gat_row_fwd_kernel<4>) : gor(gat_row_fwd_kernel<8>)));
  }
  A.L = gat_lanes(f, &T);
  auto go = [&](auto kern) {
    const int grid = rows_grid((int)std::min<long long>((long long)n * heads, 1ll << 30), kGatWarps,
                               resident_blocks(kern, kGatWarps * kWarp));
    launch_pdl(kern, dim3(grid), dim3(kGatWarps * kWarp), 0, as_stream(stream), A, out, (int)ldo, m_out, z_out);
    return launch_status();
  };
  return T == 1 ? go(gat_fwd_kernel<1>) : (T == 2 ? go(gat_fwd_kernel<2>) : go(gat_fwd_kernel<4>));
}

extern "C" int dggb_gat_aggregate_bwd(const int32_t* rowptr, const int32_t* col, int32_t n, int64_t nnz,
                                      int32_t heads, int32_t f, const float* hd, int32_t ldh, const float* pq,
                                      const float* adj_val, const float* keep, const float* bias, float alpha,
                                      float bg_count, const float* out, const float* g_out, int32_t ldo,
                                      const float* m_in, const float* z_in, float* d_hd, float* d_pq,
                                      float* d_adj_val, float* d_htot, void* stream) {
  int T = 1;
  GatArgs A{rowptr, col, n, heads, f, 1, nnz, hd, ldh, pq, adj_val, keep, nullptr, bias, alpha, bg_count};
  int rc = gat_check(A);
  if (rc != DGGB_OK) return rc;
  if (!out || !g_out || !m_in || !z_in || !d_hd || !d_pq || ldo % 4 != 0 || ((uintptr_t)g_out % 16) ||
      ((uintptr_t)out % 16) || ((uintptr_t)d_hd % 16))
    return DGGB_ERR_BAD_ARG;
  if (n == 0) return DGGB_OK;
  const size_t smem = d_htot ? (size_t)heads * f * sizeof(float) : 0;
  if (smem > 48 * 1024) return DGGB_ERR_BAD_SHAPE;
  A.L = gat_lanes(f, &T);
  auto go = [&](auto kern) {
    const int grid = rows_grid((int)std::min<long long>((long long)n * heads, 1ll << 30), kGatWarps,
                               resident_blocks(kern, kGatWarps * kWarp, smem));
    launch_pdl(kern, dim3(grid), dim3(kGatWarps * kWarp), smem, as_stream(stream), A, g_out, out, (int)ldo, m_in,
               z_in, d_hd, d_pq, d_adj_val, d_htot);
    return launch_status();
  };
  return T == 1 ? go(gat_bwd_kernel<1>) : (T == 2 ? go(gat_bwd_kernel<2>) : go(gat_bwd_kernel<4>));
}
