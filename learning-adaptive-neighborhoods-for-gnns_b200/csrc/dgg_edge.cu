// class DGG edge ranker + degree estimator + soft first-k, forward and backward.
// Replaces dgm.py:1781-1810 (gather u,v -> Linear+LeakyReLU -> sum -> sigmoid -> dense scatter ->
// row sum -> Linear(1,1) -> N-long row sort -> tanh first-k -> un-sort scatter -> to_sparse) with one
// warp-per-CSR-row kernel per direction.  HBM-bound: E*(H*4 gathered + ~24) bytes.
#include "common.cuh"

namespace dggb {

constexpr int kEdgeWarps = 8;  // warps per block

// Lane layout: the 32 lanes are split into G = 32/L groups of L lanes; a group owns one edge at a
// time and lane `lg` of the group owns float4 chunks c = 4*(lg + L*t), t < T, of the H-long row.
template <int T>
struct RowSlice {
  float4 v[T];
};

template <int T>
__device__ __forceinline__ void load_slice(RowSlice<T>& s, const float* row, int h, int lg, int L) {
#pragma unroll
  for (int t = 0; t < T; ++t) {
    int c = 4 * (lg + L * t);
    s.v[t] = (c < h) ? ldg4(row + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// z partial = sum over this lane's chunks of LeakyReLU(yb_i - y_v); out-of-range chunks are 0 - 0 = 0.
template <int T>
__device__ __forceinline__ float edge_partial(const RowSlice<T>& yb, const RowSlice<T>& yv) {
  float z = 0.f;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    z += leaky(yb.v[t].x - yv.v[t].x) + leaky(yb.v[t].y - yv.v[t].y) + leaky(yb.v[t].z - yv.v[t].z) +
         leaky(yb.v[t].w - yv.v[t].w);
  }
  return z;
}

// 0-based descending rank of element m inside R[beg, beg+deg): ties broken by lower position first.
// All 32 lanes must call this together (shuffles); lanes with m >= deg get garbage they ignore.
// R was written earlier by this same warp in this launch: read it through L2 (__ldcg), never through
// the non-coherent read-only path.
__device__ __forceinline__ int warp_rank(const float* R, int beg, int deg, int m, int lane) {
  const float mine = (m < deg) ? __ldcg(R + beg + m) : -INFINITY;
  int cnt = 0;
  for (int jb = 0; jb < deg; jb += kWarp) {
    const int jm = jb + lane;
    const float other = (jm < deg) ? __ldcg(R + beg + jm) : -INFINITY;
    const int lim = min(kWarp, deg - jb);
    for (int jj = 0; jj < lim; ++jj) {
      const float rj = __shfl_sync(0xffffffffu, other, jj);
      const int j = jb + jj;
      cnt += (rj > mine) || (rj == mine && j < m);
    }
  }
  return cnt;
}

__device__ __forceinline__ float first_k_plus_one(float r, float k) {
  // dgm.py:1801-1804: 1 - 0.5 * (1 + tanh(t - k)), then + 1.0
  float fk = 1.f - 0.5f * (1.f + tanhf(r - k));
  return fk + 1.f;
}

template <int T>
__global__ void __launch_bounds__(kEdgeWarps* kWarp)
    dgg_edge_fwd_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int n, int h, int L,
                        const float* __restrict__ y, const float* __restrict__ be, const float* __restrict__ deg_w,
                        const float* __restrict__ deg_b, const float* __restrict__ abl_noise, int hard_k,
                        float* __restrict__ R, int32_t* __restrict__ rank, float* __restrict__ s_out,
                        float* __restrict__ k_out, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int G = kWarp / L;
  const int lg = lane % L, grp = lane / L;
  RowSlice<T> bias;
  load_slice<T>(bias, be, h, lg, L);
  const float w = __ldg(deg_w), b = __ldg(deg_b);

  for (int i = blockIdx.x * kEdgeWarps + warp; i < n; i += gridDim.x * kEdgeWarps) {
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1), deg = end - beg;
    RowSlice<T> yb;
    load_slice<T>(yb, y + (size_t)i * h, h, lg, L);
#pragma unroll
    for (int t = 0; t < T; ++t) {
      yb.v[t].x += bias.v[t].x; yb.v[t].y += bias.v[t].y; yb.v[t].z += bias.v[t].z; yb.v[t].w += bias.v[t].w;
    }
    // ---- phase A: per-edge score ----
    float s_acc = 0.f;
    for (int e0 = beg; e0 < end; e0 += G) {
      const int e = e0 + grp;
      const bool valid = e < end;
      const int v = valid ? __ldg(col + e) : i;
      RowSlice<T> yv;
      load_slice<T>(yv, y + (size_t)v * h, h, lg, L);
      float z = group_sum(edge_partial<T>(yb, yv), L);
      float r = sigmoidf_(z);
      if (abl_noise != nullptr && valid) r = sigmoidf_(r + __ldg(abl_noise + e));  // dgm.py:1933-1935
      if (valid && lg == 0) {
        R[e] = r;
        s_acc += r;
      }
    }
    const float s = warp_sum(s_acc);
    __syncwarp();  // R[beg,end) written by this warp is now visible to all of its lanes
    // ---- phase B/C: rank, degree, first-k ----
    const float k = leaky(w * s + b);  // dgm.py:1791-1792
    for (int mb = 0; mb < deg; mb += kWarp) {
      const int m = mb + lane;
      const int r = warp_rank(R, beg, deg, m, lane);
      if (m < deg) {
        const float val = __ldcg(R + beg + m);
        rank[beg + m] = r;
        out[beg + m] = (hard_k >= 0) ? (r < hard_k ? val : 0.f) : val * first_k_plus_one((float)r, k);
      }
    }
    if (lane == 0) {
      s_out[i] = s;
      k_out[i] = k;
    }
  }
}

template <int T>
__global__ void __launch_bounds__(kEdgeWarps* kWarp)
    dgg_edge_bwd_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int n, int h, int L,
                        const float* __restrict__ y, const float* __restrict__ be, const float* __restrict__ deg_w,
                        const float* __restrict__ deg_b, const float* __restrict__ abl_noise, int hard_k,
                        const float* __restrict__ R, const int32_t* __restrict__ rank, const float* __restrict__ s_in,
                        const float* __restrict__ k_in, const float* __restrict__ g_out, float* __restrict__ dy,
                        float* __restrict__ dbe, float* __restrict__ ddeg) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int G = kWarp / L;
  const int lg = lane % L, grp = lane / L;
  RowSlice<T> bias;
  load_slice<T>(bias, be, h, lg, L);
  const float w = __ldg(deg_w), b = __ldg(deg_b);
  RowSlice<T> dbe_acc;  // this warp's running column sums of d(pre) over every row it owns
#pragma unroll
  for (int t = 0; t < T; ++t) dbe_acc.v[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  float dw_acc = 0.f, db_acc = 0.f;

  for (int i = blockIdx.x * kEdgeWarps + warp; i < n; i += gridDim.x * kEdgeWarps) {
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
    const float s = __ldg(s_in + i), k = __ldg(k_in + i);
    // ---- d out / d k  (A.1: 0.5 * sum g R sech^2(r - k)) and the degree-decoder chain ----
    float ds = 0.f;
    if (hard_k < 0) {
      float dk = 0.f;
      for (int e = beg + lane; e < end; e += kWarp) {
        const float th = tanhf((float)__ldg(rank + e) - k);
        dk += __ldg(g_out + e) * __ldg(R + e) * 0.5f * (1.f - th * th);
      }
      dk = warp_sum(dk);
      const float lr = leaky_grad(w * s + b);
      ds = dk * lr * w;
      dw_acc += dk * lr * s;
      db_acc += dk * lr;
    }
    RowSlice<T> yb, acc;
    load_slice<T>(yb, y + (size_t)i * h, h, lg, L);
#pragma unroll
    for (int t = 0; t < T; ++t) {
      yb.v[t].x += bias.v[t].x; yb.v[t].y += bias.v[t].y; yb.v[t].z += bias.v[t].z; yb.v[t].w += bias.v[t].w;
      acc.v[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int e0 = beg; e0 < end; e0 += G) {
      const int e = e0 + grp;
      const bool valid = e < end;
      const int v = valid ? __ldg(col + e) : i;
      RowSlice<T> yv;
      load_slice<T>(yv, y + (size_t)v * h, h, lg, L);
      const float z = group_sum(edge_partial<T>(yb, yv), L);
      const float r1 = sigmoidf_(z);
      float dz = 0.f;
      if (valid) {
        const float r_out = __ldg(R + e);
        const float g = __ldg(g_out + e);
        float dr;
        if (hard_k >= 0) dr = (__ldg(rank + e) < hard_k) ? g : 0.f;
        else dr = g * first_k_plus_one((float)__ldg(rank + e), k) + ds;
        if (abl_noise != nullptr) dr *= r_out * (1.f - r_out);  // through the second sigmoid
        dz = dr * r1 * (1.f - r1);
      }
      float* dyv = dy + (size_t)v * h;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int c = 4 * (lg + L * t);
        float4 d;
        d.x = dz * leaky_grad(yb.v[t].x - yv.v[t].x);
        d.y = dz * leaky_grad(yb.v[t].y - yv.v[t].y);
        d.z = dz * leaky_grad(yb.v[t].z - yv.v[t].z);
        d.w = dz * leaky_grad(yb.v[t].w - yv.v[t].w);
        acc.v[t].x += d.x; acc.v[t].y += d.y; acc.v[t].z += d.z; acc.v[t].w += d.w;
        if (valid && c < h) red_add4(dyv + c, make_float4(-d.x, -d.y, -d.z, -d.w));
      }
    }
    // combine the G groups' partial sums; group 0 then owns the row total
#pragma unroll
    for (int t = 0; t < T; ++t) {
      for (int o = L; o < kWarp; o <<= 1) {
        acc.v[t].x += __shfl_xor_sync(0xffffffffu, acc.v[t].x, o);
        acc.v[t].y += __shfl_xor_sync(0xffffffffu, acc.v[t].y, o);
        acc.v[t].z += __shfl_xor_sync(0xffffffffu, acc.v[t].z, o);
        acc.v[t].w += __shfl_xor_sync(0xffffffffu, acc.v[t].w, o);
      }
      dbe_acc.v[t].x += acc.v[t].x; dbe_acc.v[t].y += acc.v[t].y;
      dbe_acc.v[t].z += acc.v[t].z; dbe_acc.v[t].w += acc.v[t].w;
    }
    if (grp == 0) {
      // u-side total of the row (a self-loop edge adds +d here and -d above: they cancel in dy but
      // its +d still belongs in dbe, which is the column sum of the u-side totals)
      float* dyi = dy + (size_t)i * h;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int c = 4 * (lg + L * t);
        if (c < h) red_add4(dyi + c, acc.v[t]);
      }
    }
  }
  // ---- flush per-warp accumulators ----
  if (grp == 0) {
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int c = 4 * (lg + L * t);
      if (c < h) red_add4(dbe + c, dbe_acc.v[t]);
    }
  }
  if (lane == 0 && hard_k < 0) {
    atomicAdd(ddeg + 0, dw_acc);
    atomicAdd(ddeg + 1, db_acc);
  }
}

template <typename F>
static int dispatch_T(int h, int L, F&& f) {
  const int T = (h + 4 * L - 1) / (4 * L);
  if (T == 1) return f(std::integral_constant<int, 1>{});
  if (T == 2) return f(std::integral_constant<int, 2>{});
  if (T <= 4) return f(std::integral_constant<int, 4>{});
  return DGGB_ERR_BAD_SHAPE;
}

}  // namespace dggb

using namespace dggb;

extern "C" int dggb_dgg_edge_fwd(const int32_t* rowptr, const int32_t* col, int32_t n, int32_t h, const float* y,
                                 const float* be, const float* deg_w, const float* deg_b, const float* ablation_noise,
                                 int32_t hard_k, float* R, int32_t* rank, float* s, float* k, float* out,
                                 void* stream) {
  if (!rowptr || !col || !y || !be || !deg_w || !deg_b || !R || !rank || !s || !k || !out || n < 0 || h <= 0)
    return DGGB_ERR_BAD_ARG;
  if (h % 4 != 0 || h > 512) return DGGB_ERR_BAD_SHAPE;
  if (n == 0) return DGGB_OK;
  const int L = pow2_floor32(h / 4);
  const int grid = rows_grid(n, kEdgeWarps, 8);
  return dispatch_T(h, L, [&](auto tc) {
    constexpr int T = decltype(tc)::value;
    dgg_edge_fwd_kernel<T><<<grid, kEdgeWarps * kWarp, 0, as_stream(stream)>>>(
        rowptr, col, n, h, L, y, be, deg_w, deg_b, ablation_noise, hard_k, R, rank, s, k, out);
    return launch_status();
  });
}

extern "C" int dggb_dgg_edge_bwd(const int32_t* rowptr, const int32_t* col, int32_t n, int32_t h, const float* y,
                                 const float* be, const float* deg_w, const float* deg_b, const float* ablation_noise,
                                 int32_t hard_k, const float* R, const int32_t* rank, const float* s, const float* k,
                                 const float* g_out, float* dy, float* dbe, float* ddeg, void* stream) {
  if (!rowptr || !col || !y || !be || !deg_w || !deg_b || !R || !rank || !s || !k || !g_out || !dy || !dbe ||
      !ddeg || n < 0 || h <= 0)
    return DGGB_ERR_BAD_ARG;
  if (h % 4 != 0 || h > 512) return DGGB_ERR_BAD_SHAPE;
  if (n == 0) return DGGB_OK;
  const int L = pow2_floor32(h / 4);
  const int grid = rows_grid(n, kEdgeWarps, 4);
  return dispatch_T(h, L, [&](auto tc) {
    constexpr int T = decltype(tc)::value;
    dgg_edge_bwd_kernel<T><<<grid, kEdgeWarps * kWarp, 0, as_stream(stream)>>>(
        rowptr, col, n, h, L, y, be, deg_w, deg_b, ablation_noise, hard_k, R, rank, s, k, g_out, dy, dbe, ddeg);
    return launch_status();
  });
}
