// class DGG edge ranker + degree estimator + soft first-k, forward and backward.
// Replaces dgm.py:1781-1810 (gather u,v -> Linear+LeakyReLU -> sum -> sigmoid -> dense N x N scatter ->
// row sum -> Linear(1,1) -> N-long row sort -> tanh first-k -> un-sort scatter -> to_sparse).
//
// Two launches per direction so that power-law hub rows cannot serialise a warp:
//   fwd:  edge_score (edge-parallel, perfectly balanced)  ->  row_rank (warp per CSR row)
//   bwd:  row_dk     (warp per CSR row)                    ->  edge_grad (edge-parallel, vector reds)
// HBM-bound.  Algorithmic bytes per launch are listed in DESIGN.md ("dgg_edge").
#include "common.cuh"

namespace dggb {

constexpr int kEdgeWarps = 8;  // warps per block

// Lane layout: the 32 lanes split into G = 32/L groups of L lanes.  A warp owns 32 consecutive edges;
// group `grp` walks the PER = 32/G consecutive edges [grp*PER, (grp+1)*PER) of that chunk, and lane `lg`
// of the group owns float4 chunks c = 4*(lg + L*t), t < T, of the H-long feature row.
template <int T>
struct RowSlice {
  float4 v[T];
};

template <int T>
__device__ __forceinline__ void load_slice(RowSlice<T>& s, const float* row, int h, int lg, int L) {
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int c = 4 * (lg + L * t);
    s.v[t] = (c < h) ? ldg4(row + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// sum over this lane's chunks of LeakyReLU(y_u - y_v + be); out-of-range chunks contribute exactly 0.
template <int T>
__device__ __forceinline__ float edge_partial(const RowSlice<T>& yu, const RowSlice<T>& yv, const RowSlice<T>& b) {
  float z = 0.f;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    z += leaky(yu.v[t].x - yv.v[t].x + b.v[t].x) + leaky(yu.v[t].y - yv.v[t].y + b.v[t].y) +
         leaky(yu.v[t].z - yv.v[t].z + b.v[t].z) + leaky(yu.v[t].w - yv.v[t].w + b.v[t].w);
  }
  return z;
}

__device__ __forceinline__ float first_k_plus_one(float r, float k) {
  // dgm.py:1801-1804: 1 - 0.5 * (1 + tanh(t - k)), then + 1.0
  const float fk = 1.f - 0.5f * (1.f + tanhf(r - k));
  return fk + 1.f;
}

// ------------------------------------------------------------------------------------------------
// fwd 1/2: R_e = sigmoid(sum_c LeakyReLU(y_u - y_v + be))  [ablation: sigmoid(R_e + noise_e)]
// ------------------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(kEdgeWarps* kWarp)
    dgg_edge_score_kernel(const int32_t* __restrict__ erow, const int32_t* __restrict__ col, int nnz, int h, int L,
                          const float* __restrict__ y, const float* __restrict__ be,
                          const float* __restrict__ abl_noise, float* __restrict__ R) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int G = kWarp / L, PER = kWarp / G;  // PER == L
  const int lg = lane % L, grp = lane / L;
  RowSlice<T> bias;
  load_slice<T>(bias, be, h, lg, L);
  const int warps_total = gridDim.x * kEdgeWarps;
  for (int base = (blockIdx.x * kEdgeWarps + (threadIdx.x >> 5)) * kWarp; base < nnz; base += warps_total * kWarp) {
    const int e_l = base + lane;
    const int u_l = (e_l < nnz) ? __ldg(erow + e_l) : 0;
    const int v_l = (e_l < nnz) ? __ldg(col + e_l) : 0;
    float mine = 0.f;
#pragma unroll 4
    for (int it = 0; it < PER; ++it) {
      const int j = grp * PER + it;  // edge slot inside the chunk handled by this group now
      const int u = __shfl_sync(0xffffffffu, u_l, j);
      const int v = __shfl_sync(0xffffffffu, v_l, j);
      RowSlice<T> yu, yv;
      load_slice<T>(yu, y + (size_t)u * h, h, lg, L);
      load_slice<T>(yv, y + (size_t)v * h, h, lg, L);
      const float z = group_sum(edge_partial<T>(yu, yv, bias), L);
      // group grp walks slots [grp*L, grp*L + L) == its own lanes: lane lg keeps iteration lg's result
      if (lg == it) mine = z;
    }
    if (e_l < nnz) {
      float r = sigmoidf_(mine);
      if (abl_noise != nullptr) r = sigmoidf_(r + __ldg(abl_noise + e_l));  // dgm.py:1933-1935
      R[e_l] = r;
    }
  }
}

// 0-based descending rank of element m inside R[beg, beg+deg): ties broken by lower position first.
// All 32 lanes call this together; lanes with m >= deg get garbage they ignore.
//   deg <= 32 : one register per lane + 32 shuffles
//   deg <= kRankCap : the row is staged in this warp's shared-memory slice; every lane then sweeps it with
//                broadcast LDS (independent iterations, throughput- not latency-bound)
//   longer rows: same sweep straight from global/L1
constexpr int kRankCap = 1024;  // floats of shared memory per warp

__device__ __forceinline__ int warp_rank(const float* __restrict__ R, float* srow, bool staged, int beg, int deg,
                                         int m, int lane) {
  const float mine = (m < deg) ? __ldg(R + beg + m) : -INFINITY;
  int cnt = 0;
  if (deg <= kWarp) {
    for (int jj = 0; jj < deg; ++jj) {
      const float rj = __shfl_sync(0xffffffffu, mine, jj);   // m == lane here
      cnt += (rj > mine) || (rj == mine && jj < m);
    }
    return cnt;
  }
  if (staged) {
#pragma unroll 4
    for (int j = 0; j < deg; ++j) {
      const float rj = srow[j];
      cnt += (rj > mine) || (rj == mine && j < m);
    }
  } else {
#pragma unroll 4
    for (int j = 0; j < deg; ++j) {
      const float rj = __ldg(R + beg + j);
      cnt += (rj > mine) || (rj == mine && j < m);
    }
  }
  return cnt;
}

// stage R[beg, beg+deg) into the warp's shared slice (deg <= kRankCap)
__device__ __forceinline__ bool stage_row(const float* __restrict__ R, float* srow, int beg, int deg, int lane) {
  if (deg <= kWarp || deg > kRankCap) return false;
  __syncwarp();
  for (int j = lane; j < deg; j += kWarp) srow[j] = __ldg(R + beg + j);
  __syncwarp();
  return true;
}

// ------------------------------------------------------------------------------------------------
// fwd 2/2: s_i, k_i = LeakyReLU(w s_i + b), in-row rank, out_e = R_e * (first_k + 1)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEdgeWarps* kWarp)
    dgg_row_rank_kernel(const int32_t* __restrict__ rowptr, int n, const float* __restrict__ R,
                        const float* __restrict__ deg_w, const float* __restrict__ deg_b, int hard_k,
                        int32_t* __restrict__ rank, float* __restrict__ s_out, float* __restrict__ k_out,
                        float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float srow_all[kEdgeWarps * kRankCap];
  float* srow = srow_all + (threadIdx.x >> 5) * kRankCap;
  const int lane = threadIdx.x & 31;
  const float w = __ldg(deg_w), b = __ldg(deg_b);
  for (int i = blockIdx.x * kEdgeWarps + (threadIdx.x >> 5); i < n; i += gridDim.x * kEdgeWarps) {
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1), deg = end - beg;
    float s = 0.f;
    for (int e = beg + lane; e < end; e += kWarp) s += __ldg(R + e);
    s = warp_sum(s);
    const float k = leaky(w * s + b);  // dgm.py:1791-1792
    const bool staged = stage_row(R, srow, beg, deg, lane);
    for (int mb = 0; mb < deg; mb += kWarp) {
      const int m = mb + lane;
      const int r = warp_rank(R, srow, staged, beg, deg, m, lane);
      if (m < deg) {
        const float val = __ldg(R + beg + m);
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

// ------------------------------------------------------------------------------------------------
// bwd 1/2: d out / d k_i = 0.5 * sum_e g_e R_e sech^2(r_e - k_i)  (SURVEY A.1), chained through the
// degree decoder: ds_i = dk_i * LeakyReLU'(w s_i + b) * w;  d w, d b accumulated per warp.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEdgeWarps* kWarp)
    dgg_row_dk_kernel(const int32_t* __restrict__ rowptr, int n, const float* __restrict__ R,
                      const int32_t* __restrict__ rank, const float* __restrict__ s_in,
                      const float* __restrict__ k_in, const float* __restrict__ g_out,
                      const float* __restrict__ deg_w, const float* __restrict__ deg_b, float* __restrict__ ds,
                      float* __restrict__ ddeg) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const float w = __ldg(deg_w), b = __ldg(deg_b);
  float dw_acc = 0.f, db_acc = 0.f;
  for (int i = blockIdx.x * kEdgeWarps + (threadIdx.x >> 5); i < n; i += gridDim.x * kEdgeWarps) {
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
    const float s = __ldg(s_in + i), k = __ldg(k_in + i);
    float dk = 0.f;
    for (int e = beg + lane; e < end; e += kWarp) {
      const float th = tanhf((float)__ldg(rank + e) - k);
      dk += __ldg(g_out + e) * __ldg(R + e) * 0.5f * (1.f - th * th);
    }
    dk = warp_sum(dk);
    const float lr = leaky_grad(w * s + b);
    if (lane == 0) ds[i] = dk * lr * w;
    dw_acc += dk * lr * s;
    db_acc += dk * lr;
  }
  // one pair of atomics per block, not per warp: every warp in the grid targets the same two addresses
  __shared__ float red[2][kEdgeWarps];
  if (lane == 0) {
    red[0][threadIdx.x >> 5] = dw_acc;
    red[1][threadIdx.x >> 5] = db_acc;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < kEdgeWarps; ++wv) t += red[threadIdx.x][wv];
    atomicAdd(ddeg + threadIdx.x, t);
  }
}

// ------------------------------------------------------------------------------------------------
// bwd 2/2: per edge, recompute z (no E x H tensor is ever stored), then
//   dR = g * (first_k + 1) + ds_u  [hard: g * (rank < hard_k)]   [ablation: *= R2 (1 - R2)]
//   dz = dR * R1 (1 - R1);  d pre_c = dz * LeakyReLU'(pre_c);  dy_u += d pre;  dy_v -= d pre;  dbe += d pre
// ------------------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(kEdgeWarps* kWarp)
    dgg_edge_grad_kernel(const int32_t* __restrict__ erow, const int32_t* __restrict__ col, int nnz, int h, int L,
                         const float* __restrict__ y, const float* __restrict__ be,
                         const float* __restrict__ abl_noise, int hard_k, const float* __restrict__ R,
                         const int32_t* __restrict__ rank, const float* __restrict__ k_in,
                         const float* __restrict__ ds, const float* __restrict__ g_out, float* __restrict__ dy,
                         float* __restrict__ dbe) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int G = kWarp / L, PER = kWarp / G;
  const int lg = lane % L, grp = lane / L;
  RowSlice<T> bias, dbe_acc;
  load_slice<T>(bias, be, h, lg, L);
#pragma unroll
  for (int t = 0; t < T; ++t) dbe_acc.v[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  __shared__ float dbe_s[512];   // block-level bias-gradient accumulator (h <= 512)
  for (int c = threadIdx.x; c < h; c += blockDim.x) dbe_s[c] = 0.f;
  __syncthreads();
  const int warps_total = gridDim.x * kEdgeWarps;
  for (int base = (blockIdx.x * kEdgeWarps + (threadIdx.x >> 5)) * kWarp; base < nnz; base += warps_total * kWarp) {
    const int e_l = base + lane;
    const bool ok_l = e_l < nnz;
    const int u_l = ok_l ? __ldg(erow + e_l) : 0;
    const int v_l = ok_l ? __ldg(col + e_l) : 0;
    // per-edge scalar factor dR (everything that does not need the feature rows), one edge per lane
    float dr_l = 0.f;
    if (ok_l) {
      const float g = __ldg(g_out + e_l);
      if (hard_k >= 0) {
        dr_l = (__ldg(rank + e_l) < hard_k) ? g : 0.f;
      } else {
        dr_l = g * first_k_plus_one((float)__ldg(rank + e_l), __ldg(k_in + u_l)) + __ldg(ds + u_l);
      }
      if (abl_noise != nullptr) {
        const float r2 = __ldg(R + e_l);
        dr_l *= r2 * (1.f - r2);  // through the second sigmoid
      }
    }
    int cur_u = -1;
    RowSlice<T> acc;  // running +d pre for the current source row of this group
#pragma unroll
    for (int t = 0; t < T; ++t) acc.v[t] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int it = 0; it < PER; ++it) {
      const int j = grp * PER + it;
      const int u = __shfl_sync(0xffffffffu, u_l, j);
      const int v = __shfl_sync(0xffffffffu, v_l, j);
      const float dr = __shfl_sync(0xffffffffu, dr_l, j);
      const bool valid = (base + j) < nnz;
      RowSlice<T> yu, yv;
      load_slice<T>(yu, y + (size_t)u * h, h, lg, L);
      load_slice<T>(yv, y + (size_t)v * h, h, lg, L);
      const float z = group_sum(edge_partial<T>(yu, yv, bias), L);
      const float r1 = sigmoidf_(z);
      const float dz = valid ? dr * r1 * (1.f - r1) : 0.f;
      if (u != cur_u) {  // group-uniform branch: flush the finished source row
        if (cur_u >= 0) {
#pragma unroll
          for (int t = 0; t < T; ++t) {
            const int c = 4 * (lg + L * t);
            if (c < h) red_add4(dy + (size_t)cur_u * h + c, acc.v[t]);
            acc.v[t] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        cur_u = u;
      }
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int c = 4 * (lg + L * t);
        float4 d;
        d.x = dz * leaky_grad(yu.v[t].x - yv.v[t].x + bias.v[t].x);
        d.y = dz * leaky_grad(yu.v[t].y - yv.v[t].y + bias.v[t].y);
        d.z = dz * leaky_grad(yu.v[t].z - yv.v[t].z + bias.v[t].z);
        d.w = dz * leaky_grad(yu.v[t].w - yv.v[t].w + bias.v[t].w);
        dbe_acc.v[t].x += d.x; dbe_acc.v[t].y += d.y; dbe_acc.v[t].z += d.z; dbe_acc.v[t].w += d.w;
        if (u != v) {  // a self loop adds +d and -d to the same row: skip both (it still counts for dbe)
          acc.v[t].x += d.x; acc.v[t].y += d.y; acc.v[t].z += d.z; acc.v[t].w += d.w;
          if (valid && c < h) red_add4(dy + (size_t)v * h + c, make_float4(-d.x, -d.y, -d.z, -d.w));
        }
      }
    }
    if (cur_u >= 0) {
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int c = 4 * (lg + L * t);
        if (c < h) red_add4(dy + (size_t)cur_u * h + c, acc.v[t]);
      }
    }
  }
  // flush this warp's bias gradient: combine the G groups first, then one vector red per chunk
#pragma unroll
  for (int t = 0; t < T; ++t) {
    for (int o = L; o < kWarp; o <<= 1) {
      dbe_acc.v[t].x += __shfl_xor_sync(0xffffffffu, dbe_acc.v[t].x, o);
      dbe_acc.v[t].y += __shfl_xor_sync(0xffffffffu, dbe_acc.v[t].y, o);
      dbe_acc.v[t].z += __shfl_xor_sync(0xffffffffu, dbe_acc.v[t].z, o);
      dbe_acc.v[t].w += __shfl_xor_sync(0xffffffffu, dbe_acc.v[t].w, o);
    }
    const int c = 4 * (lg + L * t);
    if (grp == 0 && c < h) {
      atomicAdd(&dbe_s[c + 0], dbe_acc.v[t].x);
      atomicAdd(&dbe_s[c + 1], dbe_acc.v[t].y);
      atomicAdd(&dbe_s[c + 2], dbe_acc.v[t].z);
      atomicAdd(&dbe_s[c + 3], dbe_acc.v[t].w);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < h; c += blockDim.x) atomicAdd(dbe + c, dbe_s[c]);
}

// ------------------------------------------------------------------------------------------------
// select_top_k of DGG_LearnableK_debug on CSR rows (dgm.py:1402-1421, A.2): k is an input,
//   fk(r) = 1 - 0.5 * (1 + tanh(r - k_i));  mode 0: out = score * fk (k_times_edge_prob)
// Off-support entries are exact zeros that sort after every positive score, so in-row ranks are exact.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float first_k_tanh(float r, float k) { return 1.f - 0.5f * (1.f + tanhf(r - k)); }

__global__ void __launch_bounds__(kEdgeWarps* kWarp)
    row_firstk_fwd_kernel(const int32_t* __restrict__ rowptr, int n, const float* __restrict__ score,
                          const float* __restrict__ k_in, int32_t* __restrict__ rank, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float srow_all[kEdgeWarps * kRankCap];
  float* srow = srow_all + (threadIdx.x >> 5) * kRankCap;
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * kEdgeWarps + (threadIdx.x >> 5); i < n; i += gridDim.x * kEdgeWarps) {
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1), deg = end - beg;
    const float k = __ldg(k_in + i);
    const bool staged = stage_row(score, srow, beg, deg, lane);
    for (int mb = 0; mb < deg; mb += kWarp) {
      const int m = mb + lane;
      const int r = warp_rank(score, srow, staged, beg, deg, m, lane);
      if (m < deg) {
        rank[beg + m] = r;
        out[beg + m] = __ldg(score + beg + m) * first_k_tanh((float)r, k);
      }
    }
  }
}

__global__ void __launch_bounds__(kEdgeWarps* kWarp)
    row_firstk_bwd_kernel(const int32_t* __restrict__ rowptr, int n, const float* __restrict__ score,
                          const float* __restrict__ k_in, const int32_t* __restrict__ rank,
                          const float* __restrict__ g_out, float* __restrict__ dscore, float* __restrict__ dk) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * kEdgeWarps + (threadIdx.x >> 5); i < n; i += gridDim.x * kEdgeWarps) {
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
    const float k = __ldg(k_in + i);
    float acc = 0.f;
    for (int e = beg + lane; e < end; e += kWarp) {
      const float th = tanhf((float)__ldg(rank + e) - k);
      const float g = __ldg(g_out + e);
      dscore[e] = g * (1.f - 0.5f * (1.f + th));
      acc += g * __ldg(score + e) * 0.5f * (1.f - th * th);
    }
    acc = warp_sum(acc);
    if (lane == 0) dk[i] = acc;
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

static int edges_grid(long long nnz) {
  long long need = (nnz + kEdgeWarps * kWarp - 1) / (kEdgeWarps * kWarp);
  long long cap = (long long)kNumSMs * 8;
  long long g = need < cap ? need : cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace dggb

using namespace dggb;

extern "C" int dggb_dgg_edge_fwd(const int32_t* rowptr, const int32_t* erow, const int32_t* col, int32_t n,
                                 int32_t nnz, int32_t h, const float* y, const float* be, const float* deg_w,
                                 const float* deg_b, const float* ablation_noise, int32_t hard_k, float* R,
                                 int32_t* rank, float* s, float* k, float* out, void* stream) {
  if (!rowptr || !erow || !col || !y || !be || !deg_w || !deg_b || !R || !rank || !s || !k || !out || n < 0 ||
      nnz < 0 || h <= 0)
    return DGGB_ERR_BAD_ARG;
  if (h % 4 != 0 || h > 512) return DGGB_ERR_BAD_SHAPE;
  if (n == 0) return DGGB_OK;
  const int L = pow2_floor32(h / 4);
  int st = DGGB_OK;
  if (nnz > 0) {
    st = dispatch_T(h, L, [&](auto tc) {
      constexpr int T = decltype(tc)::value;
      launch_pdl(dgg_edge_score_kernel<T>, dim3(edges_grid(nnz)), dim3(kEdgeWarps * kWarp), 0, as_stream(stream), 
          erow, col, nnz, h, L, y, be, ablation_noise, R);
      return launch_status();
    });
    if (st != DGGB_OK) return st;
  }
  launch_pdl(dgg_row_rank_kernel, dim3(rows_grid(n, kEdgeWarps, 8)), dim3(kEdgeWarps * kWarp), 0, as_stream(stream), 
      rowptr, n, R, deg_w, deg_b, hard_k, rank, s, k, out);
  return launch_status();
}

extern "C" int dggb_dgg_edge_bwd(const int32_t* rowptr, const int32_t* erow, const int32_t* col, int32_t n,
                                 int32_t nnz, int32_t h, const float* y, const float* be, const float* deg_w,
                                 const float* deg_b, const float* ablation_noise, int32_t hard_k, const float* R,
                                 const int32_t* rank, const float* s, const float* k, const float* g_out,
                                 float* ds_ws, float* dy, float* dbe, float* ddeg, void* stream) {
  if (!rowptr || !erow || !col || !y || !be || !deg_w || !deg_b || !R || !rank || !s || !k || !g_out || !ds_ws ||
      !dy || !dbe || !ddeg || n < 0 || nnz < 0 || h <= 0)
    return DGGB_ERR_BAD_ARG;
  if (h % 4 != 0 || h > 512) return DGGB_ERR_BAD_SHAPE;
  if (n == 0 || nnz == 0) return DGGB_OK;
  const int L = pow2_floor32(h / 4);
  if (hard_k < 0) {
    launch_pdl(dgg_row_dk_kernel, dim3(rows_grid(n, kEdgeWarps, 8)), dim3(kEdgeWarps * kWarp), 0, as_stream(stream), 
        rowptr, n, R, rank, s, k, g_out, deg_w, deg_b, ds_ws, ddeg);
    const int st = launch_status();
    if (st != DGGB_OK) return st;
  }
  return dispatch_T(h, L, [&](auto tc) {
    constexpr int T = decltype(tc)::value;
    launch_pdl(dgg_edge_grad_kernel<T>, dim3(edges_grid(nnz)), dim3(kEdgeWarps * kWarp), 0, as_stream(stream), 
        erow, col, nnz, h, L, y, be, ablation_noise, hard_k, R, rank, k, ds_ws, g_out, dy, dbe);
    return launch_status();
  });
}

extern "C" int dggb_row_firstk_fwd(const int32_t* rowptr, int32_t n, const float* score, const float* k,
                                   int32_t* rank, float* out, void* stream) {
  if (!rowptr || !score || !k || !rank || !out || n < 0) return DGGB_ERR_BAD_ARG;
  if (n == 0) return DGGB_OK;
  launch_pdl(row_firstk_fwd_kernel, dim3(rows_grid(n, kEdgeWarps, 8)), dim3(kEdgeWarps * kWarp), 0, as_stream(stream), rowptr, n, score,
                                                                                                 k, rank, out);
  return launch_status();
}

extern "C" int dggb_row_firstk_bwd(const int32_t* rowptr, int32_t n, const float* score, const float* k,
                                   const int32_t* rank, const float* g_out, float* dscore, float* dk,
                                   void* stream) {
  if (!rowptr || !score || !k || !rank || !g_out || !dscore || !dk || n < 0) return DGGB_ERR_BAD_ARG;
  if (n == 0) return DGGB_OK;
  launch_pdl(row_firstk_bwd_kernel, dim3(rows_grid(n, kEdgeWarps, 8)), dim3(kEdgeWarps * kWarp), 0, as_stream(stream), 
      rowptr, n, score, k, rank, g_out, dscore, dk);
  return launch_status();
}
