// normalize_adj: Ahat_ij = A_ij * s_i^-1/2 * s_j^-1/2, s = row sums on BOTH sides
// (model.py:1215-1218).  O(nnz) instead of diag + two dense N^3 torch.mm.
// HBM-bound: nnz*(4 col + 4 val + 4 out) + 8N bytes.
#include "common.cuh"

namespace dggb {

constexpr int kNormWarps = 8;

__global__ void __launch_bounds__(kNormWarps* kWarp)
    rowsum_rsqrt_kernel(const int32_t* __restrict__ rowptr, const float* __restrict__ val, int n,
                        float* __restrict__ dinv) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * kNormWarps + (threadIdx.x >> 5); i < n; i += gridDim.x * kNormWarps) {
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
    float s = 0.f;
    for (int e = beg + lane; e < end; e += kWarp) s += __ldg(val + e);
    s = warp_sum(s);
    if (lane == 0) dinv[i] = 1.0f / sqrtf(s);  // row_sum ** -0.5, no zero guard (model.py:1216)
  }
}

__global__ void __launch_bounds__(kNormWarps* kWarp)
    sym_scale_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                     const float* __restrict__ val, int n, const float* __restrict__ dinv, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * kNormWarps + (threadIdx.x >> 5); i < n; i += gridDim.x * kNormWarps) {
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
    const float ai = dinv[i];
    for (int e = beg + lane; e < end; e += kWarp) out[e] = (ai * __ldg(val + e)) * dinv[__ldg(col + e)];
  }
}

// T_i = sum over entries with row OR column i of g * val * a_other
__global__ void __launch_bounds__(kNormWarps* kWarp)
    sym_bwd_t_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                     const float* __restrict__ val, int n, const float* __restrict__ dinv,
                     const float* __restrict__ g, float* __restrict__ t_ws) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * kNormWarps + (threadIdx.x >> 5); i < n; i += gridDim.x * kNormWarps) {
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
    const float ai = __ldg(dinv + i);
    float rowpart = 0.f;
    for (int e = beg + lane; e < end; e += kWarp) {
      const int j = __ldg(col + e);
      const float p = __ldg(g + e) * __ldg(val + e);
      rowpart += p * __ldg(dinv + j);
      atomicAdd(t_ws + j, p * ai);
    }
    rowpart = warp_sum(rowpart);
    if (lane == 0) atomicAdd(t_ws + i, rowpart);
  }
}

__global__ void __launch_bounds__(kNormWarps* kWarp)
    sym_bwd_apply_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int n,
                         const float* __restrict__ dinv, const float* __restrict__ g,
                         const float* __restrict__ t_ws, float* __restrict__ dval) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * kNormWarps + (threadIdx.x >> 5); i < n; i += gridDim.x * kNormWarps) {
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
    const float ai = __ldg(dinv + i);
    const float corr = -0.5f * ai * ai * ai * __ldg(t_ws + i);  // d s_i^-1/2 / d s_i = -1/2 s_i^-3/2
    for (int e = beg + lane; e < end; e += kWarp) dval[e] = __ldg(g + e) * ai * __ldg(dinv + __ldg(col + e)) + corr;
  }
}

}  // namespace dggb
using namespace dggb;

extern "C" int dggb_sym_normalize_fwd(const int32_t* rowptr, const int32_t* col, const float* val, int32_t n,
                                      float* dinv, float* out, void* stream) {
  if (!rowptr || !col || !val || !dinv || !out || n < 0) return DGGB_ERR_BAD_ARG;
  if (n == 0) return DGGB_OK;
  const int grid = rows_grid(n, kNormWarps, 8);
  launch_pdl(rowsum_rsqrt_kernel, dim3(grid), dim3(kNormWarps * kWarp), 0, as_stream(stream), rowptr, val, n, dinv);
  launch_pdl(sym_scale_kernel, dim3(grid), dim3(kNormWarps * kWarp), 0, as_stream(stream), rowptr, col, val, n, dinv, out);
  return launch_status(2);
}

extern "C" int dggb_sym_normalize_bwd(const int32_t* rowptr, const int32_t* col, const float* val, int32_t n,
                                      const float* dinv, const float* g, float* t_ws, float* dval, void* stream) {
  if (!rowptr || !col || !val || !dinv || !g || !t_ws || !dval || n < 0) return DGGB_ERR_BAD_ARG;
  if (n == 0) return DGGB_OK;
  const int grid = rows_grid(n, kNormWarps, 8);
  launch_pdl(sym_bwd_t_kernel, dim3(grid), dim3(kNormWarps * kWarp), 0, as_stream(stream), rowptr, col, val, n, dinv, g, t_ws);
  launch_pdl(sym_bwd_apply_kernel, dim3(grid), dim3(kNormWarps * kWarp), 0, as_stream(stream), rowptr, col, n, dinv, g, t_ws, dval);
  return launch_status(2);
}
