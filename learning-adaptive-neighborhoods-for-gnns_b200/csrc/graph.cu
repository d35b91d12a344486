// Graph plumbing: COO(int64, coalesced) -> int32 CSR, and A + I without the dense round trip of
// (in_adj.to_dense() + eye).to_sparse() (model.py:1249-1251, 171-173, 389-391, 710-712, 1381-1383).
#include "common.cuh"

namespace dggb {

// rowptr[r] = first position p with row[p] >= r  (row sorted ascending)
__global__ void rowptr_from_sorted_rows(const int64_t* __restrict__ row, long long nnz, int n,
                                        int32_t* __restrict__ rowptr) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p <= nnz;
       p += (long long)gridDim.x * blockDim.x) {
    const long long prev = (p == 0) ? -1 : row[p - 1];
    const long long cur = (p == nnz) ? n : row[p];
    for (long long r = prev + 1; r <= cur; ++r) rowptr[r] = (int32_t)p;
  }
}

__global__ void cast_i64_i32(const int64_t* __restrict__ src, int32_t* __restrict__ dst, long long n) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x)
    dst[p] = (int32_t)src[p];
}

// erow[e] = i for every e in [rowptr[i], rowptr[i+1])
__global__ void csr_expand_rows(const int32_t* __restrict__ rowptr, int n, int32_t* __restrict__ erow) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
    for (int e = beg + lane; e < end; e += kWarp) erow[e] = i;
  }
}

// out_count[i] = deg_i + (row i has no diagonal entry); out_count[n] = 0 (slot for the scan total)
__global__ void self_loop_count(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int n,
                                int32_t* __restrict__ out_count) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i <= n; i += gridDim.x * wpb) {
    if (i == n) {
      if (lane == 0) out_count[n] = 0;
      continue;
    }
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
    int has = 0;
    for (int e = beg + lane; e < end; e += kWarp) has |= (__ldg(col + e) == i);
    has = __any_sync(0xffffffffu, has);
    if (lane == 0) out_count[i] = (end - beg) + (has ? 0 : 1);
  }
}

__global__ void self_loop_fill(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                               const float* __restrict__ val, int n, const int32_t* __restrict__ out_rowptr,
                               int32_t* __restrict__ out_col, float* __restrict__ out_val) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
    const int obeg = __ldg(out_rowptr + i), oend = __ldg(out_rowptr + i + 1);
    const bool insert = (oend - obeg) != (end - beg);
    // number of existing columns < i == position of the diagonal in the output row
    int lt = 0;
    for (int e = beg + lane; e < end; e += kWarp) lt += (__ldg(col + e) < i);
    for (int o = 16; o > 0; o >>= 1) lt += __shfl_xor_sync(0xffffffffu, lt, o);
    for (int e = beg + lane; e < end; e += kWarp) {
      const int c = __ldg(col + e);
      const int pos = obeg + (e - beg) + ((insert && c > i) ? 1 : 0);
      out_col[pos] = c;
      out_val[pos] = __ldg(val + e) + ((c == i) ? 1.f : 0.f);
    }
    if (insert && lane == 0) {
      out_col[obeg + lt] = i;
      out_val[obeg + lt] = 1.f;
    }
  }
}

// single-block exclusive scan, in place over n+1 ints (graph sizes here: N+1 <= a few million)
__global__ void __launch_bounds__(1024) exclusive_scan_inplace(int32_t* data, int count) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < count; base += 1024) {
    const int idx = base + threadIdx.x;
    const int32_t v = idx < count ? data[idx] : 0;
    int32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      int32_t w = warp_tot[threadIdx.x];
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t t = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += t;
      }
      warp_tot[threadIdx.x] = w;
    }
    __syncthreads();
    const int32_t wprefix = (threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0;
    const int32_t c = carry;
    if (idx < count) data[idx] = c + wprefix + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + wprefix + inc;
    __syncthreads();
  }
}

}  // namespace dggb
using namespace dggb;

extern "C" int dggb_coo_rows_to_rowptr(const int64_t* row, int64_t nnz, int32_t n, int32_t* rowptr, void* stream) {
  if (!rowptr || nnz < 0 || n < 0 || (nnz > 0 && !row)) return DGGB_ERR_BAD_ARG;
  const int block = 256;
  long long g = (nnz + 1 + block - 1) / block;
  if (g > kNumSMs * 8) g = kNumSMs * 8;
  rowptr_from_sorted_rows<<<(int)g, block, 0, as_stream(stream)>>>(row, nnz, n, rowptr);
  return launch_status();
}

extern "C" int dggb_cast_i64_i32(const int64_t* src, int32_t* dst, int64_t n, void* stream) {
  if (n < 0 || (n > 0 && (!src || !dst))) return DGGB_ERR_BAD_ARG;
  if (n == 0) return DGGB_OK;
  const int block = 256;
  long long g = (n + block - 1) / block;
  if (g > kNumSMs * 8) g = kNumSMs * 8;
  cast_i64_i32<<<(int)g, block, 0, as_stream(stream)>>>(src, dst, n);
  return launch_status();
}

extern "C" int dggb_csr_expand_rows(const int32_t* rowptr, int32_t n, int32_t* erow, void* stream) {
  if (!rowptr || !erow || n < 0) return DGGB_ERR_BAD_ARG;
  if (n == 0) return DGGB_OK;
  launch_pdl(csr_expand_rows, dim3(rows_grid(n, 8, 8)), dim3(256), 0, as_stream(stream), rowptr, n, erow);
  return launch_status();
}

extern "C" int dggb_add_self_loops_count(const int32_t* rowptr, const int32_t* col, int32_t n,
                                         int32_t* out_rowcount, void* stream) {
  if (!rowptr || !out_rowcount || n < 0) return DGGB_ERR_BAD_ARG;
  const int grid = rows_grid(n + 1, 8, 8);
  self_loop_count<<<grid, 256, 0, as_stream(stream)>>>(rowptr, col, n, out_rowcount);
  exclusive_scan_inplace<<<1, 1024, 0, as_stream(stream)>>>(out_rowcount, n + 1);
  return launch_status(2);
}

extern "C" int dggb_add_self_loops_fill(const int32_t* rowptr, const int32_t* col, const float* val, int32_t n,
                                        const int32_t* out_rowptr, int32_t* out_col, float* out_val, void* stream) {
  if (!rowptr || !out_rowptr || !out_col || !out_val || n < 0) return DGGB_ERR_BAD_ARG;
  if (n == 0) return DGGB_OK;
  const int grid = rows_grid(n, 8, 8);
  launch_pdl(self_loop_fill, dim3(grid), dim3(256), 0, as_stream(stream), rowptr, col, val, n, out_rowptr, out_col, out_val);
  return launch_status();
}
