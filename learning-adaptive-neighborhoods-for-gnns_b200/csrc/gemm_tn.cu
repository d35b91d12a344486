// Split-K "TN" GEMM for weight gradients of the tall node encoders:
//   out[P,Q] += a[N,P]^T b[N,Q],   colsum[P] += sum_n a[n,:]        (N = nodes, huge; P, Q small)
// replaces the autograd  dW = dpre^T x  /  db = dpre.sum(0)  of nn.Linear (dgm.py:1741-1744, 1778), which
// cuBLAS runs on a handful of CTAs because the output is tiny and the reduction dimension is N.
// Grid = (Q tiles) x (node splits) x (P tiles): every SM streams its own slab of rows once.
// fp32 SIMT (exact fp32 products, fp32 accumulate): 64 x QT output tile per CTA (QT = 128 / 64 / 16 by the width of
// b), 8 x QT/16 per thread.
// HBM-bound target: N*(P+Q)*4 bytes read once (+ P*Q*4*splits of reductions).
#include "common.cuh"
#include <cstdlib>

namespace dggb {

constexpr int kTnP = 64;     // rows of out per CTA  (= columns of a)
constexpr int kTnQ = 128;    // cols of out per CTA  (= columns of b)
constexpr int kTnNodes = 32; // nodes per shared-memory stage
constexpr int kTnThreads = 128;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  const int bytes = pred ? 16 : 0;   // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// VEC4: P % 4 == 0 and Q % 4 == 0 (16-byte cp.async); otherwise a scalar staging path.
// QT: width of the output tile along Q (128, 64 or 16).  The kernel is FMA-bound, not bandwidth-bound, when the
// right operand is narrow: with a fixed 128-wide tile the dWe = g_y^T x_enc product (Q = 64) and the class-logit
// weight gradient (Q = 3, padded to 4) both paid for 128 columns (13 us each at Pubmed shape, independent of the
// split count).
template <bool VEC4, int QT>
__global__ void __launch_bounds__(kTnThreads)
    gemm_tn_splitk_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, int p, int q,
                          int rows_per_split, float* __restrict__ out, float* __restrict__ colsum) {
  pdl_trigger();
  pdl_wait();
  constexpr int CQ = QT / 16;          // output columns per thread: 8, 4 or 1
  __shared__ __align__(16) float as[2][kTnNodes][kTnP];
  __shared__ __align__(16) float bs[2][kTnNodes][QT];
  const int tid = threadIdx.x;
  const int q0 = blockIdx.x * QT;
  const int p0 = blockIdx.z * kTnP;
  const int r_begin = blockIdx.y * rows_per_split;
  const int r_end = min(n, r_begin + rows_per_split);
  const int tp = (tid / 16) * 8;                       // 8 rows of out (p) per thread
  const int tq = (tid % 16) * (CQ >= 4 ? 4 : 1);       // CQ == 8: columns [tq, tq+4) and [64+tq, 64+tq+4)
  float acc[8][CQ];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < CQ; ++j) acc[i][j] = 0.f;
  float csum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) csum[i] = 0.f;

  auto stage = [&](int buf, int r0) {
    if (VEC4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {   // a tile: 32 nodes x 64 floats = 512 float4 -> 4 per thread
        const int v = tid + k * kTnThreads;
        const int node = v / (kTnP / 4), c = (v % (kTnP / 4)) * 4;
        const int r = r0 + node;
        const bool ok = r < r_end && (p0 + c) < p;
        cp_async16(&as[buf][node][c], a + (size_t)(ok ? r : 0) * p + (ok ? p0 + c : 0), ok);
      }
      constexpr int kB4 = kTnNodes * QT / 4;   // float4 of the b tile
#pragma unroll
      for (int k = 0; k < (kB4 + kTnThreads - 1) / kTnThreads; ++k) {
        const int v = tid + k * kTnThreads;
        if (kB4 % kTnThreads == 0 || v < kB4) {
          const int node = v / (QT / 4), c = (v % (QT / 4)) * 4;
          const int r = r0 + node;
          const bool ok = r < r_end && (q0 + c) < q;
          cp_async16(&bs[buf][node][c], b + (size_t)(ok ? r : 0) * q + (ok ? q0 + c : 0), ok);
        }
      }
    } else {
      for (int v = tid; v < kTnNodes * kTnP; v += kTnThreads) {
        const int node = v / kTnP, c = v % kTnP, r = r0 + node;
        as[buf][node][c] = (r < r_end && p0 + c < p) ? __ldg(a + (size_t)r * p + p0 + c) : 0.f;
      }
      for (int v = tid; v < kTnNodes * QT; v += kTnThreads) {
        const int node = v / QT, c = v % QT, r = r0 + node;
        bs[buf][node][c] = (r < r_end && q0 + c < q) ? __ldg(b + (size_t)r * q + q0 + c) : 0.f;
      }
    }
    cp_async_commit();
  };

  const int steps = (r_end - r_begin + kTnNodes - 1) / kTnNodes;
  if (steps > 0) stage(0, r_begin);
  for (int s = 0; s < steps; ++s) {
    const int buf = s & 1;
    if (s + 1 < steps) {
      stage(buf ^ 1, r_begin + (s + 1) * kTnNodes);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll 8
    for (int node = 0; node < kTnNodes; ++node) {
      const float4 a0 = *reinterpret_cast<const float4*>(&as[buf][node][tp]);
      const float4 a1 = *reinterpret_cast<const float4*>(&as[buf][node][tp + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[CQ];
      if constexpr (CQ >= 4) {
        const float4 b0 = *reinterpret_cast<const float4*>(&bs[buf][node][tq]);
        bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
        if constexpr (CQ == 8) {
          const float4 b1 = *reinterpret_cast<const float4*>(&bs[buf][node][tq + 64]);   // conflict-free LDS.128
          bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
        }
      } else {
        bv[0] = bs[buf][node][tq];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        csum[i] += av[i];
#pragma unroll
        for (int j = 0; j < CQ; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
  // split-K reduction straight into the (zeroed / accumulated) output
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int pi = p0 + tp + i;
    if (pi >= p) continue;
    if constexpr (CQ >= 4) {
#pragma unroll
      for (int j = 0; j < CQ; j += 4) {
        const int qj = q0 + tq + (j ? 64 : 0);
        float* dst = out + (size_t)pi * q + qj;
        if (VEC4 && qj + 3 < q) {
          red_add4(dst, make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]));
        } else {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            if (qj + jj < q) atomicAdd(dst + jj, acc[i][j + jj]);
        }
      }
    } else {
      if (q0 + tq < q) atomicAdd(out + (size_t)pi * q + q0 + tq, acc[i][0]);
    }
    if (colsum != nullptr && blockIdx.x == 0 && (tid % 16) == 0) atomicAdd(colsum + pi, csum[i]);
  }
}

}  // namespace dggb
using namespace dggb;

extern "C" int dggb_gemm_tn_splitk(const float* a, const float* b, int32_t n, int32_t p, int32_t q, float* out,
                                   float* colsum_a, void* stream) {
  if (!a || !b || !out || n < 0 || p <= 0 || q <= 0) return DGGB_ERR_BAD_ARG;
  if (n == 0) return DGGB_OK;
  const int qt = q <= 16 ? 16 : (q <= 64 ? 64 : kTnQ);
  const int q_tiles = (q + qt - 1) / qt, p_tiles = (p + kTnP - 1) / kTnP;
  // enough node splits to put ~2 CTAs on every SM, at least one 32-node stage each
  int splits = (2 * kNumSMs + q_tiles * p_tiles - 1) / (q_tiles * p_tiles);
  const int max_splits = (n + kTnNodes - 1) / kTnNodes;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int rows_per_split = (n + splits - 1) / splits;
  rows_per_split = (rows_per_split + kTnNodes - 1) / kTnNodes * kTnNodes;
  splits = (n + rows_per_split - 1) / rows_per_split;
  const dim3 grid(q_tiles, splits, p_tiles);
  const bool vec4 = (p % 4 == 0) && (q % 4 == 0) && ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0) &&
                    ((uintptr_t)out % 16 == 0);
  cudaStream_t st = as_stream(stream);
#define DGGB_TN_LAUNCH(V_, QT_) \
  launch_pdl(gemm_tn_splitk_kernel<V_, QT_>, grid, dim3(kTnThreads), 0, st, a, b, n, p, q, rows_per_split, out, colsum_a)
  if (vec4) {
    if (qt == 16) DGGB_TN_LAUNCH(true, 16);
    else if (qt == 64) DGGB_TN_LAUNCH(true, 64);
    else DGGB_TN_LAUNCH(true, 128);
  } else {
    if (qt == 16) DGGB_TN_LAUNCH(false, 16);
    else if (qt == 64) DGGB_TN_LAUNCH(false, 64);
    else DGGB_TN_LAUNCH(false, 128);
  }
#undef DGGB_TN_LAUNCH
  return launch_status();
}
