// Tensor-core version of the weight-gradient GEMM  out[P,Q] += a[N,P]^T b[N,Q]  (dW = dpre^T x, reduction
// over the N nodes; see gemm_tn.cu for the SIMT fallback).  Same machinery as linear_tc.cu, transposed:
//
//   D[m = feature (128 per CTA), n = p] = sum_k  b[k, m] * a[k, n],   k = node
//   A operand = b^T : converter thread m gathers its column of the TMA-landed [32 nodes][32 feats] boxes
//               (swizzle-aware LDS), splits it into TF32 hi/lo in registers and hands it over through TMEM
//   B operand = a^T : a is small ([N,P]), a pre-pass writes a^T hi/lo [P, Npad] (K-major for the tensor core)
//               and the column sums of a (the bias gradient)
//   3xTF32 (hi*hi + hi*lo + lo*hi), fp32 accumulate in TMEM; split-K over node ranges, fp32 reductions out.
// b (= x, the big operand) is read from HBM exactly once: N*Q*4 bytes.
#include "common.cuh"
#include "tc05.cuh"

namespace dggb {

using tc::mma_tf32_ts;
using tc::tf32_rna;
using tc::tmem_st_32x32;
using tc::tmem_st_wait;

constexpr int kTcM = 128;        // features per CTA (TMEM lanes)
constexpr int kTcStages = 3;
constexpr int kTcThreads = 192;

// a [N,P] -> aT_hi, aT_lo [P, npad] (zero padded) and colsum[P] += sum_n a[n,:]
__global__ void __launch_bounds__(256)
    transpose_split_kernel(const float* __restrict__ a, int n, int npad, int p, float* __restrict__ hi,
                           float* __restrict__ lo, float* __restrict__ colsum) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[32][33];
  const int n0 = blockIdx.x * 32, p0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  float cs = 0.f;
  for (int r = ty; r < 32; r += 8) {
    const int nn = n0 + r, pp = p0 + tx;
    const float v = (nn < n && pp < p) ? __ldg(a + (size_t)nn * p + pp) : 0.f;
    tile[r][tx] = v;
    cs += v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int pp = p0 + r, nn = n0 + tx;
    if (pp < p && nn < npad) {
      const float v = tile[tx][r];
      const float h = tf32_rna(v);
      hi[(size_t)pp * npad + nn] = h;
      lo[(size_t)pp * npad + nn] = tf32_rna(v - h);
    }
  }
  if (colsum != nullptr) {
    __syncthreads();
    tile[ty][tx] = cs;
    __syncthreads();
    if (ty == 0) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) t += tile[r][tx];
      if (p0 + tx < p) atomicAdd(colsum + p0 + tx, t);
    }
  }
}

// STACK (P <= 64, a^T_lo stored right behind a^T_hi): the B operand of a k-block is ONE [2P rows][32 nodes] box
// [a^T_hi ; a^T_lo]; 3xTF32 is then TWO MMAs per k-step -- A_hi x [B_hi ; B_lo] (N = 2P: hi*hi | hi*lo side by side in the
// accumulator) and A_lo x B_hi (N = P, onto the hi*hi columns) -- instead of three N = P ones (64 + 47 against
// 3 x 47 cycles at P = 64, scripts/micro/mma_rate.cu); the epilogue adds the two column halves.
template <int P, bool STACK>
__global__ void __launch_bounds__(kTcThreads, 2)
    gemm_tn_tf32x3_kernel(const __grid_constant__ CUtensorMap tm_b1, const __grid_constant__ CUtensorMap tm_ahi1,
                          const __grid_constant__ CUtensorMap tm_alo1, int n, int q1, int kb_per_split,
                          float* __restrict__ out1, int m_tiles1, const __grid_constant__ CUtensorMap tm_b2,
                          const __grid_constant__ CUtensorMap tm_ahi2, const __grid_constant__ CUtensorMap tm_alo2,
                          int q2, float* __restrict__ out2) {
  // CTAs with blockIdx.x >= m_tiles1 work on the second (narrow) product of the same launch: other operands, same
  // node split
  const bool second = (int)blockIdx.x >= m_tiles1;
  const CUtensorMap& tm_b = second ? tm_b2 : tm_b1;
  const CUtensorMap& tm_ahi = second ? tm_ahi2 : tm_ahi1;
  const CUtensorMap& tm_alo = second ? tm_alo2 : tm_alo1;
  const int q = second ? q2 : q1;
  float* __restrict__ out = second ? out2 : out1;
  constexpr uint32_t kXBytes = 4 * 32 * 128;       // four [32 nodes][32 feats] boxes = 128 features x 32 nodes
  constexpr uint32_t kWBytes = P * 128;            // a^T hi (or lo): [P rows][32 nodes]
  constexpr uint32_t kStageBytes = kXBytes + 2 * kWBytes;
  static_assert(!STACK || P <= 64, "stacked B: accumulator 2P + two A buffers must fit 256 TMEM columns");
  constexpr uint32_t kAccCols = STACK ? 128 : (P <= 64 ? 64 : 128);
  constexpr uint32_t kTmemCols = 256;
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTcStages * kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kTcStages;              // 4 converter warps + MMA commit
  uint64_t* a_full = bars + 2 * kTcStages;
  uint64_t* a_empty = a_full + 2;
  uint64_t* acc_full = a_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = (second ? (int)blockIdx.x - m_tiles1 : (int)blockIdx.x) * kTcM;   // first feature of this CTA
  const int kb_begin = blockIdx.y * kb_per_split;
  const int kb_total = (n + 31) / 32;
  const int num_kb = max(0, min(kb_per_split, kb_total - kb_begin));

  if (warp == 0) {
    constexpr int kNumBars = 2 * kTcStages + 5;      // full | empty | a_full[2] | a_empty[2] | acc_full: one per lane
    if (lane < kNumBars) {
      uint32_t count = 1;
      if (lane >= kTcStages && lane < 2 * kTcStages) count = 5;                   // empty: 4 converter warps + commit
      else if (lane >= 2 * kTcStages && lane < 2 * kTcStages + 2) count = 4;      // a_full: 4 converter warps
      tc::mbar_init(bars + lane, count);
    }
    tc::fence_barrier_init();
    __syncwarp();
    if (lane == 0) {
      tc::tma_prefetch_desc(&tm_b);
      tc::tma_prefetch_desc(&tm_ahi);
      tc::tma_prefetch_desc(&tm_alo);
    }
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, kTmemCols);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_a0 = tmem_base + kAccCols;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      pdl_wait();   // a^T hi/lo come from transpose_split_kernel, b possibly from the kernel before it
      for (int i = 0; i < num_kb; ++i, (++s == kTcStages) ? (s = 0, ph ^= 1) : 0) {
        const int node0 = (kb_begin + i) * 32;
        tc::mbar_wait(empty + s, ph ^ 1);
        uint8_t* st = smem + s * kStageBytes;
        tc::mbar_arrive_expect_tx(full + s, kStageBytes);
#pragma unroll
        for (int bx = 0; bx < 4; ++bx) tc::tma_load_2d(st + bx * 4096, &tm_b, full + s, m0 + bx * 32, node0);
        if (STACK) {
          tc::tma_load_2d(st + kXBytes, &tm_ahi, full + s, node0, 0);      // [2P][32]: hi rows, then lo rows
        } else {
          tc::tma_load_2d(st + kXBytes, &tm_ahi, full + s, node0, 0);
          tc::tma_load_2d(st + kXBytes + kWBytes, &tm_alo, full + s, node0, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::idesc_tf32(kTcM, P);
      constexpr uint32_t idesc2 = tc::idesc_tf32(kTcM, STACK ? 2 * P : P);
      int s = 0;
      uint32_t ph = 0, acc = 0;
      for (int i = 0; i < num_kb; ++i, (++s == kTcStages) ? (s = 0, ph ^= 1) : 0) {
        const int ab = i & 1;
        const uint32_t aph = (i >> 1) & 1;
        tc::mbar_wait(full + s, ph);
        tc::mbar_wait(a_full + ab, aph);
        tc::fence_after_sync();
        const uint32_t wh = tc::smem_u32(smem + s * kStageBytes + kXBytes), wl = wh + kWBytes;
        const uint32_t ah = tmem_a0 + ab * 64, al = ah + 32;
        if (STACK) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {          // A_hi x [B_hi ; B_lo] -> [hi*hi | hi*lo]
            mma_tf32_ts(tmem_base, ah + ks * 8, tc::smem_desc_k128(wh + ks * 32), idesc2, acc);
            acc = 1;
          }
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)            // A_lo x B_hi onto the hi*hi columns
            mma_tf32_ts(tmem_base, al + ks * 8, tc::smem_desc_k128(wh + ks * 32), idesc, 1u);
        } else {
#pragma unroll
          for (int sp = 0; sp < 3; ++sp) {
            const uint32_t a = (sp == 2) ? al : ah;   // hi*hi, hi*lo, lo*hi
            const uint32_t b = (sp == 1) ? wl : wh;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              mma_tf32_ts(tmem_base, a + ks * 8, tc::smem_desc_k128(b + ks * 32), idesc, acc);
              acc = 1;
            }
          }
        }
        tc::mma_commit(empty + s);
        tc::mma_commit(a_empty + ab);
      }
      tc::mma_commit(acc_full);
    }
  } else {
    // ---------------- converters: thread == feature (TMEM lane); gathers its column of the node tile --------
    const int q4 = warp & 3;
    const int m = q4 * 32 + lane;                 // feature inside the CTA tile; box = m / 32 == q4, col = lane
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < num_kb; ++i, (++s == kTcStages) ? (s = 0, ph ^= 1) : 0) {
      const int ab = i & 1;
      const uint32_t aph = (i >> 1) & 1;
      tc::mbar_wait(full + s, ph);
      // box q4: [32 nodes][128 B]; element (node k, col c) at k*128 + (((c >> 2) ^ (k & 7)) << 4) + (c & 3)*4
      const uint8_t* box = smem + s * kStageBytes + q4 * 4096;
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float v = *reinterpret_cast<const float*>(box + k * 128 + ((((lane >> 2) ^ (k & 7)) << 4) | ((lane & 3) << 2)));
        const float h = tf32_rna(v);
        hi[k] = __float_as_uint(h);
        lo[k] = __float_as_uint(tf32_rna(v - h));
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(empty + s);
      tc::mbar_wait(a_empty + ab, aph ^ 1);
      tc::fence_after_sync();
      tmem_st_32x32(tmem_a0 + lane_addr + ab * 64, hi);
      tmem_st_32x32(tmem_a0 + lane_addr + ab * 64 + 32, lo);
      tmem_st_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(a_full + ab);
    }
    // ---------------- epilogue: D[feature][p] -> out[p][feature] (fp32 reductions, coalesced across lanes) ----
    if (num_kb > 0) {
      tc::mbar_wait(acc_full, 0);
      tc::fence_after_sync();
      pdl_wait();   // `out` was zeroed by an earlier kernel; returns immediately here
      const int feat = m0 + m;
#pragma unroll
      for (int c0 = 0; c0 < P; c0 += 16) {
        uint32_t r[16];
        tc::tmem_ld_32x16(tmem_base + lane_addr + c0, r);
        if (STACK) {
          uint32_t r2[16];
          tc::tmem_ld_32x16(tmem_base + lane_addr + P + c0, r2);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) r[c] = __float_as_uint(__uint_as_float(r[c]) + __uint_as_float(r2[c]));
        }
        tc::tmem_ld_wait();
        if (feat < q) {
#pragma unroll
          for (int c = 0; c < 16; ++c) atomicAdd(out + (size_t)(c0 + c) * q + feat, __uint_as_float(r[c]));
        }
      }
    }
  }
  __syncwarp();
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// a_hi / a_lo: a^T split, [P, npad] (npad % 4 == 0, zero padded beyond n)
struct TnSecond {     // optional second product of the launch (see the kernel)
  const float* a_hi = nullptr;
  const float* a_lo = nullptr;
  const float* b = nullptr;
  int q = 0;
  float* out = nullptr;
};

template <int P>
static int launch_tn_gemm(const float* a_hi, const float* a_lo, int npad, const float* b, int n, int q, float* out,
                          cudaStream_t st, const TnSecond& s2 = TnSecond()) {
  CUtensorMap tm_b, tm_ahi, tm_alo, tm_b2, tm_ahi2, tm_alo2;
  int rc = make_tmap_2d_f32(&tm_b, b, (uint64_t)n, (uint64_t)q, 32, 32);
  if (rc != DGGB_OK) return rc;
  static const bool no_stack = getenv("DGGB_TN_NO_STACK") != nullptr;     // A/B: three N = P MMAs per k-step
  const bool stack = (P <= 64) && !no_stack && a_lo == a_hi + (size_t)P * npad &&
                     (s2.b == nullptr || s2.a_lo == s2.a_hi + (size_t)P * npad);
  if (stack) {
    rc = make_tmap_2d_f32(&tm_ahi, a_hi, (uint64_t)2 * P, (uint64_t)npad, 2 * P, 32);
    if (rc != DGGB_OK) return rc;
    tm_alo = tm_ahi;
  } else {
    rc = make_tmap_2d_f32(&tm_ahi, a_hi, (uint64_t)P, (uint64_t)npad, P, 32);
    if (rc != DGGB_OK) return rc;
    rc = make_tmap_2d_f32(&tm_alo, a_lo, (uint64_t)P, (uint64_t)npad, P, 32);
    if (rc != DGGB_OK) return rc;
  }
  tm_b2 = tm_b, tm_ahi2 = tm_ahi, tm_alo2 = tm_alo;
  int m_tiles2 = 0;
  if (s2.b != nullptr) {
    rc = make_tmap_2d_f32(&tm_b2, s2.b, (uint64_t)n, (uint64_t)s2.q, 32, 32);
    if (rc != DGGB_OK) return rc;
    if (stack) {
      rc = make_tmap_2d_f32(&tm_ahi2, s2.a_hi, (uint64_t)2 * P, (uint64_t)npad, 2 * P, 32);
      if (rc != DGGB_OK) return rc;
      tm_alo2 = tm_ahi2;
    } else {
      rc = make_tmap_2d_f32(&tm_ahi2, s2.a_hi, (uint64_t)P, (uint64_t)npad, P, 32);
      if (rc != DGGB_OK) return rc;
      rc = make_tmap_2d_f32(&tm_alo2, s2.a_lo, (uint64_t)P, (uint64_t)npad, P, 32);
      if (rc != DGGB_OK) return rc;
    }
    m_tiles2 = (s2.q + kTcM - 1) / kTcM;
  }
  const int m_tiles1 = (q + kTcM - 1) / kTcM;
  const int m_tiles = m_tiles1 + m_tiles2;
  const int kb_total = (n + 31) / 32;
  int splits = (2 * kNumSMs + m_tiles - 1) / m_tiles;
  if (splits > kb_total) splits = kb_total;
  const int kb_per_split = (kb_total + splits - 1) / splits;
  splits = (kb_total + kb_per_split - 1) / kb_per_split;
  const size_t smem = kTcStages * (4 * 4096 + 2 * P * 128) + 256 + 1024;
  if constexpr (P <= 64) {
    if (stack) {
      cudaError_t e = cudaFuncSetAttribute(gemm_tn_tf32x3_kernel<P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem);
      if (e != cudaSuccess) return cuda_status(e);
      launch_pdl((gemm_tn_tf32x3_kernel<P, true>), dim3(m_tiles, splits), dim3(kTcThreads), smem, st, tm_b, tm_ahi,
                 tm_alo, n, q, kb_per_split, out, m_tiles1, tm_b2, tm_ahi2, tm_alo2, s2.q, s2.out);
      return launch_status();
    }
  }
  cudaError_t e = cudaFuncSetAttribute(gemm_tn_tf32x3_kernel<P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return cuda_status(e);
  launch_pdl((gemm_tn_tf32x3_kernel<P, false>), dim3(m_tiles, splits), dim3(kTcThreads), smem, st, tm_b, tm_ahi, tm_alo,
             n, q, kb_per_split, out, m_tiles1, tm_b2, tm_ahi2, tm_alo2, s2.q, s2.out);
  return launch_status();
}

template <int P>
static int launch_tn(const float* a, const float* b, int n, int q, float* out, float* colsum, float* ws,
                     cudaStream_t st) {
  const int npad = (n + 31) / 32 * 32;
  float* a_hi = ws;
  float* a_lo = ws + (size_t)P * npad;
  launch_pdl(transpose_split_kernel, dim3(dim3(npad / 32, (P + 31) / 32)), dim3(256), 0, st, a, n, npad, P, a_hi, a_lo, colsum);
  int rc = launch_status();
  if (rc != DGGB_OK) return rc;
  return launch_tn_gemm<P>(a_hi, a_lo, npad, b, n, q, out, st);
}

}  // namespace dggb
using namespace dggb;

extern "C" int64_t dggb_gemm_tn_tc_workspace_bytes(int32_t n, int32_t p) {
  if (n < 0 || p <= 0) return DGGB_ERR_BAD_ARG;
  return (int64_t)2 * p * (((int64_t)n + 31) / 32 * 32) * 4;
}

extern "C" int dggb_gemm_tn_tc(const float* a, const float* b, int32_t n, int32_t p, int32_t q, float* out,
                               float* colsum_a, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!a || !b || !out || !workspace || n < 0 || p <= 0 || q <= 0) return DGGB_ERR_BAD_ARG;
  if (q % 4 != 0 || ((uintptr_t)b % 16) || ((uintptr_t)workspace % 16)) return DGGB_ERR_BAD_SHAPE;
  if (workspace_bytes < dggb_gemm_tn_tc_workspace_bytes(n, p)) return DGGB_ERR_WORKSPACE;
  if (n == 0) return DGGB_OK;
  float* ws = reinterpret_cast<float*>(workspace);
  cudaStream_t st = as_stream(stream);
  switch (p) {
    case 16: return launch_tn<16>(a, b, n, q, out, colsum_a, ws, st);
    case 32: return launch_tn<32>(a, b, n, q, out, colsum_a, ws, st);
    case 64: return launch_tn<64>(a, b, n, q, out, colsum_a, ws, st);
    case 128: return launch_tn<128>(a, b, n, q, out, colsum_a, ws, st);
    default: return DGGB_ERR_BAD_SHAPE;
  }
}

// out[P,Q] += a^T b with a^T already split (a_t_hi / a_t_lo [P, npad], zero padded for nodes >= n): the operand
// dggb_encoder_bwd_dpre leaves behind -- no transpose pass
extern "C" int dggb_gemm_tn_tc_presplit(const float* a_t_hi, const float* a_t_lo, int32_t npad, const float* b,
                                        int32_t n, int32_t p, int32_t q, float* out, const float* a2_t_hi,
                                        const float* a2_t_lo, const float* b2, int32_t q2, float* out2,
                                        void* stream) {
  if (!a_t_hi || !a_t_lo || !b || !out || n < 0 || p <= 0 || q <= 0 || npad < n) return DGGB_ERR_BAD_ARG;
  const int n2 = (a2_t_hi != nullptr) + (a2_t_lo != nullptr) + (b2 != nullptr) + (out2 != nullptr);
  if (n2 != 0 && n2 != 4) return DGGB_ERR_BAD_ARG;
  if (q % 4 != 0 || npad % 4 != 0 || ((uintptr_t)b % 16) || ((uintptr_t)a_t_hi % 16) || ((uintptr_t)a_t_lo % 16))
    return DGGB_ERR_BAD_SHAPE;
  if (n2 && (q2 <= 0 || q2 % 4 != 0 || ((uintptr_t)b2 % 16) || ((uintptr_t)a2_t_hi % 16) || ((uintptr_t)a2_t_lo % 16)))
    return DGGB_ERR_BAD_SHAPE;
  if (n == 0) return DGGB_OK;
  cudaStream_t st = as_stream(stream);
  TnSecond s2;
  if (n2) s2.a_hi = a2_t_hi, s2.a_lo = a2_t_lo, s2.b = b2, s2.q = q2, s2.out = out2;
  switch (p) {
    case 16: return launch_tn_gemm<16>(a_t_hi, a_t_lo, npad, b, n, q, out, st, s2);
    case 32: return launch_tn_gemm<32>(a_t_hi, a_t_lo, npad, b, n, q, out, st, s2);
    case 64: return launch_tn_gemm<64>(a_t_hi, a_t_lo, npad, b, n, q, out, st, s2);
    case 128: return launch_tn_gemm<128>(a_t_hi, a_t_lo, npad, b, n, q, out, st, s2);
    default: return DGGB_ERR_BAD_SHAPE;
  }
}
