// Node-encoder GEMM on the 5th-gen tensor cores:  out[N,H] = act(x[N,F] W[H,F]^T + b),  act = LeakyReLU(slope)
// (nn.Sequential(nn.Linear, nn.LeakyReLU) of dgm.py:1741-1744, 1097-1100, 1123-1126, and y = x_enc We^T).
//
// fp32 in / fp32 out with ~fp32 accuracy (3xTF32: hi*hi + hi*lo + lo*hi in one fp32 TMEM accumulator):
//   * hi * [W_hi ; W_lo]: the tensor core reads the RAW fp32 x tile straight from the TMA-filled shared memory
//     ("SS" MMA).  kind::tf32 ignores the low 13 mantissa bits of its operands, i.e. it computes with
//     hi = trunc_tf32(x) -- no conversion pass sits between the TMA and this, the larger, product;
//   * lo * W_hi: four converter warps (thread == row) read the same tile, form lo = x - trunc_tf32(x) (one AND, one
//     SUB per element, exact in fp32) and hand it to the tensor core through TENSOR MEMORY (tcgen05.st -> "TS" MMA).
// x is read from HBM exactly once (no pre-split copy): N*F*4 + N*H*4 bytes.
//
// CTA = 128 rows of x.  warp 0: TMA producer (x box 128x32 ring + [W_hi ; W_lo] box 2Hx32 ring per k-block),
// warps 2-5: converters, then epilogue (tcgen05.ld -> bias -> LeakyReLU -> global), warp 1: MMA issuer + TMEM.
#include "common.cuh"
#include "tc05.cuh"

namespace dggb {

constexpr int kLinBM = 128;
constexpr int kLinThreads = 192;

using tc::tf32_rna;

// W -> (hi, lo) TF32 split, once per call (W is tiny: H x F); the chained GEMM's W2 is split by the same launch
// (transposed: w is given as [F, H] and the kernel needs W_eff[h][f] = w[f][h])
__global__ void split_w_kernel(const float* __restrict__ w, int count, int h, int f, int transposed,
                               float* __restrict__ hi, float* __restrict__ lo, const float* __restrict__ w2, int count2,
                               float* __restrict__ hi2, float* __restrict__ lo2, float* __restrict__ t2,
                               int h2, float* zero_ws, long long zero_count) {
  // wait FIRST, then let the dependent grid go: when the GEMM kernel's CTAs start, everything before this launch has
  // completed, so the GEMM may stream its x tiles (never written by this kernel) without waiting for the split --
  // only its W loads wait.  The x ring fills while W is being split.
  pdl_wait();
  pdl_trigger();
  zero_fill(zero_ws, zero_count);
  const int count3 = t2 != nullptr ? count2 : 0;   // [W2^T_hi ; W2^T_lo] stacked ([2 h2, h2]): the backward's B operand
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count + count2 + count3; i += gridDim.x * blockDim.x) {
    if (i >= count + count2) {
      const int k = i - count - count2;                          // k = r * h2 + c  ->  W2^T[r][c] = w2[c][r]
      const float v = __ldg(w2 + (size_t)(k % h2) * h2 + (k / h2));
      const float hh = tf32_rna(v);
      t2[k] = hh;
      t2[count2 + k] = tf32_rna(v - hh);
    } else if (i < count) {
      const float v = transposed ? __ldg(w + (size_t)(i % f) * h + (i / f)) : __ldg(w + i);
      const float hh = tf32_rna(v);
      hi[i] = hh;
      lo[i] = tf32_rna(v - hh);
    } else {
      const int k = i - count;
      const float v = __ldg(w2 + k);
      const float hh = tf32_rna(v);
      hi2[k] = hh;
      lo2[k] = tf32_rna(v - hh);
    }
  }
}

using tc::mma_tf32_ts;
using tc::tmem_st_32x32;
using tc::tmem_st_wait;

// Shared-memory traffic is what bounded the first version of this kernel (raw tile in, hi + lo tiles out,
// three tensor-core reads): here the converters keep the split x operand in REGISTERS and hand it to the
// tensor core through TMEM (tcgen05.st -> "TS" MMA, A from tensor memory); only W (pre-split) is read from
// shared memory by the MMA.
//
// Pipeline (clock64 trace of the previous single-ring version: the k-loop ran at ~1200 cycles per k-block because
// a stage was only recycled after TMA latency + conversion + MMA, 3 stages deep):
//   * x ring (XS x 16 KB, HBM stream): a stage is released once the converters have it in registers AND the hi MMAs
//     that read it have retired;
//   * W ring (WS x [W_hi ; W_lo] stacked along N, L2-resident) is released by the commit of the lo MMAs;
//   * the hi MMAs of k-block k+1 are issued BEFORE the lo MMAs of k-block k: the only work that waits for the
//     converters is the small lo product, and four TMEM lo buffers keep their tcgen05.st off the critical path.
//     Measured (profiles/r02_linear_trace.md): the k-block period is ~1050 cycles either way -- each tcgen05.mma
//     takes ~95 cycles to issue because the operand reads, the TMA writes and the converters' LDS share the SM's
//     shared-memory bandwidth (88 KB per k-block here, 72 KB when both A halves went through TMEM); the HBM rate
//     would allow 720;
//   * one producer thread polls both rings; the first XS + WS loads are issued before the TMEM allocation.
//   * 3xTF32 as TWO MMAs per k-step instead of three: A_hi x [W_hi ; W_lo]^T (N = 2H: hi*hi | hi*lo side by side
//     in the accumulator) and A_lo x W_hi^T (N = H, accumulated onto the hi*hi columns); the epilogue adds the two
//     column halves.  Measured tcgen05.mma kind::tf32 M=128 K=8 cost: N=64 47 cycles, N=128 64, N=256 127
//     (scripts/micro/mma_rate.cu), so 111 instead of 141 cycles per k-step at H = 64.
// FUSE2 (H <= 64): a second GEMM out2 = out W2^T (the edge-encoder projection y = x_enc We^T of dgm.py:1784) is
// chained in the epilogue: the activated tile goes registers -> TMEM as the A operand, W2 hi/lo arrive by TMA
// into the retired rings, the accumulator columns are reused.
#ifdef DGGB_LIN_TRACE
__device__ long long g_lin_trace[2][160];
__device__ int g_lin_smid[1024];
#define LTRACE(slot) do { if (blockIdx.x == 0 || blockIdx.x == 77) g_lin_trace[blockIdx.x ? 1 : 0][slot] = clock64(); } while (0)
#else
#define LTRACE(slot) do {} while (0)
#endif
template <int H, int XS, int WS, bool FUSE2>
__global__ void __launch_bounds__(kLinThreads, 1)
    linear_tf32x3_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                         const __grid_constant__ CUtensorMap tm_w2hi, const __grid_constant__ CUtensorMap tm_w2lo,
                         const float* __restrict__ bias, const float* __restrict__ addend,
                         const float* __restrict__ act_src, float slope, int n, int f, float* __restrict__ out,
                         float* __restrict__ out2, int wait_first, float* __restrict__ outT_hi,
                         float* __restrict__ outT_lo, int npad, float* __restrict__ colsum,
                         float* __restrict__ xT_hi, float* __restrict__ xT_lo, int kb_per_split,
                         float* __restrict__ part) {
  static_assert(!FUSE2 || H == 32 || H == 64, "fused second GEMM: K = H must be one or two 32-float k-blocks");
  constexpr uint32_t kXBytes = kLinBM * 128;       // one k-block of x: [128 rows][32 floats], 128-B swizzled
  constexpr uint32_t kWBytes = H * 128;            // one k-block of W hi (or lo): [H rows][32 floats]
  constexpr uint32_t kWStage = 2 * kWBytes;        // [W_hi ; W_lo]: 2H rows
  constexpr uint32_t kRing = XS * kXBytes + WS * kWStage;
  constexpr uint32_t kAccCols = 2 * H;             // hi*hi (+ lo*hi) | hi*lo
  constexpr int kAB = 4;                           // TMEM buffers of the lo operand (32 columns each)
  constexpr uint32_t kTmemCols = (kAccCols + 128 <= 256) ? 256 : 512;   // + kAB x (A lo 32 cols)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kRing);
  pdl_trigger();
  if (threadIdx.x == 0) LTRACE(0);
#ifdef DGGB_LIN_TRACE
  if (threadIdx.x == 0 && blockIdx.x < 1024) { uint32_t sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); g_lin_smid[blockIdx.x] = (int)sm; }
#endif
  uint64_t* x_full = bars;                    // [XS] TMA landed
  uint64_t* x_empty = x_full + XS;            // [XS] 4 converter warps have the tile in registers + hi MMAs retired
  uint64_t* w_full = x_empty + XS;            // [WS]
  uint64_t* w_empty = w_full + WS;            // [WS] MMAs that read it retired (commit)
  uint64_t* a_full = w_empty + WS;            // [kAB] lo operand written to TMEM by the 4 converter warps
  uint64_t* a_empty = a_full + kAB;           // [kAB] MMAs that read it retired
  uint64_t* acc_full = a_empty + kAB;
  uint64_t* w2_full = acc_full + 1;           // W2 hi/lo landed (second GEMM)
  uint64_t* a2_full = acc_full + 2;           // activated tile written to TMEM by the 4 epilogue warps
  uint64_t* acc2_full = acc_full + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kLinBM;
  // split-K (part != NULL; few row tiles, wide x): blockIdx.y owns k-blocks [kb0, kb0 + num_kb) and leaves its raw
  // partial tile in part[blockIdx.y]; bias / addend / activation are applied by linear_splitk_epilogue_kernel, which
  // sums the partials in a fixed order (deterministic, no atomics)
  const int kb0 = part != nullptr ? (int)blockIdx.y * kb_per_split : 0;
  const int num_kb = part != nullptr ? min(kb_per_split, (f + 31) / 32 - kb0) : (f + 31) / 32;

  auto load_x = [&](int kb) {
    const int s = kb % XS;
    tc::mbar_arrive_expect_tx(x_full + s, kXBytes);
    tc::tma_load_2d(smem + s * kXBytes, &tm_x, x_full + s, (kb0 + kb) * 32, row0);
  };
  auto load_w = [&](int kb) {
    const int s = kb % WS;
    tc::mbar_arrive_expect_tx(w_full + s, kWStage);
    tc::tma_load_2d(smem + XS * kXBytes + s * kWStage, &tm_w, w_full + s, (kb0 + kb) * 32, 0);
  };

  if (warp == 0) {
    // one barrier per lane (26 serial mbarrier.init by one thread were ~700 of the ~1 500 cycles before the first TMA)
    constexpr int kNumBars = 2 * XS + 2 * WS + 2 * kAB + 4;
    static_assert(kNumBars <= 32, "one barrier per lane of warp 0");
    if (lane < kNumBars) {
      uint32_t count = 1;                                                     // TMA / commit barriers
      if (lane >= XS && lane < 2 * XS) count = 5;                             // x_empty: 4 converter warps + hi commit
      else if (lane >= 2 * XS + 2 * WS && lane < 2 * XS + 2 * WS + kAB) count = 4;   // a_full: 4 converter warps
      else if (lane == kNumBars - 2) count = 4;                               // a2_full: 4 epilogue warps
      tc::mbar_init(bars + lane, count);
    }
    tc::fence_barrier_init();
    __syncwarp();
  }
  if (warp == 0 && lane == 0) {
    if (tc::smem_u32(smem) & 1023u) __trap();      // 128-B swizzle atoms need a 1024-B aligned window
    tc::tma_prefetch_desc(&tm_x);
    tc::tma_prefetch_desc(&tm_w);
    // The preceding grid is split_w_kernel of the same call, which triggers this launch only AFTER its own
    // dependency wait: x (and everything else older than the split) is complete and visible, only the split W is not.
    // The first round of both rings needs no "empty" wait: x goes in flight right away, W after the wait.
    // (wait_first: W arrives pre-split and no split kernel precedes this launch -- the preceding grid may be the
    // producer of x, so nothing is loaded before the dependency wait.)
    if (wait_first) pdl_wait();
    for (int kb = 0; kb < XS && kb < num_kb; ++kb) load_x(kb);
    if (!wait_first) pdl_wait();
    for (int kb = 0; kb < WS && kb < num_kb; ++kb) load_w(kb);
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, kTmemCols);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_a0 = tmem_base + kAccCols;   // A buffers start after the accumulator columns
  if (threadIdx.x == 0) LTRACE(1);

  if (warp == 0) {
    if (lane == 0) {
      int xk = num_kb < XS ? num_kb : XS, wk = num_kb < WS ? num_kb : WS;   // next k-block of each ring
      while (xk < num_kb || wk < num_kb) {
        bool moved = false;
        if (xk < num_kb && tc::mbar_try_wait(x_empty + xk % XS, ((xk / XS) - 1) & 1)) {
          LTRACE(8 + xk);
          load_x(xk++);
          moved = true;
        }
        if (wk < num_kb && tc::mbar_try_wait(w_empty + wk % WS, ((wk / WS) - 1) & 1)) {
          load_w(wk++);
          moved = true;
        }
        if (!moved) __nanosleep(32);
      }
      if (FUSE2) {
        tc::mbar_wait_backoff(acc_full, 0);            // GEMM 1 retired: both rings are free again
        uint8_t* w2 = smem + 36864;                    // behind the epilogue staging tile; [hi ; lo] per k-block
        tc::mbar_arrive_expect_tx(w2_full, (H / 32) * kWStage);
#pragma unroll
        for (int kb2 = 0; kb2 < H / 32; ++kb2) {
          tc::tma_load_2d(w2 + kb2 * kWStage, &tm_w2hi, w2_full, kb2 * 32, 0);
          tc::tma_load_2d(w2 + kb2 * kWStage + kWBytes, &tm_w2lo, w2_full, kb2 * 32, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc2 = tc::idesc_tf32(kLinBM, 2 * H);
      constexpr uint32_t idesc1 = tc::idesc_tf32(kLinBM, H);
      // hi(kb): raw x tile (shared memory, truncated to tf32 by the tensor core) x [W_hi ; W_lo] -> [hi*hi | hi*lo]
      auto issue_hi = [&](int kb) {
        const int xs = kb % XS, ws = kb % WS;
        tc::mbar_wait_backoff(x_full + xs, (kb / XS) & 1);
        tc::mbar_wait_backoff(w_full + ws, (kb / WS) & 1);
        LTRACE(24 + kb);
        tc::fence_after_sync();
        const uint32_t xst = tc::smem_u32(smem + xs * kXBytes);
        const uint32_t wst = tc::smem_u32(smem + XS * kXBytes + ws * kWStage);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          tc::mma_tf32(tmem_base, tc::smem_desc_k128(xst + ks * 32), tc::smem_desc_k128(wst + ks * 32), idesc2,
                       (kb | ks) ? 1u : 0u);
        tc::mma_commit(x_empty + xs);
      };
      // lo(kb): lo operand (tensor memory) x W_hi, accumulated onto the hi*hi columns
      auto issue_lo = [&](int kb) {
        const int ab = kb % kAB, ws = kb % WS;
        tc::mbar_wait_backoff(a_full + ab, (kb / kAB) & 1);
        LTRACE(136 + kb);
        tc::fence_after_sync();
        const uint32_t wst = tc::smem_u32(smem + XS * kXBytes + ws * kWStage);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          mma_tf32_ts(tmem_base, tmem_a0 + ab * 32 + ks * 8, tc::smem_desc_k128(wst + ks * 32), idesc1, 1u);
        tc::mma_commit(w_empty + ws);
        tc::mma_commit(a_empty + ab);
        LTRACE(40 + kb);
      };
      // Issue order.  DGGB_LIN_HI_AHEAD: hi(kb + 1) before lo(kb) (the order of the first versions, when the lo operand
      // was the late one).  r02b trace: x lands ~2 000 cycles before its MMAs issue and the lo operand ~450 cycles after
      // that, so nothing waits for the converters any more -- but W(kb + 3) can only be requested once lo(kb) has
      // retired, and with hi(kb + 1), hi(kb + 2) queued in front of it that request came too late for hi(kb + 3)
      // (~300 cycles of every 1 100-cycle k-block were spent waiting for the W stage).  In order, lo(kb) retires two
      // MMA groups earlier.
#ifdef DGGB_LIN_HI_AHEAD
      issue_hi(0);
      for (int kb = 0; kb < num_kb; ++kb) {
        if (kb + 1 < num_kb) issue_hi(kb + 1);
        issue_lo(kb);
      }
#else
      for (int kb = 0; kb < num_kb; ++kb) {
        issue_hi(kb);
        issue_lo(kb);
      }
#endif
      tc::mma_commit(acc_full);
      if (FUSE2) {
        tc::mbar_wait_backoff(w2_full, 0);
        tc::mbar_wait_backoff(a2_full, 0);
        tc::fence_after_sync();
        const uint32_t w2s = tc::smem_u32(smem + 36864);
        uint32_t acc2 = 0;
#pragma unroll
        for (int kb2 = 0; kb2 < H / 32; ++kb2)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t bdesc = tc::smem_desc_k128(w2s + kb2 * kWStage + ks * 32);
            mma_tf32_ts(tmem_base, tmem_a0 + kb2 * 32 + ks * 8, bdesc, idesc2, acc2);
            mma_tf32_ts(tmem_base, tmem_a0 + 64 + kb2 * 32 + ks * 8, bdesc, idesc1, 1u);
            acc2 = 1;
          }
        tc::mma_commit(acc2_full);
      }
    }
  } else {
    // ---------------- converters: thread == row.  smem (swizzled) -> registers -> hi/lo -> TMEM ------------
    const int q = warp & 3;
    const int r_in_tile = q * 32 + lane;          // == TMEM lane this thread may access
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    if (addend != nullptr || act_src != nullptr) {
      // the epilogue's operand rows: start them towards L2 now (after the dependency wait: they may come from the
      // preceding kernel), the k-loop hides the HBM latency
      pdl_wait();
      const int prow = row0 + r_in_tile;
      if (prow < n) {
#pragma unroll
        for (int c = 0; c < H * 4; c += 128) {
          if (addend != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(addend + (size_t)prow * H) + c));
          if (act_src != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(act_src + (size_t)prow * H) + c));
        }
      }
    }
    for (int kb = 0; kb < num_kb; ++kb) {
      const int ab = kb % kAB, xs = kb % XS;
      tc::mbar_wait(x_full + xs, (kb / XS) & 1);
      if (threadIdx.x == 64) LTRACE(72 + kb);
      // row r of the box: 128 B at r*128, its 16-B chunk c stored at chunk position c ^ (r & 7)
      const uint8_t* rowp = smem + xs * kXBytes + r_in_tile * 128;
      uint32_t lo[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(rowp + ((c ^ (r_in_tile & 7)) << 4));
        // lo = x - trunc_tf32(x): what the hi MMA (which ignores the low 13 mantissa bits of the raw tile) leaves out
        lo[4 * c + 0] = __float_as_uint(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u));
        lo[4 * c + 1] = __float_as_uint(v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u));
        lo[4 * c + 2] = __float_as_uint(v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u));
        lo[4 * c + 3] = __float_as_uint(v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u));
      }
      if (xT_hi != nullptr) {
        // transposed TF32 split of the INPUT tile, xT[col][node] ([F, npad]): the d pre launch leaves g_y^T behind for
        // the dWe = g_y^T x_enc product that rides along in the weight-gradient GEMM launch.  hi = trunc_tf32(x)
        // (what the tensor core sees of the raw tile), lo = x - hi; lane == row: coalesced 128-byte stores.
        const int node = row0 + r_in_tile;
        if (node < npad) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(rowp + ((c ^ (r_in_tile & 7)) << 4));
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int colx = (kb0 + kb) * 32 + 4 * c + j;
              if (colx < f) {
                const size_t o = (size_t)colx * npad + node;
                xT_hi[o] = __uint_as_float(__float_as_uint(vv[j]) & 0xffffe000u);
                xT_lo[o] = __uint_as_float(lo[4 * c + j]);
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(x_empty + xs);        // tile is in registers (the hi MMAs release it as well)
      tc::mbar_wait(a_empty + ab, ((kb / kAB) & 1) ^ 1);   // previous MMAs on this lo buffer retired
      tc::fence_after_sync();
      tmem_st_32x32(tmem_a0 + lane_addr + ab * 32, lo);
      tmem_st_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(a_full + ab);
      if (threadIdx.x == 64) LTRACE(104 + kb);
    }
    // ---------------- epilogue: TMEM -> registers -> this warp's shared-memory slice -> coalesced rows -------
    // (both rings are idle by now: all TMA loads landed and all MMAs retired before acc_full)
    tc::mbar_wait(acc_full, 0);
    if (threadIdx.x == 64) LTRACE(2);
    tc::fence_after_sync();
    pdl_wait();   // returns immediately here (the producer passed it long ago); orders this thread's global accesses
    // staging pitch H + 4 floats: 16-B aligned rows; STS.128 by thread == row and LDS.128 by (row, 16-B chunk)
    // are both conflict-free per quarter warp (row stride 4 banks mod 32)
    constexpr int kPitch = H + 4;
    float* stg = reinterpret_cast<float*>(smem) + (size_t)(q * 32) * kPitch;
    auto drain_acc = [&]() {                                        // stg[row][c] = D[c] + D[H + c]
#pragma unroll
      for (int c0 = 0; c0 < H; c0 += 16) {
        uint32_t r1[16], r2[16];
        tc::tmem_ld_32x16(tmem_base + lane_addr + c0, r1);
        tc::tmem_ld_32x16(tmem_base + lane_addr + H + c0, r2);
        tc::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; c += 4)
          *reinterpret_cast<float4*>(stg + lane * kPitch + c0 + c) =
              make_float4(__uint_as_float(r1[c]) + __uint_as_float(r2[c]),
                          __uint_as_float(r1[c + 1]) + __uint_as_float(r2[c + 1]),
                          __uint_as_float(r1[c + 2]) + __uint_as_float(r2[c + 2]),
                          __uint_as_float(r1[c + 3]) + __uint_as_float(r2[c + 3]));
      }
      __syncwarp();
    };
    drain_acc();
    if (threadIdx.x == 64) LTRACE(6);
    if (!FUSE2 && part != nullptr) {                 // split-K: raw partial tile, coalesced 128-bit rows
      constexpr int kLpr = H / 4, kRpi = 32 / kLpr;
      float* dst = part + (size_t)blockIdx.y * n * H;
      for (int it = 0; it < 32 / kRpi; ++it) {
        const int lr = it * kRpi + lane / kLpr, pc = (lane % kLpr) * 4;
        const int row = row0 + q * 32 + lr;
        if (row < n)
          *reinterpret_cast<float4*>(dst + (size_t)row * H + pc) = *reinterpret_cast<const float4*>(stg + lr * kPitch + pc);
      }
    } else {
    // coalesced 128-bit rows: lane -> (row = it * kRowsPerIt + lane / kLanesPerRow, 16-B chunk = lane % kLanesPerRow)
    constexpr int kLanesPerRow = H / 4, kRowsPerIt = 32 / kLanesPerRow, kIters = 32 / kRowsPerIt;
    const int er = lane / kLanesPerRow, ec = (lane % kLanesPerRow) * 4;
    const float4 bv = bias ? __ldg(reinterpret_cast<const float4*>(bias + ec)) : make_float4(0.f, 0.f, 0.f, 0.f);
    // global loads of a batch are issued before they are used: the backward form reads two [128, H] tiles (addend,
    // act_src) here, four batches of four were four exposed memory round trips (~1 us each) per CTA
    constexpr int kBatchMax = FUSE2 ? 4 : 8;
    constexpr int kBatch = kIters < kBatchMax ? kIters : kBatchMax;
    for (int it0 = 0; it0 < kIters; it0 += kBatch) {
      float4 ad[kBatch], ac[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int row = row0 + q * 32 + (it0 + u) * kRowsPerIt + er;
        const bool ok = row < n;
        ad[u] = (addend != nullptr && ok) ? __ldg(reinterpret_cast<const float4*>(addend + (size_t)row * H + ec))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
        ac[u] = (act_src != nullptr && ok) ? __ldg(reinterpret_cast<const float4*>(act_src + (size_t)row * H + ec))
                                           : make_float4(1.f, 1.f, 1.f, 1.f);
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int lr = (it0 + u) * kRowsPerIt + er;
        const int row = row0 + q * 32 + lr;
        float4 v = *reinterpret_cast<const float4*>(stg + lr * kPitch + ec);
        v.x += bv.x + ad[u].x; v.y += bv.y + ad[u].y; v.z += bv.z + ad[u].z; v.w += bv.w + ad[u].w;
        if (act_src != nullptr) {
          v.x *= ac[u].x > 0.f ? 1.f : slope; v.y *= ac[u].y > 0.f ? 1.f : slope;
          v.z *= ac[u].z > 0.f ? 1.f : slope; v.w *= ac[u].w > 0.f ? 1.f : slope;
        } else {
          v.x = v.x > 0.f ? v.x : slope * v.x; v.y = v.y > 0.f ? v.y : slope * v.y;
          v.z = v.z > 0.f ? v.z : slope * v.z; v.w = v.w > 0.f ? v.w : slope * v.w;
        }
        // (chained form: x_enc leaves for global memory later, while the second GEMM's MMAs run)
        if (row >= n) v = make_float4(0.f, 0.f, 0.f, 0.f);
        else if (!FUSE2 && out != nullptr) *reinterpret_cast<float4*>(out + (size_t)row * H + ec) = v;
        // the finished tile goes back to the staging slice for the chained GEMM / the transposed copy
        if (FUSE2 || outT_hi != nullptr) *reinterpret_cast<float4*>(stg + lr * kPitch + ec) = v;
      }
      if (it0 == 0 && threadIdx.x == 64) LTRACE(7);
    }
    if (threadIdx.x == 64) LTRACE(3);
    if (!FUSE2 && outT_hi != nullptr) {
      // ---- transposed TF32 hi/lo copy of the tile, outT[p][node] ([H, npad], zero padded), and its column sums:
      // what the weight-gradient GEMM dW = out^T x (gemm_tn_tc.cu) takes as its K-major B operand and the bias
      // gradient -- written here instead of by a transpose pass that re-reads `out` (one launch and 2 N H 4 bytes
      // less per backward).  lane == row of this warp's 32-row slice: 128-byte coalesced stores per p.
      __syncwarp();
      const int node = row0 + q * 32 + lane;
      if (node < npad) {
#pragma unroll 4
        for (int c = 0; c < H; c += 4) {
          const float4 v = *reinterpret_cast<const float4*>(stg + lane * kPitch + c);
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float hh = tf32_rna(vv[j]);
            outT_hi[(size_t)(c + j) * npad + node] = hh;
            outT_lo[(size_t)(c + j) * npad + node] = tf32_rna(vv[j] - hh);
          }
        }
      }
      if (colsum != nullptr) {
        // per-warp column sums -> shared memory -> ONE atomic per column and CTA (every CTA of the grid reaches this
        // point at about the same time and targets the same H addresses)
        float* cs = reinterpret_cast<float*>(smem) + (size_t)kLinBM * kPitch;    // [4][H], behind the staging tile
#pragma unroll
        for (int c0 = 0; c0 < H; c0 += 32) {
          const int c = c0 + lane;
          if (c < H) {
            float t = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) t += stg[r * kPitch + c];     // rows >= n hold zeros
            cs[q * H + c] = t;
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");                  // the four epilogue warps
        if (q == 0) {
          if ((reinterpret_cast<uintptr_t>(colsum) & 15) == 0) {        // 16-byte reductions: a quarter of the requests
            for (int c = 4 * lane; c < H; c += 128) {                   // on the two lines every CTA adds into
              float4 t;
              t.x = cs[c] + cs[H + c] + cs[2 * H + c] + cs[3 * H + c];
              t.y = cs[c + 1] + cs[H + c + 1] + cs[2 * H + c + 1] + cs[3 * H + c + 1];
              t.z = cs[c + 2] + cs[H + c + 2] + cs[2 * H + c + 2] + cs[3 * H + c + 2];
              t.w = cs[c + 3] + cs[H + c + 3] + cs[2 * H + c + 3] + cs[3 * H + c + 3];
              red_add4(colsum + c, t);
            }
          } else {
#pragma unroll
            for (int c0 = 0; c0 < H; c0 += 32) {
              const int c = c0 + lane;
              if (c < H) {
                const float t = cs[c] + cs[H + c] + cs[2 * H + c] + cs[3 * H + c];
                if (t != 0.f) atomicAdd(colsum + c, t);
              }
            }
          }
        }
      }
    }
    if (FUSE2) {
      // ---- chained GEMM: this thread's activated row -> TF32 hi/lo -> TMEM (A operand, K = H) ----
      __syncwarp();
#pragma unroll
      for (int kb2 = 0; kb2 < H / 32; ++kb2) {
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const float4 v = *reinterpret_cast<const float4*>(stg + lane * kPitch + kb2 * 32 + c);
          const float h0 = tf32_rna(v.x), h1 = tf32_rna(v.y), h2 = tf32_rna(v.z), h3 = tf32_rna(v.w);
          hi[c] = __float_as_uint(h0); hi[c + 1] = __float_as_uint(h1);
          hi[c + 2] = __float_as_uint(h2); hi[c + 3] = __float_as_uint(h3);
          lo[c] = __float_as_uint(tf32_rna(v.x - h0)); lo[c + 1] = __float_as_uint(tf32_rna(v.y - h1));
          lo[c + 2] = __float_as_uint(tf32_rna(v.z - h2)); lo[c + 3] = __float_as_uint(tf32_rna(v.w - h3));
        }
        tmem_st_32x32(tmem_a0 + lane_addr + kb2 * 32, hi);
        tmem_st_32x32(tmem_a0 + lane_addr + 64 + kb2 * 32, lo);
      }
      tmem_st_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(a2_full);
      // the activated tile (still in the staging slice) -> out, overlapped with the chained MMAs
#pragma unroll 4
      for (int it = 0; it < kIters; ++it) {
        const int lr = it * kRowsPerIt + er;
        const int row = row0 + q * 32 + lr;
        if (row < n)
          *reinterpret_cast<float4*>(out + (size_t)row * H + ec) = *reinterpret_cast<const float4*>(stg + lr * kPitch + ec);
      }
      __syncwarp();                                  // stg is overwritten by the second drain below
      tc::mbar_wait(acc2_full, 0);
      if (threadIdx.x == 64) LTRACE(4);
      tc::fence_after_sync();
      drain_acc();
#pragma unroll 4
      for (int it = 0; it < kIters; ++it) {
        const int lr = it * kRowsPerIt + er;
        const int row = row0 + q * 32 + lr;
        if (row < n)
          *reinterpret_cast<float4*>(out2 + (size_t)row * H + ec) = *reinterpret_cast<const float4*>(stg + lr * kPitch + ec);
      }
    }
    }   // !split-K
  }
  __syncwarp();
  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) LTRACE(5);
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// out = epi(sum_s part[s] + bias + addend) for the split-K form of the kernel above (fixed summation order)
__global__ void __launch_bounds__(256)
    linear_splitk_epilogue_kernel(const float* __restrict__ part, int splits, int n, int h, const float* __restrict__ bias,
                                  const float* __restrict__ addend, const float* __restrict__ act_src, float slope,
                                  float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long long total4 = (long long)n * h / 4;
  const int h4 = h / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = ldg4(part + 4 * i);
    for (int sp = 1; sp < splits; ++sp) {
      const float4 u = ldg4(part + (size_t)sp * n * h + 4 * i);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    const int c = (int)(i % h4) * 4;
    if (bias != nullptr) {
      const float4 b = ldg4(bias + c);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (addend != nullptr) {
      const float4 a = ldg4(addend + 4 * i);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    if (act_src != nullptr) {
      const float4 a = ldg4(act_src + 4 * i);
      v.x *= a.x > 0.f ? 1.f : slope; v.y *= a.y > 0.f ? 1.f : slope;
      v.z *= a.z > 0.f ? 1.f : slope; v.w *= a.w > 0.f ? 1.f : slope;
    } else {
      v.x = v.x > 0.f ? v.x : slope * v.x; v.y = v.y > 0.f ? v.y : slope * v.y;
      v.z = v.z > 0.f ? v.z : slope * v.z; v.w = v.w > 0.f ? v.w : slope * v.w;
    }
    st4(out + 4 * i, v);
  }
}

// k-splits of a call: few row tiles (small graphs with wide raw features: Cora 22 tiles x 45 k-blocks, Citeseer
// 26 x 116) leave most SMs idle and make one CTA stream its whole 128-row slab alone (72 us at Citeseer shape)
static int linear_splits(int n, int f, int* kb_per_split) {
  const int tiles = (n + kLinBM - 1) / kLinBM, total_kb = (f + 31) / 32;
  int splits = 1;
  if (tiles * 2 <= kNumSMs && total_kb >= 16) {
    splits = kNumSMs / tiles;
    if (splits > total_kb / 8) splits = total_kb / 8;
    if (splits < 1) splits = 1;
  }
  *kb_per_split = (total_kb + splits - 1) / splits;
  return (total_kb + *kb_per_split - 1) / *kb_per_split;
}

// Optional pieces of one call (all pointers may be NULL):
//   w_presplit : `w` is the stacked [W_hi ; W_lo] ([2H, F]) written by an earlier launch; no split kernel runs
//   w2t_split  : out, [2H, H]: stacked split of W2^T (the B operand of the backward's d pre GEMM)
//   outT_hi/lo : out, [H, npad] transposed TF32 split of `out` + colsum[H] += column sums (see the kernel epilogue)
struct LinearExtra {
  bool w_presplit = false;
  float* w2t_split = nullptr;
  float* outT_hi = nullptr;
  float* outT_lo = nullptr;
  int npad = 0;
  float* colsum = nullptr;
  float* xT_hi = nullptr;       // out, [F, npad]: transposed TF32 split of x (see the converter loop)
  float* xT_lo = nullptr;
  float* splitk_ws = nullptr;   // [splits, N, H] partial tiles: enables the split-K form when it pays (linear_splits)
  long long splitk_ws_bytes = 0;
};

template <int H>
static int launch_linear(const float* x, const float* w, int w_transposed, const float* b, const float* addend,
                         const float* act_src, float slope, int n, int f, float* out, float* ws, const float* w2,
                         float* out2, float* zero_ws, long long zero_count, cudaStream_t st,
                         const LinearExtra& ex = LinearExtra()) {
#ifndef DGGB_LIN_XS
#define DGGB_LIN_XS 4
#define DGGB_LIN_WS 3
#endif
  constexpr int XS = DGGB_LIN_XS, WS = DGGB_LIN_WS;   // H <= 64: 64 KB + 3 x 2H x 128 B <= 112 KB so that two CTAs share an SM
  float* w_hi = ex.w_presplit ? const_cast<float*>(w) : ws;   // [W_hi ; W_lo] stacked: one [2H, F] matrix, one tensor map
  const bool chained = (H == 32 || H == 64) && w2 != nullptr;
  float* w2_hi = ex.w_presplit ? nullptr : ws + (size_t)2 * H * f;
  float* w2_lo = ex.w_presplit ? nullptr : w2_hi + (size_t)H * H;
  int rc;
  if (!ex.w_presplit) {
    float* w_lo = ws + (size_t)H * f;
    const int n_split = H * f + (chained ? H * H * (ex.w2t_split ? 2 : 1) : 0);
    launch_pdl(split_w_kernel, dim3((n_split + 255) / 256), dim3(256), 0, st, w, H * f, H, f, w_transposed, w_hi, w_lo,
               chained ? w2 : static_cast<const float*>(nullptr), chained ? H * H : 0, w2_hi, w2_lo,
               chained ? ex.w2t_split : static_cast<float*>(nullptr), H, zero_ws, zero_count);
    rc = launch_status();
    if (rc != DGGB_OK) return rc;
  }
  const int wait_first = ex.w_presplit ? 1 : 0;
  CUtensorMap tm_x, tm_w, tm_w2hi, tm_w2lo;
  rc = make_tmap_2d_f32(&tm_x, x, (uint64_t)n, (uint64_t)f, kLinBM, 32);
  if (rc != DGGB_OK) return rc;
  rc = make_tmap_2d_f32(&tm_w, w_hi, (uint64_t)2 * H, (uint64_t)f, 2 * H, 32);
  if (rc != DGGB_OK) return rc;
  const size_t smem = XS * (kLinBM * 128) + WS * (2 * H * 128) + 256;
  const int grid = (n + kLinBM - 1) / kLinBM;
  if constexpr (H == 32 || H == 64) {
    if (w2 != nullptr) {
      if (ex.w_presplit || ex.outT_hi || ex.xT_hi) return DGGB_ERR_BAD_ARG;
      rc = make_tmap_2d_f32(&tm_w2hi, w2_hi, (uint64_t)H, (uint64_t)H, H, 32);
      if (rc != DGGB_OK) return rc;
      rc = make_tmap_2d_f32(&tm_w2lo, w2_lo, (uint64_t)H, (uint64_t)H, H, 32);
      if (rc != DGGB_OK) return rc;
      cudaError_t e = cudaFuncSetAttribute(linear_tf32x3_kernel<H, XS, WS, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_status(e);
      launch_pdl((linear_tf32x3_kernel<H, XS, WS, true>), dim3(grid), dim3(kLinThreads), smem, st, tm_x, tm_w,
                 tm_w2hi, tm_w2lo, b, addend, act_src, slope, n, f, out, out2, 0, static_cast<float*>(nullptr),
                 static_cast<float*>(nullptr), 0, static_cast<float*>(nullptr), static_cast<float*>(nullptr),
                 static_cast<float*>(nullptr), 0, static_cast<float*>(nullptr));
      return launch_status();
    }
  }
  if (w2 != nullptr) return DGGB_ERR_BAD_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(linear_tf32x3_kernel<H, XS, WS, false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e);
  int kb_per_split = 0;
  const int splits = (ex.splitk_ws && !ex.outT_hi && !ex.xT_hi && out) ? linear_splits(n, f, &kb_per_split) : 1;
  if (splits > 1 && ex.splitk_ws_bytes >= (long long)splits * n * H * (long long)sizeof(float)) {
    launch_pdl((linear_tf32x3_kernel<H, XS, WS, false>), dim3(grid, splits), dim3(kLinThreads), smem, st, tm_x, tm_w,
               tm_w, tm_w, b, addend, act_src, slope, n, f, out, static_cast<float*>(nullptr), wait_first,
               static_cast<float*>(nullptr), static_cast<float*>(nullptr), 0, static_cast<float*>(nullptr),
               static_cast<float*>(nullptr), static_cast<float*>(nullptr), kb_per_split, ex.splitk_ws);
    rc = launch_status();
    if (rc != DGGB_OK) return rc;
    const long long total4 = (long long)n * H / 4;
    const int eblocks = (int)min((long long)kNumSMs * 4, (total4 + 255) / 256);
    launch_pdl(linear_splitk_epilogue_kernel, dim3(eblocks), dim3(256), 0, st, static_cast<const float*>(ex.splitk_ws),
               splits, n, H, b, addend, act_src, slope, out);
    return launch_status();
  }
  launch_pdl((linear_tf32x3_kernel<H, XS, WS, false>), dim3(grid), dim3(kLinThreads), smem, st, tm_x, tm_w, tm_w, tm_w,
             b, addend, act_src, slope, n, f, out, static_cast<float*>(nullptr), wait_first, ex.outT_hi, ex.outT_lo,
             ex.npad, ex.colsum, ex.xT_hi, ex.xT_lo, 0, static_cast<float*>(nullptr));
  return launch_status();
}

template <typename... A>
static int dispatch_linear(int h, A... a) {
  switch (h) {
    case 16: return launch_linear<16>(a...);
    case 32: return launch_linear<32>(a...);
    case 64: return launch_linear<64>(a...);
    case 128: return launch_linear<128>(a...);
    default: return DGGB_ERR_BAD_SHAPE;
  }
}

}  // namespace dggb
using namespace dggb;

extern "C" int64_t dggb_linear_act_workspace_bytes(int32_t f, int32_t h) {
  if (f <= 0 || h <= 0) return DGGB_ERR_BAD_ARG;
  return ((int64_t)2 * h * f + (int64_t)2 * h * h) * 4;   // W hi/lo (+ W2 hi/lo of the chained GEMM)
}

// bytes of the split-K partial-tile workspace dggb_linear_fused can use for this shape (0: the call would not split)
extern "C" int64_t dggb_linear_splitk_workspace_bytes(int32_t n, int32_t f, int32_t h) {
  if (n <= 0 || f <= 0 || h <= 0) return 0;
  int kbps = 0;
  const int splits = linear_splits(n, f, &kbps);
  return splits > 1 ? (int64_t)splits * n * h * (int64_t)sizeof(float) : 0;
}

static bool misaligned(const void* p) { return p != nullptr && ((uintptr_t)p % 16) != 0; }

extern "C" int dggb_linear_fused(const float* x, const float* w, int32_t w_transposed, const float* b,
                                 const float* addend, const float* act_src, float slope, int32_t n, int32_t f,
                                 int32_t h, float* out, const float* w2, float* out2, void* workspace,
                                 int64_t workspace_bytes, float* zero_ws, int64_t zero_count, void* splitk_ws,
                                 int64_t splitk_ws_bytes, void* stream) {
  if (!x || !w || !out || !workspace || n < 0 || f <= 0 || h <= 0 || ((w2 == nullptr) != (out2 == nullptr)) ||
      zero_count < 0 || splitk_ws_bytes < 0 || misaligned(splitk_ws))
    return DGGB_ERR_BAD_ARG;
  if (workspace_bytes < dggb_linear_act_workspace_bytes(f, h)) return DGGB_ERR_WORKSPACE;
  if ((uintptr_t)workspace % 16) return DGGB_ERR_BAD_ARG;
  float* ws = reinterpret_cast<float*>(workspace);
  // TMA needs 16-byte row pitches and base addresses; the supported widths are the hidden sizes of the path
  if (f % 4 != 0 || misaligned(x) || misaligned(w) || misaligned(out) || (w2 && (h > 64 || h % 32)) ||
      misaligned(addend) || misaligned(act_src) || misaligned(b) || misaligned(out2))
    return DGGB_ERR_BAD_SHAPE;
  if (n == 0) return DGGB_OK;
  LinearExtra ex;
  if (w2 == nullptr && splitk_ws != nullptr) {
    ex.splitk_ws = reinterpret_cast<float*>(splitk_ws);
    ex.splitk_ws_bytes = (long long)splitk_ws_bytes;
  }
  return dispatch_linear(h, x, w, (int)w_transposed, b, addend, act_src, slope, (int)n, (int)f, out, ws, w2, out2,
                         zero_ws, (long long)zero_count, as_stream(stream), ex);
}

extern "C" int dggb_linear_act_fwd(const float* x, const float* w, const float* b, float slope, int32_t n, int32_t f,
                                   int32_t h, float* out, void* workspace, int64_t workspace_bytes, void* stream) {
  return dggb_linear_fused(x, w, 0, b, nullptr, nullptr, slope, n, f, h, out, nullptr, nullptr, workspace,
                           workspace_bytes, nullptr, 0, nullptr, 0, stream);
}

// (x_enc, y) = (LeakyReLU_slope(x Wn^T + bn), x_enc We^T) in one launch (+ the weight-split launch), see dggb.h
extern "C" int dggb_encoder_fwd(const float* x, const float* wn, const float* bn, float slope, int32_t n, int32_t f,
                                int32_t h, float* x_enc, const float* we, float* y, void* workspace,
                                int64_t workspace_bytes, float* we_t_split, float* zero_ws, int64_t zero_count,
                                void* stream) {
  if (!x || !wn || !x_enc || !we || !y || !workspace || n < 0 || f <= 0 || zero_count < 0) return DGGB_ERR_BAD_ARG;
  if (h != 32 && h != 64) return DGGB_ERR_BAD_SHAPE;
  if (workspace_bytes < dggb_linear_act_workspace_bytes(f, h)) return DGGB_ERR_WORKSPACE;
  if ((uintptr_t)workspace % 16) return DGGB_ERR_BAD_ARG;
  if (f % 4 != 0 || misaligned(x) || misaligned(wn) || misaligned(x_enc) || misaligned(bn) || misaligned(y) ||
      misaligned(we_t_split))
    return DGGB_ERR_BAD_SHAPE;
  if (n == 0) return DGGB_OK;
  LinearExtra ex;
  ex.w2t_split = we_t_split;
  return dispatch_linear(h, x, wn, 0, bn, static_cast<const float*>(nullptr), static_cast<const float*>(nullptr),
                         slope, (int)n, (int)f, x_enc, reinterpret_cast<float*>(workspace), we, y, zero_ws,
                         (long long)zero_count, as_stream(stream), ex);
}

// d pre = LeakyReLU'_slope(x_enc) * (g_y We + g_xenc) from the PRE-SPLIT We^T of dggb_encoder_fwd: one launch; the tile
// also leaves as the transposed TF32 split + column sums that dggb_gemm_tn_tc_presplit consumes
extern "C" int dggb_encoder_bwd_dpre(const float* g_y, const float* we_t_split, const float* g_xenc,
                                     const float* x_enc, float slope, int32_t n, int32_t h, float* dpre,
                                     float* dpre_t_hi, float* dpre_t_lo, int32_t npad, float* colsum,
                                     float* gy_t_hi, float* gy_t_lo, void* stream) {
  if (!g_y || !we_t_split || !x_enc || n < 0 || (dpre_t_hi == nullptr) != (dpre_t_lo == nullptr) ||
      (gy_t_hi == nullptr) != (gy_t_lo == nullptr) || (!dpre && !dpre_t_hi))
    return DGGB_ERR_BAD_ARG;
  if (h != 16 && h != 32 && h != 64 && h != 128) return DGGB_ERR_BAD_SHAPE;
  if ((dpre_t_hi || gy_t_hi) && (npad < n || npad % 4 != 0)) return DGGB_ERR_BAD_ARG;
  if (misaligned(g_y) || misaligned(we_t_split) || misaligned(g_xenc) || misaligned(x_enc) || misaligned(dpre))
    return DGGB_ERR_BAD_SHAPE;
  if (n == 0) return DGGB_OK;
  LinearExtra ex;
  ex.w_presplit = true;
  ex.outT_hi = dpre_t_hi;
  ex.outT_lo = dpre_t_lo;
  ex.npad = npad;
  ex.colsum = colsum;
  ex.xT_hi = gy_t_hi;
  ex.xT_lo = gy_t_lo;
  return dispatch_linear(h, g_y, we_t_split, 0, static_cast<const float*>(nullptr), g_xenc, x_enc, slope, (int)n,
                         (int)h, dpre, static_cast<float*>(nullptr), static_cast<const float*>(nullptr),
                         static_cast<float*>(nullptr), static_cast<float*>(nullptr), 0LL, as_stream(stream), ex);
}

#ifdef DGGB_LIN_TRACE
extern "C" int dggb_debug_lin_smid(int* host_out) {
  return cudaMemcpyFromSymbol(host_out, dggb::g_lin_smid, sizeof(int) * 1024) == cudaSuccess ? 0 : -1;
}
extern "C" int dggb_debug_lin_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, dggb::g_lin_trace, sizeof(long long) * 2 * 160) == cudaSuccess ? 0 : -1;
}
#endif
