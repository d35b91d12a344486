// Node-encoder GEMM on the 5th-gen tensor cores:  out[N,H] = act(x[N,F] W[H,F]^T + b),  act = LeakyReLU(slope)
// (nn.Sequential(nn.Linear, nn.LeakyReLU) of dgm.py:1741-1744, 1097-1100, 1123-1126, and y = x_enc We^T).
//
// fp32 in / fp32 out with ~fp32 accuracy: every operand tile is split IN SHARED MEMORY into TF32 hi + lo
// parts by four converter warps (position-preserving, so the TMA 128-B swizzle is untouched) and the product
// is accumulated as hi*hi + hi*lo + lo*hi in one fp32 TMEM accumulator (3xTF32).  x is read from HBM exactly
// once (no pre-split copy): HBM-bound, N*F*4 + N*H*4 bytes.
//
// CTA = 128 rows of x.  warp 0: TMA producer (x box 128x32, W box Hx32 per k-block, 4-deep raw ring so ~100 KB
// are in flight per SM), warps 2-5: converters (raw -> hi/lo double buffer), then epilogue (tcgen05.ld -> bias
// -> LeakyReLU -> global), warp 1: MMA issuer + TMEM owner.
#include "common.cuh"
#include "tc05.cuh"

namespace dggb {

constexpr int kLinBM = 128;
constexpr int kLinConv = 2;         // converted (hi/lo) double buffer feeding the tensor core
// raw TMA ring depth (x 16 KB + W H*128 B per stage): deep enough to keep ~100 KB in flight per SM
template <int H> struct LinRaw { static constexpr int value = (H <= 64) ? 4 : 2; };
constexpr int kLinThreads = 192;

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__device__ __forceinline__ void split4(const float4 r, float4& h, float4& l) {
  h.x = tf32_rna(r.x); h.y = tf32_rna(r.y); h.z = tf32_rna(r.z); h.w = tf32_rna(r.w);
  l.x = tf32_rna(r.x - h.x); l.y = tf32_rna(r.y - h.y); l.z = tf32_rna(r.z - h.z); l.w = tf32_rna(r.w - h.w);
}

template <int H>   // output width (UMMA N), multiple of 16, <= 128
__global__ void __launch_bounds__(kLinThreads, 1)
    linear_tf32x3_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                         const float* __restrict__ bias, float slope, int n, int f, float* __restrict__ out) {
  constexpr int kLinRaw = LinRaw<H>::value;
  constexpr uint32_t kXBytes = kLinBM * 128;       // one k-block of x: [128 rows][32 floats]
  constexpr uint32_t kWBytes = H * 128;            // one k-block of W: [H rows][32 floats]
  constexpr uint32_t kRawBytes = kXBytes + kWBytes;             // x | w   (as landed by TMA)
  constexpr uint32_t kConvBytes = 2 * kXBytes + 2 * kWBytes;    // x hi | x lo | w hi | w lo
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* raw0 = smem;
  uint8_t* conv0 = smem + kLinRaw * kRawBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(conv0 + kLinConv * kConvBytes);
  uint64_t* raw_full = bars;                        // [kLinRaw]  TMA landed (1 + tx)
  uint64_t* raw_empty = bars + kLinRaw;             // [kLinRaw]  converters done reading (4 warps)
  uint64_t* conv_full = bars + 2 * kLinRaw;         // [kLinConv] hi/lo written (4 warps)
  uint64_t* conv_empty = conv_full + kLinConv;      // [kLinConv] MMAs that read it retired (tcgen05.commit)
  uint64_t* acc_full = conv_empty + kLinConv;       // accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kLinBM;
  const int num_kb = (f + 31) / 32;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tm_x);
    tc::tma_prefetch_desc(&tm_w);
    for (int s = 0; s < kLinRaw; ++s) {
      tc::mbar_init(raw_full + s, 1);
      tc::mbar_init(raw_empty + s, 4);
    }
    for (int s = 0; s < kLinConv; ++s) {
      tc::mbar_init(conv_full + s, 4);
      tc::mbar_init(conv_empty + s, 1);
    }
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
  }
  constexpr uint32_t kTmemCols = H <= 32 ? 32 : (H <= 64 ? 64 : 128);
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, kTmemCols);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < num_kb; ++kb, (++s == kLinRaw) ? (s = 0, ph ^= 1) : 0) {
        tc::mbar_wait(raw_empty + s, ph ^ 1);
        uint8_t* st = raw0 + s * kRawBytes;
        tc::mbar_arrive_expect_tx(raw_full + s, kRawBytes);
        tc::tma_load_2d(st, &tm_x, raw_full + s, kb * 32, row0);
        tc::tma_load_2d(st + kXBytes, &tm_w, raw_full + s, kb * 32, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::idesc_tf32(kLinBM, H);
      int s = 0;
      uint32_t ph = 0;
      uint32_t acc = 0;
      for (int kb = 0; kb < num_kb; ++kb, (++s == kLinConv) ? (s = 0, ph ^= 1) : 0) {
        tc::mbar_wait(conv_full + s, ph);
        tc::fence_after_sync();
        const uint32_t xh = tc::smem_u32(conv0 + s * kConvBytes), xl = xh + kXBytes;
        const uint32_t wh = xh + 2 * kXBytes, wl = wh + kWBytes;
#pragma unroll
        for (int sp = 0; sp < 3; ++sp) {
          const uint32_t a = (sp == 2) ? xl : xh;   // hi*hi, hi*lo, lo*hi
          const uint32_t b = (sp == 1) ? wl : wh;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            tc::mma_tf32(tmem_base, tc::smem_desc_k128(a + ks * 32), tc::smem_desc_k128(b + ks * 32), idesc, acc);
            acc = 1;
          }
        }
        tc::mma_commit(conv_empty + s);
      }
      tc::mma_commit(acc_full);
    }
  } else {
    // ---------------- converters: raw stage -> TF32 hi + lo (same positions => swizzle preserved) ----------
    const int ct = threadIdx.x - 64;   // 0..127
    int rs = 0, cs = 0;
    uint32_t rph = 0, cph = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      tc::mbar_wait(raw_full + rs, rph);
      tc::mbar_wait(conv_empty + cs, cph ^ 1);
      const float4* xr = reinterpret_cast<const float4*>(raw0 + rs * kRawBytes);
      const float4* wr = reinterpret_cast<const float4*>(raw0 + rs * kRawBytes + kXBytes);
      uint8_t* cv = conv0 + cs * kConvBytes;
      float4* xh = reinterpret_cast<float4*>(cv);
      float4* xl = reinterpret_cast<float4*>(cv + kXBytes);
      float4* wh = reinterpret_cast<float4*>(cv + 2 * kXBytes);
      float4* wl = reinterpret_cast<float4*>(cv + 2 * kXBytes + kWBytes);
#pragma unroll
      for (int i = 0; i < (int)(kXBytes / 16) / 128; ++i) {
        const int v = ct + i * 128;
        float4 h, l;
        split4(xr[v], h, l);
        xh[v] = h;
        xl[v] = l;
      }
#pragma unroll
      for (int i = 0; i < ((int)(kWBytes / 16) + 127) / 128; ++i) {
        const int v = ct + i * 128;
        if (v < (int)(kWBytes / 16)) {
          float4 h, l;
          split4(wr[v], h, l);
          wh[v] = h;
          wl[v] = l;
        }
      }
      tc::fence_proxy_async();   // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(conv_full + cs);
        tc::mbar_arrive(raw_empty + rs);
      }
      if (++rs == kLinRaw) { rs = 0; rph ^= 1; }
      if (++cs == kLinConv) { cs = 0; cph ^= 1; }
    }
    // ---------------- epilogue: thread == output row ----------------
    const int q = warp & 3;
    const int row = row0 + q * 32 + lane;
    tc::mbar_wait(acc_full, 0);
    tc::fence_after_sync();
#pragma unroll
    for (int c0 = 0; c0 < H; c0 += 16) {
      uint32_t r[16];
      tc::tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + c0, r);
      tc::tmem_ld_wait();
      if (row < n) {
        float* dst = out + (size_t)row * H + c0;
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
          float4 v;
          v.x = __uint_as_float(r[c]) + (bias ? __ldg(bias + c0 + c) : 0.f);
          v.y = __uint_as_float(r[c + 1]) + (bias ? __ldg(bias + c0 + c + 1) : 0.f);
          v.z = __uint_as_float(r[c + 2]) + (bias ? __ldg(bias + c0 + c + 2) : 0.f);
          v.w = __uint_as_float(r[c + 3]) + (bias ? __ldg(bias + c0 + c + 3) : 0.f);
          v.x = v.x > 0.f ? v.x : slope * v.x;
          v.y = v.y > 0.f ? v.y : slope * v.y;
          v.z = v.z > 0.f ? v.z : slope * v.z;
          v.w = v.w > 0.f ? v.w : slope * v.w;
          *reinterpret_cast<float4*>(dst + c) = v;
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

template <int H>
static int launch_linear(const float* x, const float* w, const float* b, float slope, int n, int f, float* out,
                         cudaStream_t st) {
  CUtensorMap tm_x, tm_w;
  int rc = make_tmap_2d_f32(&tm_x, x, (uint64_t)n, (uint64_t)f, kLinBM, 32);
  if (rc != DGGB_OK) return rc;
  rc = make_tmap_2d_f32(&tm_w, w, (uint64_t)H, (uint64_t)f, H, 32);
  if (rc != DGGB_OK) return rc;
  const size_t smem = LinRaw<H>::value * (kLinBM * 128 + H * 128) + kLinConv * (2 * kLinBM * 128 + 2 * H * 128) + 256 + 1024;
  cudaError_t e = cudaFuncSetAttribute(linear_tf32x3_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e);
  linear_tf32x3_kernel<H><<<(n + kLinBM - 1) / kLinBM, kLinThreads, smem, st>>>(tm_x, tm_w, b, slope, n, f, out);
  return launch_status();
}

}  // namespace dggb
using namespace dggb;

extern "C" int dggb_linear_act_fwd(const float* x, const float* w, const float* b, float slope, int32_t n, int32_t f,
                                   int32_t h, float* out, void* stream) {
  if (!x || !w || !out || n < 0 || f <= 0 || h <= 0) return DGGB_ERR_BAD_ARG;
  // TMA needs 16-byte row pitches and base addresses; the supported widths are the hidden sizes of the path
  if (f % 4 != 0 || ((uintptr_t)x % 16) || ((uintptr_t)w % 16) || ((uintptr_t)out % 16)) return DGGB_ERR_BAD_SHAPE;
  if (n == 0) return DGGB_OK;
  cudaStream_t st = as_stream(stream);
  switch (h) {
    case 16: return launch_linear<16>(x, w, b, slope, n, f, out, st);
    case 32: return launch_linear<32>(x, w, b, slope, n, f, out, st);
    case 64: return launch_linear<64>(x, w, b, slope, n, f, out, st);
    case 128: return launch_linear<128>(x, w, b, slope, n, f, out, st);
    default: return DGGB_ERR_BAD_SHAPE;
  }
}
