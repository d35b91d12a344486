// One-shot all-reduce over NVLink peer memory, meant to sit INSIDE a captured training step.
//
// The data-parallel DGG step ends with the sum of ~36 k gradient floats over the ranks.  As an NCCL call that is a
// fixed ~20-35 us launch serialised behind a ~110 us step (SCALE_r01: 0.109 -> 0.143 ms at 8 GPUs).  Here every rank
// keeps its flat gradient buffer in symmetric memory (peer-mapped over NVLink 5 / NVSwitch), and one small kernel
//   1. tells every peer "my buffer is complete" (st.release.sys into the peer's signal pad) and waits for theirs,
//   2. reads all `world` buffers and sums them (128-bit volatile loads over NVLink; or ONE multimem.ld_reduce per
//      16 bytes when the buffers are bound to an NVSwitch multicast object: the switch does the sum),
//   3. tells every peer "I am done reading" and waits for theirs, so that the next step may overwrite the buffers
//      (skipped with end_barrier = 0: callers that ALTERNATE between two symmetric buffers do not need it -- a peer
//      can only signal step i + 1 after its step-i kernel, i.e. its reads of buffer i % 2, has completed, and nobody
//      overwrites buffer i % 2 before its own step i + 1 kernel has passed that barrier).
// No host involvement, no stream switch, capturable in a CUDA graph; sequence numbers live in device memory because a
// replayed graph cannot change kernel arguments.
#include "common.cuh"

namespace dggb {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 multimem_sum4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}

constexpr int kPeerMaxWorld = 16;
constexpr int kPeerThreads = 256;

// pads[r]: rank r's signal pad (uint32 slots); slots [base, base + world) = "buffer complete", [base + world,
// base + 2 world) = "done reading".  state[0] = last completed sequence number, state[1] = finished-block counter.
__global__ void __launch_bounds__(kPeerThreads)
    allreduce_oneshot_kernel(float* const* __restrict__ bufs, uint32_t* const* __restrict__ pads, int rank, int world,
                             long long count, float* __restrict__ out, uint32_t* state, const float* mc, int base,
                             int end_barrier) {
  pdl_trigger();
  pdl_wait();
  __shared__ uint32_t seq_s;
  __shared__ int last_s;
  if (threadIdx.x == 0) seq_s = *reinterpret_cast<volatile uint32_t*>(state) + 1;
  __syncthreads();
  const uint32_t seq = seq_s;
  // ---- 1. every rank's buffer is complete
  if (threadIdx.x < world) {
    if (blockIdx.x == 0) {
      __threadfence_system();
      st_release_sys(pads[threadIdx.x] + base + rank, seq);
    }
    const uint32_t* mine = pads[rank] + base + threadIdx.x;
    while (ld_acquire_sys(mine) != seq) {
    }
  }
  __syncthreads();
  // ---- 2. sum
  const long long n4 = count >> 2;
  const long long tid = (long long)blockIdx.x * kPeerThreads + threadIdx.x, nth = (long long)gridDim.x * kPeerThreads;
  // (loads are issued in batches of four before any of them is consumed: a load -> store -> load chain would cost one
  //  NVLink round trip per element)
  constexpr int kB = 4;
  for (long long i0 = tid; i0 < n4; i0 += nth * kB) {
    float4 acc[kB];
    if (mc != nullptr) {
#pragma unroll
      for (int b = 0; b < kB; ++b) {
        const long long i = i0 + b * nth;
        acc[b] = i < n4 ? multimem_sum4(mc + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int b = 0; b < kB; ++b) acc[b] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p0 = 0; p0 < world; p0 += 4) {        // same order on every rank: bit-identical sums everywhere
        float4 v[kB][4];
#pragma unroll
        for (int b = 0; b < kB; ++b)
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) {
            const long long i = i0 + b * nth;
            v[b][pp] = (i < n4 && p0 + pp < world) ? ld_volatile4(bufs[p0 + pp] + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
        for (int b = 0; b < kB; ++b)
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) {
            acc[b].x += v[b][pp].x; acc[b].y += v[b][pp].y; acc[b].z += v[b][pp].z; acc[b].w += v[b][pp].w;
          }
      }
    }
#pragma unroll
    for (int b = 0; b < kB; ++b) {
      const long long i = i0 + b * nth;
      if (i < n4) st4(out + 4 * i, acc[b]);
    }
  }
  // ---- 3. everybody is done reading (the last block of this rank speaks for it)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last_s = (atomicAdd(state + 1, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last_s) {
    if (end_barrier && threadIdx.x < world) {
      st_release_sys(pads[threadIdx.x] + base + world + rank, seq);
      const uint32_t* mine = pads[rank] + base + world + threadIdx.x;
      while (ld_acquire_sys(mine) != seq) {
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      state[1] = 0;
      *reinterpret_cast<volatile uint32_t*>(state) = seq;
    }
  }
}

}  // namespace dggb
using namespace dggb;

extern "C" int dggb_allreduce_oneshot(void* const* bufs_dev, void* const* pads_dev, int32_t rank, int32_t world,
                                      int64_t count, float* out, uint32_t* state, const void* multicast_ptr,
                                      int32_t pad_slot_base, int32_t blocks, int32_t end_barrier, void* stream) {
  if (!bufs_dev || !pads_dev || !out || !state || rank < 0 || world < 1 || rank >= world || count < 0 ||
      pad_slot_base < 0 || blocks < 1)
    return DGGB_ERR_BAD_ARG;
  if (world > kPeerMaxWorld || count % 4 != 0 || ((uintptr_t)out % 16)) return DGGB_ERR_BAD_SHAPE;
  if (count == 0) return DGGB_OK;
  launch_pdl(allreduce_oneshot_kernel, dim3(blocks), dim3(kPeerThreads), 0, as_stream(stream),
             reinterpret_cast<float* const*>(bufs_dev), reinterpret_cast<uint32_t* const*>(pads_dev), (int)rank,
             (int)world, (long long)count, out, state, reinterpret_cast<const float*>(multicast_ptr),
             (int)pad_slot_base, (int)end_barrier);
  return launch_status();
}
