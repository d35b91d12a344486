// Shared device/host helpers for libdggb (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dggb.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libdggb is written for sm_100a only"
#endif

namespace dggb {

constexpr int kWarp = 32;
constexpr int kNumSMs = 148;           // B200
constexpr float kLeaky = 0.01f;        // nn.LeakyReLU() default (dgm.py:1743, 1748, 1752)

extern int g_last_cuda_error;
extern long long g_kernel_launches;  // kernels this library has launched (bench.py's gpu_launches)

inline int cuda_status(cudaError_t e) {
  if (e == cudaSuccess) return DGGB_OK;
  g_last_cuda_error = (int)e;
  return DGGB_ERR_CUDA;
}
inline int launch_status(int kernels = 1) {
  g_kernel_launches += kernels;
  return cuda_status(cudaGetLastError());
}
inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch: every kernel of the path opens with pdl_trigger(); pdl_wait(); and is launched
// with programmatic stream serialisation, so that its CTAs are scheduled (and run their prologue: barrier init,
// TMEM allocation, descriptor prefetch) while the preceding kernel drains.  pdl_wait() returns once the preceding
// grid has completed and its writes are visible, so nothing before it may touch global memory.  The step of this
// path is a chain of ~10-30 us kernels: without this ~2 us of launch latency is exposed at every boundary.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// grid-stride zero fill of a scratch buffer the NEXT kernels accumulate into (gradient buffers): folded into a kernel
// that runs anyway, instead of a separate fill launch that would also break the dependent-launch chain
__device__ __forceinline__ void zero_fill(float* p, long long count) {
  if (p == nullptr) return;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    const long long n4 = count >> 2;
    for (long long i = tid; i < n4; i += nth) reinterpret_cast<float4*>(p)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long i = (n4 << 2) + tid; i < count; i += nth) p[i] = 0.f;
  } else {
    for (long long i = tid; i < count; i += nth) p[i] = 0.f;
  }
}
bool pdl_enabled();   // false when DGGB_NO_PDL is set in the environment (A/B measurements)

template <typename... Params, typename... Args>
inline void launch_pdl(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at = {};
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(args)...);   // errors surface through cudaGetLastError()
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// sum over the aligned group of `width` lanes (width = power of two <= 32)
__device__ __forceinline__ float group_sum(float v, int width) {
  for (int o = width >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// max(v, 0.01 v): same value as the select form for every finite v (slope < 1), two instructions instead of three
__device__ __forceinline__ float leaky(float v) { return fmaxf(v, kLeaky * v); }
__device__ __forceinline__ float leaky_grad(float v) { return v > 0.f ? 1.f : kLeaky; }
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// 128-bit vector reduction to global memory (sm_90+): one L2 atomic op for four floats.
__device__ __forceinline__ void red_add4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// largest power of two <= v (v >= 1), capped at 32
__host__ __device__ inline int pow2_floor32(int v) {
  int p = 1;
  while (p * 2 <= v && p < 32) p *= 2;
  return p;
}

// resident blocks per SM of a kernel (registers / shared memory decide, not the 2048-thread limit alone): the
// grid-stride kernels size their grid to ONE wave with it.  A grid sized for 8 blocks/SM with a 40-register kernel
// (6 resident) runs a second, mostly idle wave and doubles a latency-bound launch.
template <typename K>
inline int resident_blocks(K kernel, int threads, size_t smem = 0) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem) != cudaSuccess || occ < 1) occ = 1;
  return occ;
}

// warp-per-row grids: rows are dealt to warps round-robin from a persistent grid
inline int rows_grid(int n_rows, int warps_per_block, int blocks_per_sm) {
  long long need = ((long long)n_rows + warps_per_block - 1) / warps_per_block;
  long long cap = (long long)kNumSMs * blocks_per_sm;
  long long g = need < cap ? need : cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace dggb
