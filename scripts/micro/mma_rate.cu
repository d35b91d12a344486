// tcgen05.mma kind::tf32 issue/execute rate vs N and operand mode (SS: A from smem, TS: A from TMEM).
// One CTA per SM (or two with -2); thread 0 issues REPS x 4 MMAs (K = 8 each over one 32-float k-block), commits, waits.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../learning-adaptive-neighborhoods-for-gnns_b200/csrc/tc05.cuh"
using namespace dggb;

template <int N, bool TS>
__global__ void __launch_bounds__(128) rate_kernel(long long* out, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (tc::smem_u32(smem_raw) & 1023)) & 1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (threadIdx.x < 32) { tc::tmem_alloc(&slot, 512); tc::tmem_relinquish(); }
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = tc::idesc_tf32(128, N);
    const uint32_t a_s = tc::smem_u32(smem), b_s = a_s + 16 * 1024;   // A: 128 x 128 B, B: up to 256 x 128 B
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        if (TS) tc::mma_tf32_ts(tm, tm + 256 + ks * 8, tc::smem_desc_k128(b_s + ks * 32), idesc, 1u);
        else tc::mma_tf32(tm, tc::smem_desc_k128(a_s + ks * 32), tc::smem_desc_k128(b_s + ks * 32), idesc, 1u);
      }
    }
    const long long t1 = clock64();
    tc::mma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { tc::fence_after_sync(); tc::tmem_dealloc(tm, 512); }
}

template <int N, bool TS>
void run(int ctas_per_sm, int reps) {
  long long* d; cudaMalloc(&d, 16);
  auto k = rate_kernel<N, TS>;
  const int smem = ctas_per_sm == 1 ? 120 * 1024 : 50 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int it = 0; it < 2; ++it) k<<<148 * ctas_per_sm, 128, smem>>>(d, reps);
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const cudaError_t e = cudaGetLastError();
  const int n_mma = reps * 4;
  printf("N=%3d %s ctas/sm=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %.1f)  %s\n", N, TS ? "TS" : "SS",
         ctas_per_sm, (double)h[0] / n_mma, (double)h[1] / n_mma, 128.0 * N * 8 / 1934.0,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  // TMEM budget: 512 columns per CTA -> only one CTA per SM can hold 512; use ctas_per_sm = 1 here
  run<64, false>(1, 64); run<64, true>(1, 64);
  run<128, false>(1, 64); run<128, true>(1, 64);
  run<256, false>(1, 64); run<256, true>(1, 64);
  run<64, true>(1, 16); run<128, true>(1, 16);
  return 0;
}
