// What can the access pattern of linear_tf32x3_kernel deliver?  Reads x [19717, 500] fp32 (39.4 MB) from HBM once,
// with no math at all, in several ways, 6 rotating buffers (> L2), 24 launches per CUDA graph:
//   plain   : grid-stride float4 loads (the "copy" pattern the HBM peak was measured with)
//   tma S B : one CTA per 128-row tile (155 CTAs), ring of S stages, B 128x32-float SW128 boxes per stage
//             (consecutive k-blocks, i.e. B*128 contiguous bytes per row and stage), the consumer frees a stage as
//             soon as it has landed
//   tma64   : 64-row tiles (309 CTAs, two per SM)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../learning-adaptive-neighborhoods-for-gnns_b200/csrc/tc05.cuh"
using namespace dggb;

namespace dggb { int g_last_cuda_error = 0; }

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_enc get_enc() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  return (PFN_enc)fn;
}
static CUtensorMap tmap(const float* base, uint64_t rows, uint64_t cols, uint32_t box_rows, CUtensorMapL2promotion pr) {
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * sizeof(float)};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_enc()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

__global__ void __launch_bounds__(256) plain_kernel(const float4* __restrict__ x, long long n4, float* sink) {
  float acc = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    const float4 a = __ldg(x + i), b = __ldg(x + i + stride), c = __ldg(x + i + 2 * stride), d = __ldg(x + i + 3 * stride);
    acc += a.x + b.y + c.z + d.w;
  }
  for (; i < n4; i += stride) acc += __ldg(x + i).x;
  if (acc == 123.456f) *sink = acc;
}

template <int ROWS>
__global__ void __launch_bounds__(64) tma_kernel(const __grid_constant__ CUtensorMap tm, int num_kb, int S, int B,
                                                 float* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (tc::smem_u32(smem_raw) & 1023)) & 1023);
  __shared__ uint64_t full[16], empty[16];
  constexpr uint32_t kBox = ROWS * 128;
  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tm);
    for (int s = 0; s < S; ++s) { tc::mbar_init(full + s, 1); tc::mbar_init(empty + s, 1); }
    tc::fence_barrier_init();
  }
  __syncthreads();
  const int row0 = blockIdx.x * ROWS;
  const int groups = (num_kb + B - 1) / B;
  if (threadIdx.x == 0) {            // producer
    for (int g = 0; g < groups; ++g) {
      const int s = g % S;
      if (g >= S) tc::mbar_wait(empty + s, ((g / S) - 1) & 1);
      const int nb = min(B, num_kb - g * B);
      tc::mbar_arrive_expect_tx(full + s, nb * kBox);
      for (int b = 0; b < nb; ++b) tc::tma_load_2d(smem + (s * B + b) * kBox, &tm, full + s, (g * B + b) * 32, row0);
    }
  } else if (threadIdx.x == 32) {    // consumer
    float acc = 0.f;
    for (int g = 0; g < groups; ++g) {
      const int s = g % S;
      tc::mbar_wait(full + s, (g / S) & 1);
      acc += *reinterpret_cast<volatile float*>(smem + s * B * kBox);
      tc::mbar_arrive(empty + s);
    }
    if (acc == 123.456f) *sink = acc;
  }
}

int main() {
  const int n = 19717, f = 500, NB = 6, LAUNCHES = 24;
  float* bufs[NB];
  for (int i = 0; i < NB; ++i) { cudaMalloc(&bufs[i], (size_t)n * f * 4); cudaMemset(bufs[i], 0, (size_t)n * f * 4); }
  float* sink; cudaMalloc(&sink, 4);
  cudaStream_t st; cudaStreamCreate(&st);
  const double mb = (double)n * f * 4 / 1e6;
  auto time_graph = [&](const char* name, auto launch) {
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < LAUNCHES; ++i) launch(i % NB);
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ge, g, 0);
    for (int i = 0; i < 20; ++i) cudaGraphLaunch(ge, st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double us = ms * 1e3 / (10 * LAUNCHES);
    const cudaError_t e = cudaGetLastError();
    printf("%-28s %6.2f us/launch  %6.0f GB/s  %s\n", name, us, mb / us * 1e3, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  };
  for (int blocks_per_sm : {2, 4, 8}) {
    char nm[64]; snprintf(nm, 64, "plain %d blocks/SM", blocks_per_sm);
    time_graph(nm, [&](int b) { plain_kernel<<<148 * blocks_per_sm, 256, 0, st>>>((const float4*)bufs[b], (long long)n * f / 4, sink); });
  }
  const int num_kb = (f + 31) / 32;
  CUtensorMapL2promotion prs[3] = {CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B};
  const char* prn[3] = {"none", "128B", "256B"};
  for (int p = 0; p < 3; ++p) {
    CUtensorMap tms[NB];
    for (int i = 0; i < NB; ++i) tms[i] = tmap(bufs[i], n, f, 128, prs[p]);
    const int cfg[][2] = {{4, 1}, {6, 1}, {8, 1}, {12, 1}, {2, 2}, {4, 2}, {6, 2}, {2, 4}, {3, 4}, {2, 8}, {1, 16}};
    for (auto& c : cfg) {
      const int S = c[0], B = c[1];
      const size_t smem = (size_t)S * B * 128 * 128 + 1024;
      if (smem > 227 * 1024) continue;
      cudaFuncSetAttribute(tma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      char nm[64]; snprintf(nm, 64, "tma128 S=%d B=%d promo=%s", S, B, prn[p]);
      time_graph(nm, [&](int b) { tma_kernel<128><<<(n + 127) / 128, 64, smem, st>>>(tms[b], num_kb, S, B, sink); });
    }
    CUtensorMap tms64[NB];
    for (int i = 0; i < NB; ++i) tms64[i] = tmap(bufs[i], n, f, 64, prs[p]);
    const int cfg64[][2] = {{4, 1}, {8, 1}, {12, 1}, {4, 2}, {3, 4}, {2, 8}};
    for (auto& c : cfg64) {
      const int S = c[0], B = c[1];
      const size_t smem = (size_t)S * B * 64 * 128 + 1024;
      if (smem > 110 * 1024) continue;
      cudaFuncSetAttribute(tma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      char nm[64]; snprintf(nm, 64, "tma64  S=%d B=%d promo=%s", S, B, prn[p]);
      time_graph(nm, [&](int b) { tma_kernel<64><<<(n + 63) / 64, 64, smem, st>>>(tms64[b], num_kb, S, B, sink); });
    }
  }
  return 0;
}
