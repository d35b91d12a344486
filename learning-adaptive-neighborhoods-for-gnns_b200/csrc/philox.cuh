// Philox4x32-7 counter-based generator shared by the all-pairs Gumbel noise (allpairs.cu) and the sub-graph samplers
// (sampler.cu).  Host restatement: tests/philox_ref.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dggb {

__device__ __forceinline__ uint4 philox4x32_7(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1) {
  uint32_t x0 = c0, x1 = c1, x2 = 0u, x3 = 0u;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
    const uint32_t y0 = hi1 ^ x1 ^ k0, y1 = lo1, y2 = hi0 ^ x3 ^ k1, y3 = lo0;
    x0 = y0; x1 = y1; x2 = y2; x3 = y3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(x0, x1, x2, x3);
}

}  // namespace dggb
