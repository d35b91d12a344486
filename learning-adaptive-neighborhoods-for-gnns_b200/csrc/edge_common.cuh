// Lane layout and dispatch helpers shared by the per-edge kernels (dgg_edge.cu, edge_mlp.cu).
#pragma once
#include <type_traits>
#include "common.cuh"

namespace dggb {

// Lane layout: the 32 lanes split into G = 32/L groups of L lanes.  A warp owns 32 consecutive edges;
// group `grp` walks the PER = 32/G consecutive edges [grp*PER, (grp+1)*PER) of that chunk, and lane `lg`
// of the group owns float4 chunks c = 4*(lg + L*t), t < T, of the H-long feature row.
template <int T>
struct RowSlice {
  float4 v[T];
};

template <int T>
__device__ __forceinline__ void load_slice(RowSlice<T>& s, const float* row, int h, int lg, int L, bool live = true) {
  // live = false: a slot beyond the entry range issues NO load.  (Pointing idle slots at row 0 looks harmless, but every
  // block of the grid has such slots in its last iteration and they all hit the same two L2 lines at about the same
  // time: tens of thousands of requests on one line serialise -- measured in the GCNII stack kernel, spmm.cu.)
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int c = 4 * (lg + L * t);
    s.v[t] = (live && c < h) ? ldg4(row + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Lanes per edge.  These kernels are issue-bound, not bandwidth-bound (ncu: 4.6 M warp instructions for 108 k edges
// with 16 lanes x 1 float4 per edge at h = 64): four float4 chunks per lane amortise the index shuffles, address
// arithmetic, group reduction and sigmoid over 4x more channels per instruction.
inline int lanes_per_edge(int h) {
  int L = 1;
  while (L < 32 && 16 * L < h) L *= 2;
  return L;
}

template <typename F>
inline int dispatch_T(int h, int L, F&& f) {
  const int T = (h + 4 * L - 1) / (4 * L);
  if (T == 1) return f(std::integral_constant<int, 1>{});
  if (T == 2) return f(std::integral_constant<int, 2>{});
  if (T <= 4) return f(std::integral_constant<int, 4>{});
  return DGGB_ERR_BAD_SHAPE;
}


}  // namespace dggb
