// Sub-graph sampling for the large-graph drivers on the device: the reference feeds
// train_large_graphs.py / train_reddit.py from torch_geometric.loader.GraphSAINTRandomWalkSampler
// (train_large_graphs.py:402-413, train_reddit.py:400-411; third-party, CPU, worker processes).  Its two steps are
//   * adj.random_walk(start, walk_length)   -- torch_sparse: next = col[rowptr[cur] + floor(u * deg)], u ~ U[0,1); a node
//                                              without out-edges keeps the walker in place
//   * adj.saint_subgraph(unique(walk))      -- the sub-graph INDUCED by the visited nodes, relabelled, with the ids of
//                                              the kept entries (edge attributes / normalisation counts follow them)
// Here: one thread per walker with counter-based Philox uniforms (any walker regenerates identically, the host
// restatement in tests/test_gpu_sampler.py reproduces the walks bit for bit), and a two-pass induced-subgraph
// extraction (count, host-side scan by the caller, ordered fill) straight from the int32 CSR the rest of the path uses.
#include "common.cuh"
#include "philox.cuh"

namespace dggb {

// walk[w][0] = start[w];  step s draws x = lane (s & 3) of philox(counter = (walker_offset + w, s >> 2), key = seed)
// and moves to entry floor((x >> 8) * deg / 2^24) of the current row (exact integer arithmetic)
__global__ void __launch_bounds__(256)
    random_walk_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ start,
                       int num_walkers, int walk_length, uint32_t k0, uint32_t k1, long long walker_offset,
                       int32_t* __restrict__ walk) {
  pdl_trigger();
  pdl_wait();
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < num_walkers; w += gridDim.x * blockDim.x) {
    int cur = __ldg(start + w);
    int32_t* out = walk + (size_t)w * (walk_length + 1);
    out[0] = cur;
    const unsigned long long id = (unsigned long long)(walker_offset + w);
    uint4 bits = make_uint4(0u, 0u, 0u, 0u);
    for (int s = 0; s < walk_length; ++s) {
      if ((s & 3) == 0) bits = philox4x32_7((uint32_t)id, (uint32_t)(s >> 2) ^ (uint32_t)(id >> 32) * 0x9E3779B9u, k0, k1);
      const uint32_t x = (s & 3) == 0 ? bits.x : ((s & 3) == 1 ? bits.y : ((s & 3) == 2 ? bits.z : bits.w));
      const int beg = __ldg(rowptr + cur), deg = __ldg(rowptr + cur + 1) - beg;
      if (deg > 0) cur = __ldg(col + beg + (int)(((unsigned long long)(x >> 8) * (unsigned long long)deg) >> 24));
      out[s + 1] = cur;
    }
  }
}

__global__ void mark_nodes_kernel(const int32_t* __restrict__ nodes, int m, int32_t* __restrict__ relabel) {
  pdl_trigger();
  pdl_wait();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) relabel[__ldg(nodes + i)] = i;
}

// counts[i] = entries of row nodes[i] whose column is selected; counts[m] = 0 (slot for the caller's scan total)
__global__ void induced_count_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                     const int32_t* __restrict__ nodes, int m, const int32_t* __restrict__ relabel,
                                     int32_t* __restrict__ counts) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i <= m; i += gridDim.x * wpb) {
    if (i == m) {
      if (lane == 0) counts[m] = 0;
      continue;
    }
    const int u = __ldg(nodes + i);
    const int beg = __ldg(rowptr + u), end = __ldg(rowptr + u + 1);
    int c = 0;
    for (int e = beg + lane; e < end; e += kWarp) c += (__ldg(relabel + __ldg(col + e)) >= 0);
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) counts[i] = c;
  }
}

// ordered compaction of every selected row: the kept entries stay in column order (relabelling is monotone for a
// sorted node list), sub_eid = position of the entry in the parent CSR
__global__ void induced_fill_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                    const int32_t* __restrict__ nodes, int m, const int32_t* __restrict__ relabel,
                                    const int32_t* __restrict__ sub_rowptr, int32_t* __restrict__ sub_col,
                                    int32_t* __restrict__ sub_eid) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < m; i += gridDim.x * wpb) {
    const int u = __ldg(nodes + i);
    const int beg = __ldg(rowptr + u), end = __ldg(rowptr + u + 1);
    int pos = __ldg(sub_rowptr + i);
    for (int e0 = beg; e0 < end; e0 += kWarp) {
      const int e = e0 + lane;
      const int r = (e < end) ? __ldg(relabel + __ldg(col + e)) : -1;
      const unsigned keep = __ballot_sync(0xffffffffu, r >= 0);
      if (r >= 0) {
        const int o = pos + __popc(keep & ((1u << lane) - 1u));
        sub_col[o] = r;
        sub_eid[o] = e;
      }
      pos += __popc(keep);
    }
  }
}

}  // namespace dggb
using namespace dggb;

extern "C" int dggb_random_walk(const int32_t* rowptr, const int32_t* col, int32_t n, const int32_t* start,
                                int32_t num_walkers, int32_t walk_length, uint64_t seed, int64_t walker_offset,
                                int32_t* walk, void* stream) {
  if (!rowptr || !col || !start || !walk || n <= 0 || num_walkers < 0 || walk_length < 0 || walker_offset < 0)
    return DGGB_ERR_BAD_ARG;
  if (num_walkers == 0) return DGGB_OK;
  const int blocks = (int)min((long long)kNumSMs * 8, ((long long)num_walkers + 255) / 256);
  launch_pdl(random_walk_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), rowptr, col, start, (int)num_walkers,
             (int)walk_length, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), (long long)walker_offset, walk);
  return launch_status();
}

extern "C" int dggb_induced_subgraph_count(const int32_t* rowptr, const int32_t* col, int32_t n, const int32_t* nodes,
                                           int32_t m, int32_t* relabel, int32_t* counts, void* stream) {
  if (!rowptr || !col || !nodes || !relabel || !counts || n <= 0 || m < 0 || m > n) return DGGB_ERR_BAD_ARG;
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(relabel, 0xff, (size_t)n * sizeof(int32_t), st);      // -1 everywhere
  if (e != cudaSuccess) return cuda_status(e);
  if (m > 0) {
    launch_pdl(mark_nodes_kernel, dim3((m + 255) / 256), dim3(256), 0, st, nodes, (int)m, relabel);
    const int rc = launch_status();
    if (rc != DGGB_OK) return rc;
  }
  launch_pdl(induced_count_kernel, dim3(rows_grid(m + 1, 8, 8)), dim3(256), 0, st, rowptr, col, nodes, (int)m,
             static_cast<const int32_t*>(relabel), counts);
  return launch_status();
}

extern "C" int dggb_induced_subgraph_fill(const int32_t* rowptr, const int32_t* col, const int32_t* nodes, int32_t m,
                                          const int32_t* relabel, const int32_t* sub_rowptr, int32_t* sub_col,
                                          int32_t* sub_eid, void* stream) {
  if (!rowptr || !col || !nodes || !relabel || !sub_rowptr || !sub_col || !sub_eid || m < 0) return DGGB_ERR_BAD_ARG;
  if (m == 0) return DGGB_OK;
  launch_pdl(induced_fill_kernel, dim3(rows_grid(m, 8, 8)), dim3(256), 0, as_stream(stream), rowptr, col, nodes, (int)m,
             relabel, sub_rowptr, sub_col, sub_eid);
  return launch_status();
}
