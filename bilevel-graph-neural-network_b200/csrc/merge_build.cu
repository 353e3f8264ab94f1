// Device-side merged-batch / CSR construction (see include/bignn_b200.h).
// Replaces src/batch.py:105-144 + model/layers_util.py:100-166 + src/merged_graph.py:27-96.
//
// Three small phases over the G graphs of the batch:
//   A  per-block sums of (atoms, directed edges)            grid = ceil(G/1024)
//   B  scan of the block sums                               1 block
//   C  block-local exclusive scan + block offset -> seg_ptr / edge_ptr
//   D  one warp per graph copies its CSR slice with the node offset added,
//      writes batch ids, the optional int64 COO view and the feature rows.
// HBM-bound integer/byte work: every source array is read once, every output
// written once, all accesses are contiguous per graph.
#include "common.cuh"

namespace bignn {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;  // graphs per block

struct I2 { int n, e; };
__device__ __forceinline__ I2 add2(I2 a, I2 b) { return {a.n + b.n, a.e + b.e}; }

__device__ __forceinline__ I2 graph_size(const int32_t* __restrict__ atom_ptr,
                                         const int32_t* __restrict__ nbr_ptr,
                                         const int32_t* __restrict__ rows, int g, int G) {
  if (g >= G) return {0, 0};
  int r = rows[g];
  int a0 = atom_ptr[r], a1 = atom_ptr[r + 1];
  return {a1 - a0, nbr_ptr[a1] - nbr_ptr[a0]};
}

// exclusive scan of one I2 per thread across the block; returns the block total in `total`
__device__ __forceinline__ I2 block_excl_scan(I2 v, I2& total) {
  __shared__ I2 warp_tot[kScanThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  I2 inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int tn = __shfl_up_sync(0xffffffffu, inc.n, o);
    int te = __shfl_up_sync(0xffffffffu, inc.e, o);
    if (lane >= o) { inc.n += tn; inc.e += te; }
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  I2 base = {0, 0};
  I2 tot = {0, 0};
#pragma unroll
  for (int i = 0; i < kScanThreads / 32; ++i) {
    if (i < w) base = add2(base, warp_tot[i]);
    tot = add2(tot, warp_tot[i]);
  }
  total = tot;
  __syncthreads();
  return {base.n + inc.n - v.n, base.e + inc.e - v.e};
}

__global__ void __launch_bounds__(kScanThreads)
k_merge_blocksum(const int32_t* __restrict__ atom_ptr, const int32_t* __restrict__ nbr_ptr,
                 const int32_t* __restrict__ rows, int G, I2* __restrict__ blk) {
  int g0 = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  I2 s = {0, 0};
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) s = add2(s, graph_size(atom_ptr, nbr_ptr, rows, g0 + i, G));
  I2 tot;
  block_excl_scan(s, tot);
  if (threadIdx.x == 0) blk[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads)
k_merge_scanblocks(I2* __restrict__ blk, int nb) {
  I2 carry = {0, 0};
  for (int base = 0; base < nb; base += kScanThreads) {
    int i = base + threadIdx.x;
    I2 v = i < nb ? blk[i] : I2{0, 0};
    I2 tot;
    I2 ex = block_excl_scan(v, tot);
    if (i < nb) blk[i] = add2(ex, carry);
    carry = add2(carry, tot);
  }
}

__global__ void __launch_bounds__(kScanThreads)
k_merge_scanfinal(const int32_t* __restrict__ atom_ptr, const int32_t* __restrict__ nbr_ptr,
                  const int32_t* __restrict__ rows, int G, const I2* __restrict__ blk,
                  int32_t* __restrict__ seg_ptr, int32_t* __restrict__ edge_ptr) {
  int g0 = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  I2 sz[kScanItems];
  I2 s = {0, 0};
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    sz[i] = graph_size(atom_ptr, nbr_ptr, rows, g0 + i, G);
    s = add2(s, sz[i]);
  }
  I2 tot;
  I2 ex = block_excl_scan(s, tot);
  if (blk != nullptr) ex = add2(ex, blk[blockIdx.x]);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int g = g0 + i;
    if (g < G) { seg_ptr[g] = ex.n; edge_ptr[g] = ex.e; }
    ex = add2(ex, sz[i]);
    if (g == G - 1) { seg_ptr[G] = ex.n; edge_ptr[G] = ex.e; }
  }
  if (G == 0 && blockIdx.x == 0 && threadIdx.x == 0) { seg_ptr[0] = 0; edge_ptr[0] = 0; }
}

__global__ void __launch_bounds__(256)
k_merge_fill(const int32_t* __restrict__ atom_ptr, const int32_t* __restrict__ nbr_ptr,
             const int32_t* __restrict__ nbr_idx, const float* __restrict__ x_all, int F,
             const int32_t* __restrict__ rows, int G,
             const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ edge_ptr,
             int32_t* __restrict__ row_ptr, int32_t* __restrict__ col_idx,
             int32_t* __restrict__ batch, float* __restrict__ x,
             int64_t* __restrict__ ei, int64_t* __restrict__ batch64, int E) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int nwarps = gridDim.x * warps_per_block;
  for (int g = blockIdx.x * warps_per_block + (threadIdx.x >> 5); g < G; g += nwarps) {
    const int r = rows[g];
    const int a0 = atom_ptr[r];
    const int n = atom_ptr[r + 1] - a0;
    const int e0 = nbr_ptr[a0];
    const int s0 = seg_ptr[g];
    const int E0 = edge_ptr[g];
    for (int i = lane; i < n; i += 32) {
      const int k0 = nbr_ptr[a0 + i] - e0;
      const int k1 = nbr_ptr[a0 + i + 1] - e0;
      row_ptr[s0 + i] = E0 + k0;
      if (batch) batch[s0 + i] = g;
      if (batch64) batch64[s0 + i] = g;
      for (int k = k0; k < k1; ++k) {
        const int c = nbr_idx[e0 + k] + s0;
        col_idx[E0 + k] = c;
        if (ei) { ei[E0 + k] = s0 + i; ei[(int64_t)E + E0 + k] = c; }
      }
    }
    if (x) {
      const int64_t src = (int64_t)a0 * F, dst = (int64_t)s0 * F;
      const int cnt = n * F;
      if ((F & 3) == 0 && aligned16(x_all) && aligned16(x)) {       // whole graph block as 128-bit words
        const float4* s4 = reinterpret_cast<const float4*>(x_all + src);
        float4* d4 = reinterpret_cast<float4*>(x + dst);
        for (int t = lane; t < (cnt >> 2); t += 32) d4[t] = __ldg(s4 + t);
      } else {
        for (int t = lane; t < cnt; t += 32) x[dst + t] = x_all[src + t];
      }
    }
    if (g == G - 1 && lane == 0) row_ptr[s0 + n] = E0 + (nbr_ptr[a0 + n] - e0);
  }
  if (G == 0 && blockIdx.x == 0 && threadIdx.x == 0) row_ptr[0] = 0;
}

}  // namespace bignn

using namespace bignn;

extern "C" int64_t bignn_merge_build_workspace_bytes(int32_t G) {
  int nb = ceil_div(G > 0 ? G : 1, kScanTile);
  return (int64_t)nb * sizeof(I2) + 16;
}

extern "C" int bignn_merge_build(const int32_t* atom_ptr, const int32_t* nbr_ptr, const int32_t* nbr_idx,
                                 const float* x_all, int32_t F, const int32_t* rows, int32_t G,
                                 int32_t* seg_ptr, int32_t* edge_ptr, int32_t* row_ptr, int32_t* col_idx,
                                 int32_t* batch, float* x, int64_t* edge_index_i64, int64_t* batch_i64,
                                 int32_t A, int32_t E, void* workspace, int64_t workspace_bytes,
                                 void* stream) {
  if (G < 0 || A < 0 || E < 0 || F < 0) return BIGNN_EINVAL;
  if (!atom_ptr || !nbr_ptr || !nbr_idx || !seg_ptr || !edge_ptr || !row_ptr) return BIGNN_EINVAL;
  if (G > 0 && (!rows || (E > 0 && !col_idx))) return BIGNN_EINVAL;
  if (x && !x_all) return BIGNN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = ceil_div(G > 0 ? G : 1, kScanTile);
  I2* blk = nullptr;
  if (nb > 1) {
    if (!workspace || workspace_bytes < bignn_merge_build_workspace_bytes(G)) return BIGNN_EWORKSPACE;
    blk = (I2*)workspace;
    k_merge_blocksum<<<nb, kScanThreads, 0, st>>>(atom_ptr, nbr_ptr, rows, G, blk);
    k_merge_scanblocks<<<1, kScanThreads, 0, st>>>(blk, nb);
    BIGNN_LAUNCH_COUNT(2);
  }
  k_merge_scanfinal<<<nb, kScanThreads, 0, st>>>(atom_ptr, nbr_ptr, rows, G, blk, seg_ptr, edge_ptr);
  const int wpb = 8;
  int grid = ceil_div(G > 0 ? G : 1, wpb);
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  k_merge_fill<<<grid, wpb * 32, 0, st>>>(atom_ptr, nbr_ptr, nbr_idx, x_all, F, rows, G, seg_ptr, edge_ptr,
                                          row_ptr, col_idx, batch, x, edge_index_i64, batch_i64, E);
  BIGNN_LAUNCH_COUNT(2);
  (void)A;
  return last_launch_status();
}
