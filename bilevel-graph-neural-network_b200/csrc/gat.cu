// One-head GAT edge-softmax message passing, forward and backward (see include/bignn_b200.h).
// Replaces PyG 1.1.2 GATConv (reference model/layers.py:32-34,54):
//   h = x W;  s_e = leaky_relu(att_i . h_target + att_j . h_source)  for every edge source->target
//   (self loops removed, then one added per node);  alpha = softmax of s over the edges that share
//   the grouping node (group = 0: the SOURCE, the 1.1.x behaviour; group = 1: the TARGET, >= 1.2);
//   out_target = sum_e alpha_e h_source + bias.
// All graphs on the path are symmetric, so row r of the CSR lists both the sources of r's in-edges
// and the targets of its out-edges; every pass below is therefore row-parallel (one warp per row,
// neighbours in ascending order, self loop last) with gathers only -- no scatter, no atomics.
// Per-node scalars p = att_i.h, q = att_j.h turn the per-edge score into two scalar gathers.
#include "common.cuh"

namespace bignn {


__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }
__device__ __forceinline__ float lrelu_grad(float v, float slope) { return v > 0.f ? 1.f : slope; }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// p[n] = att_i . h[n],  q[n] = att_j . h[n]
// (several edge types are batched as one block-diagonal graph of n / n_block blocks; block b uses
//  att[b] and bias[b])
__global__ void __launch_bounds__(256)
k_gat_scores(const float* __restrict__ H, int64_t ldh, int n, int D, const float* __restrict__ att_all, int n_block,
             float* __restrict__ p, float* __restrict__ q) {
  const int lane = threadIdx.x % 32, wpb = blockDim.x / 32;
  for (int r = blockIdx.x * wpb + threadIdx.x / 32; r < n; r += gridDim.x * wpb) {
    const float* att = att_all + (int64_t)(r / n_block) * 2 * D;
    float a = 0.f, b = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float v = __ldg(H + (int64_t)r * ldh + c);
      a = fmaf(v, __ldg(att + c), a);
      b = fmaf(v, __ldg(att + D + c), b);
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) { p[r] = a; q[r] = b; }
  }
}

// score of the edge between row r and neighbour nb.  ROLE 0: r is the target (nb -> r);
// ROLE 1: r is the source (r -> nb).
template <int ROLE>
__device__ __forceinline__ float edge_score(const float* __restrict__ p, const float* __restrict__ q, int r, int nb,
                                            float slope) {
  return ROLE == 0 ? lrelu(__ldg(p + r) + __ldg(q + nb), slope) : lrelu(__ldg(p + nb) + __ldg(q + r), slope);
}
template <int ROLE>
__device__ __forceinline__ float edge_pre(const float* __restrict__ p, const float* __restrict__ q, int r, int nb) {
  return ROLE == 0 ? __ldg(p + r) + __ldg(q + nb) : __ldg(p + nb) + __ldg(q + r);
}
// grouping node of that edge: group_target ? the target : the source
template <int ROLE>
__device__ __forceinline__ int edge_group(int r, int nb, int group_target) {
  return ROLE == 0 ? (group_target ? r : nb) : (group_target ? nb : r);
}

// ---- work items -----------------------------------------------------------------------------------
// Interaction graphs have hub drugs (DrugCombo synergy graph: max degree 1 401, p99 226), so every
// neighbour pass runs over WORK ITEMS of at most `seg` neighbours (the same plan as
// bignn_spmm_planned_f32): a 16-lane sub-warp owns one item, 128-bit loads for the 64-float rows.
// Rows made of one item are finished in place (including their self-loop edge); rows made of several
// items leave per-item partials that a second kernel combines in item order (deterministic).
struct ItemPlan {
  const int32_t* item_ptr;
  const int32_t* item_row;
  const int32_t* multi_rows;
  int n_items, n_multi, seg;
};

__device__ __forceinline__ unsigned half_mask() { return (threadIdx.x & 16) ? 0xffff0000u : 0x0000ffffu; }
__device__ __forceinline__ float half_sum(float v, unsigned mask) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}
__device__ __forceinline__ void softmax_merge(float& m, float& s, float om, float os) {
  const float nm = fmaxf(m, om);
  const float a = m == -INFINITY ? 0.f : s * expf(m - nm);
  const float b = om == -INFINITY ? 0.f : os * expf(om - nm);
  m = nm;
  s = a + b;
}

// (max, sum exp) per grouping node.  ROLE = group_target ? 0 : 1
template <int ROLE>
__global__ void __launch_bounds__(256)
k_gat_stats_items(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx, ItemPlan pl,
                  const float* __restrict__ p, const float* __restrict__ q, float slope,
                  float* __restrict__ m, float* __restrict__ z, float* __restrict__ part) {
  const int lane = threadIdx.x & 15;
  const unsigned mask = half_mask();
  for (int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 4; item < pl.n_items; item += (gridDim.x * blockDim.x) >> 4) {
    const int r = __ldg(pl.item_row + item);
    const int i0 = __ldg(pl.item_ptr + r), i1 = __ldg(pl.item_ptr + r + 1);
    const int k0 = __ldg(row_ptr + r) + (item - i0) * pl.seg;
    const int k1 = min(k0 + pl.seg, __ldg(row_ptr + r + 1));
    float mx = -INFINITY, sm = 0.f;
    for (int k = k0 + lane; k < k1; k += 16) {
      const int nb = __ldg(col_idx + k);
      if (nb == r) continue;
      const float s = edge_score<ROLE>(p, q, r, nb, slope);
      softmax_merge(mx, sm, s, 1.f);
    }
    const bool single = (i1 - i0) == 1;
    if (single && lane == 0) softmax_merge(mx, sm, edge_score<ROLE>(p, q, r, r, slope), 1.f);   // self loop
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(mask, mx, o), os = __shfl_xor_sync(mask, sm, o);
      softmax_merge(mx, sm, om, os);
    }
    if (lane == 0) {
      if (single) { m[r] = mx; z[r] = sm; }
      else { part[2 * item] = mx; part[2 * item + 1] = sm; }
    }
  }
}

template <int ROLE>
__global__ void __launch_bounds__(256)
k_gat_stats_multi(ItemPlan pl, const float* __restrict__ p, const float* __restrict__ q, float slope,
                  float* __restrict__ m, float* __restrict__ z, const float* __restrict__ part) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pl.n_multi) return;
  const int r = pl.multi_rows[i];
  float mx = -INFINITY, sm = 0.f;
  for (int it = pl.item_ptr[r]; it < pl.item_ptr[r + 1]; ++it) softmax_merge(mx, sm, part[2 * it], part[2 * it + 1]);
  softmax_merge(mx, sm, edge_score<ROLE>(p, q, r, r, slope), 1.f);
  m[r] = mx;
  z[r] = sm;
}

template <int ROLE>
__device__ __forceinline__ float edge_alpha(const float* __restrict__ p, const float* __restrict__ q,
                                            const float* __restrict__ m, const float* __restrict__ z, int r, int nb,
                                            float slope, int group_target) {
  const int g = edge_group<ROLE>(r, nb, group_target);
  return expf(edge_score<ROLE>(p, q, r, nb, slope) - __ldg(m + g)) / (__ldg(z + g) + 1e-16f);
}

// out[r] = sum_nb alpha(r,nb) V[nb] (+ alpha(r,r) V[r] + bias).  ROLE 0 = forward aggregation at the
// targets (V = h); ROLE 1 = its backward at the sources (V = d out).
template <int ROLE>
__global__ void __launch_bounds__(256)
k_gat_gather_items(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx, ItemPlan pl, int D4,
                   const float* __restrict__ V, int64_t ldv, const float* __restrict__ p, const float* __restrict__ q,
                   const float* __restrict__ m, const float* __restrict__ z, float slope, int group_target,
                   const float* __restrict__ bias_all, int n_block, float* __restrict__ out, int64_t ldo,
                   float* __restrict__ part) {
  constexpr int U = 4;
  const int lane = threadIdx.x & 15;
  const bool active = lane < D4;
  for (int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 4; item < pl.n_items; item += (gridDim.x * blockDim.x) >> 4) {
    const int r = __ldg(pl.item_row + item);
    const int i0 = __ldg(pl.item_ptr + r), i1 = __ldg(pl.item_ptr + r + 1);
    const int k0 = __ldg(row_ptr + r) + (item - i0) * pl.seg;
    const int k1 = min(k0 + pl.seg, __ldg(row_ptr + r + 1));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = k0; k < k1; k += U) {
      int c[U];
      float w[U];
      float4 val[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        c[u] = (k + u < k1) ? __ldg(col_idx + k + u) : -1;
        if (c[u] == r) c[u] = -1;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        w[u] = c[u] >= 0 ? edge_alpha<ROLE>(p, q, m, z, r, c[u], slope, group_target) : 0.f;
        val[u] = (c[u] >= 0 && active) ? ldg4(V + (int64_t)c[u] * ldv + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (c[u] >= 0) {
          acc.x = __fadd_rn(acc.x, __fmul_rn(w[u], val[u].x)); acc.y = __fadd_rn(acc.y, __fmul_rn(w[u], val[u].y));
          acc.z = __fadd_rn(acc.z, __fmul_rn(w[u], val[u].z)); acc.w = __fadd_rn(acc.w, __fmul_rn(w[u], val[u].w));
        }
    }
    if (!active) continue;
    if (i1 - i0 == 1) {
      const float ws = edge_alpha<ROLE>(p, q, m, z, r, r, slope, group_target);
      const float4 sv = ldg4(V + (int64_t)r * ldv + 4 * lane);
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias_all) b = ldg4(bias_all + (int64_t)(r / n_block) * 4 * D4 + 4 * lane);
      acc.x = __fadd_rn(acc.x, __fmul_rn(ws, sv.x)) + b.x; acc.y = __fadd_rn(acc.y, __fmul_rn(ws, sv.y)) + b.y;
      acc.z = __fadd_rn(acc.z, __fmul_rn(ws, sv.z)) + b.z; acc.w = __fadd_rn(acc.w, __fmul_rn(ws, sv.w)) + b.w;
      st4(out + (int64_t)r * ldo + 4 * lane, acc);
    } else {
      st4(part + ((int64_t)item * D4 + lane) * 4, acc);
    }
  }
}

template <int ROLE>
__global__ void __launch_bounds__(256)
k_gat_gather_multi(ItemPlan pl, int D4, const float* __restrict__ V, int64_t ldv, const float* __restrict__ p,
                   const float* __restrict__ q, const float* __restrict__ m, const float* __restrict__ z, float slope,
                   int group_target, const float* __restrict__ bias_all, int n_block, float* __restrict__ out,
                   int64_t ldo, const float* __restrict__ part) {
  const int lane = threadIdx.x & 15;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 4; i < pl.n_multi; i += (gridDim.x * blockDim.x) >> 4) {
    if (lane >= D4) continue;
    const int r = pl.multi_rows[i];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = pl.item_ptr[r]; it < pl.item_ptr[r + 1]; ++it) {
      const float4 v = ldg4(part + ((int64_t)it * D4 + lane) * 4);
      acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y); acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
    }
    const float ws = edge_alpha<ROLE>(p, q, m, z, r, r, slope, group_target);
    const float4 sv = ldg4(V + (int64_t)r * ldv + 4 * lane);
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias_all) b = ldg4(bias_all + (int64_t)(r / n_block) * 4 * D4 + 4 * lane);
    acc.x = __fadd_rn(acc.x, __fmul_rn(ws, sv.x)) + b.x; acc.y = __fadd_rn(acc.y, __fmul_rn(ws, sv.y)) + b.y;
    acc.z = __fadd_rn(acc.z, __fmul_rn(ws, sv.z)) + b.z; acc.w = __fadd_rn(acc.w, __fmul_rn(ws, sv.w)) + b.w;
    st4(out + (int64_t)r * ldo + 4 * lane, acc);
  }
}

// t[r] = <A[r], B[r] - sub>     (softmax-backward group term)
__global__ void __launch_bounds__(256)
k_rowdot(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
         const float* __restrict__ sub_all, int n_block, int n, int D, float* __restrict__ t) {
  const int lane = threadIdx.x % 32, wpb = blockDim.x / 32;
  for (int r = blockIdx.x * wpb + threadIdx.x / 32; r < n; r += gridDim.x * wpb) {
    const float* sub = sub_all ? sub_all + (int64_t)(r / n_block) * D : nullptr;
    float a = 0.f;
    for (int c = lane; c < D; c += 32)
      a = fmaf(__ldg(A + (int64_t)r * lda + c), __ldg(B + (int64_t)r * ldb + c) - (sub ? __ldg(sub + c) : 0.f), a);
    a = warp_sum(a);
    if (lane == 0) t[r] = a;
  }
}

// d score sums.  ROLE 0: dp[r] = sum over in-edges (nb -> r) of u_e, own = d out, gathered = h.
//                ROLE 1: dq[r] = sum over out-edges (r -> nb) of u_e, own = h, gathered = d out.
// u_e = alpha_e (d alpha_e - t[group]) * leaky_relu'(pre_e),  d alpha_e = <d out_target, h_source>
template <int ROLE>
__device__ __forceinline__ float edge_u(const float* __restrict__ p, const float* __restrict__ q,
                                        const float* __restrict__ m, const float* __restrict__ z,
                                        const float* __restrict__ t, int r, int nb, float d, float slope,
                                        int group_target) {
  const int g = edge_group<ROLE>(r, nb, group_target);
  const float pre = edge_pre<ROLE>(p, q, r, nb);
  const float alpha = expf(lrelu(pre, slope) - __ldg(m + g)) / (__ldg(z + g) + 1e-16f);
  return alpha * (d - __ldg(t + g)) * lrelu_grad(pre, slope);
}

template <int ROLE>
__global__ void __launch_bounds__(256)
k_gat_edge_grad_items(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx, ItemPlan pl, int D4,
                      const float* __restrict__ Own, int64_t ldown, const float* __restrict__ Oth, int64_t ldoth,
                      const float* __restrict__ p, const float* __restrict__ q, const float* __restrict__ m,
                      const float* __restrict__ z, const float* __restrict__ t, float slope, int group_target,
                      float* __restrict__ out, float* __restrict__ part) {
  constexpr int U = 4;
  const int lane = threadIdx.x & 15;
  const unsigned mask = half_mask();
  const bool active = lane < D4;
  for (int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 4; item < pl.n_items; item += (gridDim.x * blockDim.x) >> 4) {
    const int r = __ldg(pl.item_row + item);
    const int i0 = __ldg(pl.item_ptr + r), i1 = __ldg(pl.item_ptr + r + 1);
    const int k0 = __ldg(row_ptr + r) + (item - i0) * pl.seg;
    const int k1 = min(k0 + pl.seg, __ldg(row_ptr + r + 1));
    const float4 own = active ? ldg4(Own + (int64_t)r * ldown + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    float acc = 0.f;
    for (int k = k0; k < k1; k += U) {
      int c[U];
      float d[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        c[u] = (k + u < k1) ? __ldg(col_idx + k + u) : -1;
        if (c[u] == r) c[u] = -1;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c[u] >= 0 && active) v = ldg4(Oth + (int64_t)c[u] * ldoth + 4 * lane);
        d[u] = fmaf(own.x, v.x, fmaf(own.y, v.y, fmaf(own.z, v.z, own.w * v.w)));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) d[u] = half_sum(d[u], mask);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (c[u] >= 0) acc += edge_u<ROLE>(p, q, m, z, t, r, c[u], d[u], slope, group_target);
    }
    if (i1 - i0 == 1) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (active) v = ldg4(Oth + (int64_t)r * ldoth + 4 * lane);
      const float ds = half_sum(fmaf(own.x, v.x, fmaf(own.y, v.y, fmaf(own.z, v.z, own.w * v.w))), mask);
      if (lane == 0) out[r] = acc + edge_u<ROLE>(p, q, m, z, t, r, r, ds, slope, group_target);
    } else if (lane == 0) {
      part[item] = acc;
    }
  }
}

template <int ROLE>
__global__ void __launch_bounds__(256)
k_gat_edge_grad_multi(ItemPlan pl, int D4, const float* __restrict__ Own, int64_t ldown,
                      const float* __restrict__ Oth, int64_t ldoth, const float* __restrict__ p,
                      const float* __restrict__ q, const float* __restrict__ m, const float* __restrict__ z,
                      const float* __restrict__ t, float slope, int group_target, float* __restrict__ out,
                      const float* __restrict__ part) {
  const int lane = threadIdx.x & 15;
  const unsigned mask = half_mask();
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const bool valid = i < pl.n_multi;
  const int r = valid ? pl.multi_rows[i] : 0;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  if (valid && lane < D4) { a = ldg4(Own + (int64_t)r * ldown + 4 * lane); b = ldg4(Oth + (int64_t)r * ldoth + 4 * lane); }
  const float ds = half_sum(fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))), mask);
  if (valid && lane == 0) {
    float acc = 0.f;
    for (int it = pl.item_ptr[r]; it < pl.item_ptr[r + 1]; ++it) acc += part[it];
    out[r] = acc + edge_u<ROLE>(p, q, m, z, t, r, r, ds, slope, group_target);
  }
}

// dH[r,:] = G[r,:] + dp[r] att_i + dq[r] att_j
__global__ void __launch_bounds__(256)
k_gat_combine(const float* __restrict__ G, int64_t ldg, const float* __restrict__ dp, const float* __restrict__ dq,
              const float* __restrict__ att_all, int n_block, int n, int D, float* __restrict__ dH, int64_t ldd) {
  const int64_t total = (int64_t)n * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / D), c = (int)(i % D);
    const float* att = att_all + (int64_t)(r / n_block) * 2 * D;
    dH[(int64_t)r * ldd + c] = G[(int64_t)r * ldg + c] + dp[r] * __ldg(att + c) + dq[r] * __ldg(att + D + c);
  }
}

static inline int warp_grid(int n) {
  int g = ceil_div(n, 8);
  const int cap = sm_count() * 8;
  return g > cap ? cap : (g < 1 ? 1 : g);
}

static inline int item_grid(int n_items) {
  int g = ceil_div(n_items, 16);
  const int cap = sm_count() * 8;
  return g > cap ? cap : (g < 1 ? 1 : g);
}

static int gat_check(int n, int D, const int32_t* row_ptr, const int32_t* item_ptr, const int32_t* item_row, int n_items,
                     int seg, const int32_t* multi_rows, int n_multi) {
  if (n < 0 || D < 0 || n_items < 0 || n_multi < 0 || seg <= 0) return BIGNN_EINVAL;
  if (D > 64 || (D & 3)) return BIGNN_EINVAL;
  if (!row_ptr || !item_ptr || !item_row || (n_multi > 0 && !multi_rows)) return BIGNN_EINVAL;
  return 0;
}

}  // namespace bignn

using namespace bignn;

extern "C" int64_t bignn_gat_fwd_workspace_bytes(int32_t n_items, int32_t D) {
  if (n_items <= 0 || D <= 0) return 0;
  return (int64_t)n_items * (D + 2) * (int64_t)sizeof(float) + 64;
}

// scratch layout (floats): p[n] q[n] m[n] z[n]  -- kept by the caller for the backward
extern "C" int bignn_gat_fwd(const int32_t* row_ptr, const int32_t* col_idx, const int32_t* item_ptr,
                             const int32_t* item_row, int32_t n_items, int32_t seg, const int32_t* multi_rows,
                             int32_t n_multi, int32_t n, int32_t n_block, int32_t D, const float* H, int64_t ldh,
                             const float* att, const float* bias, float negative_slope, int32_t group_target,
                             float* out, int64_t ldo, float* scratch4n, void* workspace, int64_t workspace_bytes,
                             void* stream) {
  if (n_block <= 0 || (n > 0 && n % n_block)) return BIGNN_EINVAL;
  int rc = gat_check(n, D, row_ptr, item_ptr, item_row, n_items, seg, multi_rows, n_multi);
  if (rc) return rc;
  if (n == 0 || D == 0) return 0;
  if (!H || !att || !out || !scratch4n || ldh < D || ldo < D) return BIGNN_EINVAL;
  if ((ldh & 3) || (ldo & 3) || !aligned16(H) || !aligned16(out) || (bias && !aligned16(bias))) return BIGNN_EALIGN;
  if (n_multi > 0 && (!workspace || workspace_bytes < bignn_gat_fwd_workspace_bytes(n_items, D))) return BIGNN_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float *p = scratch4n, *q = p + n, *m = q + n, *z = m + n;
  float* part_s = (float*)workspace;                       // [n_items][2]
  float* part_v = part_s ? part_s + (((int64_t)2 * n_items + 3) & ~(int64_t)3) : nullptr;   // [n_items][D], 16-byte aligned
  ItemPlan pl{item_ptr, item_row, multi_rows, n_items, n_multi, seg};
  const int g = item_grid(n_items), gm = item_grid(n_multi > 0 ? n_multi : 1);
  k_gat_scores<<<warp_grid(n), 256, 0, st>>>(H, ldh, n, D, att, n_block, p, q);
  if (group_target) {
    k_gat_stats_items<0><<<g, 256, 0, st>>>(row_ptr, col_idx, pl, p, q, negative_slope, m, z, part_s);
    if (n_multi) k_gat_stats_multi<0><<<ceil_div(n_multi, 256), 256, 0, st>>>(pl, p, q, negative_slope, m, z, part_s);
  } else {
    k_gat_stats_items<1><<<g, 256, 0, st>>>(row_ptr, col_idx, pl, p, q, negative_slope, m, z, part_s);
    if (n_multi) k_gat_stats_multi<1><<<ceil_div(n_multi, 256), 256, 0, st>>>(pl, p, q, negative_slope, m, z, part_s);
  }
  k_gat_gather_items<0><<<g, 256, 0, st>>>(row_ptr, col_idx, pl, D / 4, H, ldh, p, q, m, z, negative_slope, group_target,
                                            bias, n_block, out, ldo, part_v);
  if (n_multi)
    k_gat_gather_multi<0><<<gm, 256, 0, st>>>(pl, D / 4, H, ldh, p, q, m, z, negative_slope, group_target, bias, n_block,
                                              out, ldo, part_v);
  BIGNN_LAUNCH_COUNT(3 + (n_multi ? 2 : 0));
  return last_launch_status();
}

extern "C" int64_t bignn_gat_bwd_workspace_bytes(int32_t n, int32_t D, int32_t n_items) {
  if (n <= 0 || D <= 0) return 0;
  return ((int64_t)n * D + n + (int64_t)n_items * (D + 1) + 16) * (int64_t)sizeof(float);
}

// dH[n,D] and the two score-gradient vectors dpq[2n] (d att = [dp^T H ; dq^T H] is a GEMM for the caller)
extern "C" int bignn_gat_bwd(const int32_t* row_ptr, const int32_t* col_idx, const int32_t* item_ptr,
                             const int32_t* item_row, int32_t n_items, int32_t seg, const int32_t* multi_rows,
                             int32_t n_multi, int32_t n, int32_t n_block, int32_t D, const float* H, int64_t ldh,
                             const float* att, const float* bias, float negative_slope, int32_t group_target,
                             const float* out, int64_t ldo, const float* dOut, int64_t lddo, const float* scratch4n,
                             float* dH, int64_t lddh, float* dpq, void* workspace, int64_t workspace_bytes,
                             void* stream) {
  if (n_block <= 0 || (n > 0 && n % n_block)) return BIGNN_EINVAL;
  int rc = gat_check(n, D, row_ptr, item_ptr, item_row, n_items, seg, multi_rows, n_multi);
  if (rc) return rc;
  if (n == 0 || D == 0) return 0;
  if (!H || !att || !out || !dOut || !scratch4n || !dH || !dpq) return BIGNN_EINVAL;
  if (ldh < D || ldo < D || lddo < D || lddh < D) return BIGNN_EINVAL;
  if ((ldh & 3) || (lddo & 3) || !aligned16(H) || !aligned16(dOut)) return BIGNN_EALIGN;
  if (!workspace || workspace_bytes < bignn_gat_bwd_workspace_bytes(n, D, n_items)) return BIGNN_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const float *p = scratch4n, *q = p + n, *m = q + n, *z = m + n;
  float* G = (float*)workspace;                                   // [n][D]
  float* part_v = G + (int64_t)n * D;                             // [n_items][D]
  float* t = part_v + (int64_t)n_items * D;                       // [n]
  float* part_s = t + n;                                          // [n_items]
  float *dp = dpq, *dq = dpq + n;
  ItemPlan pl{item_ptr, item_row, multi_rows, n_items, n_multi, seg};
  const int g = item_grid(n_items), gm = item_grid(n_multi > 0 ? n_multi : 1), wg = warp_grid(n);
  k_gat_gather_items<1><<<g, 256, 0, st>>>(row_ptr, col_idx, pl, D / 4, dOut, lddo, p, q, m, z, negative_slope,
                                            group_target, nullptr, n_block, G, D, part_v);
  if (n_multi)
    k_gat_gather_multi<1><<<gm, 256, 0, st>>>(pl, D / 4, dOut, lddo, p, q, m, z, negative_slope, group_target, nullptr,
                                              n_block, G, D, part_v);
  if (group_target) k_rowdot<<<wg, 256, 0, st>>>(dOut, lddo, out, ldo, bias, n_block, n, D, t);
  else k_rowdot<<<wg, 256, 0, st>>>(H, ldh, G, D, nullptr, n_block, n, D, t);
  k_gat_edge_grad_items<0><<<g, 256, 0, st>>>(row_ptr, col_idx, pl, D / 4, dOut, lddo, H, ldh, p, q, m, z, t,
                                               negative_slope, group_target, dp, part_s);
  if (n_multi)
    k_gat_edge_grad_multi<0><<<gm, 256, 0, st>>>(pl, D / 4, dOut, lddo, H, ldh, p, q, m, z, t, negative_slope,
                                                 group_target, dp, part_s);
  k_gat_edge_grad_items<1><<<g, 256, 0, st>>>(row_ptr, col_idx, pl, D / 4, H, ldh, dOut, lddo, p, q, m, z, t,
                                               negative_slope, group_target, dq, part_s);
  if (n_multi)
    k_gat_edge_grad_multi<1><<<gm, 256, 0, st>>>(pl, D / 4, H, ldh, dOut, lddo, p, q, m, z, t, negative_slope,
                                                 group_target, dq, part_s);
  int eg = (int)ceil_div<int64_t>((int64_t)n * D, 256);
  const int cap = sm_count() * 8;
  if (eg > cap) eg = cap;
  k_gat_combine<<<eg, 256, 0, st>>>(G, D, dp, dq, att, n_block, n, D, dH, lddh);
  BIGNN_LAUNCH_COUNT(5 + (n_multi ? 3 : 0));
  return last_launch_status();
}
