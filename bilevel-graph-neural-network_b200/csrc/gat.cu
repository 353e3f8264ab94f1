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

constexpr int GAT_MAXS = 4;   // columns per lane -> D <= 128

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }
__device__ __forceinline__ float lrelu_grad(float v, float slope) { return v > 0.f ? 1.f : slope; }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// p[n] = att_i . h[n],  q[n] = att_j . h[n]
__global__ void __launch_bounds__(256)
k_gat_scores(const float* __restrict__ H, int64_t ldh, int n, int D, const float* __restrict__ att,
             float* __restrict__ p, float* __restrict__ q) {
  const int lane = threadIdx.x % 32, wpb = blockDim.x / 32;
  for (int r = blockIdx.x * wpb + threadIdx.x / 32; r < n; r += gridDim.x * wpb) {
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

// m[g], z[g]: max and sum of exp(s - m) over the edges grouped at g (ROLE = group_target ? 0 : 1)
template <int ROLE>
__global__ void __launch_bounds__(256)
k_gat_group_stats(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx, int n,
                  const float* __restrict__ p, const float* __restrict__ q, float slope,
                  float* __restrict__ m, float* __restrict__ z) {
  const int lane = threadIdx.x % 32, wpb = blockDim.x / 32;
  for (int r = blockIdx.x * wpb + threadIdx.x / 32; r < n; r += gridDim.x * wpb) {
    const int k0 = row_ptr[r], k1 = row_ptr[r + 1];
    float mx = -INFINITY, sm = 0.f;
    for (int k = k0 + lane; k <= k1; k += 32) {           // k == k1 is the added self loop
      const int nb = k < k1 ? __ldg(col_idx + k) : r;
      if (k < k1 && nb == r) continue;                    // remove_self_loops
      const float s = edge_score<ROLE>(p, q, r, nb, slope);
      if (s > mx) { sm = sm * expf(mx - s) + 1.f; mx = s; }
      else sm += expf(s - mx);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const float os = __shfl_xor_sync(0xffffffffu, sm, o);
      const float nm = fmaxf(mx, om);
      const float a = mx == -INFINITY ? 0.f : sm * expf(mx - nm);
      const float b = om == -INFINITY ? 0.f : os * expf(om - nm);
      mx = nm;
      sm = a + b;
    }
    if (lane == 0) { m[r] = mx; z[r] = sm; }
  }
}

// out[r] = sum_nb alpha(r,nb) V[nb] (+ bias).  ROLE 0 = forward aggregation at targets (V = h);
// ROLE 1 = backward of it at sources (V = d out).
template <int ROLE>
__global__ void __launch_bounds__(256)
k_gat_gather(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx, int n, int D,
             const float* __restrict__ V, int64_t ldv, const float* __restrict__ p, const float* __restrict__ q,
             const float* __restrict__ m, const float* __restrict__ z, float slope, int group_target,
             const float* __restrict__ bias, float* __restrict__ out, int64_t ldo) {
  const int lane = threadIdx.x % 32, wpb = blockDim.x / 32;
  for (int r = blockIdx.x * wpb + threadIdx.x / 32; r < n; r += gridDim.x * wpb) {
    const int k0 = row_ptr[r], k1 = row_ptr[r + 1];
    float acc[GAT_MAXS];
#pragma unroll
    for (int s = 0; s < GAT_MAXS; ++s) acc[s] = 0.f;
    for (int k = k0; k <= k1; ++k) {
      const int nb = k < k1 ? __ldg(col_idx + k) : r;
      if (k < k1 && nb == r) continue;
      const int g = edge_group<ROLE>(r, nb, group_target);
      const float sc = edge_score<ROLE>(p, q, r, nb, slope);
      const float alpha = expf(sc - __ldg(m + g)) / (__ldg(z + g) + 1e-16f);
#pragma unroll
      for (int s = 0; s < GAT_MAXS; ++s) {
        const int c = lane + 32 * s;
        if (c < D) acc[s] = __fadd_rn(acc[s], __fmul_rn(alpha, __ldg(V + (int64_t)nb * ldv + c)));
      }
    }
#pragma unroll
    for (int s = 0; s < GAT_MAXS; ++s) {
      const int c = lane + 32 * s;
      if (c < D) out[(int64_t)r * ldo + c] = acc[s] + (bias ? __ldg(bias + c) : 0.f);
    }
  }
}

// t[r] = <A[r], B[r] - sub>     (softmax-backward group term)
__global__ void __launch_bounds__(256)
k_rowdot(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
         const float* __restrict__ sub, int n, int D, float* __restrict__ t) {
  const int lane = threadIdx.x % 32, wpb = blockDim.x / 32;
  for (int r = blockIdx.x * wpb + threadIdx.x / 32; r < n; r += gridDim.x * wpb) {
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
__global__ void __launch_bounds__(256)
k_gat_edge_grad(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx, int n, int D,
                const float* __restrict__ Own, int64_t ldown, const float* __restrict__ Oth, int64_t ldoth,
                const float* __restrict__ p, const float* __restrict__ q, const float* __restrict__ m,
                const float* __restrict__ z, const float* __restrict__ t, float slope, int group_target,
                float* __restrict__ out) {
  const int lane = threadIdx.x % 32, wpb = blockDim.x / 32;
  for (int r = blockIdx.x * wpb + threadIdx.x / 32; r < n; r += gridDim.x * wpb) {
    const int k0 = row_ptr[r], k1 = row_ptr[r + 1];
    float own[GAT_MAXS];
#pragma unroll
    for (int s = 0; s < GAT_MAXS; ++s) {
      const int c = lane + 32 * s;
      own[s] = c < D ? __ldg(Own + (int64_t)r * ldown + c) : 0.f;
    }
    float acc = 0.f;
    for (int k = k0; k <= k1; ++k) {
      const int nb = k < k1 ? __ldg(col_idx + k) : r;
      if (k < k1 && nb == r) continue;
      float d = 0.f;
#pragma unroll
      for (int s = 0; s < GAT_MAXS; ++s) {
        const int c = lane + 32 * s;
        if (c < D) d = fmaf(own[s], __ldg(Oth + (int64_t)nb * ldoth + c), d);
      }
      d = warp_sum(d);
      const int g = edge_group<ROLE>(r, nb, group_target);
      const float pre = edge_pre<ROLE>(p, q, r, nb);
      const float alpha = expf(lrelu(pre, slope) - __ldg(m + g)) / (__ldg(z + g) + 1e-16f);
      acc += alpha * (d - __ldg(t + g)) * lrelu_grad(pre, slope);
    }
    if (lane == 0) out[r] = acc;
  }
}

// dH[r,:] = G[r,:] + dp[r] att_i + dq[r] att_j
__global__ void __launch_bounds__(256)
k_gat_combine(const float* __restrict__ G, int64_t ldg, const float* __restrict__ dp, const float* __restrict__ dq,
              const float* __restrict__ att, int n, int D, float* __restrict__ dH, int64_t ldd) {
  const int64_t total = (int64_t)n * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / D), c = (int)(i % D);
    dH[(int64_t)r * ldd + c] = G[(int64_t)r * ldg + c] + dp[r] * __ldg(att + c) + dq[r] * __ldg(att + D + c);
  }
}

static inline int warp_grid(int n) {
  int g = ceil_div(n, 8);
  const int cap = sm_count() * 8;
  return g > cap ? cap : (g < 1 ? 1 : g);
}

}  // namespace bignn

using namespace bignn;

// scratch layout (floats): p[n] q[n] m[n] z[n]  -- kept by the caller for the backward
extern "C" int bignn_gat_fwd(const int32_t* row_ptr, const int32_t* col_idx, int32_t n, int32_t D, const float* H,
                             int64_t ldh, const float* att, const float* bias, float negative_slope,
                             int32_t group_target, float* out, int64_t ldo, float* scratch4n, void* stream) {
  if (n < 0 || D < 0) return BIGNN_EINVAL;
  if (n == 0 || D == 0) return 0;
  if (D > 32 * GAT_MAXS || !row_ptr || !H || !att || !out || !scratch4n || ldh < D || ldo < D) return BIGNN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  float *p = scratch4n, *q = p + n, *m = q + n, *z = m + n;
  const int g = warp_grid(n);
  k_gat_scores<<<g, 256, 0, st>>>(H, ldh, n, D, att, p, q);
  if (group_target) k_gat_group_stats<0><<<g, 256, 0, st>>>(row_ptr, col_idx, n, p, q, negative_slope, m, z);
  else k_gat_group_stats<1><<<g, 256, 0, st>>>(row_ptr, col_idx, n, p, q, negative_slope, m, z);
  k_gat_gather<0><<<g, 256, 0, st>>>(row_ptr, col_idx, n, D, H, ldh, p, q, m, z, negative_slope, group_target, bias, out, ldo);
  BIGNN_LAUNCH_COUNT(3);
  return last_launch_status();
}

// dH[n,D] and the two score-gradient vectors dpq[2n] (d att = [dp^T H ; dq^T H] is a GEMM for the caller)
extern "C" int bignn_gat_bwd(const int32_t* row_ptr, const int32_t* col_idx, int32_t n, int32_t D, const float* H,
                             int64_t ldh, const float* att, const float* bias, float negative_slope,
                             int32_t group_target, const float* out, int64_t ldo, const float* dOut, int64_t lddo,
                             const float* scratch4n, float* dH, int64_t lddh, float* dpq, float* workspace,
                             int64_t workspace_bytes, void* stream) {
  if (n < 0 || D < 0) return BIGNN_EINVAL;
  if (n == 0 || D == 0) return 0;
  if (D > 32 * GAT_MAXS || !row_ptr || !H || !att || !out || !dOut || !scratch4n || !dH || !dpq) return BIGNN_EINVAL;
  if (ldh < D || ldo < D || lddo < D || lddh < D) return BIGNN_EINVAL;
  const int64_t need = ((int64_t)n * D + n) * (int64_t)sizeof(float);
  if (!workspace || workspace_bytes < need) return BIGNN_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const float *p = scratch4n, *q = p + n, *m = q + n, *z = m + n;
  float* G = workspace;
  float* t = G + (int64_t)n * D;
  float *dp = dpq, *dq = dpq + n;
  const int g = warp_grid(n);
  k_gat_gather<1><<<g, 256, 0, st>>>(row_ptr, col_idx, n, D, dOut, lddo, p, q, m, z, negative_slope, group_target, nullptr, G, D);
  if (group_target) k_rowdot<<<g, 256, 0, st>>>(dOut, lddo, out, ldo, bias, n, D, t);
  else k_rowdot<<<g, 256, 0, st>>>(H, ldh, G, D, nullptr, n, D, t);
  k_gat_edge_grad<0><<<g, 256, 0, st>>>(row_ptr, col_idx, n, D, dOut, lddo, H, ldh, p, q, m, z, t, negative_slope, group_target, dp);
  k_gat_edge_grad<1><<<g, 256, 0, st>>>(row_ptr, col_idx, n, D, H, ldh, dOut, lddo, p, q, m, z, t, negative_slope, group_target, dq);
  int eg = (int)ceil_div<int64_t>((int64_t)n * D, 256);
  const int cap = sm_count() * 8;
  if (eg > cap) eg = cap;
  k_gat_combine<<<eg, 256, 0, st>>>(G, D, dp, dq, att, n, D, dH, lddh);
  BIGNN_LAUNCH_COUNT(5);
  return last_launch_status();
}

extern "C" int64_t bignn_gat_bwd_workspace_bytes(int32_t n, int32_t D) {
  if (n <= 0 || D <= 0) return 0;
  return ((int64_t)n * D + n) * (int64_t)sizeof(float);
}
