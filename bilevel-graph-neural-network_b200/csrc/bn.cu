// Segmented train-mode BatchNorm1d (see include/bignn_b200.h); replaces
// torch.nn.BatchNorm1d at model/layers.py:57 for a whole all-drug pass at once:
// segment s = rows [seg_row_ptr[s], seg_row_ptr[s+1]) = one 128-graph chunk of
// src/train.py:62-71, normalised with its own batch statistics.
//
// Deterministic two-level reductions (no float atomics): each (segment, part)
// CTA accumulates sum / sum-of-squares in fp64 over its contiguous row range,
// a finalize kernel adds the parts in part order.  Running buffers are advanced
// sequentially in segment order in fp64 (as torch's CPU kernel does) so the 11
// momentum updates per step of the reference are reproduced.
// HBM-bound: forward reads X twice (stats, apply) and writes Y once.
#include <cooperative_groups.h>
#include <cstdlib>
#include "common.cuh"

namespace bignn {

constexpr int BN_ROWS = 8;

// seg_row_ptr == nullptr: one segment made of rows [0, rows) (the row-partitioned upper level)
__device__ __forceinline__ void part_range(const int32_t* __restrict__ seg_row_ptr, int rows, int s, int p, int parts,
                                           int& r0, int& r1, int& n) {
  const int a = seg_row_ptr ? seg_row_ptr[s] : 0, b = seg_row_ptr ? seg_row_ptr[s + 1] : rows;
  n = b - a;
  r0 = a + (int)(((int64_t)n * p) / parts);
  r1 = a + (int)(((int64_t)n * (p + 1)) / parts);
}

// ws_a[(s*parts+p)*C+c] = sum_r f(r,c), ws_b = sum_r g(r,c)
// STATS: f = x, g = x*x.   BWD: f = dy, g = dy * (x-mean)*rstd
template <bool BWD>
__global__ void __launch_bounds__(256)
k_bn_reduce_part(const float* __restrict__ X, int64_t ldx, const float* __restrict__ dY, int64_t lddy,
                 const int32_t* __restrict__ seg_row_ptr, int C, int parts,
                 const float* __restrict__ mean, const float* __restrict__ rstd,
                 double* __restrict__ ws_a, double* __restrict__ ws_b, int rows) {
  __shared__ double ra[BN_ROWS][33];
  __shared__ double rb[BN_ROWS][33];
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const int c = blockIdx.y * 32 + tx;
  const int s = blockIdx.x / parts, p = blockIdx.x % parts;
  int r0, r1, n;
  part_range(seg_row_ptr, rows, s, p, parts, r0, r1, n);
  double a = 0.0, b = 0.0;
  if (c < C) {
    float mu = 0.f, rs = 0.f;
    if (BWD) { mu = mean[(int64_t)s * C + c]; rs = rstd[(int64_t)s * C + c]; }
    for (int r = r0 + ty; r < r1; r += BN_ROWS) {
      const float x = __ldg(X + (int64_t)r * ldx + c);
      if (BWD) {
        const float g = __ldg(dY + (int64_t)r * lddy + c);
        a += (double)g;
        b += (double)g * (double)((x - mu) * rs);
      } else {
        a += (double)x;
        b += (double)x * (double)x;
      }
    }
  }
  ra[ty][tx] = a;
  rb[ty][tx] = b;
  __syncthreads();
  if (ty == 0 && c < C) {
    double ta = 0.0, tb = 0.0;
#pragma unroll
    for (int i = 0; i < BN_ROWS; ++i) { ta += ra[i][tx]; tb += rb[i][tx]; }
    ws_a[(int64_t)blockIdx.x * C + c] = ta;
    ws_b[(int64_t)blockIdx.x * C + c] = tb;
  }
}

// sum of the `parts` partials of (segment s, channel c): 8 threads walk contiguous slices of the parts, slices added
// in slice order (deterministic; a single thread walking 256 parts of an S = 1 batch took 100 us)
constexpr int BN_FIN_SL = 8;
__device__ __forceinline__ void bn_sum_parts(const double* __restrict__ ws_a, const double* __restrict__ ws_b, int s,
                                             int c, int C, int parts, double (*red)[BN_FIN_SL][33], double& a, double& b) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int per = (parts + BN_FIN_SL - 1) / BN_FIN_SL;
  const int p0 = ty * per, p1 = min(parts, p0 + per);
  double ta = 0.0, tb = 0.0;
  if (c < C) {
#pragma unroll 8
    for (int p = p0; p < p1; ++p) {
      ta += ws_a[((int64_t)s * parts + p) * C + c];
      tb += ws_b[((int64_t)s * parts + p) * C + c];
    }
  }
  red[0][ty][tx] = ta;
  red[1][ty][tx] = tb;
  __syncthreads();
  a = 0.0; b = 0.0;
#pragma unroll
  for (int g = 0; g < BN_FIN_SL; ++g) { a += red[0][g][tx]; b += red[1][g][tx]; }
}

// grid (S, ceil(C / 32)), block 32 x BN_FIN_SL
__global__ void __launch_bounds__(32 * BN_FIN_SL)
k_bn_finalize(const double* __restrict__ ws_a, const double* __restrict__ ws_b,
              const int32_t* __restrict__ seg_row_ptr, int S, int C, int parts, float eps,
              float* __restrict__ mean, float* __restrict__ rstd,
              double* __restrict__ mean_d, double* __restrict__ varu_d) {
  __shared__ double red[2][BN_FIN_SL][33];
  const int s = blockIdx.x, c = blockIdx.y * 32 + (threadIdx.x & 31);
  double a, b;
  bn_sum_parts(ws_a, ws_b, s, c, C, parts, red, a, b);
  if (threadIdx.x >= 32 || c >= C) return;
  const int64_t i = (int64_t)s * C + c;
  const int n = seg_row_ptr[s + 1] - seg_row_ptr[s];
  double mu = 0.0, var = 0.0;
  if (n > 0) {
    mu = a / n;
    var = b / n - mu * mu;
    if (var < 0.0) var = 0.0;
  }
  mean[i] = (float)mu;
  rstd[i] = n > 0 ? (float)(1.0 / sqrt(var + (double)eps)) : 0.f;
  mean_d[i] = mu;
  varu_d[i] = n > 1 ? var * ((double)n / (double)(n - 1)) : var;
}

// The S momentum updates of one channel are a sequential, float-rounded recurrence (the reference runs one
// BatchNorm call per chunk, src/train.py:62-71) -- kept as such, but the operands of RB segments are fetched
// together so that the dependent chain waits on arithmetic latency, not on one L2 round trip per segment.
__global__ void __launch_bounds__(64)
k_bn_running(const double* __restrict__ mean_d, const double* __restrict__ varu_d,
             const int32_t* __restrict__ seg_row_ptr, int S, int C, double momentum,
             float* __restrict__ running_mean, float* __restrict__ running_var,
             int64_t* __restrict__ nbt) {
  constexpr int RB = 16;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ int live_count;
  if (threadIdx.x == 0) live_count = 0;
  __syncthreads();
  // after W further updates a term is scaled by keep^W: the first S - W segments (and the incoming buffer value) cannot
  // change the fp32 result once keep^W < 1e-13, so the recurrence starts W segments before the end (W = 264 at the
  // default momentum 0.1; the all-drug pass has 1 563 chunks at 200 k drugs)
  const double keep = 1.0 - momentum;
  int s_begin = 0;
  if (keep > 0.0 && keep < 1.0) {
    const int W = (int)ceil(log(1e-13) / log(keep)) + 1;
    if (S > W) s_begin = S - W;
  } else if (keep == 0.0 && S > 1) {
    s_begin = S - 1;
  }
  if (c < C) {
    float rm = s_begin > 0 ? 0.f : running_mean[c], rv = s_begin > 0 ? 0.f : running_var[c];
    for (int s0 = s_begin; s0 < S; s0 += RB) {
      double mu[RB], vu[RB];
      bool live[RB];
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        const int s = s0 + j;
        live[j] = s < S && (seg_row_ptr[s + 1] - seg_row_ptr[s] > 0);
        mu[j] = s < S ? mean_d[(int64_t)s * C + c] : 0.0;
        vu[j] = s < S ? varu_d[(int64_t)s * C + c] : 0.0;
      }
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        if (live[j]) {
          rm = (float)(momentum * mu[j] + keep * (double)rm);
          rv = (float)(momentum * vu[j] + keep * (double)rv);
        }
      }
    }
    running_mean[c] = rm;
    running_var[c] = rv;
  }
  if (blockIdx.x == 0 && nbt) {                                     // number of non-empty segments (integer: exact)
    int k = 0;
    for (int s = threadIdx.x; s < S; s += blockDim.x) k += (seg_row_ptr[s + 1] - seg_row_ptr[s] > 0);
    atomicAdd(&live_count, k);
    __syncthreads();
    if (threadIdx.x == 0) *nbt += live_count;
  }
}

__global__ void __launch_bounds__(256)
k_bn_apply(const float* __restrict__ X, int64_t ldx, float* __restrict__ Y, int64_t ldy,
           const int32_t* __restrict__ seg_row_ptr, int C, int parts,
           const float* __restrict__ gamma, const float* __restrict__ beta,
           const float* __restrict__ mean, const float* __restrict__ rstd, int rows) {
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const int c = blockIdx.y * 32 + tx;
  const int s = blockIdx.x / parts, p = blockIdx.x % parts;
  int r0, r1, n;
  part_range(seg_row_ptr, rows, s, p, parts, r0, r1, n);
  if (c >= C) return;
  const float alpha = rstd[(int64_t)s * C + c] * (gamma ? gamma[c] : 1.f);
  const float bt = (beta ? beta[c] : 0.f) - mean[(int64_t)s * C + c] * alpha;
  for (int r = r0 + ty; r < r1; r += BN_ROWS)
    Y[(int64_t)r * ldy + c] = fmaf(__ldg(X + (int64_t)r * ldx + c), alpha, bt);
}

__global__ void __launch_bounds__(256)
k_bn_eval(const float* __restrict__ X, int64_t ldx, float* __restrict__ Y, int64_t ldy, int rows, int C,
          const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
          const float* __restrict__ rm, const float* __restrict__ rv) {
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const int c = blockIdx.y * 32 + tx;
  if (c >= C) return;
  const float alpha = (float)(1.0 / sqrt((double)rv[c] + (double)eps)) * (gamma ? gamma[c] : 1.f);
  const float bt = (beta ? beta[c] : 0.f) - rm[c] * alpha;
  for (int r = blockIdx.x * BN_ROWS + ty; r < rows; r += gridDim.x * BN_ROWS)
    Y[(int64_t)r * ldy + c] = fmaf(__ldg(X + (int64_t)r * ldx + c), alpha, bt);
}

// per segment sums of the backward reduction; also the parameter gradients
__global__ void __launch_bounds__(32 * BN_FIN_SL)
k_bn_bwd_finalize(const double* __restrict__ ws_a, const double* __restrict__ ws_b, int S, int C, int parts,
                  double* __restrict__ seg_a, double* __restrict__ seg_b) {
  __shared__ double red[2][BN_FIN_SL][33];
  const int s = blockIdx.x, c = blockIdx.y * 32 + (threadIdx.x & 31);
  double a, b;
  bn_sum_parts(ws_a, ws_b, s, c, C, parts, red, a, b);
  if (threadIdx.x >= 32 || c >= C) return;
  seg_a[(int64_t)s * C + c] = a;
  seg_b[(int64_t)s * C + c] = b;
}

// dgamma / dbeta = sums over the segments: 32 channels x 32 contiguous slices of segments per CTA, slices added in
// slice order (one thread per channel walking all S segments took 188 us at S = 1 563 -- 1.5 ms of the C4 step)
__global__ void __launch_bounds__(1024)
k_bn_bwd_params(const double* __restrict__ seg_a, const double* __restrict__ seg_b, int S, int C,
                float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ double red[2][32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int per = (S + 31) / 32;
  const int s0 = ty * per, s1 = min(S, s0 + per);
  double a = 0.0, b = 0.0;
  if (c < C) {
#pragma unroll 8
    for (int s = s0; s < s1; ++s) { a += seg_a[(int64_t)s * C + c]; b += seg_b[(int64_t)s * C + c]; }
  }
  red[0][ty][tx] = a;
  red[1][ty][tx] = b;
  __syncthreads();
  if (ty < 2 && c < C) {
    double t = 0.0;
#pragma unroll 8
    for (int g = 0; g < 32; ++g) t += red[ty][g][tx];
    float* out = ty ? dgamma : dbeta;
    if (out) out[c] = (float)t;
  }
}

template <int IN_ACT>
__global__ void __launch_bounds__(256)
k_bn_bwd_apply(const float* __restrict__ X, int64_t ldx, const float* __restrict__ dY, int64_t lddy,
               float* __restrict__ dX, int64_t lddx, const int32_t* __restrict__ seg_row_ptr, int C, int parts,
               const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
               const double* __restrict__ seg_a, const double* __restrict__ seg_b, int rows, int64_t n_total) {
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const int c = blockIdx.y * 32 + tx;
  const int s = blockIdx.x / parts, p = blockIdx.x % parts;
  int r0, r1, n;
  part_range(seg_row_ptr, rows, s, p, parts, r0, r1, n);
  if (c >= C || n <= 0) return;
  const float mu = mean[(int64_t)s * C + c], rs = rstd[(int64_t)s * C + c];
  const double cnt = n_total > 0 ? (double)n_total : (double)n;    // n_total: the batch spans several ranks
  const float gmean = (float)(seg_a[(int64_t)s * C + c] / cnt);
  const float dgn = (float)(seg_b[(int64_t)s * C + c] / cnt);
  const float scale = rs * (gamma ? gamma[c] : 1.f);
  for (int r = r0 + ty; r < r1; r += BN_ROWS) {
    const float x = __ldg(X + (int64_t)r * ldx + c);
    const float xhat = (x - mu) * rs;
    const float g = __ldg(dY + (int64_t)r * lddy + c);
    float d = (g - gmean - xhat * dgn) * scale;
    // X is the OUTPUT of the activation in front of this BatchNorm (model/layers.py:55-57): its derivative is
    // applied here, on the value already in a register, instead of in a separate pass (same formulas as k_act_bwd)
    if (IN_ACT == BIGNN_ACT_RELU) d = x > 0.f ? d : 0.f;
    else if (IN_ACT == BIGNN_ACT_SIGMOID) d = d * ((1.0f - x) * x);
    else if (IN_ACT == BIGNN_ACT_TANH) d = d * (1.0f - x * x);
    dX[(int64_t)r * lddx + c] = d;
  }
}

// ---- row-partitioned batch (upper level sharded by source drug): one batch spans the rows of all ranks.
// sums[0..C) = sum of parts of ws_a, sums[C..2C) = of ws_b, parts added in part order
__global__ void __launch_bounds__(256)
k_bn_rows_sum_parts(const double* __restrict__ ws_a, const double* __restrict__ ws_b, int C, int parts,
                    double* __restrict__ sums) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double a = 0.0, b = 0.0;
#pragma unroll 8
  for (int p = 0; p < parts; ++p) { a += ws_a[(int64_t)p * C + c]; b += ws_b[(int64_t)p * C + c]; }
  sums[c] = a;
  sums[C + c] = b;
}

// statistics of the whole batch from the rank-summed (sum x, sum x^2); one momentum update
__global__ void __launch_bounds__(256)
k_bn_rows_finalize(const double* __restrict__ sums, int64_t n_total, int C, float eps, double momentum,
                   float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ running_mean,
                   float* __restrict__ running_var, int64_t* __restrict__ nbt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    double mu = 0.0, var = 0.0;
    if (n_total > 0) {
      mu = sums[c] / (double)n_total;
      var = sums[C + c] / (double)n_total - mu * mu;
      if (var < 0.0) var = 0.0;
    }
    mean[c] = (float)mu;
    rstd[c] = n_total > 0 ? (float)(1.0 / sqrt(var + (double)eps)) : 0.f;
    if (running_mean && running_var && n_total > 0) {
      const double varu = n_total > 1 ? var * ((double)n_total / (double)(n_total - 1)) : var;
      running_mean[c] = (float)(momentum * mu + (1.0 - momentum) * (double)running_mean[c]);
      running_var[c] = (float)(momentum * varu + (1.0 - momentum) * (double)running_var[c]);
    }
  }
  if (c == 0 && nbt && running_mean && n_total > 0) *nbt += 1;
}

// ---- 128-bit variants of the three streaming passes (C % 4 == 0, 16-byte aligned rows): `tpr` threads own one
// row (a float4 of channels each), 256/tpr rows per pass, four rows in flight per thread.  Same reduction tree as the
// scalar kernels: per-thread fp64 partials over the thread's rows (ascending), then the rows of the block in order.
template <bool BWD>
__global__ void __launch_bounds__(256)
k_bn_reduce_part_v4(const float* __restrict__ X, int64_t ldx, const float* __restrict__ dY, int64_t lddy,
                    const int32_t* __restrict__ seg_row_ptr, int C, int tpr, int parts,
                    const float* __restrict__ mean, const float* __restrict__ rstd,
                    double* __restrict__ ws_a, double* __restrict__ ws_b, int rows) {
  extern __shared__ double bn_sm[];                 // [2][rpb][C]
  const int lane = threadIdx.x % tpr, ty = threadIdx.x / tpr, rpb = blockDim.x / tpr;
  const int s = blockIdx.x / parts, p = blockIdx.x % parts;
  int r0, r1, n;
  part_range(seg_row_ptr, rows, s, p, parts, r0, r1, n);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
  const int c = 4 * lane;
  if (c < C) {
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), rs = mu;
    if (BWD) { mu = ldg4(mean + (int64_t)s * C + c); rs = ldg4(rstd + (int64_t)s * C + c); }
#pragma unroll 4
    for (int r = r0 + ty; r < r1; r += rpb) {
      const float4 x = ldg4(X + (int64_t)r * ldx + c);
      if (BWD) {
        const float4 g = ldg4(dY + (int64_t)r * lddy + c);
        a0 += (double)g.x; a1 += (double)g.y; a2 += (double)g.z; a3 += (double)g.w;
        b0 += (double)g.x * (double)((x.x - mu.x) * rs.x); b1 += (double)g.y * (double)((x.y - mu.y) * rs.y);
        b2 += (double)g.z * (double)((x.z - mu.z) * rs.z); b3 += (double)g.w * (double)((x.w - mu.w) * rs.w);
      } else {
        a0 += (double)x.x; a1 += (double)x.y; a2 += (double)x.z; a3 += (double)x.w;
        b0 += (double)x.x * (double)x.x; b1 += (double)x.y * (double)x.y;
        b2 += (double)x.z * (double)x.z; b3 += (double)x.w * (double)x.w;
      }
    }
    double* sa = bn_sm + (int64_t)ty * C + c;
    double* sb = bn_sm + (int64_t)(rpb + ty) * C + c;
    sa[0] = a0; sa[1] = a1; sa[2] = a2; sa[3] = a3;
    sb[0] = b0; sb[1] = b1; sb[2] = b2; sb[3] = b3;
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    double ta = 0.0, tb = 0.0;
    for (int i = 0; i < rpb; ++i) { ta += bn_sm[(int64_t)i * C + ch]; tb += bn_sm[(int64_t)(rpb + i) * C + ch]; }
    ws_a[(int64_t)blockIdx.x * C + ch] = ta;
    ws_b[(int64_t)blockIdx.x * C + ch] = tb;
  }
}

__global__ void __launch_bounds__(256)
k_bn_apply_v4(const float* __restrict__ X, int64_t ldx, float* __restrict__ Y, int64_t ldy,
              const int32_t* __restrict__ seg_row_ptr, int C, int tpr, int parts,
              const float* __restrict__ gamma, const float* __restrict__ beta,
              const float* __restrict__ mean, const float* __restrict__ rstd, int rows) {
  const int lane = threadIdx.x % tpr, ty = threadIdx.x / tpr, rpb = blockDim.x / tpr;
  const int s = blockIdx.x / parts, p = blockIdx.x % parts;
  int r0, r1, n;
  part_range(seg_row_ptr, rows, s, p, parts, r0, r1, n);
  const int c = 4 * lane;
  if (c >= C) return;
  const float4 rs = ldg4(rstd + (int64_t)s * C + c), mu = ldg4(mean + (int64_t)s * C + c);
  const float4 ga = gamma ? ldg4(gamma + c) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float4 be = beta ? ldg4(beta + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 al = make_float4(rs.x * ga.x, rs.y * ga.y, rs.z * ga.z, rs.w * ga.w);
  const float4 bt = make_float4(be.x - mu.x * al.x, be.y - mu.y * al.y, be.z - mu.z * al.z, be.w - mu.w * al.w);
#pragma unroll 4
  for (int r = r0 + ty; r < r1; r += rpb) {
    const float4 x = ldg4(X + (int64_t)r * ldx + c);
    st4(Y + (int64_t)r * ldy + c, make_float4(fmaf(x.x, al.x, bt.x), fmaf(x.y, al.y, bt.y), fmaf(x.z, al.z, bt.z),
                                              fmaf(x.w, al.w, bt.w)));
  }
}

template <int IN_ACT>
__device__ __forceinline__ float bn_bwd_one(float x, float g, float mu, float rs, float gmean, float dgn, float scale) {
  const float xhat = (x - mu) * rs;
  float d = (g - gmean - xhat * dgn) * scale;
  if (IN_ACT == BIGNN_ACT_RELU) d = x > 0.f ? d : 0.f;
  else if (IN_ACT == BIGNN_ACT_SIGMOID) d = d * ((1.0f - x) * x);
  else if (IN_ACT == BIGNN_ACT_TANH) d = d * (1.0f - x * x);
  return d;
}

template <int IN_ACT>
__global__ void __launch_bounds__(256)
k_bn_bwd_apply_v4(const float* __restrict__ X, int64_t ldx, const float* __restrict__ dY, int64_t lddy,
                  float* __restrict__ dX, int64_t lddx, const int32_t* __restrict__ seg_row_ptr, int C, int tpr,
                  int parts, const float* __restrict__ gamma, const float* __restrict__ mean,
                  const float* __restrict__ rstd, const double* __restrict__ seg_a,
                  const double* __restrict__ seg_b, int rows, int64_t n_total) {
  const int lane = threadIdx.x % tpr, ty = threadIdx.x / tpr, rpb = blockDim.x / tpr;
  const int s = blockIdx.x / parts, p = blockIdx.x % parts;
  int r0, r1, n;
  part_range(seg_row_ptr, rows, s, p, parts, r0, r1, n);
  const int c = 4 * lane;
  if (c >= C || n <= 0) return;
  const double cnt = n_total > 0 ? (double)n_total : (double)n;
  const float4 mu = ldg4(mean + (int64_t)s * C + c), rs = ldg4(rstd + (int64_t)s * C + c);
  const float4 ga = gamma ? ldg4(gamma + c) : make_float4(1.f, 1.f, 1.f, 1.f);
  const double* pa = seg_a + (int64_t)s * C + c;
  const double* pb = seg_b + (int64_t)s * C + c;
  const float4 gm = make_float4((float)(pa[0] / cnt), (float)(pa[1] / cnt), (float)(pa[2] / cnt), (float)(pa[3] / cnt));
  const float4 dg = make_float4((float)(pb[0] / cnt), (float)(pb[1] / cnt), (float)(pb[2] / cnt), (float)(pb[3] / cnt));
  const float4 sc = make_float4(rs.x * ga.x, rs.y * ga.y, rs.z * ga.z, rs.w * ga.w);
#pragma unroll 4
  for (int r = r0 + ty; r < r1; r += rpb) {
    const float4 x = ldg4(X + (int64_t)r * ldx + c);
    const float4 g = ldg4(dY + (int64_t)r * lddy + c);
    st4(dX + (int64_t)r * lddx + c,
        make_float4(bn_bwd_one<IN_ACT>(x.x, g.x, mu.x, rs.x, gm.x, dg.x, sc.x),
                    bn_bwd_one<IN_ACT>(x.y, g.y, mu.y, rs.y, gm.y, dg.y, sc.y),
                    bn_bwd_one<IN_ACT>(x.z, g.z, mu.z, rs.z, gm.z, dg.z, sc.z),
                    bn_bwd_one<IN_ACT>(x.w, g.w, mu.w, rs.w, gm.w, dg.w, sc.w)));
  }
}

// threads per row of the 128-bit kernels (power of two >= C/4), 0 = take the scalar kernels
static int bn_tpr(int C, const void* a, int64_t lda, const void* b, int64_t ldb, const void* c, int64_t ldc) {
  if ((C & 3) || C > 1024) return 0;
  if (a && (!aligned16(a) || (lda & 3))) return 0;
  if (b && (!aligned16(b) || (ldb & 3))) return 0;
  if (c && (!aligned16(c) || (ldc & 3))) return 0;
  int t = 1;
  while (t < C / 4) t <<= 1;
  return t;
}

static void launch_bn_reduce(bool bwd, int nblk, cudaStream_t st, const float* X, int64_t ldx, const float* dY,
                             int64_t lddy, const int32_t* seg_row_ptr, int C, int parts, const float* mean,
                             const float* rstd, double* ws_a, double* ws_b, int rows) {
  const int tpr = bn_tpr(C, X, ldx, dY, lddy, (bwd ? (const void*)mean : nullptr), 4);
  if (tpr > 0 && (!bwd || aligned16(rstd))) {
    const size_t sm = sizeof(double) * 2 * (256 / tpr) * C;
    if (bwd) k_bn_reduce_part_v4<true><<<nblk, 256, sm, st>>>(X, ldx, dY, lddy, seg_row_ptr, C, tpr, parts, mean, rstd, ws_a, ws_b, rows);
    else k_bn_reduce_part_v4<false><<<nblk, 256, sm, st>>>(X, ldx, nullptr, 0, seg_row_ptr, C, tpr, parts, nullptr, nullptr, ws_a, ws_b, rows);
    return;
  }
  dim3 grid(nblk, ceil_div(C, 32));
  if (bwd) k_bn_reduce_part<true><<<grid, 256, 0, st>>>(X, ldx, dY, lddy, seg_row_ptr, C, parts, mean, rstd, ws_a, ws_b, rows);
  else k_bn_reduce_part<false><<<grid, 256, 0, st>>>(X, ldx, nullptr, 0, seg_row_ptr, C, parts, nullptr, nullptr, ws_a, ws_b, rows);
}

static void launch_bn_apply(int nblk, cudaStream_t st, const float* X, int64_t ldx, float* Y, int64_t ldy,
                            const int32_t* seg_row_ptr, int C, int parts, const float* gamma, const float* beta,
                            const float* mean, const float* rstd, int rows) {
  const int tpr = bn_tpr(C, X, ldx, Y, ldy, mean, 4);
  if (tpr > 0 && aligned16(rstd) && (!gamma || aligned16(gamma)) && (!beta || aligned16(beta))) {
    k_bn_apply_v4<<<nblk, 256, 0, st>>>(X, ldx, Y, ldy, seg_row_ptr, C, tpr, parts, gamma, beta, mean, rstd, rows);
    return;
  }
  dim3 grid(nblk, ceil_div(C, 32));
  k_bn_apply<<<grid, 256, 0, st>>>(X, ldx, Y, ldy, seg_row_ptr, C, parts, gamma, beta, mean, rstd, rows);
}


// ---- chunk-resident BatchNorm backward (round 2): ONE thread-block cluster of 4 (or 8) CTAs per segment (chunk).  The two passes
// of the backward -- the sums (sum dy, sum dy*xhat), then dx -- both walk the chunk's rows of X and dY; as separate
// grid-wide kernels the second pass finds nothing of a 1.5 GB tensor in the 126 MB L2 (7.7 GB of DRAM traffic per call
// at 6 M rows x 64).  Here a cluster does both passes back to back on its own chunk: every CTA sums its share of the
// rows, the CTAs exchange their [2][64] fp64 partials through distributed shared memory (added in rank order:
// deterministic), and the second pass re-reads rows that were touched microseconds ago.  One CTA per SM (1 024 threads,
// the dynamic shared-memory request keeps a second one out), so at most 148 / 4 chunks (2 MB of X + dY each; dX leaves with streaming stores) are in
// flight and the re-reads hit L2: DRAM traffic = X + dY + dX once each.  C = 64 only (the path's channel width).
namespace cg = cooperative_groups;
constexpr int BNC_C = 64;
// (shared-memory request: 1 024-thread CTAs one per SM, 512-thread CTAs two per SM)
constexpr int bnc_smem(int threads) { return threads >= 1024 ? 120 * 1024 : 100 * 1024; }

template <int IN_ACT, int BNC_THREADS, int BNC_CL>
__global__ void __cluster_dims__(BNC_CL, 1, 1) __launch_bounds__(BNC_THREADS, BNC_THREADS >= 1024 ? 1 : 2)
k_bn_bwd_chunk(const float* __restrict__ X, int64_t ldx, const float* __restrict__ dY, int64_t lddy,
               float* __restrict__ dX, int64_t lddx, const int32_t* __restrict__ seg_row_ptr,
               const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
               double* __restrict__ seg_a, double* __restrict__ seg_b) {
  constexpr int BNC_RPB = BNC_THREADS / 16, BNC_WARPS = BNC_THREADS / 32;
  extern __shared__ double bnc_sm[];
  double* red = bnc_sm;                              // [2][BNC_WARPS][64]
  double* cta_sum = red + 2 * BNC_WARPS * BNC_C;     // [2][64]  (read by the other CTAs of the cluster)
  double* tot = cta_sum + 2 * BNC_C;                 // [2][64]
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int s = blockIdx.x / BNC_CL;
  const int lane16 = threadIdx.x & 15, ty = threadIdx.x >> 4, warp = threadIdx.x >> 5;
  const int c = 4 * lane16;
  const int ra = seg_row_ptr[s], n = seg_row_ptr[s + 1] - ra;
  const int r0 = ra + (int)(((int64_t)n * rank) / BNC_CL), r1 = ra + (int)(((int64_t)n * (rank + 1)) / BNC_CL);
  const float4 mu = ldg4(mean + (int64_t)s * BNC_C + c), rs = ldg4(rstd + (int64_t)s * BNC_C + c);
  // ---- pass 1: this CTA's rows (same per-element formulas as k_bn_reduce_part_v4<true>)
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
#pragma unroll 4
  for (int r = r0 + ty; r < r1; r += BNC_RPB) {
    const float4 x = ldg4(X + (int64_t)r * ldx + c);
    const float4 g = ldg4(dY + (int64_t)r * lddy + c);
    a0 += (double)g.x; a1 += (double)g.y; a2 += (double)g.z; a3 += (double)g.w;
    b0 += (double)g.x * (double)((x.x - mu.x) * rs.x); b1 += (double)g.y * (double)((x.y - mu.y) * rs.y);
    b2 += (double)g.z * (double)((x.z - mu.z) * rs.z); b3 += (double)g.w * (double)((x.w - mu.w) * rs.w);
  }
  // the two row groups of a warp (lanes l and l + 16 own the same channels), then the warps in order, then the CTAs
  a0 += __shfl_down_sync(0xffffffffu, a0, 16); a1 += __shfl_down_sync(0xffffffffu, a1, 16);
  a2 += __shfl_down_sync(0xffffffffu, a2, 16); a3 += __shfl_down_sync(0xffffffffu, a3, 16);
  b0 += __shfl_down_sync(0xffffffffu, b0, 16); b1 += __shfl_down_sync(0xffffffffu, b1, 16);
  b2 += __shfl_down_sync(0xffffffffu, b2, 16); b3 += __shfl_down_sync(0xffffffffu, b3, 16);
  if ((threadIdx.x & 31) < 16) {
    double* pa = red + (int64_t)warp * BNC_C + c;
    double* pb = red + (int64_t)(BNC_WARPS + warp) * BNC_C + c;
    pa[0] = a0; pa[1] = a1; pa[2] = a2; pa[3] = a3;
    pb[0] = b0; pb[1] = b1; pb[2] = b2; pb[3] = b3;
  }
  __syncthreads();
  if (threadIdx.x < 2 * BNC_C) {
    const int which = threadIdx.x >> 6, ch = threadIdx.x & 63;
    double t = 0.0;
#pragma unroll 8
    for (int w = 0; w < BNC_WARPS; ++w) t += red[(int64_t)(which * BNC_WARPS + w) * BNC_C + ch];
    cta_sum[threadIdx.x] = t;
  }
  cluster.sync();
  if (threadIdx.x < 2 * BNC_C) {
    double t = 0.0;
#pragma unroll
    for (unsigned q = 0; q < (unsigned)BNC_CL; ++q) t += *cluster.map_shared_rank(cta_sum + threadIdx.x, q);
    tot[threadIdx.x] = t;
    if (rank == 0) (threadIdx.x < BNC_C ? seg_a : seg_b)[(int64_t)s * BNC_C + (threadIdx.x & 63)] = t;
  }
  cluster.sync();                    // every remote read of cta_sum is done (a CTA may exit now); tot is visible
  if (n <= 0) return;
  // ---- pass 2: dx of the same rows (L2 hits)
  const double cnt = (double)n;
  const float4 ga = gamma ? ldg4(gamma + c) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float4 gm = make_float4((float)(tot[c] / cnt), (float)(tot[c + 1] / cnt), (float)(tot[c + 2] / cnt),
                                (float)(tot[c + 3] / cnt));
  const float4 dg = make_float4((float)(tot[BNC_C + c] / cnt), (float)(tot[BNC_C + c + 1] / cnt),
                                (float)(tot[BNC_C + c + 2] / cnt), (float)(tot[BNC_C + c + 3] / cnt));
  const float4 sc = make_float4(rs.x * ga.x, rs.y * ga.y, rs.z * ga.z, rs.w * ga.w);
#pragma unroll 4
  for (int r = r0 + ty; r < r1; r += BNC_RPB) {
    const float4 x = __ldcs(reinterpret_cast<const float4*>(X + (int64_t)r * ldx + c));      // last use: evict first
    const float4 g = __ldcs(reinterpret_cast<const float4*>(dY + (int64_t)r * lddy + c));
    __stcs(reinterpret_cast<float4*>(dX + (int64_t)r * lddx + c),      // streaming: must not push the chunks in flight out of L2
        make_float4(bn_bwd_one<IN_ACT>(x.x, g.x, mu.x, rs.x, gm.x, dg.x, sc.x),
                    bn_bwd_one<IN_ACT>(x.y, g.y, mu.y, rs.y, gm.y, dg.y, sc.y),
                    bn_bwd_one<IN_ACT>(x.z, g.z, mu.z, rs.z, gm.z, dg.z, sc.z),
                    bn_bwd_one<IN_ACT>(x.w, g.w, mu.w, rs.w, gm.w, dg.w, sc.w)));
  }
}

template <int IN_ACT, int THREADS, int CL>
static cudaError_t launch_bn_bwd_chunk_v(cudaStream_t st, int S, const float* X, int64_t ldx, const float* dY, int64_t lddy,
                                         float* dX, int64_t lddx, const int32_t* seg_row_ptr, const float* gamma,
                                         const float* mean, const float* rstd, double* seg_a, double* seg_b) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_bn_bwd_chunk<IN_ACT, THREADS, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         bnc_smem(THREADS));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  k_bn_bwd_chunk<IN_ACT, THREADS, CL><<<S * CL, THREADS, bnc_smem(THREADS), st>>>(X, ldx, dY, lddy, dX, lddx, seg_row_ptr,
                                                                                 gamma, mean, rstd, seg_a, seg_b);
  return cudaSuccess;
}

template <int IN_ACT>
static cudaError_t launch_bn_bwd_chunk(cudaStream_t st, int S, const float* X, int64_t ldx, const float* dY, int64_t lddy,
                                       float* dX, int64_t lddx, const int32_t* seg_row_ptr, const float* gamma,
                                       const float* mean, const float* rstd, double* seg_a, double* seg_b) {
  // 1 024-thread CTAs (one per SM), four per chunk: 37 chunks x (X + dY = 2 MB) in flight.  Measured on B200 at 6 M rows /
  // 1 562 chunks (profiles/r2_summary.md): 1.11 ms against 1.39 ms for the two grid-wide passes; 8 CTAs per chunk 1.32 ms
  // (15 clusters fit = 120 of 148 SMs), 512-thread CTAs two per SM x 8 per chunk 1.19 ms, x 4 per chunk 1.37 ms (74
  // chunks in flight no longer fit in L2).  BIGNN_BN_CL_VARIANT selects the others.
  const char* v = getenv("BIGNN_BN_CL_VARIANT");
  switch (v ? atoi(v) : 2) {
    case 0: return launch_bn_bwd_chunk_v<IN_ACT, 1024, 8>(st, S, X, ldx, dY, lddy, dX, lddx, seg_row_ptr, gamma, mean, rstd, seg_a, seg_b);
    case 1: return launch_bn_bwd_chunk_v<IN_ACT, 512, 8>(st, S, X, ldx, dY, lddy, dX, lddx, seg_row_ptr, gamma, mean, rstd, seg_a, seg_b);
    case 3: return launch_bn_bwd_chunk_v<IN_ACT, 512, 4>(st, S, X, ldx, dY, lddy, dX, lddx, seg_row_ptr, gamma, mean, rstd, seg_a, seg_b);
    default: return launch_bn_bwd_chunk_v<IN_ACT, 1024, 4>(st, S, X, ldx, dY, lddy, dX, lddx, seg_row_ptr, gamma, mean, rstd, seg_a, seg_b);
  }
}

static void launch_bn_bwd_apply(dim3 grid, cudaStream_t st, int in_act, const float* X, int64_t ldx, const float* dY,
                                int64_t lddy, float* dX, int64_t lddx, const int32_t* seg_row_ptr, int C, int parts,
                                const float* gamma, const float* mean, const float* rstd, const double* seg_a,
                                const double* seg_b, int rows, int64_t n_total) {
  const int tpr = bn_tpr(C, X, ldx, dY, lddy, dX, lddx);
  if (tpr > 0 && aligned16(mean) && aligned16(rstd) && (!gamma || aligned16(gamma))) {
#define BIGNN_BN_BWD4(A) k_bn_bwd_apply_v4<A><<<grid.x, 256, 0, st>>>(X, ldx, dY, lddy, dX, lddx, seg_row_ptr, C, tpr, \
                                                                      parts, gamma, mean, rstd, seg_a, seg_b, rows, n_total)
    switch (in_act) {
      case BIGNN_ACT_RELU: BIGNN_BN_BWD4(BIGNN_ACT_RELU); break;
      case BIGNN_ACT_SIGMOID: BIGNN_BN_BWD4(BIGNN_ACT_SIGMOID); break;
      case BIGNN_ACT_TANH: BIGNN_BN_BWD4(BIGNN_ACT_TANH); break;
      default: BIGNN_BN_BWD4(BIGNN_ACT_IDENTITY);
    }
#undef BIGNN_BN_BWD4
    return;
  }
#define BIGNN_BN_BWD(A) k_bn_bwd_apply<A><<<grid, 256, 0, st>>>(X, ldx, dY, lddy, dX, lddx, seg_row_ptr, C, parts, \
                                                                 gamma, mean, rstd, seg_a, seg_b, rows, n_total)
  switch (in_act) {
    case BIGNN_ACT_RELU: BIGNN_BN_BWD(BIGNN_ACT_RELU); break;
    case BIGNN_ACT_SIGMOID: BIGNN_BN_BWD(BIGNN_ACT_SIGMOID); break;
    case BIGNN_ACT_TANH: BIGNN_BN_BWD(BIGNN_ACT_TANH); break;
    default: BIGNN_BN_BWD(BIGNN_ACT_IDENTITY);
  }
#undef BIGNN_BN_BWD
}

}  // namespace bignn

using namespace bignn;

extern "C" int64_t bignn_bn_workspace_bytes(int32_t S, int32_t C, int32_t parts) {
  if (S <= 0 || C <= 0 || parts <= 0) return 0;
  return (int64_t)sizeof(double) * (2 * (int64_t)S * parts * C + 2 * (int64_t)S * C);
}

extern "C" int bignn_bn_seg_fwd(const float* X, int64_t ldx, float* Y, int64_t ldy, const int32_t* seg_row_ptr,
                                int32_t S, int32_t C, int32_t parts, const float* gamma, const float* beta,
                                float eps, float momentum, float* running_mean, float* running_var,
                                int64_t* num_batches_tracked, float* mean, float* rstd, double* seg_stats_out,
                                void* workspace, int64_t workspace_bytes, void* stream) {
  if (S < 0 || C < 0 || parts <= 0) return BIGNN_EINVAL;
  if (S == 0 || C == 0) return 0;
  if (!X || !Y || !seg_row_ptr || !mean || !rstd || ldx < C || ldy < C) return BIGNN_EINVAL;
  if ((int64_t)S * parts > 2147483647LL) return BIGNN_EINVAL;
  if (!workspace || workspace_bytes < bignn_bn_workspace_bytes(S, C, parts)) return BIGNN_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  double* ws_a = (double*)workspace;
  double* ws_b = ws_a + (int64_t)S * parts * C;
  double* mean_d = seg_stats_out ? seg_stats_out : ws_b + (int64_t)S * parts * C;
  double* varu_d = mean_d + (int64_t)S * C;
  dim3 grid(S * parts, ceil_div(C, 32));
  launch_bn_reduce(false, S * parts, st, X, ldx, nullptr, 0, seg_row_ptr, C, parts, nullptr, nullptr, ws_a, ws_b, 0);
  k_bn_finalize<<<dim3(S, ceil_div(C, 32)), 32 * BN_FIN_SL, 0, st>>>(ws_a, ws_b, seg_row_ptr, S, C, parts, eps, mean, rstd, mean_d, varu_d);
  BIGNN_LAUNCH_COUNT(2);
  if (running_mean && running_var) {
    k_bn_running<<<ceil_div(C, 64), 64, 0, st>>>(mean_d, varu_d, seg_row_ptr, S, C, (double)momentum, running_mean, running_var, num_batches_tracked);
    BIGNN_LAUNCH_COUNT(1);
  }
  launch_bn_apply(S * parts, st, X, ldx, Y, ldy, seg_row_ptr, C, parts, gamma, beta, mean, rstd, 0);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_bn_running_update(const double* seg_stats, const int32_t* seg_row_ptr, int32_t S, int32_t C,
                                       float momentum, float* running_mean, float* running_var,
                                       int64_t* num_batches_tracked, void* stream) {
  if (S < 0 || C < 0) return BIGNN_EINVAL;
  if (S == 0 || C == 0) return 0;
  if (!seg_stats || !seg_row_ptr || !running_mean || !running_var) return BIGNN_EINVAL;
  k_bn_running<<<ceil_div(C, 64), 64, 0, (cudaStream_t)stream>>>(seg_stats, seg_stats + (int64_t)S * C, seg_row_ptr, S,
                                                                  C, (double)momentum, running_mean, running_var,
                                                                  num_batches_tracked);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_bn_eval_fwd(const float* X, int64_t ldx, float* Y, int64_t ldy, int32_t rows, int32_t C,
                                 const float* gamma, const float* beta, float eps, const float* running_mean,
                                 const float* running_var, void* stream) {
  if (rows < 0 || C < 0) return BIGNN_EINVAL;
  if (rows == 0 || C == 0) return 0;
  if (!X || !Y || !running_mean || !running_var || ldx < C || ldy < C) return BIGNN_EINVAL;
  int gx = ceil_div(rows, BN_ROWS);
  const int cap = sm_count() * 8;
  if (gx > cap) gx = cap;
  dim3 grid(gx, ceil_div(C, 32));
  k_bn_eval<<<grid, 256, 0, (cudaStream_t)stream>>>(X, ldx, Y, ldy, rows, C, gamma, beta, eps, running_mean, running_var);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_bn_seg_bwd(const float* X, int64_t ldx, const float* dY, int64_t lddy, float* dX,
                                int64_t lddx, const int32_t* seg_row_ptr, int32_t S, int32_t C, int32_t parts,
                                const float* gamma, const float* mean, const float* rstd, float* dgamma,
                                float* dbeta, int32_t input_act, void* workspace, int64_t workspace_bytes,
                                void* stream) {
  if (S < 0 || C < 0 || parts <= 0 || input_act < 0 || input_act > BIGNN_ACT_TANH) return BIGNN_EINVAL;
  if (S == 0 || C == 0) return 0;
  if (!X || !dY || !dX || !seg_row_ptr || !mean || !rstd || ldx < C || lddy < C || lddx < C) return BIGNN_EINVAL;
  if ((int64_t)S * parts > 2147483647LL) return BIGNN_EINVAL;
  if (!workspace || workspace_bytes < bignn_bn_workspace_bytes(S, C, parts)) return BIGNN_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  double* ws_a = (double*)workspace;
  double* ws_b = ws_a + (int64_t)S * parts * C;
  double* seg_a = ws_b + (int64_t)S * parts * C;
  double* seg_b = seg_a + (int64_t)S * C;
  dim3 grid(S * parts, ceil_div(C, 32));
  // many chunks of 64 channels (the lower level of an all-drug pass): one cluster per chunk does both passes while the
  // chunk is in L2 (BIGNN_BN_CLUSTER=0: the two grid-wide passes below)
  const char* cl = getenv("BIGNN_BN_CLUSTER");
  if (C == BNC_C && S >= 24 && (int64_t)S * 8 <= 2147483647LL && !(cl && atoi(cl) == 0) &&
      bn_tpr(C, X, ldx, dY, lddy, dX, lddx) > 0 && aligned16(mean) && aligned16(rstd) && (!gamma || aligned16(gamma))) {
    cudaError_t e;
    switch (input_act) {
      case BIGNN_ACT_RELU: e = launch_bn_bwd_chunk<BIGNN_ACT_RELU>(st, S, X, ldx, dY, lddy, dX, lddx, seg_row_ptr, gamma, mean, rstd, seg_a, seg_b); break;
      case BIGNN_ACT_SIGMOID: e = launch_bn_bwd_chunk<BIGNN_ACT_SIGMOID>(st, S, X, ldx, dY, lddy, dX, lddx, seg_row_ptr, gamma, mean, rstd, seg_a, seg_b); break;
      case BIGNN_ACT_TANH: e = launch_bn_bwd_chunk<BIGNN_ACT_TANH>(st, S, X, ldx, dY, lddy, dX, lddx, seg_row_ptr, gamma, mean, rstd, seg_a, seg_b); break;
      default: e = launch_bn_bwd_chunk<BIGNN_ACT_IDENTITY>(st, S, X, ldx, dY, lddy, dX, lddx, seg_row_ptr, gamma, mean, rstd, seg_a, seg_b);
    }
    if (e != cudaSuccess) return (int)e;
    k_bn_bwd_params<<<ceil_div(C, 32), 1024, 0, st>>>(seg_a, seg_b, S, C, dgamma, dbeta);
    BIGNN_LAUNCH_COUNT(2);
    return last_launch_status();
  }
  launch_bn_reduce(true, S * parts, st, X, ldx, dY, lddy, seg_row_ptr, C, parts, mean, rstd, ws_a, ws_b, 0);
  k_bn_bwd_finalize<<<dim3(S, ceil_div(C, 32)), 32 * BN_FIN_SL, 0, st>>>(ws_a, ws_b, S, C, parts, seg_a, seg_b);
  k_bn_bwd_params<<<ceil_div(C, 32), 1024, 0, st>>>(seg_a, seg_b, S, C, dgamma, dbeta);
  launch_bn_bwd_apply(grid, st, input_act, X, ldx, dY, lddy, dX, lddx, seg_row_ptr, C, parts, gamma, mean, rstd, seg_a,
                      seg_b, 0, 0);
  BIGNN_LAUNCH_COUNT(4);
  return last_launch_status();
}

// ---------------------------------------------------------------------------------------------------
// Row-partitioned BatchNorm (multi-GPU upper level): local sums -> (caller: all-reduce) -> apply.
extern "C" int64_t bignn_bn_rows_workspace_bytes(int32_t C, int32_t parts) {
  if (C <= 0 || parts <= 0) return 0;
  return (int64_t)sizeof(double) * 2 * (int64_t)parts * C;
}

extern "C" int bignn_bn_rows_sums(const float* X, int64_t ldx, const float* dY, int64_t lddy, int32_t rows,
                                  int32_t C, int32_t parts, const float* mean, const float* rstd, double* sums,
                                  void* workspace, int64_t workspace_bytes, void* stream) {
  if (rows < 0 || C < 0 || parts <= 0) return BIGNN_EINVAL;
  if (C == 0) return 0;
  if (!sums || (rows > 0 && (!X || ldx < C))) return BIGNN_EINVAL;
  if (dY && (!mean || !rstd || lddy < C)) return BIGNN_EINVAL;
  if (!workspace || workspace_bytes < bignn_bn_rows_workspace_bytes(C, parts)) return BIGNN_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  double* ws_a = (double*)workspace;
  double* ws_b = ws_a + (int64_t)parts * C;
  dim3 grid(parts, ceil_div(C, 32));
  launch_bn_reduce(dY != nullptr, parts, st, X, ldx, dY, lddy, nullptr, C, parts, mean, rstd, ws_a, ws_b, rows);
  k_bn_rows_sum_parts<<<ceil_div(C, 256), 256, 0, st>>>(ws_a, ws_b, C, parts, sums);
  BIGNN_LAUNCH_COUNT(2);
  return last_launch_status();
}

extern "C" int bignn_bn_rows_fwd_apply(const float* X, int64_t ldx, float* Y, int64_t ldy, int32_t rows, int32_t C,
                                       int32_t parts, const double* sums, int64_t n_total, const float* gamma,
                                       const float* beta, float eps, float momentum, float* running_mean,
                                       float* running_var, int64_t* num_batches_tracked, float* mean, float* rstd,
                                       void* stream) {
  if (rows < 0 || C < 0 || parts <= 0 || n_total < rows) return BIGNN_EINVAL;
  if (C == 0) return 0;
  if (!sums || !mean || !rstd || (rows > 0 && (!X || !Y || ldx < C || ldy < C))) return BIGNN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  k_bn_rows_finalize<<<ceil_div(C, 256), 256, 0, st>>>(sums, n_total, C, eps, (double)momentum, mean, rstd,
                                                        running_mean, running_var, num_batches_tracked);
  BIGNN_LAUNCH_COUNT(1);
  if (rows > 0) {
    dim3 grid(parts, ceil_div(C, 32));
    launch_bn_apply(parts, st, X, ldx, Y, ldy, nullptr, C, parts, gamma, beta, mean, rstd, rows);
    BIGNN_LAUNCH_COUNT(1);
  }
  return last_launch_status();
}

extern "C" int bignn_bn_rows_bwd_apply(const float* X, int64_t ldx, const float* dY, int64_t lddy, float* dX,
                                       int64_t lddx, int32_t rows, int32_t C, int32_t parts, const float* gamma,
                                       const float* mean, const float* rstd, const double* sums, int64_t n_total,
                                       int32_t input_act, void* stream) {
  if (rows < 0 || C < 0 || parts <= 0 || n_total < rows || input_act < 0 || input_act > BIGNN_ACT_TANH) return BIGNN_EINVAL;
  if (rows == 0 || C == 0) return 0;
  if (!X || !dY || !dX || !mean || !rstd || !sums || ldx < C || lddy < C || lddx < C) return BIGNN_EINVAL;
  dim3 grid(parts, ceil_div(C, 32));
  launch_bn_bwd_apply(grid, (cudaStream_t)stream, input_act, X, ldx, dY, lddy, dX, lddx, nullptr, C, parts, gamma, mean,
                      rstd, sums, sums + C, rows, n_total);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}
