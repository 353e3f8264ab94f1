// Dense fp32 transforms of the path (see include/bignn_b200.h): C = act(op(A) op(B) + bias),
// deterministic split-K for the weight gradients, column sums, activation backward.
// Replaces nn.Linear / `x @ weight` and their autograd backward (model/layers.py:26-30,
// model/layers_util.py:28-33, PyG GCNConv/GATConv).
//
// This is the exact-fp32 (FMA, CUDA-core) path: 64x64x16 shared-memory tiles, 4x4
// register micro-tiles, 256 threads.  The feature widths on this path are 49/64/320,
// so one CTA column covers the whole N and A is streamed exactly once:
// algorithmic bytes = 4*(M*K + M*N) + weights.
#include "common.cuh"

namespace bignn {

constexpr int BM = 64, BN = 64, BK = 16;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
k_gemm_f32(int M, int N, int K, const float* __restrict__ A, int64_t lda,
           const float* __restrict__ B, int64_t ldb, float* __restrict__ C, int64_t ldc,
           const float* __restrict__ bias, int act, int k_chunk, float* __restrict__ ws) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * k_chunk;
  const int kend = min(K, kbeg + k_chunk);
  const int ty = t / 16, tx = t % 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int kk = kbeg; kk < kend; kk += BK) {
    // ---- A tile -> As[k][m]
    if (!TA) {
      const int m = t / 4, kq = (t % 4) * 4;
      const int gm = m0 + m;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gk = kk + kq + j;
        As[kq + j][m] = (gm < M && gk < kend) ? __ldg(A + (int64_t)gm * lda + gk) : 0.f;
      }
    } else {
      const int k = t / 16, mq = (t % 16) * 4;
      const int gk = kk + k;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gm = m0 + mq + j;
        As[k][mq + j] = (gm < M && gk < kend) ? __ldg(A + (int64_t)gk * lda + gm) : 0.f;
      }
    }
    // ---- B tile -> Bs[k][n]
    if (!TB) {
      const int k = t / 16, nq = (t % 16) * 4;
      const int gk = kk + k;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gn = n0 + nq + j;
        Bs[k][nq + j] = (gn < N && gk < kend) ? __ldg(B + (int64_t)gk * ldb + gn) : 0.f;
      }
    } else {
      const int n = t / 4, kq = (t % 4) * 4;
      const int gn = n0 + n;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gk = kk + kq + j;
        Bs[kq + j][n] = (gn < N && gk < kend) ? __ldg(B + (int64_t)gn * ldb + gk) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool direct = (gridDim.z == 1);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (direct) {
        if (bias) v += __ldg(bias + gn);
        C[(int64_t)gm * ldc + gn] = apply_act(v, act);
      } else {
        ws[((int64_t)blockIdx.z * M + gm) * N + gn] = v;
      }
    }
  }
}

// C[i] = act(sum_z ws[z][i] + bias): 32 consecutive outputs x 8 split groups per block -- every
// warp reads 128 contiguous bytes per split, the 8 groups are combined in fixed order (deterministic).
__global__ void __launch_bounds__(256)
k_splitk_reduce(const float* __restrict__ ws, int splits, int M, int N, float* __restrict__ C, int64_t ldc,
                const float* __restrict__ bias, int act) {
  __shared__ float red[8][33];
  const int64_t total = (int64_t)M * N;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const int64_t i = (int64_t)blockIdx.x * 32 + tx;
  float s = 0.f;
  if (i < total) {
#pragma unroll 4
    for (int z = ty; z < splits; z += 8) s += __ldg(ws + (int64_t)z * total + i);
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && i < total) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += red[g][tx];
    const int m = (int)(i / N), n = (int)(i % N);
    if (bias) t += __ldg(bias + n);
    C[(int64_t)m * ldc + n] = apply_act(t, act);
  }
}

static int gemm_splits(int M, int N, int K) {
  const int tiles = ceil_div(M, BM) * ceil_div(N, BN);
  const int sms = sm_count();
  if (tiles >= sms || K <= 512) return 1;
  int want = ceil_div(2 * sms, tiles);
  int maxs = ceil_div(K, 256);
  int s = want < maxs ? want : maxs;
  return s < 1 ? 1 : s;
}

// ---- column sums -----------------------------------------------------------
constexpr int CS_ROWS = 8;   // row lanes per block (x 32 columns)
static int colsum_parts(int rows) {
  int p = ceil_div(rows, 512);
  const int cap = 4 * sm_count();
  if (p > cap) p = cap;
  return p < 1 ? 1 : p;
}

__global__ void __launch_bounds__(256)
k_colsum_part(const float* __restrict__ X, int64_t ldx, int rows, int cols, int parts,
              double* __restrict__ ws) {
  __shared__ double red[CS_ROWS][33];
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const int c = blockIdx.x * 32 + tx;
  const int p = blockIdx.y;
  const int r0 = (int)(((int64_t)rows * p) / parts), r1 = (int)(((int64_t)rows * (p + 1)) / parts);
  double s = 0.0;
  if (c < cols)
    for (int r = r0 + ty; r < r1; r += CS_ROWS) s += (double)__ldg(X + (int64_t)r * ldx + c);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < cols) {
    double tot = 0.0;
#pragma unroll
    for (int i = 0; i < CS_ROWS; ++i) tot += red[i][tx];
    ws[(int64_t)p * cols + c] = tot;
  }
}

__global__ void __launch_bounds__(256)
k_colsum_final(const double* __restrict__ ws, int parts, int cols, float* __restrict__ out) {
  __shared__ double red[8][33];
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const int c = blockIdx.x * 32 + tx;
  double s = 0.0;
  if (c < cols) {
#pragma unroll 4
    for (int p = ty; p < parts; p += 8) s += ws[(int64_t)p * cols + c];
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < cols) {
    double t = 0.0;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += red[g][tx];
    out[c] = (float)t;
  }
}

__global__ void __launch_bounds__(256)
k_act_bwd(const float* __restrict__ Y, const float* __restrict__ dY, float* __restrict__ dX, int64_t n, int act) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float y = Y[i], g = dY[i];
    float d;
    switch (act) {
      case BIGNN_ACT_RELU: d = y > 0.f ? g : 0.f; break;
      case BIGNN_ACT_SIGMOID: d = g * ((1.0f - y) * y); break;
      case BIGNN_ACT_TANH: d = g * (1.0f - y * y); break;
      default: d = g;
    }
    dX[i] = d;
  }
}

}  // namespace bignn

using namespace bignn;

extern "C" int64_t bignn_gemm_workspace_bytes(int32_t M, int32_t N, int32_t K, int32_t ta) {
  (void)ta;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int s = gemm_splits(M, N, K);
  return s > 1 ? (int64_t)s * M * N * sizeof(float) : 0;
}

extern "C" int bignn_gemm_f32(int32_t ta, int32_t tb, int32_t M, int32_t N, int32_t K, const float* A,
                              int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                              const float* bias, int32_t act, void* workspace, int64_t workspace_bytes,
                              void* stream) {
  if (M < 0 || N < 0 || K < 0) return BIGNN_EINVAL;
  if (M == 0 || N == 0) return 0;
  if (!C || ldc < N || (K > 0 && (!A || !B))) return BIGNN_EINVAL;
  if (act < 0 || act > BIGNN_ACT_TANH) return BIGNN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  int splits = K > 0 ? gemm_splits(M, N, K) : 1;
  if (splits > 1 && (!workspace || workspace_bytes < (int64_t)splits * M * N * (int64_t)sizeof(float)))
    return BIGNN_EWORKSPACE;
  int k_chunk = K > 0 ? ceil_div(ceil_div(K, splits), BK) * BK : BK;
  splits = K > 0 ? ceil_div(K, k_chunk) : 1;
  dim3 grid(ceil_div(M, BM), ceil_div(N, BN), splits);
  if (grid.y > 65535u) return BIGNN_EINVAL;  // N beyond 4.1M columns is not on this path
  float* ws = (float*)workspace;
  if (!ta && !tb) k_gemm_f32<false, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, act, k_chunk, ws);
  else if (!ta && tb) k_gemm_f32<false, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, act, k_chunk, ws);
  else if (ta && !tb) k_gemm_f32<true, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, act, k_chunk, ws);
  else k_gemm_f32<true, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, act, k_chunk, ws);
  BIGNN_LAUNCH_COUNT(1);
  if (splits > 1) {
    k_splitk_reduce<<<ceil_div(M * N, 32), 256, 0, st>>>(ws, splits, M, N, C, ldc, bias, act);
    BIGNN_LAUNCH_COUNT(1);
  }
  return last_launch_status();
}

extern "C" int64_t bignn_colsum_workspace_bytes(int32_t rows, int32_t cols) {
  if (rows <= 0 || cols <= 0) return 0;
  return (int64_t)colsum_parts(rows) * cols * sizeof(double);
}

extern "C" int bignn_colsum_f32(const float* X, int64_t ldx, int32_t rows, int32_t cols, float* out,
                                void* workspace, int64_t workspace_bytes, void* stream) {
  if (rows < 0 || cols < 0) return BIGNN_EINVAL;
  if (cols == 0) return 0;
  if (!out) return BIGNN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) { cudaMemsetAsync(out, 0, sizeof(float) * cols, st); return last_launch_status(); }
  if (!X || ldx < cols) return BIGNN_EINVAL;
  const int parts = colsum_parts(rows);
  if (!workspace || workspace_bytes < (int64_t)parts * cols * (int64_t)sizeof(double)) return BIGNN_EWORKSPACE;
  dim3 grid(ceil_div(cols, 32), parts);
  k_colsum_part<<<grid, 256, 0, st>>>(X, ldx, rows, cols, parts, (double*)workspace);
  k_colsum_final<<<ceil_div(cols, 32), 256, 0, st>>>((const double*)workspace, parts, cols, out);
  BIGNN_LAUNCH_COUNT(2);
  return last_launch_status();
}

extern "C" int bignn_act_bwd_f32(const float* Y, const float* dY, float* dX, int64_t n, int32_t act,
                                 void* stream) {
  if (n < 0 || act < 0 || act > BIGNN_ACT_TANH) return BIGNN_EINVAL;
  if (n == 0) return 0;
  if (!Y || !dY || !dX) return BIGNN_EINVAL;
  int64_t g = ceil_div<int64_t>(n, 256);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (g > cap) g = cap;
  k_act_bwd<<<(int)g, 256, 0, (cudaStream_t)stream>>>(Y, dY, dX, n, act);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}
