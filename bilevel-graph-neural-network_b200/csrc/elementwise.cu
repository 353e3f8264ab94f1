// Small row-wise / elementwise operators of the path (see include/bignn_b200.h):
// standalone activations (incl. PReLU), row L2 normalisation (NodeEmbedding normalize=True,
// model/layers.py:60-61), the gated product of the gmn_aggr readout
// (model/layers_aggregation.py:90-93), the dot-product scorer (model/layers_link_pred.py:66-67)
// and the cross-entropy head (model/layers.py:75,85-88).  All HBM-bound, one pass each.
#include "common.cuh"

namespace bignn {

__global__ void __launch_bounds__(256)
k_act_fwd(const float* __restrict__ X, float* __restrict__ Y, int64_t n, int act) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    Y[i] = apply_act(X[i], act);
}

__global__ void __launch_bounds__(256)
k_add(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ O, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    O[i] = A[i] + B[i];
}

// PReLU: y = x > 0 ? x : w[c] * x   (nw == 1: one shared slope, else one per column)
__global__ void __launch_bounds__(256)
k_prelu_fwd(const float* __restrict__ X, float* __restrict__ Y, int64_t n, int C, const float* __restrict__ w, int nw) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = X[i];
    const float a = w[nw == 1 ? 0 : (int)(i % C)];
    Y[i] = x > 0.f ? x : a * x;
  }
}

// dX = dY * (x > 0 ? 1 : w);  T = dY * (x > 0 ? 0 : x)   (column sums of T = d slope)
__global__ void __launch_bounds__(256)
k_prelu_bwd(const float* __restrict__ X, const float* __restrict__ dY, float* __restrict__ dX, float* __restrict__ T,
            int64_t n, int C, const float* __restrict__ w, int nw) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = X[i], g = dY[i];
    const float a = w[nw == 1 ? 0 : (int)(i % C)];
    dX[i] = x > 0.f ? g : a * g;
    T[i] = x > 0.f ? 0.f : g * x;
  }
}

// warp per row: y = x / max(|x|_2, 1e-12); nrm[row] = the clamped norm
__global__ void __launch_bounds__(256)
k_rownorm_fwd(const float* __restrict__ X, int64_t ldx, float* __restrict__ Y, int64_t ldy, int rows, int D,
              float* __restrict__ nrm) {
  const int lane = threadIdx.x % 32, wpb = blockDim.x / 32;
  for (int r = blockIdx.x * wpb + threadIdx.x / 32; r < rows; r += gridDim.x * wpb) {
    const float* x = X + (int64_t)r * ldx;
    float ss = 0.f;
    for (int q = lane; q < D; q += 32) { const float v = __ldg(x + q); ss = fmaf(v, v, ss); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float n = fmaxf(__fsqrt_rn(ss), 1e-12f);
    if (lane == 0) nrm[r] = n;
    float* y = Y + (int64_t)r * ldy;
    for (int q = lane; q < D; q += 32) y[q] = __fdiv_rn(__ldg(x + q), n);
  }
}

__global__ void __launch_bounds__(256)
k_rownorm_bwd(const float* __restrict__ Y, int64_t ldy, const float* __restrict__ dY, int64_t lddy,
              const float* __restrict__ nrm, float* __restrict__ dX, int64_t lddx, int rows, int D) {
  const int lane = threadIdx.x % 32, wpb = blockDim.x / 32;
  for (int r = blockIdx.x * wpb + threadIdx.x / 32; r < rows; r += gridDim.x * wpb) {
    const float* y = Y + (int64_t)r * ldy;
    const float* g = dY + (int64_t)r * lddy;
    const float n = nrm[r];
    float dot = 0.f;
    for (int q = lane; q < D; q += 32) dot = fmaf(__ldg(g + q), __ldg(y + q), dot);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const bool clamped = n <= 1e-12f;
    float* dx = dX + (int64_t)r * lddx;
    for (int q = lane; q < D; q += 32)
      dx[q] = clamped ? __fdiv_rn(__ldg(g + q), n) : __fdiv_rn(__ldg(g + q) - __ldg(y + q) * dot, n);
  }
}

// out = sigmoid(gate) * w
__global__ void __launch_bounds__(256)
k_gate_mul_fwd(const float* __restrict__ G, const float* __restrict__ W, float* __restrict__ O, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    O[i] = (1.0f / (1.0f + expf(-G[i]))) * W[i];
}
__global__ void __launch_bounds__(256)
k_gate_mul_bwd(const float* __restrict__ G, const float* __restrict__ W, const float* __restrict__ dO,
               float* __restrict__ dG, float* __restrict__ dW, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float s = 1.0f / (1.0f + expf(-G[i]));
    const float g = dO[i];
    dW[i] = g * s;
    dG[i] = g * W[i] * ((1.0f - s) * s);
  }
}

// warp per pair: out[p] = act(<Z[p,0:D], Z[p,D:2D]>)
__global__ void __launch_bounds__(256)
k_pair_dot_fwd(const float* __restrict__ Z, int64_t ldz, int P, int D, float* __restrict__ out, int act) {
  const int lane = threadIdx.x % 32, wpb = blockDim.x / 32;
  for (int p = blockIdx.x * wpb + threadIdx.x / 32; p < P; p += gridDim.x * wpb) {
    const float* z = Z + (int64_t)p * ldz;
    float dot = 0.f;
    for (int q = lane; q < D; q += 32) dot = fmaf(__ldg(z + q), __ldg(z + D + q), dot);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) out[p] = apply_act(dot, act);
  }
}
// dZ[p,0:D] = g z2, dZ[p,D:2D] = g z1 with g = dout * act'(out)
__global__ void __launch_bounds__(256)
k_pair_dot_bwd(const float* __restrict__ Z, int64_t ldz, int P, int D, const float* __restrict__ out,
               const float* __restrict__ dout, int act, float* __restrict__ dZ, int64_t lddz) {
  const int lane = threadIdx.x % 32, wpb = blockDim.x / 32;
  for (int p = blockIdx.x * wpb + threadIdx.x / 32; p < P; p += gridDim.x * wpb) {
    const float y = out[p];
    float g = dout[p];
    if (act == BIGNN_ACT_SIGMOID) g *= (1.0f - y) * y;
    else if (act == BIGNN_ACT_TANH) g *= 1.0f - y * y;
    else if (act == BIGNN_ACT_RELU) g = y > 0.f ? g : 0.f;
    const float* z = Z + (int64_t)p * ldz;
    float* d = dZ + (int64_t)p * lddz;
    for (int q = lane; q < D; q += 32) { d[q] = g * __ldg(z + D + q); d[D + q] = g * __ldg(z + q); }
  }
}

// cross entropy (mean) over P rows of K logits; single block, fp64 accumulation
__global__ void __launch_bounds__(256)
k_ce_fwd(const float* __restrict__ X, int64_t ldx, const int32_t* __restrict__ y, int P, int K, float* __restrict__ loss) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const float* x = X + (int64_t)i * ldx;
    float m = x[0];
    for (int k = 1; k < K; ++k) m = fmaxf(m, x[k]);
    float z = 0.f;
    for (int k = 0; k < K; ++k) z += expf(x[k] - m);
    s += (double)(logf(z) + m - x[y[i]]);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = (float)(red[0] / (double)(P > 0 ? P : 1));
}
__global__ void __launch_bounds__(256)
k_ce_bwd(const float* __restrict__ X, int64_t ldx, const int32_t* __restrict__ y, int P, int K,
         const float* __restrict__ dloss, float* __restrict__ dX, int64_t lddx) {
  const float g = *dloss / (float)(P > 0 ? P : 1);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    const float* x = X + (int64_t)i * ldx;
    float m = x[0];
    for (int k = 1; k < K; ++k) m = fmaxf(m, x[k]);
    float z = 0.f;
    for (int k = 0; k < K; ++k) z += expf(x[k] - m);
    for (int k = 0; k < K; ++k) dX[(int64_t)i * lddx + k] = g * (expf(x[k] - m) / z - (k == y[i] ? 1.f : 0.f));
  }
}

static inline int ew_grid(int64_t n) {
  int64_t g = ceil_div<int64_t>(n, 256);
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace bignn

using namespace bignn;

extern "C" int bignn_act_fwd_f32(const float* X, float* Y, int64_t n, int32_t act, void* stream) {
  if (n < 0 || act < 0 || act > BIGNN_ACT_TANH) return BIGNN_EINVAL;
  if (n == 0) return 0;
  if (!X || !Y) return BIGNN_EINVAL;
  k_act_fwd<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(X, Y, n, act);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_add_f32(const float* A, const float* B, float* O, int64_t n, void* stream) {
  if (n < 0) return BIGNN_EINVAL;
  if (n == 0) return 0;
  if (!A || !B || !O) return BIGNN_EINVAL;
  k_add<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(A, B, O, n);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_prelu_fwd_f32(const float* X, float* Y, int64_t rows, int32_t C, const float* w, int32_t nw,
                                   void* stream) {
  if (rows < 0 || C <= 0 || (nw != 1 && nw != C)) return BIGNN_EINVAL;
  if (rows == 0) return 0;
  if (!X || !Y || !w) return BIGNN_EINVAL;
  k_prelu_fwd<<<ew_grid(rows * C), 256, 0, (cudaStream_t)stream>>>(X, Y, rows * C, C, w, nw);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_prelu_bwd_f32(const float* X, const float* dY, float* dX, float* T, int64_t rows, int32_t C,
                                   const float* w, int32_t nw, void* stream) {
  if (rows < 0 || C <= 0 || (nw != 1 && nw != C)) return BIGNN_EINVAL;
  if (rows == 0) return 0;
  if (!X || !dY || !dX || !T || !w) return BIGNN_EINVAL;
  k_prelu_bwd<<<ew_grid(rows * C), 256, 0, (cudaStream_t)stream>>>(X, dY, dX, T, rows * C, C, w, nw);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_rownorm_fwd_f32(const float* X, int64_t ldx, float* Y, int64_t ldy, int32_t rows, int32_t D,
                                     float* nrm, void* stream) {
  if (rows < 0 || D < 0) return BIGNN_EINVAL;
  if (rows == 0 || D == 0) return 0;
  if (!X || !Y || !nrm || ldx < D || ldy < D) return BIGNN_EINVAL;
  int grid = ceil_div(rows, 8);
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  k_rownorm_fwd<<<grid, 256, 0, (cudaStream_t)stream>>>(X, ldx, Y, ldy, rows, D, nrm);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_rownorm_bwd_f32(const float* Y, int64_t ldy, const float* dY, int64_t lddy, const float* nrm,
                                     float* dX, int64_t lddx, int32_t rows, int32_t D, void* stream) {
  if (rows < 0 || D < 0) return BIGNN_EINVAL;
  if (rows == 0 || D == 0) return 0;
  if (!Y || !dY || !nrm || !dX || ldy < D || lddy < D || lddx < D) return BIGNN_EINVAL;
  int grid = ceil_div(rows, 8);
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  k_rownorm_bwd<<<grid, 256, 0, (cudaStream_t)stream>>>(Y, ldy, dY, lddy, nrm, dX, lddx, rows, D);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_gate_mul_fwd_f32(const float* G, const float* W, float* O, int64_t n, void* stream) {
  if (n < 0) return BIGNN_EINVAL;
  if (n == 0) return 0;
  if (!G || !W || !O) return BIGNN_EINVAL;
  k_gate_mul_fwd<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(G, W, O, n);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_gate_mul_bwd_f32(const float* G, const float* W, const float* dO, float* dG, float* dW,
                                      int64_t n, void* stream) {
  if (n < 0) return BIGNN_EINVAL;
  if (n == 0) return 0;
  if (!G || !W || !dO || !dG || !dW) return BIGNN_EINVAL;
  k_gate_mul_bwd<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(G, W, dO, dG, dW, n);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_pair_dot_fwd_f32(const float* Z, int64_t ldz, int32_t P, int32_t D, float* out, int32_t act,
                                      void* stream) {
  if (P < 0 || D < 0 || act < 0 || act > BIGNN_ACT_TANH) return BIGNN_EINVAL;
  if (P == 0) return 0;
  if (!Z || !out || ldz < 2 * D) return BIGNN_EINVAL;
  int grid = ceil_div(P, 8);
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  k_pair_dot_fwd<<<grid, 256, 0, (cudaStream_t)stream>>>(Z, ldz, P, D, out, act);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_pair_dot_bwd_f32(const float* Z, int64_t ldz, int32_t P, int32_t D, const float* out,
                                      const float* dout, int32_t act, float* dZ, int64_t lddz, void* stream) {
  if (P < 0 || D < 0 || act < 0 || act > BIGNN_ACT_TANH) return BIGNN_EINVAL;
  if (P == 0) return 0;
  if (!Z || !out || !dout || !dZ || ldz < 2 * D || lddz < 2 * D) return BIGNN_EINVAL;
  int grid = ceil_div(P, 8);
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  k_pair_dot_bwd<<<grid, 256, 0, (cudaStream_t)stream>>>(Z, ldz, P, D, out, dout, act, dZ, lddz);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_ce_fwd(const float* logits, int64_t ldx, const int32_t* labels, int32_t P, int32_t K,
                            float* loss, void* stream) {
  if (P < 0 || K <= 0 || !loss) return BIGNN_EINVAL;
  if (P > 0 && (!logits || !labels || ldx < K)) return BIGNN_EINVAL;
  k_ce_fwd<<<1, 256, 0, (cudaStream_t)stream>>>(logits, ldx, labels, P, K, loss);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_ce_bwd(const float* logits, int64_t ldx, const int32_t* labels, int32_t P, int32_t K,
                            const float* dloss, float* dlogits, int64_t lddx, void* stream) {
  if (P < 0 || K <= 0) return BIGNN_EINVAL;
  if (P == 0) return 0;
  if (!logits || !labels || !dloss || !dlogits || ldx < K || lddx < K) return BIGNN_EINVAL;
  k_ce_bwd<<<ew_grid(P), 256, 0, (cudaStream_t)stream>>>(logits, ldx, labels, P, K, dloss, dlogits, lddx);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}
