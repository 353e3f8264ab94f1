// Pair decoder front end and loss heads (see include/bignn_b200.h).
// Replaces F.normalize + row gather + concat of model/layers_link_pred.py:44-61 and
// sigmoid + nn.BCELoss of model/layers_link_pred.py:65 / model/layers.py:71,88.
#include "common.cuh"

namespace bignn {

// one warp per (pair, side) entry e = 2p+side
__global__ void __launch_bounds__(256)
k_pair_gather_norm_fwd(const float* __restrict__ H, int64_t ldh, const int32_t* __restrict__ ids, int P, int D,
                       float* __restrict__ Z, int64_t ldz, float* __restrict__ inv_norm) {
  const int lane = threadIdx.x % 32;
  const int wpb = blockDim.x / 32;
  for (int e = blockIdx.x * wpb + threadIdx.x / 32; e < 2 * P; e += gridDim.x * wpb) {
    const int p = e >> 1, side = e & 1;
    const float* h = H + (int64_t)ids[e] * ldh;
    float ss = 0.f;
    for (int q = lane; q < D; q += 32) { const float v = __ldg(h + q); ss = fmaf(v, v, ss); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float nrm = fmaxf(__fsqrt_rn(ss), 1e-12f);       // F.normalize: x / max(|x|_2, eps)
    if (lane == 0) inv_norm[e] = nrm;                      // the clamped norm itself
    float* z = Z + (int64_t)p * ldz + (int64_t)side * D;
    for (int q = lane; q < D; q += 32) z[q] = __fdiv_rn(__ldg(h + q), nrm);
  }
}

// dRow_e = (dz - zhat * <dz, zhat>) / nrm   (zero gradient through the clamp branch when |h| < eps)
__global__ void __launch_bounds__(256)
k_pair_gather_norm_bwd(const float* __restrict__ H, int64_t ldh, const int32_t* __restrict__ ids, int P, int D,
                       const float* __restrict__ dZ, int64_t lddz, const float* __restrict__ inv_norm,
                       float* __restrict__ dRows, int64_t lddr) {
  const int lane = threadIdx.x % 32;
  const int wpb = blockDim.x / 32;
  for (int e = blockIdx.x * wpb + threadIdx.x / 32; e < 2 * P; e += gridDim.x * wpb) {
    const int p = e >> 1, side = e & 1;
    const float* h = H + (int64_t)ids[e] * ldh;
    const float* dz = dZ + (int64_t)p * lddz + (int64_t)side * D;
    const float nrm = inv_norm[e];
    float dot = 0.f;
    for (int q = lane; q < D; q += 32) dot = fmaf(__ldg(dz + q), __fdiv_rn(__ldg(h + q), nrm), dot);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const bool clamped = nrm <= 1e-12f;
    float* out = dRows + (int64_t)e * lddr;
    for (int q = lane; q < D; q += 32) {
      const float zh = __fdiv_rn(__ldg(h + q), nrm);
      const float g = __ldg(dz + q);
      out[q] = clamped ? __fdiv_rn(g, nrm) : __fdiv_rn(g - zh * dot, nrm);
    }
  }
}

// single block: loss = mean(-(y log p + (1-y) log(1-p))), logs clamped at -100 (torch BCELoss)
__global__ void __launch_bounds__(256)
k_bce_fwd(const float* __restrict__ pred, const float* __restrict__ y, int P, float* __restrict__ loss) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const float p = pred[i];
    const float lp = fmaxf(logf(p), -100.f);
    const float lq = fmaxf(log1pf(-p), -100.f);
    const float t = y[i];
    s += (double)((t - 1.0f) * lq - t * lp);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = (float)(red[0] / (double)(P > 0 ? P : 1));
}

// torch: grad_in = grad * (p - y) / clamp_min((1 - p) * p, 1e-12), grad = dloss / P
__global__ void __launch_bounds__(256)
k_bce_bwd(const float* __restrict__ pred, const float* __restrict__ y, int P, const float* __restrict__ dloss,
          float* __restrict__ dpred) {
  const float g = *dloss / (float)(P > 0 ? P : 1);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    const float p = pred[i];
    dpred[i] = g * (p - y[i]) / fmaxf((1.0f - p) * p, 1e-12f);
  }
}

// BCEWithLogits (torch): loss_i = (1-y) x + max(-x,0) + log(exp(-max(-x,0)) + exp(-x-max(-x,0)))
__global__ void __launch_bounds__(256)
k_bce_logits_fwd(const float* __restrict__ x, const float* __restrict__ y, int P, float* __restrict__ loss) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const float v = x[i], t = y[i];
    const float m = fmaxf(-v, 0.f);
    s += (double)((1.0f - t) * v + m + logf(expf(-m) + expf(-v - m)));
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
k_bce_logits_bwd(const float* __restrict__ x, const float* __restrict__ y, int P, const float* __restrict__ dloss,
                 float* __restrict__ dx) {
  const float g = *dloss / (float)(P > 0 ? P : 1);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x)
    dx[i] = (1.0f / (1.0f + expf(-x[i])) - y[i]) * g;
}

}  // namespace bignn

using namespace bignn;

extern "C" int bignn_pair_gather_norm_fwd(const float* H, int64_t ldh, const int32_t* ids, int32_t P, int32_t D,
                                          float* Z, int64_t ldz, float* inv_norm, void* stream) {
  if (P < 0 || D < 0) return BIGNN_EINVAL;
  if (P == 0 || D == 0) return 0;
  if (!H || !ids || !Z || !inv_norm || ldh < D || ldz < 2 * D) return BIGNN_EINVAL;
  int grid = ceil_div(2 * P, 8);
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  k_pair_gather_norm_fwd<<<grid, 256, 0, (cudaStream_t)stream>>>(H, ldh, ids, P, D, Z, ldz, inv_norm);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_pair_gather_norm_bwd(const float* H, int64_t ldh, const int32_t* ids, int32_t P, int32_t D,
                                          const float* dZ, int64_t lddz, const float* inv_norm, float* dRows,
                                          int64_t lddr, void* stream) {
  if (P < 0 || D < 0) return BIGNN_EINVAL;
  if (P == 0 || D == 0) return 0;
  if (!H || !ids || !dZ || !inv_norm || !dRows || ldh < D || lddz < 2 * D || lddr < D) return BIGNN_EINVAL;
  int grid = ceil_div(2 * P, 8);
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  k_pair_gather_norm_bwd<<<grid, 256, 0, (cudaStream_t)stream>>>(H, ldh, ids, P, D, dZ, lddz, inv_norm, dRows, lddr);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

static int loss_fwd(int kind, const float* pred, const float* y, int32_t P, float* loss, void* stream) {
  if (P < 0 || !loss) return BIGNN_EINVAL;
  if (P > 0 && (!pred || !y)) return BIGNN_EINVAL;
  if (kind == 0) k_bce_fwd<<<1, 256, 0, (cudaStream_t)stream>>>(pred, y, P, loss);
  else k_bce_logits_fwd<<<1, 256, 0, (cudaStream_t)stream>>>(pred, y, P, loss);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

static int loss_bwd(int kind, const float* pred, const float* y, int32_t P, const float* dloss, float* dpred,
                    void* stream) {
  if (P < 0) return BIGNN_EINVAL;
  if (P == 0) return 0;
  if (!pred || !y || !dloss || !dpred) return BIGNN_EINVAL;
  int grid = ceil_div(P, 256);
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  if (kind == 0) k_bce_bwd<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, y, P, dloss, dpred);
  else k_bce_logits_bwd<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, y, P, dloss, dpred);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_bce_fwd(const float* pred, const float* y, int32_t P, float* loss, void* stream) {
  return loss_fwd(0, pred, y, P, loss, stream);
}
extern "C" int bignn_bce_bwd(const float* pred, const float* y, int32_t P, const float* dloss, float* dpred,
                             void* stream) {
  return loss_bwd(0, pred, y, P, dloss, dpred, stream);
}
extern "C" int bignn_bce_logits_fwd(const float* x, const float* y, int32_t P, float* loss, void* stream) {
  return loss_fwd(1, x, y, P, loss, stream);
}
extern "C" int bignn_bce_logits_bwd(const float* x, const float* y, int32_t P, const float* dloss, float* dx,
                                    void* stream) {
  return loss_bwd(1, x, y, P, dloss, dx, stream);
}
