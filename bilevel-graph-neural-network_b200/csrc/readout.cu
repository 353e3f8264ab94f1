// Segment readout: atoms -> one embedding row per drug (see include/bignn_b200.h).
// Replaces torch-scatter scatter_mean / scatter_add at model/layers_aggregation.py:17-19,
// 34-41 and the per-row Python scatter into init_x at :70-74.
// One sub-warp (L lanes x 128-bit) per graph; rows are added in ascending order
// (the reference's sequential scatter_add order), 4 independent loads in flight.
// Algorithmic bytes: 4*D*A read + 4*D*G written + 4*(G+1).
#include "common.cuh"

namespace bignn {

template <int L>
__global__ void __launch_bounds__(256)
k_readout_fwd_v4(const float* __restrict__ X, int64_t ldx, const int32_t* __restrict__ seg_ptr, int G, int D4,
                 int style, const int32_t* __restrict__ dst_row, float* __restrict__ out, int64_t ldo,
                 const float* __restrict__ fold_mean, const float* __restrict__ fold_a,
                 const float* __restrict__ fold_beta, const int32_t* __restrict__ graph_chunk) {
  const int gpb = blockDim.x / L;
  const int sub = threadIdx.x / L, lane = threadIdx.x % L;
  for (int g = blockIdx.x * gpb + sub; g < G; g += gridDim.x * gpb) {
    const int r0 = __ldg(seg_ptr + g), r1 = __ldg(seg_ptr + g + 1);
    const int n = r1 - r0;
    const int64_t orow = dst_row ? dst_row[g] : g;
    const int64_t ch = (fold_a && graph_chunk) ? graph_chunk[g] : 0;
    for (int q = lane; q < D4; q += L) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      // folded BatchNorm, centred: pool(a (x - mean) + beta) = a * pool(x - mean) + beta * (1 | n)
      const float4 mu = fold_a ? ldg4(fold_mean + (ch * D4 + q) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      int r = r0;
      for (; r + 4 <= r1; r += 4) {
        float4 v0 = ldg4(X + (int64_t)(r + 0) * ldx + 4 * q);
        float4 v1 = ldg4(X + (int64_t)(r + 1) * ldx + 4 * q);
        float4 v2 = ldg4(X + (int64_t)(r + 2) * ldx + 4 * q);
        float4 v3 = ldg4(X + (int64_t)(r + 3) * ldx + 4 * q);
        if (fold_a) {
          v0.x -= mu.x; v0.y -= mu.y; v0.z -= mu.z; v0.w -= mu.w;  v1.x -= mu.x; v1.y -= mu.y; v1.z -= mu.z; v1.w -= mu.w;
          v2.x -= mu.x; v2.y -= mu.y; v2.z -= mu.z; v2.w -= mu.w;  v3.x -= mu.x; v3.y -= mu.y; v3.z -= mu.z; v3.w -= mu.w;
        }
        acc.x = (((acc.x + v0.x) + v1.x) + v2.x) + v3.x;
        acc.y = (((acc.y + v0.y) + v1.y) + v2.y) + v3.y;
        acc.z = (((acc.z + v0.z) + v1.z) + v2.z) + v3.z;
        acc.w = (((acc.w + v0.w) + v1.w) + v2.w) + v3.w;
      }
      for (; r < r1; ++r) {
        const float4 v = ldg4(X + (int64_t)r * ldx + 4 * q);
        acc.x += v.x - mu.x; acc.y += v.y - mu.y; acc.z += v.z - mu.z; acc.w += v.w - mu.w;
      }
      if (style == BIGNN_READOUT_MEAN) {
        const float cnt = (float)(n > 1 ? n : 1);      // count.clamp(min=1)
        acc.x = __fdiv_rn(acc.x, cnt); acc.y = __fdiv_rn(acc.y, cnt);
        acc.z = __fdiv_rn(acc.z, cnt); acc.w = __fdiv_rn(acc.w, cnt);
      }
      if (fold_a) {                                    // BatchNorm affine of the pooled rows, folded in
        const float4 a = ldg4(fold_a + (ch * D4 + q) * 4), b = ldg4(fold_beta + 4 * q);
        const float w = style == BIGNN_READOUT_MEAN ? (n > 0 ? 1.f : 0.f) : (float)n;
        acc.x = fmaf(a.x, acc.x, b.x * w); acc.y = fmaf(a.y, acc.y, b.y * w);
        acc.z = fmaf(a.z, acc.z, b.z * w); acc.w = fmaf(a.w, acc.w, b.w * w);
      }
      st4(out + orow * ldo + 4 * q, acc);
    }
  }
}

__global__ void __launch_bounds__(256)
k_readout_fwd_scalar(const float* __restrict__ X, int64_t ldx, const int32_t* __restrict__ seg_ptr, int G, int D,
                     int style, const int32_t* __restrict__ dst_row, float* __restrict__ out, int64_t ldo) {
  const int gpb = blockDim.x / 32;
  const int sub = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int g = blockIdx.x * gpb + sub; g < G; g += gridDim.x * gpb) {
    const int r0 = seg_ptr[g], r1 = seg_ptr[g + 1];
    const int n = r1 - r0;
    const int64_t orow = dst_row ? dst_row[g] : g;
    for (int q = lane; q < D; q += 32) {
      float acc = 0.f;
      for (int r = r0; r < r1; ++r) acc += __ldg(X + (int64_t)r * ldx + q);
      if (style == BIGNN_READOUT_MEAN) acc = __fdiv_rn(acc, (float)(n > 1 ? n : 1));
      out[orow * ldo + q] = acc;
    }
  }
}

// dX[r, :] (+)= dOut[dst_row[g], :] / n      for every row r of graph g -- 128-bit accesses, 16 lanes per graph, four
// rows in flight (the accumulate form reads before it writes: one row at a time is a chain of 30 dependent round trips)
__global__ void __launch_bounds__(256)
k_readout_bwd_v4(const float* __restrict__ dOut, int64_t ldo, const int32_t* __restrict__ dst_row,
                 const int32_t* __restrict__ seg_ptr, int G, int D4, int style, float* __restrict__ dX, int64_t lddx,
                 int accumulate) {
  const int gpb = blockDim.x / 16;
  const int sub = threadIdx.x / 16, lane = threadIdx.x % 16;
  for (int g = blockIdx.x * gpb + sub; g < G; g += gridDim.x * gpb) {
    const int r0 = __ldg(seg_ptr + g), r1 = __ldg(seg_ptr + g + 1);
    const int n = r1 - r0;
    const int64_t orow = dst_row ? dst_row[g] : g;
    const float cnt = (float)(n > 1 ? n : 1);
    for (int q = lane; q < D4; q += 16) {
      float4 v = ldg4(dOut + orow * ldo + 4 * q);
      if (style == BIGNN_READOUT_MEAN) {
        v.x = __fdiv_rn(v.x, cnt); v.y = __fdiv_rn(v.y, cnt); v.z = __fdiv_rn(v.z, cnt); v.w = __fdiv_rn(v.w, cnt);
      }
      int r = r0;
      if (accumulate) {
        for (; r + 4 <= r1; r += 4) {
          float4 o[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) o[u] = *reinterpret_cast<const float4*>(dX + (int64_t)(r + u) * lddx + 4 * q);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            o[u].x += v.x; o[u].y += v.y; o[u].z += v.z; o[u].w += v.w;
            st4(dX + (int64_t)(r + u) * lddx + 4 * q, o[u]);
          }
        }
        for (; r < r1; ++r) {
          float4 o = *reinterpret_cast<const float4*>(dX + (int64_t)r * lddx + 4 * q);
          o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
          st4(dX + (int64_t)r * lddx + 4 * q, o);
        }
      } else {
        for (; r < r1; ++r) st4(dX + (int64_t)r * lddx + 4 * q, v);
      }
    }
  }
}

// scalar form (D % 4 != 0 or unaligned)
__global__ void __launch_bounds__(256)
k_readout_bwd(const float* __restrict__ dOut, int64_t ldo, const int32_t* __restrict__ dst_row,
              const int32_t* __restrict__ seg_ptr, int G, int D, int style, float* __restrict__ dX, int64_t lddx,
              int accumulate) {
  const int gpb = blockDim.x / 32;
  const int sub = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int g = blockIdx.x * gpb + sub; g < G; g += gridDim.x * gpb) {
    const int r0 = seg_ptr[g], r1 = seg_ptr[g + 1];
    const int n = r1 - r0;
    const int64_t orow = dst_row ? dst_row[g] : g;
    const float cnt = (float)(n > 1 ? n : 1);
    for (int q = lane; q < D; q += 32) {
      float v = __ldg(dOut + orow * ldo + q);
      if (style == BIGNN_READOUT_MEAN) v = __fdiv_rn(v, cnt);
      for (int r = r0; r < r1; ++r) {
        float* p = dX + (int64_t)r * lddx + q;
        *p = accumulate ? (*p + v) : v;
      }
    }
  }
}

// ---- gated ("attention") readout: out[g] = sum over the graph's atoms of sigmoid(gate) * weight
// (model/layers_aggregation.py:90-94: `sigmoid(gate_func(x)) * weight_func(x)` -> scatter_add), the product is not
// materialised; rows in ascending order, the product rounded to fp32 before it is added (the unfused arithmetic)
template <int L>
__global__ void __launch_bounds__(256)
k_readout_gated_fwd(const float* __restrict__ Gt, int64_t ldg, const float* __restrict__ W, int64_t ldw,
                    const int32_t* __restrict__ seg_ptr, int G, int D4, float* __restrict__ out, int64_t ldo) {
  const int gpb = blockDim.x / L;
  const int sub = threadIdx.x / L, lane = threadIdx.x % L;
  for (int g = blockIdx.x * gpb + sub; g < G; g += gridDim.x * gpb) {
    const int r0 = __ldg(seg_ptr + g), r1 = __ldg(seg_ptr + g + 1);
    for (int q = lane; q < D4; q += L) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = r0; r < r1; ++r) {
        const float4 a = ldg4(Gt + (int64_t)r * ldg + 4 * q), w = ldg4(W + (int64_t)r * ldw + 4 * q);
        acc.x += __fmul_rn(1.0f / (1.0f + expf(-a.x)), w.x); acc.y += __fmul_rn(1.0f / (1.0f + expf(-a.y)), w.y);
        acc.z += __fmul_rn(1.0f / (1.0f + expf(-a.z)), w.z); acc.w += __fmul_rn(1.0f / (1.0f + expf(-a.w)), w.w);
      }
      st4(out + (int64_t)g * ldo + 4 * q, acc);
    }
  }
}

// dW[r] = dOut[g] * s, dGate[r] = dOut[g] * w * s (1 - s),  s = sigmoid(gate[r]),  for every atom r of graph g
template <int L>
__global__ void __launch_bounds__(256)
k_readout_gated_bwd(const float* __restrict__ Gt, int64_t ldg, const float* __restrict__ W, int64_t ldw,
                    const float* __restrict__ dOut, int64_t ldo, const int32_t* __restrict__ seg_ptr, int G, int D4,
                    float* __restrict__ dG, int64_t lddg, float* __restrict__ dW, int64_t lddw) {
  const int gpb = blockDim.x / L;
  const int sub = threadIdx.x / L, lane = threadIdx.x % L;
  for (int g = blockIdx.x * gpb + sub; g < G; g += gridDim.x * gpb) {
    const int r0 = __ldg(seg_ptr + g), r1 = __ldg(seg_ptr + g + 1);
    for (int q = lane; q < D4; q += L) {
      const float4 d = ldg4(dOut + (int64_t)g * ldo + 4 * q);
      for (int r = r0; r < r1; ++r) {
        const float4 a = ldg4(Gt + (int64_t)r * ldg + 4 * q), w = ldg4(W + (int64_t)r * ldw + 4 * q);
        float4 s, ow, og;
        s.x = 1.0f / (1.0f + expf(-a.x)); s.y = 1.0f / (1.0f + expf(-a.y));
        s.z = 1.0f / (1.0f + expf(-a.z)); s.w = 1.0f / (1.0f + expf(-a.w));
        ow.x = d.x * s.x; ow.y = d.y * s.y; ow.z = d.z * s.z; ow.w = d.w * s.w;
        og.x = d.x * w.x * ((1.0f - s.x) * s.x); og.y = d.y * w.y * ((1.0f - s.y) * s.y);
        og.z = d.z * w.z * ((1.0f - s.z) * s.z); og.w = d.w * w.w * ((1.0f - s.w) * s.w);
        st4(dW + (int64_t)r * lddw + 4 * q, ow);
        st4(dG + (int64_t)r * lddg + 4 * q, og);
      }
    }
  }
}

}  // namespace bignn

using namespace bignn;

extern "C" int bignn_readout_gated_fwd(const float* gate, int64_t ldg, const float* weight, int64_t ldw,
                                       const int32_t* seg_ptr, int32_t G, int32_t D, float* out, int64_t ldo,
                                       void* stream) {
  if (G < 0 || D < 0) return BIGNN_EINVAL;
  if (G == 0 || D == 0) return 0;
  if (!gate || !weight || !seg_ptr || !out || ldg < D || ldw < D || ldo < D) return BIGNN_EINVAL;
  if ((D & 3) || (ldg & 3) || (ldw & 3) || (ldo & 3) || !aligned16(gate) || !aligned16(weight) || !aligned16(out))
    return BIGNN_EALIGN;
  int grid = ceil_div(G, 16);
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  k_readout_gated_fwd<16><<<grid, 256, 0, (cudaStream_t)stream>>>(gate, ldg, weight, ldw, seg_ptr, G, D / 4, out, ldo);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_readout_gated_bwd(const float* gate, int64_t ldg, const float* weight, int64_t ldw,
                                       const float* dOut, int64_t ldo, const int32_t* seg_ptr, int32_t G, int32_t D,
                                       float* dGate, int64_t lddg, float* dWeight, int64_t lddw, void* stream) {
  if (G < 0 || D < 0) return BIGNN_EINVAL;
  if (G == 0 || D == 0) return 0;
  if (!gate || !weight || !dOut || !seg_ptr || !dGate || !dWeight || ldg < D || ldw < D || ldo < D || lddg < D ||
      lddw < D)
    return BIGNN_EINVAL;
  if ((D & 3) || (ldg & 3) || (ldw & 3) || (ldo & 3) || (lddg & 3) || (lddw & 3) || !aligned16(gate) ||
      !aligned16(weight) || !aligned16(dOut) || !aligned16(dGate) || !aligned16(dWeight))
    return BIGNN_EALIGN;
  int grid = ceil_div(G, 16);
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  k_readout_gated_bwd<16><<<grid, 256, 0, (cudaStream_t)stream>>>(gate, ldg, weight, ldw, dOut, ldo, seg_ptr, G, D / 4,
                                                                 dGate, lddg, dWeight, lddw);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_readout_fold_fwd(const float* X, int64_t ldx, const int32_t* seg_ptr, int32_t G, int32_t D,
                                      int32_t style, const int32_t* dst_row, const float* fold_mean,
                                      const float* fold_a, const float* fold_beta, const int32_t* graph_chunk,
                                      float* out, int64_t ldo, int32_t col_off, void* stream) {
  if (G < 0 || D < 0 || col_off < 0) return BIGNN_EINVAL;
  if (G == 0 || D == 0) return 0;
  if (!X || !seg_ptr || !out || ldx < D || ldo < col_off + D) return BIGNN_EINVAL;
  if (style != BIGNN_READOUT_SUM && style != BIGNN_READOUT_MEAN) return BIGNN_EINVAL;
  if ((fold_a != nullptr) != (fold_mean != nullptr) || (fold_a != nullptr) != (fold_beta != nullptr)) return BIGNN_EINVAL;
  if (fold_a && ((D % 4) || (ldx % 4) || (ldo % 4) || (col_off % 4) || !aligned16(X) || !aligned16(out) ||
                 !aligned16(fold_a) || !aligned16(fold_mean) || !aligned16(fold_beta)))
    return BIGNN_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  float* o = out + col_off;
  const int cap = sm_count() * 8;
  const bool vec = (D % 4 == 0) && (ldx % 4 == 0) && (ldo % 4 == 0) && aligned16(X) && aligned16(o);
  if (vec) {
    const int d4 = D / 4;
    if (d4 <= 16) {
      int grid = ceil_div(G, 16);
      if (grid > cap) grid = cap;
      k_readout_fwd_v4<16><<<grid, 256, 0, st>>>(X, ldx, seg_ptr, G, d4, style, dst_row, o, ldo, fold_mean, fold_a, fold_beta, graph_chunk);
    } else {
      int grid = ceil_div(G, 8);
      if (grid > cap) grid = cap;
      k_readout_fwd_v4<32><<<grid, 256, 0, st>>>(X, ldx, seg_ptr, G, d4, style, dst_row, o, ldo, fold_mean, fold_a, fold_beta, graph_chunk);
    }
  } else {
    int grid = ceil_div(G, 8);
    if (grid > cap) grid = cap;
    k_readout_fwd_scalar<<<grid, 256, 0, st>>>(X, ldx, seg_ptr, G, D, style, dst_row, o, ldo);
  }
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_readout_fwd(const float* X, int64_t ldx, const int32_t* seg_ptr, int32_t G, int32_t D,
                                 int32_t style, const int32_t* dst_row, float* out, int64_t ldo,
                                 int32_t col_off, void* stream) {
  return bignn_readout_fold_fwd(X, ldx, seg_ptr, G, D, style, dst_row, nullptr, nullptr, nullptr, nullptr, out, ldo,
                                col_off, stream);
}

extern "C" int bignn_readout_bwd(const float* dOut, int64_t ldo, int32_t col_off, const int32_t* dst_row,
                                 const int32_t* seg_ptr, int32_t G, int32_t D, int32_t style, float* dX,
                                 int64_t lddx, int32_t accumulate, void* stream) {
  if (G < 0 || D < 0 || col_off < 0) return BIGNN_EINVAL;
  if (G == 0 || D == 0) return 0;
  if (!dOut || !seg_ptr || !dX || lddx < D || ldo < col_off + D) return BIGNN_EINVAL;
  if (style != BIGNN_READOUT_SUM && style != BIGNN_READOUT_MEAN) return BIGNN_EINVAL;
  const int cap = sm_count() * 8;
  const float* d = dOut + col_off;
  if ((D % 4 == 0) && (ldo % 4 == 0) && (lddx % 4 == 0) && (col_off % 4 == 0) && aligned16(dOut) && aligned16(dX)) {
    int grid = ceil_div(G, 16);
    if (grid > cap) grid = cap;
    k_readout_bwd_v4<<<grid, 256, 0, (cudaStream_t)stream>>>(d, ldo, dst_row, seg_ptr, G, D / 4, style, dX, lddx, accumulate);
    BIGNN_LAUNCH_COUNT(1);
    return last_launch_status();
  }
  int grid = ceil_div(G, 8);
  if (grid > cap) grid = cap;
  k_readout_bwd<<<grid, 256, 0, (cudaStream_t)stream>>>(d, ldo, dst_row, seg_ptr, G, D, style, dX, lddx, accumulate);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}
