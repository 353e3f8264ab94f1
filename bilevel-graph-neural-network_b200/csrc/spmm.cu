// Deterministic row-parallel segment SpMM (see include/bignn_b200.h).
// Replaces PyG MessagePassing.propagate (index_select + scatter_add with float
// atomics) reached from model/layers.py:52-54; because every graph on the path is
// symmetric the same kernel is its own backward.
//
// Mapping: a sub-warp of L lanes owns one output row; each lane owns NV float4
// column slots, so a 64-float row is 16 lanes x one 128-bit load and a warp covers
// two rows.  Neighbour ids are read as warp-broadcast loads, neighbour rows are
// gathered 4 at a time (4*NV independent 128-bit loads in flight per lane) and
// added in ascending neighbour order with unfused mul/add, i.e. the exact
// summation order of the reference's sequential scatter_add over the sorted COO.
// No atomics, no shared memory; grid = a multiple of the SM count (grid-stride).
// Algorithmic bytes: 4*D*(rows read + rows written) + 4*nnz + 4*(rows+1).
#include "common.cuh"

namespace bignn {

__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void acc_add(float4& a, const float4& v) {
  a.x = __fadd_rn(a.x, v.x); a.y = __fadd_rn(a.y, v.y); a.z = __fadd_rn(a.z, v.z); a.w = __fadd_rn(a.w, v.w);
}
__device__ __forceinline__ void acc_add_scaled(float4& a, float w, const float4& v) {
  a.x = __fadd_rn(a.x, __fmul_rn(w, v.x)); a.y = __fadd_rn(a.y, __fmul_rn(w, v.y));
  a.z = __fadd_rn(a.z, __fmul_rn(w, v.z)); a.w = __fadd_rn(a.w, __fmul_rn(w, v.w));
}

// neighbours [k0,k1) of `row` accumulated in ascending order into acc[NV]
template <int L, int NV, int MODE>
__device__ __forceinline__ void accumulate_range(float4 (&acc)[NV], int k0, int k1, int row, int lane, int D4,
                                                 float di, const int32_t* __restrict__ col_idx,
                                                 const float* __restrict__ X, int64_t ldx,
                                                 const float* __restrict__ dinv) {
  constexpr int U = 4;  // neighbours gathered per step
  for (int k = k0; k < k1; k += U) {
    int c[U];
    float w[U];
    float4 val[U][NV];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      c[u] = (k + u < k1) ? __ldg(col_idx + k + u) : -1;
      if (MODE != BIGNN_SPMM_SUM && c[u] == row) c[u] = -1;  // remove_self_loops
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (MODE == BIGNN_SPMM_GCN) w[u] = c[u] >= 0 ? __fmul_rn(__ldg(dinv + c[u]), di) : 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int q = lane + v * L;
        val[u][v] = (c[u] >= 0 && q < D4) ? ldg4(X + (int64_t)c[u] * ldx + 4 * q) : f4zero();
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (c[u] >= 0) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          if (MODE == BIGNN_SPMM_GCN) acc_add_scaled(acc[v], w[u], val[u][v]);
          else acc_add(acc[v], val[u][v]);
        }
      }
    }
  }
}

// self term, bias, activation, store
template <int L, int NV, int MODE>
__device__ __forceinline__ void finish_row(float4 (&acc)[NV], int row, int grow, int lane, int D4, float self_coef, float di,
                                           const float* __restrict__ X, int64_t ldx, float* __restrict__ Y,
                                           int64_t ldy, const float* __restrict__ bias, int act) {
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int q = lane + v * L;
    if (q >= D4) continue;
    float4 a = acc[v];
    if (MODE == BIGNN_SPMM_GIN) {
      const float4 s = ldg4(X + (int64_t)grow * ldx + 4 * q);
      a.x = __fadd_rn(__fmul_rn(self_coef, s.x), a.x); a.y = __fadd_rn(__fmul_rn(self_coef, s.y), a.y);
      a.z = __fadd_rn(__fmul_rn(self_coef, s.z), a.z); a.w = __fadd_rn(__fmul_rn(self_coef, s.w), a.w);
    } else if (MODE == BIGNN_SPMM_GCN) {
      const float4 s = ldg4(X + (int64_t)grow * ldx + 4 * q);
      acc_add_scaled(a, __fmul_rn(di, di), s);       // self loop is the last COO entry
    }
    if (bias) {
      const float4 b = ldg4(bias + 4 * q);
      a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
    }
    if (act != BIGNN_ACT_IDENTITY) {
      a.x = apply_act(a.x, act); a.y = apply_act(a.y, act); a.z = apply_act(a.z, act); a.w = apply_act(a.w, act);
    }
    st4(Y + (int64_t)row * ldy + 4 * q, a);
  }
}

template <int L, int NV, int MODE>
__global__ void __launch_bounds__(256)
k_spmm_v4(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
          const float* __restrict__ X, int64_t ldx, float* __restrict__ Y, int64_t ldy,
          int n_rows, int D4, float self_coef, const float* __restrict__ dinv,
          const float* __restrict__ bias, int act, int roff) {
  const int rpb = blockDim.x / L;
  const int sub = threadIdx.x / L;
  const int lane = threadIdx.x % L;
  for (int row = blockIdx.x * rpb + sub; row < n_rows; row += gridDim.x * rpb) {
    const int grow = row + roff;          // row id in the column (node) index space
    const int k0 = __ldg(row_ptr + row), k1 = __ldg(row_ptr + row + 1);
    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = f4zero();
    float di = 0.f;
    if (MODE == BIGNN_SPMM_GCN) di = __ldg(dinv + grow);
    accumulate_range<L, NV, MODE>(acc, k0, k1, grow, lane, D4, di, col_idx, X, ldx, dinv);
    finish_row<L, NV, MODE>(acc, row, grow, lane, D4, self_coef, di, X, ldx, Y, ldy, bias, act);
  }
}

// ---- long-row variant: work items of at most `seg` neighbours --------------------------------
// item i covers neighbours [row_ptr[r] + c*seg, ...) of row r = item_row[i], c = i - item_ptr[r].
// Rows with one item are finished in place; rows with several items leave per-item partial sums in
// `partial` and are finished by k_spmm_multi_v4, which adds the partials in item order
// (deterministic, no atomics).  Keeps a 190k-neighbour hub row from serialising on one sub-warp.
template <int L, int NV, int MODE>
__global__ void __launch_bounds__(256)
k_spmm_items_v4(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                const int32_t* __restrict__ item_ptr, const int32_t* __restrict__ item_row, int n_items, int seg,
                const float* __restrict__ X, int64_t ldx, float* __restrict__ Y, int64_t ldy,
                int D4, float self_coef, const float* __restrict__ dinv, const float* __restrict__ bias, int act,
                float* __restrict__ partial, int roff) {
  const int ipb = blockDim.x / L;
  const int sub = threadIdx.x / L;
  const int lane = threadIdx.x % L;
  for (int item = blockIdx.x * ipb + sub; item < n_items; item += gridDim.x * ipb) {
    const int row = __ldg(item_row + item);
    const int i0 = __ldg(item_ptr + row), i1 = __ldg(item_ptr + row + 1);
    const int rk1 = __ldg(row_ptr + row + 1);
    const int k0 = __ldg(row_ptr + row) + (item - i0) * seg;
    const int k1 = min(k0 + seg, rk1);
    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = f4zero();
    const int grow = row + roff;
    float di = 0.f;
    if (MODE == BIGNN_SPMM_GCN) di = __ldg(dinv + grow);
    accumulate_range<L, NV, MODE>(acc, k0, k1, grow, lane, D4, di, col_idx, X, ldx, dinv);
    if (i1 - i0 == 1) {
      finish_row<L, NV, MODE>(acc, row, grow, lane, D4, self_coef, di, X, ldx, Y, ldy, bias, act);
    } else {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int q = lane + v * L;
        if (q < D4) st4(partial + ((int64_t)item * D4 + q) * 4, acc[v]);
      }
    }
  }
}

// partial sums of items [a, b) added to acc in item order, four 128-bit loads in flight per lane
template <int L, int NV>
__device__ __forceinline__ void sum_items(float4 (&acc)[NV], int a, int b, int lane, int D4,
                                          const float* __restrict__ partial) {
  constexpr int U = 4;
  int it = a;
  for (; it + U <= b; it += U) {
    float4 val[U][NV];
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int q = lane + v * L;
        val[u][v] = q < D4 ? ldg4(partial + ((int64_t)(it + u) * D4 + q) * 4) : f4zero();
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int v = 0; v < NV; ++v) acc_add(acc[v], val[u][v]);
    }
  }
  for (; it < b; ++it) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int q = lane + v * L;
      if (q < D4) acc_add(acc[v], ldg4(partial + ((int64_t)it * D4 + q) * 4));
    }
  }
}

template <int L, int NV, int MODE>
__global__ void __launch_bounds__(256)
k_spmm_multi_v4(const int32_t* __restrict__ item_ptr, const int32_t* __restrict__ multi_rows, int n_multi,
                const float* __restrict__ X, int64_t ldx, float* __restrict__ Y, int64_t ldy,
                int D4, float self_coef, const float* __restrict__ dinv, const float* __restrict__ bias, int act,
                const float* __restrict__ partial, int roff) {
  const int rpb = blockDim.x / L;
  const int sub = threadIdx.x / L;
  const int lane = threadIdx.x % L;
  for (int m = blockIdx.x * rpb + sub; m < n_multi; m += gridDim.x * rpb) {
    const int row = __ldg(multi_rows + m);
    const int i0 = __ldg(item_ptr + row), i1 = __ldg(item_ptr + row + 1);
    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = f4zero();
    sum_items<L, NV>(acc, i0, i1, lane, D4, partial);
    float di = 0.f;
    if (MODE == BIGNN_SPMM_GCN) di = __ldg(dinv + row + roff);
    finish_row<L, NV, MODE>(acc, row, row + roff, lane, D4, self_coef, di, X, ldx, Y, ldy, bias, act);
  }
}

// Hub rows (more than BIGNN_SPMM_BIG_ITEMS work items, e.g. a drug that interacts with most of a 200 k-drug graph):
// one CTA per row; every sub-warp sums a contiguous chunk of the row's partials in item order, the chunk sums are
// added in chunk order by sub-warp 0.  A fixed tree per (number of items, kernel configuration) -- deterministic --
// that shortens the dependent chain of the row from n_items to n_items / (256/L) + 256/L additions.
template <int L, int NV, int MODE>
__global__ void __launch_bounds__(256)
k_spmm_multi_big_v4(const int32_t* __restrict__ item_ptr, const int32_t* __restrict__ big_rows, int n_big,
                    const float* __restrict__ X, int64_t ldx, float* __restrict__ Y, int64_t ldy,
                    int D4, float self_coef, const float* __restrict__ dinv, const float* __restrict__ bias, int act,
                    const float* __restrict__ partial, int roff) {
  extern __shared__ float4 big_red[];                 // [256/L][D4]
  const int rpb = blockDim.x / L;
  const int sub = threadIdx.x / L;
  const int lane = threadIdx.x % L;
  for (int m = blockIdx.x; m < n_big; m += gridDim.x) {
    const int row = __ldg(big_rows + m);
    const int i0 = __ldg(item_ptr + row), i1 = __ldg(item_ptr + row + 1);
    const int chunk = (i1 - i0 + rpb - 1) / rpb;
    const int a = min(i0 + sub * chunk, i1), b = min(a + chunk, i1);
    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = f4zero();
    sum_items<L, NV>(acc, a, b, lane, D4, partial);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int q = lane + v * L;
      if (q < D4) big_red[sub * D4 + q] = acc[v];
    }
    __syncthreads();
    if (sub == 0) {
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[v] = f4zero();
      for (int s = 0; s < rpb; ++s) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int q = lane + v * L;
          if (q < D4) acc_add(acc[v], big_red[s * D4 + q]);
        }
      }
      float di = 0.f;
      if (MODE == BIGNN_SPMM_GCN) di = __ldg(dinv + row + roff);
      finish_row<L, NV, MODE>(acc, row, row + roff, lane, D4, self_coef, di, X, ldx, Y, ldy, bias, act);
    }
    __syncthreads();                                   // big_red is rewritten for the CTA's next row
  }
}

// any D / any alignment: a full warp per row, lane owns columns lane + 32*s
template <int NS, int MODE>
__global__ void __launch_bounds__(256)
k_spmm_scalar(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
              const float* __restrict__ X, int64_t ldx, float* __restrict__ Y, int64_t ldy,
              int n_rows, int D, float self_coef, const float* __restrict__ dinv,
              const float* __restrict__ bias, int act, int roff) {
  constexpr int U = 4;
  const int rpb = blockDim.x / 32;
  const int sub = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  for (int row = blockIdx.x * rpb + sub; row < n_rows; row += gridDim.x * rpb) {
    const int k0 = __ldg(row_ptr + row), k1 = __ldg(row_ptr + row + 1);
    float acc[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[s] = 0.f;
    const int grow = row + roff;
    float di = 0.f;
    if (MODE == BIGNN_SPMM_GCN) di = __ldg(dinv + grow);
    for (int k = k0; k < k1; k += U) {
      int c[U];
      float w[U];
      float val[U][NS];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        c[u] = (k + u < k1) ? __ldg(col_idx + k + u) : -1;
        if (MODE != BIGNN_SPMM_SUM && c[u] == grow) c[u] = -1;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (MODE == BIGNN_SPMM_GCN) w[u] = c[u] >= 0 ? __fmul_rn(__ldg(dinv + c[u]), di) : 0.f;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const int q = lane + 32 * s;
          val[u][s] = (c[u] >= 0 && q < D) ? __ldg(X + (int64_t)c[u] * ldx + q) : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (c[u] >= 0) {
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            if (MODE == BIGNN_SPMM_GCN) acc[s] = __fadd_rn(acc[s], __fmul_rn(w[u], val[u][s]));
            else acc[s] = __fadd_rn(acc[s], val[u][s]);
          }
        }
      }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const int q = lane + 32 * s;
      if (q >= D) continue;
      float a = acc[s];
      if (MODE == BIGNN_SPMM_GIN) {
        a = __fadd_rn(__fmul_rn(self_coef, __ldg(X + (int64_t)grow * ldx + q)), a);
      } else if (MODE == BIGNN_SPMM_GCN) {
        a = __fadd_rn(a, __fmul_rn(__fmul_rn(di, di), __ldg(X + (int64_t)grow * ldx + q)));
      }
      if (bias) a = __fadd_rn(a, __ldg(bias + q));
      a = apply_act(a, act);
      Y[(int64_t)row * ldy + q] = a;
    }
  }
}

__global__ void __launch_bounds__(256)
k_gcn_dinv(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx, int n_rows,
           float* __restrict__ dinv) {
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n_rows; row += gridDim.x * blockDim.x) {
    const int k0 = row_ptr[row], k1 = row_ptr[row + 1];
    int deg = 1;  // add_self_loops (weight 1)
    for (int k = k0; k < k1; ++k) deg += (col_idx[k] != row);
    // torch pow(-0.5) on CPU == 1/sqrt with both operations correctly rounded
    dinv[row] = __fdiv_rn(1.0f, __fsqrt_rn((float)deg));
  }
}

template <int MODE>
static int launch_mode(const int32_t* row_ptr, const int32_t* col_idx, const float* X, int64_t ldx,
                       float* Y, int64_t ldy, int n_rows, int D, float self_coef, const float* dinv,
                       const float* bias, int act, int roff, cudaStream_t st) {
  const bool vec = (D % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && aligned16(X) && aligned16(Y) &&
                   (bias == nullptr || aligned16(bias));
  const int cap = sm_count() * 8;
  if (vec) {
    // column blocks of at most 128 float4 (512 floats) per launch
    for (int c0 = 0; c0 < D; c0 += 512) {
      const int d4 = ((D - c0) < 512 ? (D - c0) : 512) / 4;
      const float* Xb = X + c0;
      float* Yb = Y + c0;
      const float* bb = bias ? bias + c0 : nullptr;
#define BIGNN_SPMM_LAUNCH(L, NV)                                                              \
  {                                                                                           \
    int rpb = 256 / L;                                                                        \
    int grid = ceil_div(n_rows, rpb);                                                         \
    if (grid > cap) grid = cap;                                                               \
    k_spmm_v4<L, NV, MODE><<<grid, 256, 0, st>>>(row_ptr, col_idx, Xb, ldx, Yb, ldy, n_rows, d4, \
                                                 self_coef, dinv, bb, act, roff);             \
  }
      if (d4 <= 8) BIGNN_SPMM_LAUNCH(8, 1)
      else if (d4 <= 16) BIGNN_SPMM_LAUNCH(16, 1)
      else if (d4 <= 32) BIGNN_SPMM_LAUNCH(32, 1)
      else if (d4 <= 64) BIGNN_SPMM_LAUNCH(32, 2)
      else if (d4 <= 96) BIGNN_SPMM_LAUNCH(32, 3)
      else BIGNN_SPMM_LAUNCH(32, 4)
#undef BIGNN_SPMM_LAUNCH
      BIGNN_LAUNCH_COUNT(1);
    }
  } else {
    for (int c0 = 0; c0 < D; c0 += 128) {
      const int d = (D - c0) < 128 ? (D - c0) : 128;
      int grid = ceil_div(n_rows, 8);
      if (grid > cap) grid = cap;
      const float* bb = bias ? bias + c0 : nullptr;
      if (d <= 32) k_spmm_scalar<1, MODE><<<grid, 256, 0, st>>>(row_ptr, col_idx, X + c0, ldx, Y + c0, ldy, n_rows, d, self_coef, dinv, bb, act, roff);
      else if (d <= 64) k_spmm_scalar<2, MODE><<<grid, 256, 0, st>>>(row_ptr, col_idx, X + c0, ldx, Y + c0, ldy, n_rows, d, self_coef, dinv, bb, act, roff);
      else k_spmm_scalar<4, MODE><<<grid, 256, 0, st>>>(row_ptr, col_idx, X + c0, ldx, Y + c0, ldy, n_rows, d, self_coef, dinv, bb, act, roff);
      BIGNN_LAUNCH_COUNT(1);
    }
  }
  return last_launch_status();
}


template <int MODE>
static int launch_planned(const int32_t* row_ptr, const int32_t* col_idx, const int32_t* item_ptr,
                          const int32_t* item_row, int n_items, int seg, const int32_t* multi_rows, int n_multi,
                          const float* X, int64_t ldx, float* Y, int64_t ldy, int D, float self_coef,
                          const float* dinv, const float* bias, int act, float* partial, int roff, int n_big,
                          cudaStream_t st) {
  const int cap = sm_count() * 8;
  const int d4 = D / 4;
#define BIGNN_PLANNED(L, NV)                                                                          \
  {                                                                                                   \
    const int per = 256 / L;                                                                          \
    int grid = ceil_div(n_items, per);                                                                \
    if (grid > cap) grid = cap;                                                                       \
    k_spmm_items_v4<L, NV, MODE><<<grid, 256, 0, st>>>(row_ptr, col_idx, item_ptr, item_row, n_items, seg, X, \
                                                       ldx, Y, ldy, d4, self_coef, dinv, bias, act, partial, roff); \
    BIGNN_LAUNCH_COUNT(1);                                                                            \
    const int n_small = n_multi - n_big;               /* the hub rows are the LAST n_big of multi_rows */ \
    if (n_small > 0) {                                                                                \
      int g2 = ceil_div(n_small, per);                                                                \
      if (g2 > cap) g2 = cap;                                                                         \
      k_spmm_multi_v4<L, NV, MODE><<<g2, 256, 0, st>>>(item_ptr, multi_rows, n_small, X, ldx, Y, ldy, d4,     \
                                                       self_coef, dinv, bias, act, partial, roff);    \
      BIGNN_LAUNCH_COUNT(1);                                                                          \
    }                                                                                                 \
    if (n_big > 0) {                                                                                  \
      const int g3 = n_big > cap ? cap : n_big;                                                       \
      k_spmm_multi_big_v4<L, NV, MODE><<<g3, 256, (256 / L) * d4 * sizeof(float4), st>>>(             \
          item_ptr, multi_rows + n_small, n_big, X, ldx, Y, ldy, d4, self_coef, dinv, bias, act, partial, roff); \
      BIGNN_LAUNCH_COUNT(1);                                                                          \
    }                                                                                                 \
  }
  if (d4 <= 8) BIGNN_PLANNED(8, 1)
  else if (d4 <= 16) BIGNN_PLANNED(16, 1)
  else if (d4 <= 32) BIGNN_PLANNED(32, 1)
  else if (d4 <= 64) BIGNN_PLANNED(32, 2)
  else if (d4 <= 96) BIGNN_PLANNED(32, 3)
  else BIGNN_PLANNED(32, 4)
#undef BIGNN_PLANNED
  return last_launch_status();
}

}  // namespace bignn

using namespace bignn;

extern "C" int bignn_gcn_dinv(const int32_t* row_ptr, const int32_t* col_idx, int32_t n_rows, float* dinv,
                              void* stream) {
  if (n_rows < 0 || !row_ptr || (n_rows > 0 && !dinv)) return BIGNN_EINVAL;
  if (n_rows == 0) return 0;
  int grid = ceil_div(n_rows, 256);
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  k_gcn_dinv<<<grid, 256, 0, (cudaStream_t)stream>>>(row_ptr, col_idx, n_rows, dinv);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_spmm_f32(const int32_t* row_ptr, const int32_t* col_idx, const float* X, int64_t ldx,
                              float* Y, int64_t ldy, int32_t n_rows, int32_t D, int32_t mode,
                              float self_coef, const float* dinv, const float* bias, int32_t act,
                              void* stream) {
  return bignn_spmm_rows_f32(row_ptr, col_idx, X, ldx, Y, ldy, n_rows, 0, D, mode, self_coef, dinv, bias, act, stream);
}

extern "C" int bignn_spmm_rows_f32(const int32_t* row_ptr, const int32_t* col_idx, const float* X, int64_t ldx,
                                   float* Y, int64_t ldy, int32_t n_rows, int32_t row_offset, int32_t D,
                                   int32_t mode, float self_coef, const float* dinv, const float* bias,
                                   int32_t act, void* stream) {
  const int roff = row_offset;
  if (n_rows < 0 || D < 0 || !row_ptr) return BIGNN_EINVAL;
  if (n_rows == 0 || D == 0) return 0;
  if (!X || !Y || ldx < D || ldy < D) return BIGNN_EINVAL;
  if (mode == BIGNN_SPMM_GCN && !dinv) return BIGNN_EINVAL;
  if (act < 0 || act > BIGNN_ACT_TANH) return BIGNN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  switch (mode) {
    case BIGNN_SPMM_SUM: return launch_mode<BIGNN_SPMM_SUM>(row_ptr, col_idx, X, ldx, Y, ldy, n_rows, D, self_coef, dinv, bias, act, roff, st);
    case BIGNN_SPMM_GIN: return launch_mode<BIGNN_SPMM_GIN>(row_ptr, col_idx, X, ldx, Y, ldy, n_rows, D, self_coef, dinv, bias, act, roff, st);
    case BIGNN_SPMM_GCN: return launch_mode<BIGNN_SPMM_GCN>(row_ptr, col_idx, X, ldx, Y, ldy, n_rows, D, self_coef, dinv, bias, act, roff, st);
    default: return BIGNN_EINVAL;
  }
}

extern "C" int64_t bignn_spmm_planned_workspace_bytes(int32_t n_items, int32_t D) {
  if (n_items <= 0 || D <= 0) return 0;
  return (int64_t)n_items * D * sizeof(float);
}

extern "C" int bignn_spmm_planned_f32(const int32_t* row_ptr, const int32_t* col_idx, const int32_t* item_ptr,
                                      const int32_t* item_row, int32_t n_items, int32_t seg,
                                      const int32_t* multi_rows, int32_t n_multi, const float* X, int64_t ldx,
                                      float* Y, int64_t ldy, int32_t n_rows, int32_t D, int32_t mode,
                                      float self_coef, const float* dinv, const float* bias, int32_t act,
                                      void* workspace, int64_t workspace_bytes, void* stream) {
  return bignn_spmm_planned_rows_f32(row_ptr, col_idx, item_ptr, item_row, n_items, seg, multi_rows, n_multi, 0, X, ldx,
                                     Y, ldy, n_rows, 0, D, mode, self_coef, dinv, bias, act, workspace,
                                     workspace_bytes, stream);
}

extern "C" int bignn_spmm_planned_rows_f32(const int32_t* row_ptr, const int32_t* col_idx, const int32_t* item_ptr,
                                           const int32_t* item_row, int32_t n_items, int32_t seg,
                                           const int32_t* multi_rows, int32_t n_multi, int32_t n_big,
                                           const float* X, int64_t ldx,
                                           float* Y, int64_t ldy, int32_t n_rows, int32_t row_offset, int32_t D,
                                           int32_t mode, float self_coef, const float* dinv, const float* bias,
                                           int32_t act, void* workspace, int64_t workspace_bytes, void* stream) {
  const int roff = row_offset;
  if (n_rows < 0 || D < 0 || n_items < 0 || n_multi < 0 || seg <= 0 || !row_ptr) return BIGNN_EINVAL;
  if (n_big < 0 || n_big > n_multi) return BIGNN_EINVAL;
  if (n_rows == 0 || D == 0) return 0;
  if (!X || !Y || !item_ptr || !item_row || ldx < D || ldy < D || D > 512) return BIGNN_EINVAL;
  if (n_multi > 0 && !multi_rows) return BIGNN_EINVAL;
  if (mode == BIGNN_SPMM_GCN && !dinv) return BIGNN_EINVAL;
  if (act < 0 || act > BIGNN_ACT_TANH) return BIGNN_EINVAL;
  if ((D % 4) || (ldx % 4) || (ldy % 4) || !aligned16(X) || !aligned16(Y) || (bias && !aligned16(bias)) ||
      (workspace && !aligned16(workspace)))
    return BIGNN_EALIGN;
  if (n_multi > 0 && (!workspace || workspace_bytes < bignn_spmm_planned_workspace_bytes(n_items, D)))
    return BIGNN_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = (float*)workspace;
  switch (mode) {
    case BIGNN_SPMM_SUM: return launch_planned<BIGNN_SPMM_SUM>(row_ptr, col_idx, item_ptr, item_row, n_items, seg, multi_rows, n_multi, X, ldx, Y, ldy, D, self_coef, dinv, bias, act, partial, roff, n_big, st);
    case BIGNN_SPMM_GIN: return launch_planned<BIGNN_SPMM_GIN>(row_ptr, col_idx, item_ptr, item_row, n_items, seg, multi_rows, n_multi, X, ldx, Y, ldy, D, self_coef, dinv, bias, act, partial, roff, n_big, st);
    case BIGNN_SPMM_GCN: return launch_planned<BIGNN_SPMM_GCN>(row_ptr, col_idx, item_ptr, item_row, n_items, seg, multi_rows, n_multi, X, ldx, Y, ldy, D, self_coef, dinv, bias, act, partial, roff, n_big, st);
    default: return BIGNN_EINVAL;
  }
}
