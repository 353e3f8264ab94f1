// Shared helpers for the sm_100a kernels of the Bi-GNN path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/bignn_b200.h"

namespace bignn {

extern long long g_launch_count;

inline int last_launch_status() {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
  return 0;
}

// number of SMs of the current device (cached)
int sm_count();

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case BIGNN_ACT_RELU: return v > 0.f ? v : 0.f;         // torch.relu: max(x,0) (NaN kept out of scope)
    case BIGNN_ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
    case BIGNN_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

// 128-bit read-only gather load (L2-resident operand rows are re-read by many rows)
__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
// the same without allocating the line in L1 (the fused layer kernel leaves L1 only 28 KB beside its operand slots)
__device__ __forceinline__ float4 ldg4_na(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
// streaming 128-bit store: outputs are written once and not re-read by this kernel
__device__ __forceinline__ void st4(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}

#define BIGNN_LAUNCH_COUNT(n) (::bignn::g_launch_count += (n))

}  // namespace bignn
