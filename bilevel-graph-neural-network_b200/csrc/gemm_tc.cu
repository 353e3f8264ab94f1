// Dense node-feature transforms on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-accurate
// through a 3xTF32 split (see include/bignn_b200.h: bignn_gemm_tc_f32).
//
//   C[M,N] = act( A_eff[M,K] * B^T + bias ),   A_eff = A  or  A * act'(Y)   (fused activation backward)
//
// Replaces the nn.Linear / `x @ weight` transforms of model/layers.py:26-30 and PyG GCNConv/GATConv
// for the tall-skinny shapes of the path (M = atoms or drugs, N <= 256, any K).
//
// One CTA owns a 128-row tile.  K is walked in 32-float (128-byte) chunks, double buffered:
//   * all 256 threads load the A chunk (coalesced 128-bit loads) and the B chunk, split every fp32
//     value x into hi = x & 0xffffe000 (exactly a TF32) and lo = tf32(x - hi), and store both in
//     shared memory in the canonical K-major SWIZZLE_128B layout the UMMA descriptors expect
//     (8-row x 128-byte atoms, 16-byte chunk c of row r at position c ^ (r & 7));
//   * one elected thread issues tcgen05.mma.kind::tf32 (M=128, N, K=8): hi*hi + lo*hi + hi*lo
//     accumulated in fp32 in TMEM (the dropped lo*lo term is 2^-22 relative) and commits to an
//     mbarrier; the next chunk is staged while the tensor core runs;
//   * epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> bias + activation -> shared memory ->
//     coalesced 128-bit stores.
// No TMA descriptor is needed: operands are produced by the threads themselves (the split), which is
// also what lets later kernels fuse the neighbour aggregation in front of the transform.
#include "common.cuh"

namespace bignn {

constexpr int TC_BM = 128;          // rows per CTA tile (UMMA M)
constexpr int TC_KC = 32;           // floats per K chunk (128 bytes = one swizzle span)
constexpr int TC_THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  // start address (>>4) | LBO=1 (unused for swizzled K-major) | SBO = 1024 B (8 rows x 128 B) |
  // descriptor version 1 (sm_100) | layout type 2 = SWIZZLE_128B
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  // c_format F32 (1) @4, a_format TF32 (2) @7, b_format TF32 (2) @10, A and B K-major, N>>3 @17, M>>4 @24
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}

__device__ __forceinline__ float act_grad_from_output(float y, int act) {
  switch (act) {
    case BIGNN_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case BIGNN_ACT_SIGMOID: return (1.0f - y) * y;
    case BIGNN_ACT_TANH: return 1.0f - y * y;
    default: return 1.f;
  }
}

// byte offset of 16-byte chunk c (0..7) of row r inside a [rows x 128 B] K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

__device__ __forceinline__ void split_store(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, float4 v) {
  uint4 h, l;
  h.x = __float_as_uint(v.x) & 0xffffe000u; h.y = __float_as_uint(v.y) & 0xffffe000u;
  h.z = __float_as_uint(v.z) & 0xffffe000u; h.w = __float_as_uint(v.w) & 0xffffe000u;
  l.x = __float_as_uint(v.x - __uint_as_float(h.x)) & 0xffffe000u;
  l.y = __float_as_uint(v.y - __uint_as_float(h.y)) & 0xffffe000u;
  l.z = __float_as_uint(v.z - __uint_as_float(h.z)) & 0xffffe000u;
  l.w = __float_as_uint(v.w - __uint_as_float(h.w)) & 0xffffe000u;
  *reinterpret_cast<uint4*>(hi_base + off) = h;
  *reinterpret_cast<uint4*>(lo_base + off) = l;
}

// NPAD = N rounded up to a multiple of 32 (TMEM columns / epilogue granularity), <= 256
template <int NPAD>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_tc(int M, int N, int K, const float* __restrict__ A, int64_t lda, const float* __restrict__ Yact, int64_t ldy,
          int act_in, const float* __restrict__ B, int64_t ldb, int b_is_nk, float* __restrict__ C, int64_t ldc,
          const float* __restrict__ bias, int act) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: 2 stages x { A_hi, A_lo : 128 x 128 B ; B_hi, B_lo : NPAD x 128 B }
  constexpr int A_BYTES = TC_BM * 128;
  constexpr int B_BYTES = NPAD * 128;
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t mma_bar[2];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * TC_BM;
  constexpr int TMEM_COLS = NPAD <= 32 ? 32 : (NPAD <= 64 ? 64 : (NPAD <= 128 ? 128 : 256));

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    mbar_init(&mma_bar[0], 1);
    mbar_init(&mma_bar[1], 1);
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t idesc = umma_idesc_tf32(TC_BM, NPAD);

  const int n_chunks = (K + TC_KC - 1) / TC_KC;
  const bool a_vec = ((lda & 3) == 0) && aligned16(A) && (Yact == nullptr || (((ldy & 3) == 0) && aligned16(Yact)));
  uint32_t phase[2] = {0, 0};

  for (int ch = 0; ch < n_chunks; ++ch) {
    const int st = ch & 1;
    uint8_t* a_hi = smem + st * STAGE_BYTES;
    uint8_t* a_lo = a_hi + A_BYTES;
    uint8_t* b_hi = a_lo + A_BYTES;
    uint8_t* b_lo = b_hi + B_BYTES;
    if (ch >= 2) {             // the MMAs that read this stage two chunks ago must have finished
      mbar_wait(&mma_bar[st], phase[st]);
      phase[st] ^= 1;
    }
    const int kbase = ch * TC_KC;
    // ---- A chunk: 128 rows x 8 sixteen-byte chunks = 1024 float4, 4 per thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * TC_THREADS;
      const int r = idx >> 3, c = idx & 7;
      const int gm = m0 + r, gk = kbase + c * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gm < M && gk < K) {
        const float* ap = A + (int64_t)gm * lda + gk;
        if (a_vec && gk + 3 < K) {
          v = ldg4(ap);
          if (Yact) {
            const float4 y = ldg4(Yact + (int64_t)gm * ldy + gk);
            v.x *= act_grad_from_output(y.x, act_in); v.y *= act_grad_from_output(y.y, act_in);
            v.z *= act_grad_from_output(y.z, act_in); v.w *= act_grad_from_output(y.w, act_in);
          }
        } else {
          float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (gk + j < K) {
              t[j] = __ldg(ap + j);
              if (Yact) t[j] *= act_grad_from_output(__ldg(Yact + (int64_t)gm * ldy + gk + j), act_in);
            }
          v = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
      split_store(a_hi, a_lo, sw128_off(r, c), v);
    }
    // ---- B chunk: NPAD rows (output features) x 8 chunks; B stored [N,K] (b_is_nk) or [K,N]
    for (int idx = tid; idx < NPAD * 8; idx += TC_THREADS) {
      const int n = idx >> 3, c = idx & 7;
      const int gk = kbase + c * 4;
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      if (n < N) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (gk + j < K) t[j] = b_is_nk ? __ldg(B + (int64_t)n * ldb + gk + j) : __ldg(B + (int64_t)(gk + j) * ldb + n);
      }
      split_store(b_hi, b_lo, sw128_off(n, c), make_float4(t[0], t[1], t[2], t[3]));
    }
    // generic-proxy smem writes -> visible to the async (tensor core) proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t da_hi = umma_desc_k_sw128(smem_u32(a_hi)), da_lo = umma_desc_k_sw128(smem_u32(a_lo));
      const uint64_t db_hi = umma_desc_k_sw128(smem_u32(b_hi)), db_lo = umma_desc_k_sw128(smem_u32(b_lo));
#pragma unroll
      for (int k = 0; k < TC_KC / 8; ++k) {
        const uint64_t adv = (uint64_t)((k * 32) >> 4);       // 8 TF32 = 32 bytes along K inside the swizzle span
        umma_tf32(tmem_d, da_lo + adv, db_hi + adv, idesc, (ch | k) ? 1u : 0u);
        umma_tf32(tmem_d, da_hi + adv, db_lo + adv, idesc, 1u);
        umma_tf32(tmem_d, da_hi + adv, db_hi + adv, idesc, 1u);
      }
      umma_commit(&mma_bar[st]);
      if (ch == n_chunks - 1) umma_commit(&done_bar);
    }
  }
  // ---- epilogue: TMEM -> registers -> (bias, act) -> smem -> coalesced global stores
  mbar_wait(&done_bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float* stage_out = reinterpret_cast<float*>(smem);            // [128][NPAD + 4] floats; operand buffers are free now
  constexpr int LDS = NPAD + 4;
  {
    const int q = warp & 3;                                     // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;
    for (int cb = (warp >> 2) * 32; cb < NPAD; cb += 64) {
      uint32_t r[32];
      const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)cb;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int n = cb + j;
        float v = __uint_as_float(r[j]);
        if (bias && n < N) v += __ldg(bias + n);
        stage_out[row * LDS + n] = apply_act(v, act);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  {
    const bool c_vec = ((ldc & 3) == 0) && aligned16(C) && ((N & 3) == 0);
    if (c_vec) {
      const int n4 = N >> 2;
      for (int idx = tid; idx < TC_BM * n4; idx += TC_THREADS) {
        const int r = idx / n4, c = idx % n4;
        if (m0 + r < M)
          st4(C + (int64_t)(m0 + r) * ldc + 4 * c, *reinterpret_cast<const float4*>(&stage_out[r * LDS + 4 * c]));
      }
    } else {
      for (int idx = tid; idx < TC_BM * N; idx += TC_THREADS) {
        const int r = idx / N, c = idx % N;
        if (m0 + r < M) C[(int64_t)(m0 + r) * ldc + c] = stage_out[r * LDS + c];
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
  }
}

template <int NPAD>
static int launch_tc(int M, int N, int K, const float* A, int64_t lda, const float* Yact, int64_t ldy, int act_in,
                     const float* B, int64_t ldb, int b_is_nk, float* C, int64_t ldc, const float* bias, int act,
                     cudaStream_t st) {
  constexpr int stage = 2 * TC_BM * 128 + 2 * NPAD * 128;
  constexpr int out_stage = TC_BM * (NPAD + 4) * 4;
  constexpr int smem = (2 * stage > out_stage ? 2 * stage : out_stage) + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm_tc<NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  k_gemm_tc<NPAD><<<ceil_div(M, TC_BM), TC_THREADS, smem, st>>>(M, N, K, A, lda, Yact, ldy, act_in, B, ldb, b_is_nk, C,
                                                                ldc, bias, act);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

}  // namespace bignn

using namespace bignn;

extern "C" int bignn_gemm_tc_f32(int32_t M, int32_t N, int32_t K, const float* A, int64_t lda, const float* act_y,
                                 int64_t ldy, int32_t act_in, const float* B, int64_t ldb, int32_t b_is_nk, float* C,
                                 int64_t ldc, const float* bias, int32_t act, void* stream) {
  if (M < 0 || N < 0 || K < 0) return BIGNN_EINVAL;
  if (M == 0 || N == 0) return 0;
  if (K == 0 || !A || !B || !C || ldc < N || lda < K) return BIGNN_EINVAL;
  if (N > 128) return BIGNN_EINVAL;                          // the path's widths are <= 64; 128 keeps smem < 227 KB
  if (act < 0 || act > BIGNN_ACT_TANH || act_in < 0 || act_in > BIGNN_ACT_TANH) return BIGNN_EINVAL;
  if (act_y && ldy < K) return BIGNN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (N <= 32) return launch_tc<32>(M, N, K, A, lda, act_y, ldy, act_in, B, ldb, b_is_nk, C, ldc, bias, act, st);
  if (N <= 64) return launch_tc<64>(M, N, K, A, lda, act_y, ldy, act_in, B, ldb, b_is_nk, C, ldc, bias, act, st);
  if (N <= 96) return launch_tc<96>(M, N, K, A, lda, act_y, ldy, act_in, B, ldb, b_is_nk, C, ldc, bias, act, st);
  return launch_tc<128>(M, N, K, A, lda, act_y, ldy, act_in, B, ldb, b_is_nk, C, ldc, bias, act, st);
}
