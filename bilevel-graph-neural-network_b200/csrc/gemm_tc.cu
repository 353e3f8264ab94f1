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
#include "tc_common.cuh"

namespace bignn {

// Persistent kernel: NPAD = N rounded up to a multiple of 32 (<= 128); KCH = ceil(K/32) chunks (<= 3).
// Requires K % 4 == 0, lda % 4 == 0 and a 16-byte aligned A (128-bit cp.async).
//
// Per 128-row tile:
//   cp.async (LDGSTS, no registers) streams the raw fp32 rows straight into the swizzled A_hi operand
//   buffer -- kind::tf32 reads only the upper 19 bits of each word, so the raw data IS the hi operand;
//   one shared->shared pass derives A_lo = tf32(x - hi(x));  one elected thread issues the MMAs;
//   the next tile's cp.async is issued as soon as the MMAs have drained the buffer, so it overlaps
//   the epilogue (TMEM -> registers -> swizzled staging in the A_lo area -> coalesced 128-bit stores).
// Two CTAs per SM (96 KB smem, <= 256 TMEM columns each) overlap each other's phases.
//
// TMEM: tcgen05.mma truncates (round-toward-zero) when it writes the fp32 accumulator, so a long
// accumulation chain drifts (measured -2.1e-8 * K relative).  The hi*hi products rotate over NA
// accumulators, the 2^-11-scaled correction products use their own, and the epilogue adds them in
// fp32 round-to-nearest.
// out * act'(y) from the activation OUTPUT y (the formulas of k_act_bwd): the fused activation backward of the
// backward-input GEMM  dT = (g W) * act'(t)
__device__ __forceinline__ float mask1(float o, float y, int mask_act) {
  switch (mask_act) {
    case BIGNN_ACT_RELU: return y > 0.f ? o : 0.f;
    case BIGNN_ACT_SIGMOID: return o * ((1.0f - y) * y);
    case BIGNN_ACT_TANH: return o * (1.0f - y * y);
    default: return o;
  }
}

template <int NPAD, int KCH>
__global__ void __launch_bounds__(TC_THREADS, 2)
k_gemm_tc(int M, int N, int K, const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
          int b_is_nk, float* __restrict__ C, int64_t ldc, const float* __restrict__ bias, int act,
          const float* __restrict__ mask_y, int64_t ldmy, int mask_act) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int A_BYTES = TC_BM * 128;            // one K chunk of the A tile
  constexpr int B_BYTES = NPAD * 128;
  constexpr int NBLK = NPAD / 32;                 // 32-column blocks of the output tile
  constexpr int LO_CH = KCH > NBLK ? KCH : NBLK;  // the A_lo area doubles as the output staging tile
  constexpr int NA = (NPAD <= 64) ? 2 : 1;        // main accumulators
  constexpr int ACC_COLS = (NA + 1) * NPAD;
  constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : (ACC_COLS <= 64 ? 64 : (ACC_COLS <= 128 ? 128 : 256));
  static_assert(ACC_COLS <= 256, "two CTAs per SM share the 512 TMEM columns");
  // (aligned through an OFFSET from the __shared__ array: a round trip through uintptr_t makes every access below a
  // generic LD / ST instead of LDS / STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_hi = smem;                            // [KCH][128 x 128 B]  raw fp32 rows
  uint8_t* a_lo = a_hi + KCH * A_BYTES;            // [LO_CH][128 x 128 B]
  uint8_t* b_hi = a_lo + LO_CH * A_BYTES;          // [KCH][NPAD x 128 B]
  uint8_t* b_lo = b_hi + KCH * B_BYTES;
  __shared__ uint64_t mma_bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = (M + TC_BM - 1) / TC_BM;
  const int kch_used = (K + TC_KC - 1) / TC_KC;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    mbar_init(&mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t a_hi_s = smem_u32(a_hi);
  // this thread's chunks of a tile: 16-byte chunk pc of rows pr, pr + 32, pr + 64, pr + 96 of every K chunk.  Everything
  // that does not depend on the tile is computed once (as in k_dw_tc_ring, where the per-chunk index arithmetic was a
  // third of the instructions a warp executed per tile): r & 7 is the same for the four rows, so the swizzled offset
  // advances by 4 KB per step
  static_assert(TC_THREADS == 256 && TC_BM == 128, "chunk mapping");
  const int pr = tid >> 3, pc = tid & 7;
  const uint32_t soff = sw128_off(pr, pc);
  const float* gp = A + (int64_t)pr * lda + pc * 4;
  auto prefetch_tile = [&](int tile) {
    const int m0 = tile * TC_BM;
    const float* g0 = gp + (int64_t)m0 * lda;
#pragma unroll
    for (int i = 0; i < KCH * 4; ++i) {
      if (i >= kch_used * 4) break;
      const int ch = i >> 2, rr = (i & 3) * 32;
      const bool ok = (m0 + pr + rr < M) && (ch * TC_KC + pc * 4 < K);
      cp_async16(a_hi_s + soff + i * 4096, ok ? g0 + (int64_t)rr * lda + ch * TC_KC : A, ok ? 16u : 0u);   // 0 -> zero fill
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int tile = blockIdx.x;
  if (tile < n_tiles) prefetch_tile(tile);
  // ---- weights: split once per CTA, resident in shared memory for every tile
#pragma unroll 1
  for (int idx = tid; idx < kch_used * NPAD * 8; idx += TC_THREADS) {
    const int ch = idx / (NPAD * 8), rem = idx % (NPAD * 8);
    const int n = rem >> 3, c = rem & 7;
    const int gk = ch * TC_KC + c * 4;
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    if (n < N) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (gk + j < K) t[j] = b_is_nk ? __ldg(B + (int64_t)n * ldb + gk + j) : __ldg(B + (int64_t)(gk + j) * ldb + n);
    }
    split_store(b_hi + ch * B_BYTES, b_lo + ch * B_BYTES, sw128_off(n, c), make_float4(t[0], t[1], t[2], t[3]));
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t idesc = umma_idesc_tf32(TC_BM, NPAD);
  const bool c_vec = ((ldc & 3) == 0) && aligned16(C) && ((N & 3) == 0) &&
                     (mask_y == nullptr || (((ldmy & 3) == 0) && aligned16(mask_y)));
  const int n_main = (K + 7) / 8 < NA ? (K + 7) / 8 : NA;

  uint32_t phase = 0;
  for (; tile < n_tiles; tile += gridDim.x) {
    const int m0 = tile * TC_BM;
    // ---- this thread's cp.async chunks have landed; derive the lo operand from them (same chunks)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int i = 0; i < KCH * 4; ++i) {
      if (i >= kch_used * 4) break;
      const uint32_t off = soff + i * 4096;
      const float4 x = *reinterpret_cast<const float4*>(a_hi + off);
      uint4 l;
      l.x = __float_as_uint(x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u)) & 0xffffe000u;
      l.y = __float_as_uint(x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u)) & 0xffffe000u;
      l.z = __float_as_uint(x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u)) & 0xffffe000u;
      l.w = __float_as_uint(x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u)) & 0xffffe000u;
      *reinterpret_cast<uint4*>(a_lo + off) = l;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async (tensor core) proxy
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      int ks = 0;
#pragma unroll
      for (int ch = 0; ch < KCH; ++ch) {
        if (ch * TC_KC >= K) break;
        const uint64_t da_hi = umma_desc_k_sw128(smem_u32(a_hi + ch * A_BYTES));
        const uint64_t da_lo = umma_desc_k_sw128(smem_u32(a_lo + ch * A_BYTES));
        const uint64_t db_hi = umma_desc_k_sw128(smem_u32(b_hi + ch * B_BYTES));
        const uint64_t db_lo = umma_desc_k_sw128(smem_u32(b_lo + ch * B_BYTES));
#pragma unroll
        for (int k = 0; k < TC_KC / 8; ++k, ++ks) {
          if (ch * TC_KC + k * 8 >= K) break;
          const uint64_t adv = (uint64_t)((k * 32) >> 4);     // 8 TF32 = 32 bytes along K inside the swizzle span
          umma_tf32(tmem_d + (uint32_t)((ks % NA) * NPAD), da_hi + adv, db_hi + adv, idesc, ks >= NA ? 1u : 0u);
          umma_tf32(tmem_d + (uint32_t)(NA * NPAD), da_lo + adv, db_hi + adv, idesc, ks > 0 ? 1u : 0u);
          umma_tf32(tmem_d + (uint32_t)(NA * NPAD), da_hi + adv, db_lo + adv, idesc, 1u);
        }
      }
      umma_commit(&mma_bar);
    }
    // ---- fused ReLU backward: fetch this thread's part of the mask tile now (its copy-out elements), as one bit
    // per element, so that the loads fly while the tensor core works instead of stalling the copy-out loop
    uint64_t mbits = 0;
    const bool mask_bits = mask_y != nullptr && mask_act == BIGNN_ACT_RELU && c_vec && N == NPAD &&
                           (NPAD == 32 || NPAD == 64 || NPAD == 128);
    if (mask_bits) {
      constexpr int N4 = NPAD / 4, RSTEP = TC_THREADS / N4, NR = TC_BM / RSTEP;
      const int c4 = tid % N4, rows_here = min(TC_BM, M - m0);
      float4 y[NR];
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int r = tid / N4 + i * RSTEP;
        y[i] = r < rows_here ? ldg4(mask_y + (int64_t)(m0 + r) * ldmy + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const uint32_t b = (y[i].x > 0.f ? 1u : 0u) | (y[i].y > 0.f ? 2u : 0u) | (y[i].z > 0.f ? 4u : 0u) |
                           (y[i].w > 0.f ? 8u : 0u);
        mbits |= (uint64_t)b << (4 * i);
      }
    }
    mbar_wait(&mma_bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- the MMAs have drained A_hi: stream the next tile in while the epilogue runs
    if (tile + gridDim.x < n_tiles) prefetch_tile(tile + gridDim.x);
    // ---- epilogue: TMEM -> registers (sum of accumulators, RN) -> bias/act -> swizzled staging (A_lo area)
    {
      const int q = warp & 3;                                   // TMEM lane quadrant this warp may read
      const int row = q * 32 + lane;
#pragma unroll 1
      for (int cb = (warp >> 2) * 32; cb < NPAD; cb += 64) {
        uint32_t r[32];
        float v[32];
        const uint32_t tbase = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)cb;
        tmem_ld32(tbase + (uint32_t)(NA * NPAD), r);            // corrections first (small)
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        tmem_ld32(tbase, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __fadd_rn(__uint_as_float(r[j]), v[j]);
        if (NA == 2 && n_main == 2) {
          tmem_ld32(tbase + (uint32_t)NPAD, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __fadd_rn(__uint_as_float(r[j]), v[j]);
        }
        uint8_t* blk = a_lo + (cb >> 5) * A_BYTES;
#pragma unroll
        for (int c = 0; c < 8; ++c)      // raw sums; bias + activation are applied in the (rolled) copy-out loop
          *reinterpret_cast<float4*>(blk + sw128_off(row, c)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (c_vec && N == NPAD && (NPAD == 32 || NPAD == 64 || NPAD == 128)) {
      // fast path (the path's widths): 256 % (N/4) == 0, so a thread keeps one 16-byte column chunk for
      // all its rows -- no integer division, bias loaded once, activation branch hoisted out of the loop
      constexpr int N4 = NPAD / 4, RSTEP = TC_THREADS / N4;
      const int c4 = tid % N4;
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias) b4 = ldg4(bias + 4 * c4);
      const uint8_t* src = a_lo + (c4 >> 3) * A_BYTES;
      const int rows_here = min(TC_BM, M - m0);
      if (act == BIGNN_ACT_RELU) {
#pragma unroll 4
        for (int r = tid / N4; r < rows_here; r += RSTEP) {
          float4 o = *reinterpret_cast<const float4*>(src + sw128_off(r, c4 & 7));
          o.x = fmaxf(o.x + b4.x, 0.f); o.y = fmaxf(o.y + b4.y, 0.f); o.z = fmaxf(o.z + b4.z, 0.f); o.w = fmaxf(o.w + b4.w, 0.f);
          st4(C + (int64_t)(m0 + r) * ldc + 4 * c4, o);
        }
      } else if (act == BIGNN_ACT_IDENTITY && mask_bits) {
        constexpr int NR = TC_BM / RSTEP;
#pragma unroll
        for (int i = 0; i < NR; ++i) {                           // backward-input GEMM with the ReLU mask of t
          const int r = tid / N4 + i * RSTEP;
          if (r < rows_here) {
            float4 o = *reinterpret_cast<const float4*>(src + sw128_off(r, c4 & 7));
            const uint32_t b = (uint32_t)(mbits >> (4 * i));
            o.x = (b & 1u) ? o.x + b4.x : 0.f; o.y = (b & 2u) ? o.y + b4.y : 0.f;
            o.z = (b & 4u) ? o.z + b4.z : 0.f; o.w = (b & 8u) ? o.w + b4.w : 0.f;
            st4(C + (int64_t)(m0 + r) * ldc + 4 * c4, o);
          }
        }
      } else if (act == BIGNN_ACT_IDENTITY && mask_y == nullptr) {
#pragma unroll 4
        for (int r = tid / N4; r < rows_here; r += RSTEP) {
          float4 o = *reinterpret_cast<const float4*>(src + sw128_off(r, c4 & 7));
          o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
          st4(C + (int64_t)(m0 + r) * ldc + 4 * c4, o);
        }
      } else {
#pragma unroll 1
        for (int r = tid / N4; r < rows_here; r += RSTEP) {
          float4 o = *reinterpret_cast<const float4*>(src + sw128_off(r, c4 & 7));
          o.x = apply_act(o.x + b4.x, act); o.y = apply_act(o.y + b4.y, act);
          o.z = apply_act(o.z + b4.z, act); o.w = apply_act(o.w + b4.w, act);
          if (mask_y) {
            const float4 y = ldg4(mask_y + (int64_t)(m0 + r) * ldmy + 4 * c4);
            o.x = mask1(o.x, y.x, mask_act); o.y = mask1(o.y, y.y, mask_act);
            o.z = mask1(o.z, y.z, mask_act); o.w = mask1(o.w, y.w, mask_act);
          }
          st4(C + (int64_t)(m0 + r) * ldc + 4 * c4, o);
        }
      }
    } else if (c_vec) {
      const int n4 = N >> 2;
#pragma unroll 1
      for (int idx = tid; idx < TC_BM * n4; idx += TC_THREADS) {
        const int r = idx / n4, c4 = idx % n4;
        if (m0 + r < M) {
          float4 o = *reinterpret_cast<const float4*>(a_lo + (c4 >> 3) * A_BYTES + sw128_off(r, c4 & 7));
          if (bias) {
            o.x += __ldg(bias + 4 * c4); o.y += __ldg(bias + 4 * c4 + 1);
            o.z += __ldg(bias + 4 * c4 + 2); o.w += __ldg(bias + 4 * c4 + 3);
          }
          if (act != BIGNN_ACT_IDENTITY) {
            o.x = apply_act(o.x, act); o.y = apply_act(o.y, act); o.z = apply_act(o.z, act); o.w = apply_act(o.w, act);
          }
          if (mask_y) {
            const float4 y = ldg4(mask_y + (int64_t)(m0 + r) * ldmy + 4 * c4);
            o.x = mask1(o.x, y.x, mask_act); o.y = mask1(o.y, y.y, mask_act);
            o.z = mask1(o.z, y.z, mask_act); o.w = mask1(o.w, y.w, mask_act);
          }
          st4(C + (int64_t)(m0 + r) * ldc + 4 * c4, o);
        }
      }
    } else {
#pragma unroll 1
      for (int idx = tid; idx < TC_BM * N; idx += TC_THREADS) {
        const int r = idx / N, n = idx % N;
        if (m0 + r < M) {
          float o = *reinterpret_cast<const float*>(a_lo + (n >> 5) * A_BYTES + sw128_off(r, (n & 31) >> 2) + (n & 3) * 4);
          if (bias) o += __ldg(bias + n);
          o = apply_act(o, act);
          if (mask_y) o = mask1(o, __ldg(mask_y + (int64_t)(m0 + r) * ldmy + n), mask_act);
          C[(int64_t)(m0 + r) * ldc + n] = o;
        }
      }
    }
    __syncthreads();                                            // staging (A_lo) is rewritten by the next tile's split
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
  }
}

template <int NPAD, int KCH>
static int launch_tc(int M, int N, int K, const float* A, int64_t lda, const float* B, int64_t ldb, int b_is_nk,
                     float* C, int64_t ldc, const float* bias, int act, const float* mask_y, int64_t ldmy,
                     int mask_act, cudaStream_t st) {
  constexpr int lo_ch = KCH > NPAD / 32 ? KCH : NPAD / 32;
  constexpr int smem = (KCH + lo_ch) * TC_BM * 128 + 2 * KCH * NPAD * 128 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm_tc<NPAD, KCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int n_tiles = ceil_div(M, TC_BM);
  int grid = 2 * sm_count();
  if (grid > n_tiles) grid = n_tiles;
  k_gemm_tc<NPAD, KCH><<<grid, TC_THREADS, smem, st>>>(M, N, K, A, lda, B, ldb, b_is_nk, C, ldc, bias, act, mask_y,
                                                       ldmy, mask_act);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

template <int NPAD>
static int launch_tc_k(int M, int N, int K, const float* A, int64_t lda, const float* B, int64_t ldb, int b_is_nk,
                       float* C, int64_t ldc, const float* bias, int act, const float* mask_y, int64_t ldmy,
                       int mask_act, cudaStream_t st) {
  if (K <= 32) return launch_tc<NPAD, 1>(M, N, K, A, lda, B, ldb, b_is_nk, C, ldc, bias, act, mask_y, ldmy, mask_act, st);
  if (K <= 64) return launch_tc<NPAD, 2>(M, N, K, A, lda, B, ldb, b_is_nk, C, ldc, bias, act, mask_y, ldmy, mask_act, st);
  if (NPAD <= 64 && K <= 96)
    return launch_tc<NPAD, 3>(M, N, K, A, lda, B, ldb, b_is_nk, C, ldc, bias, act, mask_y, ldmy, mask_act, st);
  return BIGNN_EINVAL;
}

}  // namespace bignn

using namespace bignn;

extern "C" int bignn_gemm_tc_supported(int32_t M, int32_t N, int32_t K) {
  if (M <= 0 || N <= 0 || K <= 0 || N > 128 || (K & 3)) return 0;
  return (K <= 64 || (N <= 64 && K <= 96)) ? 1 : 0;
}

extern "C" int bignn_gemm_tc_masked_f32(int32_t M, int32_t N, int32_t K, const float* A, int64_t lda, const float* B,
                                        int64_t ldb, int32_t b_is_nk, float* C, int64_t ldc, const float* bias,
                                        int32_t act, const float* mask_y, int64_t ldmy, int32_t mask_act,
                                        void* stream) {
  if (M < 0 || N < 0 || K < 0) return BIGNN_EINVAL;
  if (M == 0 || N == 0) return 0;
  if (K == 0 || !A || !B || !C || ldc < N || lda < K) return BIGNN_EINVAL;
  if (!bignn_gemm_tc_supported(M, N, K)) return BIGNN_EINVAL;   // two CTAs/SM: operands must fit in ~110 KB
  if ((lda & 3) || !aligned16(A)) return BIGNN_EALIGN;          // 128-bit cp.async of the A rows
  if (act < 0 || act > BIGNN_ACT_TANH) return BIGNN_EINVAL;
  if (mask_y && (ldmy < N || mask_act < 0 || mask_act > BIGNN_ACT_TANH)) return BIGNN_EINVAL;
  if (!mask_y || mask_act == BIGNN_ACT_IDENTITY) { mask_y = nullptr; mask_act = 0; }
  cudaStream_t st = (cudaStream_t)stream;
  if (N <= 32) return launch_tc_k<32>(M, N, K, A, lda, B, ldb, b_is_nk, C, ldc, bias, act, mask_y, ldmy, mask_act, st);
  if (N <= 64) return launch_tc_k<64>(M, N, K, A, lda, B, ldb, b_is_nk, C, ldc, bias, act, mask_y, ldmy, mask_act, st);
  if (N <= 96) return launch_tc_k<96>(M, N, K, A, lda, B, ldb, b_is_nk, C, ldc, bias, act, mask_y, ldmy, mask_act, st);
  return launch_tc_k<128>(M, N, K, A, lda, B, ldb, b_is_nk, C, ldc, bias, act, mask_y, ldmy, mask_act, st);
}

extern "C" int bignn_gemm_tc_f32(int32_t M, int32_t N, int32_t K, const float* A, int64_t lda, const float* B,
                                 int64_t ldb, int32_t b_is_nk, float* C, int64_t ldc, const float* bias, int32_t act,
                                 void* stream) {
  return bignn_gemm_tc_masked_f32(M, N, K, A, lda, B, ldb, b_is_nk, C, ldc, bias, act, nullptr, 0, 0, stream);
}
