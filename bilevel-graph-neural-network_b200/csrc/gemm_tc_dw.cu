// Weight gradients on the tensor cores:  D[Np, Nq] = P^T Q  (+ column sums of P or Q), reduction over
// the M rows (atoms / drugs), fp32-accurate through 3xTF32 (see include/bignn_b200.h: bignn_dw_tc_f32).
// Replaces autograd's dW = dY^T X / X^T dY and db = sum dY of nn.Linear / `x @ weight`
// (model/layers.py:26-30, PyG GCNConv/GATConv).
//
// A row tile [BM rows x 64 floats] staged as two [rows x 128 B] blocks is the canonical MN-major UMMA
// operand (MN = feature, K = row).  For 32-bit MN-major operands the only legal shared-memory layout is
// SWIZZLE_128B_BASE32B: 4-row groups (SBO = 512 B) whose four 32-byte chunks are XOR-permuted with
// the row index; the two 32-feature blocks are LBO = BM * 128 B apart; one K step (8 TF32) = two groups.  Per 32-row
// tile 4 K-steps x 3 products (hi*hi + lo*hi + hi*lo) accumulate into TMEM; accumulators persist across all tiles of the CTA and
// the per-CTA partial [64 x 64] goes to a workspace that a fixed-order reduction sums (deterministic).
// The MMA runs with M = 128: rows 64..127 of D come from the Q blocks that follow P in shared memory
// and are ignored (a 64-row MMA costs the same tensor time and has a scattered TMEM layout).
#include <cstdlib>
#include "tc_common.cuh"

namespace bignn {

constexpr int DW_F = 64;                       // padded feature width of both operands
constexpr int DW_MMA_M = 128;                  // MMA M (features of P in rows 0..63; rows 64..127 ignored)

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, int blk_bytes) {
  // MN-major, SWIZZLE_128B_BASE32B (layout type 1): LBO = distance between 32-feature blocks,
  // SBO = distance between 4-row groups
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(blk_bytes >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
         (1ull << 46) | (1ull << 61);
}

// byte offset of 16-byte chunk c (0..7) of row r inside a [rows x 128 B] block, Swizzle<2,5,2>:
// the 32-byte chunk index (c >> 1) is XORed with the row's position in its 4-row group
__device__ __forceinline__ uint32_t sw32b_off(int r, int c) {
  return (uint32_t)((r >> 2) * 512 + (r & 3) * 128 + ((((c >> 1) ^ (r & 3)) << 5) | ((c & 1) << 4)));
}

__device__ __forceinline__ uint32_t umma_idesc_tf32_mn(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------------------
// The kernel.  Its predecessors (128-row tiles, one CTA per SM: 1.22 ms per call at 6 M rows; 64-row double-buffered
// tiles, two CTAs per SM: 0.89 ms, ncu: no dominant stall, 23 % tensor-pipe activity, one 32 KB tile in flight per CTA,
// i.e. waiting for data) lost to this one and were removed in round 2.  Here the operand tiles are 32-row stages in an NST-deep cp.async ring
// (NST-2 tiles in flight beyond the one being split) and the lo operand is double buffered, so the hi/lo split of
// tile i+1 overlaps the MMAs of tile i (one mbarrier per lo buffer; a stage is refilled only after the MMAs that
// read it have committed).  96 KB of shared memory and 256 TMEM columns per CTA -> two CTAs per SM.
template <int BM, int NST, int DW_NA>
__global__ void __launch_bounds__(TC_THREADS, 2)
k_dw_tc_ring(int M, int Np, int Nq, const float* __restrict__ P, int64_t ldp, const float* __restrict__ Q, int64_t ldq,
             int colsum_of, float* __restrict__ ws_dw, double* __restrict__ ws_cs) {
  constexpr int DW_BLK = BM * 128;
  constexpr int DW_TILE = 2 * DW_BLK;
  constexpr int STAGE = 2 * DW_TILE;               // [P][Q]
  static_assert((DW_NA + 1) * DW_F <= 256 && NST >= 3, "two CTAs per SM");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // (aligned through an OFFSET from the __shared__ array: a round trip through uintptr_t makes every access below a
  // generic LD / ST instead of LDS / STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* lo_base = smem + NST * STAGE;           // two lo buffers of STAGE bytes
  __shared__ uint64_t mma_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ double cs_red[16][DW_F];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = (M + BM - 1) / BM;
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  constexpr int TMEM_COLS = 256;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    mbar_init(&mma_bar[0], 1);
    mbar_init(&mma_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t smem_s = smem_u32(smem);
  // this thread's four 16-byte chunks of every stage: row pr, chunk pc of the blocks (P | Q) x (features 0-31 | 32-63).
  // Everything that does not depend on the tile is computed once (ncu, r2: the index arithmetic of the copy loop was a
  // third of the 375 instructions a warp executed per 32-row tile, at 49 % issue utilisation)
  static_assert(BM == 32 && TC_THREADS == 256, "chunk mapping: 8 chunks per row, 32 rows, 4 blocks");
  const int pr = tid >> 3, pc = tid & 7;
  const uint32_t soff = sw32b_off(pr, pc);
  const bool colok0 = pc * 4 < Np, colok1 = 32 + pc * 4 < Np, colok2 = pc * 4 < Nq, colok3 = 32 + pc * 4 < Nq;
  const float* gp = P + (int64_t)pr * ldp + pc * 4;
  const float* gq = Q + (int64_t)pr * ldq + pc * 4;
  auto prefetch_tile = [&](int i, int stage) {       // i = this CTA's tile index
    const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * BM;
    const uint32_t d = smem_s + stage * STAGE + soff;
    const bool rowok = m0 + pr < M;
    const float* sp = gp + (int64_t)m0 * ldp;
    const float* sq = gq + (int64_t)m0 * ldq;
    cp_async16(d, (rowok && colok0) ? sp : P, (rowok && colok0) ? 16u : 0u);
    cp_async16(d + DW_BLK, (rowok && colok1) ? sp + 32 : P, (rowok && colok1) ? 16u : 0u);
    cp_async16(d + DW_TILE, (rowok && colok2) ? sq : Q, (rowok && colok2) ? 16u : 0u);
    cp_async16(d + DW_TILE + DW_BLK, (rowok && colok3) ? sq + 32 : Q, (rowok && colok3) ? 16u : 0u);
  };
  // prologue: tiles 0 .. NST-3 in flight, one commit group per tile (empty groups keep the count uniform)
#pragma unroll
  for (int j = 0; j < NST - 2; ++j) {
    if (j < my_tiles) prefetch_tile(j, j);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t idesc = umma_idesc_tf32_mn(DW_MMA_M, DW_F);

  // column sums (the bias gradient): thread = (16-byte column group, row group); fp32 over 8 tiles (16 values), then fp64
  double cs[4] = {0.0, 0.0, 0.0, 0.0};
  float4 csf = make_float4(0.f, 0.f, 0.f, 0.f);
  const int cs_cg = tid & 15, cs_rg = tid >> 4;
  const uint32_t cs_off0 = (cs_cg >> 3) * DW_BLK + sw32b_off(cs_rg, cs_cg & 7);
  const uint32_t cs_off1 = (cs_cg >> 3) * DW_BLK + sw32b_off(cs_rg + 16, cs_cg & 7);
  uint32_t ph0 = 0, ph1 = 0;
  int ks = 0;
  for (int i = 0; i < my_tiles; ++i) {
    const int stage = i % NST, lb = i & 1;
    uint8_t* p_hi = smem + stage * STAGE;
    uint8_t* p_lo = lo_base + lb * STAGE;
    // groups committed so far: (NST-2) + i; tile i is group i -> at most NST-3 newer groups may be pending
    asm volatile("cp.async.wait_group %0;" ::"n"(NST - 3) : "memory");
    if (i >= 2) {                                      // MMAs of tile i-2: they read lo[lb] and stage (i-2) % NST
      if (lb == 0) { mbar_wait(&mma_bar[0], ph0); ph0 ^= 1; } else { mbar_wait(&mma_bar[1], ph1); ph1 ^= 1; }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (i + NST - 2 < my_tiles) prefetch_tile(i + NST - 2, (i + NST - 2) % NST);     // = the stage of tile i-2
    asm volatile("cp.async.commit_group;" ::: "memory");
    // ---- lo = tf32(x - hi(x)) for the chunks this thread copied
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t off = (uint32_t)((k >> 1) * DW_TILE + (k & 1) * DW_BLK) + soff;
      const float4 x = *reinterpret_cast<const float4*>(p_hi + off);
      uint4 l;
      l.x = __float_as_uint(x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u)) & 0xffffe000u;
      l.y = __float_as_uint(x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u)) & 0xffffe000u;
      l.z = __float_as_uint(x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u)) & 0xffffe000u;
      l.w = __float_as_uint(x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u)) & 0xffffe000u;
      *reinterpret_cast<uint4*>(p_lo + off) = l;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t dp_hi = umma_desc_mn_sw128(smem_u32(p_hi), DW_BLK), dp_lo = umma_desc_mn_sw128(smem_u32(p_lo), DW_BLK);
      const uint64_t dq_hi = umma_desc_mn_sw128(smem_u32(p_hi + DW_TILE), DW_BLK),
                     dq_lo = umma_desc_mn_sw128(smem_u32(p_lo + DW_TILE), DW_BLK);
#pragma unroll 1
      for (int k = 0; k < BM / 8; ++k, ++ks) {
        const uint64_t adv = (uint64_t)((k * 1024) >> 4);
        umma_tf32(tmem_d + (uint32_t)((ks % DW_NA) * DW_F), dp_hi + adv, dq_hi + adv, idesc, ks >= DW_NA ? 1u : 0u);
        umma_tf32(tmem_d + (uint32_t)(DW_NA * DW_F), dp_lo + adv, dq_hi + adv, idesc, ks > 0 ? 1u : 0u);
        umma_tf32(tmem_d + (uint32_t)(DW_NA * DW_F), dp_hi + adv, dq_lo + adv, idesc, 1u);
      }
      umma_commit(lb == 0 ? &mma_bar[0] : &mma_bar[1]);
    } else if (tid >= 32) {
      ks += BM / 8;
    }
    if (colsum_of >= 0) {                              // from the raw tile, while the tensor core runs
      const uint8_t* t0 = p_hi + (colsum_of ? DW_TILE : 0);
      const float4 a = *reinterpret_cast<const float4*>(t0 + cs_off0);
      const float4 b = *reinterpret_cast<const float4*>(t0 + cs_off1);
      csf.x += a.x + b.x; csf.y += a.y + b.y; csf.z += a.z + b.z; csf.w += a.w + b.w;
      if ((i & 7) == 7) {
        cs[0] += (double)csf.x; cs[1] += (double)csf.y; cs[2] += (double)csf.z; cs[3] += (double)csf.w;
        csf = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  cs[0] += (double)csf.x; cs[1] += (double)csf.y; cs[2] += (double)csf.z; cs[3] += (double)csf.w;
  // drain: the MMAs of the last two tiles
  for (int i = (my_tiles >= 2 ? my_tiles - 2 : 0); i < my_tiles; ++i) {
    if ((i & 1) == 0) { mbar_wait(&mma_bar[0], ph0); ph0 ^= 1; } else { mbar_wait(&mma_bar[1], ph1); ph1 ^= 1; }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncthreads();
  ks = __shfl_sync(0xffffffffu, ks, 0);
  const int n_steps = my_tiles * (BM / 8);
  {
    const int q = warp & 3;
    if (q < 2) {
      const int row = q * 32 + lane;
#pragma unroll 1
      for (int cb = (warp >> 2) * 32; cb < DW_F; cb += 64) {
        uint32_t r[32];
        float v[32];
        const uint32_t tbase = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)cb;
        if (n_steps > 0) {
          tmem_ld32(tbase + (uint32_t)(DW_NA * DW_F), r);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
#pragma unroll 1
        for (int a = 0; a < DW_NA; ++a) {
          if (a >= n_steps) break;
          tmem_ld32(tbase + (uint32_t)(a * DW_F), r);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __fadd_rn(__uint_as_float(r[j]), v[j]);
        }
        if (row < Np) {
          float* dst = ws_dw + ((int64_t)blockIdx.x * Np + row) * Nq;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (cb + j < Nq) dst[cb + j] = v[j];
        }
      }
    }
  }
  if (colsum_of >= 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) cs_red[cs_rg][4 * cs_cg + j] = cs[j];
    __syncthreads();
    const int ncs = colsum_of ? Nq : Np;
    if (tid < ncs) {
      double t = 0.0;
#pragma unroll
      for (int g = 0; g < 16; ++g) t += cs_red[g][tid];
      ws_cs[(int64_t)blockIdx.x * ncs + tid] = t;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
  }
}

// fixed-order sums over the per-CTA partials: blocks [0, dw_blocks) finish 32 elements of dW each, the blocks after
// them 32 column sums (the bias gradient) each -- one launch for both
__global__ void __launch_bounds__(256)
k_dw_reduce(const float* __restrict__ ws_dw, int parts, int total, float* __restrict__ out, int dw_blocks,
            const double* __restrict__ ws_cs, int cols, float* __restrict__ cs_out) {
  __shared__ double red[8][33];
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  if ((int)blockIdx.x < dw_blocks) {
    float* redf = reinterpret_cast<float*>(&red[0][0]);
    const int i = blockIdx.x * 32 + tx;
    float s = 0.f;
    if (i < total) {
#pragma unroll 4
      for (int z = ty; z < parts; z += 8) s += __ldg(ws_dw + (int64_t)z * total + i);
    }
    redf[ty * 33 + tx] = s;
    __syncthreads();
    if (ty == 0 && i < total) {
      float t = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) t += redf[g * 33 + tx];
      out[i] = t;
    }
  } else {
    const int c = ((int)blockIdx.x - dw_blocks) * 32 + tx;
    double s = 0.0;
    if (c < cols) {
#pragma unroll 4
      for (int q = ty; q < parts; q += 8) s += ws_cs[(int64_t)q * cols + c];
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && c < cols) {
      double t = 0.0;
#pragma unroll
      for (int g = 0; g < 8; ++g) t += red[g][tx];
      cs_out[c] = (float)t;
    }
  }
}

constexpr int DW_BM = 32, DW_NST = 4, DW_ACC = 3;      // rows per stage, ring depth, rotating hi*hi accumulators
constexpr int DW_SMEM = 6 * DW_NST * DW_BM * 128 + 1024;

static int dw_grid(int M) {
  const int n_tiles = ceil_div(M, DW_BM);
  const int g = sm_count() * 2;
  return g > n_tiles ? n_tiles : g;
}

}  // namespace bignn

using namespace bignn;

extern "C" int bignn_dw_tc_supported(int32_t M, int32_t Np, int32_t Nq) {
  return (M > 0 && Np > 0 && Nq > 0 && Np <= DW_F && Nq <= DW_F && (Np % 4) == 0 && (Nq % 4) == 0) ? 1 : 0;
}

extern "C" int64_t bignn_dw_tc_workspace_bytes(int32_t M, int32_t Np, int32_t Nq) {
  if (!bignn_dw_tc_supported(M, Np, Nq)) return 0;
  const int64_t g = dw_grid(M);
  return g * Np * Nq * (int64_t)sizeof(float) + g * DW_F * (int64_t)sizeof(double) + 64;
}

extern "C" int bignn_dw_tc_f32(int32_t M, int32_t Np, int32_t Nq, const float* P, int64_t ldp, const float* Q,
                               int64_t ldq, float* D, int32_t colsum_of, float* colsum, void* workspace,
                               int64_t workspace_bytes, void* stream) {
  if (!bignn_dw_tc_supported(M, Np, Nq)) return BIGNN_EINVAL;
  if (!P || !Q || !D || ldp < Np || ldq < Nq) return BIGNN_EINVAL;
  if (colsum_of > 1 || (colsum_of >= 0 && !colsum)) return BIGNN_EINVAL;
  if ((ldp & 3) || (ldq & 3) || !aligned16(P) || !aligned16(Q)) return BIGNN_EALIGN;
  if (!workspace || workspace_bytes < bignn_dw_tc_workspace_bytes(M, Np, Nq)) return BIGNN_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = dw_grid(M);
  float* ws_dw = (float*)workspace;
  double* ws_cs = (double*)((uint8_t*)workspace + (((int64_t)grid * Np * Nq * sizeof(float) + 15) & ~(int64_t)15));
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_dw_tc_ring<DW_BM, DW_NST, DW_ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  k_dw_tc_ring<DW_BM, DW_NST, DW_ACC><<<grid, TC_THREADS, DW_SMEM, st>>>(M, Np, Nq, P, ldp, Q, ldq, colsum_of < 0 ? -1 : colsum_of, ws_dw, ws_cs);
  const int ncs = colsum_of < 0 ? 0 : (colsum_of ? Nq : Np);
  const int dw_blocks = ceil_div(Np * Nq, 32);
  k_dw_reduce<<<dw_blocks + ceil_div(ncs, 32), 256, 0, st>>>(ws_dw, grid, Np * Nq, D, dw_blocks, ws_cs, ncs, colsum);
  BIGNN_LAUNCH_COUNT(2);
  return last_launch_status();
}
