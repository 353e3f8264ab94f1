// Fused gather-and-score edge decoder (see include/bignn_b200.h: bignn_pair_decoder_fwd / _bwd).
//
// Replaces model/layers_link_pred.py:43-65 (F.normalize of all drug embeddings, gather of the two rows of every
// pair, concat, the `mlp_concat` MLP, sigmoid) AND the loss head model/layers.py:79-89 (BCELoss / BCEWithLogitsLoss /
// CrossEntropyLoss, mean reduced) -- forward in ONE launch, backward in ONE launch (+ the per-drug row sum that
// transposes the gather).  The reference normalises all N rows and then reads 2P of them; here only the gathered rows
// are normalised.
//
// One warp per pair.  A lane owns the columns k = lane, lane+32, .. of the concatenated vector [z_a || z_b]
// (2D <= 256), so the two embedding rows are read coalesced, the L2 norms are warp reductions, the first layer is n1
// warp reductions (n1 <= 16) and the later layers (<= 32 wide) are computed by one lane per output.  The loss terms are
// accumulated in fp64 per warp; the last CTA to finish adds the per-warp partial sums in warp order (deterministic; no
// float atomics).  Backward: per-pair chain rule in the same mapping; every warp keeps its partial weight gradients in
// registers (dW0: n1 x (2D/32) per lane) across its pairs, writes them once, and the last CTA adds the partials in warp
// order.
#include "common.cuh"

namespace bignn {

constexpr int PD_MAX_KPL = 8;      // columns of the concatenated vector per lane (2D <= 256)
constexpr int PD_MAX_N1 = 16;      // width of the first hidden layer
constexpr int PD_WARPS = 4;        // warps per CTA

struct PairDecoderArgs {
  const float* H; int64_t ldh; const int32_t* ids; int P, D;
  const float* W[3]; const float* b[3]; int n[3]; int n_layers;      // widths after each layer; W[l] is [n[l], n[l-1]]
  int head;                                                           // 0 sigmoid+BCE, 1 logits+BCEWithLogits, 2 logits+CE
  const float* y; const int32_t* labels;                              // targets (float for BCE, int32 for CE); null: no loss
  float* scores; int64_t lds;                                         // [P, n_out]
  float* nrm; float* h1; float* h2;                                   // saved for backward: [P,2], [P,n1], [P,n2]
  float* loss; double* ws_loss; unsigned int* counter;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// loads the pair's two rows, normalises them; returns z[i] for column k = lane + 32 i and the clamped norms
__device__ __forceinline__ void load_pair(const PairDecoderArgs& a, int p, int lane, int kpl, float (&z)[PD_MAX_KPL],
                                          float& na, float& nb) {
  const int D = a.D;
  const float* ha = a.H + (int64_t)a.ids[2 * p] * a.ldh;
  const float* hb = a.H + (int64_t)a.ids[2 * p + 1] * a.ldh;
  float ssa = 0.f, ssb = 0.f;
#pragma unroll
  for (int i = 0; i < PD_MAX_KPL; ++i) {
    z[i] = 0.f;
    const int k = lane + 32 * i;
    if (i < kpl && k < 2 * D) {
      z[i] = k < D ? __ldg(ha + k) : __ldg(hb + k - D);
      if (k < D) ssa = fmaf(z[i], z[i], ssa); else ssb = fmaf(z[i], z[i], ssb);
    }
  }
  ssa = warp_sum(ssa); ssb = warp_sum(ssb);
  na = fmaxf(__fsqrt_rn(ssa), 1e-12f);                     // F.normalize: x / max(|x|_2, eps)
  nb = fmaxf(__fsqrt_rn(ssb), 1e-12f);
#pragma unroll
  for (int i = 0; i < PD_MAX_KPL; ++i) {
    const int k = lane + 32 * i;
    if (i < kpl && k < 2 * D) z[i] = __fdiv_rn(z[i], k < D ? na : nb);
  }
}

__global__ void __launch_bounds__(PD_WARPS * 32)
k_pair_decoder_fwd(const PairDecoderArgs a) {
  __shared__ float hbuf[PD_WARPS][2][32];
  __shared__ int is_last;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int gw = blockIdx.x * PD_WARPS + w, n_warps = gridDim.x * PD_WARPS;
  const int kpl = (2 * a.D + 31) / 32;
  const int n1 = a.n[0], nL = a.n_layers;
  const int n_out = a.n[nL - 1];
  double lsum = 0.0;
  for (int p = gw; p < a.P; p += n_warps) {
    float z[PD_MAX_KPL], na, nb;
    load_pair(a, p, lane, kpl, z, na, nb);
    if (lane == 0 && a.nrm) { a.nrm[2 * p] = na; a.nrm[2 * p + 1] = nb; }
    // ---- layer 1: n1 dot products over the 2D columns (warp reductions)
    float mine = 0.f;                                      // lane j keeps output j
    for (int j = 0; j < n1; ++j) {
      const float* wr = a.W[0] + (int64_t)j * 2 * a.D;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < PD_MAX_KPL; ++i) {
        const int k = lane + 32 * i;
        if (i < kpl && k < 2 * a.D) s = fmaf(__ldg(wr + k), z[i], s);
      }
      s = warp_sum(s);
      if (lane == j) mine = s + (a.b[0] ? __ldg(a.b[0] + j) : 0.f);
    }
    float v = mine;                                        // pre-activation of this lane's output of the current layer
    int width = n1;
    for (int l = 1; l < nL; ++l) {
      // activation of the previous layer (relu on every layer but the last, model/layers_util.py:44-57)
      const float hprev = lane < width ? fmaxf(v, 0.f) : 0.f;
      hbuf[w][l & 1][lane] = hprev;
      if (l == 1 && a.h1 && lane < width) a.h1[(int64_t)p * width + lane] = hprev;
      if (l == 2 && a.h2 && lane < width) a.h2[(int64_t)p * width + lane] = hprev;
      __syncwarp();
      const int nw = a.n[l];
      float s = 0.f;
      if (lane < nw) {
        const float* wr = a.W[l] + (int64_t)lane * width;
        s = a.b[l] ? __ldg(a.b[l] + lane) : 0.f;
        for (int m = 0; m < width; ++m) s = fmaf(__ldg(wr + m), hbuf[w][l & 1][m], s);
      }
      v = s;
      width = nw;
      __syncwarp();
    }
    // ---- head
    float out = v;
    if (a.head == 0) out = 1.0f / (1.0f + expf(-v));       // sigmoid (model/layers_link_pred.py:65)
    if (lane < n_out) a.scores[(int64_t)p * a.lds + lane] = out;
    if (a.loss) {
      float term = 0.f;
      if (a.head == 0) {                                   // torch BCELoss: logs clamped at -100
        if (lane == 0) {
          const float t = a.y[p];
          term = (t - 1.0f) * fmaxf(log1pf(-out), -100.f) - t * fmaxf(logf(out), -100.f);
        }
      } else if (a.head == 1) {                            // torch BCEWithLogitsLoss
        if (lane == 0) {
          const float t = a.y[p], m = fmaxf(-v, 0.f);
          term = (1.0f - t) * v + m + logf(expf(-m) + expf(-v - m));
        }
      } else {                                             // torch CrossEntropyLoss: logsumexp - x[label]
        float mx = lane < n_out ? v : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        // (sequential sum over the classes, as k_ce_fwd does)
        float zsum = 0.f;
        for (int k = 0; k < n_out; ++k) zsum += expf(__shfl_sync(0xffffffffu, v, k) - mx);
        const float xl = __shfl_sync(0xffffffffu, v, a.labels[p]);
        if (lane == 0) term = logf(zsum) + mx - xl;
      }
      if (lane == 0) lsum += (double)term;
    }
  }
  if (a.loss) {
    if (lane == 0) a.ws_loss[gw] = lsum;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
      __threadfence();
      double s = 0.0;
      for (int i = 0; i < n_warps; ++i) s += a.ws_loss[i];             // warp order: deterministic
      *a.loss = (float)(s / (double)(a.P > 0 ? a.P : 1));
      *a.counter = 0u;                                                  // ready for the next launch
    }
  }
}

struct PairDecoderBwdArgs {
  PairDecoderArgs f;
  const float* dloss;                 // scalar
  float* drows; int64_t lddr;         // [2P, D]: gradient w.r.t. the gathered (un-normalised) rows
  float* dW[3]; float* db[3];         // outputs
  float* ws; int64_t ws_stride;       // per-warp partial parameter gradients
};

__global__ void __launch_bounds__(PD_WARPS * 32)
k_pair_decoder_bwd(const PairDecoderBwdArgs g) {
  const PairDecoderArgs& a = g.f;
  __shared__ float sbuf[PD_WARPS][3][32];      // [0] h1, [1] h2 / dh (current), [2] d of the layer below
  __shared__ int is_last;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int gw = blockIdx.x * PD_WARPS + w, n_warps = gridDim.x * PD_WARPS;
  const int D = a.D, kpl = (2 * D + 31) / 32;
  const int nL = a.n_layers, n1 = a.n[0], n2 = nL == 3 ? a.n[1] : 0, n_out = a.n[nL - 1];
  const float gl = *g.dloss / (float)(a.P > 0 ? a.P : 1);
  // per-warp partial gradients, in registers across the warp's pairs
  float dW0[PD_MAX_N1][PD_MAX_KPL];
  float db0 = 0.f, dWm[8], dbm = 0.f, dWl[8], dbl = 0.f;      // lane l owns elements e = l + 32 t of the small matrices
#pragma unroll
  for (int j = 0; j < PD_MAX_N1; ++j)
#pragma unroll
    for (int i = 0; i < PD_MAX_KPL; ++i) dW0[j][i] = 0.f;
#pragma unroll
  for (int t = 0; t < 8; ++t) { dWm[t] = 0.f; dWl[t] = 0.f; }
  const int wl_in = nL == 3 ? n2 : n1;                          // input width of the last layer
  for (int p = gw; p < a.P; p += n_warps) {
    float z[PD_MAX_KPL], na, nb;
    load_pair(a, p, lane, kpl, z, na, nb);
    const float h1 = lane < n1 ? a.h1[(int64_t)p * n1 + lane] : 0.f;
    const float h2 = (nL == 3 && lane < n2) ? a.h2[(int64_t)p * n2 + lane] : 0.f;
    // ---- d loss / d (last layer's pre-activation), lane j < n_out
    float dl = 0.f;
    {
      const float s = lane < n_out ? a.scores[(int64_t)p * a.lds + lane] : 0.f;
      if (a.head == 0) {           // BCELoss backward, then sigmoid backward
        if (lane == 0) { const float t = a.y[p]; dl = gl * (s - t) / fmaxf((1.0f - s) * s, 1e-12f) * ((1.0f - s) * s); }
      } else if (a.head == 1) {
        if (lane == 0) dl = (1.0f / (1.0f + expf(-s)) - a.y[p]) * gl;
      } else {
        float mx = lane < n_out ? s : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float zsum = 0.f;
        for (int k = 0; k < n_out; ++k) zsum += expf(__shfl_sync(0xffffffffu, s, k) - mx);
        if (lane < n_out) dl = gl * (expf(s - mx) / zsum - (lane == a.labels[p] ? 1.f : 0.f));
      }
    }
    sbuf[w][0][lane] = h1;
    sbuf[w][1][lane] = nL == 3 ? h2 : h1;         // input of the last layer
    sbuf[w][2][lane] = dl;
    __syncwarp();
    // ---- last layer: parameter gradients (element e = lane + 32 t -> (j, m) = (e / wl_in, e % wl_in)), input gradient
    const int last = nL - 1;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int e = lane + 32 * t;
      if (e < n_out * wl_in) dWl[t] = fmaf(sbuf[w][2][e / wl_in], sbuf[w][1][e % wl_in], dWl[t]);
    }
    if (lane < n_out) dbl += dl;
    float dprev = 0.f;                             // gradient w.r.t. the last layer's input (post-activation), lane m < wl_in
    if (lane < wl_in) {
      for (int j = 0; j < n_out; ++j) dprev = fmaf(__ldg(a.W[last] + (int64_t)j * wl_in + lane), sbuf[w][2][j], dprev);
      dprev = sbuf[w][1][lane] > 0.f ? dprev : 0.f;           // relu backward
    }
    __syncwarp();
    float dh1 = dprev;                             // (two-layer scorer: the last layer's input is h1)
    if (nL == 3) {
      sbuf[w][2][lane] = dprev;                    // = d pre-activation of layer 2, lane j < n2
      __syncwarp();
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int e = lane + 32 * t;
        if (e < n2 * n1) dWm[t] = fmaf(sbuf[w][2][e / n1], sbuf[w][0][e % n1], dWm[t]);
      }
      if (lane < n2) dbm += dprev;
      dh1 = 0.f;
      if (lane < n1) {
        for (int j = 0; j < n2; ++j) dh1 = fmaf(__ldg(a.W[1] + (int64_t)j * n1 + lane), sbuf[w][2][j], dh1);
        dh1 = h1 > 0.f ? dh1 : 0.f;
      }
      __syncwarp();
    }
    // ---- first layer: dW0[j][k] += d1[j] z[k]; dz[k] = sum_j W0[j][k] d1[j]
    if (lane < n1) db0 += dh1;
    float dz[PD_MAX_KPL];
#pragma unroll
    for (int i = 0; i < PD_MAX_KPL; ++i) dz[i] = 0.f;
#pragma unroll
    for (int j = 0; j < PD_MAX_N1; ++j) {
      if (j < n1) {
        const float d1 = __shfl_sync(0xffffffffu, dh1, j);
        const float* wr = a.W[0] + (int64_t)j * 2 * D;
#pragma unroll
        for (int i = 0; i < PD_MAX_KPL; ++i) {
          const int k = lane + 32 * i;
          if (i < kpl && k < 2 * D) {
            dW0[j][i] = fmaf(d1, z[i], dW0[j][i]);
            dz[i] = fmaf(__ldg(wr + k), d1, dz[i]);
          }
        }
      }
    }
    // ---- through the two normalisations: dRow = (dz - zhat <dz, zhat>) / nrm   (dz / nrm through the clamp branch)
    float dota = 0.f, dotb = 0.f;
#pragma unroll
    for (int i = 0; i < PD_MAX_KPL; ++i) {
      const int k = lane + 32 * i;
      if (i < kpl && k < 2 * D) { if (k < D) dota = fmaf(dz[i], z[i], dota); else dotb = fmaf(dz[i], z[i], dotb); }
    }
    dota = warp_sum(dota); dotb = warp_sum(dotb);
#pragma unroll
    for (int i = 0; i < PD_MAX_KPL; ++i) {
      const int k = lane + 32 * i;
      if (i < kpl && k < 2 * D) {
        const bool sa = k < D;
        const float n = sa ? na : nb, dot = sa ? dota : dotb;
        const float v = n <= 1e-12f ? __fdiv_rn(dz[i], n) : __fdiv_rn(dz[i] - z[i] * dot, n);
        g.drows[(int64_t)(2 * p + (sa ? 0 : 1)) * g.lddr + (sa ? k : k - D)] = v;
      }
    }
  }
  // ---- per-warp partials -> shared memory, added in warp order into ONE partial per CTA; the last CTA to finish adds
  // the per-CTA partials in CTA order (deterministic; two short levels instead of one long one)
  extern __shared__ float part_s[];                 // [PD_WARPS][ws_stride]
  {
    float* o = part_s + (int64_t)w * g.ws_stride;
    int off = 0;
#pragma unroll
    for (int j = 0; j < PD_MAX_N1; ++j)
#pragma unroll
      for (int i = 0; i < PD_MAX_KPL; ++i) {
        const int k = lane + 32 * i;
        if (j < n1 && i < kpl && k < 2 * D) o[off + j * 2 * D + k] = dW0[j][i];
      }
    off += n1 * 2 * D;
    if (lane < n1) o[off + lane] = db0;
    off += n1;
    if (nL == 3) {
#pragma unroll
      for (int t = 0; t < 8; ++t) { const int e = lane + 32 * t; if (e < n2 * n1) o[off + e] = dWm[t]; }
      off += n2 * n1;
      if (lane < n2) o[off + lane] = dbm;
      off += n2;
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) { const int e = lane + 32 * t; if (e < n_out * wl_in) o[off + e] = dWl[t]; }
    off += n_out * wl_in;
    if (lane < n_out) o[off + lane] = dbl;
  }
  __syncthreads();
  const int total = (int)g.ws_stride;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PD_WARPS; ++i) s += part_s[(int64_t)i * g.ws_stride + e];
    g.ws[(int64_t)blockIdx.x * g.ws_stride + e] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    __threadfence();
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
      float s = 0.f;
      for (int i = 0; i < (int)gridDim.x; ++i) s += g.ws[(int64_t)i * g.ws_stride + e];
      // scatter into the caller's gradient tensors
      int off = e;
      if (off < n1 * 2 * D) { g.dW[0][off] = s; continue; }
      off -= n1 * 2 * D;
      if (off < n1) { if (g.db[0]) g.db[0][off] = s; continue; }
      off -= n1;
      if (nL == 3) {
        if (off < n2 * n1) { g.dW[1][off] = s; continue; }
        off -= n2 * n1;
        if (off < n2) { if (g.db[1]) g.db[1][off] = s; continue; }
        off -= n2;
      }
      if (off < n_out * wl_in) { g.dW[nL - 1][off] = s; continue; }
      off -= n_out * wl_in;
      if (g.db[nL - 1]) g.db[nL - 1][off] = s;
    }
    if (threadIdx.x == 0) *a.counter = 0u;
  }
}

static int pd_grid(int P) {
  int g = ceil_div(P, PD_WARPS);
  const int cap = 2 * sm_count();
  if (g > cap) g = cap;
  return g < 1 ? 1 : g;
}

static int pd_params(const int* n, int n_layers, int D) {
  int t = n[0] * 2 * D + n[0];
  for (int l = 1; l < n_layers; ++l) t += n[l] * n[l - 1] + n[l];
  return t;
}

}  // namespace bignn

using namespace bignn;

extern "C" int bignn_pair_decoder_supported(int32_t D, int32_t n_layers, int32_t n1, int32_t n2, int32_t n3) {
  if (D <= 0 || 2 * D > 32 * PD_MAX_KPL) return 0;
  if (n_layers != 2 && n_layers != 3) return 0;
  if (n1 <= 0 || n1 > PD_MAX_N1) return 0;
  const int n_out = n_layers == 3 ? n3 : n2;
  const int mid = n_layers == 3 ? n2 : n1;
  if (n_out <= 0 || n_out > 32 || mid <= 0 || mid > 32) return 0;
  if (n_layers == 3 && n2 * n1 > 256) return 0;
  if (n_out * mid > 256) return 0;
  return 1;
}

extern "C" int64_t bignn_pair_decoder_workspace_bytes(int32_t P, int32_t D, int32_t n_layers, int32_t n1, int32_t n2,
                                                      int32_t n3) {
  const int n[3] = {n1, n2, n3};
  const int64_t warps = (int64_t)pd_grid(P) * PD_WARPS;
  // [counter (16 B)] [loss partials: warps doubles] [parameter-gradient partials: warps x params floats]
  return 16 + warps * 8 + warps * (int64_t)pd_params(n, n_layers, D) * 4 + 64;
}

static int pd_fill(PairDecoderArgs& a, const float* H, int64_t ldh, const int32_t* ids, int32_t P, int32_t D,
                   int32_t n_layers, const float* W0, const float* b0, int32_t n1, const float* W1, const float* b1,
                   int32_t n2, const float* W2, const float* b2, int32_t n3, int32_t head, const float* y,
                   const int32_t* labels, float* scores, int64_t lds, float* nrm, float* h1, float* h2, void* ws,
                   int64_t ws_bytes) {
  if (P < 0 || !bignn_pair_decoder_supported(D, n_layers, n1, n2, n3)) return BIGNN_EINVAL;
  if (!H || !ids || !W0 || !W1 || (n_layers == 3 && !W2) || !scores || ldh < D) return BIGNN_EINVAL;
  if (head < 0 || head > 2) return BIGNN_EINVAL;
  if (!ws || ws_bytes < bignn_pair_decoder_workspace_bytes(P, D, n_layers, n1, n2, n3)) return BIGNN_EWORKSPACE;
  a.H = H; a.ldh = ldh; a.ids = ids; a.P = P; a.D = D;
  a.W[0] = W0; a.b[0] = b0; a.W[1] = W1; a.b[1] = b1; a.W[2] = W2; a.b[2] = b2;
  a.n[0] = n1; a.n[1] = n2; a.n[2] = n3; a.n_layers = n_layers; a.head = head; a.y = y; a.labels = labels;
  a.scores = scores; a.lds = lds; a.nrm = nrm; a.h1 = h1; a.h2 = h2;
  a.counter = (unsigned int*)ws;
  a.ws_loss = (double*)((uint8_t*)ws + 16);
  a.loss = nullptr;
  return 0;
}

extern "C" int bignn_pair_decoder_fwd(const float* H, int64_t ldh, const int32_t* ids, int32_t P, int32_t D,
                                      int32_t n_layers, const float* W0, const float* b0, int32_t n1, const float* W1,
                                      const float* b1, int32_t n2, const float* W2, const float* b2, int32_t n3,
                                      int32_t head, const float* y, const int32_t* labels, float* scores, int64_t lds,
                                      float* nrm, float* h1, float* h2, float* loss, void* workspace,
                                      int64_t workspace_bytes, void* stream) {
  PairDecoderArgs a;
  const int rc = pd_fill(a, H, ldh, ids, P, D, n_layers, W0, b0, n1, W1, b1, n2, W2, b2, n3, head, y, labels, scores, lds,
                         nrm, h1, h2, workspace, workspace_bytes);
  if (rc) return rc;
  if (P == 0) return 0;
  const int n_out = n_layers == 3 ? n3 : n2;
  if (lds < n_out) return BIGNN_EINVAL;
  if (loss && ((head == 2 && !labels) || (head != 2 && !y))) return BIGNN_EINVAL;
  if (head != 2 && n_out != 1 && loss) return BIGNN_EINVAL;
  a.loss = loss;
  k_pair_decoder_fwd<<<pd_grid(P), PD_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}

extern "C" int bignn_pair_decoder_bwd(const float* H, int64_t ldh, const int32_t* ids, int32_t P, int32_t D,
                                      int32_t n_layers, const float* W0, int32_t n1, const float* W1, int32_t n2,
                                      const float* W2, int32_t n3, int32_t head, const float* y, const int32_t* labels,
                                      const float* scores, int64_t lds, const float* nrm, const float* h1,
                                      const float* h2, const float* dloss, float* drows, int64_t lddr, float* dW0,
                                      float* db0, float* dW1, float* db1, float* dW2, float* db2, void* workspace,
                                      int64_t workspace_bytes, void* stream) {
  PairDecoderBwdArgs g;
  const int rc = pd_fill(g.f, H, ldh, ids, P, D, n_layers, W0, nullptr, n1, W1, nullptr, n2, W2, nullptr, n3, head, y,
                         labels, (float*)scores, lds, (float*)nrm, (float*)h1, (float*)h2, workspace, workspace_bytes);
  if (rc) return rc;
  if (P == 0) return 0;
  if (!dloss || !drows || lddr < D || !dW0 || !dW1 || (n_layers == 3 && !dW2) || !h1 || (n_layers == 3 && !h2) || !nrm)
    return BIGNN_EINVAL;
  if ((head == 2 && !labels) || (head != 2 && !y)) return BIGNN_EINVAL;
  const int n[3] = {n1, n2, n3};
  const int grid = pd_grid(P);
  g.dloss = dloss; g.drows = drows; g.lddr = lddr;
  g.dW[0] = dW0; g.dW[1] = dW1; g.dW[2] = dW2; g.db[0] = db0; g.db[1] = db1; g.db[2] = db2;
  g.ws_stride = pd_params(n, n_layers, D);
  g.ws = (float*)((uint8_t*)workspace + 16 + (int64_t)grid * PD_WARPS * 8);
  const int smem = PD_WARPS * (int)g.ws_stride * (int)sizeof(float);
  static int configured_smem = 0;
  if (smem > configured_smem) {
    cudaError_t e = cudaFuncSetAttribute(k_pair_decoder_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured_smem = smem;
  }
  k_pair_decoder_bwd<<<grid, PD_WARPS * 32, smem, (cudaStream_t)stream>>>(g);
  BIGNN_LAUNCH_COUNT(1);
  return last_launch_status();
}
