// Fused graph stage on the Blackwell tensor cores (tcgen05 + TMEM + bulk-TMA), sm_100a.
// Reference math: CE.forward, DN_Gray/model/dagl.py:250-272 (scores, adaptive neighbour mask,
// un-renormalised masked softmax, weighted aggregation of 7x7x16 value patches, fold).
//
// Design (see DESIGN.md §4 for the derivation and the measured MMA cost model):
//  * CTA = 128 queries x one half of the 784 value columns (TMEM holds 512 fp32 columns, so the
//    128x784 accumulator is split over two CTAs) x one slice of the keys (split-K; partials are
//    merged by merge_fold.cu).
//  * Keys are enumerated over the zero-padded width Wp = W+6 ("padded-flat" order).  Then the value
//    patch of key k' at shift (dy,dx) is theta_pad_flat[k' + dy*Wp + dx]: the whole N_k x 784 value
//    operand is a Toeplitz view of the 16-channel theta map and is never materialised.  In smem a
//    7-row halo of theta lies as [pixel][16 ch] (32 B per pixel, SWIZZLE_32B, MN-major); one
//    tcgen05.mma with N = 16*G whose N-group stride is one pixel covers G dx-shifts x 16 channels.
//  * Scores need fp32 accuracy: Q and K are split into fp16 hi + lo parts after a
//    power-of-two rescale, S = Qh.Kh + Qh.Kl + Ql.Kh (three kind::f16 MMAs, fp32 accumulate).
//  * P (fp16) is written back into the S columns of TMEM and used as the A operand of P.V.
//  * Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2-13 softmax/epilogue (a thread owns
//    one query row x 16 key columns; three warps share a TMEM lane quadrant), mbarrier pipelines
//    between them; S is double buffered so the scores of tile j+1 are computed while tile j goes
//    through the softmax.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc_utils.cuh"

namespace dagl {
using namespace tc;

constexpr int TC_BM = 128;
constexpr int TC_BN = 48;
constexpr int TC_EP = 208;                        // 196 padded to 13 k-steps of 16
constexpr int TC_KSTEPS = TC_EP / 16;
constexpr int TC_ECH = TC_EP / 8;                 // 26 16-byte chunks per row
constexpr int Q_HALF_BYTES = TC_BM * TC_EP * 2;   // 53248
constexpr int Q_TILE_BYTES = 2 * Q_HALF_BYTES;    // hi | lo
constexpr int K_HALF_BYTES = TC_BN * TC_EP * 2;   // 19968
constexpr int K_TILE_BYTES = 2 * K_HALF_BYTES;    // 39936
constexpr int TH_SEG_PIX = 64;
constexpr int TH_SEG_BYTES = TH_SEG_PIX * 32;     // 2048
constexpr int TH_SLOTS = 4;                       // dy rows per half
constexpr int TH_STAGE_BYTES = TH_SLOTS * TH_SEG_BYTES;
constexpr int TC_THREADS = 448;                 // 1 TMA + 1 MMA + 12 softmax warps
constexpr int TC_TMEM_COLS = 512;
constexpr int TC_S_COL0 = 400;                    // S/P buffers at columns 400 and 448

constexpr int SM_Q = 0;
constexpr int SM_K = SM_Q + Q_TILE_BYTES;                 // 2 stages
constexpr int SM_T = SM_K + 2 * K_TILE_BYTES;             // 2 stages
static_assert(SM_K % 1024 == 0 && SM_T % 1024 == 0 && K_TILE_BYTES % 1024 == 0, "smem carve-up alignment");

struct TcGeom {
  int Wp;        // padded width W + 6
  int NkP;       // number of padded-flat key slots = (H-1)*Wp + W
  int NT;        // key tiles of 48
  int NP;        // pixels in the packed theta array
  int nqt;       // query tiles of 128
};

static TcGeom tc_geom(const Geom& g) {
  TcGeom t;
  t.Wp = (g.W + 2 * PADK + 7) & ~7;   // multiple of 8: every dy row of a key tile starts at the same 8-pixel phase
  t.NkP = (g.H - 1) * t.Wp + g.W;
  t.NT = (t.NkP + TC_BN - 1) / TC_BN;
  int np = (g.H + 2 * PADK) * t.Wp;
  int need = TC_BN * (t.NT + 3) + 2 * PADK * t.Wp + TH_SEG_PIX + 8;   // v4 loads the theta rows of 4 tiles in one copy
  t.NP = ((np > need ? np : need) + 7) & ~7;
  t.nqt = (g.Nq + TC_BM - 1) / TC_BM;
  return t;
}

// power-of-two scale that brings absmax just below 2^target
__device__ __forceinline__ float pow2_scale(unsigned absmax_bits, int target) {
  const float a = __uint_as_float(absmax_bits);
  if (!(a > 0.f) || !isfinite(a)) return 1.f;
  int e;
  frexpf(a, &e);                      // a = m * 2^e, m in [0.5, 1)
  return ldexpf(1.f, target - e);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// byte offset of the 16-byte chunk (row, chunk kc) inside a K-major no-swizzle tile half of `rows` rows
__device__ __forceinline__ uint32_t tile_chunk_off(int row, int kc, int rows) {
  return (uint32_t)(kc * (rows / 8) * 128 + (row / 8) * 128 + (row % 8) * 16);
}

// ---------------------------------------------------------------------------------------------
// absmax reduction for the split entry (when the embeddings come from the caller)
// ---------------------------------------------------------------------------------------------
__global__ void absmax_kernel(const float* __restrict__ x, size_t n_per_img, unsigned* __restrict__ absmax, int slot) {
  const int img = blockIdx.y;
  const float* xi = x + (size_t)img * n_per_img;
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_img; i += (size_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(xi[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(absmax + img * AMAX_STRIDE + slot, __float_as_uint(m));
}

// ---------------------------------------------------------------------------------------------
// pack: fp32 rows [*,196] -> fp16 hi/lo UMMA tiles.  One CTA per tile; rows staged through smem so
// both the global reads and the tile writes are coalesced.
//   MODE 0: queries, ROWS=128, row r of tile t <- Q[t*128 + r]; also emits tA = mu*gamma, tB = beta
//   MODE 1: keys, ROWS=48, padded-flat slot k' = t*48 + r <- K[y*W + x] if x < W (y = k'/Wp, x = k'%Wp)
// ---------------------------------------------------------------------------------------------
template <int ROWS, int MODE, int RPB /*rows per block; ROWS % RPB == 0*/>
__global__ void __launch_bounds__(256)
pack_tiles_kernel(Geom g, TcGeom tg, const float* __restrict__ src, const unsigned* __restrict__ absmax,
                  uint8_t* __restrict__ tiles, unsigned long long* __restrict__ tilemask,
                  const float* __restrict__ Kbar, const float* __restrict__ gamma, const float* __restrict__ beta,
                  float4* __restrict__ thr4, float* __restrict__ colsum_partial) {
  pdl_prologue();
  extern __shared__ __align__(16) float rows_s[];            // [RPB][196]
  constexpr int BPT = ROWS / RPB;                            // blocks per tile
  const int t = blockIdx.x / BPT, r0 = (blockIdx.x % BPT) * RPB, img = blockIdx.y, tid = threadIdx.x;
  const int ntile = gridDim.x / BPT;
  const int nrows_src = MODE == 0 ? g.Nq : g.Nk;
  const float* si = src + (size_t)img * nrows_src * ED;
  const float scale = pow2_scale(absmax[img * AMAX_STRIDE + MODE], 14);

  for (int i = tid; i < RPB * (ED / 4); i += 256) {
    const int r = r0 + i / (ED / 4), e4 = i % (ED / 4);
    int srow = -1;
    if (MODE == 0) {
      const int q = t * ROWS + r;
      if (q < g.Nq) srow = q;
    } else {
      const int kp = t * ROWS + r;
      const int y = kp / tg.Wp, x = kp % tg.Wp;
      if (kp < tg.NkP && x < g.W) srow = y * g.W + x;
    }
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (srow >= 0) v = __ldg(reinterpret_cast<const float4*>(si + (size_t)srow * ED) + e4);
    reinterpret_cast<float4*>(rows_s)[i] = v;
  }
  __syncthreads();

  uint8_t* tile = tiles + ((size_t)img * ntile + t) * (size_t)(2 * ROWS * TC_EP * 2);
  constexpr int HALF = ROWS * TC_EP * 2;
  // one thread per 16-byte output chunk; consecutive threads -> consecutive rows of one chunk column (coalesced)
  for (int o = tid; o < TC_ECH * RPB; o += 256) {
    const int kc = o / RPB, rl = o % RPB;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x0 = 0.f, x1 = 0.f;
      const int e = kc * 8 + 2 * j;
      if (e < ED) { x0 = rows_s[rl * ED + e] * scale; x1 = rows_s[rl * ED + e + 1] * scale; }
      const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
      const __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1));
      hi[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      lo[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    const uint32_t off = tile_chunk_off(r0 + rl, kc, ROWS);
    *reinterpret_cast<uint4*>(tile + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(tile + HALF + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  if (MODE == 1) {
    if (tid == 0) {                                  // validity bits of the 48 key slots of this tile
      unsigned long long m = 0ull;
      for (int r = 0; r < ROWS; ++r) {
        const int kp = t * ROWS + r;
        if (kp < tg.NkP && (kp % tg.Wp) < g.W) m |= 1ull << r;
      }
      tilemask[(size_t)img * ntile + t] = m;
    }
    // column sums of this tile's rows for Kbar = mean_k K (dummy slots are zero rows); fixed order -> deterministic
    if (colsum_partial != nullptr)
      for (int e = tid; e < ED; e += 256) {
        float sum = 0.f;
#pragma unroll 8
        for (int r = 0; r < RPB; ++r) sum += rows_s[r * ED + e];
        colsum_partial[((size_t)img * ntile + t) * ED + e] = sum;
      }
  } else {
    // per-query threshold terms: mu = Q[q,:] . Kbar (fp64 accumulate), tA = mu*gamma, tB = beta  (dagl.py:256)
    const int lane = tid & 31, warp = tid >> 5;
    for (int rl = warp; rl < RPB; rl += 8) {
      const int r = r0 + rl;
      const int q = t * ROWS + r;
      double s = 0.0;
      for (int e = lane; e < ED; e += 32) s += (double)rows_s[rl * ED + e] * (double)__ldg(Kbar + (size_t)img * ED + e);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) {
        const size_t idx = ((size_t)img * ntile + t) * ROWS + r;
        const float mu = (float)s;
        thr4[idx] = (q < g.Nq) ? make_float4(mu, 0.f, __ldg(gamma + (size_t)img * g.Nq + q), __ldg(beta + (size_t)img * g.Nq + q))
                               : make_float4(0.f, 0.f, 0.f, -1.f);
      }
    }
  }
}

// theta [16][H][W] fp32 -> zero-padded flat [NP pixels][16 ch] fp16 (scaled), SWIZZLE_32B pre-applied:
// the two 16-byte halves of a pixel are swapped when (pixel & 4), which is address bit 7 once a
// segment that starts at a multiple of 8 pixels lands on a 256-byte aligned smem address.
__global__ void __launch_bounds__(256)
pack_theta_kernel(Geom g, TcGeom tg, const float* __restrict__ theta, const unsigned* __restrict__ absmax,
                  uint8_t* __restrict__ thp, unsigned long long* __restrict__ tilemask /*nullable*/) {
  pdl_prologue();
  const int img = blockIdx.y;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  if (pix >= tg.NP) return;
  if (tilemask != nullptr && pix < tg.NT) {            // validity bits of the 48 key slots of tile `pix` (NT < NP)
    unsigned long long m = 0ull;
    int kp = pix * TC_BN, x = kp % tg.Wp;
    for (int r = 0; r < TC_BN; ++r, ++kp) {
      if (kp < tg.NkP && x < g.W) m |= 1ull << r;
      if (++x == tg.Wp) x = 0;
    }
    tilemask[(size_t)img * tg.NT + pix] = m;
  }
  const float scale = pow2_scale(absmax[img * AMAX_STRIDE + AMAX_THETA], 12);
  const int r = pix / tg.Wp, cc = pix % tg.Wp;
  const int y = r - PADK, x = cc - PADK;
  const bool inb = (y >= 0 && y < g.H && x >= 0 && x < g.W);
  uint32_t w[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float a = 0.f, b = 0.f;
    if (inb) {
      a = __ldg(theta + (((size_t)img * CI + 2 * j) * g.H + y) * g.W + x) * scale;
      b = __ldg(theta + (((size_t)img * CI + 2 * j + 1) * g.H + y) * g.W + x) * scale;
    }
    w[j] = pack_half2(a, b);
  }
  uint4* dst = reinterpret_cast<uint4*>(thp + ((size_t)img * tg.NP + pix) * 32);
  const int sw = (pix >> 2) & 1;
  dst[sw] = make_uint4(w[0], w[1], w[2], w[3]);
  dst[sw ^ 1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// ---------------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------------
#ifdef DAGL_TC_TRACE
// development aid (tools/tc_trace.py): per-CTA cycle counters of the pipeline roles
__device__ long long g_tc_trace[1024][24];
__device__ long long g_tc_tl[4][32][24];   // tc4 timeline: [cluster rank][round - TL_P0][event], cycles since CTA start (cluster 0 only)
#define TL_P0 100
#define TL(round, ev) do { const int _r = (round) - TL_P0; if (tl_on && _r >= 0 && _r < 32) g_tc_tl[rank][_r][ev] = clock64() - tl_base; } while (0)
__device__ int g_tc_dbg_mode = 0;      // bit0: skip the S MMAs, bit1: skip the P.V MMAs (timing experiments only)
#define TRACE_T0() long long _t0 = clock64()
#define TRACE_ADD(var) (var) += clock64() - _t0
#else
#define TRACE_T0()
#define TRACE_ADD(var)
#define TL(round, ev)
#endif

struct PvGroup { int dy, dx0, n, col0; };
// half 0: dy 0,1,2 full (3 x 112 cols) + dy 3 dx 0..3 (64)      = 400 columns
// half 1: dy 3 dx 4..6 (48) + dy 4,5,6 full (3 x 112)            = 384 columns
__constant__ PvGroup c_groups[2][4] = {
    {{0, 0, 112, 0}, {1, 0, 112, 112}, {2, 0, 112, 224}, {3, 0, 64, 336}},
    {{3, 4, 48, 0}, {4, 0, 112, 48}, {5, 0, 112, 160}, {6, 0, 112, 272}}};

// ---------------------------------------------------------------------------------------------
// Softmax reference of a query row (shared by the clustered kernels and the exact pass below).
//
// The pre-pass gives smax_hi = max_k Qh.Kh.  With |x - hi(x)| <= 2^-11 |x| for both operands, S <= Qh.Kh * (1 + 2^-10 +
// 2^-22); RM_INFL = 1 + 1.25 * 2^-10 also covers the fp32 accumulation of the 13 k-steps.  The logit c*s*relu(s - T) is
// monotone in s >= 0, so ref = logit(RM_INFL * smax_hi) - 12 bounds every term: P <= 2^12, inside fp16.
// The logit is quadratic in s, so the bound overshoots the true row maximum by up to ~2^-8 of the maximum logit.  That
// is < 1 log2 unit for the trained heads (logits <= 140), but for logits of several thousand units (random-init networks
// a few stages deep, rgb_range = 255 models, diverging training) every fp16 P would drift into the subnormal range and
// finally flush to zero.  Query tiles with a row whose possible overshoot exceeds RM_REFINE_LOG2 units get a second,
// EXACT pass (rowmax_tc_kernel<1>): the same three split-fp16 MMA terms in the same order as the graph kernel, so the
// scores are the very values the softmax will see, and the row maximum of the LOGIT is stored in smax2.  With it the row
// maximum lands on 2^12 exactly, whatever the magnitude.
// ---------------------------------------------------------------------------------------------
constexpr int TOPK_MAX = 64;                                 // legacy fixed-top-k variant: at most this many edges per query
constexpr float RM_INFL = 1.f + 1.25f / 1024.f;
constexpr float RM_REFINE_LOG2 = 12.f;

// The softmax logit of dagl.py:256-260, z = c * S * relu(S - T) (`topk` == 0), or of the legacy fixed-top-k variant
// (GReccR2b_3mh_1-checkpoint.py:243-250), z = c * S * [S among the row's k largest] with the k-th largest score folded into
// T (`topk` != 0); both are monotone in S >= 0.
__device__ __forceinline__ float logit_factor(float s, float rl, int topk) { return topk ? (rl != 0.f ? s : 0.f) : s * rl; }
__device__ __forceinline__ float row_logit_log2(float s, float tA, float tB, float sm_scale_log2, int topk) {
  const float rl = fmaxf((s - tA) + tB, 0.f);
  return logit_factor(s, rl, topk) * sm_scale_log2;
}
__device__ __forceinline__ bool row_needs_refine(float s_hi /*unscaled Qh.Kh maximum*/, float tA, float tB, float sm_scale_log2, int topk) {
  return row_logit_log2(s_hi * RM_INFL, tA, tB, sm_scale_log2, topk) - row_logit_log2(s_hi * (2.f - RM_INFL), tA, tB, sm_scale_log2, topk) >
         RM_REFINE_LOG2;
}
// smax2_bits != 0: exact maximum of the logit (log2 units) from the exact pass
__device__ __forceinline__ float row_softmax_ref(unsigned smax_bits, unsigned smax2_bits, float inv_s, float tA, float tB,
                                                 float sm_scale_log2, int topk) {
  // (1 + 2^-22): the stored maximum is itself rounded (half an ulp: 4 log2 units at logits of 1e8), and 2^(12 + 4) would
  // overflow fp16; two to four ulps of head-room keep P <= 2^12 at every magnitude
  if (smax2_bits != 0u) return __uint_as_float(smax2_bits) * (1.f + 1.f / 4194304.f) - 12.f;
  return row_logit_log2(__uint_as_float(smax_bits) * inv_s * RM_INFL, tA, tB, sm_scale_log2, topk) - 12.f;
}

// Legacy fixed-top-k variant: per query row, the k-th largest score over all key splits becomes the selection threshold.
// thr4 = (T, 0, 1, 0) makes the graph kernels' relu(S - T) positive exactly for S >= tau (T = the float below tau); rows
// with fewer than k valid keys select everything.  Ties AT the k-th value are all kept (torch.topk keeps an arbitrary
// subset of them).
__global__ void topk_merge_kernel(Geom g, TcGeom tg, int nlists, int k, const float* __restrict__ lists,
                                  const unsigned* __restrict__ absmax, float4* __restrict__ thr4) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.B * tg.nqt * TC_BM) return;
  const int img = i / (tg.nqt * TC_BM), q = i % (tg.nqt * TC_BM);
  if (q >= g.Nq) { thr4[i] = make_float4(0.f, 0.f, 0.f, -1.f); return; }
  const float inv_s = 1.f / (pow2_scale(absmax[img * AMAX_STRIDE + AMAX_Q], 14) * pow2_scale(absmax[img * AMAX_STRIDE + AMAX_K], 14));
  const float* src = lists + (size_t)i * nlists * TOPK_MAX;
  float sel[TOPK_MAX];
  int nsel = 0, imin = 0;
  float vmin = 0.f;
  for (int j = 0; j < nlists * TOPK_MAX; ++j) {
    const float x = src[j];
    if (x < 0.f) continue;                                   // unused entry (scores are >= 0)
    if (nsel < k) {
      sel[nsel++] = x;
      if (nsel == k) { vmin = sel[0]; imin = 0; for (int u = 1; u < k; ++u) if (sel[u] < vmin) { vmin = sel[u]; imin = u; } }
    } else if (x > vmin) {
      sel[imin] = x;
      vmin = sel[0]; imin = 0;
      for (int u = 1; u < k; ++u) if (sel[u] < vmin) { vmin = sel[u]; imin = u; }
    }
  }
  float T = -1.f;                                            // fewer than k valid keys: every key is selected (S >= 0 > T)
  if (nsel == k) T = nextafterf(vmin * inv_s, -INFINITY);
  thr4[i] = make_float4(T, 0.f, 1.f, 0.f);
}

// =============================================================================================
// v2: 2-CTA cluster per query tile.  The two value-column halves (CTA rank 0 / 1) no longer both
// compute the scores: CTA h owns the key tiles j with (j & 1) == h, runs S + softmax for them and
// writes the fp16 P tile into BOTH CTAs' shared memory (DSMEM); both CTAs run P.V for every tile on
// their own half of the value columns (P read from smem; the four value-column groups of a k-step
// share it through the A collector).  Sharing P across CTAs needs one softmax reference per query row
// that both agree on, so the reference is fixed up front: a cheap pre-pass (rowmax_tc_kernel, Qh.Kh
// only) gives max_k S[q,k] to fp16 accuracy, and exponent(s) = c*s*relu(s - T) is monotone in s >= 0,
// so ref_q = exponent(1.004 * smax_q) bounds every term.  No online maximum, no accumulator rescale,
// and key-split partials merge by plain sums.
// =============================================================================================
constexpr int RM_QT = 2;                                    // query tiles per CTA (each K tile is fetched once for both)
constexpr int RM_KSTAGES = 4;
constexpr int RM_SM_K = 0;                                  // ring of Kh tiles (19968 B each); exact pass: 3 stages of hi|lo tiles
constexpr int RM_SM_BAR = RM_SM_K + RM_KSTAGES * K_HALF_BYTES;
constexpr int RM_SM_TOTAL = RM_SM_BAR + 128;
constexpr int RMX_KSTAGES = 3;
constexpr int RMX_SM_BAR = RM_SM_K + RMX_KSTAGES * K_TILE_BYTES;
constexpr int RMX_SM_TOTAL = RMX_SM_BAR + 128;
constexpr int RM_THREADS = 192;
constexpr int RM_QCOL = 0;                                  // Qh of the two query tiles: 2 x 104 TMEM columns (exact: Qh | Ql of one)
constexpr int RM_DCOL0 = RM_QT * (TC_EP / 2);               // 208: three accumulator buffers of RM_QT x 48 columns
constexpr int RM_DBUF = 3;
constexpr int RM_DCOLS = RM_QT * TC_BN;                     // 96

// EXACT = false (always runs): smax[b][qt*128 + row] = max over the keys of (Qh . Kh) in scaled units (>= 0); atomicMax on
//   float bits.  The query tiles live in TMEM (A operand of the MMAs, TS form: 27.7 instead of ~51 cycles per N = 48 MMA)
//   and a CTA serves two query tiles per K fetch (the pre-pass would otherwise be bound by the L2 -> SM traffic of the K tiles).
//   TMEM: Qh(q0) [0,104) | Qh(q1) [104,208) | D0 [208,304) | D1 [304,400) | D2 [400,496)
// EXACT = true (always launched; a CTA whose query tile has no row flagged by row_needs_refine exits at once): one query
//   tile per CTA, the full 3-term score Ql.Kh + Qh.Kl + Qh.Kh with the MMA sequence of the graph kernels, and
//   smax2[row] = max over the keys of the LOGIT c * S * relu(S - T) in log2 units (what the softmax exponent will be).
//   TMEM: Qh [0,104) | Ql [104,208) | D0 [208,256) | D1 [304,352) | D2 [400,448)
// MODE 2 (legacy fixed-top-k variant only): the exact scores again, but instead of the row maximum every row keeps its
//   `topk` largest VALID scores (scaled units) of this CTA's key range in a small per-thread selection buffer and writes
//   them to lists[row][split][TOPK_MAX] (unused entries -1); topk_merge_kernel turns them into the per-row threshold.
template <int MODE>
__global__ void __launch_bounds__(RM_THREADS, 1)
rowmax_tc_kernel(TcGeom tg, const uint8_t* __restrict__ Qp, const uint8_t* __restrict__ Kp, int nsplit,
                 int qt_base, int qt_end, unsigned* __restrict__ smax, const float4* __restrict__ thr4,
                 const unsigned* __restrict__ absmax, float sm_scale_log2, int topk,
                 unsigned* __restrict__ smax2, const unsigned long long* __restrict__ tilemask, float* __restrict__ lists) {
  pdl_prologue();
  constexpr bool EXACT = MODE != 0;
  constexpr int QT = EXACT ? 1 : RM_QT;
  constexpr int KST = EXACT ? RMX_KSTAGES : RM_KSTAGES;
  constexpr int KBYTES = EXACT ? K_TILE_BYTES : K_HALF_BYTES;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (EXACT ? RMX_SM_BAR : RM_SM_BAR));
  uint64_t* q_ready = bars + 0;   // 128 arrivals: the query tiles are in TMEM
  uint64_t* k_full = bars + 1;    // [4]
  uint64_t* k_empty = bars + 5;   // [4]
  uint64_t* d_full = bars + 9;    // [3]
  uint64_t* d_empty = bars + 12;  // [3]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 15);
  const int warp = warp_id_uniform();
  const int tid = threadIdx.x;
  const int qt0 = qt_base + blockIdx.x * QT, split = blockIdx.y, img = blockIdx.z;
  const int nq_here = min(QT, qt_end - qt0);
  const int t_begin = (int)(((long long)split * tg.NT) / nsplit);
  const int t_end = (int)(((long long)(split + 1) * tg.NT) / nsplit);
  const int nsteps = t_end - t_begin;                       // one key tile per step
  const float inv_s = EXACT ? 1.f / (pow2_scale(absmax[img * AMAX_STRIDE + AMAX_Q], 14) * pow2_scale(absmax[img * AMAX_STRIDE + AMAX_K], 14))
                            : 1.f;
  if (MODE == 1) {
    // does any row of this query tile need the exact maximum?  (block-uniform decision, before anything is allocated)
    bool flag = false;
    if (tid < TC_BM) {
      const size_t qi = ((size_t)img * tg.nqt + qt0) * TC_BM + tid;
      const float4 t4 = __ldg(thr4 + qi);
      flag = row_needs_refine(__uint_as_float(__ldg(smax + qi)) * inv_s, (t4.x + t4.y) * t4.z, t4.w, sm_scale_log2, topk);
    }
    if (!__syncthreads_or(flag ? 1 : 0)) return;
  }

  if (tid == 0) {
    mbar_init(q_ready, 128);
    for (int i = 0; i < KST; ++i) { mbar_init(k_full + i, 1); mbar_init(k_empty + i, 1); }
    for (int i = 0; i < RM_DBUF; ++i) { mbar_init(d_full + i, 1); mbar_init(d_empty + i, 4); }
    mbar_init_fence();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      const uint8_t* ksrc = Kp + ((size_t)img * tg.NT + t_begin) * K_TILE_BYTES;
      for (int st = 0; st < nsteps; ++st, ksrc += K_TILE_BYTES) {
        const int s = st % KST;
        mbar_wait(k_empty + s, ((uint32_t)(st / KST) & 1u) ^ 1u);
        mbar_arrive_expect_tx(k_full + s, KBYTES);
        bulk_g2s(smem + RM_SM_K + s * KBYTES, ksrc, KBYTES, k_full + s);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idS = instr_desc(128, TC_BN, FMT_F16, FMT_F16, 0, 0);
      const bool two_q = nq_here > 1;
      mbar_wait(q_ready, 0);
      tc_fence_after();
      for (int st = 0; st < nsteps; ++st) {
        const int s = st % KST, db = st % RM_DBUF;
        mbar_wait(k_full + s, (uint32_t)(st / KST) & 1u);
        mbar_wait(d_empty + db, ((uint32_t)(st / RM_DBUF) & 1u) ^ 1u);
        tc_fence_after();
        const uint64_t dk = smem_desc(smem_u32(smem + RM_SM_K + s * KBYTES), (TC_BN / 8) * 128, 128);
        const uint32_t d0 = tbase + RM_DCOL0 + db * RM_DCOLS;
        if (EXACT) {
          // the MMA sequence of the graph kernels (attend_tc4_kernel score issuer): Ql.Kh first, then Qh.Kl / Qh.Kh
          const uint64_t dk_lo = smem_desc(smem_u32(smem + RM_SM_K + s * KBYTES + K_HALF_BYTES), (TC_BN / 8) * 128, 128);
          const uint32_t qh = tbase + RM_QCOL, ql = qh + TC_EP / 2;
#pragma unroll
          for (int ks = 0; ks < TC_KSTEPS; ++ks) {
            const uint64_t ko = (uint64_t)(ks * 2 * (TC_BN / 8) * 128 >> 4);
            mma_f16_ts(d0, ql + ks * 8, dk + ko, idS, ks > 0);
          }
#pragma unroll
          for (int ks = 0; ks < TC_KSTEPS; ++ks) {
            const uint64_t ko = (uint64_t)(ks * 2 * (TC_BN / 8) * 128 >> 4);
            mma_f16_ts(d0, qh + ks * 8, dk_lo + ko, idS, 1);
            mma_f16_ts(d0, qh + ks * 8, dk + ko, idS, 1);
          }
        } else {
#pragma unroll
          for (int ks = 0; ks < TC_KSTEPS; ++ks) {
            const uint64_t ko = (uint64_t)(ks * 2 * (TC_BN / 8) * 128 >> 4);
            mma_f16_ts(d0, tbase + RM_QCOL + ks * 8, dk + ko, idS, ks > 0);
            if (two_q) mma_f16_ts(d0 + TC_BN, tbase + RM_QCOL + TC_EP / 2 + ks * 8, dk + ko, idS, ks > 0);
          }
        }
        mma_commit(k_empty + s);
        mma_commit(d_full + db);
      }
    }
  } else {
    const int quad = warp & 3, lane = tid & 31;
    const int row = quad * 32 + lane;
    const uint32_t trow = tbase + ((uint32_t)(quad * 32) << 16);
    // ---- query tiles -> TMEM: lane = row, column j = elements (2j, 2j+1); slab = (hi part of tile qi) or (hi / lo part) ----
    constexpr int NSLAB = 2;
#pragma unroll 1
    for (int sl = 0; sl < NSLAB; ++sl) {
      if (!EXACT && sl >= nq_here) break;
      const uint8_t* src = Qp + ((size_t)img * tg.nqt + qt0 + (EXACT ? 0 : sl)) * Q_TILE_BYTES + (EXACT ? sl * Q_HALF_BYTES : 0) + row * 16;
      uint4 c0[TC_KSTEPS], c1[TC_KSTEPS];                  // all 26 loads in flight before the first TMEM store
#pragma unroll
      for (int ks = 0; ks < TC_KSTEPS; ++ks) {
        c0[ks] = __ldg(reinterpret_cast<const uint4*>(src + (2 * ks) * (TC_BM / 8) * 128));
        c1[ks] = __ldg(reinterpret_cast<const uint4*>(src + (2 * ks + 1) * (TC_BM / 8) * 128));
      }
#pragma unroll
      for (int ks = 0; ks < TC_KSTEPS; ++ks) {
        const uint32_t v[8] = {c0[ks].x, c0[ks].y, c0[ks].z, c0[ks].w, c1[ks].x, c1[ks].y, c1[ks].z, c1[ks].w};
        tmem_st8(trow + RM_QCOL + sl * (TC_EP / 2) + ks * 8, v);
      }
    }
    tmem_wait_st();
    tc_fence_before();
    mbar_arrive(q_ready);
    const size_t qidx = ((size_t)img * tg.nqt + qt0) * TC_BM + row;
    const float4 t4 = MODE == 1 ? __ldg(thr4 + qidx) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float tA = (t4.x + t4.y) * t4.z, tB = t4.w;             // T = mu * gamma - beta (dagl.py:256), mu from two column halves
    float m[RM_QT][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};      // four chains per row: a single one is latency-bound
    float sel[MODE == 2 ? TOPK_MAX : 1];                                   // MODE 2: the row's largest scores so far (unordered)
    int nsel = 0, imin = 0;
    float vmin = 0.f;
    for (int st = 0; st < nsteps; ++st) {
      const int db = st % RM_DBUF;
      mbar_wait(d_full + db, (uint32_t)(st / RM_DBUF) & 1u);
      tc_fence_after();
#pragma unroll
      for (int qi = 0; qi < QT; ++qi) {
        if (qi >= nq_here) break;
        uint32_t v[48];
#pragma unroll
        for (int c0 = 0; c0 < TC_BN; c0 += 16) {
          uint32_t t16[16];
          tmem_ld16(trow + RM_DCOL0 + db * RM_DCOLS + qi * TC_BN + c0, t16);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[c0 + i] = t16[i];
        }
        tmem_wait_ld();
        if (MODE == 2) {
          const unsigned long long vb = __ldg(tilemask + (size_t)img * tg.NT + t_begin + st);
#pragma unroll 1
          for (int i = 0; i < TC_BN; ++i) {
            if (!((vb >> i) & 1ull)) continue;                             // dummy key slot
            const float x = __uint_as_float(v[i]);
            if (nsel < topk) {
              sel[nsel++] = x;
              if (nsel == topk) { vmin = sel[0]; imin = 0; for (int j = 1; j < topk; ++j) if (sel[j] < vmin) { vmin = sel[j]; imin = j; } }
            } else if (x > vmin) {
              sel[imin] = x;
              vmin = sel[0]; imin = 0;
              for (int j = 1; j < topk; ++j) if (sel[j] < vmin) { vmin = sel[j]; imin = j; }
            }
          }
        } else if (EXACT) {
          // dummy key slots are zero rows: S = 0, logit 0 <= the maximum (logits are >= 0)
#pragma unroll
          for (int i = 0; i < TC_BN; ++i)
            m[qi][i & 3] = fmaxf(m[qi][i & 3], row_logit_log2(__uint_as_float(v[i]) * inv_s, tA, tB, sm_scale_log2, topk));
        } else {
#pragma unroll
          for (int i = 0; i < TC_BN; ++i) m[qi][i & 3] = fmaxf(m[qi][i & 3], __uint_as_float(v[i]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d_empty + db);
    }
    if (MODE == 2) {
      float* dst = lists + ((((size_t)img * tg.nqt + qt0) * TC_BM + row) * gridDim.y + split) * TOPK_MAX;
      for (int j = 0; j < TOPK_MAX; ++j) dst[j] = j < nsel ? sel[j] : -1.f;
    } else {
#pragma unroll
      for (int qi = 0; qi < QT; ++qi)
        if (qi < nq_here)
          atomicMax((EXACT ? smax2 : smax) + ((size_t)img * tg.nqt + qt0 + qi) * TC_BM + row,
                    __float_as_uint(fmaxf(fmaxf(m[qi][0], m[qi][1]), fmaxf(m[qi][2], m[qi][3]))));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tbase);
}

constexpr int TC2_THREADS = TC_THREADS + 32;                // + one warp that forwards P tiles to the peer CTA
constexpr int P_SLOT_BYTES = TC_BM * TC_BN * 2;             // 12288: one P tile, K-major no-swizzle A operand
constexpr int S2_Q = 0;
constexpr int S2_K = S2_Q + Q_TILE_BYTES;
constexpr int S2_T = S2_K + 2 * K_TILE_BYTES;
constexpr int S2_P = S2_T + 2 * TH_STAGE_BYTES;
constexpr int S2_BAR = S2_P + 2 * P_SLOT_BYTES;
constexpr int S2_RED = S2_BAR + 256;                         // [4 quad][3 sub][32] floats, then the same as ints
constexpr int S2_TOTAL = S2_RED + 2 * 4 * 3 * 32 * 4;
static_assert(S2_P % 1024 == 0, "P slots alignment");
static_assert(S2_TOTAL <= 232448, "v2 kernel exceeds the 227 KB dynamic shared memory limit");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC2_THREADS, 1)
attend_tc2_kernel(Geom g, TcGeom tg, const uint8_t* __restrict__ Qp, const uint8_t* __restrict__ Kp,
                  const uint8_t* __restrict__ Thp, const unsigned long long* __restrict__ tilemask,
                  const float4* __restrict__ thr4 /*per query row: mu partials (x, y), gamma, beta*/,
                  const unsigned* __restrict__ absmax, const unsigned* __restrict__ smax, const unsigned* __restrict__ smax2,
                  float sm_scale_log2, int topk,
                  int nsplit, int qt_base, float* __restrict__ Opart, float* __restrict__ lpart /*[B][nsplit][2][Nq]*/,
                  uint32_t* __restrict__ mask_bits, int32_t* __restrict__ nnz,
                  int t_cut /*> 0: two unequal key splits [0, t_cut) | [t_cut, NT)*/, int split_base, int tail) {
  // tail != 0: this launch runs NEXT TO the 4-CTA kernel (its programmatic dependent: launched once every CTA of that grid
  // has started, onto the SMs the 4-CTA clusters cannot use) on the last part of the keys.  Its inputs were complete before
  // the 4-CTA kernel triggered it, so it never waits (a CTA parked in griddepcontrol.wait would keep its SM from the next
  // tail cluster); the kernel after it is launched WITHOUT the programmatic attribute, i.e. after both grids.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (!tail) asm volatile("griddepcontrol.wait;" ::: "memory");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S2_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2] ring over OWN tiles
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* t_full = bars + 5;    // [2] ring over ALL tiles
  uint64_t* t_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2] S buffers (own tiles)
  uint64_t* s_free = bars + 11;   // [2] the softmax warps have read the S buffer
  uint64_t* p_full = bars + 13;   // [2] slot r is written by CTA rank r (into both CTAs)
  uint64_t* p_free = bars + 15;   // [2] slot r consumed by the P.V of both CTAs (only slot `rank` is waited on here)
  uint64_t* pv_last = bars + 17;  // final P.V complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = warp_id_uniform();
  const int tid = threadIdx.x;
  const int img = blockIdx.z, split = blockIdx.y + split_base;
  const int qt = qt_base + (blockIdx.x >> 1);
  const int half = (int)cluster_ctarank();                  // == blockIdx.x & 1
  const int t_begin = t_cut > 0 ? (split ? t_cut : 0) : (int)(((long long)split * tg.NT) / nsplit);
  const int t_end = t_cut > 0 ? (split ? tg.NT : t_cut) : (int)(((long long)(split + 1) * tg.NT) / nsplit);
  const int ntiles = t_end - t_begin;
  const int n_own = (ntiles - half + 1) / 2;                // local tiles j with (j & 1) == half
#ifdef DAGL_TC_TRACE
  long long tr_a = 0, tr_b = 0, tr_c = 0;
  const long long tr_start = clock64();
  const int tr_cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
#endif

  if (tid == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(k_full + i, 1); mbar_init(k_empty + i, 1);
      mbar_init(t_full + i, 1); mbar_init(t_empty + i, 1);
      mbar_init(s_full + i, 1); mbar_init(s_free + i, 384);
      // P slot i is produced by CTA rank i: locally by 384 softmax threads, remotely by one bulk DSMEM copy (tx bytes)
      mbar_init(p_full + i, i == half ? 384 : 1); mbar_init(p_free + i, 2);
    }
    mbar_init(pv_last, 1);
    mbar_init_fence();
  }
  if (warp == 1) tmem_alloc<TC_TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                       // peer barriers are initialised before any remote access
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const uint8_t* qsrc = Qp + ((size_t)img * tg.nqt + qt) * Q_TILE_BYTES;
      mbar_arrive_expect_tx(q_full, Q_TILE_BYTES);
      bulk_g2s(smem + S2_Q, qsrc, Q_HALF_BYTES, q_full);
      bulk_g2s(smem + S2_Q + Q_HALF_BYTES, qsrc + Q_HALF_BYTES, Q_HALF_BYTES, q_full);
      const uint8_t* thp = Thp + (size_t)img * tg.NP * 32;
      auto load_k = [&](int i) {                            // own tile i  (local tile 2i + half)
        const int s = i & 1;
        mbar_wait(k_empty + s, ((uint32_t)(i >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(k_full + s, K_TILE_BYTES);
        bulk_g2s(smem + S2_K + s * K_TILE_BYTES, Kp + ((size_t)img * tg.NT + t_begin + 2 * i + half) * K_TILE_BYTES,
                 K_TILE_BYTES, k_full + s);
      };
      auto load_t = [&](int j) {
        const int s = j & 1, t = t_begin + j;
        mbar_wait(t_empty + s, ((uint32_t)(j >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(t_full + s, TH_STAGE_BYTES);
#pragma unroll
        for (int sl = 0; sl < TH_SLOTS; ++sl) {
          const int first = (t * TC_BN + c_groups[half][sl].dy * tg.Wp) & ~7;
          bulk_g2s(smem + S2_T + s * TH_STAGE_BYTES + sl * TH_SEG_BYTES, thp + (size_t)first * 32, TH_SEG_BYTES, t_full + s);
        }
      };
      if (n_own > 0) load_k(0);
      for (int p = 0; 2 * p < ntiles; ++p) {
        if (p + 1 < n_own) load_k(p + 1);
        load_t(2 * p);
        if (2 * p + 1 < ntiles) load_t(2 * p + 1);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idS = instr_desc(128, TC_BN, FMT_F16, FMT_F16, 0, 0);
      const uint32_t q_hi = smem_u32(smem + S2_Q), q_lo = q_hi + Q_HALF_BYTES;
      const uint64_t dq_hi = smem_desc(q_hi, (TC_BM / 8) * 128, 128);
      const uint64_t dq_lo = smem_desc(q_lo, (TC_BM / 8) * 128, 128);
      const uint32_t p_free_prod[2] = {mapa(smem_u32(p_free + 0), 0), mapa(smem_u32(p_free + 1), 1)};
      mbar_wait(q_full, 0);
      tc_fence_after();

      auto issue_S = [&](int i) {                           // own tile i
        const int s = i & 1;
        const uint32_t ph = (uint32_t)(i >> 1) & 1u;
        { TRACE_T0(); mbar_wait(k_full + s, ph); TRACE_ADD(tr_a); }
        { TRACE_T0(); mbar_wait(s_free + s, ph ^ 1u); TRACE_ADD(tr_a); }   // softmax has pulled the previous contents of this S buffer
        tc_fence_after();
        const uint32_t k_hi = smem_u32(smem + S2_K + s * K_TILE_BYTES), k_lo = k_hi + K_HALF_BYTES;
        const uint64_t dk_hi = smem_desc(k_hi, (TC_BN / 8) * 128, 128);
        const uint64_t dk_lo = smem_desc(k_lo, (TC_BN / 8) * 128, 128);
        const uint32_t d = tbase + TC_S_COL0 + s * TC_BN;
#pragma unroll
        for (int ks = 0; ks < TC_KSTEPS; ++ks) {
          const uint64_t qo = (uint64_t)(ks * 2 * (TC_BM / 8) * 128 >> 4);
          const uint64_t ko = (uint64_t)(ks * 2 * (TC_BN / 8) * 128 >> 4);
          mma_f16_ss(d, dq_lo + qo, dk_hi + ko, idS, ks > 0);             // Ql.Kh first (small terms; see v1)
        }
#pragma unroll
        for (int ks = 0; ks < TC_KSTEPS; ++ks) {
          const uint64_t qo = (uint64_t)(ks * 2 * (TC_BM / 8) * 128 >> 4);
          const uint64_t ko = (uint64_t)(ks * 2 * (TC_BN / 8) * 128 >> 4);
          mma_f16_ss_a_fill(d, dq_hi + qo, dk_lo + ko, idS, 1);           // Qh.Kl
          mma_f16_ss_a_lastuse(d, dq_hi + qo, dk_hi + ko, idS, 1);        // Qh.Kh (A from the collector)
        }
        mma_commit(s_full + s);
        mma_commit(k_empty + s);
      };

      // per-CTA constants of the four value-column groups (single-thread issue: keep the loop body lean)
      uint32_t g_idesc[TH_SLOTS], g_col[TH_SLOTS];
      int g_dywp[TH_SLOTS], g_dx0[TH_SLOTS];
#pragma unroll
      for (int sl = 0; sl < TH_SLOTS; ++sl) {
        const PvGroup gp = c_groups[half][sl];
        g_idesc[sl] = instr_desc(128, (uint32_t)gp.n, FMT_F16, FMT_F16, 0, 1);
        g_col[sl] = tbase + gp.col0;
        g_dywp[sl] = gp.dy * tg.Wp;
        g_dx0[sl] = gp.dx0;
      }
      // MN-major, SWIZZLE_32B: LBO = 32 B (next 16-channel N group = next pixel), SBO = 256 B (next 8 keys)
      const uint64_t bd_hi = ((uint64_t)(32 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)6 << 61);

      if (n_own > 0) issue_S(0);
      for (int j = 0; j < ntiles; ++j) {
        if ((j & 1) == 0 && (j >> 1) + 1 < n_own) issue_S((j >> 1) + 1);   // scores one tile pair ahead
        const int s = j & 1;
        const uint32_t ph = (uint32_t)(j >> 1) & 1u;
        const int t = t_begin + j;
        const uint32_t tstage = smem_u32(smem + S2_T + s * TH_STAGE_BYTES);
        const uint32_t pbase = smem_u32(smem + S2_P + s * P_SLOT_BYTES);
        uint32_t g_start[TH_SLOTS];
#pragma unroll
        for (int sl = 0; sl < TH_SLOTS; ++sl)
          g_start[sl] = (tstage + sl * TH_SEG_BYTES + ((((t * TC_BN + g_dywp[sl]) & 7) + g_dx0[sl]) << 5)) >> 4;
        const uint64_t ad0 = smem_desc(pbase, (TC_BM / 8) * 128, 128);
        if (s != half) mbar_arrive_expect_tx(p_full + s, P_SLOT_BYTES);   // peer tile: arrives as a bulk copy into my slot
        { TRACE_T0(); mbar_wait(p_full + s, ph); TRACE_ADD(tr_b); }
        { TRACE_T0(); mbar_wait(t_full + s, ph); TRACE_ADD(tr_c); }
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < TC_BN / 16; ++ks) {
          const uint64_t ad = ad0 + (uint64_t)(ks * 2 * (TC_BM / 8) * 128 >> 4);
          const uint32_t acc = (j > 0 || ks > 0) ? 1u : 0u;
#pragma unroll
          for (int sl = 0; sl < TH_SLOTS; ++sl) {
            const uint64_t bd = bd_hi | (uint64_t)((g_start[sl] + ks * 32) & 0x3FFF);   // + 16 keys = 512 B
            // the four value-column groups of a k-step share the P slab through the A collector
            if (sl == 0) mma_f16_ss_a_fill(g_col[sl], ad, bd, g_idesc[sl], acc);
            else if (sl == TH_SLOTS - 1) mma_f16_ss_a_lastuse(g_col[sl], ad, bd, g_idesc[sl], acc);
            else mma_f16_ss_a_use(g_col[sl], ad, bd, g_idesc[sl], acc);
          }
        }
        mma_commit(t_empty + s);
        if (j + 2 < ntiles) mma_commit_caddr(p_free_prod[s]);   // slot s may be refilled by its producer CTA
        if (j == ntiles - 1) mma_commit(pv_last);
      }
    }
  } else if (warp == TC_THREADS / 32) {
    // ===================== P forwarder =====================
    // Pushes every P tile produced in this CTA into the peer's slot with one bulk DSMEM copy that signals the
    // peer's p_full barrier through its transaction count (async proxy end to end: no cluster-scope fences in
    // the softmax warps).
    if (elect_one()) {
      const uint32_t peer = (uint32_t)(half ^ 1);
      const uint32_t src = smem_u32(smem + S2_P + half * P_SLOT_BYTES);
      const uint32_t dst = mapa(src, peer);
      const uint32_t rbar = mapa(smem_u32(p_full + half), peer);
      for (int i = 0; i < n_own; ++i) {
        mbar_wait(p_full + half, (uint32_t)(i & 1));       // all 384 softmax threads have written + fenced
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "r"(src), "r"((uint32_t)P_SLOT_BYTES), "r"(rbar) : "memory");
      }
    }
  } else {
    // ===================== softmax / epilogue warps =====================
    const int quad = warp & 3;
    const int sub = (warp - 2) >> 2;
    const int lane = tid & 31;
    const int row = quad * 32 + lane;
    const uint32_t trow = tbase + ((uint32_t)(quad * 32) << 16);
    const size_t qidx = ((size_t)img * tg.nqt + qt) * TC_BM + row;
    const float4 t4 = __ldg(thr4 + qidx);
    const float tA = (t4.x + t4.y) * t4.z, tB = t4.w;             // T = mu * gamma - beta (dagl.py:256), mu from two column halves
    const float inv_s = 1.f / (pow2_scale(absmax[img * AMAX_STRIDE + AMAX_Q], 14) *
                               pow2_scale(absmax[img * AMAX_STRIDE + AMAX_K], 14));
    const int q = qt * TC_BM + row;
    const bool qvalid = q < g.Nq;
    // fixed softmax reference: logit of an upper bound of the row maximum, minus 12: P is stored as fp16, so the row
    // maximum is placed near 2^12 (fp16 max is 2^16) to keep the long tail of small weights (which carries real mass in
    // dense rows) out of the fp16 subnormal range
    const float ref = row_softmax_ref(__ldg(smax + qidx), __ldg(smax2 + qidx), inv_s, tA, tB, sm_scale_log2, topk);
    float l_run = 0.f;
    int cnt = 0;
    const int nwords = (g.Nk + 31) / 32;
    // my 2 chunks (keys 16*sub .. 16*sub+15) of P slot `half`: chunk stride 2048 B, 16 B per row
    const uint32_t p_local = smem_u32(smem + S2_P + half * P_SLOT_BYTES) + (2 * sub) * (TC_BM / 8) * 128 + row * 16;
    const bool want_mask = (mask_bits != nullptr) || (nnz != nullptr);
    const float neg_ref = -ref;

    for (int i = 0; i < n_own; ++i) {
      const int s = i & 1;
      const uint32_t ph = (uint32_t)(i >> 1) & 1u;
      const int t = t_begin + 2 * i + half;
      const unsigned vbits = (unsigned)(__ldg(tilemask + (size_t)img * tg.NT + t) >> (16 * sub)) & 0xffffu;
      { TRACE_T0(); mbar_wait(s_full + s, ph); TRACE_ADD(tr_a); }
      tc_fence_after();
      float sv[16];
      {
        uint32_t r0[16];
        tmem_ld16(trow + TC_S_COL0 + s * TC_BN + 16 * sub, r0);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 16; ++k) sv[k] = __uint_as_float(r0[k]);
      }
      tc_fence_before();
      mbar_arrive(s_free + s);                              // the S buffer may now be overwritten (scores two own tiles ahead)
      unsigned mk = 0u;
      uint32_t pk[8];
      float psum = 0.f;
      if (vbits == 0xffffu && !want_mask) {
        // common case: no dummy key slots in my 16 columns, no mask report requested
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
          float p[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float sc = sv[k + u] * inv_s;             // exact: inv_s is a power of two
            const float rl = fmaxf((sc - tA) + tB, 0.f);    // relu(S - mu*gamma + beta), dagl.py:256
            const float pe = ex2_approx(fmaf(logit_factor(sc, rl, topk), sm_scale_log2, neg_ref));
            p[u] = (rl != 0.f) ? pe : 0.f;                  // numerator: neighbours only (mask_b, dagl.py:257)
            if (rl == 0.f) psum += pe;                      // denominator: every key ...
          }
          pk[k / 2] = pack_half2(p[0], p[1]);
          // ... with the neighbours entering as the fp16 values the tensor core will see, so that the
          // rounding of a dominant weight cancels between numerator and denominator
          const float2 pr = __half22float2(*reinterpret_cast<const __half2*>(&pk[k / 2]));
          psum += pr.x + pr.y;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
          float p[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float sc = sv[k + u] * inv_s;
            const float rl = fmaxf((sc - tA) + tB, 0.f);
            const bool valid = (vbits >> (k + u)) & 1u;
            const float pe = valid ? ex2_approx(fmaf(logit_factor(sc, rl, topk), sm_scale_log2, neg_ref)) : 0.f;   // dummy key slots contribute nothing
            const bool nb = valid && (rl != 0.f);
            if (nb) mk |= 1u << (k + u);
            p[u] = nb ? pe : 0.f;
            if (!nb) psum += pe;
          }
          pk[k / 2] = pack_half2(p[0], p[1]);
          const float2 pr = __half22float2(*reinterpret_cast<const __half2*>(&pk[k / 2]));
          psum += pr.x + pr.y;
        }
        cnt += __popc(mk);
      }
      l_run += psum;
      // slot `half` must have been drained by the P.V of BOTH CTAs for my previous tile
      { TRACE_T0(); mbar_wait(p_free + half, (uint32_t)(i & 1) ^ 1u); TRACE_ADD(tr_b); }
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_local), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_local + 2048), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
      fence_async_smem();                                   // generic-proxy writes -> visible to the async proxy (UMMA, bulk copy)
      mbar_arrive(p_full + half);                           // local consumers: my P.V and the forwarder warp

      if (mask_bits != nullptr && qvalid && mk != 0u) {     // debug path only
        uint32_t* mrow = mask_bits + ((size_t)img * g.Nq + q) * nwords;
        unsigned rem = mk;
        while (rem) {
          const int b = __ffs((int)rem) - 1;
          rem &= rem - 1;
          const int kp = t * TC_BN + 16 * sub + b;
          const int kk = (kp / tg.Wp) * g.W + (kp % tg.Wp);
          atomicOr(mrow + (kk >> 5), 1u << (kk & 31));
        }
      }
    }

    // ---- epilogue: partial accumulator -> global ----
    if (ntiles > 0) {
      mbar_wait(pv_last, 0);
      tc_fence_after();
    }
    const float inv_t = 1.f / pow2_scale(absmax[img * AMAX_STRIDE + AMAX_THETA], 12);
    const size_t prow = ((size_t)img * nsplit + split) * g.Nq;
    float* orow = Opart + (prow + (qvalid ? q : 0)) * VD;
    int chunk = 0;
#pragma unroll 1
    for (int sl = 0; sl < TH_SLOTS; ++sl) {
      const PvGroup gp = c_groups[half][sl];
      for (int gdx = 0; gdx < gp.n / 16; ++gdx, ++chunk) {
        if (chunk % 3 != sub) continue;                     // warp-uniform: the three warps of a quadrant share the columns
        uint32_t v[16];
        tmem_ld16(trow + gp.col0 + gdx * 16, v);
        tmem_wait_ld();
        if (qvalid) {
          float4* dst = reinterpret_cast<float4*>(orow + (gp.dy * KS + gp.dx0 + gdx) * CI);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            dst[k] = make_float4(__uint_as_float(v[4 * k]) * inv_t, __uint_as_float(v[4 * k + 1]) * inv_t,
                                 __uint_as_float(v[4 * k + 2]) * inv_t, __uint_as_float(v[4 * k + 3]) * inv_t);
        }
      }
    }
    // row sums / neighbour counts of the three column groups (this CTA's own tiles only)
    float* xl = reinterpret_cast<float*>(smem + S2_RED) + (quad * 3) * 32;
    int* xc = reinterpret_cast<int*>(smem + S2_RED + 4 * 3 * 32 * 4) + (quad * 3) * 32;
    xl[sub * 32 + lane] = l_run;
    xc[sub * 32 + lane] = cnt;
    asm volatile("bar.sync %0, 96;" ::"r"(1 + quad) : "memory");
    if (sub == 0 && qvalid) {
      lpart[(((size_t)img * nsplit + split) * 2 + half) * g.Nq + q] = (xl[lane] + xl[32 + lane]) + xl[64 + lane];
      if (nnz != nullptr) atomicAdd(nnz + (size_t)img * g.Nq + q, xc[lane] + xc[32 + lane] + xc[64 + lane]);
    }
  }

#ifdef DAGL_TC_TRACE
  if (tr_cta < 1024 && (tid & 31) == 0 && warp <= 2) {
    long long* o = g_tc_trace[tr_cta] + warp * 4;        // warp 0 producer, 1 mma, 2 softmax
    o[0] = tr_a; o[1] = tr_b; o[2] = tr_c; o[3] = clock64() - tr_start;
    if (warp == 0) { g_tc_trace[tr_cta][12] = ntiles; g_tc_trace[tr_cta][14] = tr_start; }
  }
#endif
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                       // nobody leaves while the peer can still write here
  if (warp == 1) tmem_dealloc<TC_TMEM_COLS>(tbase);
}

// =============================================================================================
// v4: 4-CTA cluster per query tile.  Each CTA owns one QUARTER of the value columns (<= 208), which
// leaves room in TMEM for the whole query tile (hi + lo = 208 columns): the score MMAs take their A
// operand from TMEM (no 4 KB smem re-read per MMA: 28 instead of ~51 cycles at N = 48), and the 104 KB
// of smem the query tile used to occupy go to deeper K / theta rings.  CTA r owns the key tiles j with
// (j & 3) == r, computes S + softmax for them and forwards the fp16 P tile to the three peers with bulk
// DSMEM copies; all four run P.V for every tile on their column quarter.  Same fixed softmax reference
// (rowmax pre-pass) as v2.
//   TMEM: O [0,208) | Qh [208,312) | Ql [312,416) | S0 [416,464) | S1 [464,512)
// =============================================================================================
#ifndef V4_OPT_PREWAIT
#define V4_OPT_PREWAIT 0
#endif
#ifndef V4_OPT_WARP_ARRIVE
#define V4_OPT_WARP_ARRIVE 1
#endif
constexpr int V4_THREADS = TC_THREADS + 96;                 // + forwarder warp + score-MMA issuer warp + K loader warp
#ifndef V4_PDEPTH
#define V4_PDEPTH 2
#endif
#ifndef V4_KSTAGES
#define V4_KSTAGES 2
#endif
#ifndef V4_TSTAGES
#define V4_TSTAGES 3
#endif
constexpr int V4_KST = V4_KSTAGES, V4_TST = V4_TSTAGES, V4_TSLOTS = 4;
constexpr int V4_PD = V4_PDEPTH;                            // P tiles a producer rank may have in flight
constexpr int V4_PSLOTS = 4 * V4_PD;                        // P slots: (producer rank) + 4 * (own-tile index mod V4_PD)
constexpr int V4_TSEG_BYTES = (TC_BN + TH_SEG_PIX) * 32;    // one dy row of theta for a pair of tiles: 112 pixels = 3584 B
constexpr int V4_TSTAGE_BYTES = V4_TSLOTS * V4_TSEG_BYTES;  // one stage of rank 3 = 4 theta rows of a tile pair: 14336
#ifndef V4_PAR_ISSUE
#define V4_PAR_ISSUE 1                                      // theta rows / P forwards of a tile issued by parallel lanes
#endif
#ifndef V4_TASYM
#define V4_TASYM 0                                          // ranks 0..2 need only 2 theta rows per pair: twice the ring depth in the same smem
#endif
constexpr int V4_TST_MAX = V4_TASYM ? 2 * V4_TST : V4_TST;
constexpr int V4_O_COLS = 208, V4_QH_COL = 208, V4_QL_COL = 312, V4_S_COL0 = 416;
constexpr int S4_K = 0;
constexpr int S4_T = S4_K + V4_KST * K_TILE_BYTES;          // 79872 = 78 * 1024
constexpr int S4_P = S4_T + V4_TST * V4_TSTAGE_BYTES;       // + 43008
constexpr int S4_BAR = S4_P + V4_PSLOTS * P_SLOT_BYTES;
constexpr int S4_RED = S4_BAR + 512;
constexpr int S4_LSUM = S4_RED + 2 * 4 * 3 * 32 * 4;         // rank 0: [4 ranks][128 rows] row-sum partials of the cluster
constexpr int S4_TOTAL = S4_LSUM + 4 * TC_BM * 4;
static_assert(S4_T % 1024 == 0 && S4_P % 1024 == 0, "v4 smem alignment");
static_assert(S4_TOTAL <= 232448, "v4 smem");

// Value columns per cluster rank: two MMAs per k-step for every rank (balanced tensor-pipe load).
// Ranks 0..2 own the shift rows (2r, 2r+1) except shift (2r+1, dx 6): N = 112 + 96.  Rank 3 owns row 6 (N = 112) and the
// three left-over shifts (dy 1,3,5; dx 6) as ONE MMA with N = 48 whose N-group stride is a theta row slot (the padded
// width is a multiple of 8, so all rows of a tile share the 8-pixel phase and the slots are congruent).
struct PvGroup4 { int slot, dx0, n, col0, vertical; };     // slot: index into c_rows4[rank] of the (first) theta row
__constant__ int c_rows4[4][4] = {{0, 1, 0, 0}, {2, 3, 0, 0}, {4, 5, 0, 0}, {6, 1, 3, 5}};
__constant__ PvGroup4 c_groups4[4][2] = {
    {{0, 0, 112, 0, 0}, {1, 0, 96, 112, 0}},
    {{0, 0, 112, 0, 0}, {1, 0, 96, 112, 0}},
    {{0, 0, 112, 0, 0}, {1, 0, 96, 112, 0}},
    {{0, 0, 112, 0, 0}, {1, 6, 48, 112, 1}}};

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(V4_THREADS, 1)
attend_tc4_kernel(Geom g, TcGeom tg, const uint8_t* __restrict__ Qp, const uint8_t* __restrict__ Kp,
                  const uint8_t* __restrict__ Thp, const unsigned long long* __restrict__ tilemask,
                  const float4* __restrict__ thr4 /*per query row: mu partials (x, y), gamma, beta*/,
                  const unsigned* __restrict__ absmax, const unsigned* __restrict__ smax, const unsigned* __restrict__ smax2,
                  float sm_scale_log2, int topk,
                  int nsplit, int qt_base, float* __restrict__ Opart, float* __restrict__ lpart /*[B][nsplit][nparts][Nq]: summed over the cluster*/,
                  uint32_t* __restrict__ mask_bits, int32_t* __restrict__ nnz,
                  int t_cut /*> 0: two unequal key splits [0, t_cut) | [t_cut, NT)*/, int nparts, int head) {
  // head != 0: a 2-CTA tail launch follows as the programmatic dependent and runs concurrently: it must not start before
  // the kernels in front of this one have completed, so the trigger comes after the wait
  if (head) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  } else {
    pdl_prologue();
  }
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S4_BAR);
  uint64_t* q_ready = bars + 0;   // query tile resident in TMEM (128 arrivals)
  uint64_t* k_full = bars + 1;                  // [V4_KST] ring over OWN tiles
  uint64_t* k_empty = k_full + V4_KST;          // [V4_KST]
  uint64_t* t_full = k_empty + V4_KST;          // [V4_TST] ring over PAIRS of tiles
  uint64_t* t_empty = t_full + V4_TST_MAX;      // [V4_TST_MAX]
  uint64_t* s_full = t_empty + V4_TST_MAX;      // [2]
  uint64_t* s_free = s_full + 2;                // [2]
  uint64_t* p_full = s_free + 2;                // [V4_PSLOTS] slot r + 4*d is written by cluster rank r (into all four CTAs)
  uint64_t* p_free = p_full + V4_PSLOTS;        // [V4_PSLOTS] only the slots of `rank` are waited on here: 4 commit arrivals (every CTA's P.V)
  uint64_t* pv_last = p_free + V4_PSLOTS;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_last + 1);
  static_assert((7 + 2 * V4_KST + 2 * V4_TST_MAX + 2 * V4_PSLOTS) * 8 <= 512, "v4 barrier area");

  const int warp = warp_id_uniform();
  const int tid = threadIdx.x;
  const int img = blockIdx.z, split = blockIdx.y;
  const int qt = qt_base + (blockIdx.x >> 2);
  const int rank = (int)cluster_ctarank();                  // == blockIdx.x & 3
  const int t_begin = t_cut > 0 ? (split ? t_cut : 0) : (int)(((long long)split * tg.NT) / nsplit);
  const int t_end = t_cut > 0 ? (split ? tg.NT : t_cut) : (int)(((long long)(split + 1) * tg.NT) / nsplit);
  const int ntiles = t_end - t_begin;
  const int n_own = (ntiles - rank + 3) / 4;                // local tiles j with (j & 3) == rank
  const int nslots = rank == 3 ? 4 : 2;                     // theta rows this rank needs
  // theta ring of this rank: the same bytes hold twice as many (half-size) stages for the ranks that need two rows
  const int tst = (V4_TASYM && rank != 3) ? 2 * V4_TST : V4_TST;
  const int tstage_bytes = (V4_TASYM && rank != 3) ? V4_TSTAGE_BYTES / 2 : V4_TSTAGE_BYTES;
#ifdef DAGL_TC_TRACE
  long long tr_a = 0, tr_b = 0, tr_c = 0;
  const bool tl_on = (blockIdx.x >> 2) == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  const long long tr_start = clock64();
  const int tr_cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
#endif

  if (tid == 0) {
    mbar_init(q_ready, 12);
    for (int i = 0; i < V4_KST; ++i) { mbar_init(k_full + i, 1); mbar_init(k_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(s_full + i, 1); mbar_init(s_free + i, V4_OPT_WARP_ARRIVE ? 12 : 384); }
    for (int i = 0; i < V4_TST_MAX; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, 1); }
    for (int i = 0; i < V4_PSLOTS; ++i) { mbar_init(p_full + i, (i & 3) == rank ? (V4_OPT_WARP_ARRIVE ? 12 : 384) : 1); mbar_init(p_free + i, 4); }
    mbar_init(pv_last, 1);
    mbar_init_fence();
  }
  if (warp == 1) tmem_alloc<TC_TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;
#ifdef DAGL_TC_TRACE
  const long long tl_base = clock64();      // common time origin of the cluster's timelines: right after the cluster barrier
#endif

  if (warp == 0) {
    // ===================== TMA producer =====================
#if V4_PAR_ISSUE
    {
      // one lane per theta row: the 2 (rank 3: 4) copies of a tile pair are issued together
      const int lane = tid & 31;
      const uint8_t* thp = Thp + (size_t)img * tg.NP * 32;
      const int nhalf = (ntiles + 1) >> 1;
      for (int h = 0; h < nhalf; ++h) {
        const int s = h % tst;
        if (lane == 0) {
          mbar_wait(t_empty + s, ((uint32_t)(h / tst) & 1u) ^ 1u);
          mbar_arrive_expect_tx(t_full + s, (uint32_t)nslots * V4_TSEG_BYTES);
        }
        __syncwarp();
        const int k0 = (t_begin + 2 * h) * TC_BN;            // multiple of 8, and so is Wp
        if (lane < nslots)
          bulk_g2s(smem + S4_T + s * tstage_bytes + lane * V4_TSEG_BYTES, thp + (size_t)(k0 + c_rows4[rank][lane] * tg.Wp) * 32,
                   V4_TSEG_BYTES, t_full + s);
        __syncwarp();
      }
    }
#else
    if (elect_one()) {
      const uint8_t* thp = Thp + (size_t)img * tg.NP * 32;
      // theta rows of the tile pair (2h, 2h+1): 96 consecutive key slots + the 64-pixel window, ONE copy per row
      const int nhalf = (ntiles + 1) >> 1;
      for (int h = 0; h < nhalf; ++h) {
        const int s = h % tst;
        { TRACE_T0(); mbar_wait(t_empty + s, ((uint32_t)(h / tst) & 1u) ^ 1u); TRACE_ADD(tr_b); }
#ifdef DAGL_TC_TRACE
        if (g_tc_dbg_mode & 8) {                             // timing experiment: one small segment per pair
          mbar_arrive_expect_tx(t_full + s, 1024);
          bulk_g2s(smem + S4_T + s * tstage_bytes, thp, 1024, t_full + s);
          continue;
        }
#endif
        mbar_arrive_expect_tx(t_full + s, (uint32_t)nslots * V4_TSEG_BYTES);
        const int k0 = (t_begin + 2 * h) * TC_BN;            // multiple of 8, and so is Wp
        for (int sl = 0; sl < nslots; ++sl)
          bulk_g2s(smem + S4_T + s * tstage_bytes + sl * V4_TSEG_BYTES, thp + (size_t)(k0 + c_rows4[rank][sl] * tg.Wp) * 32,
                   V4_TSEG_BYTES, t_full + s);
        TL(h >> 1, 20);
      }
    }
#endif
  } else if (warp == TC_THREADS / 32 + 2) {
    // ===================== K loader: own key tiles (local tile 4i + rank) =====================
    if (elect_one()) {
      const uint8_t* ksrc = Kp + ((size_t)img * tg.NT + t_begin + rank) * K_TILE_BYTES;
      for (int i = 0; i < n_own; ++i, ksrc += 4 * (size_t)K_TILE_BYTES) {
        const int s = i % V4_KST;
        { TRACE_T0(); mbar_wait(k_empty + s, ((uint32_t)(i / V4_KST) & 1u) ^ 1u); TRACE_ADD(tr_a); }
#ifdef DAGL_TC_TRACE
        if (g_tc_dbg_mode & 16) {                            // timing experiment: 1 KB instead of the whole K tile
          mbar_arrive_expect_tx(k_full + s, 1024);
          bulk_g2s(smem + S4_K + s * K_TILE_BYTES, ksrc, 1024, k_full + s);
          continue;
        }
#endif
        mbar_arrive_expect_tx(k_full + s, K_TILE_BYTES);
        bulk_g2s(smem + S4_K + s * K_TILE_BYTES, ksrc, K_TILE_BYTES, k_full + s);
        TL(i, 19);
      }
    }
  } else if (warp == 1) {
    // ===================== P.V MMA issuer =====================
    if (elect_one()) {
      uint32_t g_idesc[2], g_col[2], g_off[2];
      uint64_t g_bd[2];
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) {
        const PvGroup4 gp = c_groups4[rank][sl];
        g_idesc[sl] = instr_desc(128, (uint32_t)gp.n, FMT_F16, FMT_F16, 0, 1);
        g_col[sl] = tbase + gp.col0;
        g_off[sl] = (uint32_t)(gp.slot * V4_TSEG_BYTES + gp.dx0 * 32);
        // MN-major SWIZZLE_32B: SBO = 256 B (next 8 keys); LBO = stride between 16-channel N groups:
        // one pixel (next dx) for a horizontal run, one theta row slot (next owned dy) for the vertical run
        const uint32_t lbo = gp.vertical ? (uint32_t)V4_TSEG_BYTES : 32u;
        g_bd[sl] = ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
      }
      // The barrier work of tile j+1 (theta rows of a new pair, P tile) is done between the k-steps of tile j, so the
      // tensor pipe has queued work while this thread sits in try_wait / expect_tx; the waits normally pass at once.
      // The loop is unrolled over the producer rank u = j & 3 so that all slot / barrier addresses are immediates.
      uint32_t p_free_prod[4];                                               // p_free[r] (parity 0) in the producer CTA r
#pragma unroll
      for (int r = 0; r < 4; ++r) p_free_prod[r] = mapa(smem_u32(p_free + r), (uint32_t)r);
      const uint32_t t_base = smem_u32(smem + S4_T), p_base = smem_u32(smem + S4_P);
      int w_ts = -1; uint32_t w_ph = 1u;                                     // theta stage / phase of the pair last waited for
      int i_ts = -1;                                                         // theta stage of the pair being issued
#define V4_PREWAIT(JN, UN, PN)                                                                            \
      do {                                                                                                \
        if (((UN) & 1) == 0) {                                                                            \
          if (++w_ts == tst) w_ts = 0;                                                                    \
          if (w_ts == 0) w_ph ^= 1u;                                                                      \
          { TRACE_T0(); mbar_wait(t_full + w_ts, w_ph); TRACE_ADD(tr_c); }                                \
          if ((UN) == 0) TL(PN, 16);                                                                      \
        }                                                                                                 \
        const int ps_ = (UN) + 4 * ((PN) % V4_PD);                                                        \
        if ((UN) != rank) mbar_arrive_expect_tx(p_full + ps_, V4_FWD_BYTES);                              \
        { TRACE_T0(); mbar_wait(p_full + ps_, (uint32_t)((PN) / V4_PD) & 1u); TRACE_ADD(tr_b); }         \
        TL(PN, 8 + 2 * (UN));                                                                             \
        if (!skip_fence) tc_fence_after();                                                                \
      } while (0)
#ifdef DAGL_TC_TRACE
      const uint32_t V4_FWD_BYTES = (g_tc_dbg_mode & 4) ? 16u : (uint32_t)P_SLOT_BYTES;
      const bool skip_pv = (g_tc_dbg_mode & 2) != 0;
      const bool skip_fence = (g_tc_dbg_mode & 64) != 0;
#else
      constexpr bool skip_fence = false;
      constexpr uint32_t V4_FWD_BYTES = P_SLOT_BYTES;
      constexpr bool skip_pv = false;
#endif
      if (ntiles > 0) V4_PREWAIT(0, 0, 0);
      for (int p = 0; 4 * p < ntiles; ++p) {
        const uint32_t pslot0 = p_base + (p % V4_PD) * (4 * P_SLOT_BYTES);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = 4 * p + u;                                           // producer rank u, P slot u + 4 * (p & 1)
          if (j >= ntiles) break;
          if ((u & 1) == 0 && ++i_ts == tst) i_ts = 0;
          const uint32_t tile0 = t_base + i_ts * tstage_bytes + (u & 1) * (TC_BN * 32);   // 2nd tile of the pair: +48 pixels
          const uint32_t g_start0 = (tile0 + g_off[0]) >> 4, g_start1 = (tile0 + g_off[1]) >> 4;
          const uint64_t ad0 = smem_desc(pslot0 + u * P_SLOT_BYTES, (TC_BM / 8) * 128, 128);
          const uint32_t acc0 = j > 0 ? 1u : 0u;
          if (!skip_pv) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t ad = ad0 + (uint64_t)(ks * 2 * (TC_BM / 8) * 128 >> 4);
              const uint64_t b0 = g_bd[0] | (uint64_t)((g_start0 + ks * 32) & 0x3FFF);    // 16 keys = 512 B further per k-step
              const uint64_t b1 = g_bd[1] | (uint64_t)((g_start1 + ks * 32) & 0x3FFF);
              mma_f16_ss_a_fill(g_col[0], ad, b0, g_idesc[0], ks > 0 ? 1u : acc0);         // the two groups share the P slab (A collector)
              mma_f16_ss_a_lastuse(g_col[1], ad, b1, g_idesc[1], ks > 0 ? 1u : acc0);
            }
          }
#if V4_OPT_PREWAIT
          if (j + 1 < ntiles) {
            if (u == 3) V4_PREWAIT(j + 1, 0, p + 1); else V4_PREWAIT(j + 1, u + 1, p);
          }
#endif
          if (!skip_pv) {
            const uint64_t ad = ad0 + (uint64_t)(2 * 2 * (TC_BM / 8) * 128 >> 4);
            const uint64_t b0 = g_bd[0] | (uint64_t)((g_start0 + 2 * 32) & 0x3FFF);
            const uint64_t b1 = g_bd[1] | (uint64_t)((g_start1 + 2 * 32) & 0x3FFF);
            mma_f16_ss_a_fill(g_col[0], ad, b0, g_idesc[0], 1u);
            mma_f16_ss_a_lastuse(g_col[1], ad, b1, g_idesc[1], 1u);
          }
#ifdef DAGL_TC_TRACE
          if (g_tc_dbg_mode & 32) {                                              // timing experiment: plain arrives instead of tcgen05.commit
            if (j + 4 * V4_PD < ntiles)
              asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(p_free_prod[u] + 32 * (p % V4_PD)) : "memory");
            if ((u & 1) == 1 || j == ntiles - 1) mbar_arrive(t_empty + i_ts);
          } else {
#endif
          if (j + 4 * V4_PD < ntiles) mma_commit_caddr(p_free_prod[u] + 32 * (p % V4_PD));   // slot may be refilled by its producer CTA (+4 barriers)
          if ((u & 1) == 1 || j == ntiles - 1) mma_commit(t_empty + i_ts);
#ifdef DAGL_TC_TRACE
          }
#endif
          if (j == ntiles - 1) mma_commit(pv_last);
#if !V4_OPT_PREWAIT
          if (j + 1 < ntiles) {
            if (u == 3) V4_PREWAIT(j + 1, 0, p + 1); else V4_PREWAIT(j + 1, u + 1, p);
          }
#endif
          TL(p, 9 + 2 * u);
        }
      }
#undef V4_PREWAIT
    }
  } else if (warp == TC_THREADS / 32 + 1) {
    // ===================== score MMA issuer (own tiles; A operands Qh, Ql from TMEM) =====================
    if (elect_one()) {
      constexpr uint32_t idS = instr_desc(128, TC_BN, FMT_F16, FMT_F16, 0, 0);
      const uint32_t qh = tbase + V4_QH_COL, ql = tbase + V4_QL_COL;
      mbar_wait(q_ready, 0);
      tc_fence_after();
      for (int i = 0; i < n_own; ++i) {
        const int s = i & 1, ks_ = i % V4_KST;
        { TRACE_T0(); mbar_wait(k_full + ks_, (uint32_t)(i / V4_KST) & 1u); TRACE_ADD(tr_a); }
        TL(i, 5);
        { TRACE_T0(); mbar_wait(s_free + s, ((uint32_t)(i >> 1) & 1u) ^ 1u); TRACE_ADD(tr_b); }
        TL(i, 6);
        tc_fence_after();
        const uint32_t k_hi = smem_u32(smem + S4_K + ks_ * K_TILE_BYTES), k_lo = k_hi + K_HALF_BYTES;
        const uint64_t dk_hi = smem_desc(k_hi, (TC_BN / 8) * 128, 128);
        const uint64_t dk_lo = smem_desc(k_lo, (TC_BN / 8) * 128, 128);
        const uint32_t d = tbase + V4_S_COL0 + s * TC_BN;
#ifdef DAGL_TC_TRACE
        if (!(g_tc_dbg_mode & 1))
#endif
        {
#pragma unroll
          for (int ks = 0; ks < TC_KSTEPS; ++ks) {
            const uint64_t ko = (uint64_t)(ks * 2 * (TC_BN / 8) * 128 >> 4);
            mma_f16_ts(d, ql + ks * 8, dk_hi + ko, idS, ks > 0);              // Ql.Kh first (small terms)
          }
#pragma unroll
          for (int ks = 0; ks < TC_KSTEPS; ++ks) {
            const uint64_t ko = (uint64_t)(ks * 2 * (TC_BN / 8) * 128 >> 4);
            mma_f16_ts(d, qh + ks * 8, dk_lo + ko, idS, 1);                   // Qh.Kl
            mma_f16_ts(d, qh + ks * 8, dk_hi + ko, idS, 1);                   // Qh.Kh
          }
        }
        mma_commit(s_full + s);
        mma_commit(k_empty + ks_);
        TL(i, 7);
      }
    }
  } else if (warp == TC_THREADS / 32) {
    // ===================== P forwarder: one bulk DSMEM copy per peer =====================
    // A thread needs ~100 cycles per dependent bulk-copy issue (tools/bulk_copy_probe.cu): lanes 0..2 each own one peer, so
    // the three forwards of a tile leave together instead of one after the other.
#if V4_PAR_ISSUE
    {
      const int lane = tid & 31;
      const uint32_t src0 = smem_u32(smem + S4_P + rank * P_SLOT_BYTES);
      const uint32_t peer = (uint32_t)((rank + 1 + (lane < 3 ? lane : 0)) & 3);
      const uint32_t dst = mapa(src0, peer), rbar = mapa(smem_u32(p_full + rank), peer);
      for (int i = 0; i < n_own; ++i) {
        const uint32_t par = (uint32_t)(i % V4_PD);                        // slot rank + 4*par
        mbar_wait(p_full + rank + 4 * par, (uint32_t)(i / V4_PD) & 1u);
        if (lane < 3)
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst + par * 4 * P_SLOT_BYTES), "r"(src0 + par * 4 * P_SLOT_BYTES), "r"((uint32_t)P_SLOT_BYTES),
                         "r"(rbar + par * 32) : "memory");
        __syncwarp();
      }
    }
#else
    if (elect_one()) {
      const uint32_t src0 = smem_u32(smem + S4_P + rank * P_SLOT_BYTES);
      uint32_t dst[3], rbar[3];
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const uint32_t peer = (uint32_t)((rank + 1 + u) & 3);
        dst[u] = mapa(src0, peer);
        rbar[u] = mapa(smem_u32(p_full + rank), peer);
      }
      for (int i = 0; i < n_own; ++i) {
        const uint32_t par = (uint32_t)(i % V4_PD);                        // slot rank + 4*par
        mbar_wait(p_full + rank + 4 * par, (uint32_t)(i / V4_PD) & 1u);
        TL(i, 17);
#ifdef DAGL_TC_TRACE
        const uint32_t fwd_bytes = (g_tc_dbg_mode & 4) ? 16u : (uint32_t)P_SLOT_BYTES;   // timing experiment: 16-byte forwards
#else
        const uint32_t fwd_bytes = (uint32_t)P_SLOT_BYTES;
#endif
#pragma unroll
        for (int u = 0; u < 3; ++u)
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst[u] + par * 4 * P_SLOT_BYTES), "r"(src0 + par * 4 * P_SLOT_BYTES), "r"(fwd_bytes),
                         "r"(rbar[u] + par * 32) : "memory");
        TL(i, 18);
      }
    }
#endif
  } else {
    // ===================== softmax / epilogue warps =====================
    const int quad = warp & 3;
    const int sub = (warp - 2) >> 2;
    const int lane = tid & 31;
    const int row = quad * 32 + lane;
    const uint32_t trow = tbase + ((uint32_t)(quad * 32) << 16);
    const size_t qidx = ((size_t)img * tg.nqt + qt) * TC_BM + row;
    // ---- query tile -> TMEM (A operand of the score MMAs): lane = row, column j = elements (2j, 2j+1) ----
    // All twelve warps take part (the three warps of a lane quadrant split the 26 (part, k-step) slabs) and every thread
    // issues its loads back to back: with one warp per quadrant and one dependent load -> store pair at a time this
    // prologue took ~13 us per cluster, a third of a chop-leaf cluster's whole life.
    {
      const uint8_t* qsrc = Qp + ((size_t)img * tg.nqt + qt) * Q_TILE_BYTES + row * 16;
      constexpr int NSLAB = 2 * TC_KSTEPS;                 // slab = (part, k-step): 8 TMEM columns
      constexpr int PER = (NSLAB + 2) / 3;                 // 9 slabs per warp of the quadrant (the last warp: 8)
      uint4 c0[PER], c1[PER];
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int slab = sub + 3 * i;
        if (slab < NSLAB) {
          const int part = slab / TC_KSTEPS, ks = slab % TC_KSTEPS;
          const uint8_t* src = qsrc + part * Q_HALF_BYTES;
          c0[i] = __ldg(reinterpret_cast<const uint4*>(src + (2 * ks) * (TC_BM / 8) * 128));
          c1[i] = __ldg(reinterpret_cast<const uint4*>(src + (2 * ks + 1) * (TC_BM / 8) * 128));
        }
      }
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int slab = sub + 3 * i;
        if (slab < NSLAB) {
          const int part = slab / TC_KSTEPS, ks = slab % TC_KSTEPS;
          const uint32_t v[8] = {c0[i].x, c0[i].y, c0[i].z, c0[i].w, c1[i].x, c1[i].y, c1[i].z, c1[i].w};
          tmem_st8(trow + (part ? V4_QL_COL : V4_QH_COL) + ks * 8, v);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_ready);                 // one arrival per warp (12)
    }
    const float4 t4 = __ldg(thr4 + qidx);
    const float tA = (t4.x + t4.y) * t4.z, tB = t4.w;             // T = mu * gamma - beta (dagl.py:256), mu from two column halves
    const float inv_s = 1.f / (pow2_scale(absmax[img * AMAX_STRIDE + AMAX_Q], 14) *
                               pow2_scale(absmax[img * AMAX_STRIDE + AMAX_K], 14));
    const int q = qt * TC_BM + row;
    const bool qvalid = q < g.Nq;
    const float ref = row_softmax_ref(__ldg(smax + qidx), __ldg(smax2 + qidx), inv_s, tA, tB, sm_scale_log2, topk);
    float l_run = 0.f;
    int cnt = 0;
    const int nwords = (g.Nk + 31) / 32;
    const uint32_t p_local0 = smem_u32(smem + S4_P + rank * P_SLOT_BYTES) + (2 * sub) * (TC_BM / 8) * 128 + row * 16;
    const bool want_mask = (mask_bits != nullptr) || (nnz != nullptr);
    const float neg_ref = -ref;

    unsigned long long vmask_next = n_own > 0 ? __ldg(tilemask + (size_t)img * tg.NT + t_begin + rank) : 0ull;
    for (int i = 0; i < n_own; ++i) {
      const int s = i & 1;
      const uint32_t ph = (uint32_t)(i >> 1) & 1u;
      const int t = t_begin + 4 * i + rank;
      const unsigned vbits = (unsigned)(vmask_next >> (16 * sub)) & 0xffffu;
      if (i + 1 < n_own) vmask_next = __ldg(tilemask + (size_t)img * tg.NT + t + 4);     // prefetched one tile ahead
      { TRACE_T0(); mbar_wait(s_full + s, ph); TRACE_ADD(tr_a); }
      if (warp == 2 && lane == 0) TL(i, 0);
      tc_fence_after();
      float sv[16];
      {
        uint32_t r0[16];
        tmem_ld16(trow + V4_S_COL0 + s * TC_BN + 16 * sub, r0);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 16; ++k) sv[k] = __uint_as_float(r0[k]);
      }
      tc_fence_before();
#if V4_OPT_WARP_ARRIVE
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free + s);               // one arrival per warp (12 per tile)
#else
      mbar_arrive(s_free + s);
#endif
      if (warp == 2 && lane == 0) TL(i, 1);
      unsigned mk = 0u;
      uint32_t pk[8];
      float psum = 0.f;
      if (vbits == 0xffffu && !want_mask) {
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
          float p[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float sc = sv[k + u] * inv_s;
            const float rl = fmaxf((sc - tA) + tB, 0.f);
            const float pe = ex2_approx(fmaf(logit_factor(sc, rl, topk), sm_scale_log2, neg_ref));
            p[u] = (rl != 0.f) ? pe : 0.f;
            if (rl == 0.f) psum += pe;
          }
          pk[k / 2] = pack_half2(p[0], p[1]);
          const float2 pr = __half22float2(*reinterpret_cast<const __half2*>(&pk[k / 2]));
          psum += pr.x + pr.y;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
          float p[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float sc = sv[k + u] * inv_s;
            const float rl = fmaxf((sc - tA) + tB, 0.f);
            const bool valid = (vbits >> (k + u)) & 1u;
            const float pe = valid ? ex2_approx(fmaf(logit_factor(sc, rl, topk), sm_scale_log2, neg_ref)) : 0.f;
            const bool nb = valid && (rl != 0.f);
            if (nb) mk |= 1u << (k + u);
            p[u] = nb ? pe : 0.f;
            if (!nb) psum += pe;
          }
          pk[k / 2] = pack_half2(p[0], p[1]);
          const float2 pr = __half22float2(*reinterpret_cast<const __half2*>(&pk[k / 2]));
          psum += pr.x + pr.y;
        }
        cnt += __popc(mk);
      }
      l_run += psum;
      const int par = i % V4_PD;                            // my slot for this tile: rank + 4*par (last used by own tile i - V4_PD)
      const uint32_t p_local = p_local0 + par * 4 * P_SLOT_BYTES;
      if (warp == 2 && lane == 0) TL(i, 2);
      { TRACE_T0(); mbar_wait(p_free + rank + 4 * par, ((uint32_t)(i / V4_PD) & 1u) ^ 1u); TRACE_ADD(tr_b); }
      if (warp == 2 && lane == 0) TL(i, 3);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_local), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_local + 2048), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
      fence_async_smem();
#if V4_OPT_WARP_ARRIVE
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full + rank + 4 * par);  // one arrival per warp
#else
      mbar_arrive(p_full + rank + 4 * par);
#endif
      if (warp == 2 && lane == 0) TL(i, 4);

      if (mask_bits != nullptr && qvalid && mk != 0u) {     // debug path only
        uint32_t* mrow = mask_bits + ((size_t)img * g.Nq + q) * nwords;
        unsigned rem = mk;
        while (rem) {
          const int b = __ffs((int)rem) - 1;
          rem &= rem - 1;
          const int kp = t * TC_BN + 16 * sub + b;
          const int kk = (kp / tg.Wp) * g.W + (kp % tg.Wp);
          atomicOr(mrow + (kk >> 5), 1u << (kk & 31));
        }
      }
    }

    // ---- epilogue: partial accumulator -> global ----
    if (ntiles > 0) {
      mbar_wait(pv_last, 0);
      tc_fence_after();
    }
    const float inv_t = 1.f / pow2_scale(absmax[img * AMAX_STRIDE + AMAX_THETA], 12);
    const size_t prow = ((size_t)img * nsplit + split) * g.Nq;
    float* orow = Opart + (prow + (qvalid ? q : 0)) * VD;
    int chunk = 0;
#pragma unroll 1
    for (int sl = 0; sl < 2; ++sl) {
      const PvGroup4 gp = c_groups4[rank][sl];
      for (int gdx = 0; gdx < gp.n / 16; ++gdx, ++chunk) {
        if (chunk % 3 != sub) continue;
        uint32_t v[16];
        tmem_ld16(trow + gp.col0 + gdx * 16, v);
        tmem_wait_ld();
        if (qvalid) {
          const int shift = gp.vertical ? c_rows4[rank][gp.slot + gdx] * KS + gp.dx0 : c_rows4[rank][gp.slot] * KS + gp.dx0 + gdx;
          float4* dst = reinterpret_cast<float4*>(orow + shift * CI);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            dst[k] = make_float4(__uint_as_float(v[4 * k]) * inv_t, __uint_as_float(v[4 * k + 1]) * inv_t,
                                 __uint_as_float(v[4 * k + 2]) * inv_t, __uint_as_float(v[4 * k + 3]) * inv_t);
        }
      }
    }
    float* xl = reinterpret_cast<float*>(smem + S4_RED) + (quad * 3) * 32;
    int* xc = reinterpret_cast<int*>(smem + S4_RED + 4 * 3 * 32 * 4) + (quad * 3) * 32;
    xl[sub * 32 + lane] = l_run;
    xc[sub * 32 + lane] = cnt;
    asm volatile("bar.sync %0, 96;" ::"r"(1 + quad) : "memory");
    if (sub == 0) {
      // row-sum partial of this rank -> rank 0's smem (DSMEM store, made visible by the cluster barrier below); rank 0 adds
      // the four in a fixed order, so the fold sees ONE partial per key split
      const uint32_t dst = mapa(smem_u32(reinterpret_cast<float*>(smem + S4_LSUM) + rank * TC_BM + row), 0u);
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(dst), "f"((xl[lane] + xl[32 + lane]) + xl[64 + lane]) : "memory");
      if (qvalid && nnz != nullptr) atomicAdd(nnz + (size_t)img * g.Nq + q, xc[lane] + xc[32 + lane] + xc[64 + lane]);
    }
  }

#ifdef DAGL_TC_TRACE
  if (tr_cta < 1024 && (tid & 31) == 0 && warp <= 2) {
    long long* o = g_tc_trace[tr_cta] + warp * 4;
    o[0] = tr_a; o[1] = tr_b; o[2] = tr_c; o[3] = clock64() - tr_start;
    if (warp == 0) { g_tc_trace[tr_cta][12] = ntiles; g_tc_trace[tr_cta][14] = tr_start; }
  }
  if (tr_cta < 1024 && (tid & 31) == 0 && warp == TC_THREADS / 32 + 1) {   // score issuer
    long long* o = g_tc_trace[tr_cta] + 16;
    o[0] = tr_a; o[1] = tr_b; o[2] = tr_c; o[3] = clock64() - tr_start;
  }
#endif
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (rank == 0 && tid < TC_BM && qt * TC_BM + tid < g.Nq) {
    const float* ls = reinterpret_cast<const float*>(smem + S4_LSUM) + tid;
    float* lp = lpart + ((size_t)img * nsplit + split) * nparts * g.Nq + qt * TC_BM + tid;
    lp[0] = ((ls[0] + ls[TC_BM]) + ls[2 * TC_BM]) + ls[3 * TC_BM];
    if (nparts > 1) lp[g.Nq] = 0.f;                         // layout shared with a 2-CTA tail launch (two partials per split)
  }
  if (warp == 1) tmem_dealloc<TC_TMEM_COLS>(tbase);
}

// coef[b][s][q] = 1 / sum_{s,h} l   (fixed-reference partials merge by plain sums)
__global__ void merge_coef_fixed_kernel(int B, int Nq, int nsplit, int nparts, int q_begin, int q_end,
                                        const float* __restrict__ lpart, float* __restrict__ coef) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Nq) return;
  const int img = i / Nq, q = i % Nq;
  if (q < q_begin || q >= q_end) return;
  float L = 0.f;
  for (int s = 0; s < nsplit; ++s)
    for (int h = 0; h < nparts; ++h) L += lpart[(((size_t)img * nsplit + s) * nparts + h) * Nq + q];
  // L > 0 whenever the row has a valid key (the row maximum itself contributes ~2^12); guard the degenerate case so that
  // it can never turn into inf * 0 = NaN in the fold
  const float inv = L > 0.f ? 1.f / L : 0.f;
  for (int s = 0; s < nsplit; ++s) coef[((size_t)img * nsplit + s) * Nq + q] = inv;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static double split_cost() {
  static const double c = [] { const char* e = getenv("DAGL_SPLIT_COST"); return e ? atof(e) : 4.0; }();
  return c;
}
static int tc_splits(const Geom& g, const TcGeom& tg, int nqt_range, int csize = 2) {
  const long long base = (long long)g.B * nqt_range * csize;
  const int sms = csize == 4 ? 132 : 148;          // 4-CTA clusters cannot use every SM (GPC sizes 16/18/20)
  const int smax = tg.NT < 32 ? tg.NT : 32;
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= smax; ++s) {
    const double waves = (double)((base * s + sms - 1) / sms);
    // time ~ waves * (tiles per CTA + fixed per-CTA overhead of ~6 tiles), plus the fold's serial loop over the splits
    // (measured at 64^2: ~2.5 us per split = ~5 tile times; DAGL_SPLIT_COST overrides for experiments)
    const double cost = waves * ((double)tg.NT / s + 6.0) + split_cost() * s;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  return best;
}

constexpr int TOPK_LISTS = 2;                                // key splits of the top-k selection pass
struct TcWs {
  size_t absmax, Qp, Kp, Thp, tilemask, thr4, smax, colsum, kbar, lists, Opart, lpart, coef, total;
  int nsplit;
};

// `nqt_range`: number of 128-query tiles the launch will cover (0: all).  The key-split factor (and with it the size of
// the partial-result buffers, which come last so that every other offset is independent of it) is the larger of what the
// 2-CTA and the 4-CTA kernel would choose for that range.
static TcWs tc_ws(const Geom& g, const TcGeom& tg, int nqt_range = 0) {
  TcWs w;
  const int range = (nqt_range > 0 && nqt_range < tg.nqt) ? nqt_range : tg.nqt;
  {
    const int s2 = tc_splits(g, tg, range, 2), s4 = tc_splits(g, tg, range, 4);
    w.nsplit = s2 > s4 ? s2 : s4;
  }
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off += align_up(b); return o; };
  w.absmax = take((size_t)g.B * AMAX_STRIDE * sizeof(unsigned));
  w.Qp = take((size_t)g.B * tg.nqt * Q_TILE_BYTES);
  w.Kp = take((size_t)g.B * tg.NT * K_TILE_BYTES);
  w.Thp = take((size_t)g.B * tg.NP * 32);
  w.tilemask = take((size_t)g.B * tg.NT * 8);
  w.thr4 = take((size_t)g.B * tg.nqt * TC_BM * 16);
  w.smax = take(2 * (size_t)g.B * tg.nqt * TC_BM * 4);      // pre-pass maxima | exact maxima of refined rows
  w.colsum = take((size_t)g.B * tg.NT * ED * 4);
  w.kbar = take((size_t)g.B * ED * 4);
  w.lists = take((size_t)g.B * tg.nqt * TC_BM * TOPK_LISTS * TOPK_MAX * 4);      // legacy top-k variant only
  const size_t rows = (size_t)g.B * w.nsplit * g.Nq;
  w.Opart = take(rows * VD * 4);
  w.lpart = take(4 * rows * 4);                 // one row-sum partial per cluster rank
  w.coef = take(rows * 4);
  w.total = off;
  return w;
}

size_t attend_tc_workspace_bytes(const Geom& g, int nqt_range) { return tc_ws(g, tc_geom(g), nqt_range).total; }

// operand buffers inside the graph kernel's workspace that the prologue kernels write directly (none of these offsets
// depends on the key-split factor)
AttendBuffers attend_tc_buffers(const Geom& g, void* attend_ws) {
  const TcGeom tg = tc_geom(g);
  const TcWs w = tc_ws(g, tg);
  char* base = static_cast<char*>(attend_ws);
  AttendBuffers b;
  b.ktiles = reinterpret_cast<uint8_t*>(base + w.Kp);
  b.colsum = reinterpret_cast<float*>(base + w.colsum);
  b.qtiles = reinterpret_cast<uint8_t*>(base + w.Qp);
  b.thp = reinterpret_cast<uint8_t*>(base + w.Thp);
  b.np_t = tg.NP;
  b.tilemask = reinterpret_cast<unsigned long long*>(base + w.tilemask);
  b.thr4 = reinterpret_cast<float*>(base + w.thr4);
  b.kbar = reinterpret_cast<float*>(base + w.kbar);
  return b;
}


int launch_attend_tc(const Geom& g, const AttendArgs& a, const unsigned* absmax_in, int variant, cudaStream_t st) {
  const TcGeom tg = tc_geom(g);
  const int qt_begin = a.qt_begin > 0 ? a.qt_begin : 0;
  const int qt_end = (a.qt_end > 0 && a.qt_end < tg.nqt) ? a.qt_end : tg.nqt;
  if (qt_begin >= qt_end) {
    call_state().err = "empty query-tile range";
    return -1;
  }
  if (variant != 2 && variant != 4) {
    call_state().err = "unknown tensor-core variant";
    return -1;
  }
  const bool ranged = (qt_begin != 0 || qt_end != tg.nqt);
  TcWs w = tc_ws(g, tg, ranged ? qt_end - qt_begin : 0);
  {
    const int want = tc_splits(g, tg, qt_end - qt_begin, variant == 4 ? 4 : 2);
    if (want < w.nsplit) w.nsplit = want;           // never more splits than the workspace was sized for
  }
  if (a.ws_bytes < w.total) {
    call_state().err = "attend (tc) workspace too small";
    return -3;
  }
  char* base = static_cast<char*>(a.ws);
  unsigned* absmax = reinterpret_cast<unsigned*>(base + w.absmax);
  uint8_t* Qp = reinterpret_cast<uint8_t*>(base + w.Qp);
  uint8_t* Kp = reinterpret_cast<uint8_t*>(base + w.Kp);
  uint8_t* Thp = reinterpret_cast<uint8_t*>(base + w.Thp);
  unsigned long long* tilemask = reinterpret_cast<unsigned long long*>(base + w.tilemask);
  float4* thr4 = reinterpret_cast<float4*>(base + w.thr4);
  float* Opart = reinterpret_cast<float*>(base + w.Opart);
  float* lpart = reinterpret_cast<float*>(base + w.lpart);
  float* coef = reinterpret_cast<float*>(base + w.coef);

  if (absmax_in != nullptr) {
    absmax = const_cast<unsigned*>(absmax_in);      // filled by the prologue kernels of the same forward
  } else {
    DAGL_CUDA_OK(cudaMemsetAsync(absmax, 0, (size_t)g.B * AMAX_STRIDE * sizeof(unsigned), st));
    absmax_kernel<<<dim3(64, g.B), 256, 0, st>>>(a.Q, (size_t)g.Nq * ED, absmax, 0);
    DAGL_LAUNCH_CHECK();
    absmax_kernel<<<dim3(256, g.B), 256, 0, st>>>(a.K, (size_t)g.Nk * ED, absmax, 1);
    DAGL_LAUNCH_CHECK();
    absmax_kernel<<<dim3(64, g.B), 256, 0, st>>>(a.theta, (size_t)CI * g.Nk, absmax, 2);
    DAGL_LAUNCH_CHECK();
  }
  const int nwords = (g.Nk + 31) / 32;
  if (a.mask_bits) DAGL_CUDA_OK(cudaMemsetAsync(a.mask_bits, 0, (size_t)g.B * g.Nq * nwords * sizeof(uint32_t), st));
  if (a.nnz) DAGL_CUDA_OK(cudaMemsetAsync(a.nnz, 0, (size_t)g.B * g.Nq * sizeof(int32_t), st));
  unsigned* smax = reinterpret_cast<unsigned*>(base + w.smax);
  unsigned* smax2 = smax + (size_t)g.B * tg.nqt * TC_BM;
  DAGL_CUDA_OK(cudaMemsetAsync(smax, 0, 2 * (size_t)g.B * tg.nqt * TC_BM * 4, st));

  if (!a.k_packed) {
    // split entry (embeddings supplied by the caller as fp32): pack them into the operand tiles here.  Keys first: the
    // pack kernel also produces the per-tile column sums from which Kbar is formed when the caller did not supply it;
    // the query pack needs Kbar for the per-query thresholds.  (The full forward writes all of this from the epilogues of
    // the feature-map and embedding kernels: a.k_packed.)
    const float* Kbar = a.Kbar;
    float* colsum = nullptr;
    if (Kbar == nullptr) colsum = reinterpret_cast<float*>(base + w.colsum);
    auto kk = pack_tiles_kernel<TC_BN, 1, TC_BN>;
    const size_t smem_k = (size_t)TC_BN * ED * 4;
    kk<<<dim3(tg.NT, g.B), 256, smem_k, st>>>(g, tg, a.K, absmax, Kp, tilemask, nullptr, nullptr, nullptr, nullptr, colsum);
    DAGL_LAUNCH_CHECK();
    if (Kbar == nullptr) {
      float* kb = a.kbar_out ? a.kbar_out : reinterpret_cast<float*>(base + w.kbar);
      if (int rc = launch_kbar(g, colsum, tg.NT, kb, st)) return rc;
      Kbar = kb;
    }
    auto kq = pack_tiles_kernel<TC_BM, 0, 16>;
    const size_t smem = (size_t)16 * ED * 4;
    DAGL_CUDA_OK(launch_pdl(kq, dim3(tg.nqt * (TC_BM / 16), g.B), 256, smem, st, g, tg, a.Q, absmax, Qp, nullptr, Kbar, a.gamma, a.beta, thr4, nullptr));
    DAGL_LAUNCH_CHECK();
    DAGL_CUDA_OK(launch_pdl(pack_theta_kernel, dim3((tg.NP + 255) / 256, g.B), 256, 0, st, g, tg, a.theta, absmax, Thp, nullptr));
    DAGL_LAUNCH_CHECK();
  }

  const float sm_scale_log2 = a.scale * 1.4426950408889634f;
  dim3 grid((qt_end - qt_begin) * (variant == 4 ? 4 : 2), w.nsplit, g.B);
  const int topk = a.topk > 0 ? a.topk : 0;
  if (topk > TOPK_MAX) {
    call_state().err = "legacy top-k: at most 64 edges per query";
    return -2;
  }
  if (topk > 0) {
    // legacy fixed-top-k variant: exact scores once more -> per-row k largest per key split -> per-row threshold into thr4
    float* lists = reinterpret_cast<float*>(base + w.lists);
    const int nl = tg.NT < TOPK_LISTS ? tg.NT : TOPK_LISTS;
    DAGL_CUDA_OK(cudaFuncSetAttribute(rowmax_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, RMX_SM_TOTAL));
    DAGL_CUDA_OK(launch_pdl(rowmax_tc_kernel<2>, dim3(qt_end - qt_begin, nl, g.B), RM_THREADS, RMX_SM_TOTAL, st, tg, Qp, Kp, nl, qt_begin,
                            qt_end, (unsigned*)nullptr, thr4, absmax, 0.f, topk, (unsigned*)nullptr, tilemask, lists));
    DAGL_LAUNCH_CHECK();
    const int nrows = g.B * tg.nqt * TC_BM;
    DAGL_CUDA_OK(launch_pdl(topk_merge_kernel, (nrows + 127) / 128, 128, 0, st, g, tg, nl, topk, lists, absmax, thr4));
    DAGL_LAUNCH_CHECK();
  }
  // pre-pass: row maxima of the scores (Qh.Kh only) ...
  const int nqg = (qt_end - qt_begin + RM_QT - 1) / RM_QT;
  int pre_split = 148 / (nqg * g.B);
  if (pre_split < 1) pre_split = 1;
  const int max_split = tg.NT;
  if (pre_split > max_split) pre_split = max_split;
  DAGL_CUDA_OK(cudaFuncSetAttribute(rowmax_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, RM_SM_TOTAL));
  DAGL_CUDA_OK(launch_pdl(rowmax_tc_kernel<0>, dim3(nqg, pre_split, g.B), RM_THREADS, RM_SM_TOTAL, st, tg, Qp, Kp, pre_split, qt_begin,
                          qt_end, smax, thr4, absmax, sm_scale_log2, topk, smax2, tilemask, (float*)nullptr));
  DAGL_LAUNCH_CHECK();
  // ... made exact for query tiles with huge logits (normally every CTA exits at once)
  {
    int xsplit = 148 / ((qt_end - qt_begin) * g.B);
    if (xsplit < 1) xsplit = 1;
    if (xsplit > max_split) xsplit = max_split;
    DAGL_CUDA_OK(cudaFuncSetAttribute(rowmax_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, RMX_SM_TOTAL));
    DAGL_CUDA_OK(launch_pdl(rowmax_tc_kernel<1>, dim3(qt_end - qt_begin, xsplit, g.B), RM_THREADS, RMX_SM_TOTAL, st, tg, Qp, Kp, xsplit,
                            qt_begin, qt_end, smax, thr4, absmax, sm_scale_log2, topk, smax2, tilemask, (float*)nullptr));
    DAGL_LAUNCH_CHECK();
  }
  // Hybrid: 4-CTA clusters only fit 33 at a time (GPC sizes), so a launch of <= 32 clusters leaves >= 20 SMs idle.  The last
  // part of the keys then goes to a 2-CTA launch that runs on those SMs CONCURRENTLY (programmatic dependent launch: it
  // starts once all CTAs of the 4-CTA grid are running, see the kernels); the cut balances the two by their measured
  // per-tile rates.  DAGL_HYBRID=0 disables, DAGL_HYBRID=<permille> forces the tail's share of the keys.
  int t_cut = 0, nsplit_run = w.nsplit, nparts = variant == 4 ? 1 : 2;
  if (variant == 4 && w.nsplit == 1 && a.mask_bits == nullptr && a.nnz == nullptr && tg.NT >= 64) {
    const char* hyb_env = getenv("DAGL_HYBRID");                 // read per call: the tests toggle it
    int hyb = hyb_env ? atoi(hyb_env) : -1;
    // Not inside a stream capture: in a captured graph the fold would depend on the tail node only (a full edge), and the tail
    // on the 4-CTA node by a programmatic edge it never waits on — "after both grids" is a property of stream order only.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) hyb = 0;
    const int ncl4 = (qt_end - qt_begin) * g.B;
    const int idle = 148 - 4 * ncl4;
    const TcWs wfull = tc_ws(g, tg, ranged ? qt_end - qt_begin : 0);
    if (hyb != 0 && ncl4 <= 33 && idle >= 8 && wfull.nsplit >= 2) {
      const int slots2 = idle / 2;
      // measured at 1x64x256x256 (profiles/r2/r2w_hybrid_sweep.log): a 4-CTA cluster takes 0.45 us per key tile, a tail
      // cluster 1.05 us + ~15 us of start-up, ~10 tail clusters run at a time.  Balance point of
      //   (NT - x) c4 + o4 = waves2 (x c2 + o2)        [tile units of the 4-CTA kernel]
      // and 20 % less than that: past the balance point the tail is the critical path and the time climbs steeply, and the
      // number of SM pairs the 4-CTA clusters leave free depends on the chip's GPC configuration
      const double c4 = 1.0, c2 = 2.6, o4 = 22.0, o2 = 45.0;
      const double waves2 = (double)((ncl4 + slots2 - 1) / slots2);
      double x = 0.8 * ((double)tg.NT * c4 + o4 - waves2 * o2) / (c4 + waves2 * c2);
      if (hyb > 0) x = (double)tg.NT * hyb / 1000.0;
      const int xt = (int)x;
      if (xt >= 16 && xt <= tg.NT / 2) { t_cut = tg.NT - xt; nsplit_run = 2; nparts = 2; }
    }
  }
  if (int rc = prof_begin(st)) return rc;
  if (variant == 4) {
    DAGL_CUDA_OK(cudaFuncSetAttribute(attend_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S4_TOTAL));
    if (t_cut > 0) grid.y = 1;
    DAGL_CUDA_OK(launch_pdl(attend_tc4_kernel, grid, V4_THREADS, S4_TOTAL, st, g, tg, Qp, Kp, Thp, tilemask, thr4, absmax, smax, smax2,
                            sm_scale_log2, topk, nsplit_run, qt_begin, Opart, lpart, a.mask_bits, a.nnz, t_cut, nparts, t_cut > 0 ? 1 : 0));
    if (t_cut > 0) {
      DAGL_LAUNCH_CHECK();
      DAGL_CUDA_OK(cudaFuncSetAttribute(attend_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_TOTAL));
      DAGL_CUDA_OK(launch_pdl(attend_tc2_kernel, dim3((qt_end - qt_begin) * 2, 1, g.B), TC2_THREADS, S2_TOTAL, st, g, tg, Qp, Kp, Thp,
                              tilemask, thr4, absmax, smax, smax2, sm_scale_log2, topk, nsplit_run, qt_begin, Opart, lpart,
                              a.mask_bits, a.nnz, t_cut, 1, 1));
    }
  } else {
    DAGL_CUDA_OK(cudaFuncSetAttribute(attend_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_TOTAL));
    DAGL_CUDA_OK(launch_pdl(attend_tc2_kernel, grid, TC2_THREADS, S2_TOTAL, st, g, tg, Qp, Kp, Thp, tilemask, thr4, absmax, smax, smax2,
                            sm_scale_log2, topk, w.nsplit, qt_begin, Opart, lpart, a.mask_bits, a.nnz, 0, 0, 0));
  }
  DAGL_LAUNCH_CHECK();
  if (int rc = prof_end(st)) return rc;
  const int q_begin = qt_begin * TC_BM, q_end = qt_end * TC_BM < g.Nq ? qt_end * TC_BM : g.Nq;
  if (a.rows_out != nullptr) {   // sharded use: hand the merged, normalised rows to the caller (fold happens after the gather)
    const int nq_total = g.B * g.Nq;
    if (t_cut > 0)      // after a concurrent tail launch: a plain launch (= after BOTH grids), not a programmatic dependent of the tail
      merge_coef_fixed_kernel<<<(nq_total + 255) / 256, 256, 0, st>>>(g.B, g.Nq, nsplit_run, nparts, q_begin, q_end, lpart, coef);
    else
    DAGL_CUDA_OK(launch_pdl(merge_coef_fixed_kernel, (nq_total + 255) / 256, 256, 0, st, g.B, g.Nq, nsplit_run, nparts, q_begin, q_end, lpart, coef));
    DAGL_LAUNCH_CHECK();
    return launch_merge_rows(g, nsplit_run, q_begin, q_end, Opart, coef, a.rows_out, st);
  }
  return launch_fold_partials(g, nsplit_run, nparts, Opart, lpart, a.y, st, /*after_concurrent_grids=*/t_cut > 0);
}

}  // namespace dagl

#ifdef DAGL_TC_TRACE
extern "C" int dagl_debug_set_tc_mode(int mode) {
  return (int)cudaMemcpyToSymbol(dagl::g_tc_dbg_mode, &mode, sizeof(int));
}
extern "C" int dagl_debug_read_tc_timeline(long long* host_out /*[4][32][24]*/) {
  return (int)cudaMemcpyFromSymbol(host_out, dagl::g_tc_tl, sizeof(long long) * 4 * 32 * 24);
}
extern "C" int dagl_debug_read_tc_trace(long long* host_out /*[1024][24]*/) {
  return (int)cudaMemcpyFromSymbol(host_out, dagl::g_tc_trace, sizeof(long long) * 1024 * 24);
}
#endif
