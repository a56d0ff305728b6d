// Patch embeddings on the tensor cores (tcgen05), sm_100a.
// Reference: fc1 / fc2 applied to unfolded 7x7x16 patches + ReLU, DN_Gray/model/dagl.py:216-221,233-239,248-249.
//
// out[p][e] = relu(b[e] + sum_{ky,kx,c} W[e][c,ky,kx] * Gpad[c][y+ky][x+kx])      (p = pixel (y,x))
// is an implicit GEMM with M = pixels, N = 196 (padded 208), K = 49 taps x 16 channels.  With G stored
// channel-last over the zero-padded image and pixels enumerated in padded-flat order (p' = y*Wp + x),
// the A operand of tap (ky,kx) for 128 consecutive pixels is the contiguous slab
// GpadFlat[p0 + ky*Wp + kx + (0..127)][0..15] — a pixel-shifted view of a 7-row halo held in smem
// (K-major, SWIZZLE_32B, one k-step of 16 channels per tap).  Nothing is unfolded (the reference
// writes 205 MB of key patches to HBM per image at 256^2).
//
// The embeddings feed the score threshold, so they need fp32 accuracy: G and W are rescaled by powers
// of two and split into fp16 hi + lo; acc = Gh.Wh + Gh.Wl + Gl.Wh (fp32 accumulate in TMEM).
//
// Queries (stride 4, TF-SAME padding) are the same computation with fc1 at the pixels
// (4qy - top + 3, 4qx - left + 3): the kernel runs over the image rows that contain query centres and
// stores only those pixels.
#include <cuda_fp16.h>
#include <math.h>
#include "common.cuh"
#include "tc_utils.cuh"

namespace dagl {
using namespace tc;

constexpr int EB_M = 128;                          // pixels per CTA
constexpr int EB_N = 208;                          // 196 padded to 13 * 16
constexpr int EB_SEG_PIX = 144;                    // 128 + 6 (kx) + 7 (alignment) rounded up to 8
constexpr int EB_SEG_BYTES = EB_SEG_PIX * 32;      // 4608
constexpr int EB_G_BYTES = 2 * KS * EB_SEG_BYTES;  // hi | lo, 7 rows each: 64512
constexpr int EB_N0 = 112, EB_N1 = EB_N - EB_N0;   // the 208 outputs are split over two CTAs (N = 112 / 96): 2 CTAs per SM overlap
                                                   // the halo load, MMA and epilogue phases of different tiles
constexpr int EB_WPART_BYTES = EB_N * 16 * 2;      // one tap, one part (hi or lo), all 208 outputs: 6656
constexpr int EB_WTAP_BYTES = 2 * EB_WPART_BYTES;  // hi | lo: 13312
// packed weights: [output half][tap][hi | lo][2 k-chunks][rows of the half][8 ch] so that a CTA streams only its half
constexpr int EB_WTAP0_BYTES = 2 * EB_N0 * 32;     // 7168: one tap of output half 0 (hi | lo)
constexpr int EB_WTAP1_BYTES = 2 * EB_N1 * 32;     // 6144
constexpr int EB_WHALF1_OFF = KK * EB_WTAP0_BYTES; // start of output half 1 in the packed array
constexpr int EB_WSTAGE_BYTES = EB_WTAP0_BYTES;
constexpr int EB_QA_BYTES = 2 * EB_M * 16 * 2;     // 8192: one tap of a 128-query patch tile, hi | lo, [2 k-chunks][128 rows][16 B]
constexpr int EB_QSTAGE_BYTES = EB_QA_BYTES + EB_WSTAGE_BYTES;   // query launch: the A operand streams with the weights
constexpr int EB_WSTAGES = KS;                     // 7 = one row of taps: tap (ky, kx) of EVERY item uses stage kx (a compile-time
                                                   // constant in the unrolled issue loop) and every stage completes 7 phases per item.
                                                   // (14 stages measured no faster: the kernel is bound by the L2 -> SM stream of the
                                                   // weight taps, 343 KB per 128-pixel item, not by its latency.)
#ifndef EB_QSTAGES
#define EB_QSTAGES 14                              // query launch: deeper ring (no halo stages there, and its loop tap -> commit -> refill
#endif                                             // -> data is latency-bound: 20 us per item whatever the number of items)
constexpr int EB_WSTAGES_Q = EB_QSTAGES;
// how often stage s of a WST-deep ring is filled per item (49 taps)
__host__ __device__ constexpr int eb_stage_uses(int s, int wst) { return (KK - s + wst - 1) / wst; }
constexpr int EB_SM_G = 0;                                           // two halo stages
constexpr int EB_SM_W = 2 * EB_G_BYTES;                              // 129024 = 126 * 1024
constexpr int EB_SM_CSUM = EB_SM_W + EB_WSTAGES * EB_WSTAGE_BYTES;   // column-sum exchange [4][112] floats
constexpr int EB_SM_BAR = EB_SM_CSUM + 4 * EB_N0 * 4;
constexpr int EB_MU_BYTES = 2 * EB_M * 8;            // query launch: the two epilogue warps of a row exchange their partial mu (doubles)
constexpr int EB_SM_TOTAL = EB_SM_BAR + 384 + MAX_HEADS * EB_N * 4 + EB_MU_BYTES;   // mbarriers + TMEM base (384 B), bias vectors [heads][208], mu exchange
constexpr int EB_SMQ_TOTAL = EB_WSTAGES_Q * EB_QSTAGE_BYTES + 4 * EB_N0 * 4 + 384 + MAX_HEADS * EB_N * 4 + EB_MU_BYTES;   // query launch
static_assert(EB_SMQ_TOTAL <= 232448, "query embed kernel exceeds the 227 KB dynamic shared memory limit");
#ifndef EB_NPROD
#define EB_NPROD 2                                 // producer warps of the weight-tap ring (0, 7, then 12, 13)
#endif
constexpr int EB_THREADS = 384 + 32 * (EB_NPROD - 2);   // warps 0, 7 (12, 13) weight taps, 1 MMA issuer, 2-5 + 8-11 epilogue, 6 G halos
constexpr int EB_ACC_COLS = 2 * EB_N0;             // one accumulator buffer: main (hi.hi) at +0, cross terms at +112
constexpr int EB_TMEM_COLS = 512;                  // two accumulator buffers (448 columns used)
static_assert(EB_SM_W % 1024 == 0, "weight ring alignment");
static_assert(EB_SM_TOTAL <= 232448, "embed kernel exceeds the 227 KB dynamic shared memory limit");

struct EmbGeom {
  int Wp, NkP, ntile, NPG, nqt;
};
constexpr int EB_KTILE = 48;                       // key-tile size of the graph kernel (attend_tc.cu TC_BN)
static EmbGeom emb_geom(const Geom& g) {
  EmbGeom e;
  e.Wp = (g.W + 2 * PADK + 7) & ~7;                // the padded-flat key enumeration of attend_tc.cu (tc_geom)
  e.NkP = (g.H - 1) * e.Wp + g.W;
  const int slots = ((e.NkP + EB_KTILE - 1) / EB_KTILE) * EB_KTILE;     // every slot of every 48-key tile gets written
  e.ntile = (slots + EB_M - 1) / EB_M;
  int np = (g.H + 2 * PADK) * e.Wp;
  const int need = EB_M * e.ntile + 2 * PADK * e.Wp + EB_SEG_PIX + 8;
  e.NPG = ((np > need ? np : need) + 7) & ~7;
  e.nqt = (g.Nq + EB_M - 1) / EB_M;
  return e;
}

__device__ __forceinline__ float pow2_scale_e(unsigned absmax_bits, int target) {
  const float a = __uint_as_float(absmax_bits);
  if (!(a > 0.f) || !isfinite(a)) return 1.f;
  int e;
  frexpf(a, &e);
  return ldexpf(1.f, target - e);
}

// max |x| over a tensor (weights) -> absmax[slot]
__global__ void absmax_flat_kernel(const float* __restrict__ x, int n, unsigned* __restrict__ out) {
  float m = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

// fc weight [196][784] (e, (c,ky,kx)) -> per output half and tap: [hi|lo][2 chunks][rows/8 row groups][8 rows (e)][8 c] fp16
__global__ void __launch_bounds__(256)
pack_fc_kernel(const float* __restrict__ w, const unsigned* __restrict__ wmax, uint8_t* __restrict__ out) {
  const int tap = blockIdx.x;
  const float scale = pow2_scale_e(*wmax, 14);
  for (int o = threadIdx.x; o < EB_WPART_BYTES / 16; o += 256) {      // one 16-byte chunk = 8 channels of one e
    const int kc = o / EB_N, e = o % EB_N;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x0 = 0.f, x1 = 0.f;
      if (e < ED) {
        const int c = kc * 8 + 2 * j;
        x0 = __ldg(w + (size_t)e * VD + c * KK + tap) * scale;
        x1 = __ldg(w + (size_t)e * VD + (c + 1) * KK + tap) * scale;
      }
      const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
      const __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1));
      hi[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      lo[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    const int half = e >= EB_N0, el = e - half * EB_N0, rows = half ? EB_N1 : EB_N0;
    uint8_t* base = out + (half ? EB_WHALF1_OFF + (size_t)tap * EB_WTAP1_BYTES : (size_t)tap * EB_WTAP0_BYTES) +
                    (size_t)(kc * rows + el) * 16;                             // [kc][e] chunk order == K-major no-swizzle
    *reinterpret_cast<uint4*>(base) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + rows * 32) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// QG = false: keys (every pixel) -> the graph kernel's fp16 hi|lo key tiles + per-item column sums for Kbar
// QG = true : queries -> the graph kernel's fp16 hi|lo query tiles + the per-query threshold terms (mu partial, gamma, beta);
//             the A operand is the gathered patch image (gather_qpatch_kernel)
// The fp16 scales are the a-priori bounds in absmax slots 0 (Q) and 1 (K), written by the feature-map epilogue; `out`
// (fp32 [rows][196]) is only written for the debug entry.
//
// Persistent kernel: one CTA per SM walks the work items (image, 128-pixel tile, output half) with a stride of the grid.
// The G halo and the TMEM accumulators are double buffered, the weight taps stream through one ring that runs across
// items, so the tensor pipe goes from the last tap of an item straight to the first tap of the next while the epilogue
// warps drain the previous accumulator (before: two CTAs per SM whose load / MMA / epilogue phases ran in lock step).
// Work item w -> (image, tile, output half).  Keys: tile = 128 consecutive padded-flat pixel slots.  Queries: tile = 128
// consecutive queries of the gathered patch image (pack_qpatch_kernel).
template <bool QG>
__device__ __forceinline__ void embed_item(const Geom& g, const EmbGeom& eg, int w, int& img, int& tile, int& eh) {
  const int per_img = QG ? eg.nqt : eg.ntile;
  eh = w & 1;
  img = (w >> 1) / per_img;
  tile = (w >> 1) % per_img;
}

template <bool QG>
__global__ void __launch_bounds__(EB_THREADS, 1)
embed_tc_kernel(Geom g, EmbGeom eg, const uint8_t* __restrict__ ghi /*QG: the gathered query-patch image*/,
                const uint8_t* __restrict__ glo, HeadPtrs wp_h /*packed fc weights per head*/, HeadPtrs bias_h,
                const unsigned* __restrict__ absmax_in /*[B][4]: bounds on Q, K, theta, G*/, HeadPtrs wmax_h,
                float* __restrict__ out /*nullable: fp32 copy for the debug entry*/,
                uint8_t* __restrict__ ktiles /*keys: fp16 hi|lo key tiles of the graph kernel; queries: its query tiles*/,
                float* __restrict__ colsum /*keys: [B][ntile][196] column sums of an item's rows*/, EmbQOut qo /*queries only*/) {
  pdl_prologue();
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int STAGE = QG ? EB_QSTAGE_BYTES : EB_WSTAGE_BYTES;      // QG: [A hi 4 KB | A lo 4 KB | weight tap]
  constexpr int W_OFF = QG ? EB_QA_BYTES : 0;
  constexpr int SM_W = QG ? 0 : EB_SM_W;                             // queries: no halo stages
  constexpr int WST = QG ? EB_WSTAGES_Q : EB_WSTAGES;               // depth of the tap ring
  constexpr int SM_CSUM = SM_W + WST * STAGE;
  constexpr int SM_BAR = SM_CSUM + 4 * EB_N0 * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint64_t* g_full = bars + 0;                   // [2]
  uint64_t* g_empty = bars + 2;                  // [2]
  uint64_t* d_full = bars + 4;                   // [2]
  uint64_t* d_empty = bars + 6;                  // [2] 8 arrivals (one per epilogue warp)
  uint64_t* w_full = bars + 8;                   // [WST]
  uint64_t* w_empty = bars + 8 + WST;            // [WST]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8 + 2 * WST);
  static_assert((8 + 2 * WST) * 8 + 4 <= 384, "barrier area");

  const int warp = warp_id_uniform();
  const int tid = threadIdx.x;
  const int nwork = g.B * 2 * (QG ? eg.nqt : eg.ntile);

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(g_full + i, 1); mbar_init(g_empty + i, 1);
      mbar_init(d_full + i, 1); mbar_init(d_empty + i, 8);
    }
    for (int i = 0; i < WST; ++i) { mbar_init(w_full + i, 1); mbar_init(w_empty + i, 1); }
    mbar_init_fence();
  }
  if (warp == 1) tmem_alloc<EB_TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;
  float* bias_all = reinterpret_cast<float*>(smem + SM_BAR + 384);     // [NH][208] (behind the mbarriers)
  for (int e = tid; e < EB_N * g.NH; e += EB_THREADS)
    bias_all[e] = (e % EB_N) < ED ? __ldg(static_cast<const float*>(bias_h.p[e / EB_N]) + (e % EB_N)) : 0.f;
  __syncthreads();

  if (warp == 0 || warp == 7 || warp >= 12) {
    // ===================== producers: weight taps (one ring that runs across items) =====================
    // A single thread issues a tap (try_wait + expect_tx + bulk copy) every ~306 cycles whatever the copy size (measured:
    // tools/bulk_copy_probe.cu, profiles/r2/r2p_bulk_probe.log), more than the tap's three MMAs take (176), so EB_NPROD
    // warps share the taps round-robin; every stage still sees its fills in order.
    const int pidx = warp == 0 ? 0 : warp == 7 ? 1 : warp - 10;
    if (elect_one()) {
      int it = 0;
      for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        int img, tile, eh;
        embed_item<QG>(g, eg, w, img, tile, eh);
        const uint32_t tap_bytes = eh ? EB_WTAP1_BYTES : EB_WTAP0_BYTES;
        const uint8_t* wsrc = static_cast<const uint8_t*>(wp_h.p[g.head(img)]) + (eh ? EB_WHALF1_OFF : 0);
        const uint8_t* asrc = QG ? ghi + ((size_t)img * eg.nqt + tile) * (size_t)(KK * EB_QA_BYTES) : nullptr;
        for (int t = pidx; t < KK; t += EB_NPROD) {
          const int s = t % WST;
          const uint32_t use = (uint32_t)(it * eb_stage_uses(s, WST) + t / WST);        // how often stage s was filled before
          mbar_wait(w_empty + s, (use & 1u) ^ 1u);
          mbar_arrive_expect_tx(w_full + s, tap_bytes + (QG ? EB_QA_BYTES : 0));
          if (QG) bulk_g2s(smem + SM_W + s * STAGE, asrc + (size_t)t * EB_QA_BYTES, EB_QA_BYTES, w_full + s);
          bulk_g2s(smem + SM_W + s * STAGE + W_OFF, wsrc + (size_t)t * tap_bytes, tap_bytes, w_full + s);
        }
        ++it;
      }
    }
  } else if (warp == 6) {
    // ===================== producer: G halos (2 stages, one item ahead of the MMAs); keys only =====================
    if (!QG && elect_one()) {
      int it = 0;
      for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        int img, tile, eh;
        embed_item<QG>(g, eg, w, img, tile, eh);
        const int p0 = tile * EB_M, hs = it & 1;
        mbar_wait(g_empty + hs, ((uint32_t)(it >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(g_full + hs, EB_G_BYTES);
        for (int part = 0; part < 2; ++part) {
          const uint8_t* src = (part ? glo : ghi) + (size_t)img * eg.NPG * 32;
          for (int ky = 0; ky < KS; ++ky) {
            const int first = (p0 + ky * eg.Wp) & ~7;
            bulk_g2s(smem + EB_SM_G + hs * EB_G_BYTES + (part * KS + ky) * EB_SEG_BYTES, src + (size_t)first * 32,
                     EB_SEG_BYTES, g_full + hs);
          }
        }
        ++it;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      int it = 0;
      for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        int img, tile, eh;
        embed_item<QG>(g, eg, w, img, tile, eh);
        const int p0 = tile * EB_M, hs = it & 1, ab = it & 1;
        const int ncols = eh ? EB_N1 : EB_N0;
        const uint32_t idesc = instr_desc(EB_M, (uint32_t)ncols, FMT_F16, FMT_F16, 0, 0);
        const uint32_t d_main = tbase + ab * EB_ACC_COLS, d_cross = d_main + EB_N0;
        if (!QG) mbar_wait(g_full + hs, (uint32_t)(it >> 1) & 1u);
        mbar_wait(d_empty + ab, ((uint32_t)(it >> 1) & 1u) ^ 1u);          // the epilogue has drained this accumulator
        tc_fence_after();
        // Single-thread issue: the 49 taps are fully unrolled with every descriptor a pre-computed base plus an
        // immediate (a runtime-indexed loop costs ~90 cycles per MMA in descriptor arithmetic and uniform-register
        // moves and left the tensor pipe idle more than half of the time).
        const uint32_t gbase = smem_u32(smem + EB_SM_G + hs * EB_G_BYTES);
        // A: K-major SWIZZLE_32B, rows (pixels) 32 B apart, 8-row groups 256 B apart
        // (queries: the gathered patches arrive with the weight tap: K-major no swizzle, LBO 2048 between the two k-chunks)
        constexpr uint64_t a_bits = QG ? (((uint64_t)(2048 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46))
                                       : (((uint64_t)1 << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)6 << 61));
        uint32_t a_row[KS];                                       // (start address >> 4) of tap (ky, kx = 0), hi part
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) a_row[ky] = QG ? 0u : (gbase + ky * EB_SEG_BYTES + (((p0 + ky * eg.Wp) & 7) << 5)) >> 4;
        // B: K-major no swizzle, this item's output half: LBO = (ncols / 8) * 128, SBO = 128
        const uint64_t b_bits = ((uint64_t)(((uint32_t)(ncols / 8) * 128) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
        const uint32_t w_row = smem_u32(smem + SM_W) >> 4;
        const uint32_t w_lo = (uint32_t)(ncols * 32) >> 4;        // lo part follows the hi part

#pragma unroll
        for (int t = 0; t < KK; ++t) {
          const int ky = t / KS, kx = t % KS, st = t % WST;           // compile-time after unrolling
          mbar_wait(w_full + st, (uint32_t)(it * eb_stage_uses(st, WST) + t / WST) & 1u);
          tc_fence_after();
          const uint32_t a_hi = QG ? w_row + st * (STAGE >> 4) : a_row[ky] + kx * 2;
          const uint64_t da_hi = a_bits | (uint64_t)(a_hi & 0x3FFF);
          const uint64_t da_lo = a_bits | (uint64_t)((a_hi + (QG ? (EB_QA_BYTES / 2) >> 4 : (KS * EB_SEG_BYTES) >> 4)) & 0x3FFF);
          const uint64_t db_hi = b_bits | (uint64_t)((w_row + st * (STAGE >> 4) + (W_OFF >> 4)) & 0x3FFF);
          const uint64_t db_lo = b_bits | (uint64_t)((w_row + st * (STAGE >> 4) + (W_OFF >> 4) + w_lo) & 0x3FFF);
          // Tensor-core fp32 accumulation truncates relative to the running sum: the two cross terms
          // (~2^-11 of the result) get their own accumulator so that the main chain has 49 steps, not 147;
          // the epilogue adds the two in round-to-nearest fp32.
          mma_f16_ss_a_fill(d_main, da_hi, db_hi, idesc, t > 0);               // Gh.Wh
          mma_f16_ss_a_lastuse(d_cross, da_hi, db_lo, idesc, t > 0);           // Gh.Wl
          mma_f16_ss(d_cross, da_lo, db_hi, idesc, 1);                         // Gl.Wh
          mma_commit(w_empty + st);
        }
        mma_commit(d_full + ab);
        if (!QG) mma_commit(g_empty + hs);
        ++it;
      }
    }
  } else {
    // ===================== epilogue (warps 2-5 and 8-11): thread = pixel row; the two warps of a TMEM lane quadrant
    // take alternate pairs of 16-column chunks =====================
    const int egrp = warp >= 8 ? 1 : 0;
    const int quad = warp & 3, lane = tid & 31;
    const int r = quad * 32 + lane;
    const uint32_t trow0 = tbase + ((uint32_t)(quad * 32) << 16);
    const int ntile_k = (eg.NkP + EB_KTILE - 1) / EB_KTILE;
    constexpr int K_HALF = EB_KTILE * EB_N * 2;                              // 19968: hi part, then lo part
    constexpr int Q_HALF = EB_M * EB_N * 2;                                  // 53248: query tile, hi part
    float* csum_s = reinterpret_cast<float*>(smem + SM_CSUM);             // [4 warps][112]
    double* mu_s = reinterpret_cast<double*>(smem + SM_BAR + 384 + MAX_HEADS * EB_N * 4);   // [2][128]
    int it = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
      int img, tile, eh;
      embed_item<QG>(g, eg, w, img, tile, eh);
      const int ab = it & 1;
      const float winv = 1.f / pow2_scale_e(*static_cast<const unsigned*>(wmax_h.p[g.head(img)]), 14);
      const float* bias_s = bias_all + g.head(img) * EB_N;
      const int e0 = eh ? EB_N0 : 0, ncols = eh ? EB_N1 : EB_N0;
      const int p = tile * EB_M + r;
      const int y = p / eg.Wp, x = p % eg.Wp;
      bool valid;
      size_t orow;
      if (!QG) {
        valid = (p < eg.NkP) && (x < g.W);
        orow = (size_t)img * g.Nk + (size_t)y * g.W + x;
      } else {
        valid = p < g.Nq;                                              // row = query
        orow = (size_t)img * g.Nq + p;
      }
      const float inv = winv / pow2_scale_e(absmax_in[img * 4 + 3], 14);
      // fp16 scale of the tiles: an a-priori bound on Q / K (absmax slots 0 / 1), so no pass over the embeddings is needed
      const float oscale = pow2_scale_e(absmax_in[img * 4 + (QG ? 0 : 1)], 14);
      const int kt = p / EB_KTILE, kr = p % EB_KTILE;                        // key tile / row of this pixel slot
      uint8_t* ktile = (!QG && kt < ntile_k) ? ktiles + ((size_t)img * ntile_k + kt) * (size_t)(2 * K_HALF) : nullptr;
      uint8_t* qtile = QG ? ktiles + ((size_t)img * eg.nqt + tile) * (size_t)(2 * Q_HALF) : nullptr;
      const float* kbar = QG ? qo.kbar + (size_t)img * ED : nullptr;
      const uint32_t trow = trow0 + ab * EB_ACC_COLS;
      double mu = 0.0;
      mbar_wait(d_full + ab, (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      // one 16-column chunk: bias + ReLU, fp16 hi|lo tile store, and column sums (keys) / the mu partial (queries)
      auto process = [&](int c16, const uint32_t (&v)[16], const uint32_t (&vc)[16]) {
        const int eb = e0 + c16 * 16;
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int e = eb + i;
          f[i] = (valid && e < ED) ? fmaxf((__uint_as_float(v[i]) + __uint_as_float(vc[i])) * inv + bias_s[e], 0.f) : 0.f;
        }
        if (valid && out != nullptr) {
          float4* dst = reinterpret_cast<float4*>(out + orow * ED + eb);
          const int n4 = min(4, (ED - eb) / 4);                    // 196 = 12*16 + 4
          for (int i = 0; i < n4; ++i) dst[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
        }
        if (QG || ktile != nullptr) {                              // dummy key slots (x >= W, p >= NkP) get zero rows
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float x0 = f[h8 * 8 + 2 * j] * oscale, x1 = f[h8 * 8 + 2 * j + 1] * oscale;
              const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
              const __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1));
              hi[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
              lo[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
            }
            const int kc = eb / 8 + h8;                            // 16-byte chunk column of the K-major no-swizzle tile
            if (QG) {
              const uint32_t off = (uint32_t)(kc * (EB_M / 8) * 128 + r * 16);
              *reinterpret_cast<uint4*>(qtile + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(qtile + Q_HALF + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            } else {
              const uint32_t off = (uint32_t)(kc * (EB_KTILE / 8) * 128 + (kr / 8) * 128 + (kr % 8) * 16);
              *reinterpret_cast<uint4*>(ktile + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(ktile + K_HALF + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
        }
        if (QG) {
          // mu = mean_k S[q,k] = Q[q,:] . Kbar (dagl.py:256 through the row-mean identity), fp64 partial over these columns
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (eb + i < ED) mu += (double)f[i] * (double)__ldg(kbar + eb + i);
        } else {
          // column sums over this warp's 32 rows, then over the 4 warps below (Kbar = mean_k K).  Transposing butterfly:
          // every step halves the columns a lane carries and doubles the rows they cover (8+4+2+1+1 = 16 shuffles instead
          // of 16 x 5); fixed order, so the result is deterministic.  Lane l ends with column (l >> 1) & 15.
          float a8[8], a4[4], a2[2];
          const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4, u2 = lane & 2;
#pragma unroll
          for (int i = 0; i < 8; ++i) a8[i] = (u16 ? f[8 + i] : f[i]) + __shfl_xor_sync(0xffffffffu, u16 ? f[i] : f[8 + i], 16);
#pragma unroll
          for (int i = 0; i < 4; ++i) a4[i] = (u8 ? a8[4 + i] : a8[i]) + __shfl_xor_sync(0xffffffffu, u8 ? a8[i] : a8[4 + i], 8);
#pragma unroll
          for (int i = 0; i < 2; ++i) a2[i] = (u4 ? a4[2 + i] : a4[i]) + __shfl_xor_sync(0xffffffffu, u4 ? a4[i] : a4[2 + i], 4);
          float a1 = (u2 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, u2 ? a2[0] : a2[1], 2);
          a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
          if ((lane & 1) == 0) csum_s[quad * EB_N0 + c16 * 16 + (lane >> 1)] = a1;
        }
      };
      // two chunks per TMEM round trip: the epilogue is one warp per scheduler, so the load latency is otherwise exposed
      const int nch = ncols / 16;
#pragma unroll 1
      for (int c16 = 2 * egrp; c16 < nch; c16 += 4) {
        uint32_t v0[16], w0[16], v1[16], w1[16];
        const bool two = c16 + 1 < nch;
        tmem_ld16(trow + c16 * 16, v0);
        tmem_ld16(trow + EB_N0 + c16 * 16, w0);
        if (two) {
          tmem_ld16(trow + (c16 + 1) * 16, v1);
          tmem_ld16(trow + EB_N0 + (c16 + 1) * 16, w1);
        }
        tmem_wait_ld();
        process(c16, v0, w0);
        if (two) process(c16 + 1, v1, w1);
      }
      // this accumulator may be overwritten by the MMAs of the item after next
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d_empty + ab);
      if (QG) {
        mu_s[egrp * EB_M + r] = mu;
        asm volatile("bar.sync 1, 256;" ::: "memory");              // the eight epilogue warps: both column groups of every row
        if (egrp == 0) {
          const size_t idx = ((size_t)img * eg.nqt + tile) * EB_M + r;
          qo.thr4[idx * 4 + eh] = (float)(mu_s[r] + mu_s[EB_M + r]);   // the graph kernel adds the two output halves
          if (eh == 0) {
            qo.thr4[idx * 4 + 2] = valid ? __ldg(qo.gamma + orow) : 0.f;
            qo.thr4[idx * 4 + 3] = valid ? __ldg(qo.beta + orow) : -1.f;   // rows past Nq: nothing is ever selected
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      } else {
        asm volatile("bar.sync 1, 256;" ::: "memory");              // the eight epilogue warps: csum_s complete
        if (egrp == 0 && r < ncols && e0 + r < ED)
          colsum[((size_t)img * eg.ntile + tile) * ED + e0 + r] =
              ((csum_s[r] + csum_s[EB_N0 + r]) + csum_s[2 * EB_N0 + r]) + csum_s[3 * EB_N0 + r];
        asm volatile("bar.sync 1, 256;" ::: "memory");              // ... and read before the next item overwrites it
      }
      ++it;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<EB_TMEM_COLS>(tbase);
}

// Query patches gathered for the query embedding GEMM: for every 128-query tile and tap (ky,kx) the 16-channel vectors of
// the queries' patch pixel (4qy - top + ky, 4qx - left + kx) of G, fp16 hi | lo, laid out as the K-major no-swizzle UMMA
// A operand [tap][hi|lo][2 k-chunks][128 rows][8 ch].  Queries are 1/16 of the pixels, so this (0.4 MB per tile) is cheap,
// unlike unfolding the keys; it turns the query embedding (fc1 on the stride-4 unfold, dagl.py:216-221,248) into a dense
// [Nq x 784] x [784 x 196] GEMM instead of a 7x7 convolution evaluated at every pixel of the query rows.
// Source: the zero-padded fp16 hi / lo images of G written by the feature-map epilogue (the SAME padding of the stride-4
// unfold, dagl.py:126-136, never reaches outside their 3-pixel border), so this is a pure gather of 16-byte chunks.
// The first CTA row also writes the 48-bit validity masks of the graph kernel's key tiles (pure geometry).
__global__ void __launch_bounds__(256)
gather_qpatch_kernel(Geom g, EmbGeom eg, const uint8_t* __restrict__ ghi, const uint8_t* __restrict__ glo,
                     uint8_t* __restrict__ qimg, unsigned long long* __restrict__ tilemask, int ntile_k) {
  pdl_prologue();
  const int qt = blockIdx.x / KS, ky = blockIdx.x % KS, img = blockIdx.y;      // one CTA per (query tile, tap row)
  if (blockIdx.x == 0) {
    for (int t = threadIdx.x; t < ntile_k; t += 256) {
      unsigned long long m = 0ull;
      int kp = t * EB_KTILE, x = kp % eg.Wp;
      for (int r = 0; r < EB_KTILE; ++r, ++kp) {
        if (kp < eg.NkP && x < g.W) m |= 1ull << r;
        if (++x == eg.Wp) x = 0;
      }
      tilemask[(size_t)img * ntile_k + t] = m;
    }
  }
  uint8_t* tile = qimg + ((size_t)img * eg.nqt + qt) * (size_t)(KK * EB_QA_BYTES);
  const uint8_t* sh = ghi + (size_t)img * eg.NPG * 32;
  const uint8_t* sl = glo + (size_t)img * eg.NPG * 32;
  for (int o = threadIdx.x; o < KS * 2 * EB_M; o += 256) {              // one 16-byte chunk: 8 channels of (tap, k-chunk, row)
    const int row = o % EB_M, kc = (o / EB_M) & 1, t = ky * KS + o / (2 * EB_M);
    const int q = qt * EB_M + row;
    uint4 hv = make_uint4(0u, 0u, 0u, 0u), lv = hv;
    if (q < g.Nq) {
      const int qy = q / g.nqx, qx = q % g.nqx;
      const int y = qy * SQ - g.qpad_top + t / KS, x = qx * SQ - g.qpad_left + t % KS;     // in [-3, H+2] x [-3, W+2]
      const size_t rec = (size_t)(y + PADK) * eg.Wp + (x + PADK);
      const int c = kc ^ ((int)(rec >> 2) & 1);                        // undo the pre-applied SWIZZLE_32B
      hv = __ldg(reinterpret_cast<const uint4*>(sh + rec * 32) + c);
      lv = __ldg(reinterpret_cast<const uint4*>(sl + rec * 32) + c);
    }
    uint8_t* dst = tile + (size_t)t * EB_QA_BYTES + (size_t)(kc * EB_M + row) * 16;     // [kc][row] chunk order
    *reinterpret_cast<uint4*>(dst) = hv;
    *reinterpret_cast<uint4*>(dst + EB_QA_BYTES / 2) = lv;
  }
}

// ---- host side -----------------------------------------------------------------------------
static inline size_t align_up_e(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t packed_w_bytes();
size_t embed_tc_workspace_bytes(const Geom& g) {
  const EmbGeom eg = emb_geom(g);
  return 2 * align_up_e((size_t)g.B * eg.NPG * 32) +                       // fp16 hi / lo images of G (feature-map epilogue)
         align_up_e((size_t)g.B * eg.nqt * KK * EB_QA_BYTES);              // query patch tiles
}
void embed_tc_g_buffers(const Geom& g, void* ws, uint8_t** ghi, uint8_t** glo, int* npg) {
  const EmbGeom eg = emb_geom(g);
  char* p = static_cast<char*>(ws);
  *ghi = reinterpret_cast<uint8_t*>(p);
  *glo = reinterpret_cast<uint8_t*>(p + align_up_e((size_t)g.B * eg.NPG * 32));
  *npg = eg.NPG;
}
int embed_tc_num_tiles(const Geom& g) { return emb_geom(g).ntile; }

// meta[2*which + 0] = max_e sum_j |w[e][j]|, meta[2*which + 1] = max_e |bias[e]|   (which = blockIdx.x: 0 fc1, 1 fc2)
// -> a-priori bound  |fc(x)|_inf <= max|x| * meta[0] + meta[1]  used as the fp16 scale of the fused key pack
__global__ void __launch_bounds__(256)
fc_meta_kernel(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
               const float* __restrict__ b2, float* __restrict__ meta) {
  __shared__ float red[2][8];
  const float* w = blockIdx.x ? w2 : w1;
  const float* bb = blockIdx.x ? b2 : b1;
  float l1 = 0.f, bm = 0.f;
  for (int e = threadIdx.x >> 5; e < ED; e += 8) {                 // one warp per output row
    float t = 0.f;
    for (int j = threadIdx.x & 31; j < VD; j += 32) t += fabsf(__ldg(w + (size_t)e * VD + j));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    l1 = fmaxf(l1, t);
    bm = fmaxf(bm, fabsf(__ldg(bb + e)));
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = l1; red[1][threadIdx.x >> 5] = bm; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) { l1 = fmaxf(l1, red[0][k]); bm = fmaxf(bm, red[1][k]); }
    meta[2 * blockIdx.x + 0] = l1 * 1.0001f;                          // margin for the rounding of the sums
    meta[2 * blockIdx.x + 1] = bm;
  }
}

// packed fc1 | packed fc2 | wmax[2], meta[4] (fc_meta_kernel)
static size_t packed_w_bytes() { return align_up_e((size_t)KK * EB_WTAP_BYTES); }
size_t embed_tc_packed_weights_bytes() { return 2 * packed_w_bytes() + align_up_e(64); }

int launch_pack_fc_weights(const float* fc1_w, const float* fc1_b, const float* fc2_w, const float* fc2_b, void* packed,
                           size_t packed_bytes, cudaStream_t st) {
  if (packed_bytes < embed_tc_packed_weights_bytes()) {
    call_state().err = "packed-weights buffer too small";
    return -3;
  }
  uint8_t* w1 = static_cast<uint8_t*>(packed);
  uint8_t* w2 = w1 + packed_w_bytes();
  unsigned* wmax = reinterpret_cast<unsigned*>(w2 + packed_w_bytes());
  DAGL_CUDA_OK(cudaMemsetAsync(wmax, 0, 2 * sizeof(unsigned), st));
  absmax_flat_kernel<<<64, 256, 0, st>>>(fc1_w, ED * VD, wmax + 0);
  DAGL_LAUNCH_CHECK();
  absmax_flat_kernel<<<64, 256, 0, st>>>(fc2_w, ED * VD, wmax + 1);
  DAGL_LAUNCH_CHECK();
  pack_fc_kernel<<<KK, 256, 0, st>>>(fc1_w, wmax + 0, w1);
  DAGL_LAUNCH_CHECK();
  pack_fc_kernel<<<KK, 256, 0, st>>>(fc2_w, wmax + 1, w2);
  DAGL_LAUNCH_CHECK();
  fc_meta_kernel<<<2, 256, 0, st>>>(fc1_w, fc1_b, fc2_w, fc2_b, reinterpret_cast<float*>(wmax + 2));
  DAGL_LAUNCH_CHECK();
  return 0;
}

const float* embed_tc_fc_meta(const void* packed) {
  return reinterpret_cast<const float*>(static_cast<const uint8_t*>(packed) + 2 * packed_w_bytes()) + 2;   // after wmax[2]
}

// Both embeddings, written straight into the graph kernel's operand tiles (EmbTargets): key tiles + column sums, then
// Kbar = mean_k K, then query tiles + per-query threshold terms (mu partials need Kbar, hence the order).  The fp16 images
// of G in `ws` were written by the feature-map epilogue.  hw.packed[h] is never null here.  Q / K (nullable): fp32 copies
// for the debug entry.
int launch_embed_tc(const Geom& g, const HeadWeights& hw, float* Q, float* K, const unsigned* absmax, void* ws, size_t ws_bytes,
                    const float* gamma, const float* beta, const EmbTargets& out, cudaStream_t st) {
  const EmbGeom eg = emb_geom(g);
  if (ws_bytes < embed_tc_workspace_bytes(g)) {
    call_state().err = "embed (tc) workspace too small";
    return -3;
  }
  uint8_t *ghi, *glo;
  int npg;
  embed_tc_g_buffers(g, ws, &ghi, &glo, &npg);
  uint8_t* qimg = static_cast<uint8_t*>(ws) + 2 * align_up_e((size_t)g.B * eg.NPG * 32);
  HeadPtrs w1{}, w2{}, b1{}, b2{}, wmax1{}, wmax2{};
  for (int h = 0; h < g.NH; ++h) {
    const uint8_t* packed = static_cast<const uint8_t*>(hw.packed[h]);
    const unsigned* wmax = reinterpret_cast<const unsigned*>(packed + 2 * packed_w_bytes());
    w1.p[h] = packed; w2.p[h] = packed + packed_w_bytes();
    wmax1.p[h] = wmax + 0; wmax2.p[h] = wmax + 1;
    b1.p[h] = hw.fc1_b[h]; b2.p[h] = hw.fc2_b[h];
  }
  int dev = 0, sms = 148;
  DAGL_CUDA_OK(cudaGetDevice(&dev));
  DAGL_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int ntile_k = (eg.NkP + EB_KTILE - 1) / EB_KTILE;
  // queries: gather the patches (+ key-tile validity masks)
  DAGL_CUDA_OK(launch_pdl(gather_qpatch_kernel, dim3(eg.nqt * KS, g.B), 256, 0, st, g, eg, ghi, glo, qimg, out.tilemask, ntile_k));
  DAGL_LAUNCH_CHECK();
  // keys: implicit GEMM over the halo, written straight into the graph kernel's key tiles
  // (a two-tiles-per-weight-pass variant measured no faster: with one accumulator set per tile its MMA phase and its
  // epilogue run back to back)
  const EmbQOut none{};
  const int nwork_k = g.B * 2 * eg.ntile;
  DAGL_CUDA_OK(cudaFuncSetAttribute(embed_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EB_SM_TOTAL));
  DAGL_CUDA_OK(launch_pdl(embed_tc_kernel<false>, nwork_k < sms ? nwork_k : sms, EB_THREADS, EB_SM_TOTAL, st, g, eg, ghi, glo, w2,
                          b2, absmax, wmax2, K, out.ktiles, out.colsum, none));
  DAGL_LAUNCH_CHECK();
  if (int rc = launch_kbar(g, out.colsum, eg.ntile, out.kbar, st)) return rc;
  // queries: a dense GEMM of the gathered patches against fc1; epilogue = query tiles + threshold terms
  const EmbQOut qo{out.thr4, out.kbar, gamma, beta};
  DAGL_CUDA_OK(cudaFuncSetAttribute(embed_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EB_SMQ_TOTAL));
  const int nwork_q = g.B * 2 * eg.nqt;                                                   // persistent: one CTA per SM
  DAGL_CUDA_OK(launch_pdl(embed_tc_kernel<true>, nwork_q < sms ? nwork_q : sms, EB_THREADS, EB_SMQ_TOTAL, st, g, eg, qimg, nullptr, w1, b1, absmax,
                          wmax1, Q, out.qtiles, nullptr, qo));
  DAGL_LAUNCH_CHECK();
  return 0;
}

}  // namespace dagl
