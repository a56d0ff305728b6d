// ResBlock chains on the tensor cores (tcgen05), sm_100a.
// Reference: common.ResBlock (DN_Gray/model/common.py:59-79): res = conv2(PReLU(conv1(x))) * res_scale + x, both convs
// Conv2d(64, 64, 3, padding 1, bias) (common.default_conv, common.py:8-11).  CES.RBS1 / RBS2 (dagl.py:86-101, 115, 117) are
// chains of four of them; RR.body holds sixteen more (dagl.py:27-34).
//
// One 3x3 convolution is an implicit GEMM with M = pixels, N = 64 output channels and K = 9 taps x 64 input channels.  The
// activations live in HBM as zero-padded, channel-last fp16 images, one for the hi and one for the lo part: a flat
// [NPG pixel records][64 ch = 128 B] array per image, pixels enumerated in padded-flat order (pitch Wp, pad 3: the frame of
// featmap_tc.cu / embed_tc.cu), SWIZZLE_128B pre-applied.  The A operand of (tap, 16-channel group) for 128 consecutive
// pixels is then a pixel- and channel-shifted view of a 144-record row segment held in smem, and a whole halo row (all 64
// channels of one part) arrives with ONE bulk copy of 18 KB: 6 copies per tile (a copy costs ~300 cycles to issue whatever
// its size; the first version used the 16-channel-group images of featmap_tc.cu, 24 copies per tile, and was bound by that).
// Tiles are walked in RUNS down 128-wide strips where the image width fills them (see CvGeom): inside a run every tile
// needs ONE new halo row, the ring keeps the other two (a third of the L2 -> SM traffic of independent tiles).
// The epilogue of a convolution writes the NEXT convolution's images directly (bias, PReLU or residual add fused), so inside
// a chain no fp32 activation is re-packed and no elementwise kernel runs; the fp32 NCHW tensor is only written at ResBlock
// outputs (it is the exact fp32 residual of the next block and the chain's result).
//
// fp32 accuracy: activations and weights are split into fp16 hi + lo (power-of-two pre-scaled);
// acc = x_hi.W_hi + (x_hi.W_lo + x_lo.W_hi), fp32 accumulate in TMEM, summed in fp32 in the epilogue.
// The fp16 scale of a convolution's output image is an a-priori bound from the MEASURED maximum of its input
// (|out| <= max|in| * max_co sum|w[co]| + max|bias|, PReLU / residual folded in); every epilogue measures the maximum of what
// it writes (atomicMax per warp), so the bound is loose by one layer's gain only, never by the product over the chain.
//
// Two kernels from one template; in both a CTA keeps the weights of 32 output channels (hi | lo, 72 KB) resident:
//   PAIR = true   a CTA pair (cta_group::2) owns two 128-pixel tiles, each CTA loads its own tile's halo; ONE M = 256,
//                 N = 128 MMA per (tap, channel group) covers both tiles and all 64 output channels: per SM half the
//                 operand smem reads of two independent CTAs.
//   PAIR = false  one CTA computes 32 output channels of one tile (even CTAs channels 0-31, odd CTAs 32-63).  A/B aid.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc_utils.cuh"

namespace dagl {
using namespace tc;

constexpr int CV_M = 128;                            // pixels per CTA tile
constexpr int CV_C = 64;                             // channels in and out
constexpr int CV_GROUPS = CV_C / 16;
constexpr int CV_HALF = 32;                          // output channels whose weights one CTA holds
constexpr int CV_SEG_PIX = 144;                      // 128 + 2 (kx) + 7 (alignment), rounded up to 8
constexpr int CV_REC = 128;                          // bytes per pixel record: 64 channels fp16
constexpr int CV_SEG_BYTES = CV_SEG_PIX * CV_REC;    // 18432 = 18 * 1024
constexpr int CV_STAGE_BYTES = 2 * CV_SEG_BYTES;     // one halo row: hi | lo
constexpr int CV_VTAPS = 9 * CV_GROUPS;              // 36 virtual taps of K = 16
constexpr int CV_WTAP_BYTES = 2 * CV_HALF * 32;      // 2048: K-major no-swizzle [2 k-chunks][64 rows: hi 32 | lo 32][16 B]
constexpr int CV_WHALF_BYTES = CV_VTAPS * CV_WTAP_BYTES;                 // 73728
constexpr int CV_STG_BYTES = 2 * CV_M * CV_REC;      // pair kernel: per epilogue group 128 output records (16 KB) staged for a bulk store
// smem: [halo ring: nst x (hi | lo) rows][weights of 32 output channels][pair / flat tiles: output staging][barriers, bias, slope]
// nst = 4 stages, or 3 stages + the 32 KB output staging of the pair kernel's bulk stores
constexpr int CV_SM_BAR = 4 * CV_STAGE_BYTES + CV_WHALF_BYTES;          // 221184 (>= 3 stages + weights + staging = 217088)
constexpr int CV_SM_PAR = CV_SM_BAR + 256;                              // bias [64] | slope [64]
constexpr int CV_SM_TOTAL = CV_SM_PAR + 512;
static_assert(CV_SEG_BYTES % 1024 == 0 && CV_STAGE_BYTES % 1024 == 0 && CV_WHALF_BYTES % 1024 == 0, "swizzle pattern alignment");
static_assert(3 * CV_STAGE_BYTES + CV_WHALF_BYTES + CV_STG_BYTES <= CV_SM_BAR, "staging fits below the barriers");
static_assert(CV_SM_TOTAL <= 227 * 1024, "smem budget");
constexpr int CV_THREADS = 320;                      // warp 0 loads, warp 1 issues (or relays), warps 2-5 / 6-9: two epilogue groups
constexpr int CV_PADK = 3;                           // frame of the packed images (shared with featmap_tc.cu / embed_tc.cu)
#ifndef CV_EXP
#define CV_EXP 0      // development experiments (wrong results): 1 no x_lo MMA, 2 no epilogue stores, 4 no halo copies, 8 no MMAs
#endif

// Work decomposition.  A RUN is a column of `RL` vertically adjacent tiles (tile i = 128 pixel slots starting at
// p0 + i Wp): tile i needs halo rows i, i+1, i+2 of the run's RL + 2 rows, so inside a run every tile loads ONE new row
// and the ring keeps the other two.
//   strip = 0 ("flat")  runs of one tile, tiles = consecutive 128-slot ranges of the padded-flat enumeration (every tile
//                        loads its three rows).  For narrow images (chop leaves: W = 72 -> 1.6 image rows per tile).
//   strip = 1            tile (y, c) = pixels (y, 128 c .. 128 c + 127); runs walk down a strip.  For images whose width
//                        fills the 128-wide strips (W/128 rounded up wastes < 20 %): a third of the L2 -> SM traffic.
struct CvGeom { int B, H, W, Wp, NkP, ntile, ntile2, NPG, Npix, strip, nstrip, RL, nyb; };
static CvGeom cv_geom(int B, int H, int W) {
  CvGeom e;
  e.B = B; e.H = H; e.W = W; e.Npix = H * W;
  e.Wp = (W + 2 * CV_PADK + 7) & ~7;
  e.NkP = (H - 1) * e.Wp + W;                        // pixel slot p = y Wp + x is record p + 3 Wp + 3 of the padded image
  e.ntile = (e.NkP + CV_M - 1) / CV_M;
  e.ntile2 = (e.ntile + 1) & ~1;                     // the pair kernel walks tile pairs
  const int np = (H + 2 * CV_PADK) * e.Wp;
  const int need = CV_M * e.ntile2 + 2 * CV_PADK * e.Wp + CV_SEG_PIX + 8;
  e.NPG = ((np > need ? np : need) + 7) & ~7;
  e.nstrip = (W + CV_M - 1) / CV_M;
  e.strip = (10 * W >= 8 * CV_M * e.nstrip && H >= 8) ? 1 : 0;
  e.RL = 1; e.nyb = H;                               // set per launch (cv_plan)
  return e;
}
// per-launch plan: rows per run (the ring depth and the output staging follow from the kernel instance: conv64_tc_kernel)
static void cv_plan(CvGeom& e, bool pair, int workers, int force_flat) {
  if (force_flat) e.strip = 0;
  if (!e.strip) {
    e.RL = 1; e.nyb = 0;
    return;
  }
  double best = 1e30;
  for (int rl = 2; rl <= 64 && rl <= e.H; ++rl) {
    const long long nruns = (long long)e.B * e.nstrip * ((e.H + rl - 1) / rl);
    const long long items = pair ? (nruns + 1) / 2 : nruns;
    const long long waves = (items + workers - 1) / workers;
    const double cost = (double)waves * (rl + 3.0);                     // + 3: the run's two extra rows and its pipeline fill
    if (cost <= best) { best = cost; e.RL = rl; }
  }
  e.nyb = (e.H + e.RL - 1) / e.RL;
}
static inline size_t cv_align(size_t x) { return (x + 255) & ~(size_t)255; }
static inline size_t cv_image_bytes(const CvGeom& e) { return cv_align((size_t)e.B * 2 * e.NPG * CV_REC); }

__device__ __forceinline__ float cv_pow2_scale(float a, int target) {
  if (!(a > 0.f) || !isfinite(a)) return 1.f;
  int e;
  frexpf(a, &e);
  return ldexpf(1.f, target - e);
}

// ---- weights ------------------------------------------------------------------------------------------------
// conv weight [64 co][64 ci][3][3] -> [half][virtual tap (group, tap)][2 k-chunks][64 rows][8 ch] fp16, rows = hi of output
// channels 32 half .. +31, then their lo parts: ONE MMA forms x_hi.W_hi (main columns) and x_hi.W_lo (cross columns).
// meta (floats, after the two halves): [max|w|, max_co sum|w[co]| (x 1.0001), max|bias|, 0]
constexpr size_t CV_WMETA_OFF = 2 * (size_t)CV_WHALF_BYTES;
constexpr size_t CV_WPACK_BYTES = CV_WMETA_OFF + 256;

__global__ void __launch_bounds__(256)
pack_conv_w_kernel(const float* __restrict__ w, const float* __restrict__ bias, uint8_t* __restrict__ out) {
  __shared__ float red[8], red2[8];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  float m = 0.f;
  for (int i = threadIdx.x; i < CV_C * CV_C * 9; i += 256) m = fmaxf(m, fabsf(__ldg(w + i)));
  float l1 = 0.f;
  for (int co = wp; co < CV_C; co += 8) {            // one warp per output channel
    float s = 0.f;
    for (int i = lane; i < CV_C * 9; i += 32) s += fabsf(__ldg(w + (size_t)co * CV_C * 9 + i));
    s = warp_sum(s);
    l1 = fmaxf(l1, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) { red[wp] = m; red2[wp] = l1; }
  __syncthreads();
  m = red[0]; l1 = red2[0];
#pragma unroll
  for (int k = 1; k < 8; ++k) { m = fmaxf(m, red[k]); l1 = fmaxf(l1, red2[k]); }
  if (threadIdx.x == 0) {
    float bm = 0.f;
    if (bias != nullptr)
      for (int c = 0; c < CV_C; ++c) bm = fmaxf(bm, fabsf(__ldg(bias + c)));
    float* meta = reinterpret_cast<float*>(out + CV_WMETA_OFF);
    meta[0] = m; meta[1] = l1 * 1.0001f; meta[2] = bm; meta[3] = 0.f;
  }
  const float scale = cv_pow2_scale(m, 14);
  for (int o = threadIdx.x; o < 2 * CV_VTAPS * 2 * CV_HALF; o += 256) {     // one 16-byte chunk = 8 input channels of one co
    const int e = o % CV_HALF, kc = (o / CV_HALF) & 1, v = (o / (2 * CV_HALF)) % CV_VTAPS, half = o / (2 * CV_HALF * CV_VTAPS);
    const int gq = v / 9, t = v % 9, co = half * CV_HALF + e;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int ci = gq * 16 + kc * 8 + 2 * j + u;
        x[u] = __ldg(w + ((size_t)co * CV_C + ci) * 9 + t) * scale;
      }
      const __half h0 = __float2half_rn(x[0]), h1 = __float2half_rn(x[1]);
      const __half l0 = __float2half_rn(x[0] - __half2float(h0)), l1h = __float2half_rn(x[1] - __half2float(h1));
      hi[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      lo[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1h) << 16);
    }
    uint8_t* base = out + (size_t)half * CV_WHALF_BYTES + (size_t)v * CV_WTAP_BYTES + (size_t)kc * (2 * CV_HALF) * 16;
    *reinterpret_cast<uint4*>(base + (size_t)e * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + (size_t)(CV_HALF + e) * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ---- input of a chain: fp32 NCHW -> packed image ----------------------------------------------------------------
__global__ void cv_absmax_kernel(const float* __restrict__ x, size_t n_per_img, unsigned* __restrict__ amax) {
  pdl_prologue();
  const int img = blockIdx.y;
  const float* xi = x + (size_t)img * n_per_img;
  float m = 0.f;
  if ((n_per_img & 3) == 0) {
    const float4* x4 = reinterpret_cast<const float4*>(xi);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_img / 4; i += (size_t)gridDim.x * blockDim.x) {
      const float4 v = __ldg(x4 + i);
      m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_img; i += (size_t)gridDim.x * blockDim.x)
      m = fmaxf(m, fabsf(__ldg(xi + i)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(amax + img, __float_as_uint(m));
}

// layout: [img][hi|lo][NPG records][128 B = 64 ch], SWIZZLE_128B pre-applied: 16-byte chunk c of record r sits at c ^ (r & 7)
// One thread per (pixel record, 16-channel group).  (One thread per whole record — 64 loads, two contiguous 128-byte stores —
// measured slower: 17 vs 15 us at 256², 88 vs 48 us at 64x72²: a quarter of the threads.)  Border records are written as
// zeros; with `zb0` / `zb1` (strip tiles, whose epilogues write pixels only) the same border records of the two rotating
// output images are zeroed as well.
__global__ void __launch_bounds__(256)
cv_pack_kernel(CvGeom eg, const float* __restrict__ x, const unsigned* __restrict__ amax, uint8_t* __restrict__ img_out,
               uint8_t* __restrict__ zb0, uint8_t* __restrict__ zb1) {
  pdl_prologue();
  const int pblocks = (eg.NPG + 255) / 256;
  const int pb = blockIdx.x;
  const int img = pb / (pblocks * CV_GROUPS), gq = (pb / pblocks) % CV_GROUPS;
  const int pix = (pb % pblocks) * 256 + threadIdx.x;
  if (pix >= eg.NPG) return;
  const float scale = cv_pow2_scale(__uint_as_float(amax[img]), 14);
  const int r = pix / eg.Wp, cc = pix % eg.Wp;
  const int y = r - CV_PADK, xx = cc - CV_PADK;
  const bool inb = (y >= 0 && y < eg.H && xx >= 0 && xx < eg.W);
  const float* src = x + (((size_t)img * CV_C + gq * 16) * eg.H + (inb ? y : 0)) * eg.W + (inb ? xx : 0);
  uint32_t h[8], l[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float a0 = 0.f, a1 = 0.f;
    if (inb) {
      a0 = __ldg(src + (size_t)(2 * j) * eg.Npix) * scale;
      a1 = __ldg(src + (size_t)(2 * j + 1) * eg.Npix) * scale;
    }
    __half2 hh = __floats2half2_rn(a0, a1);
    const float2 f = __half22float2(hh);
    __half2 ll = __floats2half2_rn(a0 - f.x, a1 - f.y);
    h[j] = *reinterpret_cast<uint32_t*>(&hh);
    l[j] = *reinterpret_cast<uint32_t*>(&ll);
  }
  const int k0 = (2 * gq) ^ (pix & 7), k1 = k0 ^ 1;
  const size_t rec_off = ((size_t)img * 2) * (size_t)eg.NPG * CV_REC + (size_t)pix * CV_REC;
  uint4* dh = reinterpret_cast<uint4*>(img_out + rec_off);
  uint4* dl = reinterpret_cast<uint4*>(img_out + rec_off + (size_t)eg.NPG * CV_REC);
  dh[k0] = make_uint4(h[0], h[1], h[2], h[3]);
  dh[k1] = make_uint4(h[4], h[5], h[6], h[7]);
  dl[k0] = make_uint4(l[0], l[1], l[2], l[3]);
  dl[k1] = make_uint4(l[4], l[5], l[6], l[7]);
  if (!inb && zb0 != nullptr) {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    uint4* a0 = reinterpret_cast<uint4*>(zb0 + rec_off);
    uint4* a1 = reinterpret_cast<uint4*>(zb0 + rec_off + (size_t)eg.NPG * CV_REC);
    uint4* b0 = reinterpret_cast<uint4*>(zb1 + rec_off);
    uint4* b1 = reinterpret_cast<uint4*>(zb1 + rec_off + (size_t)eg.NPG * CV_REC);
    a0[k0] = z; a0[k1] = z; a1[k0] = z; a1[k1] = z;
    b0[k0] = z; b0[k1] = z; b1[k0] = z; b1[k1] = z;
  }
}

// ---- the convolution ------------------------------------------------------------------------------------------
struct CvArgs {
  const uint8_t* in_img;       // packed input image
  const unsigned* in_amax;     // [B] float bits: the bound the input image was scaled with (scale = pow2(14 - exp))
  const unsigned* in_meas;     // [B] float bits: measured max |input| (<= the bound); drives the bound of the output
  const uint8_t* wpack;        // pack_conv_w_kernel image
  const float* bias;           // [64], nullable
  const float* slope;          // PReLU weight, nullable (no activation)
  int slope_n;                 // 1 or 64
  float res_scale;             // out = conv * res_scale + res   (only with res)
  const float* res;            // fp32 NCHW residual, nullable
  const unsigned* res_meas;    // [B] float bits: max |res|
  float* out;                  // fp32 NCHW, nullable
  uint8_t* out_img;            // packed output image, nullable
  unsigned* out_amax;          // [B] written: the bound out_img is scaled with
  unsigned* out_meas;          // [B] atomicMax of |out| (zeroed by the launcher)
};

// tcgen05 wrappers of the pair (cta_group::2) forms
__device__ __forceinline__ void mma_f16_ss_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this offset in BOTH CTAs of the pair when all previously issued MMAs have completed
__device__ __forceinline__ void mma_commit_2cta(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_dst) {   // one full warp of EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(NCOLS) : "memory");
}

// bulk async copy shared -> global (TMA, 1-D) and its completion (issuing thread only)
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Persistent.  Accumulator set s (two sets alternate; epilogue group s drains set s, so the epilogue of a tile overlaps the
// MMAs and the epilogue of the next):
//   PAIR   columns s*128 + [0,128): x_hi . [W_hi | W_lo] = [main | cross] of CTA 0's channels (0-31), then of CTA 1's (32-63);
//          x_lo . W_hi (N = 64: channels 0-31 from CTA 0, 32-63 from CTA 1) accumulates onto columns [32,96), i.e. onto the
//          cross columns of channels 0-31 and the main columns of 32-63: the epilogue adds main + cross per channel anyway
//   !PAIR  columns s*64 + [0,64): [main | cross] of this CTA's 32 channels; x_lo . W_hi accumulates onto the cross columns
// Work: CTA (or pair) k of n walks items k, k + n, ...; item = one run (pair: two runs, CTA r takes run 2 i + r); see CvGeom.
// Ring: halo row j of a run sits at ring position n0 + j (stage = position % nst); tile i reads positions n0 + i .. + 2 and
// frees position n0 + i (the last tile of a run frees all three).
// Everything that does not depend on the previous kernel of the stream (barriers, TMEM, the weights, bias, slope) is set up
// BEFORE griddepcontrol.wait, i.e. while the previous convolution is still running (programmatic dependent launch).
// STRIP: runs down 128-wide strips (4 ring stages, direct stores) / flat tiles (runs of one; pair: 3 stages + output staging):
// compile-time so that the flat kernels carry none of the run logic (their single-thread loops are the critical path)
template <bool PAIR, bool STRIP>
__global__ void __launch_bounds__(CV_THREADS, 1) conv64_tc_kernel(CvGeom eg, CvArgs a) {
  constexpr int MAXST = 4;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CV_SM_BAR);
  uint64_t* w_full = bars + 0;                    // weights resident
  uint64_t* a_full = bars + 1;                    // [nst] halo row (hi | lo) resident
  uint64_t* a_empty = a_full + MAXST;             // [nst] the MMAs that read it have completed
  uint64_t* pa_full = a_empty + MAXST;            // [nst] pair leader: the peer's halo row is resident
  uint64_t* d_full = pa_full + MAXST;             // [2] accumulator set complete
  uint64_t* d_empty = d_full + 2;                 // [2] accumulator set drained: 4 arrivals (epilogue warps); pair leader: 8
  uint64_t* pw_full = d_empty + 2;                // pair leader: the peer's weights are resident
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pw_full + 1);
  float* bound_s = reinterpret_cast<float*>(pw_full + 2);    // [0] max(1, max|slope|)
  float* bias_s = reinterpret_cast<float*>(smem + CV_SM_PAR);  // [64]
  float* slope_s = bias_s + CV_C;                              // [64]

  constexpr uint32_t TCOLS = PAIR ? 256 : 128;
  constexpr uint32_t SET = PAIR ? 128 : 64, D2OFF = 32;
  constexpr uint32_t nst = (PAIR && !STRIP) ? 3u : 4u;
  constexpr int sm_w = (int)nst * CV_STAGE_BYTES, sm_stg = sm_w + CV_WHALF_BYTES;
  const int warp = warp_id_uniform();
  const int tid = threadIdx.x;
  const uint32_t rank = PAIR ? cluster_ctarank() : (blockIdx.x & 1u);       // = the channel half whose weights this CTA holds
  const int RL = STRIP ? eg.RL : 1;
  const int nruns = STRIP ? eg.B * eg.nstrip * eg.nyb : (PAIR ? eg.B * eg.ntile2 : eg.B * eg.ntile);
  const int nwork = PAIR ? (nruns + 1) / 2 : nruns;
  const int w0 = (int)(blockIdx.x >> 1), wstep = (int)(gridDim.x >> 1);
  // item -> this CTA's run: image, first pixel slot p0, first image row y0 (strip), dummy = nothing to write
  auto decode = [&](int w, int& img, int& p0, int& y0, bool& dummy) {
    int q = PAIR ? 2 * w + (int)rank : w;
    dummy = q >= nruns;
    if (dummy) q = nruns - 1;
    if (STRIP) {
      const int per_img = eg.nstrip * eg.nyb;
      img = q / per_img;
      const int c = (q / eg.nyb) % eg.nstrip;
      y0 = (q % eg.nyb) * RL;
      p0 = y0 * eg.Wp + c * CV_M;
    } else {
      const int nt = PAIR ? eg.ntile2 : eg.ntile;
      img = q / nt; p0 = (q % nt) * CV_M; y0 = 0;
    }
  };

  if (tid == 0) {
    mbar_init(w_full, 1);
    mbar_init(pw_full, 1);
    for (int i = 0; i < MAXST; ++i) { mbar_init(a_full + i, 1); mbar_init(a_empty + i, 1); mbar_init(pa_full + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(d_full + i, 1); mbar_init(d_empty + i, PAIR ? 8 : 4); }
    mbar_init_fence();
    // the weights are constant inputs: fetch them before waiting for the previous kernel
    mbar_arrive_expect_tx(w_full, CV_WHALF_BYTES);
    bulk_g2s(smem + sm_w, a.wpack + (size_t)rank * CV_WHALF_BYTES, CV_WHALF_BYTES, w_full);
    float sm = 1.f;
    if (a.slope != nullptr)
      for (int i = 0; i < a.slope_n; ++i) sm = fmaxf(sm, fabsf(__ldg(a.slope + i)));
    bound_s[0] = sm;
  }
  if (tid >= 64 && tid < 64 + CV_C) {
    const int c = tid - 64;
    bias_s[c] = a.bias != nullptr ? __ldg(a.bias + c) : 0.f;
    slope_s[c] = a.slope != nullptr ? __ldg(a.slope + (a.slope_n > 1 ? c : 0)) : 1.f;
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_2cta<TCOLS>(tmem_ptr); else tmem_alloc<TCOLS>(tmem_ptr);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ===================== producer: lanes 0 / 1 copy the hi / lo part of a halo row =========
    const int lane = tid & 31;
    // ring position -> (stage, phase) kept incrementally: these single-thread loops are the critical path of the kernel and a
    // division by the run-time stage count costs more than the copy it guards
    uint32_t s = 0, ph = 0;
    for (int w = w0; w < nwork; w += wstep) {
      int img, p0, y0; bool dummy;
      decode(w, img, p0, y0, dummy);
#pragma unroll 1
      for (int j = 0; j < RL + 2; ++j) {
        if (lane == 0) {
          mbar_wait(a_empty + s, ph ^ 1u);
          mbar_arrive_expect_tx(a_full + s, CV_STAGE_BYTES);
        }
        __syncwarp();
        if (lane < 2) {
          const uint8_t* src = a.in_img + ((size_t)img * 2 + lane) * (size_t)eg.NPG * CV_REC;
          // 3x3 / pad 1 inside the pad-3 frame: halo row j of the run = frame row (first image row) + j + 2; rows below the
          // image (runs padded to RL tiles) are clamped to a border row: loaded, never used by a valid pixel
          int start = p0 + (j + 2) * eg.Wp;
          if (STRIP && y0 + j + 2 > eg.H + 4) start = p0 + (eg.H + 4 - y0) * eg.Wp;
          const int first = start & ~7;
          bulk_g2s(smem + s * CV_STAGE_BYTES + lane * CV_SEG_BYTES, src + (size_t)first * CV_REC, CV_SEG_BYTES, a_full + s);
        }
        __syncwarp();
        if (++s == nst) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (PAIR && rank == 1) {
      // ===================== peer relay: tell the leader when this CTA's operands are resident ==========
      if (elect_one()) {
        const uint32_t l_pw = mapa(smem_u32(pw_full), 0);
        uint32_t s = 0, ph = 0;
        bool first = true;
        for (int w = w0; w < nwork; w += wstep) {
          if (first) { mbar_wait(w_full, 0); fence_proxy_async_all(); mbar_arrive_cluster(l_pw); first = false; }
          for (int j = 0; j < RL + 2; ++j) {
            mbar_wait(a_full + s, ph);
            fence_proxy_async_all();
            mbar_arrive_cluster(mapa(smem_u32(pa_full + s), 0));
            if (++s == nst) { s = 0; ph ^= 1u; }
          }
        }
      }
    } else if (elect_one()) {
      // ===================== MMA issuer =====================
      const uint32_t abase = smem_u32(smem), wbase = smem_u32(smem + sm_w);
      constexpr uint32_t id1 = PAIR ? instr_desc(256, 128, FMT_F16, FMT_F16, 0, 0) : instr_desc(CV_M, 64, FMT_F16, FMT_F16, 0, 0);
      constexpr uint32_t id2 = PAIR ? instr_desc(256, 64, FMT_F16, FMT_F16, 0, 0) : instr_desc(CV_M, 32, FMT_F16, FMT_F16, 0, 0);
      uint32_t sb = 0;                                     // stage of the current tile's first row
      uint32_t sw_ = 0, phw = 0;                           // next row to wait for: stage, phase
      int ahead = 0;                                       // rows already waited for beyond the current tile's first row
      int it = 0;                                          // tile counter: accumulator set it & 1
      for (int w = w0; w < nwork; w += wstep) {
        int img, p0, y0; bool dummy;
        decode(w, img, p0, y0, dummy);
        // the leader's run fixes the pixel shift; the peer's run starts a multiple of 8 slots away: same (p0 & 7)
        const int off = (p0 & 7) + 2;
        if (it == 0) {
          mbar_wait(w_full, 0);
          if (PAIR) mbar_wait_cluster(pw_full, 0);
        }
#pragma unroll 1
        for (int i = 0; i < RL; ++i, ++it) {
          const int ab = it & 1;
          const uint32_t d1 = tbase + ab * SET, d2 = d1 + D2OFF;
          mbar_wait(d_empty + ab, ((uint32_t)(it >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const bool last = i == RL - 1;
#pragma unroll 1
          for (int ky = 0; ky < 3; ++ky) {
            uint32_t s = sb + (uint32_t)ky;
            if (s >= nst) s -= nst;
            // rows are waited for just before their first use: inside a run only row ky = 2 is new, and the MMAs of rows
            // 0 and 1 run while it lands
            if (ahead <= ky) {
              mbar_wait(a_full + sw_, phw);
              if (PAIR) mbar_wait_cluster(pa_full + sw_, phw);
              if (++sw_ == nst) { sw_ = 0; phw ^= 1u; }
              ++ahead;
              tc_fence_after();
            }
            const uint32_t row_hi = abase + s * CV_STAGE_BYTES + off * CV_REC, row_lo = row_hi + CV_SEG_BYTES;
#pragma unroll
            for (int gq = 0; gq < CV_GROUPS; ++gq) {
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                // A: K-major SWIZZLE_128B, rows (pixels) 128 B apart, 8-row groups 1024 B apart; the 16-channel group is a
                // 32-byte step inside the row, the tap a whole-record step (the swizzle works on the address bits)
                const uint32_t a_hi = row_hi + kx * CV_REC + gq * 32, a_lo = row_lo + kx * CV_REC + gq * 32;
                const uint64_t da_hi = (uint64_t)((a_hi >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
                                       ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
                const uint64_t da_lo = (uint64_t)((a_lo >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
                                       ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
                const uint64_t db = smem_desc(wbase + (gq * 9 + ky * 3 + kx) * CV_WTAP_BYTES, 2 * CV_HALF * 16, 128);   // rows: hi 0-31 | lo 32-63
                const uint32_t acc = (ky | gq | kx) ? 1u : 0u;
                if (CV_EXP & 8) continue;
                if (PAIR) {
                  mma_f16_ss_2cta(d1, da_hi, db, id1, acc);          // x_hi . [W_hi | W_lo]   (each CTA supplies its 64 rows)
                  if (!(CV_EXP & 1)) mma_f16_ss_2cta(d2, da_lo, db, id2, 1u);           // x_lo . W_hi            (each CTA supplies its first 32 rows)
                } else {
                  mma_f16_ss(d1, da_hi, db, id1, acc);
                  if (!(CV_EXP & 1)) mma_f16_ss(d2, da_lo, db, id2, 1u);
                }
              }
            }
            // row ky of this tile is row ky - 1 of the next tile of the run: only the tile's first row is done with
            // (all three after the run's last tile)
            if (ky == 0 || last) {
              if (PAIR) mma_commit_2cta(a_empty + s); else mma_commit(a_empty + s);
            }
          }
          if (PAIR) mma_commit_2cta(d_full + ab); else mma_commit(d_full + ab);
          // next tile: inside a run one row further (two of its rows are already here), after the run three
          const uint32_t adv = last ? 3u : 1u;
          sb += adv; if (sb >= nst) sb -= nst;
          ahead -= (int)adv;
        }
      }
    }
  } else {
    // ===================== epilogue: two groups of 4 warps, group g takes the tiles with (it & 1) == g; thread = pixel slot ======
    const int grp = (warp - 2) >> 2, quad = warp & 3, lane = tid & 31;
    const int r = quad * 32 + lane;
    const float* meta = reinterpret_cast<const float*>(a.wpack + CV_WMETA_OFF);
    const float s_w = cv_pow2_scale(meta[0], 14);
    const uint32_t l_de = PAIR ? mapa(smem_u32(d_empty + grp), 0) : 0u;
    constexpr int NCHUNK = PAIR ? 4 : 2;                       // 16 output channels per chunk
    const int cbase = PAIR ? 0 : (int)rank * CV_HALF;          // first output channel this CTA writes
    const bool staged = PAIR && !STRIP && a.out_img != nullptr;     // records via smem + bulk stores
    uint8_t* stg = smem + sm_stg + grp * (CV_M * CV_REC);      // this group's staging buffer
    int cur_img = -1;
    float inv = 0.f, s_out = 0.f, bound_cur = 0.f;
    int it = 0;
    for (int w = w0; w < nwork; w += wstep) {
      int img, p0, y0; bool dummy;
      decode(w, img, p0, y0, dummy);
#pragma unroll 1
      for (int i = 0; i < RL; ++i, ++it) {
      if ((it & 1) != grp) continue;
      const int p = p0 + i * eg.Wp + r;
      const int y = p / eg.Wp, x = p % eg.Wp;
      // strip tiles: slots past the row end belong to the next row's strip 0 and rows past the image to nobody
      const bool covered = !dummy && (!STRIP || ((p0 % eg.Wp) + r < eg.Wp && y0 + i < eg.H));
      const bool valid = covered && (p < eg.NkP) && (x < eg.W);
      // flat tiles write zero records for their dummy slots (the borders between the rows); strip tiles only their pixels
      // (cv_pack_kernel zeroes the rest once per call)
      const bool wr_rec = STRIP ? valid : !dummy;
      const size_t pix = (size_t)y * eg.W + x;
      if (img != cur_img) {
        cur_img = img;
        const float in_bound = __uint_as_float(a.in_amax[img]);
        inv = 1.f / (s_w * cv_pow2_scale(in_bound, 14));
        // a-priori bound of what this launch writes, from the measured maximum of its input
        float bound = __uint_as_float(a.in_meas[img]) * meta[1] + meta[2];
        bound *= bound_s[0];
        if (a.res != nullptr) bound = bound * fabsf(a.res_scale) + __uint_as_float(a.res_meas[img]);
        bound *= 1.0001f;
        s_out = cv_pow2_scale(bound, 14);
        bound_cur = bound;
      }
      // one thread per image (pixel slot 0; the single-CTA kernel visits it once per channel half) publishes the bound
      if (a.out_amax != nullptr && p == 0 && !dummy && (PAIR || rank == 0)) a.out_amax[img] = __float_as_uint(bound_cur);
      const uint32_t trow = tbase + ((uint32_t)(quad * 32) << 16) + grp * SET;
      const size_t rec = (size_t)p + 3 * eg.Wp + 3;
      const int sw = (int)rec & 7;
      uint8_t* ihi = a.out_img + ((size_t)img * 2) * (size_t)eg.NPG * CV_REC;
      uint8_t* ilo = ihi + (size_t)eg.NPG * CV_REC;
      const float* rp = a.res + ((size_t)img * CV_C + cbase) * eg.Npix + pix;
      float* op = a.out + ((size_t)img * CV_C + cbase) * eg.Npix + pix;
      // v[]: first the residual (it does not depend on the accumulator: fetched before waiting for the MMAs), then the results
      float v[NCHUNK * 16];
      if (a.res != nullptr && valid) {
#pragma unroll
        for (int c = 0; c < NCHUNK * 16; ++c) v[c] = __ldg(rp + (size_t)c * eg.Npix);
      } else {
#pragma unroll
        for (int c = 0; c < NCHUNK * 16; ++c) v[c] = 0.f;
      }
      mbar_wait(d_full + grp, (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      // TMEM -> registers with nothing but arithmetic in between: the accumulator set is handed back to the MMA issuer as
      // early as possible (it is the resource the two epilogue groups and the issuer rotate over)
#pragma unroll
      for (int ch = 0; ch < NCHUNK; ++ch) {
        const int hh = ch >> 1, cc0 = (ch & 1) * 16, cl = ch * 16;
        uint32_t vm[16], vc[16];
        tmem_ld16(trow + (PAIR ? hh * 64 : 0) + cc0, vm);
        tmem_ld16(trow + (PAIR ? hh * 64 : 0) + 32 + cc0, vc);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          float t = (__uint_as_float(vm[c]) + __uint_as_float(vc[c])) * inv + bias_s[cbase + cl + c];
          if (a.slope != nullptr) t = t > 0.f ? t : t * slope_s[cbase + cl + c];
          if (a.res != nullptr) t = t * a.res_scale + v[cl + c];
          v[cl + c] = valid ? t : 0.f;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(l_de); else mbar_arrive(d_empty + grp);
      }
      float vmax = 0.f;
#pragma unroll
      for (int c = 0; c < NCHUNK * 16; ++c) vmax = fmaxf(vmax, fabsf(v[c]));
      if (a.out != nullptr && valid && !(CV_EXP & 2)) {
#pragma unroll
        for (int c = 0; c < NCHUNK * 16; ++c) op[(size_t)c * eg.Npix] = v[c];
      }
      if (a.out_img != nullptr && !(CV_EXP & 2)) {
        // Flat pair tiles: the tile's 128 records are contiguous in the image, so they go through a 16 KB staging buffer and
        // ONE bulk store per part (hi, then lo through the same buffer) instead of 1024 16-byte stores.
        if (staged) {                                          // the previous item's lo store has read the staging buffer
          if (r == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          named_bar_sync(1 + grp, 128);
        }
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          uint4* dst = reinterpret_cast<uint4*>(staged ? stg + r * CV_REC : (part ? ilo : ihi) + rec * CV_REC);
          if (staged || wr_rec) {
#pragma unroll
            for (int ch = 0; ch < NCHUNK; ++ch) {
              uint32_t g[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float a0 = v[ch * 16 + 2 * j] * s_out, a1 = v[ch * 16 + 2 * j + 1] * s_out;
                __half2 h = __floats2half2_rn(a0, a1);                           // one packed conversion (F2FP), not two F2F
                if (part) { const float2 f = __half22float2(h); h = __floats2half2_rn(a0 - f.x, a1 - f.y); }
                g[j] = *reinterpret_cast<uint32_t*>(&h);
              }
              const int k0 = ((cbase + ch * 16) >> 3) ^ sw;                        // swizzled 16-byte chunk positions
              dst[k0] = make_uint4(g[0], g[1], g[2], g[3]);
              dst[k0 ^ 1] = make_uint4(g[4], g[5], g[6], g[7]);
            }
          }
          if (staged) {
            fence_async_smem();
            named_bar_sync(1 + grp, 128);
            if (r == 0 && !dummy) {
              bulk_s2g((part ? ilo : ihi) + rec * CV_REC, stg, CV_M * CV_REC);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              if (part == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            if (part == 0) named_bar_sync(1 + grp, 128);
          }
        }
      }
      if (!STRIP && !dummy && a.out_img != nullptr && (p0 == 0 || p0 == (eg.ntile - 1) * CV_M)) {
        // flat tiles: head (records before the first slot) and tail (after the last slot) are zero borders too: this CTA's
        // channels (the 16-byte chunks cbase/8 .. of every record; the swizzle permutes them within the record's half)
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        auto zero_range = [&](size_t i0, size_t i1) {
          for (size_t q = i0 + r; q < i1; q += CV_M)
#pragma unroll
            for (int k = 0; k < NCHUNK * 2; ++k) {
              const int kk = ((cbase >> 3) + k) ^ ((int)q & 7);
              reinterpret_cast<uint4*>(ihi + q * CV_REC)[kk] = z;
              reinterpret_cast<uint4*>(ilo + q * CV_REC)[kk] = z;
            }
        };
        if (p0 == 0) zero_range(0, (size_t)3 * eg.Wp + 3);
        if (p0 == (eg.ntile - 1) * CV_M) zero_range((size_t)eg.ntile * CV_M + 3 * eg.Wp + 3, (size_t)eg.NPG);
      }
      if (a.out_meas != nullptr) {
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o2));
        if (lane == 0 && vmax > 0.f) atomicMax(a.out_meas + img, __float_as_uint(vmax));
      }
      }
    }
    if (staged && r == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // smem is read / the stores are done before the CTA exits
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    if (PAIR) tmem_dealloc_2cta<TCOLS>(tbase); else tmem_dealloc<TCOLS>(tbase);
  }
}

// ---- host side -----------------------------------------------------------------------------
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl_cluster2(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2] = {};
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

size_t resblock_packed_weights_bytes() { return 2 * cv_align(CV_WPACK_BYTES); }     // conv1 | conv2

int launch_pack_resblock_weights(const float* w1, const float* b1, const float* w2, const float* b2, void* packed,
                                 size_t packed_bytes, cudaStream_t st) {
  if (packed_bytes < resblock_packed_weights_bytes()) {
    call_state().err = "packed ResBlock weights buffer too small";
    return -3;
  }
  uint8_t* p = static_cast<uint8_t*>(packed);
  pack_conv_w_kernel<<<1, 256, 0, st>>>(w1, b1, p);
  DAGL_LAUNCH_CHECK();
  pack_conv_w_kernel<<<1, 256, 0, st>>>(w2, b2, p + cv_align(CV_WPACK_BYTES));
  DAGL_LAUNCH_CHECK();
  return 0;
}

// workspace of a chain: 3 packed images (block input / conv1 output, rotating), 2 fp32 tensors (residuals between blocks),
// the max slots and (unless the caller pre-packed) the packed weights of every block
struct CvWs { size_t img[3], f32[2], slots, packw, total; };
static CvWs cv_ws(const CvGeom& e, int nblocks) {
  CvWs L;
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off += cv_align(b); return o; };
  for (int i = 0; i < 3; ++i) L.img[i] = take(cv_image_bytes(e));
  for (int i = 0; i < 2; ++i) L.f32[i] = take(nblocks > 1 ? (size_t)e.B * CV_C * e.Npix * sizeof(float) : 0);
  L.slots = take((size_t)(2 * nblocks + 1) * 2 * e.B * sizeof(unsigned));
  L.packw = take((size_t)nblocks * resblock_packed_weights_bytes());
  L.total = off;
  return L;
}
size_t resblocks_workspace_bytes(int B, int H, int W, int nblocks) { return cv_ws(cv_geom(B, H, W), nblocks).total; }

static int cv_sm_count(int* sms) {
  int dev = 0;
  *sms = 148;
  DAGL_CUDA_OK(cudaGetDevice(&dev));
  DAGL_CUDA_OK(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
  return 0;
}

static int launch_conv(const CvGeom& e, const CvArgs& a, bool pair, int sms, cudaStream_t st) {
  // grid = 2 x (CTA pairs): in both kernels CTA 2k+r holds the weights of channel half r and walks items k, k + n, ...
  const int nruns = e.strip ? e.B * e.nstrip * e.nyb : (pair ? e.B * e.ntile2 : e.B * e.ntile);
  const int nwork = pair ? (nruns + 1) / 2 : nruns;
  const int ncl = nwork < sms / 2 ? nwork : sms / 2;
  if (pair) {
    auto kern = e.strip ? conv64_tc_kernel<true, true> : conv64_tc_kernel<true, false>;
    DAGL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CV_SM_TOTAL));
    DAGL_CUDA_OK(launch_pdl_cluster2(kern, dim3(2 * ncl), CV_THREADS, CV_SM_TOTAL, st, e, a));
  } else {
    auto kern = e.strip ? conv64_tc_kernel<false, true> : conv64_tc_kernel<false, false>;
    DAGL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CV_SM_TOTAL));
    DAGL_CUDA_OK(launch_pdl(kern, dim3(2 * ncl), CV_THREADS, CV_SM_TOTAL, st, e, a));
  }
  DAGL_LAUNCH_CHECK();
  return 0;
}

// x, y: fp32 [B][64][H][W] (y may alias x).  blocks[i]: the parameters of ResBlock i; packed[i] (nullable): its weights
// pre-packed by launch_pack_resblock_weights.  mode: 0 = CTA-pair kernel, 1 = single-CTA kernel, 2 = by size.
int launch_resblocks(int B, int H, int W, const float* x, float* y, int nblocks, const ResBlockParams* blocks, void* ws,
                     size_t ws_bytes, int mode, cudaStream_t st) {
  CvGeom e = cv_geom(B, H, W);
  const CvWs L = cv_ws(e, nblocks);
  if (ws_bytes < L.total) {
    call_state().err = "ResBlock chain: workspace too small";
    return -3;
  }
  int sms = 148;
  if (cv_sm_count(&sms)) return -4;
  char* base = static_cast<char*>(ws);
  uint8_t* img[3] = {reinterpret_cast<uint8_t*>(base + L.img[0]), reinterpret_cast<uint8_t*>(base + L.img[1]),
                     reinterpret_cast<uint8_t*>(base + L.img[2])};
  float* f32[2] = {reinterpret_cast<float*>(base + L.f32[0]), reinterpret_cast<float*>(base + L.f32[1])};
  unsigned* slots = reinterpret_cast<unsigned*>(base + L.slots);    // layer l: bound [l][B], measured [nl + l][B]
  const int nl = 2 * nblocks + 1;
  auto bound_of = [&](int l) { return slots + (size_t)l * B; };
  auto meas_of = [&](int l) { return slots + (size_t)(nl + l) * B; };
  // auto: the pair kernel once there are enough tiles for ~7 items per CTA pair (measured: 512^2 and the chop batches;
  // below that the single-CTA kernel's shorter start-up wins)
  const bool pair = mode == 0 || (mode == 2 && (long long)B * e.ntile >= 1024);
  static const int force_flat = getenv("DAGL_CONV_FLAT") != nullptr ? atoi(getenv("DAGL_CONV_FLAT")) : 0;   // A/B aid
  cv_plan(e, pair, sms / 2, force_flat);

  DAGL_CUDA_OK(cudaMemsetAsync(slots, 0, (size_t)nl * 2 * B * sizeof(unsigned), st));
  for (int i = 0; i < nblocks; ++i)
    if (blocks[i].packed == nullptr) {
      const int rc = launch_pack_resblock_weights(blocks[i].w1, blocks[i].b1, blocks[i].w2, blocks[i].b2,
                                                  base + L.packw + (size_t)i * resblock_packed_weights_bytes(),
                                                  resblock_packed_weights_bytes(), st);
      if (rc) return rc;
    }
  // layer 0 = the chain input: its measured maximum is also the bound its image is scaled with
  const size_t n_img = (size_t)CV_C * e.Npix;
  DAGL_CUDA_OK(launch_pdl(cv_absmax_kernel, dim3(128, B), 256, 0, st, x, n_img, meas_of(0)));
  DAGL_LAUNCH_CHECK();
  // strip tiles write pixels only: the pack launch also zeroes the border records of the two rotating output images
  DAGL_CUDA_OK(launch_pdl(cv_pack_kernel, dim3(((e.NPG + 255) / 256) * CV_GROUPS * B), 256, 0, st, e, x, (const unsigned*)meas_of(0), img[0],
                          e.strip ? img[1] : (uint8_t*)nullptr, e.strip ? img[2] : (uint8_t*)nullptr));
  DAGL_LAUNCH_CHECK();

  int cur = 0;                                   // packed image holding the current block's input
  const unsigned* cur_bound = meas_of(0);
  const float* res = x;
  for (int i = 0; i < nblocks; ++i) {
    const ResBlockParams& rb = blocks[i];
    const uint8_t* pw = rb.packed != nullptr ? static_cast<const uint8_t*>(rb.packed)
                                             : reinterpret_cast<const uint8_t*>(base + L.packw + (size_t)i * resblock_packed_weights_bytes());
    const int mid = (cur + 1) % 3, nxt = (cur + 2) % 3;
    const bool last = i == nblocks - 1;
    CvArgs c1{};
    c1.in_img = img[cur]; c1.in_amax = cur_bound; c1.in_meas = meas_of(2 * i);
    c1.wpack = pw; c1.bias = rb.b1; c1.slope = rb.slope; c1.slope_n = rb.slope_n;
    c1.res_scale = 1.f; c1.res = nullptr; c1.res_meas = nullptr; c1.out = nullptr;
    c1.out_img = img[mid]; c1.out_amax = bound_of(2 * i + 1); c1.out_meas = meas_of(2 * i + 1);
    int rc = launch_conv(e, c1, pair, sms, st);
    if (rc) return rc;
    CvArgs c2{};
    c2.in_img = img[mid]; c2.in_amax = bound_of(2 * i + 1); c2.in_meas = meas_of(2 * i + 1);
    c2.wpack = pw + cv_align(CV_WPACK_BYTES); c2.bias = rb.b2; c2.slope = nullptr; c2.slope_n = 0;
    c2.res_scale = rb.res_scale; c2.res = res; c2.res_meas = meas_of(2 * i);
    c2.out = last ? y : f32[i & 1];
    c2.out_img = last ? nullptr : img[nxt]; c2.out_amax = last ? nullptr : bound_of(2 * i + 2); c2.out_meas = last ? nullptr : meas_of(2 * i + 2);
    rc = launch_conv(e, c2, pair, sms, st);
    if (rc) return rc;
    res = c2.out; cur = nxt; cur_bound = bound_of(2 * i + 2);
  }
  return 0;
}

}  // namespace dagl
