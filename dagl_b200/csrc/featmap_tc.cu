// Feature maps on the tensor cores (tcgen05), sm_100a.
// Reference: b1 = g(b) (Conv2d 64->16, 3x3, pad 1) and b2 = theta(b) (Conv2d 64->16, 1x1), DN_Gray/model/dagl.py:208-209.
//
// [G | Theta][p][0..31] = bias + sum_{tap (ky,kx)} sum_{ci} W[.][ci,tap] * bpad[ci][y+ky-1][x+kx-1]
// is an implicit GEMM with M = pixels, N = 32 (16 g outputs + 16 theta outputs; theta only has the centre tap) and
// K = 9 taps x 64 channels.  With b repacked channel-last in groups of 16 channels over the zero-padded image and
// pixels enumerated in padded-flat order (the layout of embed_tc.cu), the A operand of (tap, channel group) for 128
// consecutive pixels is a pixel-shifted view of a 3-row halo held in smem: 36 "virtual taps" of K = 16 each.
//
// G feeds the patch embeddings and through them the neighbour threshold, so it needs fp32 accuracy: b and the
// weights are rescaled by powers of two and split into fp16 hi + lo; acc = bh.Wh + (bh.Wl + bl.Wh), the cross terms
// in their own accumulator (fp32 accumulate in TMEM), added in fp32 in the epilogue.
// The fp32 CUDA-core kernel (prologue.cu) remains for C != 64 and for impl "simt".
#include <cuda_fp16.h>
#include <math.h>
#include "common.cuh"
#include "prologue.cuh"
#include "tc_utils.cuh"

namespace dagl {
using namespace tc;

constexpr int FT_M = 128;                           // pixels per CTA
constexpr int FT_C = 64;                            // input channels handled by this kernel
constexpr int FT_GROUPS = FT_C / 16;
constexpr int FT_N = 32;                            // 16 g + 16 theta outputs
constexpr int FT_SEG_PIX = 144;                     // 128 + 2 (kx) + 7 (alignment) rounded up to 8 ... shares embed_tc's padding (PADK = 3)
constexpr int FT_SEG_BYTES = FT_SEG_PIX * 32;       // 4608
constexpr int FT_VTAPS = 9 * FT_GROUPS;             // 36 virtual taps of K = 16
constexpr int FT_WPART_BYTES = FT_N * 16 * 2;       // 1024: one virtual tap, hi or lo, K-major no-swizzle [2 kc][32 rows][16 B]
constexpr int FT_WTAP_BYTES = 2 * FT_WPART_BYTES;   // hi | lo
constexpr int FT_SM_A = 0;                          // [group][hi|lo][ky] segments
constexpr int FT_SM_W = FT_GROUPS * 2 * 3 * FT_SEG_BYTES;            // 110592 = 108 * 1024
constexpr int FT_SM_BAR = FT_SM_W + FT_VTAPS * FT_WTAP_BYTES;        // + 73728
constexpr int FT_SM_TOTAL = FT_SM_BAR + 128;
constexpr int FT_THREADS = 192;                     // warp 0 loads, warp 1 issues, warps 2-5 epilogue
static_assert(FT_SM_W % 1024 == 0, "weight image alignment");

struct FtGeom { int Wp, NkP, ntile, NPG; };
static FtGeom ft_geom(const Geom& g) {
  FtGeom e;
  e.Wp = (g.W + 2 * PADK + 7) & ~7;                 // the padded-flat pitch of embed_tc.cu / attend_tc.cu: pixel slot p of an
  e.NkP = (g.H - 1) * e.Wp + g.W;                   // item maps to record p + 3 Wp + 3 of their zero-padded fp16 images
  e.ntile = (e.NkP + FT_M - 1) / FT_M;
  const int np = (g.H + 2 * PADK) * e.Wp;
  const int need = FT_M * e.ntile + 2 * PADK * e.Wp + FT_SEG_PIX + 8;
  e.NPG = ((np > need ? np : need) + 7) & ~7;
  return e;
}

__device__ __forceinline__ float pow2_scale_f(unsigned absmax_bits, int target) {
  const float a = __uint_as_float(absmax_bits);
  if (!(a > 0.f) || !isfinite(a)) return 1.f;
  int e;
  frexpf(a, &e);
  return ldexpf(1.f, target - e);
}

// max |b| per image -> bmax[img]
__global__ void absmax_img_kernel(const float* __restrict__ x, size_t n_per_img, unsigned* __restrict__ bmax) {
  pdl_prologue();
  const int img = blockIdx.y;
  const float4* xi = reinterpret_cast<const float4*>(x + (size_t)img * n_per_img);
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_img / 4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(xi + i);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(bmax + img, __float_as_uint(m));
}

// g weight [16][64][3][3], theta weight [16][64] -> per virtual tap (group, tap): [2 k-chunks][64 rows][8 ch] fp16, rows =
// g hi | g lo | theta hi | theta lo (16 each; theta is zero except at the centre tap): the hi and lo parts of the weights are
// stacked along N, so ONE MMA forms both b_hi.W_hi (main accumulator columns) and b_hi.W_lo (cross-term columns).
// wmax[0] = max |w| (float bits).
__global__ void __launch_bounds__(256)
pack_featw_kernel(const float* __restrict__ g_w, const float* __restrict__ g_b, const float* __restrict__ th_w,
                  const float* __restrict__ th_b, uint8_t* __restrict__ out, unsigned* __restrict__ wmax) {
  __shared__ float red[8];
  // meta (floats wmax[1..4]): l1g = max_e sum |g_w[e]|, bg = max |g_b|, l1t, bt  ->  |G| <= max|b| * l1g + bg (a-priori
  // bound: the fp16 scale of the G / theta images written by the feature-map epilogue, no pass over G needed)
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    float l1g = 0.f, l1t = 0.f, bg = 0.f, bt = 0.f;
    for (int e = 0; e < CI; ++e) {
      float a = 0.f, t = 0.f;
      for (int i = lane; i < FT_C * 9; i += 32) a += fabsf(__ldg(g_w + (size_t)e * FT_C * 9 + i));
      for (int i = lane; i < FT_C; i += 32) t += fabsf(__ldg(th_w + (size_t)e * FT_C + i));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); t += __shfl_xor_sync(0xffffffffu, t, o); }
      l1g = fmaxf(l1g, a); l1t = fmaxf(l1t, t);
      bg = fmaxf(bg, fabsf(__ldg(g_b + e))); bt = fmaxf(bt, fabsf(__ldg(th_b + e)));
    }
    if (lane == 0) {
      float* meta = reinterpret_cast<float*>(wmax);
      meta[1] = l1g * 1.0001f; meta[2] = bg; meta[3] = l1t * 1.0001f; meta[4] = bt;
    }
  }
  float m = 0.f;
  for (int i = threadIdx.x; i < CI * FT_C * 9; i += 256) m = fmaxf(m, fabsf(__ldg(g_w + i)));
  for (int i = threadIdx.x; i < CI * FT_C; i += 256) m = fmaxf(m, fabsf(__ldg(th_w + i)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int k = 1; k < 8; ++k) m = fmaxf(m, red[k]);
  if (threadIdx.x == 0) *wmax = __float_as_uint(m);
  const float scale = pow2_scale_f(__float_as_uint(m), 14);
  for (int o = threadIdx.x; o < FT_VTAPS * 2 * FT_N; o += 256) {          // one 16-byte chunk = 8 channels of one output row
    const int v = o / (2 * FT_N), kc = (o / FT_N) & 1, e = o % FT_N;
    const int gq = v / 9, t = v % 9;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int ci = gq * 16 + kc * 8 + 2 * j + u;
        float w = 0.f;
        if (e < CI) w = __ldg(g_w + ((size_t)e * FT_C + ci) * 9 + t);
        else if (t == 4) w = __ldg(th_w + (size_t)(e - CI) * FT_C + ci);
        x[u] = w * scale;
      }
      const __half h0 = __float2half_rn(x[0]), h1 = __float2half_rn(x[1]);
      const __half l0 = __float2half_rn(x[0] - __half2float(h0)), l1 = __float2half_rn(x[1] - __half2float(h1));
      hi[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      lo[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    // rows of the virtual tap's B operand: [g hi 0..15 | g lo 16..31 | theta hi 32..47 | theta lo 48..63], K-major no swizzle
    const int row_hi = e < CI ? e : 2 * CI + (e - CI), row_lo = row_hi + CI;
    uint8_t* base = out + (size_t)v * FT_WTAP_BYTES + (size_t)kc * (2 * FT_N) * 16;
    *reinterpret_cast<uint4*>(base + (size_t)row_hi * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + (size_t)row_lo * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// b [64][H][W] fp32 -> per 16-channel group: zero-padded flat [NPG pixels][16 ch] fp16 hi and lo (SWIZZLE_32B pre-applied)
// layout: [img][group][hi|lo][NPG][32 B]
// The same launch also computes gamma / beta (dagl.py:213-215) in its first `n_gb` CTAs: the two parts only share the
// input, the gamma/beta part is latency-bound (19 us on its own) and hides behind the bandwidth-bound repack.
__global__ void __launch_bounds__(GB_THREADS)
pack_b_gamma_beta_kernel(Geom g, FtGeom eg, const float* __restrict__ b, const unsigned* __restrict__ bmax,
                         uint8_t* __restrict__ bimg, int n_gb, HeadPtrs thr_w, HeadPtrs thr_b, HeadPtrs bias_w, HeadPtrs bias_b,
                         float* __restrict__ gamma, float* __restrict__ beta) {
  pdl_prologue();
  extern __shared__ float gb_smem[];
  if ((int)blockIdx.x < n_gb) {                           // block-uniform branch: (query block, virtual image)
    const int qblocks = (g.Nq + 31) / 32;
    const int v = blockIdx.x / qblocks, h = g.head(v);
    gamma_beta_body(g, b, static_cast<const float*>(thr_w.p[h]), static_cast<const float*>(thr_b.p[h]),
                    static_cast<const float*>(bias_w.p[h]), static_cast<const float*>(bias_b.p[h]), gamma, beta, gb_smem,
                    blockIdx.x % qblocks, v);
    return;
  }
  const int pblocks = (eg.NPG + GB_THREADS - 1) / GB_THREADS;
  const int pb = blockIdx.x - n_gb;
  const int img = pb / (pblocks * FT_GROUPS), gq = (pb / pblocks) % FT_GROUPS;
  const int pix = (pb % pblocks) * GB_THREADS + threadIdx.x;
  if (pix >= eg.NPG) return;
  const float scale = pow2_scale_f(bmax[img], 14);
  const int r = pix / eg.Wp, cc = pix % eg.Wp;
  const int y = r - PADK, x = cc - PADK;
  const bool inb = (y >= 0 && y < g.H && x >= 0 && x < g.W);
  const float* src = b + (((size_t)img * g.C + gq * 16) * g.H + (inb ? y : 0)) * g.W + (inb ? x : 0);
  uint32_t h[8], l[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float a0 = 0.f, a1 = 0.f;
    if (inb) {
      a0 = __ldg(src + (size_t)(2 * j) * g.Nk) * scale;
      a1 = __ldg(src + (size_t)(2 * j + 1) * g.Nk) * scale;
    }
    const __half h0 = __float2half_rn(a0), h1 = __float2half_rn(a1);
    const __half l0 = __float2half_rn(a0 - __half2float(h0)), l1 = __float2half_rn(a1 - __half2float(h1));
    h[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
    l[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
  }
  const int sw = (pix >> 2) & 1;
  uint8_t* base = bimg + ((size_t)(img * FT_GROUPS + gq) * 2) * (size_t)eg.NPG * 32;
  uint4* dh = reinterpret_cast<uint4*>(base + (size_t)pix * 32);
  uint4* dl = reinterpret_cast<uint4*>(base + (size_t)eg.NPG * 32 + (size_t)pix * 32);
  dh[sw] = make_uint4(h[0], h[1], h[2], h[3]);
  dh[sw ^ 1] = make_uint4(h[4], h[5], h[6], h[7]);
  dl[sw] = make_uint4(l[0], l[1], l[2], l[3]);
  dl[sw ^ 1] = make_uint4(l[4], l[5], l[6], l[7]);
}

// gamma / beta of all NH heads of a stage in one pass over the shared input (stage calls; launched next to the repack)
template <int NH>
__global__ void __launch_bounds__(GB_THREADS)
gamma_beta_heads_kernel(Geom g, const float* __restrict__ b, HeadPtrs thr_w, HeadPtrs thr_b, HeadPtrs bias_w, HeadPtrs bias_b,
                        float* __restrict__ gamma, float* __restrict__ beta) {
  pdl_prologue();
  extern __shared__ float gbh_smem[];
  gamma_beta_heads_body<NH>(g, b, thr_w, thr_b, bias_w, bias_b, gamma, beta, gbh_smem, blockIdx.x, blockIdx.y);
}

// Persistent: one CTA per SM walks the (image, 128-pixel tile) items; the packed weights (72 KB) are loaded ONCE per CTA
// (they were 40 % of the smem fill of a one-tile CTA, and the kernel is bound by that L2 -> SM traffic), the halo is
// single-buffered and the two 64-column accumulator sets alternate so that the epilogue of a tile overlaps the halo load
// and the MMAs of the next.
__global__ void __launch_bounds__(FT_THREADS, 1)
featmap_tc_kernel(Geom g, FtGeom eg, const uint8_t* __restrict__ bimg, HeadPtrs wpack, HeadPtrs g_b, HeadPtrs th_b,
                  const unsigned* __restrict__ bmax, float* __restrict__ G /*nullable: fp32 copies for the debug entry*/,
                  float* __restrict__ Th, unsigned* __restrict__ absmax /*[B][AMAX_STRIDE]*/, HeadPtrs fcmeta,
                  uint8_t* __restrict__ ghi, uint8_t* __restrict__ glo, int npg /*records per image in ghi / glo*/,
                  uint8_t* __restrict__ thp, int np_t /*records per image in thp*/) {
  // Epilogue outputs (what the embedding and graph kernels consume; nothing is re-packed by a later pass):
  //   ghi / glo  zero-padded flat fp16 hi / lo images of G   [v][npg][16 ch], SWIZZLE_32B pre-applied   (embed_tc.cu)
  //   thp        zero-padded flat fp16 image of theta          [v][np_t][16 ch], likewise                 (attend_tc.cu)
  //   absmax[v]  the fp16 scales of everything downstream as a-priori bounds: |G|, |theta| from max|b| and the L1 norms of
  //              the filters, |Q|, |K| from the bound on |G| and the L1 norms of fc1 / fc2 (fcmeta).  A bound that is loose
  //              by a factor 2^k costs nothing for hi+lo split operands as long as k < ~17 (fp16 keeps 11 bits per part at
  //              every exponent above 2^-14: the absolute error floor stays below 2^-22 of the largest value).
  // Work item w (head-major: a persistent CTA re-loads the 72 KB of packed weights at most NH times):
  //   head = w / (nreal * ntile), real image = (w / ntile) % nreal, tile = w % ntile; virtual image = real * NH + head.
  // wpack.p[h] = packed g/theta weights of head h, followed (256-aligned) by wmax (float bits of max |w|).
  pdl_prologue();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FT_SM_BAR);
  uint64_t* w_full = bars + 0;      // weights resident
  uint64_t* a_full = bars + 1;      // [2] halo of channel groups {0,1} / {2,3} of the current item: the two halves are a 2-stage
  uint64_t* a_empty = bars + 3;     // [2] ring, so the load of one half overlaps the MMAs of the other (its MMAs have completed)
  uint64_t* d_full = bars + 5;      // [2]
  uint64_t* d_empty = bars + 7;     // [2] 4 arrivals (one per epilogue warp)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = warp_id_uniform();
  const int tid = threadIdx.x;
  const int nwork = g.B * eg.ntile;
  const int per_head = (g.B / g.NH) * eg.ntile;
  constexpr size_t WMAX_OFF = ((size_t)FT_VTAPS * FT_WTAP_BYTES + 255) & ~(size_t)255;

  if (tid == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(a_full + i, 1); mbar_init(a_empty + i, 1); mbar_init(d_full + i, 1); mbar_init(d_empty + i, 4); }
    mbar_init_fence();
  }
  if (warp == 1) tmem_alloc<128>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;

  if (warp == 0) {
    // ===================== producer: one bulk copy per lane (24 halo segments; lane 24: the weights, once) ==========
    const int lane = tid & 31;
    int it = 0, cur_head = -1;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
      const int head = w / per_head, img = (w / eg.ntile) % (g.B / g.NH), p0 = (w % eg.ntile) * FT_M;   // img: REAL image
      if (head != cur_head) {
        // new weights: every MMA of the previous item (both halves) must have completed
        if (lane == 0) {
          mbar_wait(a_empty + 1, ((uint32_t)it & 1u) ^ 1u);
          mbar_arrive_expect_tx(w_full, FT_VTAPS * FT_WTAP_BYTES);
        }
        __syncwarp();
        if (lane == FT_GROUPS * 2 * 3)
          bulk_g2s(smem + FT_SM_W, static_cast<const uint8_t*>(wpack.p[head]), FT_VTAPS * FT_WTAP_BYTES, w_full);
        cur_head = head;
      }
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        if (lane == 0) {
          mbar_wait(a_empty + half, ((uint32_t)it & 1u) ^ 1u);          // the MMAs that read this half of the previous item are complete
          mbar_arrive_expect_tx(a_full + half, (FT_GROUPS / 2) * 2 * 3 * FT_SEG_BYTES);
        }
        __syncwarp();
        if (lane / 12 == half && lane < FT_GROUPS * 2 * 3) {
          const int gq = lane / 6, part = (lane / 3) & 1, ky = lane % 3;
          const uint8_t* src = bimg + ((size_t)(img * FT_GROUPS + gq) * 2 + part) * (size_t)eg.NPG * 32;
          const int first = (p0 + (ky + 2) * eg.Wp) & ~7;                 // 3x3 / pad 1 inside the pad-3 frame: rows y+ky+2
          bulk_g2s(smem + FT_SM_A + ((gq * 2 + part) * 3 + ky) * FT_SEG_BYTES, src + (size_t)first * 32, FT_SEG_BYTES, a_full + half);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t abase = smem_u32(smem + FT_SM_A), wbase = smem_u32(smem + FT_SM_W);
      constexpr uint32_t id64 = instr_desc(FT_M, 64, FMT_F16, FMT_F16, 0, 0);
      constexpr uint32_t id32 = instr_desc(FT_M, 32, FMT_F16, FMT_F16, 0, 0);
      constexpr uint32_t id16 = instr_desc(FT_M, 16, FMT_F16, FMT_F16, 0, 0);
      int it = 0, cur_head = -1, nloads = 0;
      for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
        const int p0 = (w % eg.ntile) * FT_M, ab = it & 1, head = w / per_head;
        const uint32_t d_main = tbase + ab * 64;
        if (head != cur_head) { mbar_wait(w_full, (uint32_t)nloads & 1u); ++nloads; cur_head = head; }
        mbar_wait(d_empty + ab, ((uint32_t)(it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        // Accumulator columns: [g main | g cross | theta main | theta cross] (16 each).  Per virtual tap TWO MMAs instead of
        // three: b_hi x [W_hi | W_lo] (N = 32; centre tap N = 64 incl. theta) and b_lo x W_hi (N = 16) into the cross
        // columns (these small-N MMAs cost ~36-48 cycles each whatever N is, so the count is what matters).  The centre tap
        // goes first in every group, so that the very first MMA initialises all 64 columns.
        const int order[9] = {4, 0, 1, 2, 3, 5, 6, 7, 8};
        bool first = true;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {               // channel groups {0,1} / {2,3}: the two halves of the halo ring
          mbar_wait(a_full + half, (uint32_t)it & 1u);
          tc_fence_after();
#pragma unroll 1
          for (int gq = 2 * half; gq < 2 * half + 2; ++gq) {
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            const int t = order[i], ky = t / 3, kx = t % 3;
            const int off = ((p0 + (ky + 2) * eg.Wp) & 7) + kx + 2;
            const uint32_t a_hi = abase + ((gq * 2 + 0) * 3 + ky) * FT_SEG_BYTES + off * 32;
            const uint32_t a_lo = abase + ((gq * 2 + 1) * 3 + ky) * FT_SEG_BYTES + off * 32;
            // A: K-major SWIZZLE_32B, rows (pixels) 32 B apart, 8-row groups 256 B apart
            const uint64_t da_hi = (uint64_t)((a_hi >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(256 >> 4) << 32) |
                                   ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
            const uint64_t da_lo = (uint64_t)((a_lo >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(256 >> 4) << 32) |
                                   ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
            const uint32_t wv = wbase + (gq * 9 + t) * FT_WTAP_BYTES;
            const uint64_t db_all = smem_desc(wv, 2 * FT_N * 16, 128);                 // rows 0..: g hi, g lo, theta hi, theta lo
            const uint64_t db_thi = smem_desc(wv + 2 * CI * 16, 2 * FT_N * 16, 128);   // rows 32..47: theta hi
            const uint32_t acc = first ? 0u : 1u;
            if (t == 4) {
              mma_f16_ss(d_main, da_hi, db_all, id64, acc);                     // b_hi . [Wg_hi | Wg_lo | Wt_hi | Wt_lo]
              mma_f16_ss(d_main + CI, da_lo, db_all, id16, 1);                  // b_lo . Wg_hi -> g cross
              mma_f16_ss(d_main + 3 * CI, da_lo, db_thi, id16, 1);              // b_lo . Wt_hi -> theta cross
            } else {
              mma_f16_ss(d_main, da_hi, db_all, id32, acc);                     // b_hi . [Wg_hi | Wg_lo]
              mma_f16_ss(d_main + CI, da_lo, db_all, id16, 1);                  // b_lo . Wg_hi -> g cross
            }
            first = false;
          }
          }
          // unconditional on purpose: a uniform-predicated tcgen05.commit whose (unused) address operand is misaligned
          // still faults ("misaligned address")
          mma_commit(a_empty + half);                        // this half of the halo may be refilled (next item)
        }
        mma_commit(d_full + ab);
      }
    }
  } else {
    // ===================== epilogue: thread = pixel row =====================
    const int quad = warp & 3, lane = tid & 31;
    const int r = quad * 32 + lane;
    int it = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
      const int head = w / per_head, real = (w / eg.ntile) % (g.B / g.NH), ab = it & 1;
      const int img = real * g.NH + head;                                   // virtual image: outputs
      const int p = (w % eg.ntile) * FT_M + r;
      const int y = p / eg.Wp, x = p % eg.Wp;
      const bool valid = (p < eg.NkP) && (x < g.W);
      const float* meta = reinterpret_cast<const float*>(static_cast<const uint8_t*>(wpack.p[head]) + WMAX_OFF);
      const unsigned wmax = __float_as_uint(meta[0]);
      const float bmaxf = __uint_as_float(bmax[real]);
      const float bound_g = bmaxf * meta[1] + meta[2], bound_t = bmaxf * meta[3] + meta[4];
      const float sg = pow2_scale_f(__float_as_uint(bound_g), 14), st = pow2_scale_f(__float_as_uint(bound_t), 12);
      const float inv = 1.f / (pow2_scale_f(wmax, 14) * pow2_scale_f(bmax[real], 14));
      const float* gbias = static_cast<const float*>(g_b.p[head]);
      const float* tbias = static_cast<const float*>(th_b.p[head]);
      const int tile = w % eg.ntile;
      if (tile == 0 && r == 0) {                            // one thread per virtual image: the scales of everything downstream
        const float* fm = static_cast<const float*>(fcmeta.p[head]);       // [l1 fc1, max|b1|, l1 fc2, max|b2|]
        absmax[img * AMAX_STRIDE + AMAX_G] = __float_as_uint(bound_g);
        absmax[img * AMAX_STRIDE + AMAX_THETA] = __float_as_uint(bound_t);
        absmax[img * AMAX_STRIDE + AMAX_Q] = __float_as_uint(bound_g * fm[0] + fm[1]);
        absmax[img * AMAX_STRIDE + AMAX_K] = __float_as_uint(bound_g * fm[2] + fm[3]);
      }
      const uint32_t trow = tbase + ((uint32_t)(quad * 32) << 16) + ab * 64;
      mbar_wait(d_full + ab, (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      uint32_t v[2][16], vc[2][16];
      tmem_ld16(trow, v[0]);                             // g main | g cross | theta main | theta cross
      tmem_ld16(trow + 16, vc[0]);
      tmem_ld16(trow + 32, v[1]);
      tmem_ld16(trow + 48, vc[1]);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d_empty + ab);          // the accumulator set is in registers now
      // slot p of the padded-flat enumeration is record rec of the zero-padded images; dummy slots (x >= W, p >= NkP) are the
      // right / left zero borders of the rows, so every record between the head and the tail region is written here
      const size_t rec = (size_t)p + 3 * eg.Wp + 3;
      uint32_t gh[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u}, gl[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u}, th[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      if (valid) {
        float og[CI], ot[CI];
#pragma unroll
        for (int c = 0; c < CI; ++c) {
          og[c] = (__uint_as_float(v[0][c]) + __uint_as_float(vc[0][c])) * inv + __ldg(gbias + c);
          ot[c] = (__uint_as_float(v[1][c]) + __uint_as_float(vc[1][c])) * inv + __ldg(tbias + c);
        }
        if (G != nullptr) {                                  // debug entry: fp32 copies (dagl_ce_workspace_view)
          float* dg = G + (size_t)img * CI * g.Nk + (size_t)y * g.W + x;
          float* dt = Th + (size_t)img * CI * g.Nk + (size_t)y * g.W + x;
#pragma unroll
          for (int c = 0; c < CI; ++c) { dg[(size_t)c * g.Nk] = og[c]; dt[(size_t)c * g.Nk] = ot[c]; }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a0 = og[2 * j] * sg, a1 = og[2 * j + 1] * sg;
          const __half h0 = __float2half_rn(a0), h1 = __float2half_rn(a1);
          const __half l0 = __float2half_rn(a0 - __half2float(h0)), l1 = __float2half_rn(a1 - __half2float(h1));
          gh[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
          gl[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
          const __half t0 = __float2half_rn(ot[2 * j] * st), t1 = __float2half_rn(ot[2 * j + 1] * st);
          th[j] = (uint32_t)__half_as_ushort(t0) | ((uint32_t)__half_as_ushort(t1) << 16);
        }
      }
      {
        const int sw = (int)(rec >> 2) & 1;                  // SWIZZLE_32B pre-applied: the 16-byte halves swap when (record & 4)
        uint4* dh = reinterpret_cast<uint4*>(ghi + ((size_t)img * npg + rec) * 32);
        uint4* dl = reinterpret_cast<uint4*>(glo + ((size_t)img * npg + rec) * 32);
        uint4* dt = reinterpret_cast<uint4*>(thp + ((size_t)img * np_t + rec) * 32);
        dh[sw] = make_uint4(gh[0], gh[1], gh[2], gh[3]); dh[sw ^ 1] = make_uint4(gh[4], gh[5], gh[6], gh[7]);
        dl[sw] = make_uint4(gl[0], gl[1], gl[2], gl[3]); dl[sw ^ 1] = make_uint4(gl[4], gl[5], gl[6], gl[7]);
        dt[sw] = make_uint4(th[0], th[1], th[2], th[3]); dt[sw ^ 1] = make_uint4(th[4], th[5], th[6], th[7]);
      }
      // head (records before the first slot) and tail (after the last slot) of the padded images are zero borders too:
      // written by the first / last item of the virtual image
      auto zero_range = [&](size_t a0, size_t end_g, size_t end_t) {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (size_t i = a0 + r; i < (end_g > end_t ? end_g : end_t); i += FT_M) {
          if (i < end_g) {
            uint4* dh = reinterpret_cast<uint4*>(ghi + ((size_t)img * npg + i) * 32);
            uint4* dl = reinterpret_cast<uint4*>(glo + ((size_t)img * npg + i) * 32);
            dh[0] = z; dh[1] = z; dl[0] = z; dl[1] = z;
          }
          if (i < end_t) {
            uint4* dt = reinterpret_cast<uint4*>(thp + ((size_t)img * np_t + i) * 32);
            dt[0] = z; dt[1] = z;
          }
        }
      };
      if (tile == 0) zero_range(0, (size_t)3 * eg.Wp + 3, (size_t)3 * eg.Wp + 3);
      if (tile == eg.ntile - 1) zero_range((size_t)eg.ntile * FT_M + 3 * eg.Wp + 3, (size_t)npg, (size_t)np_t);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tbase);
}

// ---- host side -----------------------------------------------------------------------------
static inline size_t align_up_f(size_t x) { return (x + 255) & ~(size_t)255; }

bool feature_maps_tc_supported(const Geom& g) { return g.C == FT_C; }

size_t feature_maps_tc_workspace_bytes(const Geom& g) {
  if (!feature_maps_tc_supported(g)) return 0;
  const FtGeom eg = ft_geom(g);
  const size_t nreal = (size_t)(g.B / g.NH);
  return align_up_f(nreal * FT_GROUPS * 2 * eg.NPG * 32) + align_up_f(nreal * sizeof(unsigned));
}

// packed g/theta weights | wmax (float bits), then the bound meta [l1g, bg, l1t, bt] (pack_featw_kernel)
size_t feature_maps_tc_packed_weights_bytes() { return align_up_f((size_t)FT_VTAPS * FT_WTAP_BYTES) + align_up_f(64); }

int launch_pack_feat_weights(int C, const float* g_w, const float* g_b, const float* th_w, const float* th_b, void* packed,
                             size_t packed_bytes, cudaStream_t st) {
  if (C != FT_C) return 0;                                   // fp32 CUDA-core feature kernel: nothing to pack
  if (packed_bytes < feature_maps_tc_packed_weights_bytes()) {
    call_state().err = "packed-weights buffer too small";
    return -3;
  }
  uint8_t* wpack = static_cast<uint8_t*>(packed);
  unsigned* wmax = reinterpret_cast<unsigned*>(wpack + align_up_f((size_t)FT_VTAPS * FT_WTAP_BYTES));
  pack_featw_kernel<<<1, 256, 0, st>>>(g_w, g_b, th_w, th_b, wpack, wmax);
  DAGL_LAUNCH_CHECK();
  return 0;
}

// hw.packed[h]: weights of head h packed by dagl_ce_pack_weights_f32 (the g/theta image follows the fc image; never null
// here: forward_impl packs into the workspace when the caller did not).  gamma / beta (dagl.py:213-215) are computed inside
// the launch that repacks b.  `reuse_b`: the repacked input (and its maximum) left in `ws` by the previous call is still
// valid (same b), so only the gamma / beta part of that launch runs.  `out`: where the epilogue writes the fp16 images of G
// and theta that the embedding and graph kernels consume.  G / Th (nullable): fp32 copies for the debug entry.
int launch_feature_maps_tc(const Geom& g, const float* b, const HeadWeights& hw, float* G, float* Th, float* gamma,
                           float* beta, unsigned* absmax, void* ws, size_t ws_bytes, bool reuse_b, const FeatTargets& out,
                           cudaStream_t st) {
  const FtGeom eg = ft_geom(g);
  if (!feature_maps_tc_supported(g) || ws_bytes < feature_maps_tc_workspace_bytes(g)) {
    call_state().err = "feature maps (tc): unsupported channel count or workspace too small";
    return -3;
  }
  const int nreal = g.B / g.NH;
  char* p = static_cast<char*>(ws);
  uint8_t* bimg = reinterpret_cast<uint8_t*>(p); p += align_up_f((size_t)nreal * FT_GROUPS * 2 * eg.NPG * 32);
  unsigned* bmax = reinterpret_cast<unsigned*>(p); p += align_up_f((size_t)nreal * sizeof(unsigned));
  HeadPtrs wpack{}, gb{}, tb{}, thr_w{}, thr_b{}, bias_w{}, bias_b{}, fcmeta{};
  for (int h = 0; h < g.NH; ++h) {
    wpack.p[h] = static_cast<const char*>(hw.packed[h]) + embed_tc_packed_weights_bytes();
    fcmeta.p[h] = embed_tc_fc_meta(hw.packed[h]);
    gb.p[h] = hw.g_b[h]; tb.p[h] = hw.th_b[h];
    thr_w.p[h] = hw.thr_w[h]; thr_b.p[h] = hw.thr_b[h]; bias_w.p[h] = hw.bias_w[h]; bias_b.p[h] = hw.bias_b[h];
  }

  if (!reuse_b) {
    DAGL_CUDA_OK(cudaMemsetAsync(bmax, 0, (size_t)nreal * sizeof(unsigned), st));
    const size_t n_img = (size_t)g.C * g.Nk;
    DAGL_CUDA_OK(launch_pdl(absmax_img_kernel, dim3(128, nreal), 256, 0, st, b, n_img, bmax));
    DAGL_LAUNCH_CHECK();
  }
  if (g.NH > 1) {
    // stage call: gamma / beta of all heads in one pass over b (own launch: it needs NH x the filter smem, which would
    // throttle the bandwidth-bound repack CTAs if the two shared a launch as they do for a single head)
    const size_t smem_h = gamma_beta_heads_smem_bytes(g.C, g.NH);
    auto kern = g.NH == 2 ? gamma_beta_heads_kernel<2> : g.NH == 3 ? gamma_beta_heads_kernel<3> : gamma_beta_heads_kernel<4>;
    DAGL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_h));
    DAGL_CUDA_OK(launch_pdl(kern, dim3((g.Nq + 31) / 32, nreal), GB_THREADS, smem_h, st, g, b, thr_w, thr_b, bias_w, bias_b, gamma, beta));
    DAGL_LAUNCH_CHECK();
  }
  {
    const int n_gb = g.NH > 1 ? 0 : ((g.Nq + 31) / 32) * g.B;
    const int n_pack = reuse_b ? 0 : ((eg.NPG + GB_THREADS - 1) / GB_THREADS) * FT_GROUPS * nreal;
    const size_t smem = gamma_beta_smem_bytes(g.C);
    if (smem > 48 * 1024)
      DAGL_CUDA_OK(cudaFuncSetAttribute(pack_b_gamma_beta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (n_gb + n_pack > 0) {
      DAGL_CUDA_OK(launch_pdl(pack_b_gamma_beta_kernel, n_gb + n_pack, GB_THREADS, n_gb ? smem : 0, st, g, eg, b, bmax, bimg, n_gb, thr_w,
                              thr_b, bias_w, bias_b, gamma, beta));
      DAGL_LAUNCH_CHECK();
    }
  }
  DAGL_CUDA_OK(cudaFuncSetAttribute(featmap_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SM_TOTAL));
  int dev = 0, sms = 148;
  DAGL_CUDA_OK(cudaGetDevice(&dev));
  DAGL_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int nwork = g.B * eg.ntile;
  DAGL_CUDA_OK(launch_pdl(featmap_tc_kernel, dim3(nwork < sms ? nwork : sms), FT_THREADS, FT_SM_TOTAL, st, g, eg, bimg, wpack, gb, tb, bmax, G, Th,
                          absmax, fcmeta, out.ghi, out.glo, out.npg, out.thp, out.np_t));
  DAGL_LAUNCH_CHECK();
  return 0;
}

}  // namespace dagl
