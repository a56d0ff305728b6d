// gamma = thr_conv(pad4(b)), beta = bias_conv(pad4(b))  (dagl.py:213-215) as a device function, so that the tensor-core
// path can run it inside the launch that repacks b (featmap_tc.cu: the two are independent and the gamma/beta part is
// latency-bound) while the fp32 path keeps it as a kernel of its own (prologue.cu).
#pragma once
#include "common.cuh"

namespace dagl {

constexpr int GB_GROUPS = 16;
constexpr int GB_THREADS = 32 * GB_GROUPS;
inline size_t gamma_beta_smem_bytes(int C) { return (size_t)(C * KK + GB_GROUPS * 32) * sizeof(float2); }

// CTA = 32 consecutive queries (lanes) x 16 channel groups (warps) of image `img`, query block `qblock`; the SAME padding
// is a predicate, not a copy; both 7x7 filters are staged in smem and read as warp-uniform broadcasts.  The
// channel-group partials are summed in a fixed order (deterministic).  Needs GB_THREADS threads and gamma_beta_smem_bytes.
__device__ __forceinline__ void gamma_beta_body(const Geom& g, const float* __restrict__ b, const float* __restrict__ thr_w,
                                                const float* __restrict__ thr_b, const float* __restrict__ bias_w,
                                                const float* __restrict__ bias_b, float* __restrict__ gamma,
                                                float* __restrict__ beta, float* smem, int qblock, int img /*virtual image: outputs*/) {
  float2* w_s = reinterpret_cast<float2*>(smem);                      // [C*49] (thr, bias)
  float2* red = reinterpret_cast<float2*>(smem) + g.C * KK;           // [GB_GROUPS][32]
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < g.C * KK; i += GB_THREADS) w_s[i] = make_float2(__ldg(thr_w + i), __ldg(bias_w + i));
  __syncthreads();
  const int q = qblock * 32 + lane;
  const bool live = q < g.Nq;
  const int qy = live ? q / g.nqx : 0, qx = live ? q % g.nqx : 0;
  const int y0 = qy * SQ - g.qpad_top, x0 = qx * SQ - g.qpad_left;
  const float* bi = b + (size_t)g.real_img(img) * g.C * g.Nk;
  bool rok[KS], cok[KS];
#pragma unroll
  for (int k = 0; k < KS; ++k) { rok[k] = live && (y0 + k >= 0) && (y0 + k < g.H); cok[k] = (x0 + k >= 0) && (x0 + k < g.W); }
  float a0[2] = {0.f, 0.f}, a1[2] = {0.f, 0.f};              // two chains per output: the FMA latency is the bound
  // All 49 loads of a channel are issued before the first FMA (written as two loops on purpose: with the load next to its FMA
  // the compiler keeps only a handful of loads in flight and the kernel becomes a chain of L2 round trips).
  for (int ci = grp; ci < g.C; ci += GB_GROUPS) {
    const float* bc = bi + (size_t)ci * g.Nk + y0 * g.W + x0;
    const float2* wc = w_s + ci * KK;
    float v[KK];
#pragma unroll
    for (int ky = 0; ky < KS; ++ky)
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) v[ky * KS + kx] = (rok[ky] && cok[kx]) ? __ldg(bc + ky * g.W + kx) : 0.f;
#pragma unroll
    for (int t = 0; t < KK; ++t) {
      const float2 w = wc[t];
      a0[t & 1] = fmaf(v[t], w.x, a0[t & 1]);
      a1[t & 1] = fmaf(v[t], w.y, a1[t & 1]);
    }
  }
  red[grp * 32 + lane] = make_float2(a0[0] + a0[1], a1[0] + a1[1]);
  __syncthreads();
  if (grp == 0 && live) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int k = 0; k < GB_GROUPS; ++k) { const float2 r = red[k * 32 + lane]; s0 += r.x; s1 += r.y; }
    gamma[(size_t)img * g.Nq + q] = s0 + thr_b[0];
    beta[(size_t)img * g.Nq + q] = s1 + bias_b[0];
  }
}

// The same for NH heads that share the input (one CES stage): CTA = 32 queries x 16 channel groups of REAL image `real`,
// every b value is loaded once and used for the 2 x NH filters (the per-head version is bound by the load pipe: two
// load-pipe operations per two FMAs; here five per 2 NH).  Filters in smem as [ci*49 + tap][head] (thr, bias) pairs.
// Summation order per head is the same as in gamma_beta_body, so the two give bit-identical results.
inline size_t gamma_beta_heads_smem_bytes(int C, int NH) { return (size_t)(C * KK + GB_GROUPS * 32) * NH * sizeof(float2); }

template <int NH>
__device__ __forceinline__ void gamma_beta_heads_body(const Geom& g, const float* __restrict__ b, const HeadPtrs& thr_w,
                                                      const HeadPtrs& thr_b, const HeadPtrs& bias_w, const HeadPtrs& bias_b,
                                                      float* __restrict__ gamma, float* __restrict__ beta, float* smem,
                                                      int qblock, int real) {
  float2* w_s = reinterpret_cast<float2*>(smem);                      // [C*49][NH]
  float2* red = w_s + (size_t)g.C * KK * NH;                          // [GB_GROUPS][32][NH]
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < g.C * KK * NH; i += GB_THREADS) {
    const int h = i % NH, j = i / NH;
    w_s[i] = make_float2(__ldg(static_cast<const float*>(thr_w.p[h]) + j), __ldg(static_cast<const float*>(bias_w.p[h]) + j));
  }
  __syncthreads();
  const int q = qblock * 32 + lane;
  const bool live = q < g.Nq;
  const int qy = live ? q / g.nqx : 0, qx = live ? q % g.nqx : 0;
  const int y0 = qy * SQ - g.qpad_top, x0 = qx * SQ - g.qpad_left;
  const float* bi = b + (size_t)real * g.C * g.Nk;
  bool rok[KS], cok[KS];
#pragma unroll
  for (int k = 0; k < KS; ++k) { rok[k] = live && (y0 + k >= 0) && (y0 + k < g.H); cok[k] = (x0 + k >= 0) && (x0 + k < g.W); }
  float a0[NH][2], a1[NH][2];
#pragma unroll
  for (int h = 0; h < NH; ++h) { a0[h][0] = a0[h][1] = a1[h][0] = a1[h][1] = 0.f; }
  for (int ci = grp; ci < g.C; ci += GB_GROUPS) {
    const float* bc = bi + (size_t)ci * g.Nk + y0 * g.W + x0;
    const float2* wc = w_s + (size_t)ci * KK * NH;
    float v[KK];                                                      // all loads of the channel first (see gamma_beta_body)
#pragma unroll
    for (int ky = 0; ky < KS; ++ky)
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) v[ky * KS + kx] = (rok[ky] && cok[kx]) ? __ldg(bc + ky * g.W + kx) : 0.f;
#pragma unroll
    for (int t = 0; t < KK; ++t) {
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const float2 w = wc[t * NH + h];
        a0[h][t & 1] = fmaf(v[t], w.x, a0[h][t & 1]);
        a1[h][t & 1] = fmaf(v[t], w.y, a1[h][t & 1]);
      }
    }
  }
#pragma unroll
  for (int h = 0; h < NH; ++h) red[(grp * 32 + lane) * NH + h] = make_float2(a0[h][0] + a0[h][1], a1[h][0] + a1[h][1]);
  __syncthreads();
  if (grp < NH && live) {                                             // warp h finishes head h
    const int h = grp;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int k = 0; k < GB_GROUPS; ++k) { const float2 r = red[(k * 32 + lane) * NH + h]; s0 += r.x; s1 += r.y; }
    const size_t o = (size_t)(real * NH + h) * g.Nq + q;              // virtual image real * NH + h
    gamma[o] = s0 + static_cast<const float*>(thr_b.p[h])[0];
    beta[o] = s1 + static_cast<const float*>(bias_b.p[h])[0];
  }
}

}  // namespace dagl
