// Fused graph stage, fp32 CUDA-core version (reference: CE.forward, DN_Gray/model/dagl.py:250-272).
//
//   S = Q K^T -> adaptive threshold mask -> softmax(10 S mask) * mask_b -> P V -> fold / count
//
// Flash-style: the N_q x N_k score matrix (dagl.py:250) and the unfolded value
// patches (dagl.py:224-230) are never materialised.  A CTA owns 32 queries and
// streams key tiles (<= 64 consecutive keys of one image row); the value operand
// of a key tile is a 7-row halo of the 16-channel theta map held in smem, from
// which V[k][(c,dy,dx)] = theta_pad[c][ky+dy][kx+dx] is read as a sliding window.
// Keys are split across blockIdx.y; partial (max, sum, acc) triples are merged
// and folded by two small follow-up kernels (deterministic: no float atomics).
//
// This kernel keeps S in fp32 FMA arithmetic, so its neighbour mask is
// bit-faithful to the reference up to fp32 summation order.  It is the
// correctness anchor for the tcgen05 kernel and the path for shapes that one
// does not cover.
#include <math.h>
#include "common.cuh"

namespace dagl {

constexpr int AT_BM = 32;
constexpr int AT_BN = 64;
constexpr int AT_THREADS = 256;
constexpr int AT_TROW = 71;          // odd halo-row stride (70 columns used)
constexpr int AT_TROWS = CI * KS;    // 112 (c,dy) rows
constexpr int AT_PV_THREADS = 2 * AT_TROWS;

constexpr int AT_SMEM_FLOATS = AT_BM * ED + AT_BN * ED + AT_TROWS * AT_TROW + AT_BM * AT_BN + 3 * AT_BM;

__device__ __forceinline__ float half_warp_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(AT_THREADS, 1)
attend_simt_kernel(Geom g, const float* __restrict__ Q, const float* __restrict__ K,
                   const float* __restrict__ Kbar, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ theta, float scale,
                   int nsplit, int ntx, int tw, float* __restrict__ Opart, float* __restrict__ mpart,
                   float* __restrict__ lpart, uint32_t* __restrict__ mask_bits, int32_t* __restrict__ nnz) {
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;                        // [32][196]
  float* Ks = Qs + AT_BM * ED;             // [64][196]
  float* Ts = Ks + AT_BN * ED;             // [112][71]
  float* Ps = Ts + AT_TROWS * AT_TROW;     // [32][64]
  float* scs = Ps + AT_BM * AT_BN;         // [32] per-row rescale of this tile
  float* tA = scs + AT_BM;                 // [32] mu*gamma
  float* tB = tA + AT_BM;                  // [32] beta

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int img = blockIdx.z, split = blockIdx.y, q0 = blockIdx.x * AT_BM;

  // ---- query tile ---------------------------------------------------------
  {
    float4* Qs4 = reinterpret_cast<float4*>(Qs);
    for (int i = tid; i < AT_BM * (ED / 4); i += AT_THREADS) {
      int q = i / (ED / 4), e4 = i % (ED / 4);
      int qq = q0 + q;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (qq < g.Nq) v = __ldg(reinterpret_cast<const float4*>(Q + ((size_t)img * g.Nq + qq) * ED) + e4);
      Qs4[i] = v;
    }
  }
  __syncthreads();
  // per-query threshold terms: mu = mean_k S[q,:] = Q[q,:] . Kbar   (dagl.py:256)
  for (int j = 0; j < AT_BM / 8; ++j) {
    int q = warp * (AT_BM / 8) + j;
    double s = 0.0;
    for (int e = lane; e < ED; e += 32) s += (double)Qs[q * ED + e] * (double)__ldg(Kbar + (size_t)img * ED + e);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      int qq = q0 + q;
      float mu = (float)s;
      tA[q] = (qq < g.Nq) ? mu * __ldg(gamma + (size_t)img * g.Nq + qq) : 0.f;
      tB[q] = (qq < g.Nq) ? __ldg(beta + (size_t)img * g.Nq + qq) : 0.f;
    }
  }

  // S-phase mapping: 16x16 threads, thread = 2 queries x 4 keys
  const int ty = tid >> 4, tx = tid & 15;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  int cnt[2] = {0, 0};
  // PV-phase mapping: thread = one (c,dy) halo row x 16 queries x 7 dx
  const int qg = tid / AT_TROWS, trow = tid % AT_TROWS;
  float acc[16][KS];
#pragma unroll
  for (int q = 0; q < 16; ++q)
#pragma unroll
    for (int d = 0; d < KS; ++d) acc[q][d] = 0.f;

  const int NT = g.H * ntx;
  const int t_begin = (int)(((long long)split * NT) / nsplit);
  const int t_end = (int)(((long long)(split + 1) * NT) / nsplit);
  const int nwords = (g.Nk + 31) / 32;

  for (int t = t_begin; t < t_end; ++t) {
    const int ky = t / ntx, x0 = (t % ntx) * tw;
    const int len = min(tw, g.W - x0);
    __syncthreads();
    // ---- key tile: len rows of K are one contiguous chunk -------------------
    {
      const float4* src = reinterpret_cast<const float4*>(K + ((size_t)img * g.Nk + (size_t)ky * g.W + x0) * ED);
      float4* Ks4 = reinterpret_cast<float4*>(Ks);
      const int nvalid = len * (ED / 4);
      for (int i = tid; i < AT_BN * (ED / 4); i += AT_THREADS)
        Ks4[i] = (i < nvalid) ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // ---- value halo: theta rows ky-3..ky+3, cols x0-3..x0+66 ----------------
    for (int i = tid; i < AT_TROWS * 70; i += AT_THREADS) {
      int r = i / 70, xx = i % 70;
      int c = r / KS, dy = r % KS;
      int yy = ky + dy - PADK, xg = x0 + xx - PADK;
      float v = 0.f;
      if (yy >= 0 && yy < g.H && xg >= 0 && xg < g.W)
        v = __ldg(theta + (((size_t)img * CI + c) * g.H + yy) * g.W + xg);
      Ts[r * AT_TROW + xx] = v;
    }
    __syncthreads();

    // ---- scores -------------------------------------------------------------
    float s_acc[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) s_acc[r][i] = 0.f;
    {
      const float4* qa = reinterpret_cast<const float4*>(Qs) + (2 * ty) * (ED / 4);
      const float4* qb = qa + (ED / 4);
      const float4* kp = reinterpret_cast<const float4*>(Ks) + tx * (ED / 4);
#pragma unroll 7
      for (int e4 = 0; e4 < ED / 4; ++e4) {
        const float4 a0 = qa[e4], a1 = qb[e4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 kv = kp[i * 16 * (ED / 4) + e4];
          s_acc[0][i] = fmaf(a0.x, kv.x, s_acc[0][i]); s_acc[0][i] = fmaf(a0.y, kv.y, s_acc[0][i]);
          s_acc[0][i] = fmaf(a0.z, kv.z, s_acc[0][i]); s_acc[0][i] = fmaf(a0.w, kv.w, s_acc[0][i]);
          s_acc[1][i] = fmaf(a1.x, kv.x, s_acc[1][i]); s_acc[1][i] = fmaf(a1.y, kv.y, s_acc[1][i]);
          s_acc[1][i] = fmaf(a1.z, kv.z, s_acc[1][i]); s_acc[1][i] = fmaf(a1.w, kv.w, s_acc[1][i]);
        }
      }
    }
    // ---- neighbour mask + online softmax (dagl.py:256-261) -------------------
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int q = 2 * ty + r;
      const float a_ = tA[q], b_ = tB[q];
      float ev[4];
      bool mk[4];
      float tmax = -INFINITY;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool valid = (tx + 16 * i) < len;
        const float s = s_acc[r][i];
        const float rl = fmaxf((s - a_) + b_, 0.f);      // relu(S - mu*gamma + beta)
        mk[i] = valid && (rl != 0.f);                    // mask_b
        ev[i] = (s * rl) * scale;                        // (S*mask)*softmax_scale
        if (valid) tmax = fmaxf(tmax, ev[i]);
      }
      tmax = half_warp_max(tmax);
      const float m_new = fmaxf(m_run[r], tmax);
      const float sc = (m_run[r] == -INFINITY) ? 0.f : expf(m_run[r] - m_new);
      float psum = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool valid = (tx + 16 * i) < len;
        const float p = valid ? expf(ev[i] - m_new) : 0.f;
        psum += p;                                       // denominator counts masked keys too
        Ps[q * AT_BN + tx + 16 * i] = mk[i] ? p : 0.f;   // numerator only neighbours
        cnt[r] += mk[i] ? 1 : 0;
      }
      psum = half_warp_sum(psum);
      l_run[r] = l_run[r] * sc + psum;
      m_run[r] = m_new;
      if (tx == 0) scs[q] = sc;
      if (mask_bits != nullptr) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const unsigned bal = __ballot_sync(0xffffffffu, mk[i]);
          const unsigned bits16 = (bal >> (16 * ((lane >> 4) & 1))) & 0xffffu;
          const int qq = q0 + q;
          if (tx == 0 && bits16 != 0u && qq < g.Nq) {
            const int kg = ky * g.W + x0 + 16 * i;
            const int wd = kg >> 5, sh = kg & 31;
            uint32_t* row = mask_bits + ((size_t)img * g.Nq + qq) * nwords;
            atomicOr(row + wd, bits16 << sh);
            const unsigned hi = (sh > 16) ? (bits16 >> (32 - sh)) : 0u;
            if (hi != 0u) atomicOr(row + wd + 1, hi);
          }
        }
      }
    }
    __syncthreads();

    // ---- aggregation: acc[q][c,dy,dx] += P[q][k] * theta_pad[c][ky+dy][kx+dx] ---
    if (tid < AT_PV_THREADS) {
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float sc = scs[qg * 16 + q];
        if (sc != 1.f) {
#pragma unroll
          for (int d = 0; d < KS; ++d) acc[q][d] *= sc;
        }
      }
      const float* tr = Ts + trow * AT_TROW;
      const float4* P4 = reinterpret_cast<const float4*>(Ps) + (qg * 16) * (AT_BN / 4);
      float win[10];
#pragma unroll
      for (int d = 0; d < 6; ++d) win[d] = tr[d];
      const int len4 = (len + 3) >> 2;
      for (int k4 = 0; k4 < len4; ++k4) {
#pragma unroll
        for (int d = 0; d < 4; ++d) win[6 + d] = tr[k4 * 4 + 6 + d];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float4 p = P4[q * (AT_BN / 4) + k4];
#pragma unroll
          for (int d = 0; d < KS; ++d) {
            float v = acc[q][d];
            v = fmaf(p.x, win[d], v); v = fmaf(p.y, win[d + 1], v);
            v = fmaf(p.z, win[d + 2], v); v = fmaf(p.w, win[d + 3], v);
            acc[q][d] = v;
          }
        }
#pragma unroll
        for (int d = 0; d < 6; ++d) win[d] = win[d + 4];
      }
    }
  }

  // ---- partial results ------------------------------------------------------
  const size_t prow = ((size_t)img * nsplit + split) * g.Nq;
  if (tid < AT_PV_THREADS) {
    const int c = trow / KS, dy = trow % KS;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int qq = q0 + qg * 16 + q;
      if (qq < g.Nq) {
        float* o = Opart + (prow + qq) * VD + c * KK + dy * KS;
#pragma unroll
        for (int d = 0; d < KS; ++d) o[d] = acc[q][d];
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    int total = cnt[r];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    const int qq = q0 + 2 * ty + r;
    if (tx == 0 && qq < g.Nq) {
      mpart[prow + qq] = m_run[r];
      lpart[prow + qq] = l_run[r];
      if (nnz != nullptr) atomicAdd(nnz + (size_t)img * g.Nq + qq, total);
    }
  }
}

// ---- host side ---------------------------------------------------------------
static void simt_tiling(const Geom& g, int* nsplit, int* ntx, int* tw) {
  *ntx = (g.W + AT_BN - 1) / AT_BN;
  *tw = (g.W + *ntx - 1) / *ntx;
  const int NT = g.H * (*ntx);
  const long long base = (long long)g.B * ((g.Nq + AT_BM - 1) / AT_BM);
  const int smax = NT < 16 ? NT : 16;
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= smax; ++s) {
    const long long ctas = base * s;
    const double waves = (double)((ctas + 147) / 148);
    const double cost = waves / s + 0.01 * s;     // time ~ waves * (work per CTA ~ 1/s); mild penalty on partial traffic
    if (cost < best_cost - 1e-12) { best_cost = cost; best = s; }
  }
  *nsplit = best;
}

static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

size_t attend_simt_workspace_bytes(const Geom& g) {
  int ns, ntx, tw;
  simt_tiling(g, &ns, &ntx, &tw);
  const size_t rows = (size_t)g.B * ns * g.Nq;
  return align_up(rows * VD * sizeof(float)) + 3 * align_up(rows * sizeof(float)) + align_up(merge_fold_scratch_bytes(g));
}

int launch_attend_simt(const Geom& g, const AttendArgs& a, cudaStream_t st) {
  int ns, ntx, tw;
  simt_tiling(g, &ns, &ntx, &tw);
  if (a.ws_bytes < attend_simt_workspace_bytes(g)) {
    call_state().err = "attend workspace too small";
    return -3;
  }
  const size_t rows = (size_t)g.B * ns * g.Nq;
  char* p = static_cast<char*>(a.ws);
  float* Opart = reinterpret_cast<float*>(p); p += align_up(rows * VD * sizeof(float));
  float* mpart = reinterpret_cast<float*>(p); p += align_up(rows * sizeof(float));
  float* lpart = reinterpret_cast<float*>(p); p += align_up(rows * sizeof(float));
  float* coef = reinterpret_cast<float*>(p); p += align_up(rows * sizeof(float));
  float* Om = reinterpret_cast<float*>(p);

  const int nwords = (g.Nk + 31) / 32;
  if (a.mask_bits) DAGL_CUDA_OK(cudaMemsetAsync(a.mask_bits, 0, (size_t)g.B * g.Nq * nwords * sizeof(uint32_t), st));
  if (a.nnz) DAGL_CUDA_OK(cudaMemsetAsync(a.nnz, 0, (size_t)g.B * g.Nq * sizeof(int32_t), st));

  const size_t smem = (size_t)AT_SMEM_FLOATS * sizeof(float);
  DAGL_CUDA_OK(cudaFuncSetAttribute(attend_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((g.Nq + AT_BM - 1) / AT_BM, ns, g.B);
  if (int rc = prof_begin(st)) return rc;
  attend_simt_kernel<<<grid, AT_THREADS, smem, st>>>(g, a.Q, a.K, a.Kbar, a.gamma, a.beta, a.theta, a.scale,
                                                     ns, ntx, tw, Opart, mpart, lpart, a.mask_bits, a.nnz);
  DAGL_LAUNCH_CHECK();
  if (int rc = prof_end(st)) return rc;
  return launch_merge_fold(g, ns, Opart, mpart, lpart, coef, Om, a.y, /*log2_units=*/0, /*shift_major=*/0, 1.f, st);
}

}  // namespace dagl
