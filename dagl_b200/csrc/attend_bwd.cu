// Backward of the fused graph stage (reference: autograd through dagl.py:250-272; the reference trains through plain
// autograd, DN_Gray/trainer.py:51-57).  fp32 CUDA-core kernels, deterministic (no float atomics).
//
// Given the embeddings Q [B][Nq][196], K [B][Nk][196] (post-ReLU), the value map theta [B][16][H][W], the per-query
// gamma / beta and the output gradient dy [B][16][H][W], computes dQ, dK, dtheta, dgamma, dbeta.  With
//   T_q = mu_q gamma_q - beta_q,  mu_q = mean_k S_qk,  rl = relu(S - T),  z = c S rl,  p = softmax_k(z),  P = p [rl > 0],
//   O = P V,  y = fold(O) / count
// the chain rule gives (gradient semantics of the reference: none through the 0/1 factor [rl > 0]):
//   dO   = unfold(dy / count)                                   dP_qk = dO_q . V_k
//   D_q  = sum_k P_qk dP_qk                                     dz_qk = p_qk ([rl > 0] dP_qk - D_q)
//   dS_qk (direct) = c dz_qk (rl + S [rl > 0])                  dT_q  = - sum_k [rl > 0] c S_qk dz_qk
//   dmu_q = gamma_q dT_q,  dgamma_q = mu_q dT_q,  dbeta_q = -dT_q
//   dQ = dS K + dmu Kbar^T        dK = dS^T Q + (1/Nk) 1 (sum_q dmu_q Q_q)^T        dV = P^T dO  ->  dtheta = fold of dV
// Every product is a dense GEMM over a chunk of query rows whose [rows x Nk] score / probability / dP matrices live in
// the workspace (bounded by the caller's workspace: the chunk shrinks for large images), so the kernels are one generic
// tiled fp32 GEMM (128 x 128 x 8, 8 x 8 per thread) with operand loaders (plain, transposed, and the Toeplitz view of theta for V) plus row-wise passes.
// The convolutions / linears in front of the graph stage are differentiated by PyTorch (dagl_b200/autograd.py).
#include <math.h>
#include <stdlib.h>
#include "../../include/dagl_b200.h"
#include "common.cuh"

namespace dagl {

constexpr int GB_M = 128, GB_N = 128, GB_K = 8, GB_THREADS_ = 256;
constexpr int TP = PADK;                         // theta is zero-padded by 3 on every side (dagl.py:224-230)

struct BwdGeom {
  int H, W, Nq, Nk, nqx, Wp6;                   // Wp6 = W + 6
  int rc;                                        // rows per chunk
  int r0, rows;                                  // current chunk
  long long img0;                                // first image of the current group
};

enum { BW_S = 0, BW_DP = 1, BW_DQ = 2, BW_DK = 3, BW_DV = 4 };

struct BwdPtrs {
  const float* Q; const float* K; const float* thpad; const float* dO;
  float* S; float* P; float* dP;                // chunk matrices [group image][rc][Nk]
  float* dQ; float* dK; float* dV;
};

// element loaders: A(m, k) and B(k, n) of the five products (z = image inside the group)
template <int MODE>
__device__ __forceinline__ float load_a(const BwdGeom& g, const BwdPtrs& p, int z, int m, int k, int M, int K) {
  if (m >= M || k >= K) return 0.f;
  const long long img = g.img0 + z;
  if (MODE == BW_S) return __ldg(p.Q + ((size_t)img * g.Nq + g.r0 + m) * ED + k);
  if (MODE == BW_DP) return __ldg(p.dO + ((size_t)img * g.Nq + g.r0 + m) * VD + k);
  if (MODE == BW_DQ) return p.S[((size_t)z * g.rc + m) * g.Nk + k];                  // dS (written over S)
  if (MODE == BW_DK) return p.S[((size_t)z * g.rc + k) * g.Nk + m];                  // dS^T
  return p.P[((size_t)z * g.rc + k) * g.Nk + m];                                     // BW_DV: P^T
}
template <int MODE>
__device__ __forceinline__ float load_b(const BwdGeom& g, const BwdPtrs& p, int z, int k, int n, int K, int N) {
  if (k >= K || n >= N) return 0.f;
  const long long img = g.img0 + z;
  if (MODE == BW_S) return __ldg(p.K + ((size_t)img * g.Nk + n) * ED + k);           // K^T
  if (MODE == BW_DP) {                                                               // V^T: Toeplitz view of padded theta
    const int sh = k >> 4, c = k & 15, dy = sh / KS, dx = sh % KS;
    const int ky = n / g.W, kx = n % g.W;
    return __ldg(p.thpad + (((size_t)img * (g.H + 2 * TP) + ky + dy) * g.Wp6 + kx + dx) * CI + c);
  }
  if (MODE == BW_DQ) return __ldg(p.K + ((size_t)img * g.Nk + k) * ED + n);
  if (MODE == BW_DK) return __ldg(p.Q + ((size_t)img * g.Nq + g.r0 + k) * ED + n);
  return __ldg(p.dO + ((size_t)img * g.Nq + g.r0 + k) * VD + n);                      // BW_DV
}

// C[M x N] (+)= A[M x K] B[K x N]: 128 x 128 x 8 tiles, 8 x 8 outputs per thread (64 FMAs per 4 shared-memory loads).
// AFK / BFK: the operand's k index is the contiguous one in memory (consecutive threads take consecutive k), else its m / n
// index is.
template <int MODE>
__global__ void __launch_bounds__(GB_THREADS_)
bwd_gemm_kernel(BwdGeom g, BwdPtrs p, int M, int N, int K, int accumulate) {
  constexpr bool AFK = (MODE == BW_S || MODE == BW_DP || MODE == BW_DQ);
  constexpr bool BFK = (MODE == BW_S || MODE == BW_DP);
  __shared__ __align__(16) float As[GB_K][GB_M + 4];
  __shared__ __align__(16) float Bs[GB_K][GB_N + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * GB_M, n0 = blockIdx.x * GB_N, z = blockIdx.z;
  float acc[8][8] = {};
  for (int k0 = 0; k0 < K; k0 += GB_K) {
#pragma unroll
    for (int r = 0; r < (GB_M * GB_K) / GB_THREADS_; ++r) {
      const int i = tid + r * GB_THREADS_;
      const int ak = AFK ? (i & (GB_K - 1)) : (i / GB_M), am = AFK ? (i / GB_K) : (i & (GB_M - 1));
      As[ak][am] = load_a<MODE>(g, p, z, m0 + am, k0 + ak, M, K);
      const int bk = BFK ? (i & (GB_K - 1)) : (i / GB_N), bn = BFK ? (i / GB_K) : (i & (GB_N - 1));
      Bs[bk][bn] = load_b<MODE>(g, p, z, k0 + bk, n0 + bn, K, N);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GB_K; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]), a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8]), b1 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const long long img = g.img0 + z;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tx * 8 + j;
      if (n >= N) continue;
      float* dst;
      if (MODE == BW_S) dst = p.S + ((size_t)z * g.rc + m) * g.Nk + n;
      else if (MODE == BW_DP) dst = p.dP + ((size_t)z * g.rc + m) * g.Nk + n;
      else if (MODE == BW_DQ) dst = p.dQ + ((size_t)img * g.Nq + g.r0 + m) * ED + n;
      else if (MODE == BW_DK) dst = p.dK + ((size_t)img * g.Nk + m) * ED + n;
      else dst = p.dV + ((size_t)z * g.Nk + m) * VD + n;
      *dst = accumulate ? *dst + acc[i][j] : acc[i][j];
    }
  }
}

// ---- small kernels ------------------------------------------------------------------------------------------------
// Kbar partial sums: part[b][pz][e] = sum over the keys of slice pz
__global__ void bwd_kbar_part_kernel(const float* __restrict__ K, int Nk, float* __restrict__ part) {
  const int e = threadIdx.x, pz = blockIdx.x, b = blockIdx.y, np = gridDim.x;
  if (e >= ED) return;
  const int k0 = (int)((long long)pz * Nk / np), k1 = (int)((long long)(pz + 1) * Nk / np);
  float s = 0.f;
  for (int k = k0; k < k1; ++k) s += __ldg(K + ((size_t)b * Nk + k) * ED + e);
  part[((size_t)b * np + pz) * ED + e] = s;
}
// Kbar = mean; mu_q = Q_q . Kbar; T_q = mu_q gamma_q - beta_q
__global__ void bwd_mu_kernel(const float* __restrict__ Q, const float* __restrict__ part, int np, int Nq, int Nk,
                              const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ Kbar,
                              float* __restrict__ mu, float* __restrict__ T) {
  __shared__ float kb[ED];
  const int b = blockIdx.y;
  for (int e = threadIdx.x; e < ED; e += blockDim.x) {
    double s = 0.0;
    for (int pz = 0; pz < np; ++pz) s += (double)part[((size_t)b * np + pz) * ED + e];
    kb[e] = (float)(s / (double)Nk);
    if (blockIdx.x == 0) Kbar[(size_t)b * ED + e] = kb[e];
  }
  __syncthreads();
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Nq) return;
  double s = 0.0;
  for (int e = 0; e < ED; ++e) s += (double)__ldg(Q + ((size_t)b * Nq + q) * ED + e) * (double)kb[e];
  const float m = (float)s;
  mu[(size_t)b * Nq + q] = m;
  T[(size_t)b * Nq + q] = m * gamma[(size_t)b * Nq + q] - beta[(size_t)b * Nq + q];
}
// dO[b][q][sh][c] = dy[b][c][py][px] / count(py, px)     (backward of fold + count normalisation, dagl.py:265-272)
__global__ void bwd_unfold_dy_kernel(Geom g, const float* __restrict__ dy, float* __restrict__ dO) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)g.B * g.Nq * VD;
  if (i >= total) return;
  const int c = (int)(i % CI), sh = (int)((i / CI) % KK), q = (int)((i / VD) % g.Nq), b = (int)(i / ((size_t)VD * g.Nq));
  const int qy = q / g.nqx, qx = q % g.nqx;
  const int py = qy * SQ - PADK + sh / KS, px = qx * SQ - PADK + sh % KS;
  float v = 0.f;
  if (py >= 0 && py < g.H && px >= 0 && px < g.W) {
    const int cy = min(g.nqy - 1, (py + PADK) >> 2) - (py >> 2) + 1, cx = min(g.nqx - 1, (px + PADK) >> 2) - (px >> 2) + 1;
    v = __ldg(dy + (((size_t)b * CI + c) * g.H + py) * g.W + px) / (float)(cy * cx);
  }
  dO[i] = v;
}
// thpad[b][y+3][x+3][c] = theta[b][c][y][x], zero border
__global__ void bwd_thpad_kernel(Geom g, const float* __restrict__ theta, float* __restrict__ thpad) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int Hp = g.H + 2 * TP, Wp = g.W + 2 * TP;
  const size_t total = (size_t)g.B * Hp * Wp * CI;
  if (i >= total) return;
  const int c = (int)(i % CI), x = (int)((i / CI) % Wp) - TP, y = (int)((i / ((size_t)CI * Wp)) % Hp) - TP, b = (int)(i / ((size_t)CI * Wp * Hp));
  thpad[i] = (y >= 0 && y < g.H && x >= 0 && x < g.W) ? __ldg(theta + (((size_t)b * CI + c) * g.H + y) * g.W + x) : 0.f;
}
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const float t = __shfl_xor_sync(0xffffffffu, v, o); v = is_max ? fmaxf(v, t) : v + t; }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  __syncthreads();
  return r;
}
// row pass 1: p = softmax_k(c S relu(S - T)) over ALL keys (masked keys contribute exp(0 - max)), written to P
__global__ void __launch_bounds__(256)
bwd_softmax_kernel(BwdGeom g, const float* __restrict__ T, float scale, const float* __restrict__ S, float* __restrict__ P) {
  __shared__ float red[8];
  const int m = blockIdx.x, z = blockIdx.y;
  if (m >= g.rows) return;
  const float t = T[(size_t)(g.img0 + z) * g.Nq + g.r0 + m];
  const float* s = S + ((size_t)z * g.rc + m) * g.Nk;
  float* p = P + ((size_t)z * g.rc + m) * g.Nk;
  float mx = 0.f;                                              // logits are >= 0 and every row has a key
  for (int k = threadIdx.x; k < g.Nk; k += 256) { const float v = s[k]; mx = fmaxf(mx, scale * v * fmaxf(v - t, 0.f)); }
  mx = block_reduce(mx, red, true);
  float sum = 0.f;
  for (int k = threadIdx.x; k < g.Nk; k += 256) {
    const float v = s[k], e = expf(scale * v * fmaxf(v - t, 0.f) - mx);
    p[k] = e;
    sum += e;
  }
  sum = block_reduce(sum, red, false);
  const float inv = 1.f / sum;
  for (int k = threadIdx.x; k < g.Nk; k += 256) p[k] *= inv;
}
// row pass 2: D, dz, dS (over S), P <- P [rl > 0] (for dV), dT -> dgamma, dbeta, dmu
__global__ void __launch_bounds__(256)
bwd_dz_kernel(BwdGeom g, const float* __restrict__ T, const float* __restrict__ mu, const float* __restrict__ gamma, float scale,
              float* __restrict__ S, float* __restrict__ P, const float* __restrict__ dP, float* __restrict__ dgamma,
              float* __restrict__ dbeta, float* __restrict__ dmu) {
  __shared__ float red[8];
  const int m = blockIdx.x, z = blockIdx.y;
  if (m >= g.rows) return;
  const size_t qi = (size_t)(g.img0 + z) * g.Nq + g.r0 + m;
  const float t = T[qi];
  float* s = S + ((size_t)z * g.rc + m) * g.Nk;
  float* p = P + ((size_t)z * g.rc + m) * g.Nk;
  const float* dp = dP + ((size_t)z * g.rc + m) * g.Nk;
  float d = 0.f;
  for (int k = threadIdx.x; k < g.Nk; k += 256)
    if (s[k] - t > 0.f) d += p[k] * dp[k];
  d = block_reduce(d, red, false);
  float dt = 0.f;
  for (int k = threadIdx.x; k < g.Nk; k += 256) {
    const float v = s[k], rl = fmaxf(v - t, 0.f), pk = p[k];
    const bool keep = rl > 0.f;
    const float dz = pk * ((keep ? dp[k] : 0.f) - d);
    s[k] = scale * dz * (rl + (keep ? v : 0.f));               // dS (direct part)
    if (keep) dt -= scale * v * dz;
    p[k] = keep ? pk : 0.f;
  }
  dt = block_reduce(dt, red, false);
  if (threadIdx.x == 0) {
    dgamma[qi] = mu[qi] * dt;
    dbeta[qi] = -dt;
    dmu[qi] = gamma[qi] * dt;
  }
}
// rank-one terms of the row mean: dQ_q += dmu_q Kbar;  dkb[b][e] = (1/Nk) sum_q dmu_q Q_qe  (added to every dK row below)
__global__ void bwd_mean_terms_kernel(int Nq, int Nk, const float* __restrict__ Q, const float* __restrict__ Kbar,
                                      const float* __restrict__ dmu, float* __restrict__ dQ, float* __restrict__ dkb) {
  const int b = blockIdx.x, e = threadIdx.x;
  if (e >= ED) return;
  const float kb = Kbar[(size_t)b * ED + e];
  double s = 0.0;
  for (int q = 0; q < Nq; ++q) {
    const size_t i = ((size_t)b * Nq + q) * ED + e;
    const float dm = dmu[(size_t)b * Nq + q];
    dQ[i] += dm * kb;
    s += (double)dm * (double)__ldg(Q + i);
  }
  dkb[(size_t)b * ED + e] = (float)(s / (double)Nk);
}
__global__ void bwd_add_dkb_kernel(size_t n_per_img, const float* __restrict__ dkb, float* __restrict__ dK) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_per_img) return;
  dK[(size_t)blockIdx.y * n_per_img + i] += dkb[(size_t)blockIdx.y * ED + (i % ED)];
}
// dtheta[b][c][y][x] = sum over the shifts of dV[key (y + 3 - dy, x + 3 - dx)][sh][c]   (transpose of the Toeplitz view)
__global__ void bwd_fold_dv_kernel(Geom g, long long img0, int nimg, const float* __restrict__ dV, float* __restrict__ dtheta) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)nimg * CI * g.Nk;
  if (i >= total) return;
  const int x = (int)(i % g.W), y = (int)((i / g.W) % g.H), c = (int)((i / g.Nk) % CI), z = (int)(i / ((size_t)CI * g.Nk));
  float s = 0.f;
  for (int dy = 0; dy < KS; ++dy) {
    const int ky = y + TP - dy;
    if (ky < 0 || ky >= g.H) continue;
    for (int dx = 0; dx < KS; ++dx) {
      const int kx = x + TP - dx;
      if (kx < 0 || kx >= g.W) continue;
      s += __ldg(dV + ((size_t)z * g.Nk + (size_t)ky * g.W + kx) * VD + (dy * KS + dx) * CI + c);
    }
  }
  dtheta[(((size_t)(img0 + z) * CI + c) * g.H + y) * g.W + x] = s;
}

// ---- host ---------------------------------------------------------------------------------------------------------
static inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }
constexpr int KB_PARTS = 32;

struct BwdWs { size_t part, Kbar, mu, T, dmu, dkb, dO, thpad, S, P, dP, dV, total; int group, rc; };

// images per group / rows per chunk so that the chunk matrices stay under ~768 MB and dV under ~1 GB
static BwdWs bwd_ws(const Geom& g) {
  BwdWs w;
  const size_t per_img_dv = (size_t)g.Nk * VD * 4;
  int group = (int)((size_t)(1ull << 30) / per_img_dv);
  if (group < 1) group = 1;
  if (group > g.B) group = g.B;
  long long rc = (long long)((size_t)(256ull << 20) / ((size_t)group * g.Nk * 4));
  if (const char* e = getenv("DAGL_BWD_RC")) { group = 1; rc = atoi(e); }      // tests: force several row chunks / image groups
  if (rc > g.Nq) rc = g.Nq;
  if (rc < 64) rc = g.Nq < 64 ? g.Nq : 64;
  rc = (rc + 63) / 64 * 64;
  w.group = group; w.rc = (int)rc;
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off += al(b); return o; };
  w.part = take((size_t)g.B * KB_PARTS * ED * 4);
  w.Kbar = take((size_t)g.B * ED * 4);
  w.mu = take((size_t)g.B * g.Nq * 4);
  w.T = take((size_t)g.B * g.Nq * 4);
  w.dmu = take((size_t)g.B * g.Nq * 4);
  w.dkb = take((size_t)g.B * ED * 4);
  w.dO = take((size_t)g.B * g.Nq * VD * 4);
  w.thpad = take((size_t)g.B * (g.H + 2 * TP) * (g.W + 2 * TP) * CI * 4);
  const size_t chunk = (size_t)group * rc * g.Nk * 4;
  w.S = take(chunk); w.P = take(chunk); w.dP = take(chunk);
  w.dV = take((size_t)group * per_img_dv);
  w.total = off;
  return w;
}

template <int MODE>
static int gemm(const BwdGeom& bg, const BwdPtrs& p, int M, int N, int K, int nimg, int accumulate, cudaStream_t st) {
  dim3 grid((N + GB_N - 1) / GB_N, (M + GB_M - 1) / GB_M, nimg);
  bwd_gemm_kernel<MODE><<<grid, GB_THREADS_, 0, st>>>(bg, p, M, N, K, accumulate);
  DAGL_LAUNCH_CHECK();
  return 0;
}

}  // namespace dagl

using namespace dagl;

extern "C" {

size_t dagl_graph_attend_backward_workspace_bytes(int32_t B, int32_t H, int32_t W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return bwd_ws(make_geom(B, 64, H, W)).total;
}

int32_t dagl_graph_attend_backward_f32(const float* Q, const float* K, const float* theta, const float* gamma, const float* beta,
                                       const float* dy, float* dQ, float* dK, float* dtheta, float* dgamma, float* dbeta,
                                       int32_t B, int32_t H, int32_t W, float softmax_scale, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  call_state().launches = 0;
  call_state().impl = "bwd";
  if (B <= 0 || H <= 0 || W <= 0 || (long long)H * W > (1 << 22)) { call_state().err = "bad shape"; return DAGL_ERR_INVALID_ARG; }
  if (!Q || !K || !theta || !gamma || !beta || !dy || !dQ || !dK || !dtheta || !dgamma || !dbeta || !workspace) {
    call_state().err = "null buffer";
    return DAGL_ERR_INVALID_ARG;
  }
  const Geom g = make_geom(B, 64, H, W);
  const BwdWs w = bwd_ws(g);
  if (workspace_bytes < w.total) { call_state().err = "workspace too small"; return DAGL_ERR_WORKSPACE; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(workspace);
  float* part = reinterpret_cast<float*>(base + w.part);
  float* Kbar = reinterpret_cast<float*>(base + w.Kbar);
  float* mu = reinterpret_cast<float*>(base + w.mu);
  float* T = reinterpret_cast<float*>(base + w.T);
  float* dmu = reinterpret_cast<float*>(base + w.dmu);
  float* dkb = reinterpret_cast<float*>(base + w.dkb);
  float* dO = reinterpret_cast<float*>(base + w.dO);
  float* thpad = reinterpret_cast<float*>(base + w.thpad);

  bwd_kbar_part_kernel<<<dim3(KB_PARTS, B), 224, 0, st>>>(K, g.Nk, part);
  DAGL_LAUNCH_CHECK();
  bwd_mu_kernel<<<dim3((g.Nq + 255) / 256, B), 256, 0, st>>>(Q, part, KB_PARTS, g.Nq, g.Nk, gamma, beta, Kbar, mu, T);
  DAGL_LAUNCH_CHECK();
  {
    const size_t n = (size_t)B * g.Nq * VD;
    bwd_unfold_dy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, dy, dO);
    DAGL_LAUNCH_CHECK();
    const size_t n2 = (size_t)B * (H + 2 * TP) * (W + 2 * TP) * CI;
    bwd_thpad_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(g, theta, thpad);
    DAGL_LAUNCH_CHECK();
  }
  BwdPtrs p;
  p.Q = Q; p.K = K; p.thpad = thpad; p.dO = dO;
  p.S = reinterpret_cast<float*>(base + w.S); p.P = reinterpret_cast<float*>(base + w.P); p.dP = reinterpret_cast<float*>(base + w.dP);
  p.dQ = dQ; p.dK = dK; p.dV = reinterpret_cast<float*>(base + w.dV);
  BwdGeom bg;
  bg.H = H; bg.W = W; bg.Nq = g.Nq; bg.Nk = g.Nk; bg.nqx = g.nqx; bg.Wp6 = W + 2 * TP; bg.rc = w.rc;
  for (int i0 = 0; i0 < B; i0 += w.group) {
    const int nimg = B - i0 < w.group ? B - i0 : w.group;
    bg.img0 = i0;
    int first = 1;
    for (int r0 = 0; r0 < g.Nq; r0 += w.rc) {
      bg.r0 = r0; bg.rows = g.Nq - r0 < w.rc ? g.Nq - r0 : w.rc;
      int rc;
      if ((rc = gemm<BW_S>(bg, p, bg.rows, g.Nk, ED, nimg, 0, st))) return rc;
      bwd_softmax_kernel<<<dim3(bg.rows, nimg), 256, 0, st>>>(bg, T, softmax_scale, p.S, p.P);
      DAGL_LAUNCH_CHECK();
      if ((rc = gemm<BW_DP>(bg, p, bg.rows, g.Nk, VD, nimg, 0, st))) return rc;
      bwd_dz_kernel<<<dim3(bg.rows, nimg), 256, 0, st>>>(bg, T, mu, gamma, softmax_scale, p.S, p.P, p.dP, dgamma, dbeta, dmu);
      DAGL_LAUNCH_CHECK();
      if ((rc = gemm<BW_DQ>(bg, p, bg.rows, ED, g.Nk, nimg, 0, st))) return rc;
      if ((rc = gemm<BW_DK>(bg, p, g.Nk, ED, bg.rows, nimg, first ? 0 : 1, st))) return rc;
      if ((rc = gemm<BW_DV>(bg, p, g.Nk, VD, bg.rows, nimg, first ? 0 : 1, st))) return rc;
      first = 0;
    }
    const size_t n = (size_t)nimg * CI * g.Nk;
    bwd_fold_dv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, i0, nimg, p.dV, dtheta);
    DAGL_LAUNCH_CHECK();
  }
  bwd_mean_terms_kernel<<<B, 224, 0, st>>>(g.Nq, g.Nk, Q, Kbar, dmu, dQ, dkb);
  DAGL_LAUNCH_CHECK();
  {
    const size_t n = (size_t)g.Nk * ED;
    bwd_add_dkb_kernel<<<dim3((unsigned)((n + 255) / 256), B), 256, 0, st>>>(n, dkb, dK);
    DAGL_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
