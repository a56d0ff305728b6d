// Prologue kernels of the graph block: feature maps, per-query threshold/bias,
// and the learned patch embeddings (reference: CE.forward, DN_Gray/model/dagl.py:208-249).
//
// Nothing the reference materialises through nn.Unfold (dagl.py:216-239) is ever
// written to HBM here: fc1/fc2 applied to unfolded patches are evaluated as
// implicit-GEMM 7x7 convolutions that gather straight from the 16-channel map G.
#include "common.cuh"
#include "prologue.cuh"

namespace dagl {

// ---------------------------------------------------------------------------
// G = conv3x3(b; g) and Theta = conv1x1(b; theta)          (dagl.py:208-209)
// One thread per pixel, 32 accumulators; weights transposed into smem as
// [ci][tap][16] so each tap costs four broadcast LDS.128.
// ---------------------------------------------------------------------------
constexpr int FM_THREADS = 128;

__global__ void __launch_bounds__(FM_THREADS)
feature_maps_kernel(Geom g, const float* __restrict__ b, const float* __restrict__ g_w,
                    const float* __restrict__ g_b, const float* __restrict__ th_w,
                    const float* __restrict__ th_b, float* __restrict__ G, float* __restrict__ Th,
                    unsigned* __restrict__ absmax /*nullable: [B][AMAX_STRIDE]*/) {
  extern __shared__ float smem[];
  float* gw_s = smem;                       // [C][9][16]
  float* tw_s = smem + g.C * 9 * CI;        // [C][16]
  const int C = g.C;
  for (int i = threadIdx.x; i < C * 9 * CI; i += FM_THREADS) {
    int co = i % CI, t = (i / CI) % 9, ci = i / (CI * 9);
    gw_s[i] = g_w[(co * C + ci) * 9 + t];
  }
  for (int i = threadIdx.x; i < C * CI; i += FM_THREADS) {
    int co = i % CI, ci = i / CI;
    tw_s[i] = th_w[co * C + ci];
  }
  __syncthreads();

  const int img = blockIdx.y;
  const int p = blockIdx.x * FM_THREADS + threadIdx.x;
  const bool live = p < g.Nk;
  const int y = p / g.W, x = p % g.W;
  const float* bi = b + (size_t)img * C * g.Nk;

  float ag[CI], at[CI];
#pragma unroll
  for (int c = 0; c < CI; ++c) { ag[c] = g_b[c]; at[c] = th_b[c]; }

  for (int ci = 0; live && ci < C; ++ci) {
    const float* bc = bi + (size_t)ci * g.Nk;
    float v[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
      v[t] = (yy >= 0 && yy < g.H && xx >= 0 && xx < g.W) ? __ldg(bc + yy * g.W + xx) : 0.f;
    }
    const float4* w4 = reinterpret_cast<const float4*>(gw_s + ci * 9 * CI);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 w = w4[t * 4 + j];
        ag[4 * j + 0] = fmaf(v[t], w.x, ag[4 * j + 0]);
        ag[4 * j + 1] = fmaf(v[t], w.y, ag[4 * j + 1]);
        ag[4 * j + 2] = fmaf(v[t], w.z, ag[4 * j + 2]);
        ag[4 * j + 3] = fmaf(v[t], w.w, ag[4 * j + 3]);
      }
    }
    const float4* t4 = reinterpret_cast<const float4*>(tw_s + ci * CI);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float4 w = t4[j];
      at[4 * j + 0] = fmaf(v[4], w.x, at[4 * j + 0]);
      at[4 * j + 1] = fmaf(v[4], w.y, at[4 * j + 1]);
      at[4 * j + 2] = fmaf(v[4], w.z, at[4 * j + 2]);
      at[4 * j + 3] = fmaf(v[4], w.w, at[4 * j + 3]);
    }
  }
  float tmax = 0.f, gmax = 0.f;
  if (live) {
    float* Go = G + (size_t)img * CI * g.Nk + p;
    float* To = Th + (size_t)img * CI * g.Nk + p;
#pragma unroll
    for (int c = 0; c < CI; ++c) {
      Go[(size_t)c * g.Nk] = ag[c];
      To[(size_t)c * g.Nk] = at[c];
      tmax = fmaxf(tmax, fabsf(at[c]));
      gmax = fmaxf(gmax, fabsf(ag[c]));
    }
  }
  if (absmax != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
      gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMax(absmax + img * AMAX_STRIDE + AMAX_THETA, __float_as_uint(tmax));
      atomicMax(absmax + img * AMAX_STRIDE + AMAX_G, __float_as_uint(gmax));
    }
  }
}

int launch_feature_maps(const Geom& g, const float* b, const float* g_w, const float* g_b,
                        const float* th_w, const float* th_b, float* G, float* Th, unsigned* absmax, cudaStream_t st) {
  size_t smem = (size_t)(g.C * 9 * CI + g.C * CI) * sizeof(float);
  if (smem > 48 * 1024)
    DAGL_CUDA_OK(cudaFuncSetAttribute(feature_maps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((g.Nk + FM_THREADS - 1) / FM_THREADS, g.B);
  feature_maps_kernel<<<grid, FM_THREADS, smem, st>>>(g, b, g_w, g_b, th_w, th_b, G, Th, absmax);
  DAGL_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// gamma = thr_conv(pad4(b)), beta = bias_conv(pad4(b))     (dagl.py:213-215)
// CTA = 32 consecutive queries (lanes) x 16 channel groups (warps); the SAME padding is a predicate, not a
// copy; both 7x7 filters are staged in smem and read as warp-uniform broadcasts.  The channel-group
// partials are summed in a fixed order (deterministic).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(GB_THREADS)
gamma_beta_kernel(Geom g, const float* __restrict__ b, const float* __restrict__ thr_w,
                  const float* __restrict__ thr_b, const float* __restrict__ bias_w,
                  const float* __restrict__ bias_b, float* __restrict__ gamma, float* __restrict__ beta) {
  extern __shared__ float smem[];
  gamma_beta_body(g, b, thr_w, thr_b, bias_w, bias_b, gamma, beta, smem, blockIdx.x, blockIdx.y);
}

int launch_gamma_beta(const Geom& g, const float* b, const float* thr_w, const float* thr_b,
                      const float* bias_w, const float* bias_b, float* gamma, float* beta, cudaStream_t st) {
  const size_t smem = gamma_beta_smem_bytes(g.C);
  if (smem > 48 * 1024)
    DAGL_CUDA_OK(cudaFuncSetAttribute(gamma_beta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((g.Nq + 31) / 32, g.B);
  gamma_beta_kernel<<<grid, GB_THREADS, smem, st>>>(g, b, thr_w, thr_b, bias_w, bias_b, gamma, beta);
  DAGL_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// Patch embedding  out[p][e] = relu(fc_b[e] + sum_{c,ky,kx} fc_w[e][c,ky,kx] * Gpad[c][oy*s-off_y+ky][ox*s-off_x+kx])
//   queries: s=4, off=SAME pad   (fc1 on unfold(pad4(G)),  dagl.py:216-221,248)
//   keys   : s=1, off=3          (fc2 on unfold(pad3(G)),  dagl.py:233-239,249)
// fp32 CUDA-core implicit GEMM: CTA = 64 positions x 196(224) outputs, K = 784 in 16 channel chunks.
// Optionally emits per-CTA column sums for Kbar = mean_k K (row-mean trick).
// ---------------------------------------------------------------------------
constexpr int EM_POS = 64;
constexpr int EM_THREADS = 256;
constexpr int EM_EP = 224;          // 196 padded to 7*32
constexpr int EM_WS_STRIDE = 225;   // odd stride: conflict-free transposed stores
constexpr int EM_WS_FLOATS = (KK * EM_WS_STRIDE + 3) & ~3;   // keep As 16-byte aligned

__global__ void __launch_bounds__(EM_THREADS, 2)
embed_kernel(Geom g, const float* __restrict__ G, const float* __restrict__ fc_w,
             const float* __restrict__ fc_b, float* __restrict__ out,
             int ny, int nx, int s, int off_y, int off_x, float* __restrict__ colsum_partial,
             unsigned* __restrict__ absmax /*nullable: [B][AMAX_STRIDE]*/, int absmax_slot) {
  extern __shared__ float smem[];
  float* Ws = smem;                              // [49][225]
  float* As = smem + EM_WS_FLOATS;               // [49][64]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int img = blockIdx.y;
  const int npos = ny * nx;
  const int p0 = blockIdx.x * EM_POS;
  const float* Gi = G + (size_t)img * CI * g.Nk;

  float acc[8][7];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 7; ++i) acc[j][i] = 0.f;

  // zero the e >= 196 padding columns once
  for (int i = tid; i < KK * (EM_EP - ED); i += EM_THREADS) {
    int kk = i / (EM_EP - ED), e = ED + i % (EM_EP - ED);
    Ws[kk * EM_WS_STRIDE + e] = 0.f;
  }

  for (int c = 0; c < CI; ++c) {
    __syncthreads();
    for (int i = tid; i < KK * ED; i += EM_THREADS) {
      int e = i / KK, kk = i % KK;
      Ws[kk * EM_WS_STRIDE + e] = __ldg(fc_w + (size_t)e * VD + c * KK + kk);
    }
    for (int i = tid; i < KK * EM_POS; i += EM_THREADS) {
      int kk = i / EM_POS, pp = i % EM_POS;
      int p = p0 + pp;
      float v = 0.f;
      if (p < npos) {
        int oy = p / nx, ox = p % nx;
        int yy = oy * s - off_y + kk / KS, xx = ox * s - off_x + kk % KS;
        if (yy >= 0 && yy < g.H && xx >= 0 && xx < g.W) v = __ldg(Gi + (size_t)c * g.Nk + yy * g.W + xx);
      }
      As[kk * EM_POS + pp] = v;
    }
    __syncthreads();
#pragma unroll 7
    for (int kk = 0; kk < KK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(As + kk * EM_POS + warp * 8);
      const float4 a1 = *reinterpret_cast<const float4*>(As + kk * EM_POS + warp * 8 + 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float w[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) w[i] = Ws[kk * EM_WS_STRIDE + lane + 32 * i];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 7; ++i) acc[j][i] = fmaf(a[j], w[i], acc[j][i]);
    }
  }

  // epilogue: bias + ReLU, store, optional column sums
  float csum[7];
  float vmax = 0.f;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const int e = lane + 32 * i;
    csum[i] = 0.f;
    if (e < ED) {
      const float bias = __ldg(fc_b + e);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int p = p0 + warp * 8 + j;
        if (p < npos) {
          float v = fmaxf(acc[j][i] + bias, 0.f);
          out[((size_t)img * npos + p) * ED + e] = v;
          csum[i] += v;
          vmax = fmaxf(vmax, v);
        }
      }
    }
  }
  if (absmax != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0) atomicMax(absmax + img * AMAX_STRIDE + absmax_slot, __float_as_uint(vmax));   // values are >= 0: uint order == float order
  }
  if (colsum_partial != nullptr) {
    __syncthreads();
    float* red = smem;                             // [8][224]
#pragma unroll
    for (int i = 0; i < 7; ++i) red[warp * EM_EP + lane + 32 * i] = csum[i];
    __syncthreads();
    for (int e = tid; e < ED; e += EM_THREADS) {
      float sum = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) sum += red[w8 * EM_EP + e];
      colsum_partial[((size_t)img * gridDim.x + blockIdx.x) * ED + e] = sum;
    }
  }
}

int embed_num_blocks(int npos) { return (npos + EM_POS - 1) / EM_POS; }

int launch_embed(const Geom& g, const float* G, const float* fc_w, const float* fc_b, float* out,
                 int ny, int nx, int s, int off_y, int off_x, float* colsum_partial, unsigned* absmax,
                 int absmax_slot, cudaStream_t st) {
  size_t smem = (size_t)(EM_WS_FLOATS + KK * EM_POS) * sizeof(float);
  DAGL_CUDA_OK(cudaFuncSetAttribute(embed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(embed_num_blocks(ny * nx), g.B);
  embed_kernel<<<grid, EM_THREADS, smem, st>>>(g, G, fc_w, fc_b, out, ny, nx, s, off_y, off_x, colsum_partial, absmax, absmax_slot);
  DAGL_LAUNCH_CHECK();
  return 0;
}

// Kbar[e] = (1/Nk) * sum over CTAs of the partial column sums; fixed summation order (32 interleaved
// fp64 chains per column, then a fixed tree), so the result is deterministic.  grid = (7, B): 32 columns
// per block, 1024 threads.
__global__ void __launch_bounds__(1024) kbar_kernel(const float* __restrict__ partial, int nblk, int Nk, float* __restrict__ Kbar) {
  pdl_prologue();
  __shared__ double acc_s[32][33];
  const int img = blockIdx.y;
  const int chain = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * 32 + lane;
  double s = 0.0;
  if (e < ED) {
#pragma unroll 4
    for (int k = chain; k < nblk; k += 32) s += (double)__ldg(partial + ((size_t)img * nblk + k) * ED + e);
  }
  acc_s[chain][lane] = s;
  __syncthreads();
  if (chain == 0 && e < ED) {
    double t = 0.0;
#pragma unroll
    for (int c = 0; c < 32; ++c) t += acc_s[c][lane];
    Kbar[(size_t)img * ED + e] = (float)(t / (double)Nk);
  }
}

int launch_kbar(const Geom& g, const float* colsum_partial, int nblk, float* Kbar, cudaStream_t st) {
  DAGL_CUDA_OK(launch_pdl(kbar_kernel, dim3((ED + 31) / 32, g.B), 1024, 0, st, colsum_partial, nblk, g.Nk, Kbar));
  DAGL_LAUNCH_CHECK();
  return 0;
}

}  // namespace dagl
