// Merge of the key-split partial results and the fold epilogue, shared by the fp32 and the
// tensor-core graph kernels (reference: dagl.py:265-272).
#include <math.h>
#include "common.cuh"

namespace dagl {

// coef[s][q] = w_s / sum_s w_s l_s with w_s = exp(m_s - max_s m_s): log-sum-exp merge of the key splits.
// `log2_units`: the running maxima are in log2 units (tensor-core kernel) instead of natural log.
__global__ void merge_coef_kernel(int B, int Nq, int nsplit, int log2_units, float out_scale,
                                  const float* __restrict__ mpart, const float* __restrict__ lpart,
                                  float* __restrict__ coef) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Nq) return;
  const int img = i / Nq, q = i % Nq;
  float M = -INFINITY;
  for (int s = 0; s < nsplit; ++s) M = fmaxf(M, mpart[((size_t)img * nsplit + s) * Nq + q]);
  float L = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const size_t j = ((size_t)img * nsplit + s) * Nq + q;
    const float w = log2_units ? exp2f(mpart[j] - M) : expf(mpart[j] - M);
    L += w * lpart[j];
  }
  const float inv = out_scale / L;
  for (int s = 0; s < nsplit; ++s) {
    const size_t j = ((size_t)img * nsplit + s) * Nq + q;
    const float w = log2_units ? exp2f(mpart[j] - M) : expf(mpart[j] - M);
    coef[j] = w * inv;
  }
}

// Omerged[q][d] = sum_s coef[s][q] * Opart[s][q][d]   (streaming, float4)
__global__ void __launch_bounds__(256)
merge_rows_kernel(int B, int Nq, int nsplit, int q_begin, int q_end, const float* __restrict__ Opart,
                  const float* __restrict__ coef, float* __restrict__ Om) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // over B*(q_end-q_begin)*(VD/4)
  const int nq = q_end - q_begin;
  const size_t total = (size_t)B * nq * (VD / 4);
  if (i >= total) return;
  const int d4 = (int)(i % (VD / 4));
  const size_t bqr = i / (VD / 4);
  const int img = (int)(bqr / nq), q = q_begin + (int)(bqr % nq);
  const size_t bq = (size_t)img * Nq + q;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < nsplit; ++s) {
    const size_t j = ((size_t)img * nsplit + s) * Nq + q;
    const float c = __ldg(coef + j);
    const float4 v = __ldg(reinterpret_cast<const float4*>(Opart + j * VD) + d4);
    acc.x = fmaf(c, v.x, acc.x); acc.y = fmaf(c, v.y, acc.y); acc.z = fmaf(c, v.z, acc.z); acc.w = fmaf(c, v.w, acc.w);
  }
  reinterpret_cast<float4*>(Om + bq * VD)[d4] = acc;
}

// y[c][py][px] = (1/cnt) * sum over the <=2x2 queries whose folded 7x7 patch covers the pixel
// (F.fold with kernel 7, padding 3, stride 4, then / coverage count: dagl.py:265-272).
// Row layout: shift_major=0 -> [c][dy][dx] (reference order), 1 -> [dy*7+dx][c].
__global__ void fold_kernel(Geom g, int shift_major, const float* __restrict__ Om, float* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = g.B * CI * g.Nk;
  if (i >= total) return;
  const int px = i % g.W, py = (i / g.W) % g.H, c = (i / g.Nk) % CI, img = i / (CI * g.Nk);
  const int qy_lo = py >> 2, qy_hi = min(g.nqy - 1, (py + PADK) >> 2);
  const int qx_lo = px >> 2, qx_hi = min(g.nqx - 1, (px + PADK) >> 2);
  float sum = 0.f;
  for (int qy = qy_lo; qy <= qy_hi; ++qy)
    for (int qx = qx_lo; qx <= qx_hi; ++qx) {
      const int q = qy * g.nqx + qx;
      const int sh = (py - (qy * SQ - PADK)) * KS + (px - (qx * SQ - PADK));
      const int d = shift_major ? sh * CI + c : c * KK + sh;
      sum += __ldg(Om + ((size_t)img * g.Nq + q) * VD + d);
    }
  const float cntf = (float)((qy_hi - qy_lo + 1) * (qx_hi - qx_lo + 1));
  y[g.y_offset(img) + (size_t)c * g.Nk + (size_t)py * g.W + px] = sum / cntf;
}

// Fused merge + fold for the fixed-reference kernels (shift-major partial rows): every output pixel gathers its <= 2x2
// covering queries straight from the key-split partials and scales them by 1 / (sum of the row-sum partials): merge
// coefficients, merge and fold + coverage normalisation (dagl.py:265-272) in ONE pass.  Thread = (pixel, 4 channels) with
// the channel quad fastest, so four neighbouring threads read the 64 contiguous bytes of one (query, shift) record.
// Fixed summation order, no atomics.  lpart: [B][nsplit][nparts][Nq].
__global__ void __launch_bounds__(256)
fold_partials_kernel(Geom g, int nsplit, int nparts, const float* __restrict__ Opart, const float* __restrict__ lpart,
                     float* __restrict__ y) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = g.B * 4 * g.Nk;
  if (i >= total) return;
  const int c4 = i & 3, pix = (i >> 2) % g.Nk, img = (i >> 2) / g.Nk;
  const int px = pix % g.W, py = pix / g.W;
  const int qy_lo = py >> 2, qy_hi = min(g.nqy - 1, (py + PADK) >> 2);
  const int qx_lo = px >> 2, qx_hi = min(g.nqx - 1, (px + PADK) >> 2);
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int qy = qy_lo; qy <= qy_hi; ++qy)
    for (int qx = qx_lo; qx <= qx_hi; ++qx) {
      const int q = qy * g.nqx + qx;
      const int sh = (py - (qy * SQ - PADK)) * KS + (px - (qx * SQ - PADK));
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      float L = 0.f;
#pragma unroll 4
      for (int s = 0; s < nsplit; ++s) {
        const size_t row = ((size_t)img * nsplit + s) * g.Nq + q;
        const float4 v = __ldg(reinterpret_cast<const float4*>(Opart + row * VD + sh * CI) + c4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        for (int h = 0; h < nparts; ++h) L += __ldg(lpart + (((size_t)img * nsplit + s) * nparts + h) * g.Nq + q);
      }
      // L > 0 whenever the row has a valid key (the row maximum itself contributes ~2^12); the guard keeps a degenerate
      // row from turning into inf * 0 = NaN
      const float c = L > 0.f ? 1.f / L : 0.f;
      sum.x = fmaf(c, acc.x, sum.x); sum.y = fmaf(c, acc.y, sum.y); sum.z = fmaf(c, acc.z, sum.z); sum.w = fmaf(c, acc.w, sum.w);
    }
  const float inv = 1.f / (float)((qy_hi - qy_lo + 1) * (qx_hi - qx_lo + 1));
  float* yo = y + g.y_offset(img) + (size_t)(4 * c4) * g.Nk + pix;
  yo[0] = sum.x * inv; yo[(size_t)g.Nk] = sum.y * inv; yo[2 * (size_t)g.Nk] = sum.z * inv; yo[3 * (size_t)g.Nk] = sum.w * inv;
}

// `serialize`: the partials come from two grids that ran concurrently (attend_tc.cu, hybrid launch): a plain launch, which
// starts after everything in front of it in the stream, instead of a programmatic dependent of the last one only
int launch_fold_partials(const Geom& g, int nsplit, int nparts, const float* Opart, const float* lpart, float* y, cudaStream_t st,
                         bool serialize) {
  const int total = g.B * 4 * g.Nk;
  if (serialize)
    fold_partials_kernel<<<(total + 255) / 256, 256, 0, st>>>(g, nsplit, nparts, Opart, lpart, y);
  else
  DAGL_CUDA_OK(launch_pdl(fold_partials_kernel, (total + 255) / 256, 256, 0, st, g, nsplit, nparts, Opart, lpart, y));
  DAGL_LAUNCH_CHECK();
  return 0;
}

size_t merge_fold_scratch_bytes(const Geom& g) { return (size_t)g.B * g.Nq * VD * sizeof(float); }

// Omerged rows [q_begin, q_end) of every image <- sum over splits of coef * partial
int launch_merge_rows(const Geom& g, int nsplit, int q_begin, int q_end, const float* Opart, const float* coef,
                      float* Omerged, cudaStream_t st) {
  const size_t n4 = (size_t)g.B * (q_end - q_begin) * (VD / 4);
  merge_rows_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(g.B, g.Nq, nsplit, q_begin, q_end, Opart, coef, Omerged);
  DAGL_LAUNCH_CHECK();
  return 0;
}

// y = fold(rows) / coverage count
int launch_fold_rows(const Geom& g, const float* Omerged, float* y, int shift_major, cudaStream_t st) {
  const int total = g.B * CI * g.Nk;
  fold_kernel<<<(total + 255) / 256, 256, 0, st>>>(g, shift_major, Omerged, y);
  DAGL_LAUNCH_CHECK();
  return 0;
}

// merge (given coefficients) + fold
int launch_rows_fold(const Geom& g, int nsplit, const float* Opart, const float* coef, float* Omerged, float* y,
                     int shift_major, cudaStream_t st) {
  if (int rc = launch_merge_rows(g, nsplit, 0, g.Nq, Opart, coef, Omerged, st)) return rc;
  return launch_fold_rows(g, Omerged, y, shift_major, st);
}

int launch_merge_fold(const Geom& g, int nsplit, const float* Opart, const float* mpart, const float* lpart,
                      float* coef, float* Omerged, float* y, int log2_units, int shift_major, float out_scale,
                      cudaStream_t st) {
  const int nq_total = g.B * g.Nq;
  merge_coef_kernel<<<(nq_total + 255) / 256, 256, 0, st>>>(g.B, g.Nq, nsplit, log2_units, out_scale, mpart, lpart, coef);
  DAGL_LAUNCH_CHECK();
  return launch_rows_fold(g, nsplit, Opart, coef, Omerged, y, shift_major, st);
}

}  // namespace dagl
