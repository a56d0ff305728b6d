// tc_probe — on-device validation of the tcgen05 building blocks the fused graph
// kernel relies on, and MMA issue-rate measurements for its two MMA shapes.
//   P1  SS MMA, A and B K-major, no-swizzle core-matrix layout, 2 k-steps
//   P2  SS MMA, B MN-major no-swizzle read at pixel-shifted start addresses (Toeplitz value operand)
//   P3  TS MMA, A (packed fp16 pairs) written to TMEM with tcgen05.st
//   P4  issue-rate / throughput of the S-shaped and PV-shaped MMA streams
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tc_probe tools/tc_probe.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include "../dagl_b200/csrc/tc_utils.cuh"

using namespace dagl::tc;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

// byte offset of element (row, k) in a K-major no-swizzle tile with `rows` rows
__host__ __device__ inline uint32_t kmajor_off(int row, int k, int rows) {
  return (uint32_t)((k / 8) * (rows / 8) * 128 + (row / 8) * 128 + (row % 8) * 16 + (k % 8) * 2);
}

// ---------------------------------------------------------------------------------
// P1/P2/P3 functional kernel.  128 threads.  Results: D[128][ncols] fp32.
//   mode 1: D[:, 0:48]   = A(128x32) * B1(48x32)^T           (SS, K-major, 2 k-steps)
//   mode 2: D[:, 16i..]  = A(128x16) * Theta[:, s_i + k]     (SS, B MN-major shifted), i over 5 shifts
//   mode 3: same as mode 2 but A read from TMEM (TS)
// ---------------------------------------------------------------------------------
constexpr int TH_PIX = 64;                         // theta tile pixels
__constant__ int c_shifts[5] = {0, 1, 3, 7, 13};

__global__ void __launch_bounds__(128) probe_func(int mode, const __half* __restrict__ A /*[128][32]*/,
                                                  const __half* __restrict__ B1 /*[48][32]*/,
                                                  const __half* __restrict__ Th /*[16][64]*/, float* __restrict__ D,
                                                  int ncols) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                  // 128x32 fp16 K-major: 4 chunks * 16 rowgroups * 128 B = 8192
  uint8_t* sB = smem + 8192;           // 48x32: 4 * 6 * 128 = 3072
  uint8_t* sT = smem + 8192 + 3072;    // theta [2 chunks][64 px][8 ch] fp16 = 2048
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < 128 * 32; i += 128) {
    int r = i / 32, k = i % 32;
    *reinterpret_cast<__half*>(sA + kmajor_off(r, k, 128)) = A[i];
  }
  for (int i = tid; i < 48 * 32; i += 128) {
    int r = i / 32, k = i % 32;
    *reinterpret_cast<__half*>(sB + kmajor_off(r, k, 48)) = B1[i];
  }
  for (int i = tid; i < 16 * TH_PIX; i += 128) {
    int c = i / TH_PIX, p = i % TH_PIX;
    *reinterpret_cast<__half*>(sT + (c / 8) * (TH_PIX * 16) + p * 16 + (c % 8) * 2) = Th[i];
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init_fence();
  }
  if (warp == 0) tmem_alloc<128>(&tmem_base_s);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  if (mode == 3) {
    // A (first 16 k) -> TMEM columns [96, 104): lane = row, column j = {A[row][2j], A[row][2j+1]}
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) {
      __half2 h = __halves2half2(A[tid * 32 + 2 * j], A[tid * 32 + 2 * j + 1]);
      v[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    tmem_st8(tbase + ((uint32_t)(warp * 32) << 16) + 96, v);
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  if (tid == 0) {
    if (mode == 1) {
      const uint32_t idesc = instr_desc(128, 48, FMT_F16, FMT_F16, 0, 0);
      for (int ks = 0; ks < 2; ++ks) {
        uint64_t ad = smem_desc(smem_u32(sA) + ks * 2 * (16 * 128), 16 * 128, 128);
        uint64_t bd = smem_desc(smem_u32(sB) + ks * 2 * (6 * 128), 6 * 128, 128);
        mma_f16_ss(tbase, ad, bd, idesc, ks > 0);
      }
    } else {
      const uint32_t idesc = instr_desc(128, 16, FMT_F16, FMT_F16, 0, 1);
      for (int i = 0; i < 5; ++i) {
        uint64_t bd = smem_desc(smem_u32(sT) + c_shifts[i] * 16, /*LBO: next 8 pixels*/ 128, /*SBO: next 8 channels*/ TH_PIX * 16);
        if (mode == 2) {
          uint64_t ad = smem_desc(smem_u32(sA), 16 * 128, 128);
          mma_f16_ss(tbase + 16 * i, ad, bd, idesc, 0);
        } else {
          mma_f16_ts(tbase + 16 * i, tbase + 96, bd, idesc, 0);
        }
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  // read back: thread = row
  for (int c0 = 0; c0 < ncols; c0 += 8) {
    uint32_t v[8];
    tmem_ld8(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_wait_ld();
    for (int j = 0; j < 8; ++j) D[tid * ncols + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tbase);
}

// ---------------------------------------------------------------------------------
// P4 timing kernel: one CTA per SM, thread 0 issues; depth-2 commit pipeline.
//   mode 0: 39 x (M128 N48 K16 SS)           "S stream" per 48-key tile
//   mode 1: 75 x (M128 N16 K16 TS, shifted B) "PV stream" (25 shifts x 3 k-steps)
//   mode 2: mode 0 followed by mode 1        (one key tile of the fused kernel)
//   mode 3: 16 x (M128 N256 K16 SS)          reference shape
//   mode 4: 75 x (M128 N16 K16 SS)           PV stream with A from smem
//   mode 5: 38 x (M128 N32 K16 TS)           PV stream, 2 shifts fused?  (N=32 probe)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) probe_time(int mode, int iters, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  // 64 KB of small fp16 values
  for (int i = tid; i < 32768; i += 128) reinterpret_cast<__half*>(smem)[i] = __float2half(((i * 37) % 17 - 8) * 0.0625f);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init_fence();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  {  // P operand region in TMEM columns [448, 472)
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) v[j] = 0x2c002c00u;  // small fp16 pairs
    for (int c = 0; c < 24; c += 8) tmem_st8(tbase + ((uint32_t)(warp * 32) << 16) + 448 + c, v);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp_id_uniform() == 0 && elect_one()) {
    const uint32_t sa = smem_u32(smem);            // A / Q region (128 x 208 would be 52 KB; reuse)
    const uint32_t sb = smem_u32(smem) + 32768;    // B / K region
    const uint32_t idS = instr_desc(128, 48, FMT_F16, FMT_F16, 0, 0);
    const uint32_t idP = instr_desc(128, 16, FMT_F16, FMT_F16, 0, 1);
    const uint32_t idP32 = instr_desc(128, 32, FMT_F16, FMT_F16, 0, 1);
    const uint32_t idR = instr_desc(128, 256, FMT_F16, FMT_F16, 0, 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode == 0 || mode == 2) {
        for (int ks = 0; ks < 13; ++ks)
          for (int part = 0; part < 3; ++part) {
            uint64_t ad = smem_desc(sa + ((ks * 2) % 8) * 2048 + (part & 1) * 16384, 2048, 128);
            uint64_t bd = smem_desc(sb + ((ks * 2) % 8) * 768 + (part >> 1) * 8192, 768, 128);
            mma_f16_ss(tbase + 400 + 48 * 0, ad, bd, idS, 1);
          }
      }
      if (mode == 1 || mode == 2 || mode == 4) {
        for (int sft = 0; sft < 25; ++sft)
          for (int ks = 0; ks < 3; ++ks) {
            uint64_t bd = smem_desc(sb + 16384 + (sft * 3 + ks * 8) * 16, 128, 1024);
            if (mode == 4) {
              uint64_t ad = smem_desc(sa + ks * 4096, 2048, 128);
              mma_f16_ss(tbase + 16 * sft, ad, bd, idP, 1);
            } else {
              mma_f16_ts(tbase + 16 * sft, tbase + 448 + 8 * ks, bd, idP, 1);
            }
          }
      }
      if (mode == 3) {
        for (int ks = 0; ks < 16; ++ks) {
          uint64_t ad = smem_desc(sa + (ks % 8) * 4096, 2048, 128);
          uint64_t bd = smem_desc(sb + (ks % 4) * 8192, 4096, 128);
          mma_f16_ss(tbase, ad, bd, idR, 1);
        }
      }
      if (mode == 5) {
        for (int sft = 0; sft < 13; ++sft)
          for (int ks = 0; ks < 3; ++ks) {
            uint64_t bd = smem_desc(sb + 16384 + (sft * 3 + ks * 8) * 16, 128, 1024);
            mma_f16_ts(tbase + 32 * sft, tbase + 448 + 8 * ks, bd, idP32, 1);
          }
      }
      mma_commit(&bar[it & 1]);
      if (it > 0) mbar_wait(&bar[(it - 1) & 1], ((it - 1) >> 1) & 1);
    }
    mbar_wait(&bar[(iters - 1) & 1], ((iters - 1) >> 1) & 1);
    long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

// ---------------------------------------------------------------------------------
// P5: parametrised issue-rate sweep.  cnt MMAs (M=128, N, K=16) per iteration.
//   b_mn   : B operand MN-major (shifted theta view) or K-major
//   a_tmem : A from TMEM (TS) or smem (SS)
//   ndist  : number of distinct accumulators the stream rotates over (1 = fully dependent chain)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) probe_rate(int N, int b_mn, int a_tmem, int ndist, int cnt, int iters,
                                                  long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32768; i += 128) reinterpret_cast<__half*>(smem)[i] = __float2half(((i * 37) % 17 - 8) * 0.0625f);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init_fence();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  {
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) v[j] = 0x2c002c00u;
    for (int c = 0; c < 32; c += 8) tmem_st8(tbase + ((uint32_t)(warp * 32) << 16) + 480 + c, v);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp_id_uniform() == 0 && elect_one()) {
    const uint32_t sa = smem_u32(smem);
    const uint32_t sb = smem_u32(smem) + 32768;
    const uint32_t idesc = instr_desc(128, N, FMT_F16, FMT_F16, 0, b_mn);
    const uint32_t b_lbo = b_mn ? 128 : (N / 8) * 128, b_sbo = b_mn ? 1024 : 128;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int i = 0; i < cnt; ++i) {
        const uint32_t d = tbase + (i % ndist) * N;
        uint64_t bd = smem_desc(sb + (i % 4) * (b_mn ? 16 : 8192), b_lbo, b_sbo);
        if (a_tmem) {
          mma_f16_ts(d, tbase + 480 + 8 * (i % 4), bd, idesc, 1);
        } else {
          uint64_t ad = smem_desc(sa + (i % 4) * 4096, 2048, 128);
          mma_f16_ss(d, ad, bd, idesc, 1);
        }
      }
      mma_commit(&bar[it & 1]);
      if (it > 0) mbar_wait(&bar[(it - 1) & 1], ((it - 1) >> 1) & 1);
    }
    mbar_wait(&bar[(iters - 1) & 1], ((iters - 1) >> 1) & 1);
    long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

// ---------------------------------------------------------------------------------
// P6/P7: wide-N Toeplitz value operand.
//   mode 6: theta tile [64 px][16 ch] (32 B / pixel), MN-major SWIZZLE_32B, N-group stride (LBO) = 32 B = one
//           pixel, so one MMA with N = 16*G covers G consecutive dx shifts x 16 channels.  The tile is stored
//           with the 32B-swizzle XOR (address bit 4 ^= bit 7) applied on the absolute smem address.
//   mode 7: theta tile [2 chunks][64 px][8 ch], no swizzle, N-group stride (SBO) = 16 B = one pixel: N = 8*G covers
//           G shifts of one 8-channel chunk.
//   variant: bit0 -> set descriptor base_offset = (start >> 7) & 7
// D[m][g*CW + c] = sum_k A[m][k] * theta[c][shift + g + k]
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t smem_desc_ex(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout,
                                                 uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

__global__ void __launch_bounds__(128) probe_wide(int mode, int variant, int shift, const __half* __restrict__ A,
                                                  const __half* __restrict__ Th, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;             // 8192
  uint8_t* sT = smem + 8192;      // 2048 (1024-aligned)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 32; i += 128) {
    int r = i / 32, k = i % 32;
    *reinterpret_cast<__half*>(sA + kmajor_off(r, k, 128)) = A[i];
  }
  for (int i = tid; i < 16 * TH_PIX; i += 128) {
    int c = i / TH_PIX, p = i % TH_PIX;
    uint32_t off;
    if (mode == 6) {
      off = p * 32 + c * 2;
      uint32_t abs_addr = smem_u32(sT) + off;
      off ^= ((abs_addr >> 7) & 1) << 4;       // Swizzle<1,4,3>
    } else {
      off = (c / 8) * (TH_PIX * 16) + p * 16 + (c % 8) * 2;
    }
    *reinterpret_cast<__half*>(sT + off) = Th[i];
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
  if (warp == 0) tmem_alloc<128>(&tmem_base_s);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  if (warp_id_uniform() == 0 && elect_one()) {
    uint64_t ad = smem_desc(smem_u32(sA), 16 * 128, 128);
    if (mode == 6) {
      const uint32_t start = smem_u32(sT) + shift * 32;
      const uint32_t bo = (variant & 1) ? ((start >> 7) & 7) : 0;
      uint64_t bd = smem_desc_ex(start, /*LBO*/ 32, /*SBO*/ 256, /*SWIZZLE_32B*/ 6, bo);
      mma_f16_ss(tbase, ad, bd, instr_desc(128, 112, FMT_F16, FMT_F16, 0, 1), 0);
    } else {
      const uint32_t start = smem_u32(sT) + shift * 16;
      uint64_t bd = smem_desc_ex(start, /*LBO: K groups*/ 128, /*SBO: N groups*/ 16, 0, 0);
      mma_f16_ss(tbase, ad, bd, instr_desc(128, 48, FMT_F16, FMT_F16, 0, 1), 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 112; c0 += 8) {
    uint32_t v[8];
    tmem_ld8(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_wait_ld();
    for (int j = 0; j < 8; ++j) D[tid * 112 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tbase);
}

// compile-time-shaped rate kernel: CNT MMAs of (M=128, N, K=16) per iteration
template <int N, int A_TMEM, int B_MODE /*0 K-major, 1 MN none, 2 MN sw32*/, int NDIST>
__global__ void __launch_bounds__(128) probe_rate_ct(int iters, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32768; i += 128) reinterpret_cast<__half*>(smem)[i] = __float2half(((i * 37) % 17 - 8) * 0.0625f);
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init_fence(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  {
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) v[j] = 0x2c002c00u;
    for (int c = 0; c < 32; c += 8) tmem_st8(tbase + ((uint32_t)(warp * 32) << 16) + 480 + c, v);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  constexpr int CNT = 24;
  if (warp_id_uniform() == 0 && elect_one()) {
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem) + 32768;
    constexpr uint32_t idesc = instr_desc(128, N, FMT_F16, FMT_F16, 0, B_MODE != 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < CNT; ++i) {
        const uint32_t d = tbase + (i % NDIST) * N;
        uint64_t bd;
        if (B_MODE == 0) bd = smem_desc(sb + (i % 3) * 8192, (N / 8) * 128, 128);
        else if (B_MODE == 1) bd = smem_desc_ex(sb + (i % 7) * 16, 128, 16, 0, 0);
        else bd = smem_desc_ex(sb + (i % 7) * 32, 32, 256, 6, 0);
        if (A_TMEM) mma_f16_ts(d, tbase + 480 + 8 * (i % 3), bd, idesc, 1);
        else mma_f16_ss(d, smem_desc(sa + (i % 3) * 4096, 2048, 128), bd, idesc, 1);
      }
      mma_commit(&bar[it & 1]);
      if (it > 0) mbar_wait(&bar[(it - 1) & 1], ((it - 1) >> 1) & 1);
    }
    mbar_wait(&bar[(iters - 1) & 1], ((iters - 1) >> 1) & 1);
    cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

template <int N, int A_TMEM, int B_MODE, int NDIST>
static void run_rate_ct(const char* name, long long* dC) {
  CK(cudaFuncSetAttribute(probe_rate_ct<N, A_TMEM, B_MODE, NDIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  const int iters = 1000;
  probe_rate_ct<N, A_TMEM, B_MODE, NDIST><<<148, 128, 65536>>>(20, dC);
  probe_rate_ct<N, A_TMEM, B_MODE, NDIST><<<148, 128, 65536>>>(iters, dC);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("P8 %s failed: %s\n", name, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> cyc(148);
  CK(cudaMemcpy(cyc.data(), dC, 148 * 8, cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (auto c : cyc) mx = c > mx ? c : mx;
  printf("P8 %-34s N=%3d : %7.1f cyc/MMA (ideal %5.1f)\n", name, N, (double)mx / iters / 24, N / 2.0);
}


// P9: does tcgen05.mma issue run ahead of execution?  Issue CNT MMAs (N = 112, SS, MN-major sw32 B, A from smem with the
// collector flags of the P.V loop when COLL != 0), read the clock, commit, read the clock, wait for completion.
template <int CNT, int COLL>
__global__ void __launch_bounds__(128) probe_issue(long long* __restrict__ out /*[grid][4]*/) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32768; i += 128) reinterpret_cast<__half*>(smem)[i] = __float2half(((i * 37) % 17 - 8) * 0.0625f);
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init_fence(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  if (warp_id_uniform() == 0 && elect_one()) {
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem) + 32768;
    constexpr uint32_t idesc = instr_desc(128, 112, FMT_F16, FMT_F16, 0, 1);
    long long t_issue = 0, t_commit = 0, t_drain = 0, t_wait_done = 0;
    for (int it = 0; it < 6; ++it) {
      const long long t0 = clock64();
#pragma unroll
      for (int i = 0; i < CNT; ++i) {
        const uint32_t d = tbase + (i % 3) * 112;
        const uint64_t bd = smem_desc_ex(sb + (i % 7) * 32, 32, 256, 6, 0);
        const uint64_t ad = smem_desc(sa + ((i / 2) % 3) * 4096, 2048, 128);
        if (COLL == 0) mma_f16_ss(d, ad, bd, idesc, 1);
        else if ((i & 1) == 0) mma_f16_ss_a_fill(d, ad, bd, idesc, 1);
        else mma_f16_ss_a_lastuse(d, ad, bd, idesc, 1);
      }
      const long long t1 = clock64();
      mma_commit(&bar[it & 1]);
      const long long t2 = clock64();
      mbar_wait(&bar[it & 1], (it >> 1) & 1);
      const long long t3 = clock64();
      mbar_wait(&bar[it & 1], (it >> 1) & 1);            // already complete: cost of a passing try_wait
      const long long t4 = clock64();
      if (it >= 2) { t_issue += t1 - t0; t_commit += t2 - t1; t_drain += t3 - t2; t_wait_done += t4 - t3; }
    }
    out[blockIdx.x * 4 + 0] = t_issue / 4; out[blockIdx.x * 4 + 1] = t_commit / 4;
    out[blockIdx.x * 4 + 2] = t_drain / 4; out[blockIdx.x * 4 + 3] = t_wait_done / 4;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

template <int CNT, int COLL>
static void run_issue(long long* dC) {
  CK(cudaFuncSetAttribute(probe_issue<CNT, COLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  probe_issue<CNT, COLL><<<8, 128, 65536>>>(dC);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("P9 failed: %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[4];
  CK(cudaMemcpy(h, dC, sizeof(h), cudaMemcpyDeviceToHost));
  printf("P9 %2d MMAs (N=112 SS%s): issue %5lld cyc (%5.1f/MMA)  commit %4lld  drain after commit %5lld  passing try_wait %3lld   [exec model %d]\n",
         CNT, COLL ? ", A collector" : "", h[0], (double)h[0] / CNT, h[1], h[2], h[3], CNT * 58);
}

static float h2f(__half h) { return __half2float(h); }

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device: %s, %d SMs, cc %d.%d\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor);

  std::vector<__half> A(128 * 32), B1(48 * 32), Th(16 * TH_PIX);
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < 32; ++k) A[m * 32 + k] = __float2half((float)((m * 3 + k * 5) % 7 - 3) * 0.25f);
  for (int n = 0; n < 48; ++n)
    for (int k = 0; k < 32; ++k) B1[n * 32 + k] = __float2half((float)((n * 5 + k) % 5 - 2) * 0.5f);
  for (int c = 0; c < 16; ++c)
    for (int p = 0; p < TH_PIX; ++p) Th[c * TH_PIX + p] = __float2half((float)((c * 7 + p * 3) % 11 - 5) * 0.125f);
  __half *dA, *dB, *dT;
  float* dD;
  CK(cudaMalloc(&dA, A.size() * 2));
  CK(cudaMalloc(&dB, B1.size() * 2));
  CK(cudaMalloc(&dT, Th.size() * 2));
  CK(cudaMalloc(&dD, 128 * 128 * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B1.data(), B1.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dT, Th.data(), Th.size() * 2, cudaMemcpyHostToDevice));
  const int shifts[5] = {0, 1, 3, 7, 13};

  for (int mode = 1; mode <= 3; ++mode) {
    const int ncols = (mode == 1) ? 48 : 80;
    CK(cudaMemset(dD, 0xff, 128 * 128 * 4));
    probe_func<<<1, 128, 8192 + 3072 + 2048>>>(mode, dA, dB, dT, dD, ncols);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("P%d: kernel failed: %s\n", mode, cudaGetErrorString(e));
      return 1;
    }
    std::vector<float> D(128 * ncols);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    int bad = 0, fm = -1, fn = -1;
    float fg = 0, fe = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < ncols; ++n) {
        double ref = 0;
        if (mode == 1) {
          for (int k = 0; k < 32; ++k) ref += (double)h2f(A[m * 32 + k]) * h2f(B1[n * 32 + k]);
        } else {
          int i = n / 16, c = n % 16;
          for (int k = 0; k < 16; ++k) ref += (double)h2f(A[m * 32 + k]) * h2f(Th[c * TH_PIX + shifts[i] + k]);
        }
        double err = fabs(ref - D[m * ncols + n]);
        if (!(err <= 1e-4)) {
          if (bad == 0) { fm = m; fn = n; fg = D[m * ncols + n]; fe = (float)ref; }
          bad++;
        }
        if (err > maxerr) maxerr = err;
      }
    printf("P%d: %s  max_err=%.3g  mismatches=%d", mode, bad ? "FAIL" : "PASS", maxerr, bad);
    if (bad) printf("  first at (m=%d,n=%d): got %g expected %g", fm, fn, fg, fe);
    printf("\n");
    if (bad) {
      printf("   row0 got:");
      for (int n = 0; n < 16 && n < ncols; ++n) printf(" %g", D[n]);
      printf("\n");
    }
  }

  // ---- timing ----
  long long* dC;
  CK(cudaMalloc(&dC, 148 * 8));
  CK(cudaFuncSetAttribute(probe_time, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  const char* names[6] = {"S stream  39x(128x48x16 SS)", "PV stream 75x(128x16x16 TS)", "S+PV (one key tile)",
                          "ref 16x(128x256x16 SS)", "PV stream 75x(128x16x16 SS)", "PV32 39x(128x32x16 TS)"};
  const double macs[6] = {39.0 * 128 * 48 * 16, 75.0 * 128 * 16 * 16, 39.0 * 128 * 48 * 16 + 75.0 * 128 * 16 * 16,
                          16.0 * 128 * 256 * 16, 75.0 * 128 * 16 * 16, 39.0 * 128 * 32 * 16};
  for (int grid : {1, 148}) {
    for (int mode = 0; mode < 6; ++mode) {
      const int iters = 2000;
      probe_time<<<grid, 128, 65536>>>(mode, 50, dC);   // warm
      probe_time<<<grid, 128, 65536>>>(mode, iters, dC);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("P4 mode %d failed: %s\n", mode, cudaGetErrorString(e));
        return 1;
      }
      std::vector<long long> cyc(grid);
      CK(cudaMemcpy(cyc.data(), dC, grid * 8, cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (auto c : cyc) mx = c > mx ? c : mx;
      double per = (double)mx / iters;
      printf("P4 grid=%3d %-30s : %8.1f cyc/iter  -> %7.1f MAC/clk/SM (fp16 peak 4096)\n", grid, names[mode], per,
             macs[mode] / per);
    }
  }

  // ---- P6 / P7 wide-N Toeplitz operand ----
  {
    float* dW;
    CK(cudaMalloc(&dW, 128 * 112 * 4));
    for (int mode = 6; mode <= 7; ++mode)
      for (int variant = 0; variant < (mode == 6 ? 2 : 1); ++variant)
        for (int shift : {0, 1, 4, 5, 11}) {
          const int G = mode == 6 ? 7 : 6, CW = mode == 6 ? 16 : 8, ncols = G * CW;
          CK(cudaMemset(dW, 0xff, 128 * 112 * 4));
          probe_wide<<<1, 128, 8192 + 2048>>>(mode, variant, shift, dA, dT, dW);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("P%d failed: %s\n", mode, cudaGetErrorString(e)); return 1; }
          std::vector<float> Dw(128 * 112);
          CK(cudaMemcpy(Dw.data(), dW, Dw.size() * 4, cudaMemcpyDeviceToHost));
          int bad = 0, fm = -1, fn = -1; float fg = 0, fe = 0;
          for (int m = 0; m < 128; ++m)
            for (int n = 0; n < ncols; ++n) {
              int g = n / CW, c = n % CW;
              double ref = 0;
              for (int k = 0; k < 16; ++k) ref += (double)h2f(A[m * 32 + k]) * h2f(Th[c * TH_PIX + shift + g + k]);
              if (!(fabs(ref - Dw[m * 112 + n]) <= 1e-4)) { if (!bad) { fm = m; fn = n; fg = Dw[m * 112 + n]; fe = (float)ref; } bad++; }
            }
          printf("P%d variant=%d shift=%2d: %s mismatches=%d", mode, variant, shift, bad ? "FAIL" : "PASS", bad);
          if (bad) printf(" first (m=%d,n=%d) got %g exp %g", fm, fn, fg, fe);
          printf("\n");
        }
  }
  if (argc > 1 && std::string(argv[1]) == "p9") {
    run_issue<1, 0>(dC); run_issue<2, 0>(dC); run_issue<4, 0>(dC); run_issue<6, 0>(dC); run_issue<12, 0>(dC);
    run_issue<24, 0>(dC); run_issue<48, 0>(dC); run_issue<6, 1>(dC); run_issue<24, 1>(dC);
    return 0;
  }
  // ---- P8 compile-time rate probes ----
  run_rate_ct<48, 0, 0, 1>("S-like  SS Kmajor dep-chain", dC);
  run_rate_ct<48, 0, 0, 2>("S-like  SS Kmajor 2 accs", dC);
  run_rate_ct<96, 0, 0, 1>("SS Kmajor", dC);
  run_rate_ct<128, 0, 0, 1>("SS Kmajor", dC);
  run_rate_ct<16, 1, 1, 8>("PV TS MN-none shifts", dC);
  run_rate_ct<48, 1, 1, 4>("PV TS MN-none SBO16", dC);
  run_rate_ct<64, 1, 1, 4>("PV TS MN-none SBO16", dC);
  run_rate_ct<112, 1, 2, 3>("PV TS MN-sw32", dC);
  run_rate_ct<112, 0, 2, 3>("PV SS MN-sw32", dC);
  run_rate_ct<64, 1, 2, 4>("PV TS MN-sw32", dC);
  run_rate_ct<48, 1, 2, 4>("PV TS MN-sw32", dC);
  run_rate_ct<112, 1, 2, 1>("PV TS MN-sw32 dep-chain", dC);
  return 0;
  // ---- P5 rate sweep (runtime-parametrised; issue-bound, kept for reference) ----
  CK(cudaFuncSetAttribute(probe_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  printf("P5: cycles per MMA (M=128,K=16 fp16); ideal = N/2\n");
  printf("   N  Bmajor A     ndist  cyc/MMA   ideal\n");
  for (int N : {16, 32, 48, 64, 96, 128, 192, 256})
    for (int b_mn = 0; b_mn < 2; ++b_mn)
      for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
        for (int nd = 0; nd < 3; ++nd) {
          int ndist = nd == 0 ? 1 : (nd == 1 ? 2 : 8);
          if (ndist * N > 448) continue;
          if (b_mn && N > 64) continue;      // MN-major probe region is only 8 chunks wide
          const int cnt = 64, iters = 500;
          probe_rate<<<148, 128, 65536>>>(N, b_mn, a_tmem, ndist, cnt, 20, dC);
          probe_rate<<<148, 128, 65536>>>(N, b_mn, a_tmem, ndist, cnt, iters, dC);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("P5 failed: %s\n", cudaGetErrorString(e)); return 1; }
          std::vector<long long> cyc(148);
          CK(cudaMemcpy(cyc.data(), dC, 148 * 8, cudaMemcpyDeviceToHost));
          long long mx = 0;
          for (auto c : cyc) mx = c > mx ? c : mx;
          printf("  %3d  %s     %s  %2d    %8.1f  %6.1f\n", N, b_mn ? "MN" : "K ", a_tmem ? "tmem" : "smem", ndist,
                 (double)mx / iters / cnt, N / 2.0);
        }
  return 0;
}
