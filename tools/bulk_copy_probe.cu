// Per-SM throughput of cp.async.bulk (global/L2 -> shared) as a function of copy size and copies in flight (development
// probe).  One CTA per SM; one producer thread keeps D copies of S bytes in flight into a ring; every copy reads a different
// L2-resident address.  Prints bytes / clock / SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory"); } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(n), "r"(smem_u32(b)) : "memory");
}

__global__ void probe(const uint8_t* src, size_t src_bytes, int S, int D, int iters, int nprod, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);          // [D]
  uint8_t* ring = smem + 1024;
  if (threadIdx.x == 0) { for (int i = 0; i < D; ++i) mbar_init(bars + i, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
  if (warp < nprod && lane == 0) {
    // producer `warp` handles copies i with i % nprod == warp; consumer == producer here (waits for the copy issued D earlier)
    size_t off = ((size_t)blockIdx.x * 7919 * 4096) % (src_bytes - (size_t)S * 64);
    for (int i = warp; i < iters + D; i += nprod) {
      const int s = i % D;
      if (i >= D) mbar_wait(bars + s, ((i / D) - 1) & 1);
      if (i < iters) {
        mbar_expect(bars + s, S);
        bulk_g2s(ring + (size_t)s * S, src + off + (size_t)(i % 64) * S, S, bars + s);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
  const size_t src_bytes = 64ull << 20;
  uint8_t* src; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int sizes[] = {2048, 3584, 7168, 14336, 19968, 39936};
  const int depths[] = {1, 2, 4, 7, 14};
  printf("%8s %6s %6s %6s | %10s %10s\n", "bytes", "depth", "nprod", "grid", "B/clk/SM", "us/copy");
  for (int grid : {148, 1}) for (int nprod : {1, 2}) for (int S : sizes) for (int D : depths) {
    if ((size_t)S * D > 200 * 1024) continue;
    if (D % nprod) continue;
    const int iters = 2000;
    probe<<<grid, 128, 1024 + S * D, 0>>>(src, src_bytes, S, D, iters, nprod, cyc);   // warm
    probe<<<grid, 128, 1024 + S * D, 0>>>(src, src_bytes, S, D, iters, nprod, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
    printf("%8d %6d %6d %6d | %10.1f %10.3f\n", S, D, nprod, grid, (double)S * iters / avg, avg / iters / 1965.0);
  }
  return 0;
}
