// Does a programmatic dependent launch start on idle SMs WHILE the primary grid is still running?
// primary: NA CTAs (cluster size CA, big smem: one CTA per SM) spin for ~300 us after griddepcontrol.launch_dependents;
// secondary: NB CTAs (cluster size CB, big smem) record their start time.  Build: nvcc -arch=sm_100a -o pdl_overlap_probe ...
#include <cstdio>
#include <cuda_runtime.h>
__device__ unsigned long long g_a_end, g_b_first = ~0ull, g_b_last;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void nop_kernel() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); asm volatile("griddepcontrol.wait;" ::: "memory"); }
__global__ void primary(int spin_us, int trigger) {
  extern __shared__ char sm[];
  if (trigger == 2) asm volatile("griddepcontrol.wait;" ::: "memory");      // wait for the predecessor first, then trigger
  if (trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const unsigned long long t0 = gtime();
  while (gtime() - t0 < (unsigned long long)spin_us * 1000ull) { sm[threadIdx.x] = 1; }
  if (threadIdx.x == 0) atomicMax(&g_a_end, gtime());
}
__global__ void secondary(int spin_us, int wait_first) {
  extern __shared__ char sm[];
  if (wait_first) asm volatile("griddepcontrol.wait;" ::: "memory");
  const unsigned long long t0 = gtime();
  if (threadIdx.x == 0) { atomicMin(&g_b_first, t0); atomicMax(&g_b_last, t0); }
  while (gtime() - t0 < (unsigned long long)spin_us * 1000ull) { sm[threadIdx.x] = 1; }
  if (!wait_first) asm volatile("griddepcontrol.wait;" ::: "memory");
}
static void launch(void (*k)(int, int), int n, int cluster, size_t smem, cudaStream_t st, bool pdl, int a0, int a1) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2]; int na = 0;
  if (pdl) { at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[na].val.programmaticStreamSerializationAllowed = 1; ++na; }
  if (cluster > 1) { at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = cluster; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1; ++na; }
  cfg.attrs = at; cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, a0, a1);
  if (e != cudaSuccess) printf("launch error %s\n", cudaGetErrorString(e));
}
int main() {
  const size_t smem = 200 * 1024;
  cudaFuncSetAttribute(primary, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(secondary, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(primary, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaStream_t st; cudaStreamCreate(&st);
  printf("%6s %4s %6s %4s %8s %8s | %14s %14s\n", "NA", "CA", "NB", "CB", "trigger", "b_wait", "B first - A end", "B last - A end");
  const int cfgs[][6] = {{128, 4, 20, 2, 1, 0}, {128, 4, 20, 1, 1, 0}, {128, 1, 20, 1, 1, 0}, {128, 1, 20, 2, 1, 0}, {128, 4, 64, 2, 1, 0},
                         {128, 4, 20, 2, 0, 0}, {128, 4, 20, 2, 1, 1}, {148, 1, 20, 1, 1, 0}, {132, 4, 16, 2, 1, 0},
                         {128, 4, 20, 2, 2, 0}, {128, 4, 20, 2, 2, 2}, {128, 4, 64, 2, 2, 2}};
  for (auto& c : cfgs) {
    for (int rep = 0; rep < 2; ++rep) {
      unsigned long long z = 0, m = ~0ull;
      cudaMemcpyToSymbol(g_a_end, &z, 8); cudaMemcpyToSymbol(g_b_last, &z, 8); cudaMemcpyToSymbol(g_b_first, &m, 8);
      cudaDeviceSynchronize();
      if (c[4] == 2) { cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(64); cfg.blockDim = dim3(128); cfg.stream = st;
        cudaLaunchAttribute at; at.id = cudaLaunchAttributeProgrammaticStreamSerialization; at.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &at; cfg.numAttrs = 1; cudaLaunchKernelEx(&cfg, nop_kernel); }
      launch(primary, c[0], c[1], smem, st, true, 300, c[4]);
      if (c[5] == 2) cudaFuncSetAttribute(secondary, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // host call between the launches
      launch(secondary, c[2], c[3], smem, st, true, 50, c[5] == 1);
      cudaError_t e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      unsigned long long a, bf, bl;
      cudaMemcpyFromSymbol(&a, g_a_end, 8); cudaMemcpyFromSymbol(&bf, g_b_first, 8); cudaMemcpyFromSymbol(&bl, g_b_last, 8);
      if (rep == 1)
        printf("%6d %4d %6d %4d %8d %8d | %11.1f us %11.1f us\n", c[0], c[1], c[2], c[3], c[4], c[5], ((double)bf - (double)a) / 1e3, ((double)bl - (double)a) / 1e3);
    }
  }
  return 0;
}
