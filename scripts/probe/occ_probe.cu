// Probe: does the occupancy calculator limit kernels that use tcgen05.alloc / mbarriers / printf to one CTA per SM?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int COLS, bool USE_TMEM, bool USE_PRINTF>
__global__ void __launch_bounds__(128, 2) k(int* out, int spin) {
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
  if (USE_TMEM && threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  while (clock64() - t0 < spin) {}
  if (USE_PRINTF && spin < 0) printf("never\n");
  unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (threadIdx.x == 0) out[blockIdx.x] = smid;
  __syncthreads();
  if (USE_TMEM && threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "n"(COLS) : "memory");
}
template <class K> void run(const char* name, K kern) {
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 128, 0);
  int* d; cudaMalloc(&d, 296 * 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  kern<<<296, 128>>>(d, 1000); cudaDeviceSynchronize();
  cudaEventRecord(a); kern<<<296, 128>>>(d, 2000000); cudaEventRecord(b); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("%-28s occupancy API %d CTAs/SM; 296 CTAs x 2M-cycle spin took %.2f ms (1 wave ~1.0 ms, 2 waves ~2.0 ms) err=%s\n", name, nb, ms,
         cudaGetErrorString(cudaGetLastError()));
}
int main() {
  run("plain", k<256, false, false>);
  run("printf", k<256, false, true>);
  run("tmem256", k<256, true, false>);
  run("tmem128", k<128, true, false>);
  run("tmem512", k<512, true, false>);
  return 0;
}
