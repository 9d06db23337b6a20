// Issue/execution rate of tcgen05.mma kind::tf32 as a function of N (M = 128, K = 8 per instruction), SS and TS forms:
// one CTA per SM, one warp issues ROUNDS x 16 MMAs into one accumulator and waits for the commit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../plankassembly_b200/csrc mma_rate.cu -o mma_rate && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"

template <int N, bool TS, bool BMN = false>
__global__ void __launch_bounds__(128, 1) k(long long* out, int rounds) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc<512>(&slot);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    constexpr uint32_t idesc = tc::make_idesc_tf32(128, N, 0, BMN ? 1 : 0);
    const uint64_t da = tc::make_smem_desc(tc::smem_u32(smem), 16, 1024);
    // MN-major tf32 B (the V / dO / Q operands of the attention kernels): 32-wide MN blocks of 128 k-rows x 128 B, 4-row atoms
    const uint64_t db = BMN ? tc::make_smem_desc(tc::smem_u32(smem + 16384), 128 * 128, 512, tc::kLayoutSw128Base32)
                            : tc::make_smem_desc(tc::smem_u32(smem + 16384), 16, 1024);
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      if (tc::elect_one()) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint64_t dbi = tc::desc_advance(db, BMN ? (i & 15) * 1024 : (i & 3) * 32);
          if (TS) tc::mma_tf32_ts(tm, tm + 256 + (i & 3) * 8, dbi, idesc, 1u);
          else tc::mma_tf32_ss(tm, tc::desc_advance(da, (i & 3) * 32), dbi, idesc, 1u);
        }
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (tc::elect_one()) tc::tc_commit(&bar);
    __syncwarp();
    tc::mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tm);
}

template <int N, bool TS, bool BMN = false>
void run(long long* d, int grid) {
  const int rounds = 64;
  cudaFuncSetAttribute(k<N, TS, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<N, TS, BMN><<<grid, 128, 100 * 1024>>>(d, rounds);
  k<N, TS, BMN><<<grid, 128, 100 * 1024>>>(d, rounds);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const double n = rounds * 16.0;
  printf("M=128 N=%3d %s%s grid %3d: issue %6.1f cyc/MMA, issue+execute %6.1f cyc/MMA (formula floor %d)  %s\n", N, TS ? "TS" : "SS", BMN ? " B MN-major" : "", grid, h[0] / n, h[1] / n,
         128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  for (int grid : {1, 148}) {
    run<32, false>(d, grid); run<64, false>(d, grid); run<128, false>(d, grid); run<256, false>(d, grid);
    run<64, true>(d, grid); run<128, true>(d, grid);
    run<64, true, true>(d, grid); run<64, false, true>(d, grid); run<128, false, true>(d, grid);
  }
  return 0;
}
