// tcgen05.ld / tcgen05.st bandwidth of one SM (32x32b.x32: 4 KB per warp instruction), alone and under a stream of MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../plankassembly_b200/csrc tmem_rate.cu -o tmem_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"

template <int MODE>   // 0: loads, 1: stores, 2: load + store pairs
__global__ void __launch_bounds__(320, 1) k(long long* out, int rounds, int nw, int with_mma) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ long long tmax[10];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 320) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (warp == 1) tc::tmem_alloc<512>(&slot);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = slot;
  long long t0 = clock64();
  if (warp == 1 && with_mma) {
    constexpr uint32_t idesc = tc::make_idesc_tf32(128, 64, 0, 0);
    const uint64_t db = tc::make_smem_desc(tc::smem_u32(smem + 16384), 16, 1024);
    for (int r = 0; r < rounds; ++r) {        // 8 TS MMAs (128x64x8) per round: 256 cycles of tensor work
      if (tc::elect_one()) {
#pragma unroll
        for (int i = 0; i < 8; ++i) tc::mma_tf32_ts(tm + 384, tm + 448 + (i & 3) * 8, tc::desc_advance(db, (i & 3) * 32), idesc, 1u);
      }
      __syncwarp();
    }
    if (tc::elect_one()) tc::tc_commit(&bar);
    __syncwarp();
    tc::mbar_wait(&bar, 0);
    if (lane == 0) tmax[1] = clock64() - t0;
  } else if (warp >= 2 && warp < 2 + nw) {
    const uint32_t addr = tm + ((uint32_t)((warp & 3) * 32) << 16) + ((warp - 2) >> 2) * 64;
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = i;
    for (int it = 0; it < rounds; ++it) {
      if (MODE == 0 || MODE == 2) { tc::tmem_ld_32x32(addr + (it & 1) * 32, r); tc::tmem_ld_wait(); }
      if (MODE == 1 || MODE == 2) { tc::tmem_st_32x32(addr + (it & 1) * 32, r); tc::tmem_st_wait(); }
    }
    if (r[3] == 0x12345) out[7] = 1;
    if (lane == 0) tmax[warp] = clock64() - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long m = 0;
    for (int w = 2; w < 2 + nw; ++w) m = tmax[w] > m ? tmax[w] : m;
    out[0] = m; out[1] = with_mma ? tmax[1] : 0;
  }
  if (warp == 1) tc::tmem_dealloc<512>(tm);
}

template <int MODE>
void run(long long* d, int nw, int with_mma) {
  const int rounds = 256;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<MODE><<<148, 320, 100 * 1024>>>(d, rounds, nw, with_mma);
  k<MODE><<<148, 320, 100 * 1024>>>(d, rounds, nw, with_mma);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const double bytes = (double)rounds * nw * 4096 * (MODE == 2 ? 2 : 1);
  printf("%-12s %d warps%s: %7.1f cyc per warp-iteration, %6.1f B/clk per SM", MODE == 0 ? "ld" : MODE == 1 ? "st" : "ld+st", nw,
         with_mma ? " + MMA stream" : "", (double)h[0] / rounds, bytes / h[0]);
  if (with_mma) printf(" | MMAs: %6.1f cyc per 128x64x8 TS MMA (32 alone)", (double)h[1] / (rounds * 8.0));
  printf("  %s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  for (int mma = 0; mma < 2; ++mma)
    for (int nw : {1, 4, 8}) { run<0>(d, nw, mma); run<1>(d, nw, mma); run<2>(d, nw, mma); }
  return 0;
}
