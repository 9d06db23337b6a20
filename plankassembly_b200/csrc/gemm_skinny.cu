// Exact-fp32 projections for the decode step: C[M,N] = X[M,K] W[N,K]^T + bias (optional ReLU), M = number
// of sequences decoded together (<= 64 per launch row-block).  These are weight-streaming, latency-bound
// GEMMs (2*M*N*K = 0.1 GFLOP each, 33 per step); cuBLAS spends ~16 us on each, which made them 60 % of the
// decode step.  Here one CTA owns 8 output columns: its W slice (8 x K) is staged in smem once, a thread
// owns 2 rows x 8 columns over an interleaved 1/8 of K, and the 8 K-slices of a row pair sit in adjacent
// lanes so the final reduction is three shuffles.  FMA order differs from cuBLAS only in association (fp32 throughout).
#include "common.cuh"

namespace {

constexpr int kCols = 8, kSlices = 8, kThreads = 256;      // 32 row pairs x 8 K-slices

__global__ void __launch_bounds__(kThreads) gemm_skinny_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w, int64_t ldw,
                                                                 const float* __restrict__ bias, float* __restrict__ c, int64_t ldc, int M, int N,
                                                                 int K, int relu) {
  extern __shared__ __align__(16) float ws[];               // [kCols][K]
  const int n0 = blockIdx.x * kCols;
  const int m_base = blockIdx.y * 64;
  for (int i = threadIdx.x * 4; i < kCols * K; i += kThreads * 4) {
    const int col = i / K, k = i - col * K;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 + col < N) v = __ldg(reinterpret_cast<const float4*>(w + (int64_t)(n0 + col) * ldw + k));
    *reinterpret_cast<float4*>(ws + i) = v;
  }
  __syncthreads();
  const int slice = threadIdx.x & (kSlices - 1), pair = threadIdx.x >> 3;
  const int r0 = m_base + pair * 2, r1 = r0 + 1;
  // K is interleaved over the 8 slices in 16-byte chunks (chunk i*8 + slice): the 8 lanes of a row pair read
  // consecutive chunks of x (one 128-byte line) and of the smem W rows (no bank conflicts).
  const float* x0 = x + (int64_t)min(r0, M - 1) * ldx + slice * 4;
  const float* x1 = x + (int64_t)min(r1, M - 1) * ldx + slice * 4;
  float acc0[kCols], acc1[kCols];
#pragma unroll
  for (int j = 0; j < kCols; ++j) acc0[j] = acc1[j] = 0.f;
#pragma unroll 2
  for (int k = 0; k < K; k += 32) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x0 + k));
    const float4 b = __ldg(reinterpret_cast<const float4*>(x1 + k));
#pragma unroll
    for (int j = 0; j < kCols; ++j) {
      const float4 wv = *reinterpret_cast<const float4*>(ws + j * K + k + slice * 4);
      acc0[j] = fmaf(a.x, wv.x, acc0[j]); acc0[j] = fmaf(a.y, wv.y, acc0[j]); acc0[j] = fmaf(a.z, wv.z, acc0[j]); acc0[j] = fmaf(a.w, wv.w, acc0[j]);
      acc1[j] = fmaf(b.x, wv.x, acc1[j]); acc1[j] = fmaf(b.y, wv.y, acc1[j]); acc1[j] = fmaf(b.z, wv.z, acc1[j]); acc1[j] = fmaf(b.w, wv.w, acc1[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < kCols; ++j) {
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) {
      acc0[j] += __shfl_xor_sync(0xffffffffu, acc0[j], off);
      acc1[j] += __shfl_xor_sync(0xffffffffu, acc1[j], off);
    }
  }
  // lane `slice` of each 8-lane group writes column `slice`
  const int col = n0 + slice;
  if (col < N) {
    float v0 = 0.f, v1 = 0.f;
#pragma unroll
    for (int j = 0; j < kCols; ++j)
      if (j == slice) { v0 = acc0[j]; v1 = acc1[j]; }
    const float bb = bias != nullptr ? __ldg(bias + col) : 0.f;
    v0 += bb; v1 += bb;
    if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
    if (r0 < M) c[(int64_t)r0 * ldc + col] = v0;
    if (r1 < M) c[(int64_t)r1 * ldc + col] = v1;
  }
}

}  // namespace

extern "C" int pa_gemm_skinny_f32(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, float* c,
                                  int64_t ldc, int M, int N, int K, int relu, void* stream) {
  PA_CHECK_ARG(M > 0 && N > 0 && K > 0 && K % 32 == 0 && ldx % 4 == 0 && ldw % 4 == 0);
  PA_CHECK_ARG((((uintptr_t)x | (uintptr_t)w) & 15) == 0);
  const size_t smem = (size_t)kCols * K * sizeof(float);
  PA_CHECK_ARG(smem <= 48 * 1024);
  dim3 grid((N + kCols - 1) / kCols, (M + 63) / 64);
  gemm_skinny_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(x, ldx, w, ldw, bias, c, ldc, M, N, K, relu);
  PA_CHECK_LAUNCH();
  return PA_OK;
}
