// Exact-fp32 projections for the decode step: C[M,N] = X[M,K] W[N,K]^T + bias (optional ReLU), M = number
// of sequences decoded together.  These are weight-streaming, latency-bound GEMMs (0.1 GFLOP each, ~39 per
// step); cuBLAS spends ~16 us on each.  Here a CTA owns a 16-row x 8-column output block: the W slice
// (8 x K) is staged in smem while the x loads are already in flight, a WARP owns a row pair with K
// interleaved over its 32 lanes in 16-byte chunks (coalesced x reads, conflict-free smem reads), so each
// lane has only K/128 dependent-free iterations; the lanes are reduced with shuffles.  fp32 FMA throughout.
#include "common.cuh"

namespace {

constexpr int kCols = 8, kRows = 16, kThreads = 256;       // 8 warps = 8 row pairs
constexpr int kMaxChunks = 12;                              // K <= 1536

__global__ void __launch_bounds__(kThreads) gemm_skinny_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w, int64_t ldw,
                                                                 const float* __restrict__ bias, float* __restrict__ c, int64_t ldc, int M, int N,
                                                                 int K, int relu) {
  extern __shared__ __align__(16) float ws[];               // [kCols][K]
  const int n0 = blockIdx.x * kCols;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = blockIdx.y * kRows + warp * 2, r1 = r0 + 1;
  const int nchunk = K / 128;                                // float4 chunks per lane
  // issue this lane's x loads first: they overlap the W staging below
  const float* x0 = x + (int64_t)min(r0, M - 1) * ldx + lane * 4;
  const float* x1 = x + (int64_t)min(r1, M - 1) * ldx + lane * 4;
  float4 a[kMaxChunks], b[kMaxChunks];
#pragma unroll
  for (int i = 0; i < kMaxChunks; ++i)
    if (i < nchunk) { a[i] = __ldg(reinterpret_cast<const float4*>(x0 + i * 128)); b[i] = __ldg(reinterpret_cast<const float4*>(x1 + i * 128)); }
  for (int i = threadIdx.x * 4; i < kCols * K; i += kThreads * 4) {
    const int col = i / K, k = i - col * K;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 + col < N) v = __ldg(reinterpret_cast<const float4*>(w + (int64_t)(n0 + col) * ldw + k));
    *reinterpret_cast<float4*>(ws + i) = v;
  }
  __syncthreads();
  float acc0[kCols], acc1[kCols];
#pragma unroll
  for (int j = 0; j < kCols; ++j) acc0[j] = acc1[j] = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxChunks; ++i) {
    if (i < nchunk) {
#pragma unroll
      for (int j = 0; j < kCols; ++j) {
        const float4 wv = *reinterpret_cast<const float4*>(ws + j * K + i * 128 + lane * 4);
        acc0[j] = fmaf(a[i].x, wv.x, acc0[j]); acc0[j] = fmaf(a[i].y, wv.y, acc0[j]); acc0[j] = fmaf(a[i].z, wv.z, acc0[j]); acc0[j] = fmaf(a[i].w, wv.w, acc0[j]);
        acc1[j] = fmaf(b[i].x, wv.x, acc1[j]); acc1[j] = fmaf(b[i].y, wv.y, acc1[j]); acc1[j] = fmaf(b[i].z, wv.z, acc1[j]); acc1[j] = fmaf(b[i].w, wv.w, acc1[j]);
      }
    }
  }
  // lanes 0..7 end up with row r0's 8 columns, lanes 8..15 with row r1's
  float mine = 0.f;
#pragma unroll
  for (int j = 0; j < kCols; ++j) {
    const float s0 = warp_sum(acc0[j]), s1 = warp_sum(acc1[j]);
    if (lane == j) mine = s0;
    if (lane == j + 8) mine = s1;
  }
  if (lane < 16) {
    const int col = n0 + (lane & 7), row = lane < 8 ? r0 : r1;
    if (col < N && row < M) {
      float v = mine + (bias != nullptr ? __ldg(bias + col) : 0.f);
      if (relu) v = fmaxf(v, 0.f);
      c[(int64_t)row * ldc + col] = v;
    }
  }
}

}  // namespace

extern "C" int pa_gemm_skinny_f32(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, float* c,
                                  int64_t ldc, int M, int N, int K, int relu, void* stream) {
  PA_CHECK_ARG(M > 0 && N > 0 && K > 0 && K % 128 == 0 && K <= 128 * kMaxChunks && ldx % 4 == 0 && ldw % 4 == 0);
  PA_CHECK_ARG((((uintptr_t)x | (uintptr_t)w) & 15) == 0);
  const size_t smem = (size_t)kCols * K * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    PA_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCols * 128 * kMaxChunks * 4));
    attr_done = true;
  }
  dim3 grid((N + kCols - 1) / kCols, (M + kRows - 1) / kRows);
  gemm_skinny_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(x, ldx, w, ldw, bias, c, ldc, M, N, K, relu);
  PA_CHECK_LAUNCH();
  return PA_OK;
}
