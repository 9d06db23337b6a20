// Exact-fp32 projections for the decode step: C[M,N] = X[M,K] W[N,K]^T + bias (optional ReLU), M = number
// of sequences decoded together.  These are weight-streaming, latency-bound GEMMs (0.1 GFLOP each, ~39 per
// step); cuBLAS spends ~16 us on each.  Here a CTA owns a 16-row x 8-column output block: the W slice
// (8 x K) is staged in smem while the x loads are already in flight, a WARP owns a row pair with K
// interleaved over its 32 lanes in 16-byte chunks (coalesced x reads, conflict-free smem reads), so each
// lane has only K/128 dependent-free iterations; the lanes are reduced with shuffles.  fp32 FMA throughout.
#include "common.cuh"

namespace {

constexpr int kCols = 8, kThreads = 256;                   // 8 warps; a warp owns R rows of the CTA's 8 columns
constexpr int kMaxChunks = 12;                              // K <= 1536

// R = rows per warp (CTA = 8*R rows x 8 columns), NCH = upper bound of K/128 (x chunks kept in registers).
// R = 2 made every 8 FFMA wait for one LDS.128 of W (the kernel ran at the shared-memory pipe); R = 4 halves that for
// the K <= 512 projections (all but FFN2), whose 4 x 4 x-chunks still fit in registers.
template <int R, int NCH>
__global__ void __launch_bounds__(kThreads) gemm_skinny_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w, int64_t ldw,
                                                                 const float* __restrict__ bias, float* __restrict__ c, int64_t ldc, int M, int N,
                                                                 int K, int relu) {
  extern __shared__ __align__(16) float ws[];               // [kCols][K]
  const int n0 = blockIdx.x * kCols;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = blockIdx.y * (8 * R) + warp * R;
  const int nchunk = K / 128;                                // float4 chunks per lane
  // issue this lane's x loads first: they overlap the W staging below
  float4 a[R][NCH];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const float* xr = x + (int64_t)min(r0 + r, M - 1) * ldx + lane * 4;
#pragma unroll
    for (int i = 0; i < NCH; ++i)
      if (i < nchunk) a[r][i] = __ldg(reinterpret_cast<const float4*>(xr + i * 128));
  }
  for (int i = threadIdx.x * 4; i < kCols * K; i += kThreads * 4) {
    const int col = i / K, k = i - col * K;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 + col < N) v = __ldg(reinterpret_cast<const float4*>(w + (int64_t)(n0 + col) * ldw + k));
    *reinterpret_cast<float4*>(ws + i) = v;
  }
  __syncthreads();
  float acc[R][kCols];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int j = 0; j < kCols; ++j) acc[r][j] = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    if (i < nchunk) {
#pragma unroll
      for (int j = 0; j < kCols; ++j) {
        const float4 wv = *reinterpret_cast<const float4*>(ws + j * K + i * 128 + lane * 4);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float t = acc[r][j];
          t = fmaf(a[r][i].x, wv.x, t); t = fmaf(a[r][i].y, wv.y, t); t = fmaf(a[r][i].z, wv.z, t); t = fmaf(a[r][i].w, wv.w, t);
          acc[r][j] = t;
        }
      }
    }
  }
  // lane (r * 8 + j) ends up with row r0 + r, column n0 + j
  float mine = 0.f;
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int j = 0; j < kCols; ++j) {
      const float sv = warp_sum(acc[r][j]);
      if (lane == r * kCols + j) mine = sv;
    }
  if (lane < R * kCols) {
    const int col = n0 + (lane & 7), row = r0 + (lane >> 3);
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
  static SmemAttrCache attr;
  if (int rc_attr = pa_set_max_smem(gemm_skinny_kernel<2, kMaxChunks>, kCols * 128 * kMaxChunks * 4, attr)) return rc_attr;
  if (K <= 512 && M >= 32) {
    dim3 grid((N + kCols - 1) / kCols, (M + 31) / 32);
    gemm_skinny_kernel<4, 4><<<grid, kThreads, smem, (cudaStream_t)stream>>>(x, ldx, w, ldw, bias, c, ldc, M, N, K, relu);
  } else {
    dim3 grid((N + kCols - 1) / kCols, (M + 15) / 16);
    gemm_skinny_kernel<2, kMaxChunks><<<grid, kThreads, smem, (cudaStream_t)stream>>>(x, ldx, w, ldw, bias, c, ldc, M, N, K, relu);
  }
  PA_CHECK_LAUNCH();
  return PA_OK;
}
