// Error-compensated ("3xTF32") operands for the exact-mode projections of inference.
//
// tcgen05 kind::tf32 keeps 11 significant bits of each operand.  Writing x = hi + lo with hi = rn_tf32(x) and
// lo = rn_tf32(x - hi) (x - hi is exact in fp32), the product x*w = hi_x*hi_w + lo_x*hi_w + hi_x*lo_w + O(2^-22 |x||w|),
// accumulated in FP32 inside the tensor core: fp32-class accuracy from three TF32 MMAs.  Instead of a new GEMM kernel the
// three terms are laid out ALONG K, so that ONE ordinary pa_gemm_tf32 call over K' = 3K computes them:
//   activations (mode 0): out[r] = [ hi(x[r]) | lo(x[r]) | hi(x[r]) ]
//   weights     (mode 1): out[n] = [ hi(w[n]) | hi(w[n]) | lo(w[n]) ]
// Replaces the cuBLAS SGEMMs of round 1's eval path (prefill encoder, cross-K/V projection) and feeds the
// large-batch decode step (ref models.py:279, 293; torch nn/functional.py _in_projection_packed).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ out, int64_t rows, int K, int mode) {
  const int k4 = K >> 2;
  const int64_t n4 = rows * k4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / k4;
    const int c = (int)(i - r * k4) << 2;
    const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
    const float4 hi = tf32_rn4(v);
    const float4 lo = tf32_rn4(make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w));
    float* o = out + r * (3 * (int64_t)K) + c;
    *reinterpret_cast<float4*>(o) = hi;
    *reinterpret_cast<float4*>(o + K) = mode == 0 ? lo : hi;
    *reinterpret_cast<float4*>(o + 2 * K) = mode == 0 ? hi : lo;
  }
}

}  // namespace

extern "C" int pa_split3_tf32(const float* x, int64_t ldx, float* out, int64_t rows, int K, int weights, void* stream) {
  PA_CHECK_ARG(x != nullptr && out != nullptr && rows > 0 && K > 0 && K % 4 == 0 && ldx % 4 == 0);
  PA_CHECK_ARG((((uintptr_t)x | (uintptr_t)out) & 15) == 0);
  const int64_t n4 = rows * (K / 4);
  const int grid = (int)((n4 + 255) / 256 < pa_num_sms() * 8 ? (n4 + 255) / 256 : pa_num_sms() * 8);
  split3_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, out, rows, K, weights ? 1 : 0);
  PA_CHECK_LAUNCH();
  return PA_OK;
}
