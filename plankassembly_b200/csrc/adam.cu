// ONE fused Adam launch over a flat fp32 master buffer that also refreshes the TF32-rounded shadow copies the
// tensor-core GEMMs read (ref: trainer_complete.py:127-129 -- torch.optim.Adam(lr), default betas/eps, no weight decay).
//
// torch's fused Adam is ~8 multi_tensor_apply launches per step plus, on this path, ~100 pa_round_tf32 launches to rebuild
// the weight shadows; here the parameters, both moments and the shadows are four parallel flat buffers and the gradients are
// reached through a per-step pointer table (autograd hands out a fresh tensor per parameter), so the whole update is one
// HBM-bound pass: 16 B read + 16 B written per parameter (p, g, m, v -> p, m, v, shadow).
//
// Arithmetic follows torch/optim/adam.py::_single_tensor_adam (no amsgrad, no weight decay, maximize = False):
//   m <- m + (g - m)(1 - b1);  v <- b2 v + (1 - b2) g^2;  p <- p - step_size * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "common.cuh"

namespace {

constexpr int kChunk = 2048;      // elements per CTA pass (256 threads x 2 float4)

__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                        float* __restrict__ shadow, const float* const* __restrict__ grads,
                                                        const int64_t* __restrict__ chunk_off, const int* __restrict__ chunk_param,
                                                        const int64_t* __restrict__ param_off, const int64_t* __restrict__ param_len,
                                                        const float* __restrict__ scalars, int n_chunks, float b1, float b2, float eps) {
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const int pi = chunk_param[c];
    const float* g = grads[pi];
    if (g == nullptr) continue;                         // parameter without a gradient this step: untouched, like torch
    // torch keeps one step count PER PARAMETER (a skipped parameter lags behind): its bias corrections come per parameter
    const float step_size = scalars[2 * pi], inv_bc2_sqrt = scalars[2 * pi + 1];
    const int64_t base = param_off[pi], start = chunk_off[c];          // start: offset of this chunk inside the parameter
    const int64_t end = min(start + kChunk, param_len[pi]);
    for (int64_t i = start + threadIdx.x * 4; i < end; i += 256 * 4) {
      if (i + 4 <= end && ((((uintptr_t)(g + i)) & 15) == 0)) {
        float4 gg = *reinterpret_cast<const float4*>(g + i);
        float4 pp = *reinterpret_cast<float4*>(p + base + i), mm = *reinterpret_cast<float4*>(m + base + i), vv = *reinterpret_cast<float4*>(v + base + i);
        float* ge = &gg.x; float* pe = &pp.x; float* me = &mm.x; float* ve = &vv.x;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          me[e] = me[e] + (ge[e] - me[e]) * (1.f - b1);
          ve[e] = ve[e] * b2 + (1.f - b2) * ge[e] * ge[e];
          pe[e] = pe[e] - step_size * (me[e] / (sqrtf(ve[e]) * inv_bc2_sqrt + eps));
        }
        *reinterpret_cast<float4*>(p + base + i) = pp;
        *reinterpret_cast<float4*>(m + base + i) = mm;
        *reinterpret_cast<float4*>(v + base + i) = vv;
        if (shadow != nullptr) *reinterpret_cast<float4*>(shadow + base + i) = tf32_rn4(pp);
      } else {
        for (int64_t j = i; j < min(i + 4, end); ++j) {
          const float ge = g[j];
          const float me = m[base + j] + (ge - m[base + j]) * (1.f - b1);
          const float ve = v[base + j] * b2 + (1.f - b2) * ge * ge;
          const float pe = p[base + j] - step_size * (me / (sqrtf(ve) * inv_bc2_sqrt + eps));
          m[base + j] = me; v[base + j] = ve; p[base + j] = pe;
          if (shadow != nullptr) shadow[base + j] = tf32_rn(pe);
        }
      }
    }
  }
}

}  // namespace

extern "C" int pa_adam_chunk_elems(void) { return kChunk; }

extern "C" int pa_adam_flat(float* p, float* m, float* v, float* shadow, const float* const* grads, const int64_t* chunk_off,
                            const int* chunk_param, const int64_t* param_off, const int64_t* param_len, const float* scalars,
                            int n_chunks, float beta1, float beta2, float eps, void* stream) {
  PA_CHECK_ARG(p != nullptr && m != nullptr && v != nullptr && grads != nullptr && scalars != nullptr && n_chunks > 0);
  PA_CHECK_ARG((((uintptr_t)p | (uintptr_t)m | (uintptr_t)v | (uintptr_t)shadow) & 15) == 0);
  const int grid = n_chunks < pa_num_sms() * 8 ? n_chunks : pa_num_sms() * 8;
  adam_flat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, m, v, shadow, grads, chunk_off, chunk_param, param_off, param_len, scalars, n_chunks,
                                                            beta1, beta2, eps);
  PA_CHECK_LAUNCH();
  return PA_OK;
}
