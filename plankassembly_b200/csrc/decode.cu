// K12 (first cut): single-query attention over the persistent K/V caches, one launch per
// (self|cross) attention per layer per step.  HBM-bound: each (sequence, head) block streams its
// K and V slices once (2 * len * dh * 4 bytes) -- the reference instead recomputes every earlier
// position and re-projects the encoder memory on every step (ref models.py:284-307).
#include "common.cuh"

namespace {

constexpr int kThreads = 128;

template <int DH>
__global__ void __launch_bounds__(kThreads) decode_attn_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k_new,
                                                                 const float* __restrict__ v_new, int64_t ld_new, float* k_cache,
                                                                 float* v_cache, int64_t cache_len, int64_t ldc, int t_host, int len_host, const int* __restrict__ t_dev,
                                                                 const uint8_t* __restrict__ kpm, const int* __restrict__ kv_len, int H, float scale,
                                                                 float* __restrict__ o) {
  extern __shared__ float s_p[];                // [len] scores -> probabilities
  __shared__ float red[kThreads / 32];
  __shared__ float s_o[kThreads / 32][DH];
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = H * DH;
  // self-attention under a CUDA graph: the step index lives in device memory (len = t + 1)
  const int t = (t_dev != nullptr && k_new != nullptr) ? *t_dev : t_host;
  // cross-attention over a padded memory: keys beyond kv_len[b] (1 + last non-PAD key) are all PAD (-inf): not even read
  const int len = (t_dev != nullptr && k_new != nullptr) ? t + 1 : (kv_len != nullptr ? min(len_host, max(kv_len[b], 1)) : len_host);
  float* kc = k_cache + (int64_t)b * cache_len * ldc + h * DH;
  float* vc = v_cache + (int64_t)b * cache_len * ldc + h * DH;
  if (k_new != nullptr) {                       // append this step's key/value (self-attention)
    if (tid < DH) kc[(int64_t)t * ldc + tid] = k_new[(int64_t)b * ld_new + h * DH + tid];
    else if (tid < 2 * DH) vc[(int64_t)t * ldc + tid - DH] = v_new[(int64_t)b * ld_new + h * DH + tid - DH];
    __syncthreads();
  }
  // scores: 8 lanes per key, DH/8 floats per lane
  constexpr int F = DH / 8;
  const int sub = lane & 7, grp = lane >> 3;
  float qr[F];
#pragma unroll
  for (int c = 0; c < F; ++c) qr[c] = q[(int64_t)b * ldq + h * DH + sub * F + c];
  float mx = -INFINITY;
  // 4 keys per warp pass, 4 passes unrolled: 8 independent 16-byte loads in flight per lane
  for (int j0 = warp * 4; j0 < len; j0 += (kThreads / 32) * 4 * 4) {
    float acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * (kThreads / 32) * 4 + grp;
      acc[u] = 0.f;
      if (j < len) {
        const float* kr = kc + (int64_t)j * ldc + sub * F;
#pragma unroll
        for (int c = 0; c < F; c += 4) {
          float4 kk = *reinterpret_cast<const float4*>(kr + c);
          acc[u] += qr[c] * kk.x + qr[c + 1] * kk.y + qr[c + 2] * kk.z + qr[c + 3] * kk.w;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * (kThreads / 32) * 4 + grp;
      float a = acc[u];
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      if (j < len && sub == 0) {
        float s = a * scale;
        if (kpm != nullptr && kpm[(int64_t)b * len_host + j]) s = -INFINITY;
        s_p[j] = s;
        mx = fmaxf(mx, s);
      }
    }
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  const float m_safe = mx == -INFINITY ? 0.f : mx;
  float sum = 0.f;
  for (int j = tid; j < len; j += kThreads) {
    float p = expf(s_p[j] - m_safe);
    s_p[j] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = (red[0] + red[1]) + (red[2] + red[3]);
  // O = P V : each warp takes keys j = warp, warp+4, ...; lanes span the head dim
  constexpr int C = DH / 32;   // floats per lane (1 or 2)
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
  for (int j0 = warp; j0 < len; j0 += (kThreads / 32) * 8) {      // 8 key rows in flight per warp
    float vv[8][C], pp[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + u * (kThreads / 32);
      pp[u] = j < len ? s_p[j] : 0.f;
      const float* vr = vc + (int64_t)min(j, len - 1) * ldc + lane * C;
#pragma unroll
      for (int c = 0; c < C; ++c) vv[u][c] = vr[c];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] += pp[u] * vv[u][c];
  }
#pragma unroll
  for (int c = 0; c < C; ++c) s_o[warp][lane * C + c] = acc[c];
  __syncthreads();
  if (tid < DH) {
    float r = (s_o[0][tid] + s_o[1][tid]) + (s_o[2][tid] + s_o[3][tid]);
    o[(int64_t)b * d + h * DH + tid] = sum > 0.f ? r / sum : 0.f;
  }
}

__global__ void decode_advance_kernel(int* t_dev) { *t_dev += 1; }

}  // namespace

extern "C" int pa_decode_advance(int* t_dev, void* stream) {
  decode_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(t_dev);
  PA_CHECK_LAUNCH();
  return PA_OK;
}

extern "C" int pa_decode_attn(const float* q, int64_t ldq, const float* k_new, const float* v_new, int64_t ld_new,
                              float* k_cache, float* v_cache, int64_t cache_len, int64_t ld_cache, int t, int len,
                              const int* t_dev, const uint8_t* kpm, const int* kv_len, int B, int H, int dh, float scale, float* o, void* stream) {
  PA_CHECK_ARG(ld_cache % 4 == 0 && B > 0 && H > 0 && len > 0 && len <= cache_len && (k_new == nullptr || (t >= 0 && t < cache_len)));
  dim3 grid(H, B);
  size_t smem = (size_t)(t_dev != nullptr ? cache_len : len) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dh) {
    case 32: decode_attn_kernel<32><<<grid, kThreads, smem, st>>>(q, ldq, k_new, v_new, ld_new, k_cache, v_cache, cache_len, ld_cache, t, len, t_dev, kpm, kv_len, H, scale, o); break;
    case 64: decode_attn_kernel<64><<<grid, kThreads, smem, st>>>(q, ldq, k_new, v_new, ld_new, k_cache, v_cache, cache_len, ld_cache, t, len, t_dev, kpm, kv_len, H, scale, o); break;
    default: pa_set_error("pa_decode_attn: head dim %d unsupported (32, 64)", dh); return PA_ERR_UNSUPPORTED;
  }
  PA_CHECK_LAUNCH();
  return PA_OK;
}
