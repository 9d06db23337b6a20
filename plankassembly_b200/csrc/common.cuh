// Shared device/host helpers for the plank_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/plank_b200.h"

void pa_set_error(const char* fmt, ...);

#define PA_CHECK_ARG(cond)                                                        \
  do {                                                                            \
    if (!(cond)) {                                                                \
      pa_set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, #cond);         \
      return PA_ERR_ARG;                                                          \
    }                                                                             \
  } while (0)

#define PA_CHECK_LAUNCH()                                                         \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      pa_set_error("%s:%d: CUDA: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return PA_ERR_CUDA;                                                         \
    }                                                                             \
  } while (0)

#define PA_CUDA(call)                                                             \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      pa_set_error("%s:%d: CUDA: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return PA_ERR_CUDA;                                                         \
    }                                                                             \
  } while (0)

constexpr int kNumSMs = 148;  // B200 (compile-time default; launch code sizes its grids with pa_num_sms())
int pa_num_sms();             // SM count of the CURRENT device (cached per device; falls back to kNumSMs)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: a process that drives several GPUs (tests on
// cuda:1, DataParallel) must set it on each of them.  One cache per kernel (a function-local static at the call site).
#ifdef __cplusplus
#include <mutex>
constexpr int kMaxDevices = 32;
struct SmemAttrCache {
  std::mutex mu;
  int bytes[kMaxDevices] = {};
};
template <typename Kern>
inline int pa_set_max_smem(Kern kern, int bytes, SmemAttrCache& cache) {
  int dev = 0;
  PA_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(cache.mu);
  const bool tracked = dev >= 0 && dev < kMaxDevices;
  if (!tracked || bytes > cache.bytes[dev]) {
    PA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (tracked) cache.bytes[dev] = bytes;
  }
  return PA_OK;
}
#endif

// Round-to-nearest conversion to TF32 (10-bit mantissa, low 13 bits zero).  tcgen05 kind::tf32 reads raw
// fp32 bits and TRUNCATES them, which biases every product by ~2^-11; operands that were rounded here
// are already exactly representable, so the MMA sees unbiased values.
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
// Same rounding (nearest, ties away from zero) for values known to be finite: add half an ulp of the 10-bit mantissa
// and clear the low 13 bits -- two integer instructions instead of cvt.rna's compare + add + mask.
__device__ __forceinline__ uint32_t tf32_rn_finite_bits(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ float4 tf32_rn4(float4 v) { return make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w)); }

// 2^x on the MUFU in one instruction (max rel. error 2^-22; -inf -> 0), for softmax probabilities
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 256-bit global store (sm_100+): p must be 32-byte aligned.  The tensor-core kernels read accumulators out of TMEM as
// (lane = row, registers = consecutive columns), so each lane writes its own 128-byte row segment; the LSU works per
// touched line and instruction, and the wide form halves the number of store instructions.
__device__ __forceinline__ void st_global_v8(float* p, float4 a, float4 b) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Reduce-scatter of 32 per-lane values over the warp in 31 shuffles: on return lane l holds the sum over
// all lanes of v[l] (column sums of a 32x32 register tile whose rows are the lanes).
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float keep = up ? v[i + off] : v[i];
      const float send = up ? v[i] : v[i + off];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG: dropout masks are a pure function of (seed, offset, element
// index), so backward kernels regenerate the forward mask instead of storing it.
// ---------------------------------------------------------------------------------------------
template <int ROUNDS = 10>
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// Attention-probability dropout, bit-sliced: the keep bits of (query row, 32 consecutive keys) are ONE word, built from 16
// random words (4 calls of Philox4x32-7, the Crush-resistant round count of the Random123 paper; the 10-round default only adds
// margin and this kernel is pure ALU) by folding them over the binary digits of the keep probability q = keep16 / 2^16, least
// significant digit first:  W <- digit ? (W | r) : (W & r).  Every bit of W is then 1 with probability exactly q,
// independently of the others -- the same distribution as comparing 16 random bits per score with a threshold, for 16 logic
// operations per 32 scores instead of ~100.  `row_pitch_words` = ceil(Lk / 32).  Shared by dropmask.cu (the bit planes the
// tensor-core kernels read) and attn_simt.cu (which recomputes single words).
__device__ __forceinline__ uint32_t drop_keep_word(uint64_t seed, uint64_t offset, int64_t row_global, int row_pitch_words, int kw, uint32_t keep16) {
  if (keep16 >= 65536u) return 0xffffffffu;
  uint32_t W = 0u;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const uint4 r = philox4x32<7>(seed, (uint64_t)((row_global * row_pitch_words + kw) * 4 + g), offset);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) W = ((keep16 >> (4 * g + e)) & 1u) ? (W | w[e]) : (W & w[e]);
  }
  return W;
}

// Dropout probabilities are quantised to 16 bits: residual-row dropout (LayerNorm kernels) draws 16 random bits per element
// and keeps iff half-word >= p * 2^16; the attention bit planes use keep16 = 2^16 - threshold (drop_keep_word above).  Either
// way the drop probability is p rounded down to a multiple of 2^-16.
__host__ __device__ __forceinline__ uint32_t drop_threshold16(float p) {
  double t = (double)p * 65536.0;
  return t >= 65535.0 ? 0xFFFFu : (uint32_t)t;
}

// keep-decision for one element: u32 random >= threshold  (threshold = p * 2^32)
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
  double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}
