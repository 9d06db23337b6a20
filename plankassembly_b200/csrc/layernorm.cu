// K6/K7 epilogues: y = LayerNorm_eps(x + dropout_p(a)), forward and backward, plus the FFN
// relu+dropout elementwise kernel.  HBM-bound: one warp owns a row, 16-byte accesses, the row
// lives in registers between the two reductions (single pass over HBM).
//   Reference: post-norm blocks of torch nn/modules/transformer.py (:952-956, :1144-1153) with the
//   eps = 1.0 quirk introduced at ref models.py:60-61,66-67; final norms (eps 1e-5) at :62,:68.
// Algorithmic bytes per row: fwd reads 2*d*4 (x, a) and writes 2*d*4 (y, s) + 8; bwd reads 2*d*4
// (dy, s) and writes up to 2*d*4 (dx, da).
#include "common.cuh"
#include <stdlib.h>

constexpr int kLnWarps = 4;
// Resident blocks per SM the backward kernel is compiled for: unconstrained ptxas takes 148 registers (3 blocks = 12 warps per
// SM, 3.9 TB/s on 32768 x 512); 4 blocks = 125 registers, no spills, 4.5 TB/s; 5 blocks = 96 registers with spills, 3.7 TB/s.
#ifndef PA_LN_BWD_MIN_BLOCKS
#define PA_LN_BWD_MIN_BLOCKS 4
#endif
// The row-per-warp kernels issue a row's loads, reduce, store, and only then touch the next row: with few resident warps
// (the backward needs 148 registers: 12 warps per SM) too few bytes are in flight to cover the DRAM latency (ncu: 3.0 TB/s).
// Each warp therefore pulls the rows it will process NEXT into L2 while it works on the current one (no registers, no smem).
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <int NV>
__device__ __forceinline__ void prefetch_rows(int lane, int64_t row, const float4* t0, const float4* t1, const float4* t2) {
  constexpr int kLines = NV * 4;                       // 128-byte lines per row (d * 4 / 128)
  const float4* ts[3] = {t0, t1, t2};
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    if (ts[t] == nullptr) continue;
    for (int l = lane; l < kLines; l += 32) prefetch_l2(reinterpret_cast<const char*>(ts[t] + row * (NV * 32)) + l * 128);
  }
}
static inline int ln_debug() {
  static const int v = [] { const char* e = getenv("PLANK_B200_LN_DEBUG"); return e ? atoi(e) : 0; }();
  return v;
}
// Measured on the encoder shape (32768 x 512, scripts/bench_ln.py): backward 89 -> 69 us with one row of look-ahead (3.0 -> 3.9 TB/s);
// the forward (56 registers, 32 warps per SM, 5.4 TB/s = torch's copy at this size) gets slower with it (49 -> 56 us): off there.
static inline int ln_prefetch_depth() {
  static const int v = [] { const char* e = getenv("PLANK_B200_LN_PREFETCH"); return e ? atoi(e) : 1; }();
  return v;
}
static inline int ln_prefetch_fwd() {
  static const int v = [] { const char* e = getenv("PLANK_B200_LN_PREFETCH_FWD"); return e ? atoi(e) : 0; }();
  return v;
}
// Blocks of 4 warps per SM for the row-per-warp LayerNorm kernels.  The ncu capture of round 2 showed add_ln_fwd at 3.4 TB/s
// with 24 % of the warp slots occupied (4 blocks per SM: too few rows in flight to cover the HBM latency, the loads of a
// warp's next row are only issued after its current row is stored), hence 8 (measured per step: 4 -> 22.71 ms, 8 -> 21.90, 16 -> 21.95, 32 -> 22.06); PLANK_B200_LN_BLOCKS_PER_SM overrides.
#include <stdlib.h>
static inline int ln_max_blocks() {
  static const int per_sm = [] { const char* e = getenv("PLANK_B200_LN_BLOCKS_PER_SM"); int v = e ? atoi(e) : 8; return v > 0 ? v : 8; }();
  return pa_num_sms() * per_sm;
}

// Sub-block output dropout of the residual rows: 16 random bits per element, one Philox4x32-7 call (the Crush-resistant round
// count, as in the attention planes) per PAIR of float4
// chunks of a lane (chunk i uses the low halves of the four words when i is even, the high halves when it is odd).
// Forward and backward derive the same keep-mask from (seed, offset, row, lane, i).
__device__ __forceinline__ uint4 ln_drop_words(uint64_t seed, uint64_t offset, int64_t row, int d4, int lane, int i) {
  return philox4x32<7>(seed, (uint64_t)(row * (d4 / 2 + 32) + lane + (i >> 1) * 32), offset);
}
__device__ __forceinline__ uint32_t ln_half(uint32_t w, int i) { return (i & 1) ? (w >> 16) : (w & 0xffffu); }

template <int NV>  // float4 per lane; d = NV*128
__global__ void __launch_bounds__(kLnWarps * 32) add_ln_fwd_kernel(const float4* __restrict__ x, const float4* __restrict__ a, const float4* __restrict__ abias,
                                                                     const float4* __restrict__ gamma, const float4* __restrict__ beta,
                                                                     float eps, float p_drop, uint64_t seed, uint64_t offset, int64_t rows,
                                                                     float4* __restrict__ y, float4* __restrict__ y_r, float4* __restrict__ s_out, float2* __restrict__ stats,
                                                                     int prefetch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d4 = NV * 32;
  const float inv_d = 1.f / (float)(NV * 128);
  const uint32_t thr = drop_threshold16(p_drop);
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  for (int64_t row = (int64_t)blockIdx.x * kLnWarps + warp; row < rows; row += (int64_t)gridDim.x * kLnWarps) {
    const int64_t ahead = row + (int64_t)prefetch * gridDim.x * kLnWarps;
    if (prefetch > 0 && ahead < rows) prefetch_rows<NV>(lane, ahead, x, a, nullptr);
    float4 v[NV];
    float sum = 0.f;
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c = lane + i * 32;
      v[i] = __ldg(x + row * d4 + c);
      if (a != nullptr) {
        float4 av = __ldg(a + row * d4 + c);
        if (abias != nullptr) {           // bias of the linear that produced `a` (its GEMM ran bias-free)
          float4 bb = __ldg(abias + c);
          av.x += bb.x; av.y += bb.y; av.z += bb.z; av.w += bb.w;
        }
        if (p_drop > 0.f) {
          if ((i & 1) == 0) r = ln_drop_words(seed, offset, row, d4, lane, i);
          av.x = ln_half(r.x, i) >= thr ? av.x * keep_scale : 0.f;
          av.y = ln_half(r.y, i) >= thr ? av.y * keep_scale : 0.f;
          av.z = ln_half(r.z, i) >= thr ? av.z * keep_scale : 0.f;
          av.w = ln_half(r.w, i) >= thr ? av.w * keep_scale : 0.f;
        }
        v[i].x += av.x; v[i].y += av.y; v[i].z += av.z; v[i].w += av.w;
      }
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(sum) * inv_d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      sq += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    const float rstd = rsqrtf(warp_sum(sq) * inv_d + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c = lane + i * 32;
      if (s_out != nullptr) s_out[row * d4 + c] = v[i];
      float4 g = __ldg(gamma + c), b = __ldg(beta + c), o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      y[row * d4 + c] = o;
      if (y_r != nullptr) y_r[row * d4 + c] = tf32_rn4(o);
    }
    if (stats != nullptr && lane == 0) stats[row] = make_float2(mean, rstd);
  }
}

template <int NV>
__global__ void __launch_bounds__(kLnWarps * 32, NV <= 4 ? PA_LN_BWD_MIN_BLOCKS : 2) add_ln_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ dy2, const float4* __restrict__ s,
                                                                     const float2* __restrict__ stats, const float4* __restrict__ gamma, const float4* __restrict__ beta,
                                                                     float p_drop, uint64_t seed, uint64_t offset, int64_t rows,
                                                                     float4* __restrict__ dx, float4* __restrict__ da, int round_da, int want_dabias, float* __restrict__ partial,
                                                                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dabias,
                                                                     int prefetch, int debug) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d4 = NV * 32, d = NV * 128;
  const float inv_d = 1.f / (float)d;
  const uint32_t thr = drop_threshold16(p_drop);
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  float4 dg[NV], db[NV], dab[NV];      // dab: column sums of da = gradient of the folded linear bias
#pragma unroll
  for (int i = 0; i < NV; ++i) dg[i] = db[i] = dab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t row = (int64_t)blockIdx.x * kLnWarps + warp; row < rows; row += (int64_t)gridDim.x * kLnWarps) {
    const int64_t ahead = row + (int64_t)prefetch * gridDim.x * kLnWarps;
    if (prefetch > 0 && ahead < rows) prefetch_rows<NV>(lane, ahead, dy, s, dy2);
    const float2 st = __ldg(stats + row);
    float4 g[NV], xh[NV];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c = lane + i * 32;
      float4 dyv = __ldg(dy + row * d4 + c), sv = __ldg(s + row * d4 + c), gm = __ldg(gamma + c);
      if (dy2 != nullptr) {
        float4 e = __ldg(dy2 + row * d4 + c);
        dyv.x += e.x; dyv.y += e.y; dyv.z += e.z; dyv.w += e.w;
      }
      if (beta != nullptr) {          // `s` holds the layer's OUTPUT y: x_hat = (y - beta) / gamma, no pre-norm tensor was saved
        const float4 bt = __ldg(beta + c);
        // fast division (MUFU.RCP + FMUL): the kernel must stay HBM-bound; a gamma of exactly 0 gives x_hat = 0
        xh[i].x = gm.x != 0.f ? __fdividef(sv.x - bt.x, gm.x) : 0.f; xh[i].y = gm.y != 0.f ? __fdividef(sv.y - bt.y, gm.y) : 0.f;
        xh[i].z = gm.z != 0.f ? __fdividef(sv.z - bt.z, gm.z) : 0.f; xh[i].w = gm.w != 0.f ? __fdividef(sv.w - bt.w, gm.w) : 0.f;
      } else {
        xh[i].x = (sv.x - st.x) * st.y; xh[i].y = (sv.y - st.x) * st.y; xh[i].z = (sv.z - st.x) * st.y; xh[i].w = (sv.w - st.x) * st.y;
      }
      g[i].x = dyv.x * gm.x; g[i].y = dyv.y * gm.y; g[i].z = dyv.z * gm.z; g[i].w = dyv.w * gm.w;
      dg[i].x += dyv.x * xh[i].x; dg[i].y += dyv.y * xh[i].y; dg[i].z += dyv.z * xh[i].z; dg[i].w += dyv.w * xh[i].w;
      db[i].x += dyv.x; db[i].y += dyv.y; db[i].z += dyv.z; db[i].w += dyv.w;
      c1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      c2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
    }
    c1 = warp_sum(c1) * inv_d;
    c2 = warp_sum(c2) * inv_d;
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c = lane + i * 32;
      float4 o;
      o.x = st.y * (g[i].x - c1 - xh[i].x * c2);
      o.y = st.y * (g[i].y - c1 - xh[i].y * c2);
      o.z = st.y * (g[i].z - c1 - xh[i].z * c2);
      o.w = st.y * (g[i].w - c1 - xh[i].w * c2);
      dx[row * d4 + c] = o;
      if (da != nullptr) {
        if (p_drop > 0.f) {
          if ((i & 1) == 0) r = ln_drop_words(seed, offset, row, d4, lane, i);
          o.x = ln_half(r.x, i) >= thr ? o.x * keep_scale : 0.f;
          o.y = ln_half(r.y, i) >= thr ? o.y * keep_scale : 0.f;
          o.z = ln_half(r.z, i) >= thr ? o.z * keep_scale : 0.f;
          o.w = ln_half(r.w, i) >= thr ? o.w * keep_scale : 0.f;
        }
        da[row * d4 + c] = round_da ? tf32_rn4(o) : o;
        if (want_dabias) { dab[i].x += o.x; dab[i].y += o.y; dab[i].z += o.z; dab[i].w += o.w; }
      }
    }
  }
  // block reduce the per-warp gamma/beta/(bias) partials -> partial[block][3][d]
  __shared__ float4 sm[kLnWarps][3][NV * 32];
#pragma unroll
  for (int i = 0; i < NV; ++i) { sm[warp][0][lane + i * 32] = dg[i]; sm[warp][1][lane + i * 32] = db[i]; sm[warp][2][lane + i * 32] = dab[i]; }
  __syncthreads();
  // one vector RED per block and column group straight into dgamma / dbeta / (linear bias) -- replaces the per-block
  // partial rows + a second reduction kernel (32 extra launches and 0.38 ms per step); `partial` is unused now
  (void)partial;
  for (int i = threadIdx.x; i < (want_dabias ? 3 : 2) * d4; i += blockDim.x) {
    int which = i / d4, c = i % d4;
    float4 acc = sm[0][which][c];
#pragma unroll
    for (int w = 1; w < kLnWarps; ++w) {
      float4 t = sm[w][which][c];
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : dabias);
    if (!(debug & 1)) atomicAdd(reinterpret_cast<float4*>(dst) + c, acc);
  }
}

static int ln_grid(int64_t rows) {
  int64_t nb = (rows + kLnWarps - 1) / kLnWarps;
  return (int)(nb < ln_max_blocks() ? nb : ln_max_blocks());
}

extern "C" size_t pa_add_ln_bwd_workspace(int64_t rows, int d) { (void)rows; (void)d; return 0; }   // kept for ABI stability: no workspace needed any more

extern "C" int pa_add_ln_fwd(const float* x, const float* a, const float* a_bias, const float* gamma, const float* beta, float eps,
                             float p_drop, uint64_t seed, uint64_t offset, int64_t rows, int d, float* y, float* y_tf32, float* s,
                             float* stats, void* stream) {
  PA_CHECK_ARG(rows >= 0 && d % 128 == 0 && d <= 1024 && p_drop >= 0.f && p_drop < 1.f);
  if (rows == 0) return PA_OK;
  int grid = ln_grid(rows);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(NV)                                                                                                   \
  add_ln_fwd_kernel<NV><<<grid, kLnWarps * 32, 0, st>>>((const float4*)x, (const float4*)a, (const float4*)a_bias, (const float4*)gamma, \
                                                         (const float4*)beta, eps, p_drop, seed, offset, rows,          \
                                                         (float4*)y, (float4*)y_tf32, (float4*)s, (float2*)stats, ln_prefetch_fwd())
  switch (d / 128) {
    case 1: LAUNCH(1); break;
    case 2: LAUNCH(2); break;
    case 4: LAUNCH(4); break;
    case 8: LAUNCH(8); break;
    default: pa_set_error("pa_add_ln_fwd: d=%d unsupported (128,256,512,1024)", d); return PA_ERR_UNSUPPORTED;
  }
#undef LAUNCH
  PA_CHECK_LAUNCH();
  return PA_OK;
}

extern "C" int pa_add_ln_bwd(const float* dy, const float* dy2, const float* s, const float* stats, const float* gamma, const float* beta, float p_drop,
                             uint64_t seed, uint64_t offset, int64_t rows, int d, float* dx, float* da, int round_da, float* dgamma,
                             float* dbeta, float* d_a_bias, void* partial, void* stream) {
  PA_CHECK_ARG(rows >= 0 && d % 128 == 0 && d <= 1024 && dgamma != nullptr && dbeta != nullptr && !(d_a_bias != nullptr && da == nullptr));
  PA_CHECK_ARG((((uintptr_t)dgamma | (uintptr_t)dbeta | (uintptr_t)d_a_bias) & 15) == 0);
  if (rows == 0) return PA_OK;
  int grid = ln_grid(rows);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(NV)                                                                                                   \
  add_ln_bwd_kernel<NV><<<grid, kLnWarps * 32, 0, st>>>((const float4*)dy, (const float4*)dy2, (const float4*)s, (const float2*)stats,    \
                                                         (const float4*)gamma, (const float4*)beta, p_drop, seed, offset, rows, (float4*)dx, \
                                                         (float4*)da, round_da, d_a_bias != nullptr, (float*)partial, dgamma, dbeta, d_a_bias, ln_prefetch_depth(), ln_debug())
  switch (d / 128) {
    case 1: LAUNCH(1); break;
    case 2: LAUNCH(2); break;
    case 4: LAUNCH(4); break;
    case 8: LAUNCH(8); break;
    default: pa_set_error("pa_add_ln_bwd: d=%d unsupported", d); return PA_ERR_UNSUPPORTED;
  }
#undef LAUNCH
  PA_CHECK_LAUNCH();
  return PA_OK;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) relu_dropout_fwd_kernel(float4* __restrict__ z, int64_t n4, float p_drop, uint64_t seed, uint64_t offset) {
  const uint32_t thr = drop_threshold(p_drop);
  const float ks = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = z[i];
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    if (p_drop > 0.f) {
      uint4 r = philox4x32(seed, (uint64_t)i, offset);
      v.x = r.x >= thr ? v.x * ks : 0.f; v.y = r.y >= thr ? v.y * ks : 0.f;
      v.z = r.z >= thr ? v.z * ks : 0.f; v.w = r.w >= thr ? v.w * ks : 0.f;
    }
    z[i] = v;
  }
}

__global__ void __launch_bounds__(256) relu_dropout_bwd_kernel(const float4* __restrict__ out, float4* __restrict__ g, int64_t n4, float ks, int round_out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 o = __ldg(out + i), v = g[i];
    v.x = o.x > 0.f ? v.x * ks : 0.f; v.y = o.y > 0.f ? v.y * ks : 0.f;
    v.z = o.z > 0.f ? v.z * ks : 0.f; v.w = o.w > 0.f ? v.w * ks : 0.f;
    g[i] = round_out ? tf32_rn4(v) : v;
  }
}

extern "C" int pa_relu_dropout_fwd(float* z, int64_t n, float p_drop, uint64_t seed, uint64_t offset, void* stream) {
  PA_CHECK_ARG(n >= 0 && n % 4 == 0 && p_drop >= 0.f && p_drop < 1.f);
  if (n == 0) return PA_OK;
  int64_t n4 = n / 4;
  int grid = (int)((n4 + 255) / 256 < pa_num_sms() * 16 ? (n4 + 255) / 256 : pa_num_sms() * 16);
  relu_dropout_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((float4*)z, n4, p_drop, seed, offset);
  PA_CHECK_LAUNCH();
  return PA_OK;
}

// Same as relu_dropout_bwd_kernel for a [rows, N] gradient, plus its column sums (= bias gradient of the
// linear that produced the activation): a thread keeps one float4 column group, blocks stride over rows.
__global__ void __launch_bounds__(256) relu_dropout_bwd_colsum_kernel(const float4* __restrict__ out, float4* __restrict__ g, int64_t rows, int cols4,
                                                                        float ks, int round_out, float* __restrict__ dbias) {
  const int c = threadIdx.x % cols4, r_in = threadIdx.x / cols4, rpi = blockDim.x / cols4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r = (int64_t)blockIdx.x * rpi + r_in; r < rows; r += (int64_t)gridDim.x * rpi) {
    const int64_t i = r * cols4 + c;
    float4 o = __ldg(out + i), v = g[i];
    v.x = o.x > 0.f ? v.x * ks : 0.f; v.y = o.y > 0.f ? v.y * ks : 0.f;
    v.z = o.z > 0.f ? v.z * ks : 0.f; v.w = o.w > 0.f ? v.w * ks : 0.f;
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    g[i] = round_out ? tf32_rn4(v) : v;
  }
  atomicAdd(reinterpret_cast<float4*>(dbias) + c, acc);
}

extern "C" int pa_relu_dropout_bwd_colsum(const float* out, float* g, int64_t rows, int N, float p_drop, int round_tf32,
                                          float* dbias, void* stream) {
  PA_CHECK_ARG(rows >= 0 && N % 4 == 0 && N / 4 <= 256 && 256 % (N / 4) == 0 && dbias != nullptr && p_drop >= 0.f && p_drop < 1.f);
  if (rows == 0) return PA_OK;
  const int cols4 = N / 4, rpi = 256 / cols4;
  int64_t nb = (rows + rpi - 1) / rpi;
  int grid = (int)(nb < pa_num_sms() * 4 ? nb : pa_num_sms() * 4);
  relu_dropout_bwd_colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)out, (float4*)g, rows, cols4,
                                                                          p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f, round_tf32, dbias);
  PA_CHECK_LAUNCH();
  return PA_OK;
}

extern "C" int pa_relu_dropout_bwd(const float* out, float* g, int64_t n, float p_drop, int round_tf32, void* stream) {
  PA_CHECK_ARG(n >= 0 && n % 4 == 0 && p_drop >= 0.f && p_drop < 1.f);
  if (n == 0) return PA_OK;
  int64_t n4 = n / 4;
  int grid = (int)((n4 + 255) / 256 < pa_num_sms() * 16 ? (n4 + 255) / 256 : pa_num_sms() * 16);
  relu_dropout_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)out, (float4*)g, n4, p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f, round_tf32);
  PA_CHECK_LAUNCH();
  return PA_OK;
}

// dst = round-to-nearest TF32 of src (weight shadow copies for the tensor-core path)
__global__ void __launch_bounds__(256) round_tf32_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) dst[i] = tf32_rn4(__ldg(src + i));
}
__global__ void round_tf32_tail_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t begin, int64_t n) {
  int64_t i = begin + threadIdx.x;
  if (i < n) dst[i] = tf32_rn(src[i]);
}

extern "C" int pa_round_tf32(const float* src, float* dst, int64_t n, void* stream) {
  PA_CHECK_ARG(n >= 0);
  if (n == 0) return PA_OK;
  int64_t n4 = n / 4;
  if (n4 > 0) {
    int grid = (int)((n4 + 255) / 256 < pa_num_sms() * 16 ? (n4 + 255) / 256 : pa_num_sms() * 16);
    round_tf32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)src, (float4*)dst, n4);
    PA_CHECK_LAUNCH();
  }
  if (n4 * 4 < n) {
    round_tf32_tail_kernel<<<1, 4, 0, (cudaStream_t)stream>>>(src, dst, n4 * 4, n);
    PA_CHECK_LAUNCH();
  }
  return PA_OK;
}
