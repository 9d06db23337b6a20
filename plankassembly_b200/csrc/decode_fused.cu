// K12: the whole KV-cached greedy decode loop as ONE persistent cooperative kernel.
//
// The reference (ref models.py:284-307) re-embeds the prefix, re-runs every decoder layer over all positions,
// re-projects the encoder memory and rebuilds the [B,t,V+t] distribution on every step, with >= 3 host syncs
// per step.  Here one launch runs every step of every sequence: the grid (2 CTAs per SM, all co-resident:
// cudaLaunchCooperativeKernel) walks a fixed list of phases per step and meets at a grid-wide barrier
// between them; nothing returns to the host until every sequence has emitted END (or T steps are done).
//
// Phases of step t (d = model width, B sequences, all fp32 -- greedy tokens must be bit-exact):
//   GEMM   split-K projections C[B,N] = X[B,K] W[N,K]^T on the CUDA cores.  A work item is 64 rows x 32
//          columns x 64 k; the k-slices go to different CTAs so that every CTA stages only 16 KB of X and
//          8 KB of W (column-splitting alone would make each of the 296 CTAs re-read all of X).  The
//          partial sums land in a [K/64][B][N] workspace and are added, in slice order (deterministic), by the
//          consumer: the attention phase (q / new k,v), the row phase (residual + LayerNorm) or the next
//          GEMM's operand staging (FFN: relu(sum + bias)).
//   ATTN   one (sequence, head) item per CTA pass: appends k,v to the self cache / streams the cross K,V
//          (projected once from the encoder memory) with 16-byte loads, PAD keys skipped; softmax in smem.
//   ROW    one sequence per CTA pass: residual + bias + LayerNorm(eps=1) [+ final LayerNorm(1e-5) -> Hfin[b,t]].
//   HEAD   vocab | pointer-feature | switch logits as ONE GEMM over the concatenated head weights, then per
//          sequence: eval distribution (ref models.py:168-186), argmax, pointer resolution, END bookkeeping
//          (ref :235-256, :306) and the NEXT step's input embedding (ref :114-138).
// Every phase is latency-bound (a dependent chain of ~70 phases per step, each: barrier -> L2/HBM loads -> math ->
// stores), so the batch is cut into `chains` independent sub-batches: CTA i works for chain i % chains, chains
// have their own barrier counter and workspace and drift apart in time, and the SM's second resident CTA belongs
// to a different chain -- one chain's barrier waits and load latencies are filled with another chain's work.
// HBM traffic per step is dominated by the cross K/V stream (2 * L * d * S_valid * 4 B per sequence); weights
// (~80 MB fp32) are read once per step and mostly stay in the 126 MB L2 because the K/V stream uses
// evict-first loads.
#include "common.cuh"

namespace {

constexpr int kThreads = 256, kWarps = kThreads / 32;
constexpr int RB = 64, NC = 32, KC = 64, XS = KC + 4;    // GEMM item: 64 rows x 32 cols x 64 k; XS = padded smem row
constexpr float kEps = 1e-6f;

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
struct ArgMax { float v; int i; };
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {   // larger value; ties -> lower index
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < kWarps; ++w) r = fmaxf(r, red[w]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) r += red[w];
  __syncthreads();
  return r;
}
__device__ __forceinline__ ArgMax block_argmax(ArgMax a, float* redv, int* redi) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax b{__shfl_xor_sync(0xffffffffu, a.v, o), __shfl_xor_sync(0xffffffffu, a.i, o)};
    a = better(a, b);
  }
  if ((threadIdx.x & 31) == 0) { redv[threadIdx.x >> 5] = a.v; redi[threadIdx.x >> 5] = a.i; }
  __syncthreads();
  ArgMax r{redv[0], redi[0]};
#pragma unroll
  for (int w = 1; w < kWarps; ++w) r = better(r, ArgMax{redv[w], redi[w]});
  __syncthreads();
  return r;
}
// 6-periodic admissibility of pointer column j for row i (ref models.py:91-101); j < i assumed
__device__ __forceinline__ bool ptr_allowed(int i, int j, int dof) {
  if (i < dof) return false;
  const int half = dof / 2;
  return j < dof ? (j == i % dof) : (j % dof == (i % dof + half) % dof);
}
__device__ __forceinline__ float4 ld_stream4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ unsigned long long g_gemm_prof[4];   // debug: ns of CTA 0 in [staging, math, write] of the GEMM items

// Grid-wide barrier: monotonically increasing arrival counter (zeroed by the host wrapper before the launch).
// Bounded spin: a protocol error must surface as a trap, never as a hung GPU.
__device__ __forceinline__ void grid_barrier(int* counter, unsigned& target, int ncta) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += ncta;
    __threadfence();
    atomicAdd(counter, 1);
    unsigned spins = 0;
    for (;;) {
      int v;
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if ((unsigned)v >= target) break;
      if (++spins > (1u << 22)) {
        printf("plank_b200: decode grid barrier timeout (block %d, target %u, seen %d)\n", blockIdx.x, target, v);
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// GEMM phase
// ------------------------------------------------------------------------------------------------
struct GemmSrc {
  const float* x; int64_t ld;        // plain: x[r*ld + k]
  const float* bias; int ksp;        // ksp > 0: x[r][k] = relu(bias[k] + sum_{s<ksp} x[(s*B + r)*ld + k])
};

// RBT = rows per item (64 / 32 / 16, picked from the chain's row count); the 256 threads form
// (RBT/4 row groups) x (8 column groups) x (KH = 128/RBT k-slices); thread tile = 4 rows x 4 columns (cg + 8c).
template <int RBT>
__device__ __forceinline__ void gemm_phase_t(float* sm_x, float* sm_w, const GemmSrc src, const float* __restrict__ W, int N, int K,
                                             int B, float* __restrict__ part, int cta, int ncta) {
  constexpr int KH = 128 / RBT;                         // k-slices inside the CTA
  const int tid = threadIdx.x;
  const int ks = K / KC, ncb = (N + NC - 1) / NC, nrb = (B + RBT - 1) / RBT;
  const int items = nrb * ncb * ks;
  const int kh = tid % KH, cg = (tid / KH) & 7, rg = tid / (8 * KH);
  for (int item = cta; item < items; item += ncta) {
    const int s = item % ks, rest = item / ks, cb = rest % ncb, rb = rest / ncb;
    const int k0 = s * KC, n0 = cb * NC, r0 = rb * RBT;
    __syncthreads();                                   // the previous item's tile reads are done
    const bool gp = blockIdx.x == 0 && tid == 0;
    unsigned long long tp0 = gp ? gtime() : 0;
    for (int idx = tid; idx < RBT * (KC / 4); idx += kThreads) {
      const int row = idx / (KC / 4), kq = idx % (KC / 4), r = r0 + row;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < B) {
        if (src.ksp == 0) {
          v = ld_cg4(src.x + (int64_t)r * src.ld + k0 + kq * 4);
        } else {
          v = __ldg(reinterpret_cast<const float4*>(src.bias + k0 + kq * 4));
          for (int s2 = 0; s2 < src.ksp; ++s2) v = add4(v, ld_cg4(src.x + ((int64_t)s2 * B + r) * src.ld + k0 + kq * 4));
          v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
        }
      }
      *reinterpret_cast<float4*>(sm_x + row * XS + kq * 4) = v;
    }
    for (int idx = tid; idx < NC * (KC / 4); idx += kThreads) {
      const int col = idx / (KC / 4), kq = idx % (KC / 4), n = n0 + col;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < N) v = __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * K + k0 + kq * 4));
      *reinterpret_cast<float4*>(sm_w + col * XS + kq * 4) = v;
    }
    __syncthreads();
    if (gp) { const unsigned long long n = gtime(); g_gemm_prof[0] += n - tp0; tp0 = n; }
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
#pragma unroll
    for (int i = 0; i < KC / 4 / KH; ++i) {
      const int ko = (KH * i + kh) * 4;
      float4 xv[4], wv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) xv[a] = *reinterpret_cast<const float4*>(sm_x + (rg * 4 + a) * XS + ko);
#pragma unroll
      for (int c = 0; c < 4; ++c) wv[c] = *reinterpret_cast<const float4*>(sm_w + (cg + 8 * c) * XS + ko);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float r = acc[a][c];
          r = fmaf(xv[a].x, wv[c].x, r); r = fmaf(xv[a].y, wv[c].y, r);
          r = fmaf(xv[a].z, wv[c].z, r); r = fmaf(xv[a].w, wv[c].w, r);
          acc[a][c] = r;
        }
    }
#pragma unroll
    for (int off = 1; off < KH; off <<= 1)
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] += __shfl_xor_sync(0xffffffffu, acc[a][c], off);
    if (gp) { const unsigned long long n = gtime(); g_gemm_prof[1] += n - tp0; tp0 = n; }
    // all KH lanes hold the same sums now; lane kh writes rows a = kh, kh + KH, ...
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int r = r0 + rg * 4 + a;
      if ((a % KH) == (kh % 4) && r < B) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int n = n0 + cg + 8 * c;
          if (n < N) part[((int64_t)s * B + r) * N + n] = acc[a][c];
        }
      }
    }
    if (gp) { const unsigned long long n = gtime(); g_gemm_prof[2] += n - tp0; g_gemm_prof[3] += 1; }
  }
}

__device__ __forceinline__ void gemm_phase(float* sm_x, float* sm_w, const GemmSrc src, const float* __restrict__ W, int N, int K,
                                           int B, float* __restrict__ part, int cta, int ncta) {
  if (B <= 16) gemm_phase_t<16>(sm_x, sm_w, src, W, N, K, B, part, cta, ncta);
  else if (B <= 32) gemm_phase_t<32>(sm_x, sm_w, src, W, N, K, B, part, cta, ncta);
  else gemm_phase_t<64>(sm_x, sm_w, src, W, N, K, B, part, cta, ncta);
}

// ------------------------------------------------------------------------------------------------
// ROW phase: y[r,:] = LN_eps(resid[r,:] + bias + sum_s part[s][r,:]); optional second LN -> out2[r*ld2 + :]
// ------------------------------------------------------------------------------------------------
constexpr int kMaxPerThread = 4;   // d <= 1024

__device__ __forceinline__ void row_phase(float* red, const float* resid, const float* part, int ks, const float* __restrict__ bias,
                                          const float* __restrict__ g, const float* __restrict__ be, float eps, float* y_out, int B,
                                          int d, const float* __restrict__ g2, const float* __restrict__ be2, float eps2,
                                          float* out2, int64_t ld2, int cta, int ncta) {
  const int tid = threadIdx.x;
  for (int r = cta; r < B; r += ncta) {
    float v[kMaxPerThread];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerThread; ++i) {
      const int c = tid + i * kThreads;
      v[i] = 0.f;
      if (c < d) {
        float a = __ldg(bias + c);
        for (int s = 0; s < ks; ++s) a += __ldcg(part + ((int64_t)s * B + r) * d + c);
        v[i] = __ldcg(resid + (int64_t)r * d + c) + a;
        sum += v[i];
      }
    }
    const float mean = block_sum(sum, red) / (float)d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerThread; ++i) {
      const int c = tid + i * kThreads;
      if (c < d) { const float u = v[i] - mean; sq += u * u; }
    }
    const float rstd = 1.f / sqrtf(block_sum(sq, red) / (float)d + eps);
    float sum2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerThread; ++i) {
      const int c = tid + i * kThreads;
      if (c < d) {
        v[i] = (v[i] - mean) * rstd * __ldg(g + c) + __ldg(be + c);
        y_out[(int64_t)r * d + c] = v[i];
        sum2 += v[i];
      }
    }
    if (g2 != nullptr) {
      const float mean2 = block_sum(sum2, red) / (float)d;
      float sq2 = 0.f;
#pragma unroll
      for (int i = 0; i < kMaxPerThread; ++i) {
        const int c = tid + i * kThreads;
        if (c < d) { const float u = v[i] - mean2; sq2 += u * u; }
      }
      const float rstd2 = 1.f / sqrtf(block_sum(sq2, red) / (float)d + eps2);
#pragma unroll
      for (int i = 0; i < kMaxPerThread; ++i) {
        const int c = tid + i * kThreads;
        if (c < d) out2[(int64_t)r * ld2 + c] = (v[i] - mean2) * rstd2 * __ldg(g2 + c) + __ldg(be2 + c);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// ATTN phase: single-query attention of every (sequence, head) over a K/V cache
// ------------------------------------------------------------------------------------------------
// scr layout: [0, DH) q | [DH, DH + kWarps*DH) per-warp partial outputs | [.., +len) scores/probabilities
template <int DH, bool SELF>
__device__ __forceinline__ void attn_phase(float* scr, float* red, const float* part, int ks, int partN, int q_off,
                                           const float* __restrict__ bias, float* k_cache, float* v_cache, int64_t ldc,
                                           int64_t cache_rows, int t, int len, const uint8_t* __restrict__ kpm, int B, int H,
                                           float scale, float* o, int cta, int ncta) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = H * DH;
  float* s_q = scr;
  float* s_o = scr + DH;
  float* s_p = scr + DH + kWarps * DH;
  for (int item = cta; item < B * H; item += ncta) {
    const int b = item / H, h = item % H;
    float* kc = k_cache + (int64_t)b * cache_rows * ldc + h * DH;
    float* vc = v_cache + (int64_t)b * cache_rows * ldc + h * DH;
    __syncthreads();                                   // scratch of the previous item is free
    if (tid < (SELF ? 3 * DH : DH)) {                  // q (and this step's k, v): bias + split-K partial sums, in slice order
      const int which = tid / DH, c = tid % DH;
      const int col = q_off + which * d + h * DH + c;
      float v = __ldg(bias + col);
      for (int s = 0; s < ks; ++s) v += __ldcg(part + ((int64_t)s * B + b) * partN + col);
      if (which == 0) s_q[c] = v;
      else if (which == 1) kc[(int64_t)t * ldc + c] = v;
      else vc[(int64_t)t * ldc + c] = v;
    }
    __syncthreads();
    // scores: 8 lanes per key, DH/8 floats per lane; 4 keys per warp pass, 4 passes in flight
    constexpr int F = DH / 8;
    const int sub = lane & 7, grp = lane >> 3;
    float qr[F];
#pragma unroll
    for (int c = 0; c < F; ++c) qr[c] = s_q[sub * F + c];
    float mx = -INFINITY;
    for (int j0 = warp * 4; j0 < len; j0 += kWarps * 4 * 4) {
      float acc[4];
      bool live[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * (kWarps * 4) + grp;
        acc[u] = 0.f;
        live[u] = j < len && !(!SELF && kpm != nullptr && __ldg(kpm + (int64_t)b * len + j));
        if (live[u]) {
          const float* kr = kc + (int64_t)j * ldc + sub * F;
#pragma unroll
          for (int c = 0; c < F; c += 4) {
            const float4 kk = SELF ? ld_cg4(kr + c) : ld_stream4(kr + c);
            acc[u] += qr[c] * kk.x + qr[c + 1] * kk.y + qr[c + 2] * kk.z + qr[c + 3] * kk.w;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * (kWarps * 4) + grp;
        float a = acc[u];
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        if (j < len && sub == 0) {
          const float sc = live[u] ? a * scale : -INFINITY;
          s_p[j] = sc;
          mx = fmaxf(mx, sc);
        }
      }
    }
    mx = block_max(mx, red);
    const float m_safe = mx == -INFINITY ? 0.f : mx;
    float sum = 0.f;
    for (int j = tid; j < len; j += kThreads) {
      const float p = expf(s_p[j] - m_safe);
      s_p[j] = p;
      sum += p;
    }
    sum = block_sum(sum, red);                          // (its barriers also publish s_p)
    // O = P V: DH/4 lanes span a row with 16-byte loads, 32/(DH/4) keys per warp instruction, 8 instructions in flight
    constexpr int LPR = DH / 4, KPI = 32 / LPR, U = 8;
    const int kg = lane / LPR, cl = (lane % LPR) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j0 = warp * KPI + kg; j0 < len; j0 += kWarps * KPI * U) {
      float4 vv[U];
      float pp[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u * (kWarps * KPI);
        pp[u] = j < len ? s_p[j] : 0.f;
        vv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pp[u] != 0.f) vv[u] = SELF ? ld_cg4(vc + (int64_t)j * ldc + cl) : ld_stream4(vc + (int64_t)j * ldc + cl);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        acc.x = fmaf(pp[u], vv[u].x, acc.x); acc.y = fmaf(pp[u], vv[u].y, acc.y);
        acc.z = fmaf(pp[u], vv[u].z, acc.z); acc.w = fmaf(pp[u], vv[u].w, acc.w);
      }
    }
#pragma unroll
    for (int off = LPR; off < 32; off <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
    }
    if (lane < LPR) *reinterpret_cast<float4*>(s_o + warp * DH + cl) = acc;
    __syncthreads();
    if (tid < DH) {
      float r = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) r += s_o[w * DH + tid];
      o[(int64_t)b * d + h * DH + tid] = sum > 0.f ? r / sum : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// HEAD phase (per sequence): eval distribution, sampling, END bookkeeping, next input embedding
// ------------------------------------------------------------------------------------------------
// scr layout: [0,V) vocab logits | [Vp, Vp+d) pointer feature (Vp = V rounded up to 4) | [Vp+d, Vp+d+T) pointer scores
struct HeadIO {                    // chain-local views (row 0 = the chain's first sequence)
  const float* hfin; int64_t* samples; int64_t* attach; int32_t* first_end; float* y;
};
__device__ __forceinline__ void head_phase(float* scr, float* red, int* redi, const pa_decode_fused_args& a, const HeadIO io,
                                           int B, const float* part, int ks, int t, int cta, int ncta) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int V = a.V, d = a.d, T = a.T, NH = V + d + 1;
  const int Vp = (V + 3) & ~3;
  float* s_lv = scr;
  float* s_pf = scr + Vp;
  float* s_lp = scr + Vp + d;
  __shared__ float s_sw;
  __shared__ long long s_tok;
  for (int b = cta; b < B; b += ncta) {
    __syncthreads();
    for (int c = tid; c < NH; c += kThreads) {
      float v = __ldg(a.b_heads + c);
      for (int s = 0; s < ks; ++s) v += __ldcg(part + ((int64_t)s * B + b) * NH + c);
      if (c < V) s_lv[c] = v;
      else if (c < V + d) s_pf[c - V] = v;
      else s_sw = v;
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int c = tid; c < V; c += kThreads) mx = fmaxf(mx, s_lv[c]);
    mx = block_max(mx, red);
    float sum = 0.f;
    for (int c = tid; c < V; c += kThreads) sum += expf(s_lv[c] - mx);
    sum = block_sum(sum, red);
    const int sz = t + 1;
    const bool with_ptr = sz >= a.dof;
    const float pi = with_ptr ? 1.f / (1.f + expf(-s_sw)) : 0.f;
    ArgMax best{-INFINITY, 0x7fffffff};
    for (int c = tid; c < V; c += kThreads) {
      float p = expf(s_lv[c] - mx) / sum;
      if (with_ptr) p *= (1.f - pi);
      best = better(best, ArgMax{p, c});
    }
    if (with_ptr) {
      const float inv_d = 1.f / (float)d;
      const float* hf = io.hfin + (int64_t)b * T * d;
      for (int j = warp; j < t; j += kWarps) {          // scores against the cached final hiddens of positions j < t
        float acc = 0.f;
        for (int c = lane * 4; c < d; c += 128) {
          const float4 p4 = *reinterpret_cast<const float4*>(s_pf + c);
          const float4 k4 = ld_cg4(hf + (int64_t)j * d + c);
          acc += p4.x * k4.x + p4.y * k4.y + p4.z * k4.z + p4.w * k4.w;
        }
        acc = warp_sum(acc);
        if (lane == 0) s_lp[j] = acc * inv_d;
      }
      __syncthreads();
      float pm = -INFINITY;
      for (int j = tid; j < t; j += kThreads) pm = fmaxf(pm, s_lp[j]);
      pm = block_max(pm, red);
      float ps = 0.f;
      for (int j = tid; j < t; j += kThreads) ps += expf(s_lp[j] - pm);
      ps = block_sum(ps, red);
      const int i = sz - 1;
      for (int j = tid; j < sz; j += kThreads) {
        float p = kEps;
        if (j < t && ptr_allowed(i, j, a.dof)) p = expf(s_lp[j] - pm) / ps * pi;
        best = better(best, ArgMax{p, V + j});
      }
    }
    best = block_argmax(best, red, redi);
    if (tid == 0) {
      long long tok = best.i, att = -1;
      if (best.i >= V) { att = best.i - V; tok = io.samples[(int64_t)b * T + att]; }
      io.samples[(int64_t)b * T + t] = tok;
      io.attach[(int64_t)b * T + t] = att;
      if (tok == a.end_token && io.first_end[b] >= T) {        // first END of this sequence
        io.first_end[b] = t;
        atomicMax(a.state + 3, t);                             // the reference stops after step max_b first_end
        __threadfence();
        atomicAdd(a.state + 1, 1);
      }
      s_tok = tok;
    }
    __syncthreads();
    if (t + 1 < T) {                                    // input of step t+1 (ref models.py:114-138 with the shift)
      const long long tok = s_tok;
      const float* ev = a.e_val + tok * d;
      const float* ec = a.e_coord + (int64_t)(t % a.dof) * d;
      const float* ep = a.e_pos + (int64_t)(t / a.dof) * d;
      for (int c = tid; c < d; c += kThreads) io.y[(int64_t)b * d + c] = (__ldg(ev + c) + __ldg(ec + c)) + __ldg(ep + c);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
// Optional phase profile (state[4] != 0 on entry): thread 0 of block 0 accumulates nanoseconds per phase kind
// into g_decode_prof: 0 gemm, 1 self-attn, 2 row, 3 cross-attn, 4 head, 5 barriers.
__device__ unsigned long long g_decode_prof[16];
#define PROF(slot, ...)                                                    \
  do {                                                                     \
    if (prof_on) { const unsigned long long t0__ = gtime(); __VA_ARGS__; prof[slot] += gtime() - t0__; } \
    else { __VA_ARGS__; }                                                         \
  } while (0)

__host__ __device__ inline size_t part_row_floats(int d, int ff, int V) {   // floats per sequence in ONE workspace half
  const size_t ksd = d / KC, ksf = ff / KC;
  size_t n = 3 * (size_t)d;
  if ((size_t)ff > n) n = ff;
  if ((size_t)(V + d + 1) > n) n = V + d + 1;
  const size_t a = ksd * n, b = (ksd > ksf ? ksd : ksf) * (size_t)d;
  return a > b ? a : b;
}

template <int DH>
__global__ void __launch_bounds__(kThreads, 2) decode_fused_kernel(const __grid_constant__ pa_decode_fused_args a) {
  extern __shared__ __align__(16) float smem[];
  float* sm_x = smem;                     // GEMM tiles; the ATTN/HEAD scratch aliases them (phases are barrier-separated)
  float* sm_w = smem + RB * XS;
  float* scr = smem;
  __shared__ float red[kWarps];
  __shared__ int redi[kWarps];
  const int tid = threadIdx.x;
  const int d = a.d, T = a.T, S = a.S, H = a.H, ff = a.ff;
  const int ksd = d / KC, ksf = ff / KC, NH = a.V + d + 1;
  // this CTA's chain: an independent sub-batch with its own barrier counter and workspace
  const int nch = a.chains, chain = blockIdx.x % nch, cta = blockIdx.x / nch, ncta = gridDim.x / nch;
  const int rpc = (a.B + nch - 1) / nch, r0 = chain * rpc;
  const int B = min(rpc, a.B - r0);                       // rows of this chain (<= 0: nothing to do)
  if (B <= 0) return;
  int* bar = a.state + 8 + chain;
  float* partA = a.part + (size_t)chain * 2 * rpc * part_row_floats(d, ff, a.V);
  float* partB = partA + (size_t)rpc * part_row_floats(d, ff, a.V);
  float* y = a.y + (int64_t)r0 * d;
  float* o = a.o + (int64_t)r0 * d;
  float* hfin = a.hfin + (int64_t)r0 * T * d;
  const uint8_t* kpm = a.kpm + (int64_t)r0 * S;
  const HeadIO io{hfin, a.samples + (int64_t)r0 * T, a.attach + (int64_t)r0 * T, a.first_end + r0, y};
  unsigned target = 0;
  const float scale = rsqrtf((float)DH);
  const bool prof_on = a.profile != 0 && blockIdx.x == 0;
  unsigned long long prof[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

  // prologue: output buffers and the first input row (zeros: ref models.py:130-137 shift)
  for (int64_t i = (int64_t)cta * kThreads + tid; i < (int64_t)B * T; i += (int64_t)ncta * kThreads) {
    io.samples[i] = 0;
    io.attach[i] = -1;
  }
  for (int64_t i = (int64_t)cta * kThreads + tid; i < (int64_t)B * d; i += (int64_t)ncta * kThreads) y[i] = 0.f;
  for (int i = cta * kThreads + tid; i < B; i += ncta * kThreads) io.first_end[i] = T;
  grid_barrier(bar, target, ncta);

  int t = 0;
  for (; t < T; ++t) {
    for (int l = 0; l < a.L; ++l) {
      const pa_decode_layer& ly = a.layers[l];
      float* self_k = ly.self_k + (int64_t)r0 * T * d;
      float* self_v = ly.self_v + (int64_t)r0 * T * d;
      float* cross_k = const_cast<float*>(ly.cross_kv) + (int64_t)r0 * S * 2 * d;
      PROF(0, gemm_phase(sm_x, sm_w, GemmSrc{y, d, nullptr, 0}, ly.w_sqkv, 3 * d, d, B, partA, cta, ncta));
      PROF(5, grid_barrier(bar, target, ncta));
      PROF(1, attn_phase<DH, true>(scr, red, partA, ksd, 3 * d, 0, ly.b_sqkv, self_k, self_v, d, T, t, t + 1, nullptr, B, H, scale, o, cta, ncta));
      PROF(5, grid_barrier(bar, target, ncta));
      PROF(6, gemm_phase(sm_x, sm_w, GemmSrc{o, d, nullptr, 0}, ly.w_so, d, d, B, partB, cta, ncta));
      PROF(5, grid_barrier(bar, target, ncta));
      PROF(2, row_phase(red, y, partB, ksd, ly.b_so, ly.g1, ly.be1, a.layer_eps, y, B, d, nullptr, nullptr, 0.f, nullptr, 0, cta, ncta));
      PROF(5, grid_barrier(bar, target, ncta));
      PROF(7, gemm_phase(sm_x, sm_w, GemmSrc{y, d, nullptr, 0}, ly.w_cq, d, d, B, partA, cta, ncta));
      PROF(5, grid_barrier(bar, target, ncta));
      PROF(3, attn_phase<DH, false>(scr, red, partA, ksd, d, 0, ly.b_cq, cross_k, cross_k + d, 2 * d, S, 0, S, kpm, B, H, scale, o, cta, ncta));
      PROF(5, grid_barrier(bar, target, ncta));
      PROF(6, gemm_phase(sm_x, sm_w, GemmSrc{o, d, nullptr, 0}, ly.w_co, d, d, B, partB, cta, ncta));
      PROF(5, grid_barrier(bar, target, ncta));
      PROF(2, row_phase(red, y, partB, ksd, ly.b_co, ly.g2, ly.be2, a.layer_eps, y, B, d, nullptr, nullptr, 0.f, nullptr, 0, cta, ncta));
      PROF(5, grid_barrier(bar, target, ncta));
      PROF(8, gemm_phase(sm_x, sm_w, GemmSrc{y, d, nullptr, 0}, ly.w_f1, ff, d, B, partA, cta, ncta));
      PROF(5, grid_barrier(bar, target, ncta));
      PROF(9, gemm_phase(sm_x, sm_w, GemmSrc{partA, ff, ly.b_f1, ksd}, ly.w_f2, d, ff, B, partB, cta, ncta));
      PROF(5, grid_barrier(bar, target, ncta));
      const bool last = l + 1 == a.L;
      PROF(2, row_phase(red, y, partB, ksf, ly.b_f2, ly.g3, ly.be3, a.layer_eps, y, B, d, last ? a.gf : nullptr, a.bf, a.final_eps,
                        hfin + (int64_t)t * d, (int64_t)T * d, cta, ncta));
      PROF(5, grid_barrier(bar, target, ncta));
    }
    PROF(10, gemm_phase(sm_x, sm_w, GemmSrc{hfin + (int64_t)t * d, (int64_t)T * d, nullptr, 0}, a.w_heads, NH, d, B, partA, cta, ncta));
    PROF(5, grid_barrier(bar, target, ncta));
    PROF(4, head_phase(scr, red, redi, a, io, B, partA, ksd, t, cta, ncta));
    PROF(5, grid_barrier(bar, target, ncta));
    // The reference stops after the step at which EVERY sequence of the batch has emitted END (ref models.py:306) and
    // keeps decoding finished rows until then; chains run at their own pace, so a chain goes on until all chains'
    // rows have ended AND it has itself reached the last of those steps (state[3] is final once state[1] == batch).
    if (__ldcg(a.state + 1) >= a.B && t >= __ldcg(a.state + 3)) { ++t; break; }
  }
  if (cta == 0 && tid == 0) {
    atomicMax(a.state + 2, t);                           // steps executed by the slowest chain
    if (prof_on)
      for (int i = 0; i < 12; ++i) g_decode_prof[i] = prof[i];
  }
}

template <int DH>
int launch(const pa_decode_fused_args& a_in, cudaStream_t st) {
  pa_decode_fused_args a = a_in;
  auto kern = decode_fused_kernel<DH>;
  size_t gemm_f = (size_t)(RB + NC) * XS;
  size_t attn_f = (size_t)DH + (size_t)kWarps * DH + (size_t)(a.S > a.T ? a.S : a.T);
  size_t head_f = (size_t)((a.V + 3) & ~3) + a.d + a.T;
  size_t fl = gemm_f > attn_f ? gemm_f : attn_f;
  if (head_f > fl) fl = head_f;
  const size_t smem = fl * sizeof(float);
  if (smem > 48 * 1024) PA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0, per_sm = 0;
  PA_CUDA(cudaGetDevice(&dev));
  PA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  PA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
  if (per_sm < 1) { pa_set_error("pa_decode_fused: kernel does not fit on an SM (smem %zu B)", smem); return PA_ERR_UNSUPPORTED; }
  if (per_sm > 2) per_sm = 2;
  if (a.chains <= 0) a.chains = a.B >= 32 ? 2 : 1;       // sub-batches of >= 16 sequences
  if (a.chains > PA_MAX_DEC_CHAINS) a.chains = PA_MAX_DEC_CHAINS;
  if (a.chains > a.B) a.chains = a.B;
  int grid = per_sm * sms;
  grid -= grid % a.chains;                               // every chain gets the same number of CTAs
  PA_CHECK_ARG(grid >= a.chains);
  PA_CUDA(cudaMemsetAsync(a.state, 0, PA_DEC_STATE_INTS * sizeof(int32_t), st));
  void* params[] = {(void*)&a};
  PA_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(kThreads), params, smem, st));
  return PA_OK;
}

}  // namespace

extern "C" int pa_debug_decode_prof(unsigned long long* out8_host) {
  PA_CUDA(cudaMemcpyFromSymbol(out8_host, g_decode_prof, sizeof(unsigned long long) * 12));
  PA_CUDA(cudaMemcpyFromSymbol(out8_host + 12, g_gemm_prof, sizeof(unsigned long long) * 4));
  unsigned long long z[4] = {0, 0, 0, 0};
  PA_CUDA(cudaMemcpyToSymbol(g_gemm_prof, z, sizeof(z)));
  return PA_OK;
}

extern "C" size_t pa_decode_fused_workspace(int B, int d, int ff, int V) {
  // two halves per chain; chains round their row count up, hence the PA_MAX_DEC_CHAINS extra rows
  return 2 * (size_t)(B + PA_MAX_DEC_CHAINS) * part_row_floats(d, ff, V) * sizeof(float);
}

extern "C" int pa_decode_fused(const pa_decode_fused_args* a, void* stream) {
  PA_CHECK_ARG(a->B > 0 && a->S > 0 && a->T > 0 && a->L > 0 && a->L <= PA_MAX_DEC_LAYERS && a->H > 0);
  PA_CHECK_ARG(a->d % KC == 0 && a->ff % KC == 0 && a->d <= kMaxPerThread * kThreads && a->d % a->H == 0 && a->dof >= 2);
  PA_CHECK_ARG(a->part != nullptr && a->state != nullptr &&
               (size_t)a->part_bytes >= pa_decode_fused_workspace(a->B, a->d, a->ff, a->V));
  switch (a->d / a->H) {
    case 32: return launch<32>(*a, (cudaStream_t)stream);
    case 64: return launch<64>(*a, (cudaStream_t)stream);
    default: pa_set_error("pa_decode_fused: head dim %d unsupported (32, 64)", a->d / a->H); return PA_ERR_UNSUPPORTED;
  }
}
