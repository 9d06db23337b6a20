// K3/K4/K5, exact mode: flash-style multi-head attention in fp32 on the CUDA cores.
// Forward + backward (dK/dV kernel, dQ kernel), key-padding + causal masks, Philox dropout on the
// probabilities.  This is the bit-faithful (fp32 accumulate, fp32 operands) path used for parity
// and for the greedy-decode prefill; the tensor-core path lives in attn_tc.cu.
//   Reference semantics: torch nn/functional.py multi_head_attention_forward as called by the
//   post-norm layers built at ref models.py:60-69.
//
// Tiling: 64 queries x 64 keys per step, 256 threads as a 16x16 grid, each thread a 4x4 patch of
// the score tile.  K/V tiles are stored row-permuted (key k at smem row (k%4)*16 + k/4) so that a
// thread's four score columns are the four CONSECUTIVE keys 4*tx..4*tx+3 (one Philox group) while
// its 16-byte smem reads stay bank-conflict free.
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, NT = 256;
constexpr int PS = BN + 4;  // row stride of the P / dS tiles

__device__ __forceinline__ int perm_row(int k) { return (k & 3) * 16 + (k >> 2); }

template <int W> __device__ __forceinline__ void load_w(const float* __restrict__ p, float (&v)[W]);
template <> __device__ __forceinline__ void load_w<4>(const float* __restrict__ p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void load_w<2>(const float* __restrict__ p, float (&v)[2]) {
  float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y;
}

// load a [64 x DH] tile (rows r0.., head slice) into smem with row stride DH+4; rows >= L -> 0
template <int DH, bool PERM>
__device__ __forceinline__ void load_tile(float* __restrict__ dst, const float* __restrict__ src, int64_t ld, int r0, int L) {
  constexpr int C4 = DH / 4;
  for (int i = threadIdx.x; i < 64 * C4; i += NT) {
    int r = i / C4, c = i % C4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < L) v = __ldg(reinterpret_cast<const float4*>(src + (int64_t)(r0 + r) * ld) + c);
    int rr = PERM ? perm_row(r) : r;
    *reinterpret_cast<float4*>(dst + rr * (DH + 4) + c * 4) = v;
  }
}

// acc[r][j] = sum_c A[ty*4+r][c] * Bp[j*16+tx][c]   (Bp row-permuted => column = key 4*tx+j)
template <int DH>
__device__ __forceinline__ void tile_dot(const float* __restrict__ A, const float* __restrict__ Bp, int ty, int tx, float acc[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;
#pragma unroll 4
  for (int c = 0; c < DH; c += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(A + (ty * 4 + r) * (DH + 4) + c);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(Bp + (j * 16 + tx) * (DH + 4) + c);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        acc[r][j] += a[r].x * b[j].x + a[r].y * b[j].y + a[r].z * b[j].z + a[r].w * b[j].w;
  }
}

// out[r][.] += sum_k P[ty*4+r][k] * B[row(k)][cols of this thread];  thread owns DH/16 contiguous cols
template <int DH, bool PERM>
__device__ __forceinline__ void tile_pv(const float* __restrict__ P, const float* __restrict__ B, int ty, int tx, float (*acc)[DH / 16]) {
  constexpr int W = DH / 16;
#pragma unroll 2
  for (int k = 0; k < BN; k += 4) {
    float4 p[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) p[r] = *reinterpret_cast<const float4*>(P + (ty * 4 + r) * PS + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float* brow = B + (PERM ? perm_row(k + kk) : (k + kk)) * (DH + 4) + tx * W;
      float bv[W];
      load_w<W>(brow, bv);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float pv = kk == 0 ? p[r].x : kk == 1 ? p[r].y : kk == 2 ? p[r].z : p[r].w;
#pragma unroll
        for (int w = 0; w < W; ++w) acc[r][w] += pv * bv[w];
      }
    }
  }
}

// out[r][.] += sum_i A[i][ty*4+r] * B[i][cols]   (transposed-A product used for dK/dV)
template <int DH>
__device__ __forceinline__ void tile_atb(const float* __restrict__ A, const float* __restrict__ B, int ty, int tx, float (*acc)[DH / 16]) {
  constexpr int W = DH / 16;
#pragma unroll 4
  for (int i = 0; i < BM; ++i) {
    float4 a = *reinterpret_cast<const float4*>(A + i * PS + ty * 4);
    const float* brow = B + i * (DH + 4) + tx * W;
    float bv[W];
    load_w<W>(brow, bv);
#pragma unroll
    for (int w = 0; w < W; ++w) {
      acc[0][w] += a.x * bv[w]; acc[1][w] += a.y * bv[w]; acc[2][w] += a.z * bv[w]; acc[3][w] += a.w * bv[w];
    }
  }
}

struct DropCtx {
  uint64_t seed, offset;
  uint32_t thr;
  float ks;
  bool on;
};

// keep-mask for the 4 consecutive keys kj0..kj0+3 (kj0 % 4 == 0) of query row `row_global` (= (b*H+h)*Lq + qi): bits of the
// row's keep word for key word kj0/32, recomputed here by the same rule as the bit planes of dropmask.cu (drop_keep_word)
__device__ __forceinline__ void drop_mask4(const DropCtx& D, int64_t row_global, int LkW, int kj0, float m[4]) {
  const uint32_t w = drop_keep_word(D.seed, D.offset, row_global, LkW, kj0 >> 5, 65536u - D.thr) >> (kj0 & 31);
  m[0] = (w & 1u) ? D.ks : 0.f; m[1] = (w & 2u) ? D.ks : 0.f; m[2] = (w & 4u) ? D.ks : 0.f; m[3] = (w & 8u) ? D.ks : 0.f;
}

template <int DH>
__global__ void __launch_bounds__(NT) attn_fwd_simt_kernel(pa_attn_fwd_args A) {
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;                    // [64][DH+4]
  float* Ks = Qs + 64 * (DH + 4);      // permuted
  float* Vs = Ks + 64 * (DH + 4);      // natural
  float* Ps = Vs + 64 * (DH + 4);      // [64][PS]
  float* kb = Ps + 64 * PS;            // [64] additive key bias 0 / -inf
  constexpr int W = DH / 16;
  const int q0 = blockIdx.x * BM, h = blockIdx.y, b = blockIdx.z;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* qp = A.q + (int64_t)b * A.Lq * A.ldq + h * DH;
  const float* kp = A.k + (int64_t)b * A.Lk * A.ldk + h * DH;
  const float* vp = A.v + (int64_t)b * A.Lk * A.ldv + h * DH;
  DropCtx D{A.seed, A.offset, drop_threshold16(A.p_drop), A.p_drop > 0.f ? 1.f / (1.f - A.p_drop) : 1.f, A.p_drop > 0.f};
  const int Lk4 = (A.Lk + 31) / 32;    // keep words (32 keys) per row

  load_tile<DH, false>(Qs, qp, A.ldq, q0, A.Lq);
  float m_run[4], l_run[4], o[4][W];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    m_run[r] = -INFINITY; l_run[r] = 0.f;
#pragma unroll
    for (int w = 0; w < W; ++w) o[r][w] = 0.f;
  }
  int kv_end = A.causal ? min(A.Lk, q0 + BM) : A.Lk;
  // keys beyond kv_len[b] (1 + last non-PAD key) carry -inf: whole key tiles of padding are skipped (exact)
  if (A.kv_len != nullptr && A.kpm != nullptr) kv_end = min(kv_end, max(A.kv_len[b], 1));
  for (int k0 = 0; k0 < kv_end; k0 += BN) {
    __syncthreads();
    load_tile<DH, true>(Ks, kp, A.ldk, k0, A.Lk);
    load_tile<DH, false>(Vs, vp, A.ldv, k0, A.Lk);
    if (threadIdx.x < BN) {
      int kj = k0 + threadIdx.x;
      bool ok = kj < A.Lk && !(A.kpm != nullptr && A.kpm[(int64_t)b * A.Lk + kj]);
      kb[threadIdx.x] = ok ? 0.f : -INFINITY;
    }
    __syncthreads();
    float s[4][4];
    tile_dot<DH>(Qs, Ks, ty, tx, s);
    float4 bias = *reinterpret_cast<const float4*>(kb + tx * 4);
    const float bz[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int qi = q0 + ty * 4 + r;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int kj = k0 + tx * 4 + j;
        float v = s[r][j] * A.scale + bz[j];
        if (A.causal && kj > qi) v = -INFINITY;
        s[r][j] = v;
        mx = fmaxf(mx, v);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m_run[r], mx);
      const float m_safe = m_new == -INFINITY ? 0.f : m_new;
      const float corr = __expf(m_run[r] - m_safe);
      float rs = 0.f, pr[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { pr[j] = __expf(s[r][j] - m_safe); rs += pr[j]; }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      l_run[r] = l_run[r] * corr + rs;
      m_run[r] = m_new;
#pragma unroll
      for (int w = 0; w < W; ++w) o[r][w] *= corr;
      if (D.on) {
        float mk[4];
        drop_mask4(D, ((int64_t)(b * A.H + h) * A.Lq + qi), Lk4, k0 + tx * 4, mk);
#pragma unroll
        for (int j = 0; j < 4; ++j) pr[j] *= mk[j];
      }
      *reinterpret_cast<float4*>(Ps + (ty * 4 + r) * PS + tx * 4) = make_float4(pr[0], pr[1], pr[2], pr[3]);
    }
    __syncthreads();
    tile_pv<DH, false>(Ps, Vs, ty, tx, o);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int qi = q0 + ty * 4 + r;
    if (qi >= A.Lq) continue;
    const float inv = l_run[r] > 0.f ? 1.f / l_run[r] : 0.f;
    float* op = A.o + ((int64_t)b * A.Lq + qi) * A.ldo + h * DH + tx * W;
#pragma unroll
    for (int w = 0; w < W; ++w) op[w] = A.round_out ? tf32_rn(o[r][w] * inv) : o[r][w] * inv;
    if (A.lse != nullptr && tx == 0)
      A.lse[((int64_t)b * A.H + h) * A.Lq + qi] = l_run[r] > 0.f ? m_run[r] + __logf(l_run[r]) : -INFINITY;
  }
}

// delta[b,h,i] = sum_c dO[b,i,h,c] * O[b,i,h,c]
// delta[b,h,i] = sum_c O[b,i,h,c] * dO[b,i,h,c].  HBM-bound (reads O and dO once): four lanes share one (row, head)
// slice so that a warp keeps 8 x 16-byte loads per lane in flight over fully coalesced 256/128-byte segments.
__global__ void __launch_bounds__(256) attn_delta_kernel(const float* __restrict__ o, const float* __restrict__ d_o, int64_t ldo, int B,
                                                          int H, int Lq, int dh, float* __restrict__ delta) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t idx = gid >> 2;                       // (b, i, h)
  const int part = (int)(gid & 3);
  const bool live = idx < (int64_t)B * Lq * H;
  float acc = 0.f;
  int h = 0, i = 0, b = 0;
  if (live) {
    h = (int)(idx % H);
    const int64_t bi = idx / H;
    i = (int)(bi % Lq); b = (int)(bi / Lq);
    const float4* po = reinterpret_cast<const float4*>(o + bi * ldo + h * dh);
    const float4* pd = reinterpret_cast<const float4*>(d_o + bi * ldo + h * dh);
    const int n4 = dh / 4;                            // 8 or 16 float4 per slice; lane `part` takes c = part, part + 4, ...
    float4 x[4], y[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = part + 4 * u;
      x[u] = c < n4 ? __ldg(po + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      y[u] = c < n4 ? __ldg(pd + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += x[u].x * y[u].x + x[u].y * y[u].y + x[u].z * y[u].z + x[u].w * y[u].w;
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  if (live && part == 0) delta[((int64_t)b * H + h) * Lq + i] = acc;
}

// Recompute one 64x64 tile of P (dropped, -> Ps) and dS (-> dSs) from Q,K,V,dO tiles in smem.
template <int DH>
__device__ __forceinline__ void recompute_p_ds(const pa_attn_bwd_args& A, const DropCtx& D, const float* Qs, const float* dOs,
                                               const float* Ks, const float* Vs, const float* kb, const float* lse_s,
                                               const float* delta_s, float* Ps, float* dSs, int b, int h, int q0, int k0, int ty,
                                               int tx, int Lk4) {
  float s[4][4], dp[4][4];
  tile_dot<DH>(Qs, Ks, ty, tx, s);
  tile_dot<DH>(dOs, Vs, ty, tx, dp);
  float4 bias = *reinterpret_cast<const float4*>(kb + tx * 4);
  const float bz[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int qi = q0 + ty * 4 + r;
    const float lse = lse_s[ty * 4 + r], dl = delta_s[ty * 4 + r];
    float mk[4] = {1.f, 1.f, 1.f, 1.f};
    if (D.on) drop_mask4(D, ((int64_t)(b * A.H + h) * A.Lq + qi), Lk4, k0 + tx * 4, mk);
    float pd[4], ds[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int kj = k0 + tx * 4 + j;
      float v = s[r][j] * A.scale + bz[j];
      bool dead = (A.causal && kj > qi) || v == -INFINITY || lse == -INFINITY || qi >= A.Lq;
      float p = dead ? 0.f : __expf(v - lse);
      pd[j] = p * mk[j];
      ds[j] = p * (dp[r][j] * mk[j] - dl);
    }
    *reinterpret_cast<float4*>(Ps + (ty * 4 + r) * PS + tx * 4) = make_float4(pd[0], pd[1], pd[2], pd[3]);
    *reinterpret_cast<float4*>(dSs + (ty * 4 + r) * PS + tx * 4) = make_float4(ds[0], ds[1], ds[2], ds[3]);
  }
}

template <int DH>
__global__ void __launch_bounds__(NT) attn_bwd_dkdv_simt_kernel(pa_attn_bwd_args A) {
  extern __shared__ __align__(16) float smem[];
  constexpr int TS = 64 * (DH + 4);
  float* Qs = smem; float* dOs = Qs + TS; float* Ks = dOs + TS; float* Vs = Ks + TS;
  float* Ps = Vs + TS; float* dSs = Ps + 64 * PS;
  float* kb = dSs + 64 * PS; float* lse_s = kb + 64; float* delta_s = lse_s + 64;
  constexpr int W = DH / 16;
  const int k0 = blockIdx.x * BN, h = blockIdx.y, b = blockIdx.z;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* qp = A.q + (int64_t)b * A.Lq * A.ldq + h * DH;
  const float* dop = A.d_o + (int64_t)b * A.Lq * A.ldo + h * DH;
  const float* kp = A.k + (int64_t)b * A.Lk * A.ldk + h * DH;
  const float* vp = A.v + (int64_t)b * A.Lk * A.ldv + h * DH;
  DropCtx D{A.seed, A.offset, drop_threshold16(A.p_drop), A.p_drop > 0.f ? 1.f / (1.f - A.p_drop) : 1.f, A.p_drop > 0.f};
  const int Lk4 = (A.Lk + 31) / 32;    // keep words (32 keys) per row
  load_tile<DH, true>(Ks, kp, A.ldk, k0, A.Lk);
  load_tile<DH, true>(Vs, vp, A.ldv, k0, A.Lk);
  if (threadIdx.x < BN) {
    int kj = k0 + threadIdx.x;
    bool ok = kj < A.Lk && !(A.kpm != nullptr && A.kpm[(int64_t)b * A.Lk + kj]);
    kb[threadIdx.x] = ok ? 0.f : -INFINITY;
  }
  float dk[4][W], dv[4][W];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int w = 0; w < W; ++w) dk[r][w] = dv[r][w] = 0.f;
  const int q_begin = A.causal ? (k0 / BM) * BM : 0;
  for (int q0 = q_begin; q0 < A.Lq; q0 += BM) {
    __syncthreads();
    load_tile<DH, false>(Qs, qp, A.ldq, q0, A.Lq);
    load_tile<DH, false>(dOs, dop, A.ldo, q0, A.Lq);
    if (threadIdx.x < BM) {
      int qi = q0 + threadIdx.x;
      lse_s[threadIdx.x] = qi < A.Lq ? A.lse[((int64_t)b * A.H + h) * A.Lq + qi] : -INFINITY;
      delta_s[threadIdx.x] = qi < A.Lq ? A.delta[((int64_t)b * A.H + h) * A.Lq + qi] : 0.f;
    }
    __syncthreads();
    recompute_p_ds<DH>(A, D, Qs, dOs, Ks, Vs, kb, lse_s, delta_s, Ps, dSs, b, h, q0, k0, ty, tx, Lk4);
    __syncthreads();
    tile_atb<DH>(Ps, dOs, ty, tx, dv);
    tile_atb<DH>(dSs, Qs, ty, tx, dk);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int kj = k0 + ty * 4 + r;
    if (kj >= A.Lk) continue;
    float* pk = A.dk + ((int64_t)b * A.Lk + kj) * A.lddk + h * DH + tx * W;
    float* pv = A.dv + ((int64_t)b * A.Lk + kj) * A.lddv + h * DH + tx * W;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      float a = dk[r][w] * A.scale, c = dv[r][w];
      pk[w] = A.round_out ? tf32_rn(a) : a; pv[w] = A.round_out ? tf32_rn(c) : c;
    }
  }
}

template <int DH>
__global__ void __launch_bounds__(NT) attn_bwd_dq_simt_kernel(pa_attn_bwd_args A) {
  extern __shared__ __align__(16) float smem[];
  constexpr int TS = 64 * (DH + 4);
  float* Qs = smem; float* dOs = Qs + TS; float* Ks = dOs + TS; float* Vs = Ks + TS;
  float* Ps = Vs + TS; float* dSs = Ps + 64 * PS;
  float* kb = dSs + 64 * PS; float* lse_s = kb + 64; float* delta_s = lse_s + 64;
  constexpr int W = DH / 16;
  const int q0 = blockIdx.x * BM, h = blockIdx.y, b = blockIdx.z;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* qp = A.q + (int64_t)b * A.Lq * A.ldq + h * DH;
  const float* dop = A.d_o + (int64_t)b * A.Lq * A.ldo + h * DH;
  const float* kp = A.k + (int64_t)b * A.Lk * A.ldk + h * DH;
  const float* vp = A.v + (int64_t)b * A.Lk * A.ldv + h * DH;
  DropCtx D{A.seed, A.offset, drop_threshold16(A.p_drop), A.p_drop > 0.f ? 1.f / (1.f - A.p_drop) : 1.f, A.p_drop > 0.f};
  const int Lk4 = (A.Lk + 31) / 32;    // keep words (32 keys) per row
  load_tile<DH, false>(Qs, qp, A.ldq, q0, A.Lq);
  load_tile<DH, false>(dOs, dop, A.ldo, q0, A.Lq);
  if (threadIdx.x < BM) {
    int qi = q0 + threadIdx.x;
    lse_s[threadIdx.x] = qi < A.Lq ? A.lse[((int64_t)b * A.H + h) * A.Lq + qi] : -INFINITY;
    delta_s[threadIdx.x] = qi < A.Lq ? A.delta[((int64_t)b * A.H + h) * A.Lq + qi] : 0.f;
  }
  float dq[4][W];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int w = 0; w < W; ++w) dq[r][w] = 0.f;
  const int kv_end = A.causal ? min(A.Lk, q0 + BM) : A.Lk;
  for (int k0 = 0; k0 < kv_end; k0 += BN) {
    __syncthreads();
    load_tile<DH, true>(Ks, kp, A.ldk, k0, A.Lk);
    load_tile<DH, true>(Vs, vp, A.ldv, k0, A.Lk);
    if (threadIdx.x < BN) {
      int kj = k0 + threadIdx.x;
      bool ok = kj < A.Lk && !(A.kpm != nullptr && A.kpm[(int64_t)b * A.Lk + kj]);
      kb[threadIdx.x] = ok ? 0.f : -INFINITY;
    }
    __syncthreads();
    recompute_p_ds<DH>(A, D, Qs, dOs, Ks, Vs, kb, lse_s, delta_s, Ps, dSs, b, h, q0, k0, ty, tx, Lk4);
    __syncthreads();
    tile_pv<DH, true>(dSs, Ks, ty, tx, dq);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int qi = q0 + ty * 4 + r;
    if (qi >= A.Lq) continue;
    float* p = A.dq + ((int64_t)b * A.Lq + qi) * A.lddq + h * DH + tx * W;
#pragma unroll
    for (int w = 0; w < W; ++w) p[w] = A.round_out ? tf32_rn(dq[r][w] * A.scale) : dq[r][w] * A.scale;
  }
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) { pa_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PA_ERR_CUDA; }
  return PA_OK;
}

template <int DH>
int launch_fwd(const pa_attn_fwd_args& A, cudaStream_t st) {
  size_t smem = (size_t)(3 * 64 * (DH + 4) + 64 * PS + 64) * sizeof(float);
  int rc = set_smem(attn_fwd_simt_kernel<DH>, smem);
  if (rc) return rc;
  dim3 grid((A.Lq + BM - 1) / BM, A.H, A.B);
  attn_fwd_simt_kernel<DH><<<grid, NT, smem, st>>>(A);
  PA_CHECK_LAUNCH();
  return PA_OK;
}

template <int DH>
int launch_bwd(const pa_attn_bwd_args& A, cudaStream_t st) {
  size_t smem = (size_t)(4 * 64 * (DH + 4) + 2 * 64 * PS + 3 * 64) * sizeof(float);
  int rc = set_smem(attn_bwd_dkdv_simt_kernel<DH>, smem);
  if (rc) return rc;
  rc = set_smem(attn_bwd_dq_simt_kernel<DH>, smem);
  if (rc) return rc;
  int64_t n = (int64_t)A.B * A.Lq * A.H;
  attn_delta_kernel<<<(unsigned)((4 * n + 255) / 256), 256, 0, st>>>(A.o, A.d_o, A.ldo, A.B, A.H, A.Lq, A.dh, A.delta);
  PA_CHECK_LAUNCH();
  dim3 gk((A.Lk + BN - 1) / BN, A.H, A.B);
  attn_bwd_dkdv_simt_kernel<DH><<<gk, NT, smem, st>>>(A);
  PA_CHECK_LAUNCH();
  dim3 gq((A.Lq + BM - 1) / BM, A.H, A.B);
  attn_bwd_dq_simt_kernel<DH><<<gq, NT, smem, st>>>(A);
  PA_CHECK_LAUNCH();
  return PA_OK;
}

}  // namespace

int pa_attn_fwd_tc(const pa_attn_fwd_args* a, void* stream);  // attn_tc.cu
int pa_attn_bwd_tc(const pa_attn_bwd_args* a, void* stream);  // attn_bwd_tc.cu

int pa_attn_delta_launch(const float* o, const float* d_o, int64_t ldo, int B, int H, int Lq, int dh, float* delta, cudaStream_t st) {
  int64_t n = (int64_t)B * Lq * H;
  attn_delta_kernel<<<(unsigned)((4 * n + 255) / 256), 256, 0, st>>>(o, d_o, ldo, B, H, Lq, dh, delta);
  PA_CHECK_LAUNCH();
  return PA_OK;
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

extern "C" int pa_attn_fwd(const pa_attn_fwd_args* a, void* stream) {
  PA_CHECK_ARG(a != nullptr && a->B > 0 && a->H > 0 && a->Lq > 0 && a->Lk > 0);
  PA_CHECK_ARG(a->ldq % 4 == 0 && a->ldk % 4 == 0 && a->ldv % 4 == 0 && a->ldo % 4 == 0);
  PA_CHECK_ARG(aligned16(a->q) && aligned16(a->k) && aligned16(a->v) && aligned16(a->o));
  PA_CHECK_ARG(a->p_drop >= 0.f && a->p_drop < 1.f);
  if (a->impl == 1) return pa_attn_fwd_tc(a, stream);
  PA_CHECK_ARG(a->impl == 0);
  switch (a->dh) {
    case 32: return launch_fwd<32>(*a, (cudaStream_t)stream);
    case 64: return launch_fwd<64>(*a, (cudaStream_t)stream);
    default: pa_set_error("pa_attn_fwd: head dim %d unsupported (32, 64)", a->dh); return PA_ERR_UNSUPPORTED;
  }
}

extern "C" int pa_attn_bwd(const pa_attn_bwd_args* a, void* stream) {
  PA_CHECK_ARG(a != nullptr && a->B > 0 && a->H > 0 && a->Lq > 0 && a->Lk > 0 && a->delta != nullptr && a->lse != nullptr);
  PA_CHECK_ARG(a->ldq % 4 == 0 && a->ldk % 4 == 0 && a->ldv % 4 == 0 && a->ldo % 4 == 0);
  PA_CHECK_ARG(a->lddq % 4 == 0 && a->lddk % 4 == 0 && a->lddv % 4 == 0);
  PA_CHECK_ARG(aligned16(a->q) && aligned16(a->k) && aligned16(a->v) && aligned16(a->o) && aligned16(a->d_o));
  if (a->impl == 1) return pa_attn_bwd_tc(a, stream);
  PA_CHECK_ARG(a->impl == 0);
  switch (a->dh) {
    case 32: return launch_bwd<32>(*a, (cudaStream_t)stream);
    case 64: return launch_bwd<64>(*a, (cudaStream_t)stream);
    default: pa_set_error("pa_attn_bwd: head dim %d unsupported (32, 64)", a->dh); return PA_ERR_UNSUPPORTED;
  }
}
