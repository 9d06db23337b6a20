// K9/K10/K11: pointer/attach head epilogues.
//   train: fused distribution + NLL + argmax + accuracy counters, forward and backward
//          (ref models.py:156-166, 219-227) -- the [B,T,V+T] distribution is never materialised.
//   eval : per-step distribution, argmax, pointer resolution, END bookkeeping
//          (ref models.py:168-186, 235-256, 306).
// HBM-bound row kernels: one warp per (b,t) row in training (reads V+T logits, writes 4 stats),
// one block per sequence in decode.
#include "common.cuh"

namespace {

constexpr float kEps = 1e-6f;
constexpr int kRowsPerBlock = 8;

struct ArgMax {
  float v;
  int i;
};
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {  // larger value; ties -> lower index
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ ArgMax warp_argmax(ArgMax a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax b{__shfl_xor_sync(0xffffffffu, a.v, o), __shfl_xor_sync(0xffffffffu, a.i, o)};
    a = better(a, b);
  }
  return a;
}

// pointer score of column j for row i after the training-time fill (ref models.py:160-161)
__device__ __forceinline__ float ptr_train(const float* __restrict__ lp_row, int i, int j, float inv_d) {
  return j < i ? __ldg(lp_row + j) * inv_d : kEps;
}

__global__ void __launch_bounds__(kRowsPerBlock * 32) dist_loss_fwd_kernel(const float* __restrict__ lv, const float* __restrict__ lp,
                                                                             const float* __restrict__ sw, const int64_t* __restrict__ label,
                                                                             int N, int T, int V, int pad, float inv_d,
                                                                             float4* __restrict__ rowstat, int64_t* __restrict__ predict,
                                                                             float* __restrict__ accum) {
  __shared__ float s_part[kRowsPerBlock][3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * kRowsPerBlock + warp;
  float loss = 0.f, valid = 0.f, correct = 0.f;
  if (n < N) {
    const int i = n % T;
    const float* lv_row = lv + (int64_t)n * V;
    const float* lp_row = lp + (int64_t)n * T;
    ArgMax av{-INFINITY, 0x7fffffff}, ap{-INFINITY, 0x7fffffff};
    for (int c = lane; c < V; c += 32) av = better(av, ArgMax{__ldg(lv_row + c), c});
    for (int j = lane; j < T; j += 32) ap = better(ap, ArgMax{ptr_train(lp_row, i, j, inv_d), j});
    av = warp_argmax(av);
    ap = warp_argmax(ap);
    float sv = 0.f, sp = 0.f;
    for (int c = lane; c < V; c += 32) sv += __expf(__ldg(lv_row + c) - av.v);
    for (int j = lane; j < T; j += 32) sp += __expf(ptr_train(lp_row, i, j, inv_d) - ap.v);
    const float lse_v = av.v + __logf(warp_sum(sv));
    const float lse_p = ap.v + __logf(warp_sum(sp));
    const float pi = 1.f / (1.f + __expf(-__ldg(sw + n)));
    const float log_1mpi = __logf(fmaxf(1.f - pi, kEps)), log_pi = __logf(fmaxf(pi, kEps));
    const float best_v = av.v - lse_v + log_1mpi, best_p = ap.v - lse_p + log_pi;
    const int64_t pred = best_v >= best_p ? (int64_t)av.i : (int64_t)(V + ap.i);
    const int64_t lab = label[n];
    float logp = 0.f;
    if (lab != pad) {
      logp = lab < V ? __ldg(lv_row + lab) - lse_v + log_1mpi : ptr_train(lp_row, i, (int)(lab - V), inv_d) - lse_p + log_pi;
      loss = -logp; valid = 1.f; correct = pred == lab ? 1.f : 0.f;
    }
    if (lane == 0) {
      rowstat[n] = make_float4(lse_v, lse_p, pi, logp);
      predict[n] = pred;
    }
  }
  if (lane == 0) { s_part[warp][0] = loss; s_part[warp][1] = valid; s_part[warp][2] = correct; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kRowsPerBlock; ++w) acc += s_part[w][threadIdx.x];
    if (acc != 0.f) atomicAdd(accum + threadIdx.x, acc);
  }
}

__global__ void __launch_bounds__(kRowsPerBlock * 32) dist_loss_bwd_kernel(const float* __restrict__ lv, const float* __restrict__ lp,
                                                                             const float* __restrict__ sw, const int64_t* __restrict__ label,
                                                                             const float4* __restrict__ rowstat, const float* __restrict__ accum,
                                                                             const float* __restrict__ gout, int N, int T, int V, int pad,
                                                                             float inv_d, float* __restrict__ dlv, float* __restrict__ dlp,
                                                                             float* __restrict__ dsw, int rnd) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * kRowsPerBlock + warp;
  if (n >= N) return;
  const int i = n % T;
  const int64_t lab = label[n];
  float* dlv_row = dlv + (int64_t)n * V;
  float* dlp_row = dlp + (int64_t)n * T;
  if (lab == pad) {
    for (int c = lane; c < V; c += 32) dlv_row[c] = 0.f;
    for (int j = lane; j < T; j += 32) dlp_row[j] = 0.f;
    if (lane == 0) dsw[n] = 0.f;
    return;
  }
  const float scale = __ldg(gout) / __ldg(accum + 1);   // d loss / d(-logp of this row)
  const float4 st = rowstat[n];
  const float pi = st.z;
  const float* lv_row = lv + (int64_t)n * V;
  const float* lp_row = lp + (int64_t)n * T;
  float dpi;
  if (lab < V) {
    for (int c = lane; c < V; c += 32) {
      float p = __expf(__ldg(lv_row + c) - st.x);
      float gv = scale * (p - (c == lab ? 1.f : 0.f));
      dlv_row[c] = rnd ? tf32_rn(gv) : gv;
    }
    for (int j = lane; j < T; j += 32) dlp_row[j] = 0.f;
    dpi = (1.f - pi) > kEps ? scale / (1.f - pi) : 0.f;
  } else {
    const int jl = (int)(lab - V);
    for (int c = lane; c < V; c += 32) dlv_row[c] = 0.f;
    for (int j = lane; j < T; j += 32) {
      float g = 0.f;
      if (j < i) g = scale * (__expf(__ldg(lp_row + j) * inv_d - st.y) - (j == jl ? 1.f : 0.f)) * inv_d;
      dlp_row[j] = rnd ? tf32_rn(g) : g;
    }
    dpi = pi > kEps ? -scale / pi : 0.f;
  }
  if (lane == 0) dsw[n] = dpi * pi * (1.f - pi);
}

__global__ void __launch_bounds__(kRowsPerBlock * 32) dist_train_full_kernel(const float* __restrict__ lv, const float* __restrict__ lp,
                                                                               const float* __restrict__ sw, int N, int T, int V, float inv_d,
                                                                               float* __restrict__ dists) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * kRowsPerBlock + warp;
  if (n >= N) return;
  const int i = n % T;
  const float* lv_row = lv + (int64_t)n * V;
  const float* lp_row = lp + (int64_t)n * T;
  float mv = -INFINITY, mp = -INFINITY;
  for (int c = lane; c < V; c += 32) mv = fmaxf(mv, __ldg(lv_row + c));
  for (int j = lane; j < T; j += 32) mp = fmaxf(mp, ptr_train(lp_row, i, j, inv_d));
  mv = warp_max(mv); mp = warp_max(mp);
  float sv = 0.f, sp = 0.f;
  for (int c = lane; c < V; c += 32) sv += __expf(__ldg(lv_row + c) - mv);
  for (int j = lane; j < T; j += 32) sp += __expf(ptr_train(lp_row, i, j, inv_d) - mp);
  const float lse_v = mv + __logf(warp_sum(sv)), lse_p = mp + __logf(warp_sum(sp));
  const float pi = 1.f / (1.f + __expf(-__ldg(sw + n)));
  const float a = __logf(fmaxf(1.f - pi, kEps)), b = __logf(fmaxf(pi, kEps));
  float* out = dists + (int64_t)n * (V + T);
  for (int c = lane; c < V; c += 32) out[c] = __ldg(lv_row + c) - lse_v + a;
  for (int j = lane; j < T; j += 32) out[V + j] = ptr_train(lp_row, i, j, inv_d) - lse_p + b;
}

// ------------------------------------------------------------------------------------ decode head
constexpr int kHeadThreads = 256;

__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < kHeadThreads / 32; ++w) r = fmaxf(r, red[w]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < kHeadThreads / 32; ++w) r += red[w];
  __syncthreads();
  return r;
}
__device__ __forceinline__ ArgMax block_argmax(ArgMax a, float* redv, int* redi) {
  a = warp_argmax(a);
  if ((threadIdx.x & 31) == 0) { redv[threadIdx.x >> 5] = a.v; redi[threadIdx.x >> 5] = a.i; }
  __syncthreads();
  ArgMax r{redv[0], redi[0]};
#pragma unroll
  for (int w = 1; w < kHeadThreads / 32; ++w) r = better(r, ArgMax{redv[w], redi[w]});
  __syncthreads();
  return r;
}

// 6-periodic admissibility of pointer column j for the row i (ref models.py:91-101); j < i assumed
__device__ __forceinline__ bool ptr_allowed(int i, int j) {
  if (i < 6) return false;
  return j < 6 ? (j == i % 6) : (j % 6 == (i % 6 + 3) % 6);
}

__global__ void __launch_bounds__(kHeadThreads) decode_head_kernel(const float* __restrict__ h, const float* __restrict__ lv,
                                                                     const float* __restrict__ pf, const float* __restrict__ sw,
                                                                     float* __restrict__ hfin, int64_t Tmax, int d, int V, int t_host,
                                                                     const int* __restrict__ t_dev, int end_token, int64_t* __restrict__ samples, int64_t* __restrict__ attach,
                                                                     int64_t ld, int32_t* __restrict__ first_end) {
  extern __shared__ float s_lp[];  // [t+1] pointer scores / probs
  __shared__ float redv[kHeadThreads / 32];
  __shared__ int redi[kHeadThreads / 32];
  const int t = t_dev != nullptr ? *t_dev : t_host;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* hb = h + (int64_t)b * d;
  float* hf = hfin + (int64_t)b * Tmax * d;
  for (int c = tid; c < d; c += kHeadThreads) hf[(int64_t)t * d + c] = hb[c];

  const float* lvb = lv + (int64_t)b * V;
  float mx = -INFINITY;
  for (int c = tid; c < V; c += kHeadThreads) mx = fmaxf(mx, lvb[c]);
  mx = block_max(mx, redv);
  float sum = 0.f;
  for (int c = tid; c < V; c += kHeadThreads) sum += expf(lvb[c] - mx);
  sum = block_sum(sum, redv);
  const int sz = t + 1;
  const float pi = sz < 6 ? 0.f : 1.f / (1.f + expf(-sw[b]));
  ArgMax best{-INFINITY, 0x7fffffff};
  for (int c = tid; c < V; c += kHeadThreads) {
    float p = expf(lvb[c] - mx) / sum;
    if (sz >= 6) p *= (1.f - pi);
    best = better(best, ArgMax{p, c});
  }
  if (sz >= 6) {
    // scores against the cached final hiddens of the earlier positions j < t
    const float inv_d = 1.f / (float)d;
    const float* pfb = pf + (int64_t)b * d;
    for (int j = warp; j < t; j += kHeadThreads / 32) {
      float acc = 0.f;
      for (int c = lane * 4; c < d; c += 128) {
        float4 a = *reinterpret_cast<const float4*>(pfb + c);
        float4 k = *reinterpret_cast<const float4*>(hf + (int64_t)j * d + c);
        acc += a.x * k.x + a.y * k.y + a.z * k.z + a.w * k.w;
      }
      acc = warp_sum(acc);
      if (lane == 0) s_lp[j] = acc * inv_d;
    }
    __syncthreads();
    float pm = -INFINITY;
    for (int j = tid; j < t; j += kHeadThreads) pm = fmaxf(pm, s_lp[j]);
    pm = block_max(pm, redv);
    float ps = 0.f;
    for (int j = tid; j < t; j += kHeadThreads) ps += expf(s_lp[j] - pm);
    ps = block_sum(ps, redv);
    const int i = sz - 1;
    for (int j = tid; j < sz; j += kHeadThreads) {
      float p = kEps;
      if (j < t && ptr_allowed(i, j)) p = expf(s_lp[j] - pm) / ps * pi;
      best = better(best, ArgMax{p, V + j});
    }
  }
  best = block_argmax(best, redv, redi);
  if (tid == 0) {
    int64_t tok = best.i, att = -1;
    if (best.i >= V) { att = best.i - V; tok = samples[(int64_t)b * ld + att]; }
    samples[(int64_t)b * ld + t] = tok;
    attach[(int64_t)b * ld + t] = att;
    if (tok == end_token && first_end[b] > t) first_end[b] = t;
  }
}

}  // namespace

extern "C" int pa_dist_loss_fwd(const float* lv, const float* lp, const float* sw, const int64_t* label, int B, int T,
                                int V, int pad, float inv_d, float* rowstat, int64_t* predict, float* accum,
                                void* stream) {
  PA_CHECK_ARG(B > 0 && T > 0 && V > 0);
  int N = B * T;
  dist_loss_fwd_kernel<<<(N + kRowsPerBlock - 1) / kRowsPerBlock, kRowsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      lv, lp, sw, label, N, T, V, pad, inv_d, (float4*)rowstat, predict, accum);
  PA_CHECK_LAUNCH();
  return PA_OK;
}

extern "C" int pa_dist_loss_bwd(const float* lv, const float* lp, const float* sw, const int64_t* label,
                                const float* rowstat, const float* accum, const float* gout, int B, int T, int V,
                                int pad, float inv_d, float* dlv, float* dlp, float* dsw, int round_tf32, void* stream) {
  PA_CHECK_ARG(B > 0 && T > 0 && V > 0);
  int N = B * T;
  dist_loss_bwd_kernel<<<(N + kRowsPerBlock - 1) / kRowsPerBlock, kRowsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      lv, lp, sw, label, (const float4*)rowstat, accum, gout, N, T, V, pad, inv_d, dlv, dlp, dsw, round_tf32);
  PA_CHECK_LAUNCH();
  return PA_OK;
}

extern "C" int pa_dist_train_full(const float* lv, const float* lp, const float* sw, int B, int T, int V, float inv_d,
                                  float* dists, void* stream) {
  PA_CHECK_ARG(B > 0 && T > 0 && V > 0);
  int N = B * T;
  dist_train_full_kernel<<<(N + kRowsPerBlock - 1) / kRowsPerBlock, kRowsPerBlock * 32, 0, (cudaStream_t)stream>>>(lv, lp, sw, N, T, V, inv_d, dists);
  PA_CHECK_LAUNCH();
  return PA_OK;
}

extern "C" int pa_decode_head(const float* h, const float* lv, const float* pf, const float* sw, float* hfin,
                              int64_t Tmax, int B, int d, int V, int t, const int* t_dev, int end_token, int64_t* samples,
                              int64_t* attach, int64_t ld, int32_t* first_end, void* stream) {
  PA_CHECK_ARG(B > 0 && d % 4 == 0 && t >= 0 && t < Tmax);
  decode_head_kernel<<<B, kHeadThreads, (size_t)(t_dev != nullptr ? Tmax : t + 1) * sizeof(float), (cudaStream_t)stream>>>(
      h, lv, pf, sw, hfin, Tmax, d, V, t, t_dev, end_token, samples, attach, ld, first_end);
  PA_CHECK_LAUNCH();
  return PA_OK;
}
