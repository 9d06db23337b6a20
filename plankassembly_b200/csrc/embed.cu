// K1/K2: fused embedding gather-sum kernels (HBM-bound; 16-byte vector loads/stores, one pass).
//   K1 replaces ref models.py:103-112 (5 gathers + 4 adds = 9 launches, 5 HBM round trips).
//   K2 replaces ref models.py:114-138 (gather x3, arange/remainder/div, adds, zeros, concat).
// Algorithmic bytes per token: n_tables*8 B ids + d*4 B written (table rows are L2-resident).
#include "common.cuh"

struct EmbedTables {
  const int64_t* ids[PA_MAX_TABLES];
  const float* tab[PA_MAX_TABLES];
  float* dtab[PA_MAX_TABLES];
  int rows[PA_MAX_TABLES];
  int n;
};

__global__ void __launch_bounds__(256) embed_input_fwd_kernel(EmbedTables T, int64_t n_tokens, int d4, float4* __restrict__ out, float4* __restrict__ out_r) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_tokens * d4) return;
  int64_t tok = idx / d4;
  int c = (int)(idx - tok * d4);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < PA_MAX_TABLES; ++k) {
    if (k < T.n) {
      int64_t id = __ldg(T.ids[k] + tok);
      float4 v = __ldg(reinterpret_cast<const float4*>(T.tab[k]) + id * d4 + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  out[idx] = acc;
  if (out_r != nullptr) out_r[idx] = tf32_rn4(acc);
}

// Backward: one block owns a chunk of tokens; thread c owns float4 column c.  Large tables take
// one vector RED per token; tables with <= 8 rows (coord/view/type) are reduced in registers
// first so the few hot rows do not serialise in L2.
constexpr int kBwdChunk = 64;

__device__ __forceinline__ void red_add4(float* p, float4 v) {
  atomicAdd(reinterpret_cast<float4*>(p), v);
}

__global__ void __launch_bounds__(128) embed_input_bwd_kernel(EmbedTables T, int64_t n_tokens, int d4, const float4* __restrict__ dout) {
  __shared__ int s_ids[PA_MAX_TABLES][kBwdChunk];
  int64_t tok0 = (int64_t)blockIdx.x * kBwdChunk;
  int n_here = (int)min((int64_t)kBwdChunk, n_tokens - tok0);
  for (int i = threadIdx.x; i < T.n * kBwdChunk; i += blockDim.x) {
    int k = i / kBwdChunk, j = i % kBwdChunk;
    s_ids[k][j] = j < n_here ? (int)T.ids[k][tok0 + j] : -1;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d4; c += blockDim.x) {
    float4 small[PA_MAX_TABLES];  // reused per small table below
    for (int k = 0; k < T.n; ++k) {
      if (T.rows[k] > 8) {
        for (int j = 0; j < n_here; ++j) {
          float4 g = __ldg(dout + (tok0 + j) * d4 + c);
          red_add4(T.dtab[k] + ((int64_t)s_ids[k][j] * d4 + c) * 4, g);
        }
      } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) small[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < n_here; ++j) {
          float4 g = __ldg(dout + (tok0 + j) * d4 + c);
          int id = s_ids[k][j];
#pragma unroll
          for (int r = 0; r < 8; ++r)
            if (id == r) { small[r].x += g.x; small[r].y += g.y; small[r].z += g.z; small[r].w += g.w; }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
          if (r < T.rows[k]) red_add4(T.dtab[k] + ((int64_t)r * d4 + c) * 4, small[r]);
      }
    }
  }
}

__global__ void __launch_bounds__(256) embed_output_fwd_kernel(const int64_t* __restrict__ value, int64_t ld, int B, int Tn, int dof,
                                                                  const float4* __restrict__ e_val, const float4* __restrict__ e_coord,
                                                                  const float4* __restrict__ e_pos, int d4, float4* __restrict__ out, float4* __restrict__ out_r) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * Tn * d4) return;
  int c = (int)(idx % d4);
  int64_t row = idx / d4;
  int t = (int)(row % Tn);
  int b = (int)(row / Tn);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t > 0) {
    int64_t id = __ldg(value + (int64_t)b * ld + (t - 1));
    float4 v = __ldg(e_val + id * d4 + c);
    float4 q = __ldg(e_coord + (int64_t)((t - 1) % dof) * d4 + c);
    float4 p = __ldg(e_pos + (int64_t)((t - 1) / dof) * d4 + c);
    // same association order as the reference: (value + coord) + pos
    acc.x = (v.x + q.x) + p.x; acc.y = (v.y + q.y) + p.y; acc.z = (v.z + q.z) + p.z; acc.w = (v.w + q.w) + p.w;
  }
  out[idx] = acc;
  if (out_r != nullptr) out_r[idx] = tf32_rn4(acc);
}

// Backward: block = one decoder position t (> 0) x a slice of the batch.  coord/pos rows depend
// on t only, so they are reduced over the batch slice in registers and hit memory once.
__global__ void __launch_bounds__(128) embed_output_bwd_kernel(const float4* __restrict__ dout, const int64_t* __restrict__ value, int64_t ld,
                                                                  int B, int Tn, int dof, float* d_val, float* d_coord, float* d_pos, int d4) {
  int t = blockIdx.x + 1;
  int b0 = blockIdx.y * 16, b1 = min(B, b0 + 16);
  for (int c = threadIdx.x; c < d4; c += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = b0; b < b1; ++b) {
      float4 g = __ldg(dout + ((int64_t)b * Tn + t) * d4 + c);
      int64_t id = __ldg(value + (int64_t)b * ld + (t - 1));
      red_add4(d_val + (id * d4 + c) * 4, g);
      acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
    }
    red_add4(d_coord + ((int64_t)((t - 1) % dof) * d4 + c) * 4, acc);
    red_add4(d_pos + ((int64_t)((t - 1) / dof) * d4 + c) * 4, acc);
  }
}

__global__ void __launch_bounds__(128) decode_embed_kernel(const int64_t* __restrict__ samples, int64_t ld, int B, int t_host, const int* __restrict__ t_dev, int dof,
                                                              const float4* __restrict__ e_val, const float4* __restrict__ e_coord,
                                                              const float4* __restrict__ e_pos, int d4, float4* __restrict__ y) {
  const int t = t_dev != nullptr ? *t_dev : t_host;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * d4) return;
  int c = idx % d4, b = idx / d4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t > 0) {
    int64_t id = __ldg(samples + (int64_t)b * ld + (t - 1));
    float4 v = __ldg(e_val + id * d4 + c);
    float4 q = __ldg(e_coord + (int64_t)((t - 1) % dof) * d4 + c);
    float4 p = __ldg(e_pos + (int64_t)((t - 1) / dof) * d4 + c);
    acc.x = (v.x + q.x) + p.x; acc.y = (v.y + q.y) + p.y; acc.z = (v.z + q.z) + p.z; acc.w = (v.w + q.w) + p.w;
  }
  y[idx] = acc;
}

extern "C" int pa_embed_input_fwd(const int64_t* const* ids_host, const float* const* tables_host, int n_tables,
                                  int64_t n_tokens, int d, float* out, float* out_tf32, void* stream) {
  PA_CHECK_ARG(n_tables >= 1 && n_tables <= PA_MAX_TABLES && d % 4 == 0 && n_tokens >= 0);
  if (n_tokens == 0) return PA_OK;
  EmbedTables T{};
  T.n = n_tables;
  for (int k = 0; k < n_tables; ++k) { T.ids[k] = ids_host[k]; T.tab[k] = tables_host[k]; }
  int64_t total = n_tokens * (d / 4);
  embed_input_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(T, n_tokens, d / 4, (float4*)out, (float4*)out_tf32);
  PA_CHECK_LAUNCH();
  return PA_OK;
}

extern "C" int pa_embed_input_bwd(const float* dout, const int64_t* const* ids_host, float* const* dtables_host,
                                  const int* table_rows_host, int n_tables, int64_t n_tokens, int d, void* stream) {
  PA_CHECK_ARG(n_tables >= 1 && n_tables <= PA_MAX_TABLES && d % 4 == 0 && n_tokens >= 0);
  if (n_tokens == 0) return PA_OK;
  EmbedTables T{};
  T.n = n_tables;
  for (int k = 0; k < n_tables; ++k) { T.ids[k] = ids_host[k]; T.dtab[k] = dtables_host[k]; T.rows[k] = table_rows_host[k]; }
  embed_input_bwd_kernel<<<(unsigned)((n_tokens + kBwdChunk - 1) / kBwdChunk), 128, 0, (cudaStream_t)stream>>>(T, n_tokens, d / 4, (const float4*)dout);
  PA_CHECK_LAUNCH();
  return PA_OK;
}

extern "C" int pa_embed_output_fwd(const int64_t* value, int64_t ld, int B, int T, int dof, const float* e_val,
                                   const float* e_coord, const float* e_pos, int d, float* out, float* out_tf32, void* stream) {
  PA_CHECK_ARG(B > 0 && T > 0 && dof > 0 && d % 4 == 0);
  int64_t total = (int64_t)B * T * (d / 4);
  embed_output_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      value, ld, B, T, dof, (const float4*)e_val, (const float4*)e_coord, (const float4*)e_pos, d / 4, (float4*)out, (float4*)out_tf32);
  PA_CHECK_LAUNCH();
  return PA_OK;
}

extern "C" int pa_embed_output_bwd(const float* dout, const int64_t* value, int64_t ld, int B, int T, int dof,
                                   float* d_val, float* d_coord, float* d_pos, int d, void* stream) {
  PA_CHECK_ARG(B > 0 && T > 0 && dof > 0 && d % 4 == 0);
  if (T == 1) return PA_OK;
  dim3 grid(T - 1, (B + 15) / 16);
  embed_output_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const float4*)dout, value, ld, B, T, dof, d_val, d_coord, d_pos, d / 4);
  PA_CHECK_LAUNCH();
  return PA_OK;
}

extern "C" int pa_decode_embed(const int64_t* samples, int64_t ld, int B, int t, const int* t_dev, int dof, const float* e_val,
                               const float* e_coord, const float* e_pos, int d, float* y, void* stream) {
  PA_CHECK_ARG(B > 0 && t >= 0 && d % 4 == 0);
  int total = B * (d / 4);
  decode_embed_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(samples, ld, B, t, t_dev, dof, (const float4*)e_val,
                                                                           (const float4*)e_coord, (const float4*)e_pos, d / 4, (float4*)y);
  PA_CHECK_LAUNCH();
  return PA_OK;
}
