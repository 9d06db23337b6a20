/* plank_b200 C ABI -- the drop-in boundary of the B200 hot path.
 *
 * The reference (manycore-research/PlankAssembly) has no FFI: its hot path is the Python
 * surface of plankassembly/models.py and every FLOP is a PyTorch library call.  Each entry
 * point below replaces the library op sequence issued at the cited reference line(s); the
 * host-side mirror (plankassembly_b200/models.py) binds them through ctypes.
 *
 * Conventions: every pointer is a DEVICE pointer unless the name ends in _host; the caller
 * owns all buffers (kernels never allocate); tensors are row-major with a contiguous last
 * dimension; ids/labels are int64 exactly as the reference's batches hold them; masks are
 * uint8 (1 = PAD key).  `stream` is a cudaStream_t.  Return 0 on success, a negative
 * pa_status otherwise; pa_last_error() returns a thread-local message.  No hidden syncs.
 */
#ifndef PLANK_B200_H
#define PLANK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PA_ABI_VERSION 3
#define PA_MAX_TABLES 8

enum pa_status { PA_OK = 0, PA_ERR_ARG = -1, PA_ERR_CUDA = -2, PA_ERR_UNSUPPORTED = -3 };

int pa_abi_version(void);
const char* pa_last_error(void);
/* 1 if the current device is compute capability 10.x (the only one this library targets). */
int pa_device_ok(void);

/* ---- K1: fused input-embedding gather-sum.  Replaces models.py:103-112 (5 nn.Embedding
 * gathers + 4 adds).  out[n,:] = sum_k tables[k][ids[k][n], :].  bwd accumulates (+=) into
 * dtables[k] (the value table is shared with the decoder side, models.py:120). */
int pa_embed_input_fwd(const int64_t* const* ids_host, const float* const* tables_host, int n_tables,
                       int64_t n_tokens, int d, float* out, float* out_tf32, void* stream);
int pa_embed_input_bwd(const float* dout, const int64_t* const* ids_host, float* const* dtables_host,
                       const int* table_rows_host, int n_tables, int64_t n_tokens, int d, void* stream);

/* ---- K2: decoder-input embedding, shifted right by one with a zero row (models.py:114-138).
 * out[b,0,:] = 0; out[b,t,:] = e_val[value[b*ld+t-1]] + e_coord[(t-1)%dof] + e_pos[(t-1)/dof]. */
int pa_embed_output_fwd(const int64_t* value, int64_t ld, int B, int T, int dof, const float* e_val,
                        const float* e_coord, const float* e_pos, int d, float* out, float* out_tf32, void* stream);
int pa_embed_output_bwd(const float* dout, const int64_t* value, int64_t ld, int B, int T, int dof,
                        float* d_val, float* d_coord, float* d_pos, int d, void* stream);

/* TF32 policy: tcgen05 kind::tf32 truncates raw fp32 operands (a systematic ~2^-11 bias per product, which
 * breaks the 1e-3 parity bar), so every tensor that feeds a tensor-core operand is ROUNDED TO NEAREST TF32
 * by the kernel that produces it: `*_tf32` outputs are rounded copies, `round_*` flags round in place. */

/* ---- K6/K7: y = LayerNorm_eps(x + dropout_p(a)) -- the post-norm residual blocks of
 * torch nn/modules/transformer.py (encoder :952-956, decoder :1144-1153) as configured by
 * models.py:60-69 (eps = 1.0 in layers, 1e-5 in the two final norms).  a may be NULL
 * (final norms).  a_bias (may be NULL) is the bias of the linear that produced `a`, folded in here so
 * that its gradient (column sums of da -> d_a_bias, accumulated +=) comes out of the backward kernel for free.
 * s receives the pre-norm sum (saved for bwd; may alias nothing, may be NULL
 * in inference); stats = [rows,2] (mean, rstd), may be NULL in inference. */
int pa_add_ln_fwd(const float* x, const float* a, const float* a_bias, const float* gamma, const float* beta, float eps,
                  float p_drop, uint64_t seed, uint64_t offset, int64_t rows, int d, float* y, float* y_tf32, float* s,
                  float* stats, void* stream);
/* dx = grad wrt x (and wrt s); da = dropout-masked copy (may be NULL); dgamma/dbeta are
 * accumulated (+=, vector RED per block: 16-byte aligned).  dy2 (may be NULL) is a second incoming gradient added to dy (the
 * TF32 copy's).  partial: unused since ABI 3 (pass NULL; pa_add_ln_bwd_workspace returns 0).  beta == NULL: `s` is the pre-norm sum the forward saved; beta != NULL: `s` is the
 * layer's OUTPUT y (or its TF32 copy, which the next GEMM keeps alive anyway) and x_hat is recovered as (y - beta) / gamma,
 * so the forward need not write a pre-norm tensor at all (20 % less LayerNorm traffic, one [rows,d] tensor less per layer). */
size_t pa_add_ln_bwd_workspace(int64_t rows, int d);
int pa_add_ln_bwd(const float* dy, const float* dy2, const float* s, const float* stats, const float* gamma, const float* beta,
                  float p_drop, uint64_t seed, uint64_t offset, int64_t rows, int d, float* dx, float* da,
                  int round_da, float* dgamma, float* dbeta, float* d_a_bias, void* partial, void* stream);

/* ---- FFN activation: z <- dropout_p(relu(z)) in place (transformer.py _ff_block). */
int pa_relu_dropout_fwd(float* z, int64_t n, float p_drop, uint64_t seed, uint64_t offset, void* stream);
int pa_relu_dropout_bwd(const float* out, float* g, int64_t n, float p_drop, int round_tf32, void* stream);
/* same on a [rows, N] gradient, also accumulating its column sums into dbias[N] (the linear's bias gradient) */
int pa_relu_dropout_bwd_colsum(const float* out, float* g, int64_t rows, int N, float p_drop, int round_tf32,
                               float* dbias, void* stream);
/* dst = round-to-nearest-TF32(src): shadow copies of the weights for the tensor-core path. */
int pa_round_tf32(const float* src, float* dst, int64_t n, void* stream);
/* Error-compensated TF32 operands ("3xTF32", exact-mode inference GEMMs; replaces the cuBLAS fp32 calls torch makes for
 * models.py:279,293): out[rows, 3K] = [hi | lo | hi] of x[rows, K] (row pitch ldx) for activations (weights = 0) or
 * [hi | hi | lo] for weights (weights = 1), hi = rn_tf32(x), lo = rn_tf32(x - hi).  One pa_gemm_tf32 over K' = 3K on such
 * a pair computes hi*hi + lo*hi + hi*lo with FP32 accumulation.  K % 4 == 0, 16-byte aligned pointers. */
int pa_split3_tf32(const float* x, int64_t ldx, float* out, int64_t rows, int K, int weights, void* stream);
/* x <- dropout_p(x) in place (sub-block output dropouts are fused into pa_add_ln_*). */

/* ---- K3/K4/K5: multi-head attention core (torch nn/functional.py multi_head_attention_forward
 * as called from models.py:206,212,279,293): O = dropout(softmax(Q K^T * scale + masks)) V.
 * q[(b*Lq+i)*ldq + h*dh + c] etc., so the packed in-proj output [B,L,3d] is consumed in place.
 * kpm: [B,Lk] uint8, 1 = PAD key (NULL = none); causal: key j visible to query i iff j <= i.
 * lse: [B,H,Lq] natural-log sum-exp of the masked scaled scores (saved for bwd; may be NULL).
 * impl: 0 = fp32 SIMT (exact mode), 1 = tcgen05 TF32 tensor-core path (dh 32 or 64; Lk <= 2048 and, for the backward,
 * Lq <= 1280: per-item bias / statistics tables live in shared memory; longer sequences return PA_ERR_UNSUPPORTED). */
/* Attention-probability dropout keep-masks for the tensor-core kernels, generated ONCE per attention call
 * (all warps of the chip) instead of three times inside the forward / dQ / dK-dV kernels (4 warps per SM):
 * 16 random bits per score: bit = half-word k%2 of word (k%8)/2 of Philox4x32-10(seed, counter = (bh*Lq + q)*ceil(Lk/8)
 * + k/8, offset) >= p*2^16, i.e. exactly the mask the fp32 kernels derive inline.  rows: [BH, Lq, ceil(Lk/32)] words, bit k%32 of word k/32;
 * cols: [BH, 32*ceil(Lk/32), ceil(Lq/32)] words, bit q%32 of word q/32 (the key-stationary backward's view). */
size_t pa_dropout_mask_words(int BH, int Lq, int Lk, int cols);
int pa_dropout_mask(uint32_t* rows, uint32_t* cols, int BH, int Lq, int Lk, float p_drop, uint64_t seed,
                    uint64_t offset, void* stream);

typedef struct {
  const float* q; const float* k; const float* v;
  int64_t ldq, ldk, ldv;
  float* o; int64_t ldo;
  float* lse;
  const uint8_t* kpm;
  int B, H, Lq, Lk, dh;
  int causal;
  float scale;
  float p_drop; uint64_t seed, offset;
  int impl;
  int round_out;               /* write O rounded to TF32 (it feeds a tensor-core GEMM) */
  const uint32_t* drop_rows;   /* keep-bit masks from pa_dropout_mask (required by impl 1 when p_drop > 0) */
  const uint32_t* drop_cols;
  const int32_t* kv_len;       /* optional [B] (both impls, forward): every key j >= kv_len[b] is PAD in kpm -> whole key tiles beyond it are
                                  skipped (ragged batches padded to a fixed length, LineDataset layout); NULL = Lk */
} pa_attn_fwd_args;
int pa_attn_fwd(const pa_attn_fwd_args* args, void* stream);

typedef struct {
  const float* q; const float* k; const float* v;
  int64_t ldq, ldk, ldv;
  const float* o; const float* d_o; int64_t ldo;
  const float* lse;
  float* delta;                 /* workspace [B,H,Lq] */
  float* dq; float* dk; float* dv;
  int64_t lddq, lddk, lddv;
  const uint8_t* kpm;
  int B, H, Lq, Lk, dh;
  int causal;
  float scale;
  float p_drop; uint64_t seed, offset;
  int impl;
  int round_out;               /* write dq/dk/dv rounded to TF32 */
  const uint32_t* drop_rows;   /* the masks the forward used */
  const uint32_t* drop_cols;
  float* dbias;                /* impl 1 only, may be NULL: [3*H*dh] += column sums of (dq | dk | dv) = in-proj bias gradient */
  const int32_t* kv_len;       /* as in pa_attn_fwd_args */
} pa_attn_bwd_args;
int pa_attn_bwd(const pa_attn_bwd_args* args, void* stream);

/* ---- Dense projections on the tcgen05 tensor cores (TF32 in, FP32 accumulate in TMEM):
 *   C[M,N] = alpha * A * B^T (+ bias[N]) (relu) (dropout_p)     -- every nn.Linear on the path
 * (in/out projections, FFN, heads: models.py:60-74,145-153) and, with the MN-major operand forms,
 * their input/weight gradients.  a_mn = 0: A is [M,K] row-major; a_mn = 1: A is stored [K,M].
 * b_mn = 0: B is [N,K] row-major (a torch Linear weight); b_mn = 1: B is stored [K,N].
 * batch > 1: A/B row offsets advance by a_batch_rows/b_batch_rows, C by c_batch_stride elements
 * (pointer scoring bmm, models.py:149).  split_k > 1 requires accumulate = 1 (C += via fp32 RED). */
typedef struct {
  const float* a; int64_t lda; int a_mn;
  const float* b; int64_t ldb; int b_mn;
  float* c; int64_t ldc;
  const float* bias;
  int relu; float p_drop; uint64_t seed, offset;
  float alpha;
  int M, N, K;
  int batch; int64_t a_batch_rows, b_batch_rows, c_batch_stride;
  int split_k; int accumulate;
  int round_out;               /* write C rounded to TF32 (when C only feeds tensor-core operands) */
  /* Fused FFN activation backward (replaces a separate relu/dropout-backward pass over the [tokens, ff] gradient; batch = 1,
   * N % 32 == 0, no split-K).  Forward GEMM (relu [+ dropout] epilogue): mask_out [M, N/32] receives one bit per output,
   * set where C > 0.  Backward dX GEMM: mask_in = that plane, C <- mask ? C * mask_scale : 0 (mask_scale = 1/(1-p)), and
   * colsum [N] (may be NULL) accumulates the column sums of the masked C -- the bias gradient of the forward linear. */
  uint32_t* mask_out; const uint32_t* mask_in; float* colsum; float mask_scale;
} pa_gemm_args;
int pa_gemm_tf32(const pa_gemm_args* args, void* stream);

/* ---- K9/K10: fused training distribution + NLL + argmax (models.py:156-166, 219-227).
 * lv [N,V] vocab logits; lp [N,T] RAW pointer scores pf.h (the kernel applies inv_d and the
 * 1e-6 fill of entries j >= i); sw [N] switch logits; label [N] (N = B*T, row n = b*T+i).
 * Never materialises the [B,T,V+T] distribution.  rowstat [N,4] = (lse_v, lse_p, pi, logp).
 * accum[3] += (sum of -logp over valid rows, #valid, #correct). */
int pa_dist_loss_fwd(const float* lv, const float* lp, const float* sw, const int64_t* label, int B, int T,
                     int V, int pad, float inv_d, float* rowstat, int64_t* predict, float* accum,
                     void* stream);
/* gout: device scalar dL/dloss.  Writes dlv [N,V], dlp [N,T] (wrt the RAW scores), dsw [N]. */
int pa_dist_loss_bwd(const float* lv, const float* lp, const float* sw, const int64_t* label,
                     const float* rowstat, const float* accum, const float* gout, int B, int T, int V,
                     int pad, float inv_d, float* dlv, float* dlp, float* dsw, int round_tf32, void* stream);
/* Optional full distribution (parity tests only): dists [N, V+T] as models.py:186 builds it. */
int pa_dist_train_full(const float* lv, const float* lp, const float* sw, int B, int T, int V, float inv_d,
                       float* dists, void* stream);

/* ---- batched tokeniser with varlen input (replaces LineDataset.prepare_input_sequence /
 * prepare_output_sequence, plankassembly/datasets/line_data.py:34-83, 85-109, data_utils.py:6-12, per batch instead of per
 * sample in CPU workers).  lines [n_total, 4] fp64 (x1, y1, x2, y2 in [-1, 1]), views / types [n_total] int64 (types may be
 * NULL: sideface batches, then `type` may be NULL too), line_off [B+1] int32 = first line of every drawing.  Outputs: the
 * six planes of the reference's batch dict, [B, S] int64 / uint8 (S = MAX_INPUT_LENGTH - 1), kv_len [B] int32 = 4 n + 1
 * (the valid length the attention kernels accept), err: 0, or 1 + index of a drawing with more lines than fit.
 * pa_tokenize_planks: coords [c_total] fp64 + attach [c_total] int64 (-1 = not attached) with coord_off [B+1] ->
 * output_value / output_label [B, T] int64, output_mask [B, T] uint8, out_len [B] int32 = n + 1. */
int pa_tokenize_lines(const double* lines, const int64_t* views, const int64_t* types, const int* line_off, int B, int S,
                      int n_bits, int end_token, int pad_token, int64_t* value, int64_t* pos, int64_t* coord, int64_t* view,
                      int64_t* type, uint8_t* mask, int* kv_len, int* err, void* stream);
int pa_tokenize_planks(const double* coords, const int64_t* attach, const int* coord_off, int B, int T, int n_bits,
                       int end_token, int pad_token, int vocab, int64_t* value, int64_t* label, uint8_t* mask, int* out_len,
                       int* err, void* stream);

/* ---- batched validation/test post-processing (replaces the per-sample loops of models.py:258-265,309-315,
 * trainer_complete.py:76-80,97-101 and third_party/boxes.py:197-242; the Hungarian matching stays on the CPU).
 * pa_parse_sequences: seq [B, n] int64 (row pitch ld) -> planks [B, p_max, dof] int64 = the whole planks before the first
 *   end_token (0 beyond), n_planks [B] int32 (clipped to p_max), keep [B, p_max] uint8 = plank 0, or all extents
 *   (coords dof/2.. minus coords 0..) non-zero -- the trainer's zero-extent filter.
 * pa_plank_iou (dof = 6): for drawing b, rows = the kept predicted planks after plank 0 (in order), columns = ground-truth
 *   planks 1..n_gt-1; iou [B, p_max-1, g_max-1] fp32 (0 beyond the valid block), bit-identical to pairwise_iou();
 *   n_rows [B] int32; row_src [B, p_max-1] int32 = source plank index of every row (-1 beyond). */
int pa_parse_sequences(const int64_t* seq, int64_t ld, int B, int n, int end_token, int dof, int64_t* planks, int p_max,
                       int* n_planks, uint8_t* keep, void* stream);
int pa_plank_iou(const int64_t* pred, const uint8_t* keep, const int* n_pred, int p_max, const int64_t* gt, const int* n_gt,
                 int g_max, int B, float* iou, int* n_rows, int* row_src, void* stream);

/* ---- fused Adam over a flat fp32 master buffer that also writes the TF32 shadow weights
 * (replaces torch.optim.Adam's multi_tensor_apply launches of trainer_complete.py:127-129 plus one pa_round_tf32 per weight).
 * p, m, v, shadow (may be NULL): flat buffers of the same length; parameter i occupies [param_off[i], param_off[i] +
 * param_len[i]) (offsets multiples of 4 elements); grads[i]: device pointer of its gradient (NULL = skip, as torch does for
 * .grad None); the chunk tables split every parameter into pa_adam_chunk_elems()-element pieces: chunk c belongs to parameter
 * chunk_param[c] and starts at element chunk_off[c] of it; scalars[2i] = lr / (1 - beta1^t_i), scalars[2i+1] =
 * 1 / sqrt(1 - beta2^t_i) with t_i the step count of parameter i (torch counts steps per parameter).  All tables live in
 * device memory.  Arithmetic of torch/optim/adam.py::_single_tensor_adam. */
int pa_adam_chunk_elems(void);
int pa_adam_flat(float* p, float* m, float* v, float* shadow, const float* const* grads, const int64_t* chunk_off,
                 const int* chunk_param, const int64_t* param_off, const int64_t* param_len, const float* scalars,
                 int n_chunks, float beta1, float beta2, float eps, void* stream);

/* Exact-fp32 small-M projection for the decode step: C[M,N] = X[M,K] W[N,K]^T + bias (relu optional).
 * K % 32 == 0 and K <= 1536 (the W slice of a CTA lives in smem). */
int pa_gemm_skinny_f32(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, float* c,
                       int64_t ldc, int M, int N, int K, int relu, void* stream);

/* ---- K11/K12: KV-cached greedy decode step pieces (replace the O(T^3) loop models.py:284-307).
 * All state lives in caller-owned device buffers; `t` is the 0-based step.  When `t_dev` is not NULL the
 * step index is read from device memory instead, so that ONE captured CUDA graph replays every step;
 * pa_decode_advance increments it at the end of a step. */
int pa_decode_advance(int* t_dev, void* stream);
int pa_decode_embed(const int64_t* samples, int64_t ld, int B, int t, const int* t_dev, int dof, const float* e_val,
                    const float* e_coord, const float* e_pos, int d, float* y, void* stream);
/* Self-attention for the new position: appends k_new/v_new ([B,ld_new]) at slot t of the caches
 * (cache_len rows of stride ld_cache per sequence) and attends over slots 0..len-1 (len = t+1).
 * Cross-attention: pass k_new = NULL, len = S and kpm ([B,len]); kv_len ([B] int32, may be NULL) = 1 + index of the last
 * non-PAD key of each sequence: keys beyond it are not streamed at all (exact: they carry -inf). */
int pa_decode_attn(const float* q, int64_t ldq, const float* k_new, const float* v_new, int64_t ld_new,
                   float* k_cache, float* v_cache, int64_t cache_len, int64_t ld_cache, int t, int len,
                   const int* t_dev, const uint8_t* kpm, const int* kv_len, int B, int H, int dh, float scale, float* o, void* stream);
/* Heads + eval distribution + sampling for step t (models.py:168-186, 235-256).
 * h [B,d] final-normed hidden (also stored into hfin[:,t,:]); lv [B,V]; pf [B,d]; sw [B].
 * Writes samples[b*ld+t], attach[b*ld+t]; first_end[b] = min(first_end[b], t) when the emitted
 * token is END (the host derives the reference's stop step, models.py:306, as max_b first_end). */
int pa_decode_head(const float* h, const float* lv, const float* pf, const float* sw, float* hfin,
                   int64_t Tmax, int B, int d, int V, int t, const int* t_dev, int end_token, int64_t* samples,
                   int64_t* attach, int64_t ld, int32_t* first_end, void* stream);

/* ---- K12, fused: the WHOLE greedy decode loop (every step of every sequence, models.py:284-307) as one
 * persistent cooperative kernel: per step 6 x [QKV proj, cached self-attention, out proj, residual+LN, q proj,
 * cross-attention over the once-projected memory K/V, out proj, residual+LN, FFN, residual+LN], final LN, heads,
 * eval distribution, sampling, END bookkeeping and the next input embedding, with grid-wide barriers between
 * the phases and no host involvement.  All fp32.  The caller owns every buffer:
 *   layers[l].self_k/self_v [B,T,d] (filled by the kernel), cross_kv [B,S,2d] (K | V of the encoder memory,
 *   projected by the caller with multihead_attn.in_proj rows d..3d), w_heads [V+d+1, d] = vocab_head |
 *   pointer_head | switch_head weights stacked (b_heads likewise), part = part_bytes = pa_decode_fused_workspace()
 *   bytes, state = PA_DEC_STATE_INTS int32 (zeroed by the call; on return state[2] = number of steps executed =
 *   the reference's output length, state[1] = sequences that emitted END, state[3] = last first-END step).
 * chains: number of independent sub-batches the grid is divided into (0 = automatic), see decode_fused.cu.
 * Outputs: samples/attach [B,T] (attach = -1 where the token was not copied), first_end [B], hfin [B,T,d]. */
#define PA_MAX_DEC_LAYERS 8
#define PA_MAX_DEC_CHAINS 8
#define PA_DEC_STATE_INTS 16
typedef struct {
  const float *w_sqkv, *b_sqkv;    /* self_attn.in_proj_weight [3d,d], in_proj_bias [3d] */
  const float *w_so, *b_so;        /* self_attn.out_proj */
  const float *g1, *be1;           /* norm1 */
  const float *w_cq, *b_cq;        /* multihead_attn.in_proj rows 0..d-1 (query projection) */
  const float *w_co, *b_co;        /* multihead_attn.out_proj */
  const float *g2, *be2;           /* norm2 */
  const float *w_f1, *b_f1;        /* linear1 [ff,d] */
  const float *w_f2, *b_f2;        /* linear2 [d,ff] */
  const float *g3, *be3;           /* norm3 */
  float* self_k; float* self_v;
  const float* cross_kv;
} pa_decode_layer;
typedef struct {
  int B, S, T, d, H, ff, V, L, dof, end_token;
  float layer_eps, final_eps;
  pa_decode_layer layers[PA_MAX_DEC_LAYERS];
  const float *gf, *bf;            /* decoder.norm */
  const float *w_heads, *b_heads;
  const float *e_val, *e_coord, *e_pos;
  const uint8_t* kpm;              /* [B,S], 1 = PAD memory key */
  float* y; float* o;              /* [B,d] scratch rows */
  float* part; int64_t part_bytes;
  float* hfin;
  int64_t* samples; int64_t* attach;
  int32_t* first_end;
  int32_t* state;
  int chains;
  int profile;                     /* != 0: CTA 0 accumulates nanoseconds per phase kind, read back with pa_debug_decode_prof */
} pa_decode_fused_args;
size_t pa_decode_fused_workspace(int B, int d, int ff, int V);
int pa_decode_fused(const pa_decode_fused_args* args, void* stream);
/* Debug: ns spent by CTA 0 of the last profiled pa_decode_fused in [gemm, self-attn, row, cross-attn, head, barriers, -, -]. */
int pa_debug_decode_prof(unsigned long long* out8_host);

/* Debug: per-role barrier wait cycles of CTA 0 of the last tensor-core attention forward that ran with
 * PLANK_B200_ATTN_DEBUG bit 64 set (32 counters, see attn_tc.cu). */
int pa_debug_attn_prof(unsigned long long* out32_host);

#ifdef __cplusplus
}
#endif
#endif /* PLANK_B200_H */
