// batched tokeniser -- the device-side equivalent of LineDataset.prepare_input_sequence /
// prepare_output_sequence (ref: plankassembly/datasets/line_data.py:34-83, 85-109; quantisation
// plankassembly/datasets/data_utils.py:6-12), for a whole batch of drawings given in VARLEN form (all lines of all
// drawings concatenated + an offsets array), so that the data loader ships raw geometry (32 B per line + 2 ids) instead
// of six padded int64 planes, and the per-sequence valid length (kv_len) comes out with the tokens: the attention
// kernels skip the padding it describes.
//
// Per drawing (one CTA): quantise the 4 coordinates of every line in fp64 exactly as numpy does
// (((v - (-1)) * 511) / 2, truncated), stable-sort the lines by (view, x1, x2, y1, y2) -- np.lexsort over
// line_with_view.T[[3, 1, 2, 0, 4]], last key primary -- with a bitonic network over 64-bit keys that carry the original
// index in their low bits (=> stable), position = rank inside the view group, coordinate id = i % 4, then END, then PAD
// (value plane) / 0 (the other planes); mask = (value == PAD).
#include "common.cuh"

namespace {

constexpr int kMaxLines = 512;       // lines per drawing handled by one CTA (MAX_INPUT_LENGTH 1200 -> 299)

__device__ __forceinline__ long long quantize_f64(double v, double range_q) {
  return (long long)(((v - (-1.0)) * range_q) / (1.0 - (-1.0)));          // numpy: astype('long') truncates toward zero
}

__global__ void __launch_bounds__(256) tokenize_lines_kernel(const double* __restrict__ lines, const int64_t* __restrict__ views,
                                                             const int64_t* __restrict__ types, const int* __restrict__ line_off, int S,
                                                             int n_bits, int end_token, int pad_token, int64_t* __restrict__ o_value,
                                                             int64_t* __restrict__ o_pos, int64_t* __restrict__ o_coord, int64_t* __restrict__ o_view,
                                                             int64_t* __restrict__ o_type, uint8_t* __restrict__ o_mask, int* __restrict__ kv_len,
                                                             int* __restrict__ err) {
  __shared__ unsigned long long key[kMaxLines];
  __shared__ int view_count[4];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int l0 = line_off[b], n = line_off[b + 1] - l0;
  const double range_q = (double)((1 << n_bits) - 1);
  if (n > kMaxLines || 4 * n + 1 > S) {          // the reference would fail in np.pad with a negative width
    if (tid == 0) { atomicExch(err, b + 1); kv_len[b] = 0; }
    return;
  }
  if (tid < 4) view_count[tid] = 0;
  __syncthreads();
  for (int i = tid; i < kMaxLines; i += blockDim.x) {
    unsigned long long k = ~0ull;                                           // padding sorts last
    if (i < n) {
      const double* ln = lines + (int64_t)(l0 + i) * 4;
      const unsigned long long x1 = quantize_f64(ln[0], range_q) & 0x3ff, y1 = quantize_f64(ln[1], range_q) & 0x3ff;
      const unsigned long long x2 = quantize_f64(ln[2], range_q) & 0x3ff, y2 = quantize_f64(ln[3], range_q) & 0x3ff;
      const unsigned long long vw = (unsigned long long)views[l0 + i] & 0x3;
      k = (vw << 50) | (x1 << 40) | (x2 << 30) | (y1 << 20) | (y2 << 10) | (unsigned long long)i;
      atomicAdd(&view_count[vw], 1);
    }
    key[i] = k;
  }
  __syncthreads();
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int k2 = 2; k2 <= np2; k2 <<= 1)
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < np2; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const unsigned long long a = key[i], c = key[p];
          const bool up = (i & k2) == 0;
          if ((a > c) == up) { key[i] = c; key[p] = a; }
        }
      }
      __syncthreads();
    }
  const int start1 = view_count[0], start2 = view_count[0] + view_count[1], start3 = start2 + view_count[2];
  int64_t* vrow = o_value + (int64_t)b * S;
  int64_t* prow = o_pos + (int64_t)b * S;
  int64_t* crow = o_coord + (int64_t)b * S;
  int64_t* wrow = o_view + (int64_t)b * S;
  int64_t* trow = o_type != nullptr ? o_type + (int64_t)b * S : nullptr;
  uint8_t* mrow = o_mask + (int64_t)b * S;
  for (int t = tid; t < S; t += blockDim.x) {
    int64_t val = pad_token, pos = 0, coord = 0, vw = 0, ty = 0;
    if (t < 4 * n) {
      const int i = t >> 2, c = t & 3;
      const unsigned long long k = key[i];
      const int src = (int)(k & 0x3ff);
      vw = (int64_t)((k >> 50) & 0x3);
      const int shift = c == 0 ? 40 : (c == 1 ? 20 : (c == 2 ? 30 : 10));       // token order x1, y1, x2, y2
      val = (int64_t)((k >> shift) & 0x3ff);
      pos = i - (vw == 0 ? 0 : (vw == 1 ? start1 : (vw == 2 ? start2 : start3)));
      coord = c;
      ty = types != nullptr ? types[l0 + src] : 0;
    } else if (t == 4 * n) {
      val = end_token;
    }
    vrow[t] = val; prow[t] = pos; crow[t] = coord; wrow[t] = vw;
    if (trow != nullptr) trow[t] = ty;
    mrow[t] = val == pad_token ? 1 : 0;
  }
  if (tid == 0) kv_len[b] = 4 * n + 1;
}

// prepare_output_sequence: value = quantised plank coordinates, END, PAD; label = attach + vocab where attach != -1 else value
__global__ void __launch_bounds__(256) tokenize_planks_kernel(const double* __restrict__ coords, const int64_t* __restrict__ attach,
                                                              const int* __restrict__ coord_off, int T, int n_bits, int end_token, int pad_token,
                                                              int vocab, int64_t* __restrict__ o_value, int64_t* __restrict__ o_label,
                                                              uint8_t* __restrict__ o_mask, int* __restrict__ out_len, int* __restrict__ err) {
  const int b = blockIdx.x;
  const int c0 = coord_off[b], n = coord_off[b + 1] - c0;
  const double range_q = (double)((1 << n_bits) - 1);
  if (n + 1 > T) {
    if (threadIdx.x == 0) { atomicExch(err, b + 1); out_len[b] = 0; }
    return;
  }
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    int64_t val = pad_token, lab = pad_token;
    if (t < n) {
      val = quantize_f64(coords[c0 + t], range_q);
      const int64_t a = attach[c0 + t];
      lab = a != -1 ? a + vocab : val;
    } else if (t == n) {
      val = end_token; lab = end_token;
    }
    o_value[(int64_t)b * T + t] = val;
    o_label[(int64_t)b * T + t] = lab;
    o_mask[(int64_t)b * T + t] = val == pad_token ? 1 : 0;
  }
  if (threadIdx.x == 0) out_len[b] = n + 1;
}

}  // namespace

extern "C" int pa_tokenize_lines(const double* lines, const int64_t* views, const int64_t* types, const int* line_off, int B, int S,
                                 int n_bits, int end_token, int pad_token, int64_t* value, int64_t* pos, int64_t* coord, int64_t* view,
                                 int64_t* type, uint8_t* mask, int* kv_len, int* err, void* stream) {
  PA_CHECK_ARG(lines != nullptr && views != nullptr && line_off != nullptr && value != nullptr && pos != nullptr && coord != nullptr);
  PA_CHECK_ARG(view != nullptr && mask != nullptr && kv_len != nullptr && err != nullptr && B > 0 && S > 0 && n_bits > 0 && n_bits <= 10);
  tokenize_lines_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(lines, views, types, line_off, S, n_bits, end_token, pad_token, value, pos, coord,
                                                              view, type, mask, kv_len, err);
  PA_CHECK_LAUNCH();
  return PA_OK;
}

extern "C" int pa_tokenize_planks(const double* coords, const int64_t* attach, const int* coord_off, int B, int T, int n_bits,
                                  int end_token, int pad_token, int vocab, int64_t* value, int64_t* label, uint8_t* mask, int* out_len,
                                  int* err, void* stream) {
  PA_CHECK_ARG(coords != nullptr && attach != nullptr && coord_off != nullptr && value != nullptr && label != nullptr && mask != nullptr);
  PA_CHECK_ARG(out_len != nullptr && err != nullptr && B > 0 && T > 0 && n_bits > 0 && n_bits <= 10);
  tokenize_planks_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(coords, attach, coord_off, T, n_bits, end_token, pad_token, vocab, value, label,
                                                               mask, out_len, err);
  PA_CHECK_LAUNCH();
  return PA_OK;
}
