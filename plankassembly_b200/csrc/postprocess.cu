// validation/test post-processing on the device, batched.
//
// The reference parses every decoded sequence on its own (ref models.py:258-265: cumsum over == END, boolean-mask gather,
// truncate to whole planks, reshape; called twice per sample from eval_step :309-315), then the trainer drops zero-extent
// boxes (ref trainer_complete.py:76-80, 97-101) and builds the 3-D IoU matrix (ref third_party/boxes.py:197-242) per sample
// -- ~15 tiny launches and several host syncs per drawing.  Here: one launch parses a whole batch of sequences, one launch
// filters + computes every IoU matrix; the Hungarian assignment (third_party/matcher.py, scipy) stays on the CPU.
//
// IoU arithmetic is the reference's, in the same fp32 order, so the matrices are bit-identical to pairwise_iou():
//   lwh = min(max1, max2) - max(min1, min2), clamped at 0;  inter = (l*w)*h;  vol = ((x1-x0)*(y1-y0))*(z1-z0);
//   iou = inter > 0 ? inter / ((vol1 + vol2) - inter) : 0
#include "common.cuh"

namespace {

// One warp per sequence.  planks[b, j, c] = seq[b, j*dof + c] for the whole planks before the first END, 0 beyond;
// keep[b, j] = 1 for plank 0 (the bounding box is never filtered) and for planks whose three extents are all non-zero.
__global__ void __launch_bounds__(128) parse_sequences_kernel(const int64_t* __restrict__ seq, int64_t ld, int B, int n, int end_token, int dof,
                                                              int64_t* __restrict__ planks, int p_max, int* __restrict__ n_planks,
                                                              uint8_t* __restrict__ keep) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const int64_t* s = seq + (int64_t)b * ld;
  int first_end = n;
  for (int j0 = 0; j0 < n && first_end == n; j0 += 32) {
    const int j = j0 + lane;
    const unsigned hit = __ballot_sync(0xffffffffu, j < n && s[j] == end_token);
    if (hit) first_end = j0 + __ffs(hit) - 1;
  }
  const int np = min(first_end / dof, p_max);
  if (lane == 0) n_planks[b] = np;
  int64_t* out = planks + (int64_t)b * p_max * dof;
  for (int i = lane; i < p_max * dof; i += 32) out[i] = i < np * dof ? s[i] : 0;
  const int half = dof / 2;
  for (int j = lane; j < p_max; j += 32) {
    bool ok = j < np;
    if (ok && j > 0)
      for (int c = 0; c < half; ++c) ok = ok && (s[j * dof + half + c] - s[j * dof + c] != 0);
    keep[(int64_t)b * p_max + j] = ok ? 1 : 0;
  }
}

// One CTA per drawing: rows = kept predictions after plank 0, columns = ground-truth planks after plank 0.
__global__ void __launch_bounds__(128) plank_iou_kernel(const int64_t* __restrict__ pred, const uint8_t* __restrict__ keep, const int* __restrict__ n_pred,
                                                        int p_max, const int64_t* __restrict__ gt, const int* __restrict__ n_gt, int g_max,
                                                        float* __restrict__ iou, int* __restrict__ n_rows, int* __restrict__ row_src) {
  extern __shared__ int s_rows[];                     // [p_max] source plank index of every kept row
  __shared__ int s_n;
  const int b = blockIdx.x;
  const int np = n_pred[b], ng = max(n_gt[b] - 1, 0);
  if (threadIdx.x == 0) {
    int r = 0;
    for (int j = 1; j < np; ++j)
      if (keep[(int64_t)b * p_max + j]) s_rows[r++] = j;
    s_n = r;
    n_rows[b] = r;
  }
  __syncthreads();
  const int nr = s_n, R = p_max - 1, Cc = g_max - 1;
  for (int i = threadIdx.x; i < R; i += blockDim.x) row_src[(int64_t)b * R + i] = i < nr ? s_rows[i] : -1;
  for (int i = threadIdx.x; i < R * Cc; i += blockDim.x) {
    const int r = i / Cc, c = i - r * Cc;
    float v = 0.f;
    if (r < nr && c < ng) {
      const int64_t* p = pred + ((int64_t)b * p_max + s_rows[r]) * 6;
      const int64_t* g = gt + ((int64_t)b * g_max + c + 1) * 6;
      float pm[6], gm[6];
#pragma unroll
      for (int e = 0; e < 6; ++e) { pm[e] = (float)p[e]; gm[e] = (float)g[e]; }
      const float vp = __fmul_rn(__fmul_rn(pm[3] - pm[0], pm[4] - pm[1]), pm[5] - pm[2]);
      const float vg = __fmul_rn(__fmul_rn(gm[3] - gm[0], gm[4] - gm[1]), gm[5] - gm[2]);
      float l[3];
#pragma unroll
      for (int e = 0; e < 3; ++e) l[e] = fmaxf(fminf(pm[3 + e], gm[3 + e]) - fmaxf(pm[e], gm[e]), 0.f);
      const float inter = __fmul_rn(__fmul_rn(l[0], l[1]), l[2]);
      v = inter > 0.f ? __fdiv_rn(inter, __fsub_rn(__fadd_rn(vp, vg), inter)) : 0.f;
    }
    iou[(int64_t)b * R * Cc + i] = v;
  }
}

}  // namespace

extern "C" int pa_parse_sequences(const int64_t* seq, int64_t ld, int B, int n, int end_token, int dof, int64_t* planks, int p_max,
                                  int* n_planks, uint8_t* keep, void* stream) {
  PA_CHECK_ARG(seq != nullptr && planks != nullptr && n_planks != nullptr && keep != nullptr);
  PA_CHECK_ARG(B > 0 && n >= 0 && dof > 0 && dof % 2 == 0 && p_max > 0 && ld >= n);
  parse_sequences_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(seq, ld, B, n, end_token, dof, planks, p_max, n_planks, keep);
  PA_CHECK_LAUNCH();
  return PA_OK;
}

extern "C" int pa_plank_iou(const int64_t* pred, const uint8_t* keep, const int* n_pred, int p_max, const int64_t* gt, const int* n_gt,
                            int g_max, int B, float* iou, int* n_rows, int* row_src, void* stream) {
  PA_CHECK_ARG(pred != nullptr && keep != nullptr && n_pred != nullptr && gt != nullptr && n_gt != nullptr && iou != nullptr);
  PA_CHECK_ARG(n_rows != nullptr && row_src != nullptr && B > 0 && p_max > 1 && g_max > 1);
  plank_iou_kernel<<<B, 128, p_max * sizeof(int), (cudaStream_t)stream>>>(pred, keep, n_pred, p_max, gt, n_gt, g_max, iou, n_rows, row_src);
  PA_CHECK_LAUNCH();
  return PA_OK;
}
