// Attention-probability dropout masks as bit planes, generated once per attention call by the whole chip.
// Inside the tensor-core attention kernels only 4 warps per SM do elementwise work, and Philox4x32-10 was
// ~2/3 of their instructions (and was recomputed in the forward, dQ and dK/dV kernels).  Here every thread
// owns (query q, 32 consecutive keys): 4 Philox calls folded bit-sliced into one row-major mask word (drop_keep_word, common.cuh);
// a shuffle bit-matrix transpose of the 32x32 block gives the key-stationary backward its (key, 32 queries) word.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(128) dropout_mask_kernel(uint32_t* __restrict__ rows, uint32_t* __restrict__ cols, int Lq, int Lk,
                                                            int LkW, int LqW, uint32_t thr, uint64_t seed, uint64_t offset) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kw = blockIdx.x * 4 + warp;            // key word (32 keys)
  const int qb = blockIdx.y;                       // query block (32 queries)
  const int bh = blockIdx.z;
  if (kw >= LkW) return;
  const int q = qb * 32 + lane;
  uint32_t word = 0;
  if (q < Lq) {
    const int64_t rg = (int64_t)bh * Lq + q;
    word = drop_keep_word(seed, offset, rg, LkW, kw, 65536u - thr);
    rows[rg * LkW + kw] = word;
  }
  if (cols != nullptr) {
    // 32x32 bit-matrix transpose across the warp (lane = query, bit = key  ->  lane = key, bit = query) in five
    // shuffle stages instead of 32 ballots: stage j swaps the off-diagonal j x j blocks of every 2j x 2j block.
    uint32_t w = word;
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
      // m selects the bit positions whose bit j is clear: 0x0000FFFF, 0x00FF00FF, 0x0F0F0F0F, 0x33333333, 0x55555555
      const uint32_t m = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
      const uint32_t other = __shfl_xor_sync(0xffffffffu, w, j);
      // lanes with bit j clear keep their low-j columns and take the partner's low-j columns as their high ones;
      // lanes with bit j set keep their high-j columns and take the partner's high ones as their low ones
      w = (lane & j) ? ((w & ~m) | ((other >> j) & m)) : ((w & m) | ((other << j) & ~m));
    }
    cols[((int64_t)bh * (LkW * 32) + kw * 32 + lane) * LqW + qb] = w;     // bit q%32 of key kw*32+lane
  }
}

}  // namespace

extern "C" size_t pa_dropout_mask_words(int BH, int Lq, int Lk, int cols) {
  const size_t LkW = (Lk + 31) / 32, LqW = (Lq + 31) / 32;
  return cols ? (size_t)BH * LkW * 32 * LqW : (size_t)BH * Lq * LkW;
}

extern "C" int pa_dropout_mask(uint32_t* rows, uint32_t* cols, int BH, int Lq, int Lk, float p_drop, uint64_t seed,
                               uint64_t offset, void* stream) {
  PA_CHECK_ARG(rows != nullptr && BH > 0 && Lq > 0 && Lk > 0 && p_drop > 0.f && p_drop < 1.f);
  const int LkW = (Lk + 31) / 32, LqW = (Lq + 31) / 32;
  dim3 grid((LkW + 3) / 4, LqW, BH);
  dropout_mask_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(rows, cols, Lq, Lk, LkW, LqW, drop_threshold16(p_drop), seed, offset);
  PA_CHECK_LAUNCH();
  return PA_OK;
}
