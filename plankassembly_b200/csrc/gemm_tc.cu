// Dense projections on the 5th-gen tensor cores: C = alpha * op(A) op(B)^T (+bias)(relu)(dropout),
// TF32 inputs read straight from the fp32 tensors, FP32 accumulation in TMEM.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer  (cp.async.bulk.tensor, 128B-swizzled tiles, mbarrier complete_tx)
//   warp 1      MMA issuer    (tcgen05.mma.cta_group::1.kind::tf32, one elected lane) + TMEM owner
//   warps 2..9  epilogue      (tcgen05.ld 32x32b -> alpha / bias / ReLU / bit-plane mask / bit-sliced dropout / column sums /
//                              TF32 rounding, each stage a 32-wide loop of independent operations -> 128B-swizzled smem
//                              tile -> TMA store or reduce-add); two warps per TMEM lane quarter take alternate
//                              32-column chunks, each warp with two staging tiles
// Three pipelines: smem full/empty ring (TMA <-> MMA), double-buffered TMEM accumulators
// (MMA <-> epilogue), static persistent tile schedule.
//
// Operand forms (all row-major fp32 in HBM):
//   A K-major : A[M,K]   (activations x, dY)          A MN-major: stored [K,M]  (dY for dW = dY^T X)
//   B K-major : B[N,K]   (weights W, or W^T copies)   B MN-major: stored [K,N]  (x for dW)
// Used for: every nn.Linear on the path (ref models.py:60-74 via torch transformer.py), their
// input/weight gradients, and the pointer scoring bmm (ref models.py:149) in batched mode.
#include "common.cuh"
#include "tc_common.cuh"

#include <mutex>
#include <stdlib.h>

// Debug timeline (PLANK_B200_NVCC_FLAGS=-DPA_GEMM_TRACE, PLANK_B200_GEMM_DEBUG bit 8): clock64 per event of CTA 0
// (roles: 0 TMA producer, 1 MMA issuer, 2 epilogue warp 2); read back with pa_debug_gemm_trace().
#ifdef PA_GEMM_TRACE
__device__ unsigned long long g_gemm_trace[3][2048];
__device__ int g_gemm_trace_n[3];
#define GTRACE(role, ev)                                                                          \
  do {                                                                                           \
    if (gtrace_on && gtrace_n < 2048) g_gemm_trace[role][gtrace_n++] = ((unsigned long long)(ev) << 56) | (clock64() & 0xffffffffffffffull); \
  } while (0)
#define GTRACE_DECL(cond) const bool gtrace_on = (p.debug & 8) && blockIdx.x == 0 && (cond); int gtrace_n = 0
#define GTRACE_END(role) do { if (gtrace_on) g_gemm_trace_n[role] = gtrace_n; } while (0)
#else
#define GTRACE(role, ev) do {} while (0)
#define GTRACE_DECL(cond) do {} while (0)
#define GTRACE_END(role) do {} while (0)
#endif

namespace {

constexpr int BM = 128;        // UMMA M (cta_group::1)
constexpr int BK = 32;         // 32 tf32 = 128 B = one swizzle row
constexpr int UK = 8;          // K per tcgen05.mma.kind::tf32
#ifndef PA_GEMM_STORE_SLOTS
#define PA_GEMM_STORE_SLOTS 2
#endif
constexpr int kThreads = 320;      // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)

struct GemmParams {
  float* c; int64_t ldc; int64_t c_batch_stride;
  const float* bias;
  int M, N, K;
  int batch; int a_batch_rows, b_batch_rows;
  int split_k; int accumulate;
  int relu; float p_drop; uint64_t seed, offset;
  float alpha;
  int round_out;
  int debug;              // bit 0: skip the epilogue body (mainloop-only timing experiments)
  int tma_store;          // epilogue stores through smem + TMA (needs a 16-byte row pitch)
  int epi_direct;         // 1: through the same swizzled smem tile and coalesced st.global.v4 (measured slower); 2: register-direct
  int wide8;              // C rows are 32-byte aligned: the register-direct epilogue may use 256-bit stores
  uint32_t* mask_out; const uint32_t* mask_in; float* colsum; float mask_scale;   // fused FFN activation backward (see the header)
  int m_tiles, n_tiles;
};

// TWO = 2-SM UMMA (tcgen05 cta_group::2): a CTA pair works on ONE 256 x BN tile; each CTA stages its 128 rows of A and only
// HALF of the B tile (BN/2 rows), the tensor cores of the two SMs exchange the halves -- a third less shared-memory traffic
// per flop than the single-SM form (fp32 operands make these GEMMs smem-bandwidth bound) and room for deeper pipelines.
template <int BN, bool TWO = false> struct Cfg {
  static constexpr int kSlots = PA_GEMM_STORE_SLOTS;     // staging tiles per epilogue warp (TMA-store epilogue)
  static constexpr int kStages = kSlots > 1 ? (TWO ? (BN == 256 ? 5 : 6) : (BN == 256 ? 3 : 5))
                                            : (TWO ? (BN == 256 ? 6 : 8) : (BN == 256 ? 4 : 6));
  static constexpr int kABytes = BM * BK * 4;           // 16 KB
  static constexpr int kBBytes = (TWO ? BN / 2 : BN) * BK * 4;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStageOut = 8 * kSlots * 32 * 128;   // per epilogue warp: kSlots 32x32 fp32 tiles, 128B-swizzled
  static constexpr int kSmem = kStages * kStageBytes + kStageOut + 1024 /*align*/ + 256 /*barriers*/ + BN * 4 /*bias*/;
  static constexpr int kTmemCols = 2 * BN;              // double-buffered accumulator
  static_assert(kSmem <= 232448, "dynamic shared memory budget of one CTA");
};

// CL = 2: CTA pairs (cluster of 2 along M) share every B (weight) tile: each CTA fetches half of it and TMA
// multicasts the half into both CTAs' smem -> 1/3 less L2->smem traffic per CTA (the K=512 GEMMs are L2-bound).
template <int BN, bool A_MN, bool B_MN, int CL, bool TWO>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_c, const GemmParams p) {
  static_assert(!TWO || CL == 2, "the 2-SM UMMA form runs on CTA pairs");
  using C = Cfg<BN, TWO>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* out_stage = smem + C::kStages * C::kStageBytes;                 // [8 warps][kSlots][32 rows][128 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_stage + C::kStageOut);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tmem_full = empty_bar + C::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_s = reinterpret_cast<float*>(out_stage + C::kStageOut + 256);   // [BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tma_a);
    tc::tma_prefetch_desc(&tma_b);
    if (p.tma_store) tc::tma_prefetch_desc(&tma_c);
    // TWO: full / tmem_empty are used in the leader (cluster rank 0) only; empty / tmem_full get one multicast commit each
    for (int s = 0; s < C::kStages; ++s) { tc::mbar_init(full_bar + s, 1); tc::mbar_init(empty_bar + s, TWO ? 1 : CL); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(tmem_full + s, 1); tc::mbar_init(tmem_empty + s, TWO ? 16 : 8); }
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (TWO) tc::tmem_alloc_pair<C::kTmemCols>(tmem_slot);
    else tc::tmem_alloc<C::kTmemCols>(tmem_slot);
  }
  tc::tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) tc::cluster_sync();      // the peer's barriers exist before anything is multicast into it
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int crank = CL > 1 ? (int)tc::cluster_ctarank() : 0;
  const int first_tile = blockIdx.x / CL, tile_step = gridDim.x / CL;
  const int m_groups = (p.m_tiles + CL - 1) / CL;   // an m group = the CL consecutive m tiles of one cluster step

  const int tiles_per_batch = m_groups * p.n_tiles * p.split_k;
  const int num_tiles = tiles_per_batch * p.batch;
  const int k_blocks_total = (p.K + BK - 1) / BK;
  const int k_per_split = (k_blocks_total + p.split_k - 1) / p.split_k;

  auto decode = [&](int tile, int& b, int& mt, int& nt, int& kb0, int& kb1) {
    b = tile / tiles_per_batch;
    int r = tile - b * tiles_per_batch;
    int sp = r / (m_groups * p.n_tiles);
    r -= sp * m_groups * p.n_tiles;
    mt = (r / p.n_tiles) * CL + crank;   // n fastest: an activation tile is reused by all its n tiles out of L2
    nt = r % p.n_tiles;
    kb0 = sp * k_per_split;
    kb1 = min(k_blocks_total, kb0 + k_per_split);
  };

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      GTRACE_DECL(true);
      int stage = 0; uint32_t phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        int b, mt, nt, kb0, kb1;
        decode(tile, b, mt, nt, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          tc::mbar_wait(empty_bar + stage, phase ^ 1);
          GTRACE(0, 0);
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kABytes;
          if constexpr (TWO) {
            // both CTAs fill their own slot and complete the bytes on the LEADER's full barrier (it expects both halves)
            const uint32_t lead_full = tc::mapa_u32(tc::smem_u32(full_bar + stage), 0);
            if (crank == 0) tc::mbar_arrive_expect_tx(full_bar + stage, 2 * C::kStageBytes);
            if constexpr (!A_MN) {
              tc::tma_load_2d_pair(sa, &tma_a, kb * BK, b * p.a_batch_rows + mt * BM, lead_full);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 32; ++j) tc::tma_load_2d_pair(sa + j * (BK * 128), &tma_a, mt * BM + j * 32, b * p.a_batch_rows + kb * BK, lead_full);
            }
            if constexpr (!B_MN) {
              tc::tma_load_2d_pair(sb, &tma_b, kb * BK, b * p.b_batch_rows + nt * BN + crank * (BN / 2), lead_full);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tc::tma_load_2d_pair(sb + j * (BK * 128), &tma_b, nt * BN + (crank * (BN / 64) + j) * 32, b * p.b_batch_rows + kb * BK, lead_full);
            }
            if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            continue;
          }
          tc::mbar_arrive_expect_tx(full_bar + stage, C::kStageBytes);
          if constexpr (!A_MN) {
            tc::tma_load_2d(sa, &tma_a, kb * BK, b * p.a_batch_rows + mt * BM, full_bar + stage);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 32; ++j) tc::tma_load_2d(sa + j * (BK * 128), &tma_a, mt * BM + j * 32, b * p.a_batch_rows + kb * BK, full_bar + stage);
          }
          if constexpr (CL == 1) {
            if constexpr (!B_MN) {
              tc::tma_load_2d(sb, &tma_b, kb * BK, b * p.b_batch_rows + nt * BN, full_bar + stage);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 32; ++j) tc::tma_load_2d(sb + j * (BK * 128), &tma_b, nt * BN + j * 32, b * p.b_batch_rows + kb * BK, full_bar + stage);
            }
          } else {   // this CTA fetches its half of the B tile and multicasts it into both CTAs of the pair
            if constexpr (!B_MN) {
              tc::tma_load_2d_mcast(sb + crank * (BN / 2) * 128, &tma_b, kb * BK, b * p.b_batch_rows + nt * BN + crank * (BN / 2), full_bar + stage, 0x3);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j) {
                const int jj = crank * (BN / 64) + j;
                tc::tma_load_2d_mcast(sb + jj * (BK * 128), &tma_b, nt * BN + jj * 32, b * p.b_batch_rows + kb * BK, full_bar + stage, 0x3);
              }
            }
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
      GTRACE_END(0);
    }
  } else if (warp == 1 && (!TWO || crank == 0)) {
    // ===================================== MMA issuer (TWO: the leader CTA of the pair only) =
    // The whole warp walks the loop (warp-uniform control flow and addresses) and one elected lane issues: inside a
    // divergent `if (lane == 0)` ptxas wraps every UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~70 cycles).
    {
      constexpr uint32_t idesc = tc::make_idesc_tf32(TWO ? 2 * BM : BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      GTRACE_DECL(lane == 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        int b, mt, nt, kb0, kb1;
        decode(tile, b, mt, nt, kb0, kb1);
        tc::mbar_wait(tmem_empty + acc, acc_phase ^ 1);
        GTRACE(1, 0);
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          tc::mbar_wait(full_bar + stage, phase);
          GTRACE(1, 1);
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(smem + stage * C::kStageBytes);
          const uint32_t sb = sa + C::kABytes;
          // K-major: LBO unused, SBO = 1024 (8 rows x 128 B).  MN-major (tf32 => 128B_BASE32B): LBO = one
          // 32-wide MN block (BK rows x 128 B), SBO = 512 (4 K rows); one MMA (K=8) spans two K groups.
          const uint64_t da = A_MN ? tc::make_smem_desc(sa, BK * 128, 512, tc::kLayoutSw128Base32) : tc::make_smem_desc(sa, 16, 1024);
          const uint64_t db = B_MN ? tc::make_smem_desc(sb, BK * 128, 512, tc::kLayoutSw128Base32) : tc::make_smem_desc(sb, 16, 1024);
          if (tc::elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              const uint64_t dak = tc::desc_advance(da, A_MN ? k * 1024 : k * UK * 4);
              const uint64_t dbk = tc::desc_advance(db, B_MN ? k * 1024 : k * UK * 4);
              if constexpr (TWO) tc::mma_tf32_ss_pair(d_tmem, dak, dbk, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              else tc::mma_tf32_ss(d_tmem, dak, dbk, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            if constexpr (TWO) {
              tc::tc_commit_pair(empty_bar + stage, 0x3);                         // frees the slot in both CTAs
              if (kb + 1 == kb1) tc::tc_commit_pair(tmem_full + acc, 0x3);        // both halves of the accumulator -> both epilogues
            } else {
              if constexpr (CL > 1) tc::tc_commit_mcast(empty_bar + stage, 0x3);   // the slot is shared: both CTAs must be done with it
              else tc::tc_commit(empty_bar + stage);        // frees the smem slot when these MMAs retire
              if (kb + 1 == kb1) tc::tc_commit(tmem_full + acc);                  // accumulator complete -> epilogue
            }
          }
          __syncwarp();
          GTRACE(1, 2);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (kb0 >= kb1) {
          if (tc::elect_one()) { if constexpr (TWO) tc::tc_commit_pair(tmem_full + acc, 0x3); else tc::tc_commit(tmem_full + acc); }
          __syncwarp();
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      GTRACE_END(1);
    }
  } else if (warp >= 2) {
    // ===================================== epilogue =========================================
    const uint32_t bias_u32 = tc::smem_u32(bias_s);
    const int q = warp & 3;                            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;                  // which of the quarter's two warps: takes chunks half, half+2, ...
    int acc = 0; uint32_t acc_phase = 0;
    int slot = 0;
    const uint32_t keep16 = 65536u - drop_threshold16(p.p_drop);
    const float ks = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;
    GTRACE_DECL(warp == 2 && lane == 0);
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      int b, mt, nt, kb0, kb1;
      decode(tile, b, mt, nt, kb0, kb1);
      // stage this tile's bias slice in smem once (per-element global loads serialised the epilogue)
      if (p.bias != nullptr) {
        asm volatile("bar.sync 1, 256;" ::: "memory");    // everyone is done with the previous tile's bias
        const int et = threadIdx.x - 64;
        if (et < BN) bias_s[et] = (nt * BN + et < p.N) ? __ldg(p.bias + nt * BN + et) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      GTRACE(2, 0);
      tc::mbar_wait(tmem_full + acc, acc_phase);
      GTRACE(2, 1);
      tc::tc_fence_after();
      const int row = mt * BM + q * 32 + lane;
      const bool row_ok = row < p.M && kb1 > kb0;
      float* crow = p.c + (int64_t)b * p.c_batch_stride + (int64_t)row * p.ldc;
#pragma unroll 1
      for (int c0 = half * 32; c0 < ((p.debug & 1) ? 0 : BN); c0 += 64) {
        // Every flag below is warp-uniform and every stage is its own 32-wide loop of INDEPENDENT operations: with two
        // epilogue warps per scheduler the per-element form (load bias -> add -> relu -> ... one element after the other)
        // was a dependent chain of ~400 instructions per chunk and, not the stores, what made the epilogue longer than a
        // K = 512 mainloop (ablation: scripts/ablate_gemm.py).
        uint32_t r[32];
        if (!(p.debug & 4)) {
          tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + c0, r);
          tc::tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = i;
        }
        const int col0 = nt * BN + c0;
        if (col0 >= p.N || kb1 <= kb0) continue;          // warp-uniform
        float x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(r[i]);
        if (p.alpha != 1.f) {
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] *= p.alpha;
        }
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bv = tc::ld_shared_v4(bias_u32 + 4 * (c0 + 4 * j));      // broadcast: all lanes read the same 16 bytes
            x[4 * j] += bv.x; x[4 * j + 1] += bv.y; x[4 * j + 2] += bv.z; x[4 * j + 3] += bv.w;
          }
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = fmaxf(x[i], 0.f);
        }
        const int mask_ld = p.N >> 5;
        if (p.mask_in != nullptr) {       // fused FFN activation backward: a lane's 32 columns are one word of the relu/dropout bit plane
          const uint32_t min_w = row_ok ? __ldg(p.mask_in + (int64_t)row * mask_ld + (col0 >> 5)) : 0u;
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = ((min_w >> i) & 1u) ? x[i] * p.mask_scale : 0.f;
        }
        if (p.p_drop > 0.f) {
          // a lane's 32 columns take ONE bit-sliced keep word (drop_keep_word, common.cuh: 4 Philox4x32-7 calls) -- the
          // per-element form cost 8 Philox4x32-10 calls per chunk on the two epilogue warps of a scheduler
          const uint32_t kw = drop_keep_word(p.seed, p.offset, row, (p.N + 31) >> 5, col0 >> 5, keep16);
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = ((kw >> i) & 1u) ? x[i] * ks : 0.f;
        }
        if (p.mask_out != nullptr && row_ok) {
          uint32_t mout_w = 0u;
#pragma unroll
          for (int i = 0; i < 32; ++i) mout_w |= (uint32_t)(x[i] > 0.f) << i;
          p.mask_out[(int64_t)row * mask_ld + (col0 >> 5)] = mout_w;
        }
        if (p.colsum != nullptr) {        // column sums over the warp's 32 rows (bias gradient) take the UNROUNDED values
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = (row_ok && col0 + i < p.N) ? x[i] : 0.f;
          const float cs = warp_colsum32(f, lane);
          if (col0 + lane < p.N) atomicAdd(p.colsum + col0 + lane, cs);
        }
        if (p.round_out) {
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = tf32_rn(x[i]);
        }
        if (p.tma_store) {
          uint8_t* my_stage = out_stage + ((warp - 2) * C::kSlots + slot) * (32 * 128);
          if (C::kSlots > 1) slot ^= 1;
          const bool direct = p.epi_direct && !p.accumulate;
          // the smem tile written kSlots chunks ago has been read out by its TMA store
          if (!direct && lane == 0 && !(p.debug & 2)) { if (C::kSlots > 1) tc::tma_store_wait_read1(); else tc::tma_store_wait_read(); }
          __syncwarp();
          GTRACE(2, 2);
          // 128B swizzle: 16-byte chunk index XOR (row & 7) -- the layout the C tensor map expects
#pragma unroll
          for (int j = 0; j < 8; ++j)
            tc::st_shared_v4(tc::smem_u32(my_stage) + lane * 128 + ((j ^ (lane & 7)) << 4), x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
          if (direct) {
            // read the tile back row-wise: 8 lanes cover one 128-byte row segment (conflict-free: a quarter warp = one
            // swizzled row), 4 rows per instruction -> fully coalesced 16-byte global stores, nothing queued behind the
            // producer's TMA loads and no async-proxy fence
            __syncwarp();
            const int rr = lane >> 3, cc = lane & 7;
            const int64_t grow0 = (int64_t)b * p.M + mt * BM + q * 32;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int r2 = it * 4 + rr;
              const float4 o = tc::ld_shared_v4(tc::smem_u32(my_stage) + r2 * 128 + ((cc ^ (r2 & 7)) << 4));
              const int col = col0 + cc * 4;
              if (mt * BM + q * 32 + r2 < p.M && !(p.debug & 2)) {
                float* dst = p.c + (grow0 + r2) * p.ldc + col;        // batched C is [batch*M, N] here (tma_store precondition)
                if (col + 3 < p.N) *reinterpret_cast<float4*>(dst) = o;
                else {
                  if (col < p.N) dst[0] = o.x;
                  if (col + 1 < p.N) dst[1] = o.y;
                  if (col + 2 < p.N) dst[2] = o.z;
                }
              }
            }
            continue;
          }
          if (!(p.debug & 16)) tc::fence_proxy_async();      // (debug bit 16: timing experiment without the proxy fence)
          __syncwarp();
          if (lane == 0 && !(p.debug & 2)) {
            const int y = b * p.M + mt * BM + q * 32;
            if (p.accumulate) tc::tma_reduce_add_2d(&tma_c, my_stage, col0, y);
            else tc::tma_store_2d(&tma_c, my_stage, col0, y);
            tc::tma_store_commit();
          }
        } else if (row_ok && !(p.debug & 2)) {
          // register-direct stores: every lane owns one output row and writes its 32 consecutive columns itself -- no
          // shared-memory traffic at all.  256-bit stores = whole 32-byte sectors per lane.
          if (p.wide8 && !p.accumulate && col0 + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              st_global_v8(crow + col0 + 8 * j, make_float4(x[8 * j], x[8 * j + 1], x[8 * j + 2], x[8 * j + 3]),
                           make_float4(x[8 * j + 4], x[8 * j + 5], x[8 * j + 6], x[8 * j + 7]));
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c4 = col0 + 4 * j;
              if (c4 + 3 < p.N && ((p.ldc & 3) == 0)) {
                const float4 o = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
                if (p.accumulate) atomicAdd(reinterpret_cast<float4*>(crow + c4), o);
                else *reinterpret_cast<float4*>(crow + c4) = o;
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (c4 + e < p.N) {
                    if (p.accumulate) atomicAdd(crow + c4 + e, x[4 * j + e]);
                    else crow[c4 + e] = x[4 * j + e];
                  }
                }
              }
            }
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (TWO) tc::mbar_arrive_cluster(tc::mapa_u32(tc::smem_u32(tmem_empty + acc), 0));   // the leader's MMA warp waits for both CTAs
        else tc::mbar_arrive(tmem_empty + acc);
      }
      GTRACE(2, 3);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    GTRACE_END(2);
  }
  if (p.tma_store && warp >= 2 && lane == 0) tc::tma_store_wait_all();
  tc::tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) tc::cluster_sync();      // do not exit while the peer may still multicast / arrive here
  if (warp == 1) {
    if constexpr (TWO) tc::tmem_dealloc_pair<C::kTmemCols>(tmem_base);
    else tc::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn g_encode = nullptr;
std::once_flag g_encode_once;

void resolve_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
    g_encode = (EncodeFn)fn;
}

template <int BN, bool A_MN, bool B_MN, int CL, bool TWO = false>
int launch_cl(const pa_gemm_args& a, cudaStream_t st) {
  using C = Cfg<BN, TWO>;
  CUtensorMap ta, tb, tc_map;
  int rc;
  if (!A_MN) rc = pa_make_tmap_2d(&ta, a.a, (uint64_t)a.K, (uint64_t)(a.batch > 1 ? (int64_t)a.batch * a.a_batch_rows : a.M), (uint64_t)a.lda * 4, BK, BM);
  else rc = pa_make_tmap_2d(&ta, a.a, (uint64_t)a.M, (uint64_t)(a.batch > 1 ? (int64_t)a.batch * a.a_batch_rows : a.K), (uint64_t)a.lda * 4, 32, BK, true);
  if (rc) return rc;
  if (!B_MN) rc = pa_make_tmap_2d(&tb, a.b, (uint64_t)a.K, (uint64_t)(a.batch > 1 ? (int64_t)a.batch * a.b_batch_rows : a.N), (uint64_t)a.ldb * 4, BK, BN / CL);
  else rc = pa_make_tmap_2d(&tb, a.b, (uint64_t)a.N, (uint64_t)(a.batch > 1 ? (int64_t)a.batch * a.b_batch_rows : a.K), (uint64_t)a.ldb * 4, 32, BK, true);
  if (rc) return rc;
  GemmParams p{};
  p.c = a.c; p.ldc = a.ldc; p.c_batch_stride = a.c_batch_stride; p.bias = a.bias;
  p.M = a.M; p.N = a.N; p.K = a.K; p.batch = a.batch < 1 ? 1 : a.batch;
  p.a_batch_rows = (int)a.a_batch_rows; p.b_batch_rows = (int)a.b_batch_rows;
  p.split_k = a.split_k < 1 ? 1 : a.split_k; p.accumulate = a.accumulate;
  p.relu = a.relu; p.p_drop = a.p_drop; p.seed = a.seed; p.offset = a.offset; p.alpha = a.alpha; p.round_out = a.round_out;
  p.m_tiles = (a.M + BM - 1) / BM; p.n_tiles = (a.N + BN - 1) / BN;
  { const char* dbg = getenv("PLANK_B200_GEMM_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
  // TMA-store epilogue: C seen as [batch*M, N] rows of pitch ldc (clipped at the bounds by the TMA)
  p.tma_store = (a.ldc % 4 == 0) && (((uintptr_t)a.c & 15) == 0) &&
                (p.batch == 1 || (a.M % BM == 0 && a.c_batch_stride == (int64_t)a.M * a.ldc));
  // PLANK_B200_GEMM_EPI: 0 (default) = staged smem tile + TMA store, 1 = staged tile + coalesced st.global, 2 = register-direct
  // 256-bit stores (split-K keeps the TMA reduce-add).  The fp32-operand mainloop already keeps the shared-memory port busy
  // (64 KB written + read per 512-cycle k-block), so what an epilogue costs is its L1/smem datapath cycles per 128 KB tile:
  // ~2k for 0 (conflict-free st.shared + full-width TMA read), ~4k for 2 (every 256-bit store touches 32 different lines);
  // measured on the QKV projection (32768 x 1536 x 512, burst clocks): mainloop only 66.6 us, 0: 93.4, 1: 98.7, 2: 99.5.
  { const char* e = getenv("PLANK_B200_GEMM_EPI"); p.epi_direct = e == nullptr ? 0 : atoi(e); }
  p.wide8 = (a.ldc % 8 == 0) && (((uintptr_t)a.c & 31) == 0) && (a.c_batch_stride % 8 == 0);
  if (p.epi_direct == 2 && !p.accumulate) p.tma_store = 0;
  p.mask_out = a.mask_out; p.mask_in = a.mask_in; p.colsum = a.colsum; p.mask_scale = a.mask_scale;
  // (the bit-plane mask / column-sum stages run before the store stage: the fused FFN GEMMs take the TMA-store path too)
  if (p.tma_store) {
    rc = pa_make_tmap_2d(&tc_map, a.c, (uint64_t)a.N, (uint64_t)p.batch * a.M, (uint64_t)a.ldc * 4, 32, 32);
    if (rc) return rc;
  } else {
    tc_map = ta;
  }
  auto kern = gemm_tf32_kernel<BN, A_MN, B_MN, CL, TWO>;
  static SmemAttrCache attr;
  if ((rc = pa_set_max_smem(kern, C::kSmem, attr))) return rc;
  int tiles = ((p.m_tiles + CL - 1) / CL) * p.n_tiles * p.split_k * p.batch;      // cluster steps
  int grid = (tiles * CL < pa_num_sms() ? tiles * CL : (pa_num_sms() / CL) * CL);
  { const char* g = getenv("PLANK_B200_GEMM_GRID"); if (g != nullptr && atoi(g) > 0 && atoi(g) < grid) grid = atoi(g) / CL * CL; }   // experiments: fewer SMs
  if constexpr (CL == 1) {
    kern<<<grid, kThreads, C::kSmem, st>>>(ta, tb, tc_map, p);
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = C::kSmem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    PA_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tc_map, p));
  }
  PA_CHECK_LAUNCH();
  return PA_OK;
}

template <int BN, bool A_MN, bool B_MN>
int launch(const pa_gemm_args& a, cudaStream_t st) {
  // CTA pairs pay off once there are enough m tiles to pair up; tiny problems keep the single-CTA kernel.
  // PLANK_B200_GEMM_PAIR: 0 = single CTAs, 1 = pairs sharing B through TMA multicast (cta_group::1 MMAs),
  // 2 (default) = pairs issuing 2-SM UMMAs (cta_group::2, M = 256)
  const char* e = getenv("PLANK_B200_GEMM_PAIR");
  const int mode = e == nullptr ? 2 : atoi(e);
  const int m_tiles = (a.M + BM - 1) / BM;
  if (mode >= 1 && m_tiles >= 2 && m_tiles % 2 == 0)
    return mode >= 2 ? launch_cl<BN, A_MN, B_MN, 2, true>(a, st) : launch_cl<BN, A_MN, B_MN, 2, false>(a, st);
  return launch_cl<BN, A_MN, B_MN, 1>(a, st);
}

}  // namespace

#ifdef PA_GEMM_TRACE
extern "C" int pa_debug_gemm_trace(unsigned long long* out_host /*[3][2048]*/, int* n_host /*[3]*/) {
  PA_CUDA(cudaMemcpyFromSymbol(out_host, g_gemm_trace, sizeof(unsigned long long) * 3 * 2048));
  PA_CUDA(cudaMemcpyFromSymbol(n_host, g_gemm_trace_n, sizeof(int) * 3));
  return PA_OK;
}
#endif

int pa_make_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                    uint32_t box_inner, uint32_t box_outer, bool atom32) {
  std::call_once(g_encode_once, resolve_encode);
  if (g_encode == nullptr) { pa_set_error("cuTensorMapEncodeTiled unavailable (driver too old?)"); return PA_ERR_CUDA; }
  if (((uintptr_t)base & 15) || (pitch_bytes & 15)) { pa_set_error("TMA needs 16-byte aligned base and pitch"); return PA_ERR_ARG; }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { pa_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return PA_ERR_CUDA; }
  return PA_OK;
}

extern "C" int pa_gemm_tf32(const pa_gemm_args* a, void* stream) {
  PA_CHECK_ARG(a != nullptr && a->M > 0 && a->N > 0 && a->K > 0);
  PA_CHECK_ARG(a->lda % 4 == 0 && a->ldb % 4 == 0);
  PA_CHECK_ARG(!(a->batch > 1 && a->split_k > 1));
  // batched MN-major operands: the contraction must not run past a batch's rows (no zero fill between batches)
  PA_CHECK_ARG(!(a->batch > 1 && (a->a_mn || a->b_mn) && a->K % 32 != 0));
  PA_CHECK_ARG(!(a->split_k > 1 && !a->accumulate));
  PA_CHECK_ARG(!(a->accumulate && (a->bias != nullptr || a->relu || a->p_drop > 0.f)));
  if (a->mask_out != nullptr || a->mask_in != nullptr || a->colsum != nullptr)
    PA_CHECK_ARG(a->batch <= 1 && a->N % 32 == 0 && !a->accumulate && a->split_k <= 1);
  cudaStream_t st = (cudaStream_t)stream;
  const bool wide = a->N > 128 && (a->N % 256 == 0 || a->N > 1024);
  if (!a->a_mn && !a->b_mn) return wide ? launch<256, false, false>(*a, st) : launch<128, false, false>(*a, st);
  if (a->a_mn && a->b_mn) return (a->N % 256 == 0) ? launch<256, true, true>(*a, st) : launch<128, true, true>(*a, st);
  if (!a->a_mn && a->b_mn) return (a->N % 256 == 0) ? launch<256, false, true>(*a, st) : launch<128, false, true>(*a, st);
  pa_set_error("pa_gemm_tf32: mixed operand majors are not built (a_mn=%d b_mn=%d)", a->a_mn, a->b_mn);
  return PA_ERR_UNSUPPORTED;
}
