// K3/K4/K5 on the 5th-gen tensor cores: flash-style attention forward, TF32 operands, FP32 accumulate.
//
//   S = Q K^T   : tcgen05.mma kind::tf32, A = Q tile (smem, K-major, 128B swizzle), B = K tile (smem, K-major)
//   P = softmax : one thread per query row reads its S row from TMEM (tcgen05.ld), applies scale +
//                 key-padding / causal masks + online max/sum + Philox dropout in registers, rounds
//                 P to nearest TF32 and writes it back over S in TMEM (tcgen05.st)
//   O_j = P V   : tcgen05.mma with A = P straight from TMEM, B = V tile (smem, MN-major, 128B_BASE32B)
//   O += O_j    : rescaled accumulation in registers (no TMEM read-modify-write of O)
//
// Persistent CTAs (one per SM), 320 threads: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
// warps 2..9 = two softmax/accumulate warpgroups (each: 128 threads = 128 query rows = 128 TMEM lanes,
// owning one half of the key columns of every tile and one half of the head-dim columns of O).
// Pipelines: K ring (3 stages) and V ring (2 stages) fed by TMA; S/P and O double-buffered in TMEM so
// QK^T of tile j+1 and P V of tile j overlap the softmax of tile j.
//   Reference: torch nn/functional.py multi_head_attention_forward as reached from ref models.py:206,212.
#include "common.cuh"
#include "tc_common.cuh"

#include <stdlib.h>

// Wait-time profile of block 0 (cycles spent in each mbarrier / named-barrier wait, per role); filled only when
// PLANK_B200_ATTN_DEBUG has bit 64 set, read back with pa_debug_attn_prof().
__device__ unsigned long long g_attn_prof[32];

// Build with PLANK_B200_NVCC_FLAGS=-DPA_ATTN_TRACE to record an event timeline of CTA 0 (clock64 per event and role:
// 0 TMA producer, 1 MMA issuer, 2 softmax warp 2, 3 softmax warp 6), read back with pa_debug_attn_trace().
#ifdef PA_ATTN_TRACE
constexpr int kTraceLen = 1024;
__device__ unsigned long long g_attn_trace[4][kTraceLen];
__device__ int g_attn_trace_n[4];
#define TRACE(role, ev)                                                                          \
  do {                                                                                           \
    if (trace_on && trace_n < kTraceLen) g_attn_trace[role][trace_n++] = ((unsigned long long)(ev) << 56) | (clock64() & 0xffffffffffffffull); \
  } while (0)
#define TRACE_DECL const bool trace_on = (p.debug & 1024) && blockIdx.x == 0; int trace_n = 0
#define TRACE_END(role) do { if (trace_on) g_attn_trace_n[role] = trace_n; } while (0)
#else
#define TRACE(role, ev) do {} while (0)
#define TRACE_DECL do {} while (0)
#define TRACE_END(role) do {} while (0)
#endif

namespace {

#define TIMED_WAIT(slot, stmt)                                             \
  do {                                                                     \
    if (prof_on) { long long t0__ = clock64(); stmt; prof[slot] += clock64() - t0__; } \
    else { stmt; }                                                         \
  } while (0)

constexpr int BQ = 128, BKV = 128;
constexpr int kThreads = 320;      // warp 0 TMA, warp 1 MMA, warps 2..9 softmax (two warpgroups)
constexpr int kKStages = 3, kVStages = 2;

template <int DH> struct Cfg {
  static constexpr int kChunks = DH / 32;                 // 32-float (128 B) column chunks per row
  static constexpr int kTileBytes = kChunks * BQ * 128;   // one Q / K / V tile
  static constexpr int kOffQ = 0;
  static constexpr int kOffK = kTileBytes;
  static constexpr int kOffV = kOffK + kKStages * kTileBytes;
  static constexpr int kOffXch = kOffV + kVStages * kTileBytes;    // [2 bufs x 2 halves + 2][128] floats: max / sum exchange
  static constexpr int kOffFlag = kOffXch + 6 * BQ * 4;             // [2 items][64]: all 32 keys of the group valid
  static constexpr int kOffBar = kOffFlag + 512;
  static constexpr int kOffBias = kOffBar + 256;                    // [2 items][LkPad] additive key bias, sized at launch
  static constexpr int kSmemFixed = kOffBias + 1024;                // + 2 * LkPad * 4
  static constexpr int kTmemCols = 512;
  static constexpr int kColS = 0;                          // 2 x 128 columns  S / P
  static constexpr int kColO = 256;                        // 2 x DH columns   O_j
};

struct Params {
  float* o; int64_t ldo; float* lse; const uint8_t* kpm;
  int B, H, Lq, Lk, causal, round_out;
  float scale_log2;   // scale * log2(e)
  float p_drop; const uint32_t* drop_rows; int LkW;
  int q_tiles, items;
  const int32_t* kv_len;   // optional [B]: keys >= kv_len[b] are all PAD -> key tiles beyond it are skipped
  int wide_st;        // o is 32-byte aligned with a row pitch that is a multiple of 8 floats: 256-bit stores
  int LkPad;          // keys rounded up to the key tile (per-item bias table length)
  int debug;          // ablation bits for timing experiments (PLANK_B200_ATTN_DEBUG); 0 in production
};

template <int DH>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const Params p) {
  using C = Cfg<DH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* bias_s = reinterpret_cast<float*>(smem + C::kOffBias);
  float* xch_s = reinterpret_cast<float*>(smem + C::kOffXch);
  const uint32_t xch_u32 = tc::smem_u32(xch_s);
  int* flag_s = reinterpret_cast<int*>(smem + C::kOffFlag);      // [2 bufs][4 warps]: all 32 keys valid
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;            // [<= 3]
  uint64_t* k_empty = bars + 5;           // [<= 3]
  uint64_t* v_full = bars + 8;            // [<= 3]
  uint64_t* v_empty = bars + 11;          // [<= 3]
  uint64_t* s_full = bars + 14;           // [2]
  uint64_t* p_full = bars + 16;           // [2]
  uint64_t* o_full = bars + 18;           // [2]
  uint64_t* o_empty = bars + 20;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool prof_on = (p.debug & 64) && blockIdx.x == 0;
  long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long t_kernel0 = clock64();
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tm_q); tc::tma_prefetch_desc(&tm_k); tc::tma_prefetch_desc(&tm_v);
    tc::mbar_init(q_full, 1); tc::mbar_init(q_empty, 1);
    for (int s = 0; s < kKStages; ++s) { tc::mbar_init(k_full + s, 1); tc::mbar_init(k_empty + s, 1); }
    for (int s = 0; s < kVStages; ++s) { tc::mbar_init(v_full + s, 1); tc::mbar_init(v_empty + s, 1); }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(s_full + s, 1); tc::mbar_init(p_full + s, 8);
      tc::mbar_init(o_full + s, 1); tc::mbar_init(o_empty + s, 8);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<C::kTmemCols>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto item_coords = [&](int item, int& b, int& h, int& q0, int& ntiles) {
    int bh = item / p.q_tiles;
    // rotate the tile index with bh: a CTA's successive items (stride = grid size, a multiple of 4) would otherwise all
    // have the SAME tile index, and tiles differ in work (causal rows, key tiles skipped through kv_len)
    int qt = (item % p.q_tiles + bh) % p.q_tiles;
    h = bh % p.H; b = bh / p.H;
    q0 = qt * BQ;
    const int lk = p.kv_len != nullptr ? min(p.Lk, max(1, __ldg(p.kv_len + b))) : p.Lk;
    int all = (lk + BKV - 1) / BKV;
    ntiles = p.causal ? min(all, qt + 1) : all;
  };

  if (warp == 0) {
    // ======================================= TMA producer =======================================
    if (lane == 0) {
      TRACE_DECL;
      uint32_t kc = 0, vc = 0, ic = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++ic) {
        int b, h, q0, n;
        item_coords(item, b, h, q0, n);
        TIMED_WAIT(0, tc::mbar_wait(q_empty, (ic & 1) ^ 1));
        TRACE(0, 0);
        tc::mbar_arrive_expect_tx(q_full, C::kTileBytes);
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c) tc::tma_load_2d(smem + C::kOffQ + c * (BQ * 128), &tm_q, h * DH + c * 32, b * p.Lq + q0, q_full);
        for (int j = 0; j < n; ++j) {
          const int ks = kc % kKStages;
          TIMED_WAIT(1, tc::mbar_wait(k_empty + ks, ((kc / kKStages) & 1) ^ 1));
          TRACE(0, 2);
          tc::mbar_arrive_expect_tx(k_full + ks, C::kTileBytes);
#pragma unroll
          for (int c = 0; c < C::kChunks; ++c)
            tc::tma_load_2d(smem + C::kOffK + ks * C::kTileBytes + c * (BKV * 128), &tm_k, h * DH + c * 32, b * p.Lk + j * BKV, k_full + ks);
          ++kc;
          const int vs = vc % kVStages;
          TIMED_WAIT(2, tc::mbar_wait(v_empty + vs, ((vc / kVStages) & 1) ^ 1));
          TRACE(0, 3);
          tc::mbar_arrive_expect_tx(v_full + vs, C::kTileBytes);
#pragma unroll
          for (int c = 0; c < C::kChunks; ++c)
            tc::tma_load_2d(smem + C::kOffV + vs * C::kTileBytes + c * (BKV * 128), &tm_v, h * DH + c * 32, b * p.Lk + j * BKV, v_full + vs);
          ++vc;
        }
      }
      TRACE_END(0);
    }
  } else if (warp == 1) {
    // ======================================= MMA issuer =========================================
    // The whole warp walks the loop (warp-uniform control flow and addresses); one elected lane issues.
    {
      TRACE_DECL;
      constexpr uint32_t idesc_qk = tc::make_idesc_tf32(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_pv = tc::make_idesc_tf32(BQ, DH, 0, 1);
      uint32_t kc = 0, vc = 0, st = 0 /*QK tiles issued*/, pt = 0 /*PV tiles issued*/, ic = 0;
      const uint32_t sq = tc::smem_u32(smem + C::kOffQ);
      auto issue_qk = [&]() {
        const int ks = kc % kKStages;
        TIMED_WAIT(1, tc::mbar_wait(k_full + ks, (kc / kKStages) & 1));
        TRACE(1, 1);
        tc::tc_fence_after();
        const uint32_t sk = tc::smem_u32(smem + C::kOffK + ks * C::kTileBytes);
        const uint32_t d_tmem = tmem_base + C::kColS + (st & 1) * BKV;
        if (tc::elect_one()) {
          if (!(p.debug & 512))
#pragma unroll
          for (int c = 0; c < C::kChunks; ++c) {
            const uint64_t dq = tc::make_smem_desc(sq + c * (BQ * 128), 16, 1024);
            const uint64_t dk = tc::make_smem_desc(sk + c * (BKV * 128), 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc::mma_tf32_ss(d_tmem, tc::desc_advance(dq, k * 32), tc::desc_advance(dk, k * 32), idesc_qk, (c > 0 || k > 0) ? 1u : 0u);
          }
          tc::tc_commit(k_empty + ks);
          tc::tc_commit(s_full + (st & 1));
        }
        __syncwarp();
        TRACE(1, 2);
        ++kc; ++st;
      };
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++ic) {
        int b, h, q0, n;
        item_coords(item, b, h, q0, n);
        TIMED_WAIT(0, tc::mbar_wait(q_full, ic & 1));
        TRACE(1, 0);
        tc::tc_fence_after();
        issue_qk();
        if (n > 1) issue_qk();
        if (n <= 2) { if (tc::elect_one()) tc::tc_commit(q_empty); __syncwarp(); }
        for (int j = 0; j < n; ++j) {
          const int buf = pt & 1;
          TIMED_WAIT(2, tc::mbar_wait(p_full + buf, (pt >> 1) & 1));
          TRACE(1, 3);
          TIMED_WAIT(3, tc::mbar_wait(o_empty + buf, ((pt >> 1) & 1) ^ 1));
          TRACE(1, 4);
          const int vs = vc % kVStages;
          TIMED_WAIT(4, tc::mbar_wait(v_full + vs, (vc / kVStages) & 1));
          TRACE(1, 5);
          tc::tc_fence_after();
          const uint32_t sv = tc::smem_u32(smem + C::kOffV + vs * C::kTileBytes);
          // V tile: kChunks MN blocks (32 head-dim columns each) of 128 key rows x 128 B; 4-row swizzle atoms
          const uint64_t dv = tc::make_smem_desc(sv, BKV * 128, 512, tc::kLayoutSw128Base32);
          const uint32_t a_tmem = tmem_base + C::kColS + buf * BKV;
          const uint32_t d_tmem = tmem_base + C::kColO + buf * DH;
          if (tc::elect_one()) {
            if (!(p.debug & 256))
#pragma unroll
            for (int k = 0; k < BKV / 8; ++k)
              tc::mma_tf32_ts(d_tmem, a_tmem + k * 8, tc::desc_advance(dv, k * 1024), idesc_pv, k > 0 ? 1u : 0u);
            tc::tc_commit(v_empty + vs);
            tc::tc_commit(o_full + buf);
          }
          __syncwarp();
          TRACE(1, 6);
          ++vc; ++pt;
          if (j + 2 < n) {
            issue_qk();
            if (j + 3 == n) { if (tc::elect_one()) tc::tc_commit(q_empty); __syncwarp(); }     // that was the last QK^T of this item
          }
        }
      }
      TRACE_END(1);
    }
  } else {
    // ============================ softmax / accumulate: two warpgroups ============================
    // Warps 2..5 and 6..9 share TMEM lane quarters pairwise; warpgroup `half` owns key columns
    // [64*half, 64*half+64) of every 128-key tile and head-dim columns [32*half, 32*half+32) of O.
    // Row maxima are exchanged through smem once per tile; each half keeps a partial row sum.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int tid = threadIdx.x - 64;                    // 0..255
    const float ks = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;
    constexpr int HC = BKV / 2;                          // score columns per half
    const bool has_o = half * 32 < DH;
    uint32_t sc = 0, oc = 0;
#ifdef PA_ATTN_TRACE
    const bool trace_on = (p.debug & 1024) && blockIdx.x == 0 && lane == 0 && (warp == 2 || warp == 6);
    int trace_n = 0;
    const int trole = warp == 2 ? 2 : 3;
#endif
    // The additive key bias and the "tile needs no masking" flags of a whole item live in a per-item smem table (two
    // tables alternate; the next item's is written before the current item's last accumulate / read-out), so a tile
    // costs one named barrier (the row-max exchange) only.
    auto build_table = [&](int item, uint32_t parity) {  // additive key bias + group flags of one item
      int b_, h_, q0_, n_;
      item_coords(item, b_, h_, q0_, n_);
      float* bt = bias_s + (parity & 1) * p.LkPad;
      int* ft = flag_s + (parity & 1) * 64;
      for (int k = tid; k < n_ * BKV; k += 256) {
        const bool ok = k < p.Lk && !(p.kpm != nullptr && p.kpm[(int64_t)b_ * p.Lk + k]);
        bt[k] = ok ? 0.f : -INFINITY;
        const bool all_ok = __all_sync(0xffffffffu, ok);
        if (lane == 0) ft[k >> 5] = all_ok ? 1 : 0;
      }
    };
    if ((int)blockIdx.x < p.items) build_table(blockIdx.x, 0);
    uint32_t itc = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++itc) {
      int b, h, q0, n;
      item_coords(item, b, h, q0, n);
      const int qi = q0 + row;
      const int64_t row_global = ((int64_t)(b * p.H + h) * p.Lq + qi);
      const float* bias_it = bias_s + (itc & 1) * p.LkPad;
      const uint32_t bias_u32 = tc::smem_u32(bias_it);
      const int* flag_it = flag_s + (itc & 1) * 64;
      float m_run = -INFINITY, l_run = 0.f, corr_prev = 1.f;
      float o_acc[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) o_acc[c] = 0.f;

      auto accumulate_o = [&]() {
        const int obuf = oc & 1;
        TIMED_WAIT(3, tc::mbar_wait(o_full + obuf, (oc >> 1) & 1));
        TRACE(trole, 8);
        tc::tc_fence_after();
        if (has_o && !(p.debug & 4)) {
          uint32_t r[32];
          tc::tmem_ld_32x32(tmem_base + lane_addr + C::kColO + obuf * DH + half * 32, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) o_acc[c] = o_acc[c] * corr_prev + __uint_as_float(r[c]);
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(o_empty + obuf);
        TRACE(trole, 9);
        ++oc;
      };

      // the dropout words of a tile are requested one tile ahead so that their latency never sits on the per-tile chain
      auto load_w = [&](int k0n, uint32_t (&wd)[HC / 32]) {
#pragma unroll
        for (int i = 0; i < HC / 32; ++i) {
          const int wi = ((k0n + half * HC) >> 5) + i;
          wd[i] = (qi < p.Lq && wi < p.LkW) ? __ldg(p.drop_rows + row_global * p.LkW + wi) : 0u;
        }
      };
      uint32_t w_pref[HC / 32];
      if (p.p_drop > 0.f) load_w(0, w_pref);
      TIMED_WAIT(0, asm volatile("bar.sync 1, 256;" ::: "memory"));      // the item's bias table is complete
      for (int j = 0; j < n; ++j) {
        const int buf = sc & 1;
        const int k0 = j * BKV;
        TRACE(trole, 0);
        uint32_t w[HC / 32];                   // keep-bits of this row's 64 keys (precomputed Philox bit plane)
#pragma unroll
        for (int i = 0; i < HC / 32; ++i) w[i] = w_pref[i];
        if (p.p_drop > 0.f && j + 1 < n) load_w(k0 + BKV, w_pref);
        TRACE(trole, 1);
        TIMED_WAIT(1, tc::mbar_wait(s_full + buf, (sc >> 1) & 1));
        TRACE(trole, 2);
        tc::tc_fence_after();
        const uint32_t s_addr = tmem_base + lane_addr + C::kColS + buf * BKV + half * HC;
        float s[HC];
        {
          uint32_t r0[32], r1[32];               // both 32-column chunks in flight before the single wait
          if (!(p.debug & 32)) {
            tc::tmem_ld_32x32(s_addr, r0);
            tc::tmem_ld_32x32(s_addr + 32, r1);
            tc::tmem_ld_wait();
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) { r0[c] = 0x3f800000u + c; r1[c] = 0x3f800000u; }
          }
#pragma unroll
          for (int c = 0; c < 32; ++c) { s[c] = __uint_as_float(r0[c]); s[32 + c] = __uint_as_float(r1[c]); }
        }
        TRACE(trole, 3);
        const bool diag = p.causal && (k0 + BKV - 1 > q0);
        // fast path: every key of the tile is valid and no causal boundary crosses it -> no per-element masking
        const int fg = k0 >> 5;
        const bool clean = !diag && (flag_it[fg] & flag_it[fg + 1] & flag_it[fg + 2] & flag_it[fg + 3]);
        float mx = -INFINITY;                  // running in the RAW score domain; scale folded into the exp2 FFMA
        if (clean) {
#pragma unroll
          for (int c = 0; c < HC; ++c) mx = fmaxf(mx, s[c]);
        } else {
#pragma unroll
          for (int c = 0; c < HC; ++c) {
            float v = s[c] + tc::ld_shared_f32(bias_u32 + 4 * (k0 + half * HC + c));
            if (diag && k0 + half * HC + c > qi) v = -INFINITY;
            s[c] = v;
            mx = fmaxf(mx, v);
          }
        }
        if (!(p.debug & 8)) {
          tc::st_shared_f32(xch_u32 + 4 * ((buf * 2 + half) * BQ + row), mx);          // exchange the half-row maxima (explicit STS/LDS)
          TIMED_WAIT(2, asm volatile("bar.sync 1, 256;" ::: "memory"));
          mx = fmaxf(mx, tc::ld_shared_f32(xch_u32 + 4 * ((buf * 2 + (half ^ 1)) * BQ + row)));
        }
        TRACE(trole, 4);
        const float m_new = fmaxf(m_run, mx);
        const float m_safe = m_new == -INFINITY ? 0.f : m_new;
        const float corr = fast_exp2((m_run - m_safe) * p.scale_log2);
        const float neg_ms = -m_safe * p.scale_log2;
        float rs = 0.f;
#pragma unroll
        for (int c = 0; c < HC; ++c) { s[c] = (p.debug & 1) ? fmaf(s[c], p.scale_log2, neg_ms) : fast_exp2(fmaf(s[c], p.scale_log2, neg_ms)); rs += s[c]; }
        TRACE(trole, 5);
        l_run = l_run * corr + rs;                        // partial sum over this half's columns
        m_run = m_new;
        if (p.p_drop > 0.f) {
#pragma unroll
          for (int c = 0; c < HC; ++c) s[c] = ((w[c >> 5] >> (c & 31)) & 1u) ? s[c] : 0.f;   // x 1/(1-p) folded into the final scale
        }
        if (!(p.debug & 2)) {
#pragma unroll
        for (int c0 = 0; c0 < HC; c0 += 32) {
          uint32_t r[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) r[c] = tf32_rn_finite_bits(s[c0 + c]);      // P is in [0, 1]
          tc::tmem_st_32x32(s_addr + c0, r);
        }
        tc::tmem_st_wait();
        } else if (rs == 123.f) { p.o[0] = s[1]; }
        TRACE(trole, 6);
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(p_full + buf);
        TRACE(trole, 7);
        ++sc;
        if (j > 0) accumulate_o();        // O_{j-1}: its P V overlapped this tile's softmax
        corr_prev = corr;
      }
      // the next item's table is built here: its key-padding loads overlap the wait for this item's last P V
      if (item + (int)gridDim.x < p.items) build_table(item + gridDim.x, itc + 1);
      accumulate_o();
      // total row sum = sum of the two halves' partial sums
      xch_s[(4 + half) * BQ + row] = l_run;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float l_tot = l_run + xch_s[(4 + (half ^ 1)) * BQ + row];
      if (qi < p.Lq) {
        const float inv = l_tot > 0.f ? ks / l_tot : 0.f;
        if (has_o) {
          float* op = p.o + ((int64_t)b * p.Lq + qi) * p.ldo + h * DH + half * 32;
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            float4 v = make_float4(o_acc[c] * inv, o_acc[c + 1] * inv, o_acc[c + 2] * inv, o_acc[c + 3] * inv);
            float4 w = make_float4(o_acc[c + 4] * inv, o_acc[c + 5] * inv, o_acc[c + 6] * inv, o_acc[c + 7] * inv);
            if (p.round_out) { v = tf32_rn4(v); w = tf32_rn4(w); }
            if (p.wide_st) st_global_v8(op + c, v, w);
            else { *reinterpret_cast<float4*>(op + c) = v; *reinterpret_cast<float4*>(op + c + 4) = w; }
          }
        }
        if (p.lse != nullptr && half == 0)
          p.lse[((int64_t)b * p.H + h) * p.Lq + qi] = l_tot > 0.f ? (m_run * p.scale_log2 + log2f(l_tot)) * 0.6931471805599453f : -INFINITY;
      }
      TRACE(trole, 10);   // (the next item's bias-table barrier also orders the reuse of xch_s)
    }
#ifdef PA_ATTN_TRACE
    if (trace_on) g_attn_trace_n[trole] = trace_n;
#endif
  }
  if (prof_on && lane == 0 && warp <= 2) {
    for (int i = 0; i < 8; ++i) g_attn_prof[warp * 8 + i] = (unsigned long long)prof[i];
    if (warp == 2) g_attn_prof[31] = (unsigned long long)(clock64() - t_kernel0);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<C::kTmemCols>(tmem_base);
}

template <int DH>
int launch(const pa_attn_fwd_args& a, cudaStream_t st) {
  using C = Cfg<DH>;
  CUtensorMap tq, tk, tv;
  int rc = pa_make_tmap_2d(&tq, a.q, (uint64_t)a.H * DH, (uint64_t)a.B * a.Lq, (uint64_t)a.ldq * 4, 32, BQ);
  if (rc) return rc;
  rc = pa_make_tmap_2d(&tk, a.k, (uint64_t)a.H * DH, (uint64_t)a.B * a.Lk, (uint64_t)a.ldk * 4, 32, BKV);
  if (rc) return rc;
  rc = pa_make_tmap_2d(&tv, a.v, (uint64_t)a.H * DH, (uint64_t)a.B * a.Lk, (uint64_t)a.ldv * 4, 32, BKV, true);
  if (rc) return rc;
  Params p{};
  p.o = a.o; p.ldo = a.ldo; p.lse = a.lse; p.kpm = a.kpm;
  p.B = a.B; p.H = a.H; p.Lq = a.Lq; p.Lk = a.Lk; p.causal = a.causal; p.round_out = a.round_out;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.p_drop = a.p_drop; p.drop_rows = a.drop_rows; p.LkW = (a.Lk + 31) / 32;
  if (a.p_drop > 0.f && a.drop_rows == nullptr) {
    pa_set_error("pa_attn_fwd (tc): p_drop > 0 needs drop_rows from pa_dropout_mask");
    return PA_ERR_ARG;
  }
  { const char* dbg = getenv("PLANK_B200_ATTN_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
  p.q_tiles = (a.Lq + BQ - 1) / BQ;
  p.items = p.q_tiles * a.H * a.B;
  p.LkPad = (a.Lk + BKV - 1) / BKV * BKV;
  p.wide_st = (((uintptr_t)a.o & 31) == 0 && a.ldo % 8 == 0) ? 1 : 0;
  p.kv_len = a.kpm != nullptr ? a.kv_len : nullptr;
  const int smem_bytes = C::kSmemFixed + 2 * p.LkPad * 4;
  if (smem_bytes > 227 * 1024 || p.LkPad > 2048) { pa_set_error("pa_attn_fwd (tc): Lk = %d too long for the bias table", a.Lk); return PA_ERR_UNSUPPORTED; }
  auto kern = attn_fwd_tc_kernel<DH>;
  static SmemAttrCache attr;
  if (int rc_attr = pa_set_max_smem(kern, smem_bytes, attr)) return rc_attr;
  int grid = p.items < pa_num_sms() ? p.items : pa_num_sms();
  kern<<<grid, kThreads, smem_bytes, st>>>(tq, tk, tv, p);
  PA_CHECK_LAUNCH();
  return PA_OK;
}

}  // namespace

// The tensor maps address q/k/v as 2-D [B*L rows, H*dh columns] arrays (row pitch ld) from their own base
// pointers, so the packed in-projection output [B,L,3d] is consumed in place.
bool pa_attn_fwd_pp_ok(const pa_attn_fwd_args* a);              // attn_pp.cu: two query tiles per CTA in ping-pong
int pa_attn_fwd_pp(const pa_attn_fwd_args* a, void* stream);

int pa_attn_fwd_tc(const pa_attn_fwd_args* a, void* stream) {
  if (pa_attn_fwd_pp_ok(a)) return pa_attn_fwd_pp(a, stream);
  switch (a->dh) {
    case 32: return launch<32>(*a, (cudaStream_t)stream);
    case 64: return launch<64>(*a, (cudaStream_t)stream);
    default: pa_set_error("pa_attn_fwd (tc): head dim %d unsupported (32, 64)", a->dh); return PA_ERR_UNSUPPORTED;
  }
}

#ifdef PA_ATTN_TRACE
extern "C" int pa_debug_attn_trace(unsigned long long* out_host /*[4][1024]*/, int* n_host /*[4]*/) {
  PA_CUDA(cudaMemcpyFromSymbol(out_host, g_attn_trace, sizeof(unsigned long long) * 4 * kTraceLen));
  PA_CUDA(cudaMemcpyFromSymbol(n_host, g_attn_trace_n, sizeof(int) * 4));
  return PA_OK;
}
#endif

// debug only (not part of the documented ABI surface used by the model): copy the wait-time profile to the host
extern "C" int pa_debug_attn_prof(unsigned long long* out32_host) {
  PA_CUDA(cudaMemcpyFromSymbol(out32_host, g_attn_prof, sizeof(unsigned long long) * 32));
  return PA_OK;
}
