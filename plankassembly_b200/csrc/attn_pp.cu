// K3/K4/K5 forward, second organisation: TWO query tiles per CTA in ping-pong (one softmax warpgroup per tile).
//
// attn_tc.cu splits the key columns of ONE query tile between its two softmax warpgroups: both run the same phase of the
// same tile at the same time (row-max exchange through a named barrier every tile), the MUFU and the FMA pipes are used in
// bursts, and QK^T / softmax / PV of a tile are one dependent chain -- the timeline (scripts/trace_attn.py) shows ~3.9 k
// cycles per 128x128 tile against ~1.1 k issue cycles and 1 k MUFU cycles.  Here a CTA owns a PAIR of query tiles of one
// (batch, head):
//   * warpgroup 1+g (warps 4+4g .. 7+4g) owns query tile g: one thread = one query row = one TMEM lane, a whole row of a
//     64-key step in registers -- no cross-group exchange, no per-tile named barrier;
//   * K and V tiles (128 keys) are loaded ONCE for both query tiles and consumed as two 64-key steps; S is double-buffered
//     per query tile (S_g[0] | S_g[1] = the two halves), so QK^T of step t+2 is issued right behind P V of step t and the
//     softmax of a tile never waits for its next S; the tensor pipe alternates  PV_A QK_A | PV_B QK_B | ...  and the MMAs of
//     one tile run under the softmax of the other;
//   * the stationary Q tiles are copied into TMEM by their softmax warpgroups (TMEM A-operand: 32 instead of 48 cycles per
//     128x64x8 MMA) and their shared-memory tiles are released at once for the next item's loads;
//   * O accumulates in TMEM across key steps (tcgen05.mma accumulate) instead of a per-tile read-out into registers;
//     it is rescaled there only when a row's running maximum grows by more than 2^8 (exact: probabilities are then taken
//     relative to a stale maximum and may exceed 1, which fp32 / TF32 hold without loss), so the per-tile
//     O handshake disappears;
//   * P overwrites S in place (the in-order tensor pipe makes PV_g(t) read it before QK_g(t+2) overwrites the buffer).
// TMEM: S_A | S_B (2 x 2 x 64 columns), O_A | O_B (2 x DH), Q_A | Q_B (2 x DH).  Shared memory: Q_A, Q_B, two K stages, two
// V stages (32 KB each at DH = 64) + the per-item key-bias tables.
// Used when the number of query tiles is even; attn_tc.cu keeps odd counts (a single 128-row tile: T = 128).
//   Reference: torch nn/functional.py multi_head_attention_forward as reached from ref models.py:206,212.
#include "common.cuh"
#include "tc_common.cuh"

#include <stdlib.h>

#ifdef PA_ATTN_TRACE
constexpr int kTraceLen2 = 1024;
__device__ unsigned long long g_attn2_trace[4][kTraceLen2];
__device__ int g_attn2_trace_n[4];
#define TRACE(role, ev)                                                                          \
  do {                                                                                           \
    if (trace_on && trace_n < kTraceLen2) g_attn2_trace[role][trace_n++] = ((unsigned long long)(ev) << 56) | (clock64() & 0xffffffffffffffull); \
  } while (0)
#else
#define TRACE(role, ev) do {} while (0)
#endif

namespace {

constexpr int BQ = 128, BKV = 128;
constexpr int BS = 64;             // keys per softmax step: a K / V tile (BKV keys) is consumed as two halves
constexpr int kThreads = 384;      // warpgroup 0: warp 0 TMA, warp 1 MMA (2, 3 idle); warpgroups 1 / 2: softmax of tile A / tile B
// A softmax thread keeps a whole 128-key score row in registers: the warpgroups trade registers at kernel start
// (setmaxnreg: 4 warps x 80 + 8 warps x 208, within the 12 x 32 x 168 the CTA is launched with); with a uniform 168 the row spilled to local memory.
constexpr int kRegsLow = 80, kRegsHigh = 208;
constexpr int kKStages = 2, kVStages = 2;
constexpr float kRescaleLog2 = 8.f;      // O is rescaled when the running maximum grows by more than this (log2 domain)

template <int DH> struct Cfg {
  static constexpr int kChunks = DH / 32;                 // 32-float (128 B) column chunks per row
  static constexpr int kTileBytes = kChunks * BQ * 128;   // one Q / K / V tile
  static constexpr int kOffQ = 0;                         // two tiles
  static constexpr int kOffK = 2 * kTileBytes;
  static constexpr int kOffV = kOffK + kKStages * kTileBytes;
  static constexpr int kOffFlag = kOffV + kVStages * kTileBytes;    // [2 items][64]: all 32 keys of the group valid
  static constexpr int kOffBar = kOffFlag + 512;
  static constexpr int kOffBias = kOffBar + 256;                    // [2 items][LkPad] additive key bias, sized at launch
  static constexpr int kSmemFixed = kOffBias + 1024;                // + 2 * LkPad * 4
  static constexpr int kTmemCols = 512;
  static constexpr int kColS = 0;                          // + g * 128
  static constexpr int kColO = 256;                        // + g * DH
  static constexpr int kColQ = 384;                        // + g * DH: the stationary Q tiles as TMEM A-operands of QK^T
};

struct Params {
  float* o; int64_t ldo; float* lse; const uint8_t* kpm;
  int B, H, Lq, Lk, causal, round_out;
  float scale_log2;   // scale * log2(e)
  float p_drop; const uint32_t* drop_rows; int LkW;
  int q_pairs, items;
  const int32_t* kv_len;
  int wide_st;
  int LkPad;
  int debug;
};

template <int DH, bool DROP>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_pp_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const Params p) {
  using C = Cfg<DH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* bias_s = reinterpret_cast<float*>(smem + C::kOffBias);
  int* flag_s = reinterpret_cast<int*>(smem + C::kOffFlag);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;            // [2]
  uint64_t* k_empty = bars + 4;           // [2]
  uint64_t* v_full = bars + 6;            // [2]
  uint64_t* v_empty = bars + 8;           // [2]
  uint64_t* s_full = bars + 10;           // [2 groups][2 bufs]  S_g(t) complete
  uint64_t* p_full = bars + 14;           // [2 groups][2 bufs]  P_g(t) stored (and O_g rescaled if it had to be)
  uint64_t* o_final = bars + 18;          // [2 groups]  last PV_g of the item complete
  uint64_t* o_read = bars + 20;           // [2 groups]  O_g read out: the next item's PV_g(0) may overwrite it
  uint64_t* pv_done = bars + 22;          // [2 groups]  PV_g(t) complete (waited for only before a rescale of O_g)
  uint64_t* qt_full = bars + 24;          // [2 groups]  Q_g copied into TMEM by its softmax warpgroup
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tm_q); tc::tma_prefetch_desc(&tm_k); tc::tma_prefetch_desc(&tm_v);
    tc::mbar_init(q_full, 1); tc::mbar_init(q_empty, 8);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(k_full + s, 1); tc::mbar_init(k_empty + s, 1);
      tc::mbar_init(v_full + s, 1); tc::mbar_init(v_empty + s, 1);
      tc::mbar_init(s_full + 2 * s, 1); tc::mbar_init(s_full + 2 * s + 1, 1);
      tc::mbar_init(p_full + 2 * s, 4); tc::mbar_init(p_full + 2 * s + 1, 4);
      tc::mbar_init(o_final + s, 1); tc::mbar_init(o_read + s, 4); tc::mbar_init(pv_done + s, 1); tc::mbar_init(qt_full + s, 4);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<C::kTmemCols>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> (batch, head, first query row of tile A, key tiles of tile A / tile B)
  auto item_coords = [&](int item, int& b, int& h, int& q0, int& nA, int& nB) {
    int bh = item / p.q_pairs;
    int qp = (item % p.q_pairs + bh) % p.q_pairs;     // rotated with bh: a CTA's successive items differ in work (causal, kv_len)
    h = bh % p.H; b = bh / p.H;
    q0 = qp * 2 * BQ;
    const int lk = p.kv_len != nullptr ? min(p.Lk, max(1, __ldg(p.kv_len + b))) : p.Lk;
    const int all = (lk + BKV - 1) / BKV;
    nA = p.causal ? min(all, 2 * qp + 1) : all;
    nB = p.causal ? min(all, 2 * qp + 2) : all;
  };

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsLow));
  if (warp == 0) {
    // ======================================= TMA producer =======================================
    if (lane == 0) {
      uint32_t kc = 0, vc = 0, ic = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++ic) {
        int b, h, q0, nA, nB;
        item_coords(item, b, h, q0, nA, nB);
        tc::mbar_wait(q_empty, (ic & 1) ^ 1);
        tc::mbar_arrive_expect_tx(q_full, 2 * C::kTileBytes);
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int c = 0; c < C::kChunks; ++c)
            tc::tma_load_2d(smem + C::kOffQ + t * C::kTileBytes + c * (BQ * 128), &tm_q, h * DH + c * 32, b * p.Lq + q0 + t * BQ, q_full);
        for (int j = 0; j < nB; ++j) {
          const int ks = kc & 1;
          tc::mbar_wait(k_empty + ks, ((kc >> 1) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(k_full + ks, C::kTileBytes);
#pragma unroll
          for (int c = 0; c < C::kChunks; ++c)
            tc::tma_load_2d(smem + C::kOffK + ks * C::kTileBytes + c * (BKV * 128), &tm_k, h * DH + c * 32, b * p.Lk + j * BKV, k_full + ks);
          ++kc;
          const int vs = vc & 1;
          tc::mbar_wait(v_empty + vs, ((vc >> 1) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(v_full + vs, C::kTileBytes);
#pragma unroll
          for (int c = 0; c < C::kChunks; ++c)
            tc::tma_load_2d(smem + C::kOffV + vs * C::kTileBytes + c * (BKV * 128), &tm_v, h * DH + c * 32, b * p.Lk + j * BKV, v_full + vs);
          ++vc;
        }
      }
    }
  } else if (warp == 1) {
    // ======================================= MMA issuer =========================================
    // The whole warp walks the loop (warp-uniform control flow and addresses); one elected lane issues.
    constexpr uint32_t idesc_qk = tc::make_idesc_tf32(BQ, BS, 0, 0);
    constexpr uint32_t idesc_pv = tc::make_idesc_tf32(BQ, DH, 0, 1);
    uint32_t kc = 0, vc = 0, ic = 0;
    uint32_t oc[2] = {0, 0};          // items finished per group
    uint32_t tb[2] = {0, 0};          // key tiles of earlier items per group (phase base of s_full / p_full)
    // S_g[buf] = Q_g K^T over the 64 keys `hh` of the key tile in stage ks
    auto issue_qk = [&](int g, int ks, int hh) {
      const uint32_t sk = tc::smem_u32(smem + C::kOffK + ks * C::kTileBytes) + hh * (BS * 128);
      const uint32_t d_tmem = tmem_base + C::kColS + g * BKV + hh * BS;
      const uint32_t a_tmem = tmem_base + C::kColQ + g * DH;
      if (tc::elect_one()) {
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c) {
          const uint64_t dk = tc::make_smem_desc(sk + c * (BKV * 128), 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)      // A = Q_g from TMEM: 32 cycles per 128x64x8 step instead of 48 with A in shared memory
            tc::mma_tf32_ts(d_tmem, a_tmem + c * 32 + k * 8, tc::desc_advance(dk, k * 32), idesc_qk, (c > 0 || k > 0) ? 1u : 0u);
        }
        tc::tc_commit(s_full + 2 * g + hh);
      }
      __syncwarp();
    };
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++ic) {
      int b, h, q0, nA, nB;
      item_coords(item, b, h, q0, nA, nB);
      const int ng[2] = {nA, nB};
      tc::mbar_wait(qt_full + 0, ic & 1);               // Q_A / Q_B are in TMEM (their smem tiles are already released)
      tc::mbar_wait(qt_full + 1, ic & 1);
      {   // key tile 0: both halves for both query tiles (S is double-buffered: the softmax never waits for its next S)
        const int ks = kc & 1;
        tc::mbar_wait(k_full + ks, (kc >> 1) & 1);
        tc::tc_fence_after();
        issue_qk(0, ks, 0); issue_qk(1, ks, 0);
        issue_qk(0, ks, 1); issue_qk(1, ks, 1);
        if (tc::elect_one()) tc::tc_commit(k_empty + ks);
        __syncwarp();
        ++kc;
      }
      for (int j = 0; j < nB; ++j) {
        const int vs = vc & 1;
        tc::mbar_wait(v_full + vs, (vc >> 1) & 1);
        const bool more = j + 1 < nB;
        const int ks = kc & 1;
        bool k_ready = false;
        const uint32_t sv = tc::smem_u32(smem + C::kOffV + vs * C::kTileBytes);
        // V tile: kChunks MN blocks (32 head-dim columns each) of 128 key rows x 128 B; 4-row swizzle atoms
        const uint64_t dv = tc::make_smem_desc(sv, BKV * 128, 512, tc::kLayoutSw128Base32);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (j >= ng[g]) continue;
            tc::mbar_wait(p_full + 2 * g + hh, (tb[g] + j) & 1);
            if (j == 0 && hh == 0) tc::mbar_wait(o_read + g, (oc[g] & 1) ^ 1);        // the previous item's O_g has been read out
            tc::tc_fence_after();
            const uint32_t a_tmem = tmem_base + C::kColS + g * BKV + hh * BS;
            const uint32_t d_tmem = tmem_base + C::kColO + g * DH;
            const bool last = j + 1 == ng[g] && hh == 1;
            if (tc::elect_one()) {
#pragma unroll
              for (int k = 0; k < BS / 8; ++k)
                tc::mma_tf32_ts(d_tmem, a_tmem + k * 8, tc::desc_advance(dv, (hh * (BS / 8) + k) * 1024), idesc_pv, (j > 0 || hh > 0 || k > 0) ? 1u : 0u);
              tc::tc_commit(pv_done + g);
              if (last) tc::tc_commit(o_final + g);
            }
            __syncwarp();
            if (last) ++oc[g];
            if (j + 1 < ng[g]) {       // the same half of the NEXT key tile into the S buffer this P V has just read
              if (!k_ready) { tc::mbar_wait(k_full + ks, (kc >> 1) & 1); tc::tc_fence_after(); k_ready = true; }
              issue_qk(g, ks, hh);
            }
          }
        }
        if (tc::elect_one()) {
          tc::tc_commit(v_empty + vs);
          if (more) tc::tc_commit(k_empty + ks);
        }
        __syncwarp();
        ++vc;
        if (more) ++kc;
      }
      tb[0] += nA; tb[1] += nB;
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsHigh));
    // ============================ softmax: one warpgroup per query tile ============================
    const int g = (warp - 4) >> 2;                       // query tile of the pair
    const int quarter = warp & 3;                        // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int tid = threadIdx.x - 128;                   // 0..255
    const float ks_drop = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;
    const uint32_t s_addr = tmem_base + lane_addr + C::kColS + g * BKV;
    const uint32_t o_addr = tmem_base + lane_addr + C::kColO + g * DH;
    uint32_t sc = 0, fc = 0;
#ifdef PA_ATTN_TRACE
    const bool trace_on = (p.debug & 1024) && blockIdx.x == 0 && lane == 0 && (warp == 4 || warp == 8);
    int trace_n = 0;
    const int trole = warp == 4 ? 2 : 3;
#endif
    // additive key bias (0 / -inf) + "all 32 keys valid" flags of one item (both query tiles share batch and head), two
    // tables alternating; the next item's table is written while the tensor pipe finishes the current item
    auto build_table = [&](int item, uint32_t parity) {
      int b_, h_, q0_, nA_, nB_;
      item_coords(item, b_, h_, q0_, nA_, nB_);
      float* bt = bias_s + (parity & 1) * p.LkPad;
      int* ft = flag_s + (parity & 1) * 64;
      for (int k = tid; k < nB_ * BKV; k += 256) {
        const bool ok = k < p.Lk && !(p.kpm != nullptr && p.kpm[(int64_t)b_ * p.Lk + k]);
        bt[k] = ok ? 0.f : -INFINITY;
        const bool all_ok = __all_sync(0xffffffffu, ok);
        if (lane == 0) ft[k >> 5] = all_ok ? 1 : 0;
      }
    };
    if ((int)blockIdx.x < p.items) build_table(blockIdx.x, 0);
    uint32_t itc = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++itc) {
      int b, h, q0, nA, nB;
      item_coords(item, b, h, q0, nA, nB);
      const int n = g == 0 ? nA : nB;
      const int q0g = q0 + g * BQ;
      const int qi = q0g + row;
      const int64_t row_global = ((int64_t)(b * p.H + h) * p.Lq + qi);
      const uint32_t bias_u32 = tc::smem_u32(bias_s + (itc & 1) * p.LkPad);
      const int* flag_it = flag_s + (itc & 1) * 64;
      float m_used = -INFINITY, l_run = 0.f;            // m_used: RAW-score maximum the probabilities are taken relative to

      auto load_w = [&](int k0n, uint32_t (&wd)[4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int wi = (k0n >> 5) + i;
          wd[i] = (qi < p.Lq && wi < p.LkW) ? __ldg(p.drop_rows + row_global * p.LkW + wi) : 0u;
        }
      };
      uint32_t w_pref[4] = {0u, 0u, 0u, 0u};
      if (DROP) load_w(0, w_pref);
      {
        // stationary operand: this thread's row of Q_g from the TMA tile in shared memory (128B swizzle: 16-byte piece i of
        // row r sits at piece i ^ (r & 7)) into TMEM.  Every QK^T of the previous item has retired (this group consumed all
        // of its S tiles), so the TMEM columns are free; the smem tile is released as soon as both groups have copied.
        tc::mbar_wait(q_full, itc & 1);
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c) {
          const uint32_t qrow = tc::smem_u32(smem + C::kOffQ + g * C::kTileBytes + c * (BQ * 128) + row * 128);
          uint32_t rq[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 a4 = tc::ld_shared_v4(qrow + ((i ^ (row & 7)) << 4));
            rq[4 * i] = __float_as_uint(a4.x); rq[4 * i + 1] = __float_as_uint(a4.y); rq[4 * i + 2] = __float_as_uint(a4.z); rq[4 * i + 3] = __float_as_uint(a4.w);
          }
          tc::tmem_st_32x32(tmem_base + lane_addr + C::kColQ + g * DH + c * 32, rq);
        }
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) { tc::mbar_arrive(qt_full + g); tc::mbar_arrive(q_empty); }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");      // the item's bias table is complete (both groups)
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      for (int t = 0; t < 2 * n; ++t) {                  // 64-key steps: half hh of key tile j
        const int j = t >> 1, hh = t & 1;
        const int k0 = j * BKV + hh * BS;
        TRACE(trole, 0);
        if (hh == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) w[i] = w_pref[i];
          if (DROP && j + 1 < n) load_w((j + 1) * BKV, w_pref);
        }
        tc::mbar_wait(s_full + 2 * g + hh, (sc / 2 + j) & 1);
        TRACE(trole, 2);
        tc::tc_fence_after();
        const uint32_t s_tm = s_addr + hh * BS;
        float s[BS];
        {
          uint32_t r0[32], r1[32];                         // both 32-column chunks in flight before the single wait
          tc::tmem_ld_32x32(s_tm, r0);
          tc::tmem_ld_32x32(s_tm + 32, r1);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) { s[c] = __uint_as_float(r0[c]); s[32 + c] = __uint_as_float(r1[c]); }
        }
        TRACE(trole, 3);
        const bool diag = p.causal && (k0 + BS - 1 > q0g);
        const int fg = k0 >> 5;
        const bool clean = !diag && (flag_it[fg] & flag_it[fg + 1]);
        if (!clean) {
#pragma unroll
          for (int c = 0; c < BS; c += 4) {
            const float4 bv = tc::ld_shared_v4(bias_u32 + 4 * (k0 + c));      // broadcast
            s[c] += bv.x; s[c + 1] += bv.y; s[c + 2] += bv.z; s[c + 3] += bv.w;
          }
          if (diag) {
#pragma unroll
            for (int c = 0; c < BS; ++c) s[c] = (k0 + c > qi) ? -INFINITY : s[c];
          }
        }
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < BS; ++c) m4[c & 3] = fmaxf(m4[c & 3], s[c]);
        const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        TRACE(trole, 4);
        // lazy rescale: keep the stale maximum unless the new one exceeds it by more than 2^8 (or there is none yet)
        const bool grow = mx > m_used && ((mx - m_used) * p.scale_log2 > kRescaleLog2 || m_used == -INFINITY);
        if (__any_sync(0xffffffffu, grow)) {
          const float corr = grow ? fast_exp2((m_used - mx) * p.scale_log2) : 1.f;      // m_used = -inf -> 0
          if (grow) { l_run *= corr; m_used = mx; }
          if (t > 0) {
            // O_g holds PV_g(0..t-1): the last of them must have retired (PV_g(t) cannot start before this step's P)
            tc::mbar_wait(pv_done + g, (sc + t - 1) & 1);
            tc::tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < DH; c0 += 32) {
              uint32_t r[32];
              tc::tmem_ld_32x32(o_addr + c0, r);
              tc::tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 32; ++c) r[c] = __float_as_uint(__uint_as_float(r[c]) * corr);
              tc::tmem_st_32x32(o_addr + c0, r);
            }
          }
        }
        const float m_safe = m_used == -INFINITY ? 0.f : m_used;
        const float neg_ms = -m_safe * p.scale_log2;
        // exp2, row sum, dropout, TF32 rounding and the TMEM store of P chunk by chunk in ONE basic block (DROP is a template
        // constant): the scheduler overlaps the MUFU queue of one 32-column chunk with the ALU work of its neighbours
        float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c0 = 0; c0 < BS; c0 += 32) {
          uint32_t r[32];
          const uint32_t wd = w[2 * hh + (c0 >> 5)];
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float e = fast_exp2(fmaf(s[c0 + c], p.scale_log2, neg_ms));
            rs4[c & 3] += e;
            if (DROP) e = ((wd >> c) & 1u) ? e : 0.f;               // x 1/(1-p) folded into the final scale
            r[c] = tf32_rn_finite_bits(e);                          // P >= 0, finite
          }
          tc::tmem_st_32x32(s_tm + c0, r);
        }
        l_run += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
        TRACE(trole, 5);
        tc::tmem_st_wait();
        TRACE(trole, 6);
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(p_full + 2 * g + hh);
        TRACE(trole, 7);
      }
      sc += 2 * n;                                       // P V steps of this group so far (pv_done phases)
      // the next item's table is built here: its key-padding loads overlap the wait for this item's last P V
      if (item + (int)gridDim.x < p.items) build_table(item + gridDim.x, itc + 1);
      tc::mbar_wait(o_final + g, fc & 1);
      ++fc;
      tc::tc_fence_after();
      TRACE(trole, 8);
      {
        const float inv = l_run > 0.f ? ks_drop / l_run : 0.f;
        float* op = p.o + ((int64_t)b * p.Lq + qi) * p.ldo + h * DH;
#pragma unroll
        for (int c0 = 0; c0 < DH; c0 += 32) {
          uint32_t r[32];
          tc::tmem_ld_32x32(o_addr + c0, r);
          tc::tmem_ld_wait();
          if (qi < p.Lq) {
#pragma unroll
            for (int c = 0; c < 32; c += 8) {
              float4 v = make_float4(__uint_as_float(r[c]) * inv, __uint_as_float(r[c + 1]) * inv, __uint_as_float(r[c + 2]) * inv, __uint_as_float(r[c + 3]) * inv);
              float4 u = make_float4(__uint_as_float(r[c + 4]) * inv, __uint_as_float(r[c + 5]) * inv, __uint_as_float(r[c + 6]) * inv, __uint_as_float(r[c + 7]) * inv);
              if (p.round_out) { v = tf32_rn4(v); u = tf32_rn4(u); }
              if (p.wide_st) st_global_v8(op + c0 + c, v, u);
              else { *reinterpret_cast<float4*>(op + c0 + c) = v; *reinterpret_cast<float4*>(op + c0 + c + 4) = u; }
            }
          }
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(o_read + g);
        if (p.lse != nullptr && qi < p.Lq)
          p.lse[((int64_t)b * p.H + h) * p.Lq + qi] = l_run > 0.f ? (m_used * p.scale_log2 + log2f(l_run)) * 0.6931471805599453f : -INFINITY;
      }
      TRACE(trole, 10);
    }
#ifdef PA_ATTN_TRACE
    if (trace_on) g_attn2_trace_n[trole] = trace_n;
#endif
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
  const int q_tiles = (a.Lq + BQ - 1) / BQ;
  p.q_pairs = q_tiles / 2;
  p.items = p.q_pairs * a.H * a.B;
  p.LkPad = (a.Lk + BKV - 1) / BKV * BKV;
  p.wide_st = (((uintptr_t)a.o & 31) == 0 && a.ldo % 8 == 0) ? 1 : 0;
  p.kv_len = a.kpm != nullptr ? a.kv_len : nullptr;
  const int smem_bytes = C::kSmemFixed + 2 * p.LkPad * 4;
  if (smem_bytes > 227 * 1024 || p.LkPad > 2048) { pa_set_error("pa_attn_fwd (tc): Lk = %d too long for the bias table", a.Lk); return PA_ERR_UNSUPPORTED; }
  int grid = p.items < pa_num_sms() ? p.items : pa_num_sms();
  if (a.p_drop > 0.f) {
    auto kern = attn_fwd_pp_kernel<DH, true>;
    static SmemAttrCache attr;
    if (int rc_attr = pa_set_max_smem(kern, smem_bytes, attr)) return rc_attr;
    kern<<<grid, kThreads, smem_bytes, st>>>(tq, tk, tv, p);
  } else {
    auto kern = attn_fwd_pp_kernel<DH, false>;
    static SmemAttrCache attr;
    if (int rc_attr = pa_set_max_smem(kern, smem_bytes, attr)) return rc_attr;
    kern<<<grid, kThreads, smem_bytes, st>>>(tq, tk, tv, p);
  }
  PA_CHECK_LAUNCH();
  return PA_OK;
}

}  // namespace

// 1 when the pair kernel serves this shape (an even number of 128-row query tiles), else attn_tc.cu does
bool pa_attn_fwd_pp_ok(const pa_attn_fwd_args* a) {
  static const int mode = [] { const char* e = getenv("PLANK_B200_ATTN_FWD"); return e ? atoi(e) : 2; }();      // 1 = attn_tc.cu always
  const int q_tiles = (a->Lq + BQ - 1) / BQ;
  return mode >= 2 && q_tiles % 2 == 0 && (a->dh == 32 || a->dh == 64);
}

int pa_attn_fwd_pp(const pa_attn_fwd_args* a, void* stream) {
  switch (a->dh) {
    case 32: return launch<32>(*a, (cudaStream_t)stream);
    case 64: return launch<64>(*a, (cudaStream_t)stream);
    default: pa_set_error("pa_attn_fwd (tc): head dim %d unsupported (32, 64)", a->dh); return PA_ERR_UNSUPPORTED;
  }
}

#ifdef PA_ATTN_TRACE
extern "C" int pa_debug_attn2_trace(unsigned long long* out_host /*[4][1024]*/, int* n_host /*[4]*/) {
  PA_CUDA(cudaMemcpyFromSymbol(out_host, g_attn2_trace, sizeof(unsigned long long) * 4 * kTraceLen2));
  PA_CUDA(cudaMemcpyFromSymbol(n_host, g_attn2_trace_n, sizeof(int) * 4));
  return PA_OK;
}
#endif
