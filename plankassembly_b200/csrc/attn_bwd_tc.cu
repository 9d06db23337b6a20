// Attention backward on the 5th-gen tensor cores (TF32 operands, FP32 accumulation in TMEM).
// Two persistent, warp-specialised kernels (warp 0 TMA, warp 1 MMA issuer, warps 2..9 elementwise: two
// warpgroups, each owning one 32-column half of every 64-column score tile -- warps w and w+4 share a
// TMEM lane quarter):
//
//  dq kernel  (query-stationary, 128 queries x 64-key steps; TMEM lanes = queries)
//     S  = Q K_j^T,  dP = dO V_j^T                       (tcgen05.mma SS, K-major operands)
//     dS = P o (dropout(dP) - delta),  P = exp2(S*c - lse)  in registers, written back over S in TMEM
//     dQ += dS K_j                                        (A = dS from TMEM, B = K_j MN-major); dQ stays in
//                                                          TMEM for the whole key loop -> one store, no atomics
//  dkdv kernel (key-stationary, 128 keys x 64-query steps; TMEM lanes = KEYS: scores are computed
//              transposed so that P^T and dS^T are directly usable as TMEM A-operands)
//     S^T = K Q_i^T,  dP^T = V dO_i^T
//     P^T, dS^T in registers (thread = key, columns = queries; lse/delta per column from smem;
//              dropout keep-bits come from the column-major bit plane written by dropmask.cu)
//     dV += P^T dO_i,  dK += dS^T Q_i                      (B operands MN-major); both stay in TMEM
//
// TF32 MN-major operands need the 128B_BASE32B smem layout while K-major ones need plain 128B swizzle,
// so tiles used both ways (K_j in the dq kernel, Q_i / dO_i in the dkdv kernel) are loaded twice by TMA
// with the two swizzle modes.  Same math as attn_simt.cu (which is the fp32 check for these kernels).
#include "common.cuh"
#include "tc_common.cuh"

int pa_attn_delta_launch(const float* o, const float* d_o, int64_t ldo, int B, int H, int Lq, int dh, float* delta, cudaStream_t st);

// Debug timeline (PLANK_B200_NVCC_FLAGS=-DPA_ATTN_TRACE, PLANK_B200_ATTN_DEBUG=1024): clock64 per event of CTA 0 of the
// dQ kernel (roles: 0 TMA producer, 1 MMA issuer, 2 elementwise warp 2, 3 elementwise warp 6); pa_debug_attn_bwd_trace().
#ifdef PA_ATTN_TRACE
#include <stdlib.h>
__device__ unsigned long long g_bwd_trace[4][1024];
__device__ int g_bwd_trace_n[4];
#define TRACE(role, ev)                                                                          \
  do {                                                                                           \
    if (trace_on && trace_n < 1024) g_bwd_trace[role][trace_n++] = ((unsigned long long)(ev) << 56) | (clock64() & 0xffffffffffffffull); \
  } while (0)
#define TRACE_DECL(cond) const bool trace_on = (p.trace & kTraceBit) && blockIdx.x == 0 && (cond); int trace_n = 0
#define TRACE_END(role) do { if (trace_on) g_bwd_trace_n[role] = trace_n; } while (0)
#else
#define TRACE(role, ev) do {} while (0)
#define TRACE_DECL(cond) do {} while (0)
#define TRACE_END(role) do {} while (0)
#endif

namespace {

constexpr int kThreads = 320;      // warp 0 TMA, warp 1 MMA, warps 2..9 = two elementwise warpgroups (column halves)
constexpr float kLog2e = 1.4426950408889634f;

struct BwdParams {
  const float* lse; const float* delta; const uint8_t* kpm;
  float* dq; float* dk; float* dv; int64_t lddq, lddk, lddv;
  int B, H, Lq, Lk, causal, round_out;
  float scale, scale_log2;
  float p_drop; const uint32_t* drop_rows; const uint32_t* drop_cols; int LkW, LqW;
  float* dbias;          // [3*H*DH] += column sums of dq | dk | dv (in-projection bias gradient), may be NULL
  int tiles, items;
  const int32_t* kv_len; // optional [B]: keys >= kv_len[b] are all PAD -> key tiles beyond it are skipped
  int wide_st;           // outputs are 32-byte aligned with row pitches that are multiples of 8 floats: 256-bit stores
  int LkPad;             // keys rounded up to the key tile (per-item bias table length)
  int LqPad;             // queries rounded up to the query tile (per-item lse/delta table length)
  int trace;
  int l2_prefetch;       // producers pull the tiles of a stage's NEXT use into L2 (PLANK_B200_ATTN_L2PF=1; measured: no gain, 347.4 vs 348.5 us, off)
};

// ================================================================================================
//                                           dQ kernel
// ================================================================================================
template <int DH> struct CfgQ {
  static constexpr int BQ = 128, BK = 64, kStages = 3;
  static constexpr int kChunks = DH / 32;
  static constexpr int kQBytes = kChunks * BQ * 128;      // Q_i or dO_i (K-major)
  static constexpr int kKBytes = kChunks * BK * 128;      // one 64-key tile in one layout
  static constexpr int kStageBytes = 3 * kKBytes;         // K (K-major) | K (MN-major) | V (K-major)
  static constexpr int kOffQ = 0, kOffDO = kQBytes, kOffKV = 2 * kQBytes;
  static constexpr int kOffBar = kOffKV + kStages * kStageBytes;
  static constexpr int kOffFlag = kOffBar + 256;          // [2 items][64] "all 32 keys valid" flags
  static constexpr int kOffBias = kOffFlag + 512;         // [2 items][LkPad] additive key bias (0 / -inf), sized at launch
  static constexpr int kSmemFixed = kOffBias + 1024;      // + 2 * LkPad * 4
  static constexpr int kColS = 0, kColDP = 128, kColDQ = 256;   // S: 2x64, dP: 2x64, dQ: DH
  static constexpr int kColQ = 320, kColDO = 384;               // the stationary Q_i / dO_i tiles as TMEM A-operands
  static constexpr int kTmemCols = 512;
};

template <int DH>
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                      const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_k_mn,
                      const __grid_constant__ CUtensorMap tm_v, const BwdParams p) {
  using C = CfgQ<DH>;
  constexpr int kTraceBit = 1024; (void)kTraceBit;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* bias_s = reinterpret_cast<float*>(smem + C::kOffBias);
  uint32_t* flags_s = reinterpret_cast<uint32_t*>(smem + C::kOffFlag);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
  uint64_t* qdo_full = bars + 0;  uint64_t* qdo_empty = bars + 1;
  uint64_t* kv_full = bars + 2;   uint64_t* kv_empty = bars + 5;     // [3] each
  uint64_t* sdp_full = bars + 8;  uint64_t* ds_full = bars + 10;     // [2] each
  uint64_t* dq_full = bars + 12;  uint64_t* dq_empty = bars + 13;
  uint64_t* qt_full = bars + 14;                                     // Q_i / dO_i copied into TMEM by the elementwise warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tm_q); tc::tma_prefetch_desc(&tm_do); tc::tma_prefetch_desc(&tm_k);
    tc::tma_prefetch_desc(&tm_k_mn); tc::tma_prefetch_desc(&tm_v);
    tc::mbar_init(qdo_full, 1); tc::mbar_init(qdo_empty, 8); tc::mbar_init(qt_full, 8);
    for (int s = 0; s < C::kStages; ++s) { tc::mbar_init(kv_full + s, 1); tc::mbar_init(kv_empty + s, 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(sdp_full + s, 1); tc::mbar_init(ds_full + s, 8); }
    tc::mbar_init(dq_full, 1); tc::mbar_init(dq_empty, 8);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<C::kTmemCols>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto coords = [&](int item, int& b, int& h, int& q0, int& n) {
    int bh = item / p.tiles;
    int qt = (item % p.tiles + bh) % p.tiles;   // rotated with bh: balances causal / kv_len-shortened items over the CTAs
    h = bh % p.H; b = bh / p.H; q0 = qt * C::BQ;
    const int lk = p.kv_len != nullptr ? min(p.Lk, max(1, __ldg(p.kv_len + b))) : p.Lk;
    int all = (lk + C::BK - 1) / C::BK;
    n = p.causal ? min(all, (q0 + C::BQ - 1) / C::BK + 1) : all;
  };

  if (warp == 0) {
    if (lane == 0) {
      TRACE_DECL(true);
      uint32_t kc = 0, ic = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++ic) {
        int b, h, q0, n;
        coords(item, b, h, q0, n);
        tc::mbar_wait(qdo_empty, (ic & 1) ^ 1);
        TRACE(0, 0);
        tc::mbar_arrive_expect_tx(qdo_full, 2 * C::kQBytes);
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c) {
          tc::tma_load_2d(smem + C::kOffQ + c * (C::BQ * 128), &tm_q, h * DH + c * 32, b * p.Lq + q0, qdo_full);
          tc::tma_load_2d(smem + C::kOffDO + c * (C::BQ * 128), &tm_do, h * DH + c * 32, b * p.Lq + q0, qdo_full);
        }
        for (int j = 0; j < n; ++j, ++kc) {
          const int s = kc % C::kStages;
          if (p.l2_prefetch && j + C::kStages < n) {      // the tiles this stage will receive NEXT time, into L2 now
#pragma unroll
            for (int c = 0; c < C::kChunks; ++c) {
              tc::tma_prefetch_l2_2d(&tm_k, h * DH + c * 32, b * p.Lk + (j + C::kStages) * C::BK);
              tc::tma_prefetch_l2_2d(&tm_v, h * DH + c * 32, b * p.Lk + (j + C::kStages) * C::BK);
            }
          }
          tc::mbar_wait(kv_empty + s, ((kc / C::kStages) & 1) ^ 1);
          TRACE(0, 1);
          tc::mbar_arrive_expect_tx(kv_full + s, C::kStageBytes);
          uint8_t* base = smem + C::kOffKV + s * C::kStageBytes;
#pragma unroll
          for (int c = 0; c < C::kChunks; ++c) {
            tc::tma_load_2d(base + c * (C::BK * 128), &tm_k, h * DH + c * 32, b * p.Lk + j * C::BK, kv_full + s);
            tc::tma_load_2d(base + C::kKBytes + c * (C::BK * 128), &tm_k_mn, h * DH + c * 32, b * p.Lk + j * C::BK, kv_full + s);
            tc::tma_load_2d(base + 2 * C::kKBytes + c * (C::BK * 128), &tm_v, h * DH + c * 32, b * p.Lk + j * C::BK, kv_full + s);
          }
        }
      }
      TRACE_END(0);
    }
  } else if (warp == 1) {
    // The whole warp walks the loop (warp-uniform control flow and addresses); one elected lane issues the MMAs.
    {
      constexpr uint32_t idesc_s = tc::make_idesc_tf32(C::BQ, C::BK, 0, 0);
      constexpr uint32_t idesc_dq = tc::make_idesc_tf32(C::BQ, DH, 0, 1);
      TRACE_DECL(lane == 0);
      uint32_t kc = 0, ic = 0, st = 0, dt = 0;
      // Q_i and dO_i are A-operands in TMEM (written once per item by the elementwise warps): in the SS form every
      // 128x64x8 MMA re-read 4 KB of them from shared memory and the kernel ran at the shared-memory bandwidth
      auto issue_sdp = [&](uint32_t kcs) {       // S and dP of the tile held in KV stage kcs % kStages
        const int s = kcs % C::kStages;
        tc::mbar_wait(kv_full + s, (kcs / C::kStages) & 1);
        TRACE(1, 1);
        tc::tc_fence_after();
        const uint32_t sk = tc::smem_u32(smem + C::kOffKV + s * C::kStageBytes);
        const uint32_t sv = sk + 2 * C::kKBytes;
        const int buf = st & 1;
        if (tc::elect_one()) {
#pragma unroll
          for (int c = 0; c < C::kChunks; ++c) {
            const uint64_t db = tc::make_smem_desc(sk + c * (C::BK * 128), 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc::mma_tf32_ts(tmem_base + C::kColS + buf * C::BK, tmem_base + C::kColQ + c * 32 + k * 8, tc::desc_advance(db, k * 32), idesc_s, (c > 0 || k > 0) ? 1u : 0u);
          }
#pragma unroll
          for (int c = 0; c < C::kChunks; ++c) {
            const uint64_t db = tc::make_smem_desc(sv + c * (C::BK * 128), 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc::mma_tf32_ts(tmem_base + C::kColDP + buf * C::BK, tmem_base + C::kColDO + c * 32 + k * 8, tc::desc_advance(db, k * 32), idesc_s, (c > 0 || k > 0) ? 1u : 0u);
          }
          tc::tc_commit(sdp_full + buf);
        }
        __syncwarp();
        TRACE(1, 2);
        ++st;
      };
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++ic) {
        int b, h, q0, n;
        coords(item, b, h, q0, n);
        tc::mbar_wait(qt_full, ic & 1);                  // Q_i / dO_i are in TMEM
        TRACE(1, 0);
        tc::tc_fence_after();
        issue_sdp(kc);
        for (int j = 0; j < n; ++j) {
          if (j + 1 < n) issue_sdp(kc + 1);
          const int buf = dt & 1;
          if (j == 0) tc::mbar_wait(dq_empty, (ic & 1) ^ 1);   // previous item's dQ has been read out of TMEM
          tc::mbar_wait(ds_full + buf, (dt >> 1) & 1);
          TRACE(1, 3);
          tc::tc_fence_after();
          const int s = kc % C::kStages;
          const uint32_t skm = tc::smem_u32(smem + C::kOffKV + s * C::kStageBytes + C::kKBytes);
          const uint64_t db = tc::make_smem_desc(skm, C::BK * 128, 512, tc::kLayoutSw128Base32);
          if (tc::elect_one()) {
#pragma unroll
            for (int k = 0; k < C::BK / 8; ++k)
              tc::mma_tf32_ts(tmem_base + C::kColDQ, tmem_base + C::kColS + buf * C::BK + k * 8, tc::desc_advance(db, k * 1024), idesc_dq,
                              (j > 0 || k > 0) ? 1u : 0u);
            tc::tc_commit(kv_empty + s);
            if (j + 1 == n) tc::tc_commit(dq_full);
          }
          __syncwarp();
          TRACE(1, 4);
          ++kc; ++dt;
        }
      }
      TRACE_END(1);
    }
  } else {
    const int quarter = warp & 3, row = quarter * 32 + lane, tid = threadIdx.x - 64;
    const int half = (warp - 2) >> 2;            // which 32-column half of each tile this warpgroup owns
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float ks = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;
    uint32_t sc = 0, ic = 0;
    TRACE_DECL(lane == 0 && (warp == 2 || warp == 6));
#ifdef PA_ATTN_TRACE
    const int trole = warp == 2 ? 2 : 3;
#endif
    // The items are software-pipelined: the next item's per-row scalars, key-padding bytes and first dropout word are
    // requested, and its Q/dO rows are copied into TMEM, BEFORE the current item's dQ epilogue -- the global-load
    // latencies and the MMA pipeline fill of item i+1 hide behind the TMEM read-out and the stores of item i.
    constexpr int KPT = 8;                               // key-padding bytes per thread (Lk <= 2048)
    struct ItemRegs {
      int b, h, q0, n, qi; bool q_ok; int64_t rg;
      float lse, dl; uint8_t kp[KPT]; uint32_t mw0;
    };
    auto prefetch = [&](int item, ItemRegs& r) {         // only ISSUES the loads
      coords(item, r.b, r.h, r.q0, r.n);
      r.qi = r.q0 + row;
      r.q_ok = r.qi < p.Lq;
      r.rg = ((int64_t)(r.b * p.H + r.h) * p.Lq + r.qi);
      r.lse = r.q_ok ? p.lse[r.rg] : -INFINITY;
      r.dl = r.q_ok ? p.delta[r.rg] : 0.f;
#pragma unroll
      for (int m = 0; m < KPT; ++m) {
        const int k = tid + 256 * m;
        r.kp[m] = (p.kpm != nullptr && k < p.Lk && k < r.n * C::BK) ? p.kpm[(int64_t)r.b * p.Lk + k] : (uint8_t)0;
      }
      const int wi = (half * 32) >> 5;
      r.mw0 = p.p_drop > 0.f ? ((r.q_ok && wi < p.LkW) ? __ldg(p.drop_rows + r.rg * p.LkW + wi) : 0u) : 0xffffffffu;
    };
    // stationary operands: this thread's row of Q_i and dO_i (its warpgroup's 32-column chunk) from the TMA tile in
    // shared memory (128B swizzle: 16-byte piece i of row r sits at piece i ^ (r & 7)) into TMEM.  Called when every
    // S/dP MMA of the previous item has been consumed by these warps, so the TMEM columns are free.
    auto copy_qdo = [&](uint32_t icn) {
      TRACE(trole, 9);
      tc::mbar_wait(qdo_full, icn & 1);
      TRACE(trole, 10);
      if (half < C::kChunks) {
        const uint8_t* qrow = smem + C::kOffQ + half * (C::BQ * 128) + row * 128;
        const uint8_t* drow = smem + C::kOffDO + half * (C::BQ * 128) + row * 128;
        uint32_t rq[32], rdo[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int off = (i ^ (row & 7)) << 4;
          const uint4 a4 = *reinterpret_cast<const uint4*>(qrow + off);
          const uint4 d4 = *reinterpret_cast<const uint4*>(drow + off);
          rq[4 * i] = a4.x; rq[4 * i + 1] = a4.y; rq[4 * i + 2] = a4.z; rq[4 * i + 3] = a4.w;
          rdo[4 * i] = d4.x; rdo[4 * i + 1] = d4.y; rdo[4 * i + 2] = d4.z; rdo[4 * i + 3] = d4.w;
        }
        tc::tmem_st_32x32(tmem_base + lane_addr + C::kColQ + half * 32, rq);
        tc::tmem_st_32x32(tmem_base + lane_addr + C::kColDO + half * 32, rdo);
        tc::tmem_st_wait();
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) { tc::mbar_arrive(qt_full); tc::mbar_arrive(qdo_empty); }
      TRACE(trole, 11);
    };

    ItemRegs nx;
    if ((int)blockIdx.x < p.items) { prefetch(blockIdx.x, nx); copy_qdo(0); }
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++ic) {
      const ItemRegs cu = nx;
      const int b = cu.b, h = cu.h, q0 = cu.q0, n = cu.n, qi = cu.qi;
      const bool q_ok = cu.q_ok;
      const int64_t rg = cu.rg;
      const float lse2 = cu.lse == -INFINITY ? INFINITY : cu.lse * kLog2e;     // +inf => P = 0
      const float dl = cu.dl;
      // additive key bias (0 / -inf: PAD keys and keys beyond Lk) of the WHOLE item, written once; the two tables
      // alternate between items, so one named barrier per item is all the elementwise warps need
      float* bias_it = bias_s + (ic & 1) * p.LkPad;
      const uint32_t bias_u32 = tc::smem_u32(bias_it);
      uint32_t* flag_it = flags_s + (ic & 1) * 64;          // per 32-key group: 1 = every key valid (tiles of valid keys skip the bias)
#pragma unroll
      for (int m = 0; m < KPT; ++m) {
        const int k = tid + 256 * m;
        if (k < n * C::BK) {
          const bool ok = k < p.Lk && !cu.kp[m];
          bias_it[k] = ok ? 0.f : -INFINITY;
          const bool all_ok = __all_sync(0xffffffffu, ok);
          if (lane == 0) flag_it[k >> 5] = all_ok ? 1u : 0u;
        }
      }
      auto load_mw = [&](int k0n) -> uint32_t {            // keep-bits of this thread's 32 keys for this query row
        const int wi = (k0n + half * 32) >> 5;
        return (q_ok && wi < p.LkW) ? __ldg(p.drop_rows + rg * p.LkW + wi) : 0u;
      };
      uint32_t mw_pref = cu.mw0;                           // global loads run one tile ahead of their use
      TRACE(trole, 12);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      TRACE(trole, 13);
      for (int j = 0; j < n; ++j, ++sc) {
        const int buf = sc & 1, k0 = j * C::BK;
        TRACE(trole, 0);
        const uint32_t mw = mw_pref;
        if (p.p_drop > 0.f && j + 1 < n) mw_pref = load_mw(k0 + C::BK);
        TRACE(trole, 1);
        tc::mbar_wait(sdp_full + buf, (sc >> 1) & 1);
        TRACE(trole, 2);
        tc::tc_fence_after();
        const bool diag = p.causal && (k0 + C::BK - 1 > q0);
        const bool clean = !diag && (flag_it[2 * j] & flag_it[2 * j + 1]) != 0;   // no masking needed anywhere in this tile
#pragma unroll
        for (int c0 = half * 32; c0 < half * 32 + 32; c0 += 32) {
          uint32_t rs[32], rd[32];
          tc::tmem_ld_32x32(tmem_base + lane_addr + C::kColS + buf * C::BK + c0, rs);
          tc::tmem_ld_32x32(tmem_base + lane_addr + C::kColDP + buf * C::BK + c0, rd);
          tc::tmem_ld_wait();
          TRACE(trole, 3);
          if (clean) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float mk = ((mw >> c) & 1u) ? ks : 0.f;
              const float pr = fast_exp2(fmaf(__uint_as_float(rs[c]), p.scale_log2, -lse2));
              const float ds = pr * fmaf(__uint_as_float(rd[c]), mk, -dl);
              rs[c] = tf32_rn_finite_bits(ds);
            }
          } else {
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              float mk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) mk[e] = ((mw >> (c + e)) & 1u) ? ks : 0.f;
              const float4 b4 = tc::ld_shared_v4(bias_u32 + 4 * (k0 + c0 + c));      // explicit LDS (a generic LD.E costs ~17 more cycles)
              const float bz[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int kk = c0 + c + e;
                float v = fmaf(__uint_as_float(rs[c + e]), p.scale_log2, bz[e] - lse2);
                if (diag && k0 + kk > qi) v = -INFINITY;
                const float pr = fast_exp2(v);
                const float ds = pr * fmaf(__uint_as_float(rd[c + e]), mk[e], -dl);
                rs[c + e] = tf32_rn_finite_bits(ds);
              }
            }
          }
          TRACE(trole, 4);
          tc::tmem_st_32x32(tmem_base + lane_addr + C::kColS + buf * C::BK + c0, rs);
        }
        tc::tmem_st_wait();
        TRACE(trole, 5);
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(ds_full + buf);
        TRACE(trole, 6);
      }
      // item boundary: get the next item going before this item's read-out
      const int next_item = item + gridDim.x;
      if (next_item < p.items) { prefetch(next_item, nx); copy_qdo(ic + 1); }
      // dQ_i is complete in TMEM
      tc::mbar_wait(dq_full, ic & 1);
      TRACE(trole, 7);
      tc::tc_fence_after();
      float* out = p.dq + ((int64_t)b * p.Lq + qi) * p.lddq + h * DH;
#pragma unroll
      for (int c0 = half * 32; c0 < DH; c0 += 64) {
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem_base + lane_addr + C::kColDQ + c0, r);
        tc::tmem_ld_wait();
        if (q_ok) {
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            float4 v = make_float4(__uint_as_float(r[c]) * p.scale, __uint_as_float(r[c + 1]) * p.scale,
                                   __uint_as_float(r[c + 2]) * p.scale, __uint_as_float(r[c + 3]) * p.scale);
            float4 w = make_float4(__uint_as_float(r[c + 4]) * p.scale, __uint_as_float(r[c + 5]) * p.scale,
                                   __uint_as_float(r[c + 6]) * p.scale, __uint_as_float(r[c + 7]) * p.scale);
            if (p.round_out) { v = tf32_rn4(v); w = tf32_rn4(w); }
            if (p.wide_st) st_global_v8(out + c0 + c, v, w);
            else { *reinterpret_cast<float4*>(out + c0 + c) = v; *reinterpret_cast<float4*>(out + c0 + c + 4) = w; }
          }
        }
        if (p.dbias != nullptr) {          // q-bias gradient: column sums over this warp's 32 query rows
          float cs[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) cs[c] = q_ok ? __uint_as_float(r[c]) * p.scale : 0.f;
          const float t = warp_colsum32(cs, lane);
          atomicAdd(p.dbias + h * DH + c0 + lane, t);
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(dq_empty);
      TRACE(trole, 8);
    }
#ifdef PA_ATTN_TRACE
    if (trace_on) g_bwd_trace_n[trole] = trace_n;
#endif
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<C::kTmemCols>(tmem_base);
}

// ================================================================================================
//                                         dK / dV kernel
// ================================================================================================
template <int DH> struct CfgK {
  static constexpr int BKV = 128, BQ = 64, kStages = 2;
  static constexpr int kChunks = DH / 32;
  static constexpr int kKVBytes = kChunks * BKV * 128;    // K_j or V_j (K-major)
  static constexpr int kQBytes = kChunks * BQ * 128;      // one 64-query tile in one layout
  static constexpr int kStageBytes = 4 * kQBytes;         // Q K-major | Q MN-major | dO K-major | dO MN-major
  static constexpr int kOffK = 0, kOffV = kKVBytes, kOffQ = 2 * kKVBytes;
  static constexpr int kOffBar = kOffQ + kStages * kStageBytes;
  static constexpr int kOffStat = kOffBar + 256;                      // [2 items][lse2 | delta][LqPad], sized at launch
  static constexpr int kSmemFixed = kOffStat + 1024;                  // + 2 * 2 * LqPad * 4
  static constexpr int kColS = 0, kColDP = 128, kColDK = 256, kColDV = 256 + DH;
  static constexpr int kColK = 384, kColV = 448;         // the stationary K_j / V_j tiles as TMEM A-operands of S^T / dP^T
  static constexpr int kTmemCols = 512;
};

template <int DH>
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_dkdv_tc_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                        const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_q_mn,
                        const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_do_mn,
                        const BwdParams p) {
  using C = CfgK<DH>;
  constexpr int kTraceBit = 2048; (void)kTraceBit;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* stat_s = reinterpret_cast<float*>(smem + C::kOffStat);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
  uint64_t* kv_full = bars + 0;   uint64_t* kv_empty = bars + 1;
  // Each query-tile stage is two half-stages with their own barriers: the K-major copies of Q_i / dO_i are released as
  // soon as S^T / dP^T have been computed (one tile earlier than the MN-major copies, which dV / dK read) -- with a
  // single release per stage the TMA refill of a slot started exactly when its data was needed next.
  uint64_t* q_full = bars + 2;    uint64_t* q_empty = bars + 4;      // [2] each: K-major halves
  uint64_t* qm_full = bars + 12;  uint64_t* qm_empty = bars + 14;    // [2] each: MN-major halves
  uint64_t* st_full = bars + 6;   uint64_t* pds_full = bars + 8;     // [2] each
  uint64_t* acc_full = bars + 10; uint64_t* acc_empty = bars + 11;
  uint64_t* kt_full = bars + 16;                                     // K_j / V_j copied into TMEM by the elementwise warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tm_k); tc::tma_prefetch_desc(&tm_v); tc::tma_prefetch_desc(&tm_q);
    tc::tma_prefetch_desc(&tm_q_mn); tc::tma_prefetch_desc(&tm_do); tc::tma_prefetch_desc(&tm_do_mn);
    tc::mbar_init(kv_full, 1); tc::mbar_init(kv_empty, 8); tc::mbar_init(kt_full, 8);
    for (int s = 0; s < C::kStages; ++s) {
      tc::mbar_init(q_full + s, 1); tc::mbar_init(q_empty + s, 1);
      tc::mbar_init(qm_full + s, 1); tc::mbar_init(qm_empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(st_full + s, 1); tc::mbar_init(pds_full + s, 8); }
    tc::mbar_init(acc_full, 1); tc::mbar_init(acc_empty, 8);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<C::kTmemCols>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int q_tiles_all = (p.Lq + C::BQ - 1) / C::BQ;
  auto coords = [&](int item, int& b, int& h, int& k0, int& i0) {
    int bh = item / p.tiles;
    int kt = (item % p.tiles + bh) % p.tiles;   // rotated with bh: balances causal / all-PAD key tiles over the CTAs
    h = bh % p.H; b = bh / p.H; k0 = kt * C::BKV;
    i0 = p.causal ? k0 / C::BQ : 0;          // first query tile that can see these keys
    if (p.kv_len != nullptr && k0 >= __ldg(p.kv_len + b) && k0 > 0) i0 = q_tiles_all;   // a key tile of PAD only: dK = dV = 0
  };

  if (warp == 0) {
    // Two independent producers (lanes 0 and 1): the K-major copies of Q_i / dO_i (operands of S^T / dP^T) are released one
    // MMA block earlier than the MN-major copies (operands of dV / dK).  One thread walking both in turn waited for the late
    // release before it could issue the next early load, and the MMA issuer then waited ~500 cycles per step for that load
    // (timeline: profiles/r2b_attn_bwd_dkdv_trace.txt).
    if (lane == 0) {
      uint32_t qc = 0, ic = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++ic) {
        int b, h, k0, i0;
        coords(item, b, h, k0, i0);
        tc::mbar_wait(kv_empty, (ic & 1) ^ 1);
        tc::mbar_arrive_expect_tx(kv_full, 2 * C::kKVBytes);
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c) {
          tc::tma_load_2d(smem + C::kOffK + c * (C::BKV * 128), &tm_k, h * DH + c * 32, b * p.Lk + k0, kv_full);
          tc::tma_load_2d(smem + C::kOffV + c * (C::BKV * 128), &tm_v, h * DH + c * 32, b * p.Lk + k0, kv_full);
        }
        for (int i = i0; i < q_tiles_all; ++i, ++qc) {
          const int s = qc % C::kStages;
          uint8_t* base = smem + C::kOffQ + s * C::kStageBytes;
          const int y = b * p.Lq + i * C::BQ;
          if (p.l2_prefetch && i + C::kStages < q_tiles_all) {      // the tiles this stage will receive NEXT time, into L2 now
#pragma unroll
            for (int c = 0; c < C::kChunks; ++c) {
              tc::tma_prefetch_l2_2d(&tm_q, h * DH + c * 32, y + C::kStages * C::BQ);
              tc::tma_prefetch_l2_2d(&tm_do, h * DH + c * 32, y + C::kStages * C::BQ);
            }
          }
          tc::mbar_wait(q_empty + s, ((qc / C::kStages) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(q_full + s, 2 * C::kQBytes);
#pragma unroll
          for (int c = 0; c < C::kChunks; ++c) {
            const int x = h * DH + c * 32, off = c * (C::BQ * 128);
            tc::tma_load_2d(base + off, &tm_q, x, y, q_full + s);
            tc::tma_load_2d(base + 2 * C::kQBytes + off, &tm_do, x, y, q_full + s);
          }
        }
      }
    } else if (lane == 1) {
      uint32_t qc = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        int b, h, k0, i0;
        coords(item, b, h, k0, i0);
        for (int i = i0; i < q_tiles_all; ++i, ++qc) {
          const int s = qc % C::kStages;
          uint8_t* base = smem + C::kOffQ + s * C::kStageBytes;
          const int y = b * p.Lq + i * C::BQ;
          tc::mbar_wait(qm_empty + s, ((qc / C::kStages) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(qm_full + s, 2 * C::kQBytes);
#pragma unroll
          for (int c = 0; c < C::kChunks; ++c) {
            const int x = h * DH + c * 32, off = c * (C::BQ * 128);
            tc::tma_load_2d(base + C::kQBytes + off, &tm_q_mn, x, y, qm_full + s);
            tc::tma_load_2d(base + 3 * C::kQBytes + off, &tm_do_mn, x, y, qm_full + s);
          }
        }
      }
    }
  } else if (warp == 1) {
    // The whole warp walks the loop (warp-uniform control flow and addresses); one elected lane issues the MMAs.
    {
      constexpr uint32_t idesc_s = tc::make_idesc_tf32(C::BKV, C::BQ, 0, 0);      // [128 keys x 64 queries]
      constexpr uint32_t idesc_acc = tc::make_idesc_tf32(C::BKV, DH, 0, 1);       // [128 keys x DH]
      TRACE_DECL(lane == 0);
      uint32_t qc = 0, ic = 0, st = 0, dt = 0;
      // S^T / dP^T take K_j / V_j from TMEM (tcgen05.mma with a TMEM A-operand: 32 cycles per 128x64x8 step; the SS form also
      // reads 4 KB of A from shared memory per step and needs 48 -- scripts/probe/mma_rate.cu -- and this kernel is bound by
      // its 32 MMAs per 64-query step)
      auto issue_st = [&](uint32_t qcs) {
        const int s = qcs % C::kStages;
        tc::mbar_wait(q_full + s, (qcs / C::kStages) & 1);
        TRACE(1, 1);
        tc::tc_fence_after();
        const uint32_t sq = tc::smem_u32(smem + C::kOffQ + s * C::kStageBytes);
        const uint32_t sdo = sq + 2 * C::kQBytes;
        const int buf = st & 1;
        if (tc::elect_one()) {
#pragma unroll
          for (int c = 0; c < C::kChunks; ++c) {
            const uint64_t db = tc::make_smem_desc(sq + c * (C::BQ * 128), 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc::mma_tf32_ts(tmem_base + C::kColS + buf * C::BQ, tmem_base + C::kColK + c * 32 + k * 8, tc::desc_advance(db, k * 32), idesc_s, (c > 0 || k > 0) ? 1u : 0u);
          }
#pragma unroll
          for (int c = 0; c < C::kChunks; ++c) {
            const uint64_t db = tc::make_smem_desc(sdo + c * (C::BQ * 128), 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc::mma_tf32_ts(tmem_base + C::kColDP + buf * C::BQ, tmem_base + C::kColV + c * 32 + k * 8, tc::desc_advance(db, k * 32), idesc_s, (c > 0 || k > 0) ? 1u : 0u);
          }
          tc::tc_commit(q_empty + s);                    // the K-major halves are free once these retire
          tc::tc_commit(st_full + buf);
        }
        __syncwarp();
        TRACE(1, 2);
        ++st;
      };
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++ic) {
        int b, h, k0, i0;
        coords(item, b, h, k0, i0);
        const int n = q_tiles_all - i0;
        tc::mbar_wait(kt_full, ic & 1);                  // K_j / V_j are in TMEM (the elementwise warps released the smem tiles)
        tc::tc_fence_after();
        if (n > 0) issue_st(qc);
        else tc::mbar_wait(acc_empty, (ic & 1) ^ 1);
        for (int j = 0; j < n; ++j) {
          if (j + 1 < n) issue_st(qc + 1);
          const int buf = dt & 1;
          if (j == 0) tc::mbar_wait(acc_empty, (ic & 1) ^ 1);    // previous item's dK/dV have been read out of TMEM
          tc::mbar_wait(pds_full + buf, (dt >> 1) & 1);
          TRACE(1, 3);
          tc::tc_fence_after();
          const int s = qc % C::kStages;
          tc::mbar_wait(qm_full + s, (qc / C::kStages) & 1);
          tc::tc_fence_after();
          const uint32_t sqm = tc::smem_u32(smem + C::kOffQ + s * C::kStageBytes + C::kQBytes);
          const uint32_t sdom = sqm + 2 * C::kQBytes;
          const uint64_t dqm = tc::make_smem_desc(sqm, C::BQ * 128, 512, tc::kLayoutSw128Base32);
          const uint64_t ddom = tc::make_smem_desc(sdom, C::BQ * 128, 512, tc::kLayoutSw128Base32);
          if (tc::elect_one()) {
#pragma unroll
            for (int k = 0; k < C::BQ / 8; ++k)      // dV += P^T dO_i
              tc::mma_tf32_ts(tmem_base + C::kColDV, tmem_base + C::kColS + buf * C::BQ + k * 8, tc::desc_advance(ddom, k * 1024), idesc_acc,
                              (j > 0 || k > 0) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < C::BQ / 8; ++k)      // dK += dS^T Q_i
              tc::mma_tf32_ts(tmem_base + C::kColDK, tmem_base + C::kColDP + buf * C::BQ + k * 8, tc::desc_advance(dqm, k * 1024), idesc_acc,
                              (j > 0 || k > 0) ? 1u : 0u);
            tc::tc_commit(qm_empty + s);
          }
          __syncwarp();
          TRACE(1, 4);
          ++qc; ++dt;
        }
        if (tc::elect_one()) tc::tc_commit(acc_full);
        __syncwarp();
      }
      TRACE_END(1);
    }
  } else {
    const int quarter = warp & 3, row = quarter * 32 + lane, tid = threadIdx.x - 64;
    const int half = (warp - 2) >> 2;            // which 32-column half of each tile this warpgroup owns
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float ks = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;
    uint32_t sc = 0, ic = 0;
    // Items are software-pipelined like in the dQ kernel: the next item's key-padding byte, the lse/delta of ALL its
    // queries and its first dropout word are requested before the current item's dK/dV read-out; the stats go into a
    // per-item smem table (two tables alternate), so the tile loop has no named barrier and no global-load latency.
    TRACE_DECL(lane == 0 && (warp == 2 || warp == 6));
#ifdef PA_ATTN_TRACE
    const int trole = warp == 2 ? 2 : 3;
#endif
    constexpr int SPT = 10;                              // table values per thread: 2 * LqPad <= 2560
    struct ItemRegs {
      int b, h, k0, i0, kj; uint8_t kp; float st[SPT]; uint32_t mw0;
    };
    auto load_mw_at = [&](int b_, int h_, int kj_, int q0n) -> uint32_t {   // keep-bits of 32 queries for this key (column plane)
      const int wi = (q0n + half * 32) >> 5;
      return (kj_ < p.LkW * 32 && wi < p.LqW)
                 ? __ldg(p.drop_cols + ((int64_t)(b_ * p.H + h_) * (p.LkW * 32) + kj_) * p.LqW + wi) : 0u;
    };
    auto prefetch = [&](int item, ItemRegs& r) {         // only ISSUES the loads
      coords(item, r.b, r.h, r.k0, r.i0);
      r.kj = r.k0 + row;
      r.kp = (p.kpm != nullptr && r.kj < p.Lk) ? p.kpm[(int64_t)r.b * p.Lk + r.kj] : (uint8_t)0;
      const int64_t rows = (int64_t)(r.b * p.H + r.h) * p.Lq;
#pragma unroll
      for (int m = 0; m < SPT; ++m) {
        const int v = tid + 256 * m;
        float x = 0.f;
        if (v < p.LqPad) x = v < p.Lq ? p.lse[rows + v] : -INFINITY;
        else if (v < 2 * p.LqPad) x = (v - p.LqPad) < p.Lq ? p.delta[rows + v - p.LqPad] : 0.f;
        r.st[m] = x;
      }
      r.mw0 = p.p_drop > 0.f ? load_mw_at(r.b, r.h, r.kj, r.i0 * C::BQ) : 0xffffffffu;
    };
    // stationary operands: this thread's row of K_j and V_j (its warpgroup's 32-column chunk) from the TMA tiles in shared
    // memory (128B swizzle: 16-byte piece i of row r sits at piece i ^ (r & 7)) into TMEM.  Called when these warps have
    // consumed every S^T / dP^T of the previous item, so the TMEM columns are free; the smem tiles are released at once.
    auto copy_kv = [&](uint32_t icn) {
      tc::mbar_wait(kv_full, icn & 1);
      if (half < C::kChunks) {
        const uint32_t krow = tc::smem_u32(smem + C::kOffK + half * (C::BKV * 128) + row * 128);
        const uint32_t vrow = tc::smem_u32(smem + C::kOffV + half * (C::BKV * 128) + row * 128);
        uint32_t rk[32], rv[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int off = (i ^ (row & 7)) << 4;
          const float4 a4 = tc::ld_shared_v4(krow + off);
          const float4 d4 = tc::ld_shared_v4(vrow + off);
          rk[4 * i] = __float_as_uint(a4.x); rk[4 * i + 1] = __float_as_uint(a4.y); rk[4 * i + 2] = __float_as_uint(a4.z); rk[4 * i + 3] = __float_as_uint(a4.w);
          rv[4 * i] = __float_as_uint(d4.x); rv[4 * i + 1] = __float_as_uint(d4.y); rv[4 * i + 2] = __float_as_uint(d4.z); rv[4 * i + 3] = __float_as_uint(d4.w);
        }
        tc::tmem_st_32x32(tmem_base + lane_addr + C::kColK + half * 32, rk);
        tc::tmem_st_32x32(tmem_base + lane_addr + C::kColV + half * 32, rv);
        tc::tmem_st_wait();
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) { tc::mbar_arrive(kt_full); tc::mbar_arrive(kv_empty); }
    };
    ItemRegs nx;
    if ((int)blockIdx.x < p.items) { prefetch(blockIdx.x, nx); copy_kv(0); }
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++ic) {
      const ItemRegs cu = nx;
      const int b = cu.b, h = cu.h, k0 = cu.k0, i0 = cu.i0, kj = cu.kj;
      const bool k_ok = kj < p.Lk && !cu.kp;
      const int n = q_tiles_all - i0;
      float* stat_it = stat_s + (ic & 1) * 2 * p.LqPad;   // [lse in the log2 domain (+inf kills the row) | delta]
#pragma unroll
      for (int m = 0; m < SPT; ++m) {
        const int v = tid + 256 * m;
        if (v < p.LqPad) stat_it[v] = cu.st[m] == -INFINITY ? INFINITY : cu.st[m] * kLog2e;
        else if (v < 2 * p.LqPad) stat_it[v] = cu.st[m];
      }
      uint32_t mw_pref = cu.mw0;                           // global loads run one tile ahead of their use
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int j = 0; j < n; ++j, ++sc) {
        const int buf = sc & 1, q0 = (i0 + j) * C::BQ;
        const uint32_t mw = mw_pref;
        TRACE(trole, 0);
        if (p.p_drop > 0.f && j + 1 < n) mw_pref = load_mw_at(b, h, kj, q0 + C::BQ);
        tc::mbar_wait(st_full + buf, (sc >> 1) & 1);
        TRACE(trole, 2);
        tc::tc_fence_after();
        const bool diag = p.causal && (q0 < k0 + C::BKV - 1);
        const float* lse2 = stat_it + q0;
        const float* dl = stat_it + p.LqPad + q0;
        const uint32_t lse2_u32 = tc::smem_u32(lse2), dl_u32 = tc::smem_u32(dl);
#pragma unroll
        for (int c0 = half * 32; c0 < half * 32 + 32; c0 += 32) {
          uint32_t rs[32], rd[32];
          tc::tmem_ld_32x32(tmem_base + lane_addr + C::kColS + buf * C::BQ + c0, rs);
          tc::tmem_ld_32x32(tmem_base + lane_addr + C::kColDP + buf * C::BQ + c0, rd);
          tc::tmem_ld_wait();
          TRACE(trole, 3);
          // rows of masked keys produce P = dS = 0; a tile below the causal diagonal needs no per-element test
          if (!k_ok) {
#pragma unroll
            for (int c = 0; c < 32; ++c) { rs[c] = 0u; rd[c] = 0u; }
          } else if (!diag) {
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              const float4 l4 = tc::ld_shared_v4(lse2_u32 + 4 * (c0 + c));     // per-query stats, 4 columns at a time (explicit LDS)
              const float4 d4 = tc::ld_shared_v4(dl_u32 + 4 * (c0 + c));
              const float lq[4] = {l4.x, l4.y, l4.z, l4.w}, dq4[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float mk = ((mw >> (c + e)) & 1u) ? ks : 0.f;
                const float pr = fast_exp2(fmaf(__uint_as_float(rs[c + e]), p.scale_log2, -lq[e]));
                const float ds = pr * fmaf(__uint_as_float(rd[c + e]), mk, -dq4[e]);
                rs[c + e] = tf32_rn_finite_bits(pr * mk);
                rd[c + e] = tf32_rn_finite_bits(ds);
              }
            }
          } else {
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              float mk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) mk[e] = ((mw >> (c + e)) & 1u) ? ks : 0.f;
              const float4 l4 = tc::ld_shared_v4(lse2_u32 + 4 * (c0 + c));
              const float4 d4 = tc::ld_shared_v4(dl_u32 + 4 * (c0 + c));
              const float lq[4] = {l4.x, l4.y, l4.z, l4.w}, dq4[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int qq = c0 + c + e;
                float v = fmaf(__uint_as_float(rs[c + e]), p.scale_log2, -lq[e]);
                if (kj > q0 + qq) v = -INFINITY;
                const float pr = fast_exp2(v);
                const float ds = pr * fmaf(__uint_as_float(rd[c + e]), mk[e], -dq4[e]);
                rs[c + e] = tf32_rn_finite_bits(pr * mk[e]);
                rd[c + e] = tf32_rn_finite_bits(ds);
              }
            }
          }
          TRACE(trole, 4);
          tc::tmem_st_32x32(tmem_base + lane_addr + C::kColS + buf * C::BQ + c0, rs);
          tc::tmem_st_32x32(tmem_base + lane_addr + C::kColDP + buf * C::BQ + c0, rd);
        }
        tc::tmem_st_wait();
        TRACE(trole, 5);
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(pds_full + buf);
        TRACE(trole, 6);
      }
      if (item + (int)gridDim.x < p.items) {
        prefetch(item + gridDim.x, nx);                  // next item's loads fly during the read-out
        copy_kv(ic + 1);                                 // and its K / V go into TMEM: the tensor pipe starts on it meanwhile
      }
      tc::mbar_wait(acc_full, ic & 1);
      TRACE(trole, 7);
      tc::tc_fence_after();
      const bool row_ok = kj < p.Lk;
      float* outk = p.dk + ((int64_t)b * p.Lk + kj) * p.lddk + h * DH;
      float* outv = p.dv + ((int64_t)b * p.Lk + kj) * p.lddv + h * DH;
#pragma unroll
      for (int c0 = half * 32; c0 < DH; c0 += 64) {
        uint32_t rk[32], rv[32];
        tc::tmem_ld_32x32(tmem_base + lane_addr + C::kColDK + c0, rk);
        tc::tmem_ld_32x32(tmem_base + lane_addr + C::kColDV + c0, rv);
        tc::tmem_ld_wait();
        TRACE(trole, 14);
        if (row_ok) {
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            float4 a = make_float4(__uint_as_float(rk[c]) * p.scale, __uint_as_float(rk[c + 1]) * p.scale,
                                   __uint_as_float(rk[c + 2]) * p.scale, __uint_as_float(rk[c + 3]) * p.scale);
            float4 a2 = make_float4(__uint_as_float(rk[c + 4]) * p.scale, __uint_as_float(rk[c + 5]) * p.scale,
                                    __uint_as_float(rk[c + 6]) * p.scale, __uint_as_float(rk[c + 7]) * p.scale);
            float4 v = make_float4(__uint_as_float(rv[c]), __uint_as_float(rv[c + 1]), __uint_as_float(rv[c + 2]), __uint_as_float(rv[c + 3]));
            float4 v2 = make_float4(__uint_as_float(rv[c + 4]), __uint_as_float(rv[c + 5]), __uint_as_float(rv[c + 6]), __uint_as_float(rv[c + 7]));
            if (n == 0) { a = make_float4(0.f, 0.f, 0.f, 0.f); a2 = a; v = a; v2 = a; }
            if (p.round_out) { a = tf32_rn4(a); a2 = tf32_rn4(a2); v = tf32_rn4(v); v2 = tf32_rn4(v2); }
            if (p.wide_st) {
              st_global_v8(outk + c0 + c, a, a2);
              st_global_v8(outv + c0 + c, v, v2);
            } else {
              *reinterpret_cast<float4*>(outk + c0 + c) = a; *reinterpret_cast<float4*>(outk + c0 + c + 4) = a2;
              *reinterpret_cast<float4*>(outv + c0 + c) = v; *reinterpret_cast<float4*>(outv + c0 + c + 4) = v2;
            }
          }
        }
        TRACE(trole, 15);
        if (p.dbias != nullptr) {          // k- and v-bias gradients: column sums over this warp's 32 key rows
          float ck[32], cv[32];
          const bool use = row_ok && n > 0;
#pragma unroll
          for (int c = 0; c < 32; ++c) { ck[c] = use ? __uint_as_float(rk[c]) * p.scale : 0.f; cv[c] = use ? __uint_as_float(rv[c]) : 0.f; }
          const float tk = warp_colsum32(ck, lane), tv = warp_colsum32(cv, lane);
          const int dmodel = p.H * DH;
          atomicAdd(p.dbias + dmodel + h * DH + c0 + lane, tk);
          atomicAdd(p.dbias + 2 * dmodel + h * DH + c0 + lane, tv);
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(acc_empty);
      TRACE(trole, 8);
    }
#ifdef PA_ATTN_TRACE
    if (trace_on) g_bwd_trace_n[trole] = trace_n;
#endif
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<C::kTmemCols>(tmem_base);
}

template <int DH>
int launch(const pa_attn_bwd_args& a, cudaStream_t st) {
  int rc = pa_attn_delta_launch(a.o, a.d_o, a.ldo, a.B, a.H, a.Lq, a.dh, a.delta, st);
  if (rc) return rc;
  const uint64_t d = (uint64_t)a.H * DH, rq = (uint64_t)a.B * a.Lq, rk = (uint64_t)a.B * a.Lk;
  BwdParams p{};
  p.lse = a.lse; p.delta = a.delta; p.kpm = a.kpm;
  p.dq = a.dq; p.dk = a.dk; p.dv = a.dv; p.lddq = a.lddq; p.lddk = a.lddk; p.lddv = a.lddv;
  p.B = a.B; p.H = a.H; p.Lq = a.Lq; p.Lk = a.Lk; p.causal = a.causal; p.round_out = a.round_out;
  p.scale = a.scale; p.scale_log2 = a.scale * kLog2e;
  p.p_drop = a.p_drop; p.drop_rows = a.drop_rows; p.drop_cols = a.drop_cols; p.dbias = a.dbias;
  p.LkW = (a.Lk + 31) / 32; p.LqW = (a.Lq + 31) / 32;
  p.kv_len = a.kpm != nullptr ? a.kv_len : nullptr;
  p.wide_st = ((((uintptr_t)a.dq | (uintptr_t)a.dk | (uintptr_t)a.dv) & 31) == 0 && a.lddq % 8 == 0 && a.lddk % 8 == 0 && a.lddv % 8 == 0) ? 1 : 0;
#ifdef PA_ATTN_TRACE
  { const char* dbg = getenv("PLANK_B200_ATTN_DEBUG"); p.trace = dbg ? (atoi(dbg) & (1024 | 2048)) : 0; }
  { const char* e = getenv("PLANK_B200_ATTN_L2PF"); p.l2_prefetch = e ? atoi(e) : 0; }
#endif
  if (a.p_drop > 0.f && (a.drop_rows == nullptr || a.drop_cols == nullptr)) {
    pa_set_error("pa_attn_bwd (tc): p_drop > 0 needs drop_rows/drop_cols from pa_dropout_mask");
    return PA_ERR_ARG;
  }
  {
    using C = CfgQ<DH>;
    CUtensorMap tq, tdo, tk, tkm, tv;
    if ((rc = pa_make_tmap_2d(&tq, a.q, d, rq, (uint64_t)a.ldq * 4, 32, C::BQ))) return rc;
    if ((rc = pa_make_tmap_2d(&tdo, a.d_o, d, rq, (uint64_t)a.ldo * 4, 32, C::BQ))) return rc;
    if ((rc = pa_make_tmap_2d(&tk, a.k, d, rk, (uint64_t)a.ldk * 4, 32, C::BK))) return rc;
    if ((rc = pa_make_tmap_2d(&tkm, a.k, d, rk, (uint64_t)a.ldk * 4, 32, C::BK, true))) return rc;
    if ((rc = pa_make_tmap_2d(&tv, a.v, d, rk, (uint64_t)a.ldv * 4, 32, C::BK))) return rc;
    auto kern = attn_bwd_dq_tc_kernel<DH>;
    p.LkPad = (a.Lk + C::BK - 1) / C::BK * C::BK;
    const int smem_q = C::kSmemFixed + 2 * p.LkPad * 4;
    if (smem_q > 227 * 1024 || p.LkPad > 2048) { pa_set_error("pa_attn_bwd (tc): Lk = %d too long for the dQ kernel's bias table", a.Lk); return PA_ERR_UNSUPPORTED; }
    static SmemAttrCache attr_q;
    if ((rc = pa_set_max_smem(kern, smem_q, attr_q))) return rc;
    p.tiles = (a.Lq + C::BQ - 1) / C::BQ;
    p.items = p.tiles * a.H * a.B;
    kern<<<p.items < pa_num_sms() ? p.items : pa_num_sms(), kThreads, smem_q, st>>>(tq, tdo, tk, tkm, tv, p);
    PA_CHECK_LAUNCH();
  }
  {
    using C = CfgK<DH>;
    CUtensorMap tk, tv, tq, tqm, tdo, tdom;
    if ((rc = pa_make_tmap_2d(&tk, a.k, d, rk, (uint64_t)a.ldk * 4, 32, C::BKV))) return rc;
    if ((rc = pa_make_tmap_2d(&tv, a.v, d, rk, (uint64_t)a.ldv * 4, 32, C::BKV))) return rc;
    if ((rc = pa_make_tmap_2d(&tq, a.q, d, rq, (uint64_t)a.ldq * 4, 32, C::BQ))) return rc;
    if ((rc = pa_make_tmap_2d(&tqm, a.q, d, rq, (uint64_t)a.ldq * 4, 32, C::BQ, true))) return rc;
    if ((rc = pa_make_tmap_2d(&tdo, a.d_o, d, rq, (uint64_t)a.ldo * 4, 32, C::BQ))) return rc;
    if ((rc = pa_make_tmap_2d(&tdom, a.d_o, d, rq, (uint64_t)a.ldo * 4, 32, C::BQ, true))) return rc;
    auto kern = attn_bwd_dkdv_tc_kernel<DH>;
    p.LqPad = (a.Lq + C::BQ - 1) / C::BQ * C::BQ;
    const int smem_k = C::kSmemFixed + 2 * 2 * p.LqPad * 4;
    if (smem_k > 227 * 1024 || 2 * p.LqPad > 2560) { pa_set_error("pa_attn_bwd (tc): Lq = %d too long for the dK/dV kernel's stat table", a.Lq); return PA_ERR_UNSUPPORTED; }
    static SmemAttrCache attr_k;
    if ((rc = pa_set_max_smem(kern, smem_k, attr_k))) return rc;
    p.tiles = (a.Lk + C::BKV - 1) / C::BKV;
    p.items = p.tiles * a.H * a.B;
    kern<<<p.items < pa_num_sms() ? p.items : pa_num_sms(), kThreads, smem_k, st>>>(tk, tv, tq, tqm, tdo, tdom, p);
    PA_CHECK_LAUNCH();
  }
  return PA_OK;
}

}  // namespace

#ifdef PA_ATTN_TRACE
extern "C" int pa_debug_attn_bwd_trace(unsigned long long* out_host /*[4][1024]*/, int* n_host /*[4]*/) {
  PA_CUDA(cudaMemcpyFromSymbol(out_host, g_bwd_trace, sizeof(unsigned long long) * 4 * 1024));
  PA_CUDA(cudaMemcpyFromSymbol(n_host, g_bwd_trace_n, sizeof(int) * 4));
  return PA_OK;
}
#endif

int pa_attn_bwd_tc(const pa_attn_bwd_args* a, void* stream) {
  switch (a->dh) {
    case 32: return launch<32>(*a, (cudaStream_t)stream);
    case 64: return launch<64>(*a, (cudaStream_t)stream);
    default: pa_set_error("pa_attn_bwd (tc): head dim %d unsupported (32, 64)", a->dh); return PA_ERR_UNSUPPORTED;
  }
}
