// Flash-style attention forward for head_dim 64 on tcgen05 / TMEM (sm_100a).
//
//   O = softmax(Q K^T * scale + mask) V        per (batch, head); bf16 in/out, fp32 softmax
//
// Replaces F.scaled_dot_product_attention in MultiHeadAttention.forward (olmoasr/model.py:331-340):
//   encoder self-attention (no mask), decoder self-attention (causal + per-sample key length, derived
//   from the dense additive mask the reference passes, model.py:740-743) and cross-attention (no mask).
//
// CTA = 256 query rows (two 128-row tiles) x one (b, h); loops over 128-key tiles.  384 threads:
//   warps 0..3   softmax warpgroup of query tile 0        warps 4..7   softmax warpgroup of query tile 1
//   warp 8       TMA producer (Q once, K/V in a 3-stage ring; every K/V tile serves both query tiles)
//   warp 9       tcgen05.mma issuer + TMEM owner           warps 10, 11 idle (keep the control warpgroup 4-aligned)
// setmaxnreg moves registers from the control warpgroup (56) to the softmax warpgroups (224) so that a thread can
// hold its whole 128-column score row: it reads the row from TMEM ONCE, immediately releases the S buffer
// (bar_sfree), and the MMA warp issues Q K^T of the NEXT key tile while this tile's softmax is still running.
// The softmax warpgroups therefore run back to back (MUFU-bound) and the tensor pipe works underneath them.
//
// Thread r of a warpgroup owns score row r (= TMEM lane r): the row maximum and sum need no shuffles.
// TMEM (512 cols): S0 0..127 | S1 128..255 | O0 256..319 | O1 320..383 | P0 384..447 | P1 448..511.  O is accumulated by the P*V MMAs
// directly in TMEM; it is rescaled (tcgen05.ld -> scale -> tcgen05.st) only when a row maximum grows by more
// than 2^8 -- otherwise the stale maximum is kept (probabilities stay <= 256; exact after the final 1/l).
// Scale/subtract and the row sum use packed FFMA2 / FADD2, the maximum FMNMX3.  P goes to smem as bf16 in the
// 128B-swizzled K-major layout and is the A operand of the P*V MMA; V is consumed MN-major from its TMA tile.
#include "common.cuh"
#include "ptx_sm100.cuh"

#include <type_traits>

namespace oasr {
namespace {

constexpr int HD = 64;
constexpr int BQ = 128;                      // rows per query tile (two tiles per CTA)
constexpr int BKV = 128;
constexpr int KV_STAGES = 3;
constexpr int TILE_BYTES = 128 * HD * 2;     // 16 KB: a [128 rows][64 bf16] swizzled tile
constexpr int P_BYTES = BQ * BKV * 2;        // 32 KB: two 64-key halves of [128][128B]
// P (bf16) as the A operand of the P V MMA: 1 = written back to TMEM (tcgen05.st, consumed by a TMEM-A MMA), 0 = through
// swizzled shared memory.  Through smem a key tile moves 256 KB over the 128 B/clk shared-memory port (Q K^T operands
// 64, P stores 64, P V operands 96, K/V TMA writes 32); in TMEM the P stores and the P operand reads (128 KB) disappear.
// Measured (profiles/r02_attention_fwd_ptmem_ab.txt): parity green, but 0.528 ms vs 0.485 ms for the encoder shape -- the
// forward is paced by the softmax warps (MUFU), not by the shared-memory port, and tcgen05.st + wait::st is a longer
// tail than 16 STS.128.  Kept as an A/B switch, default off.
#ifndef OASR_FWD_P_TMEM
#define OASR_FWD_P_TMEM 0
#endif
constexpr bool P_TMEM = OASR_FWD_P_TMEM != 0;
// The two softmax warpgroups share the SM's MUFU units.  1 = they take turns in the exponential phase (named barriers
// 2 / 3, the FlashAttention-3 "ping-pong"): one warpgroup's TMEM reads / row maximum / O rescale then run under the
// other's exponentials instead of both idling the MUFU pipe at the same time.
#ifndef OASR_FWD_PINGPONG
#define OASR_FWD_PINGPONG 0
#endif
constexpr bool PINGPONG = OASR_FWD_PINGPONG != 0;
constexpr int ATT_TILES = 2 * TILE_BYTES /*Q0,Q1*/ + KV_STAGES * 2 * TILE_BYTES /*K,V*/ + (P_TMEM ? 0 : 2 * P_BYTES);  // 128 / 192 KB
constexpr int ATT_SMEM = ATT_TILES + 256;
constexpr int TMEM_COLS = 512;
constexpr int S_COL = 0, O_COL = 256, P_COL = 384;   // + t*128 / + t*64 / + t*64 (128 bf16 keys = 64 columns)
constexpr float RESCALE_LOG2 = 8.0f;
constexpr int SOFTMAX_REGS = 224, CONTROL_REGS = 56;   // 8 * SOFTMAX + 4 * CONTROL == 12 * 168
// Share of the exponentials of an unmasked tile evaluated on the FMA pipe (exp2_poly2) instead of MUFU.EX2:
// 0 = none, 1 = every fourth pair (25 %), 2 = every second pair (50 %).  A/B builds: python -m olmoasr_b200.build --variant.
#ifndef OASR_ATTN_POLY
#define OASR_ATTN_POLY 0
#endif

#ifdef OASR_ATTN_TRACE
// Debug build only: one CTA in the middle of the grid stamps clock64() at phase boundaries (tools/trace_attention.py).
__device__ unsigned long long g_fwd_trace[4][512];
#define TRACE_DECL(role_)                                                                                     \
  const bool tr_on = (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) == (gridDim.x * gridDim.y * gridDim.z) / 2 + gridDim.x / 2 && \
                     lane == 0 && (role_) >= 0;                                                               \
  const int tr_role = (role_) < 0 ? 0 : (role_);                                                              \
  int tr_n = 0;
#define TR(id_)                                                                                               \
  do {                                                                                                        \
    if (tr_on && tr_n < 511) g_fwd_trace[tr_role][++tr_n] = (static_cast<unsigned long long>(id_) << 48) | (clock64() & 0xffffffffffffull); \
  } while (0)
#define TRACE_END()                                                                                           \
  do {                                                                                                        \
    if (tr_on) g_fwd_trace[tr_role][0] = tr_n;                                                                \
  } while (0)
#else
#define TRACE_DECL(role_)
#define TR(id_)
#define TRACE_END()
#endif

struct AttnParams {
  bf16* o;
  float* lse;        // (B, H, Tq) log2-domain log-sum-exp, nullable
  const int32_t* kv_len;  // (B,) valid keys per sample, nullable
  int64_t ldo;
  int B, H, Tq, Tkv;
  int causal;
  float scale_log2;  // scale * log2(e)
};

__global__ void __launch_bounds__(384, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + ATT_TILES);
  uint64_t& bar_q = bars[0];
  uint64_t* bar_s = bars + 1;                 // [2] S_t ready (MMA commit)
  uint64_t* bar_sfree = bars + 3;             // [2] S_t read into registers (4 warp arrivals)
  uint64_t* bar_p = bars + 5;                 // [2] P_t in smem, O_t rescaled (4 warp arrivals)
  uint64_t* bar_pv = bars + 7;                // [2] P_t V retired (MMA commit)
  uint64_t* bar_kv_full = bars + 9;           // [KV_STAGES]
  uint64_t* bar_kv_empty = bars + 9 + KV_STAGES;
  uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(bars + 9 + 2 * KV_STAGES);

  const uint32_t sbase = ptx::smem_u32(smem_raw);
  if ((sbase & 1023u) != 0) {  // swizzled tiles need 1 KB alignment; the declaration above should guarantee it
    if (threadIdx.x == 0) printf("oasr attention: dynamic smem base %u not 1 KB aligned\n", sbase);
    __trap();
  }
  const uint32_t sQ = sbase;                          // tile t at sQ + t*16K
  const uint32_t sK0 = sQ + 2 * TILE_BYTES;           // stage s: K at sK0 + s*32K, V right after K
  const uint32_t sP = sK0 + KV_STAGES * 2 * TILE_BYTES;  // tile t at sP + t*32K
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q_base = blockIdx.x * 2 * BQ;

  int kv_valid = p.Tkv;
  if (p.kv_len) kv_valid = min(kv_valid, max(1, p.kv_len[b]));
  int n_kv[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int q0 = q_base + t * BQ;
    int kv_end = kv_valid;
    if (p.causal) kv_end = min(kv_end, q0 + BQ);
    n_kv[t] = (q0 < p.Tq) ? (kv_end + BKV - 1) / BKV : 0;
  }
  const int n_max = max(n_kv[0], n_kv[1]);

  if (warp == 8 && lane == 0) {
    ptx::tma_prefetch_desc(&tmQ); ptx::tma_prefetch_desc(&tmK); ptx::tma_prefetch_desc(&tmV);
    ptx::mbar_init(ptx::smem_u32(&bar_q), 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bar_s[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bar_sfree[s]), 4);
      ptx::mbar_init(ptx::smem_u32(&bar_p[s]), 4);
      ptx::mbar_init(ptx::smem_u32(&bar_pv[s]), 1);
    }
    for (int s = 0; s < KV_STAGES; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bar_kv_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bar_kv_empty[s]), 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 9) {
    ptx::tmem_alloc<TMEM_COLS>(ptx::smem_u32(&tmem_slot));
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp >= 8) {
    ptx::setmaxnreg_dec<CONTROL_REGS>();
    if (warp == 8 && lane == 0) {
      // ------------------------------ TMA producer ------------------------------
      ptx::mbar_arrive_expect_tx(ptx::smem_u32(&bar_q), 2 * TILE_BYTES);
      ptx::tma_load_2d(sQ, &tmQ, ptx::smem_u32(&bar_q), h * HD, b * p.Tq + q_base);  // box 64 x 256 rows
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_max; ++j) {
        ptx::mbar_wait(ptx::smem_u32(&bar_kv_empty[s]), ph ^ 1);
        const uint32_t full = ptx::smem_u32(&bar_kv_full[s]);
        ptx::mbar_arrive_expect_tx(full, 2 * TILE_BYTES);
        const int krow = b * p.Tkv + j * BKV;
        ptx::tma_load_2d(sK0 + s * 2 * TILE_BYTES, &tmK, full, h * HD, krow);
        ptx::tma_load_2d(sK0 + s * 2 * TILE_BYTES + TILE_BYTES, &tmV, full, h * HD, krow);
        if (++s == KV_STAGES) { s = 0; ph ^= 1; }
      }
    } else if (warp == 9 && n_max > 0) {
      // ------------------------------ MMA issuer (whole warp convergent; elect.sync picks the issuing lane) ------------------------------
      TRACE_DECL(1)
      TR(1);
      constexpr uint32_t idesc_qk = ptx::umma_idesc_bf16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_pv = ptx::umma_idesc_bf16(BQ, HD, 0, 1);
      uint32_t q_lo[2], p_lo[2], k_lo0, v_lo0, hi_k, hi_v, unused;
      ptx::umma_desc_sw128_lh(sQ, 16, 1024, q_lo[0], hi_k);
      ptx::umma_desc_sw128_lh(sQ + TILE_BYTES, 16, 1024, q_lo[1], unused);
      ptx::umma_desc_sw128_lh(sP, 16, 1024, p_lo[0], unused);
      ptx::umma_desc_sw128_lh(sP + P_BYTES, 16, 1024, p_lo[1], unused);
      ptx::umma_desc_sw128_lh(sK0, 16, 1024, k_lo0, unused);
      ptx::umma_desc_sw128_lh(sK0 + TILE_BYTES, BKV * 128, 1024, v_lo0, hi_v);
      constexpr uint32_t STAGE_LO = (2 * TILE_BYTES) >> 4;
      auto issue_qk = [&](int t, int stage) {
        const uint32_t klo = k_lo0 + stage * STAGE_LO;
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            ptx::tc_mma_f16_lh(tmem + S_COL + t * BKV, q_lo[t] + k * 2, hi_k, klo + k * 2, hi_k, idesc_qk, k > 0);
          ptx::tc_commit(ptx::smem_u32(&bar_s[t]));
        }
        __syncwarp();
      };
      auto issue_pv = [&](int t, int stage, int j) {
        const uint32_t vlo = v_lo0 + stage * STAGE_LO;
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k) {
            if (P_TMEM)   // 16 keys = 8 TMEM columns of packed bf16 pairs
              ptx::tc_mma_f16_ts_lh(tmem + O_COL + t * HD, tmem + P_COL + t * (BKV / 2) + k * 8, vlo + k * (2048 >> 4), hi_v,
                                    idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
            else
              ptx::tc_mma_f16_lh(tmem + O_COL + t * HD, p_lo[t] + (k >> 2) * ((P_BYTES / 2) >> 4) + (k & 3) * 2, hi_k,
                                 vlo + k * (2048 >> 4), hi_v, idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
          }
          ptx::tc_commit(ptx::smem_u32(&bar_pv[t]));
        }
        __syncwarp();
      };
      ptx::mbar_wait(ptx::smem_u32(&bar_q), 0);
      ptx::mbar_wait(ptx::smem_u32(&bar_kv_full[0]), 0);
      ptx::tc_fence_after();
#pragma unroll
      for (int t = 0; t < 2; ++t)
        if (n_kv[t] > 0) issue_qk(t, 0);
      TR(2);
      int st = 0;            // ring stage of key tile j
      uint32_t st_ph = 0;    // its phase
      for (int j = 0; j < n_max; ++j) {
        int st_n = st + 1;
        uint32_t ph_n = st_ph;
        if (st_n == KV_STAGES) { st_n = 0; ph_n ^= 1; }
        // (1) next tile's scores as soon as this tile's S has been read into registers
        if (j + 1 < n_max) {
          ptx::mbar_wait(ptx::smem_u32(&bar_kv_full[st_n]), ph_n);
          TR(3);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (j + 1 < n_kv[t]) {
              ptx::mbar_wait(ptx::smem_u32(&bar_sfree[t]), j & 1);
              TR(4 + 2 * t);
              ptx::tc_fence_after();
              issue_qk(t, st_n);
              TR(5 + 2 * t);
            }
          }
        }
        // (2) this tile's P V
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (j < n_kv[t]) {
            ptx::mbar_wait(ptx::smem_u32(&bar_p[t]), j & 1);   // P_t(j) in smem, O_t rescaled
            TR(8 + 2 * t);
            ptx::tc_fence_after();
            issue_pv(t, st, j);
            TR(9 + 2 * t);
          }
        }
        if (ptx::elect_one()) ptx::tc_commit(ptx::smem_u32(&bar_kv_empty[st]));   // K/V stage free once every MMA issued so far retires
        __syncwarp();
        st = st_n; st_ph = ph_n;
      }
      TRACE_END();
    }
  } else {
    // ----------------------------- softmax / output warpgroups -----------------------------
    ptx::setmaxnreg_inc<SOFTMAX_REGS>();
    const int t = warp >> 2;                         // query tile of this warpgroup
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;               // row within the tile == TMEM lane
    const int qi = q_base + t * BQ + r;              // query index within the sequence
    const uint32_t t_lane = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t t_s = tmem + t_lane + S_COL + t * BKV;
    const uint32_t t_o = tmem + t_lane + O_COL + t * HD;
    const uint32_t sPt = sP + t * P_BYTES;
    const uint32_t t_p = tmem + t_lane + P_COL + t * (BKV / 2);
    const float c = p.scale_log2;
    const int n_t = n_kv[t];
    const int n_common = min(n_kv[0], n_kv[1]);      // key tiles on which both warpgroups work: turns are taken there
    float m = -INFINITY, l = 0.f;
    TRACE_DECL(warp == 0 ? 0 : (warp == 4 ? 3 : -1))
    TR(1);
    if (PINGPONG && t == 1 && n_common > 0) asm volatile("bar.arrive 2, 256;" ::: "memory");   // warpgroup 0 goes first

    for (int j = 0; j < n_t; ++j) {
      TR(10);
      ptx::mbar_wait(ptx::smem_u32(&bar_s[t]), j & 1);
      TR(11);
      ptx::tc_fence_after();
      float v[BKV];
      {
        uint32_t (&u)[BKV] = reinterpret_cast<uint32_t (&)[BKV]>(v);
#pragma unroll
        for (int cc = 0; cc < BKV / 32; ++cc)
          ptx::tc_ld_32x32b_x32(t_s + cc * 32, reinterpret_cast<uint32_t (&)[32]>(u[cc * 32]));
        ptx::tc_wait_ld();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bar_sfree[t]));   // S_t may be overwritten by Q K^T (j+1)
      TR(12);

      const int k0 = j * BKV;
      int limit = kv_valid - k0;                     // keys [0, limit) of this tile are visible
      if (p.causal) limit = min(limit, qi - k0 + 1);
      if (limit < BKV) {
#pragma unroll
        for (int i = 0; i < BKV; ++i)
          if (i >= limit) v[i] = -INFINITY;
      }
      // 8 independent FMNMX3 chains (a single chain of 63 dependent ops costs ~250 cycles of pure latency per tile)
      float mxs[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) mxs[u] = fmaxf(v[2 * u], v[2 * u + 1]);
#pragma unroll
      for (int i = 16; i < BKV; i += 16)
#pragma unroll
        for (int u = 0; u < 8; ++u) mxs[u] = fmaxf(fmaxf(mxs[u], v[i + 2 * u]), v[i + 2 * u + 1]);
      const float mx = fmaxf(fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3])), fmaxf(fmaxf(mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7])));
      const float m_new = fmaxf(m, mx);
      const bool need = (j == 0) || ((m_new - m) * c > RESCALE_LOG2);
      const float m_next = need ? m_new : m;
      TR(13);
      if (j > 0) {
        ptx::mbar_wait(ptx::smem_u32(&bar_pv[t]), (j - 1) & 1);   // P_t V (j-1) retired: O_t valid, sP_t reusable
        TR(14);
        ptx::tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? fast_exp2((m - m_next) * c) : 1.0f;
#pragma unroll 1
          for (int cc = 0; cc < HD / 16; ++cc) {
            uint32_t o[16];
            ptx::tc_ld_32x32b_x16(t_o + cc * 16, o);
            ptx::tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            ptx::tc_st_32x32b_x16(t_o + cc * 16, o);
          }
          ptx::tc_wait_st();
          l *= alpha;
        }
      }
      m = m_next;
      TR(15);
      // p = 2^(s*c - m*c) (packed FFMA2), row sum (FADD2), bf16 -> swizzled smem (A operand of P V)
      const float neg = (m == -INFINITY) ? 0.f : -m * c;
      const float2 c2 = make_float2(c, c), n2 = make_float2(neg, neg);
      float2 sums[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      // POLY: odd pairs go through the FMA-pipe polynomial, even pairs through MUFU.EX2 (the two pipes run side by
      // side).  Masked tiles keep MUFU for every element so that -inf maps to exactly 0.
      auto emit_p = [&](auto poly_tag) {
        constexpr int POLY = decltype(poly_tag)::value;
        uint32_t pk[32];   // P_TMEM: 64 keys of this row as packed bf16 pairs
#pragma unroll
        for (int cc = 0; cc < BKV / 32; ++cc) {
          const uint32_t half_base = sPt + (cc >> 1) * (P_BYTES / 2) + r * 128;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = cc * 32 + q4 * 8 + e * 2;
              const float2 x = __ffma2_rn(make_float2(v[i], v[i + 1]), c2, n2);
              float2 pe;
              if ((POLY == 2 && (e & 1)) || (POLY == 1 && e == 3)) pe = exp2_poly2(x);
              else pe = make_float2(fast_exp2(x.x), fast_exp2(x.y));
              sums[e] = __fadd2_rn(sums[e], pe);
              w[e] = pack_bf16x2(pe.x, pe.y);
            }
            if (P_TMEM) {
#pragma unroll
              for (int e = 0; e < 4; ++e) pk[(cc & 1) * 16 + q4 * 4 + e] = w[e];
            } else {
              const int chunk = (cc & 1) * 4 + q4;       // 16-byte chunk index within the 128-byte row
              const uint32_t addr = half_base + ((chunk ^ (r & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                           : "memory");
            }
          }
          if (P_TMEM && (cc & 1)) ptx::tc_st_32x32b_x32(t_p + (cc >> 1) * 32, pk);
        }
        if (P_TMEM) ptx::tc_wait_st();
      };
      // ncu of the MUFU-only build (profiles/r01_ncu_kernel_metrics.txt): XU pipe 59.5 %, issue slots 36.7 %, tensor
      // pipe 29.7 % -- 128 x 128 ex2 per tile at 16 / clk / SM is twice the tile's MMA time, so a share of the
      // exponentials moves to the FMA pipe (OASR_ATTN_POLY; measured per build in profiles/r02_attention_poly_ab.txt)
      if (PINGPONG && j < n_common) {
        if (t == 0) asm volatile("bar.sync 2, 256;" ::: "memory");
        else asm volatile("bar.sync 3, 256;" ::: "memory");
      }
      // (warp-uniform choice: with P_TMEM both paths contain .sync.aligned tcgen05.st)
      if (OASR_ATTN_POLY != 0 && __all_sync(0xffffffffu, limit >= BKV)) emit_p(std::integral_constant<int, OASR_ATTN_POLY>{});
      else emit_p(std::integral_constant<int, 0>{});
      if (PINGPONG) {
        if (t == 0 && j < n_common) asm volatile("bar.arrive 3, 256;" ::: "memory");
        if (t == 1 && j + 1 < n_common) asm volatile("bar.arrive 2, 256;" ::: "memory");
      }
      if (!P_TMEM) ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bar_p[t]));
      TR(16);
      l += ((sums[0].x + sums[0].y) + (sums[1].x + sums[1].y)) + ((sums[2].x + sums[2].y) + (sums[3].x + sums[3].y));
    }
    if (n_t > 0) {
      ptx::mbar_wait(ptx::smem_u32(&bar_pv[t]), (n_t - 1) & 1);
      ptx::tc_fence_after();
      const float inv = 1.f / l;
      bf16* dst = p.o + (static_cast<int64_t>(b) * p.Tq + qi) * p.ldo + h * HD;
#pragma unroll
      for (int cc = 0; cc < HD / 32; ++cc) {
        uint32_t o[32];
        ptx::tc_ld_32x32b_x32(t_o + cc * 32, o);
        ptx::tc_wait_ld();
        if (qi < p.Tq) {
#pragma unroll
          for (int q8 = 0; q8 < 4; ++q8) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(o[8 * q8 + 0]) * inv, __uint_as_float(o[8 * q8 + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(o[8 * q8 + 2]) * inv, __uint_as_float(o[8 * q8 + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(o[8 * q8 + 4]) * inv, __uint_as_float(o[8 * q8 + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(o[8 * q8 + 6]) * inv, __uint_as_float(o[8 * q8 + 7]) * inv);
            reinterpret_cast<uint4*>(dst + cc * 32)[q8] = u;
          }
        }
      }
      if (qi < p.Tq && p.lse) p.lse[(static_cast<int64_t>(b) * p.H + h) * p.Tq + qi] = m * c + log2f(l);
    }
    TR(30);
    TRACE_END();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem);
  }
}

}  // namespace
}  // namespace oasr

using namespace oasr;

#ifdef OASR_ATTN_TRACE
extern "C" OASR_API int oasr_debug_fwd_trace(unsigned long long* host_out) {   // 4 x 512 words
  return cudaMemcpyFromSymbol(host_out, g_fwd_trace, sizeof(unsigned long long) * 4 * 512) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int oasr_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                  void* o, int64_t ldo, float* lse, int64_t B, int64_t H, int64_t Tq, int64_t Tkv,
                                  int64_t head_dim, int causal, const int32_t* kv_len, float scale, void* stream) {
  OASR_REQUIRE(head_dim == HD, "attention: head_dim %ld unsupported (every OLMoASR variant uses 64)", (long)head_dim);
  OASR_REQUIRE(B > 0 && H > 0 && Tq > 0 && Tkv > 0, "attention: empty problem");
  OASR_REQUIRE((ldq & 7) == 0 && (ldk & 7) == 0 && (ldv & 7) == 0 && (ldo & 7) == 0, "attention: strides must be multiples of 8");
  OASR_REQUIRE(!causal || Tq == Tkv, "attention: causal needs Tq == Tkv");
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_tmap_2d(&tmQ, q, 2, (uint64_t)(H * HD), (uint64_t)(B * Tq), (uint64_t)ldq * 2, HD, 2 * BQ, true))) return rc;
  if ((rc = make_tmap_2d(&tmK, k, 2, (uint64_t)(H * HD), (uint64_t)(B * Tkv), (uint64_t)ldk * 2, HD, BKV, true))) return rc;
  if ((rc = make_tmap_2d(&tmV, v, 2, (uint64_t)(H * HD), (uint64_t)(B * Tkv), (uint64_t)ldv * 2, HD, BKV, true))) return rc;
  AttnParams p;
  p.o = (bf16*)o; p.lse = lse; p.kv_len = kv_len; p.ldo = ldo;
  p.B = (int)B; p.H = (int)H; p.Tq = (int)Tq; p.Tkv = (int)Tkv; p.causal = causal;
  p.scale_log2 = scale * 1.4426950408889634f;
  static bool attr_set = false;
  if (!attr_set) {
    OASR_CUDA_OK(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    attr_set = true;
  }
  dim3 grid((unsigned)ceil_div(Tq, 2 * BQ), (unsigned)H, (unsigned)B);
  attention_fwd_kernel<<<grid, 384, ATT_SMEM, (cudaStream_t)stream>>>(tmQ, tmK, tmV, p);
  OASR_LAUNCH_CHECK();
  return OASR_OK;
}
