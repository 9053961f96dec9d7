// Flash-style attention forward for head_dim 64 on tcgen05 / TMEM (sm_100a).
//
//   O = softmax(Q K^T * scale + mask) V        per (batch, head); bf16 in/out, fp32 softmax
//
// Replaces F.scaled_dot_product_attention in MultiHeadAttention.forward (olmoasr/model.py:331-340):
//   encoder self-attention (no mask), decoder self-attention (causal + per-sample key length, derived
//   from the dense additive mask the reference passes, model.py:740-743) and cross-attention (no mask).
//
// Work item = 256 query rows (two 128-row tiles) x one (b, h), looping over 128-key tiles.  Persistent: one CTA per SM
// walks over the items with stride gridDim.x, so the next item's Q / K / V loads and first Q K^T MMAs run under the
// current item's output epilogue.  384 threads:
//   warps 0..3   softmax warpgroup of query tile 0        warps 4..7   softmax warpgroup of query tile 1
//   warp 8       TMA producer (Q per item, K/V in a 3-stage ring that keeps running across items; every K/V tile serves
//                both query tiles)
//   warp 9       tcgen05.mma issuer + TMEM owner           warps 10, 11 idle (keep the control warpgroup 4-aligned)
// setmaxnreg moves registers from the control warpgroup (56) to the softmax warpgroups (224) so that a thread can
// hold its whole 128-column score row: it reads the row from TMEM ONCE, immediately releases the S buffer
// (bar_sfree), and the MMA warp issues Q K^T of the NEXT key tile while this tile's softmax is still running.
//
// Thread r of a warpgroup owns score row r (= TMEM lane r): the row maximum and sum need no shuffles.
// TMEM (512 cols): S0 0..127 | S1 128..255 | O0 256..319 | O1 320..383.  O is accumulated by the P*V MMAs
// directly in TMEM; it is rescaled (tcgen05.ld -> scale -> tcgen05.st) only when a row maximum grows by more
// than 2^8 -- otherwise the stale maximum is kept (probabilities stay <= 256; exact after the final 1/l).
// Scale/subtract and the row sum use packed FFMA2 / FADD2, the maximum FMNMX3.  P goes to smem as bf16 in the
// 128B-swizzled K-major layout and is the A operand of the P*V MMA; V is consumed MN-major from its TMA tile.
// The output tile leaves through the (then free) P buffer as a swizzled bf16 tile + one TMA store.
//
// Measured and rejected (profiles/r02_attention_fwd_ptmem_ab.txt, r02_attention_ab_pingpong_poly.txt): P as a
// TMEM-resident A operand (parity green, 9 % slower: the kernel is paced by the exponentials, not by the shared-memory
// port), FlashAttention-3 style ping-pong of the two softmax warpgroups (5 % slower: one warp per scheduler cannot
// keep the MUFU pipe full), exponentials partly on the FMA pipe (OASR_ATTN_POLY, no gain).
// mbarrier parities: per key-tile barriers follow c_t (key tiles query tile t has consumed so far in this CTA), per item
// barriers follow the item count, so every role derives them the same way.
#include "common.cuh"
#include "ptx_sm100.cuh"

#include <stdlib.h>
#include <type_traits>

namespace oasr {
namespace {

constexpr int HD = 64;
constexpr int BQ = 128;                      // rows per query tile (two tiles per item)
constexpr int BKV = 128;
constexpr int KV_STAGES = 3;
constexpr int TILE_BYTES = 128 * HD * 2;     // 16 KB: a [128 rows][64 bf16] swizzled tile
constexpr int P_BYTES = BQ * BKV * 2;        // 32 KB: two 64-key halves of [128][128B]
constexpr int ATT_TILES = 2 * TILE_BYTES /*Q0,Q1*/ + KV_STAGES * 2 * TILE_BYTES /*K,V*/ + 2 * P_BYTES;  // 192 KB
constexpr int ATT_SMEM = ATT_TILES + 256;
constexpr int TMEM_COLS = 512;
constexpr int S_COL = 0, O_COL = 256;        // + t*128 / + t*64
constexpr float RESCALE_LOG2 = 8.0f;
constexpr int SOFTMAX_REGS = 216, CONTROL_REGS = 72;   // 8 * SOFTMAX + 4 * CONTROL == 12 * 168
// Share of the exponentials of an unmasked tile evaluated on the FMA pipe (exp2_poly2) instead of MUFU.EX2:
// 0 = none, 1 = every fourth pair (25 %), 2 = every second pair (50 %).  A/B builds: python -m olmoasr_b200.build --variant.
#ifndef OASR_ATTN_POLY
#define OASR_ATTN_POLY 0
#endif

#ifdef OASR_ATTN_TRACE
// Debug build only: one CTA in the middle of the grid stamps clock64() at phase boundaries (tools/trace_attention.py).
__device__ unsigned long long g_fwd_trace[4][512];
#define TRACE_DECL(role_)                                                                                     \
  const bool tr_on = blockIdx.x == gridDim.x / 2 && lane == 0 && (role_) >= 0;                                \
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
  float* lse;        // (B, H, Tq) log2-domain log-sum-exp, nullable
  const int32_t* kv_len;  // (B,) valid keys per sample, nullable
  int B, H, Tq, Tkv;
  int causal;
  float scale_log2;  // scale * log2(e)
};

struct Item {
  int h, b, q_base, kv_valid, n_kv[2], n_max;
};

__device__ __forceinline__ Item decode_item(const AttnParams& p, int id, int n_qblk) {
  Item w;
  const int rest = id / n_qblk;
  const int qblk = (id + rest) % n_qblk;   // rotate the query block per (head, sample): causal blocks differ in cost
  w.h = rest % p.H;
  w.b = rest / p.H;
  w.q_base = qblk * 2 * BQ;
  w.kv_valid = p.Tkv;
  if (p.kv_len) w.kv_valid = min(w.kv_valid, max(1, p.kv_len[w.b]));
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int q0 = w.q_base + t * BQ;
    int kv_end = w.kv_valid;
    if (p.causal) kv_end = min(kv_end, q0 + BQ);
    w.n_kv[t] = (q0 < p.Tq) ? (kv_end + BKV - 1) / BKV : 0;
  }
  w.n_max = max(w.n_kv[0], w.n_kv[1]);   // >= 1: query tile 0 of an item always exists and sees at least one key
  return w;
}

__global__ void __launch_bounds__(384, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const AttnParams p,
                     const int n_items) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + ATT_TILES);
  uint64_t& bar_q = bars[0];                  // Q of an item landed                          (parity: item)
  uint64_t* bar_s = bars + 1;                 // [2] S_t ready (MMA commit)                   (parity: c_t)
  uint64_t* bar_sfree = bars + 3;             // [2] S_t read into registers (4 warp arrivals)
  uint64_t* bar_p = bars + 5;                 // [2] P_t in smem, O_t rescaled (4 warp arrivals)
  uint64_t* bar_pv = bars + 7;                // [2] P_t V retired (MMA commit)
  uint64_t* bar_kv_full = bars + 9;           // [KV_STAGES]
  uint64_t* bar_kv_empty = bars + 9 + KV_STAGES;
  uint64_t& bar_qfree = bars[9 + 2 * KV_STAGES];      // every Q K^T of the item retired: Q smem reusable   (parity: item)
  uint64_t* bar_ofree = bars + 10 + 2 * KV_STAGES;    // [2] O_t read out of TMEM (4 warp arrivals)         (parity: items of tile t)
  uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(bars + 12 + 2 * KV_STAGES);

  const uint32_t sbase = ptx::smem_u32(smem_raw);
  if ((sbase & 1023u) != 0) {  // swizzled tiles need 1 KB alignment; the declaration above should guarantee it
    if (threadIdx.x == 0) printf("oasr attention: dynamic smem base %u not 1 KB aligned\n", sbase);
    __trap();
  }
  const uint32_t sQ = sbase;                          // tile t at sQ + t*16K
  const uint32_t sK0 = sQ + 2 * TILE_BYTES;           // stage s: K at sK0 + s*32K, V right after K
  const uint32_t sP = sK0 + KV_STAGES * 2 * TILE_BYTES;  // tile t at sP + t*32K
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qblk = (p.Tq + 2 * BQ - 1) / (2 * BQ);
  const int item0 = blockIdx.x, item_step = gridDim.x;

  if (warp == 8 && lane == 0) {
    ptx::tma_prefetch_desc(&tmQ); ptx::tma_prefetch_desc(&tmK); ptx::tma_prefetch_desc(&tmV); ptx::tma_prefetch_desc(&tmO);
    ptx::mbar_init(ptx::smem_u32(&bar_q), 1);
    ptx::mbar_init(ptx::smem_u32(&bar_qfree), 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bar_s[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bar_sfree[s]), 4);
      ptx::mbar_init(ptx::smem_u32(&bar_p[s]), 4);
      ptx::mbar_init(ptx::smem_u32(&bar_pv[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bar_ofree[s]), 4);
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
      int s = 0;
      uint32_t ph = 0, wi = 0;
      for (int id = item0; id < n_items; id += item_step, ++wi) {
        const Item w = decode_item(p, id, n_qblk);
        if (wi > 0) ptx::mbar_wait(ptx::smem_u32(&bar_qfree), (wi - 1) & 1);   // the previous item's Q K^T MMAs no longer read Q
        ptx::mbar_arrive_expect_tx(ptx::smem_u32(&bar_q), 2 * TILE_BYTES);
        ptx::tma_load_2d(sQ, &tmQ, ptx::smem_u32(&bar_q), w.h * HD, w.b * p.Tq + w.q_base);  // box 64 x 256 rows
        for (int j = 0; j < w.n_max; ++j) {
          ptx::mbar_wait(ptx::smem_u32(&bar_kv_empty[s]), ph ^ 1);
          const uint32_t full = ptx::smem_u32(&bar_kv_full[s]);
          ptx::mbar_arrive_expect_tx(full, 2 * TILE_BYTES);
          const int krow = w.b * p.Tkv + j * BKV;
          ptx::tma_load_2d(sK0 + s * 2 * TILE_BYTES, &tmK, full, w.h * HD, krow);
          ptx::tma_load_2d(sK0 + s * 2 * TILE_BYTES + TILE_BYTES, &tmV, full, w.h * HD, krow);
          if (++s == KV_STAGES) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp == 9) {
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
          for (int k = 0; k < BKV / 16; ++k)
            ptx::tc_mma_f16_lh(tmem + O_COL + t * HD, p_lo[t] + (k >> 2) * ((P_BYTES / 2) >> 4) + (k & 3) * 2, hi_k,
                               vlo + k * (2048 >> 4), hi_v, idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
          ptx::tc_commit(ptx::smem_u32(&bar_pv[t]));
        }
        __syncwarp();
      };
      int st = 0;            // ring stage of the current key tile
      uint32_t st_ph = 0;    // its phase
      uint32_t wi = 0, c[2] = {0, 0}, items_t[2] = {0, 0};
      for (int id = item0; id < n_items; id += item_step, ++wi) {
        const Item w = decode_item(p, id, n_qblk);
        ptx::mbar_wait(ptx::smem_u32(&bar_q), wi & 1);
        ptx::mbar_wait(ptx::smem_u32(&bar_kv_full[st]), st_ph);
        TR(2);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (w.n_kv[t] > 0) {
            if (c[t] > 0) ptx::mbar_wait(ptx::smem_u32(&bar_sfree[t]), (c[t] - 1) & 1);   // the previous item's last S_t is in registers
            ptx::tc_fence_after();
            issue_qk(t, st);
          }
        }
        if (w.n_max == 1 && ptx::elect_one()) ptx::tc_commit(ptx::smem_u32(&bar_qfree));
        __syncwarp();
        TR(3);
        for (int j = 0; j < w.n_max; ++j) {
          int st_n = st + 1;
          uint32_t ph_n = st_ph;
          if (st_n == KV_STAGES) { st_n = 0; ph_n ^= 1; }
          // (1) next tile's scores as soon as this tile's S has been read into registers
          if (j + 1 < w.n_max) {
            ptx::mbar_wait(ptx::smem_u32(&bar_kv_full[st_n]), ph_n);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              if (j + 1 < w.n_kv[t]) {
                ptx::mbar_wait(ptx::smem_u32(&bar_sfree[t]), (c[t] + j) & 1);
                TR(4 + 2 * t);
                ptx::tc_fence_after();
                issue_qk(t, st_n);
                TR(5 + 2 * t);
              }
            }
            if (j + 2 == w.n_max && ptx::elect_one()) ptx::tc_commit(ptx::smem_u32(&bar_qfree));   // last Q K^T of the item issued
            __syncwarp();
          }
          // (2) this tile's P V
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (j < w.n_kv[t]) {
              ptx::mbar_wait(ptx::smem_u32(&bar_p[t]), (c[t] + j) & 1);   // P_t(j) in smem, O_t rescaled
              if (j == 0 && items_t[t] > 0) ptx::mbar_wait(ptx::smem_u32(&bar_ofree[t]), (items_t[t] - 1) & 1);   // previous O_t left TMEM
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
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          c[t] += w.n_kv[t];
          items_t[t] += (w.n_kv[t] > 0) ? 1u : 0u;
        }
      }
      TRACE_END();
    }
  } else {
    // ----------------------------- softmax / output warpgroups -----------------------------
    ptx::setmaxnreg_inc<SOFTMAX_REGS>();
    const int t = warp >> 2;                         // query tile of this warpgroup
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;               // row within the tile == TMEM lane
    const uint32_t t_lane = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t t_s = tmem + t_lane + S_COL + t * BKV;
    const uint32_t t_o = tmem + t_lane + O_COL + t * HD;
    const uint32_t sPt = sP + t * P_BYTES;
    const float c = p.scale_log2;
    const bool leader = (quarter == 0 && lane == 0);
    TRACE_DECL(warp == 0 ? 0 : (warp == 4 ? 3 : -1))
    TR(1);
    uint32_t ct = 0;           // key tiles this warpgroup has consumed so far
    bool out_pending = false;  // the previous item's output tile may still be read out of sP_t by the TMA engine

    for (int id = item0; id < n_items; id += item_step) {
      const Item w = decode_item(p, id, n_qblk);
      const int qi = w.q_base + t * BQ + r;          // query index within the sequence
      const int n_t = t ? w.n_kv[1] : w.n_kv[0];   // (no runtime-indexed array: that would live in local memory)
      float m = -INFINITY, l = 0.f;

      for (int j = 0; j < n_t; ++j) {
        TR(10);
        ptx::mbar_wait(ptx::smem_u32(&bar_s[t]), (ct + j) & 1);
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
        int limit = w.kv_valid - k0;                     // keys [0, limit) of this tile are visible
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
          ptx::mbar_wait(ptx::smem_u32(&bar_pv[t]), (ct + j - 1) & 1);   // P_t V (j-1) retired: O_t valid, sP_t reusable
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
        } else if (out_pending) {   // first P store of an item: the TMA engine must have read the previous output tile out of sP_t
          if (leader) ptx::tma_store_wait_read<0>();
          asm volatile("bar.sync %0, 128;" ::"r"(1 + t) : "memory");
          out_pending = false;
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
#pragma unroll
          for (int cc = 0; cc < BKV / 32; ++cc) {
            const uint32_t half_base = sPt + (cc >> 1) * (P_BYTES / 2) + r * 128;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              uint32_t wv[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int i = cc * 32 + q4 * 8 + e * 2;
                const float2 x = __ffma2_rn(make_float2(v[i], v[i + 1]), c2, n2);
                float2 pe;
                if ((POLY == 2 && (e & 1)) || (POLY == 1 && e == 3)) pe = exp2_poly2(x);
                else pe = make_float2(fast_exp2(x.x), fast_exp2(x.y));
                sums[e] = __fadd2_rn(sums[e], pe);
                wv[e] = pack_bf16x2(pe.x, pe.y);
              }
              const int chunk = (cc & 1) * 4 + q4;       // 16-byte chunk index within the 128-byte row
              const uint32_t addr = half_base + ((chunk ^ (r & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(wv[0]), "r"(wv[1]), "r"(wv[2]), "r"(wv[3])
                           : "memory");
            }
          }
        };
        if (OASR_ATTN_POLY != 0 && __all_sync(0xffffffffu, limit >= BKV)) emit_p(std::integral_constant<int, OASR_ATTN_POLY>{});
        else emit_p(std::integral_constant<int, 0>{});
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bar_p[t]));
        TR(16);
        l += ((sums[0].x + sums[0].y) + (sums[1].x + sums[1].y)) + ((sums[2].x + sums[2].y) + (sums[3].x + sums[3].y));
      }
      if (n_t > 0) {
        // ---- output tile: O_t / l -> bf16 -> swizzled [128][128 B] tile in sP_t -> one TMA store (the 3-D map clips rows past Tq)
        ptx::mbar_wait(ptx::smem_u32(&bar_pv[t]), (ct + n_t - 1) & 1);
        ptx::tc_fence_after();
        const float inv = 1.f / l;
        uint32_t o[64];
        ptx::tc_ld_32x32b_x32(t_o, reinterpret_cast<uint32_t (&)[32]>(o[0]));
        ptx::tc_ld_32x32b_x32(t_o + 32, reinterpret_cast<uint32_t (&)[32]>(o[32]));
        ptx::tc_wait_ld();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bar_ofree[t]));   // the next item may start accumulating O_t
#pragma unroll
        for (int q8 = 0; q8 < 8; ++q8) {
          ptx::st_shared_v4(sPt + r * 128 + ((q8 ^ (r & 7)) << 4),
                            pack_bf16x2(__uint_as_float(o[8 * q8 + 0]) * inv, __uint_as_float(o[8 * q8 + 1]) * inv),
                            pack_bf16x2(__uint_as_float(o[8 * q8 + 2]) * inv, __uint_as_float(o[8 * q8 + 3]) * inv),
                            pack_bf16x2(__uint_as_float(o[8 * q8 + 4]) * inv, __uint_as_float(o[8 * q8 + 5]) * inv),
                            pack_bf16x2(__uint_as_float(o[8 * q8 + 6]) * inv, __uint_as_float(o[8 * q8 + 7]) * inv));
        }
        ptx::fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(1 + t) : "memory");
        if (leader) {
          ptx::tma_store_3d(&tmO, sPt, w.h * HD, w.q_base + t * BQ, w.b);
          ptx::tma_store_commit();
        }
        out_pending = true;
        if (qi < p.Tq && p.lse) p.lse[(static_cast<int64_t>(w.b) * p.H + w.h) * p.Tq + qi] = m * c + log2f(l);
        ct += n_t;
      }
      TR(30);
    }
    if (leader) ptx::tma_store_wait_read<0>();
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
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc;
  if ((rc = make_tmap_2d(&tmQ, q, 2, (uint64_t)(H * HD), (uint64_t)(B * Tq), (uint64_t)ldq * 2, HD, 2 * BQ, true))) return rc;
  if ((rc = make_tmap_2d(&tmK, k, 2, (uint64_t)(H * HD), (uint64_t)(B * Tkv), (uint64_t)ldk * 2, HD, BKV, true))) return rc;
  if ((rc = make_tmap_2d(&tmV, v, 2, (uint64_t)(H * HD), (uint64_t)(B * Tkv), (uint64_t)ldv * 2, HD, BKV, true))) return rc;
  // the output leaves per (sample, query tile): 3-D map (columns, queries of one sample, batch) so that the 128-row box of a
  // sample's last query tile is clipped at Tq instead of running into the next sample
  if ((rc = make_tmap_3d(&tmO, o, 2, (uint64_t)(H * HD), (uint64_t)Tq, (uint64_t)B, (uint64_t)ldo * 2, (uint64_t)Tq * ldo * 2, HD, BQ, 1,
                         true))) return rc;
  AttnParams p;
  p.lse = lse; p.kv_len = kv_len;
  p.B = (int)B; p.H = (int)H; p.Tq = (int)Tq; p.Tkv = (int)Tkv; p.causal = causal;
  p.scale_log2 = scale * 1.4426950408889634f;
  static bool attr_set = false;
  if (!attr_set) {
    OASR_CUDA_OK(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    attr_set = true;
  }
  // one CTA per SM walking over (query block, head, sample) items; OASR_FWD_PERSISTENT=0: one CTA per item (A/B)
  const int64_t n_work = ceil_div(Tq, 2 * BQ) * H * B;
  OASR_REQUIRE(n_work < (int64_t(1) << 31), "attention: too many work items");
  const char* pe = getenv("OASR_FWD_PERSISTENT");   // read on every call: the host layer flips it when collectives share the GPU
  const bool persistent = !(pe && pe[0] == '0');
  int64_t ctas = num_sms();
  if (const char* cap = getenv("OASR_ATTN_MAX_CTAS")) {   // tests: few CTAs, so that small problems walk the multi-item path
    const long v = atol(cap);
    if (v > 0 && v < ctas) ctas = v;
  }
  const unsigned grid = persistent ? (unsigned)(n_work < ctas ? n_work : ctas) : (unsigned)n_work;
  attention_fwd_kernel<<<grid, 384, ATT_SMEM, (cudaStream_t)stream>>>(tmQ, tmK, tmV, tmO, p, (int)n_work);
  OASR_LAUNCH_CHECK();
  return OASR_OK;
}
