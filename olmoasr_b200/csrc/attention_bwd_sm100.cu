// Flash-style attention backward for head_dim 64 on tcgen05 / TMEM (sm_100a).
//
// Autograd of F.scaled_dot_product_attention (olmoasr/model.py:331-340) for the three training uses
// (encoder self, decoder causal+length self, cross).  Given Q, K, V, O, dO and the forward's log2-domain
// LSE it produces dQ, dK, dV without ever materialising the (Tq x Tkv) score matrix in HBM.
//
// Work item = one 128-key tile of one (b, h), looping over the 128-query tiles that can see it (a persistent CTA per SM
// walks over the items, see the kernel's comment):
//     S  = Q K^T            dP = dO V^T                      (tcgen05, accumulators in TMEM)
//     P  = exp2(S c - LSE)  dS = P o (dP - D)                (thread r owns query row r: no reductions)
//     dV += P^T dO          dK += dS^T Q        dQ_i = dS K  (P / dS go through swizzled smem as bf16;
//                                                             the "transposes" are MN-major descriptors)
// dV and dK stay resident in TMEM for the whole loop and leave through swizzled smem tiles + one TMA store each; dQ_i is
// drained per query tile into a swizzled fp32 smem tile and reduced across key tiles with ONE bulk TMA reduce-add per
// 32-column half (cp.reduce.async.bulk.tensor .add) into an fp32 scratch buffer (converted to bf16 afterwards).  (Per-thread red.global.add.v4 made the whole kernel
// atomics-bound: ~6500 cycles per tile pair against ~1300 of tensor work.)
// 8 compute warps (two per TMEM lane quarter, each taking half of the 128 key columns; packed FFMA2 / FADD2 /
// FMUL2 math) + TMA warp + MMA warp + 4 dQ-drain warps; S / dP are released to the MMA warp as soon as they sit in
// registers.  The dQ drain (TMEM -> fp32 smem tile -> TMA reduce-add) has its own warpgroup: on the compute warps it
// cost 1300 of 3850 cycles per query tile (profiles/r02_attention_bwd_trace_before.txt).
// TMEM map (512 cols): S 0..127 | dP 128..255 | dV 256..319 | dK 320..383 | dQ 384..447.
#include "common.cuh"
#include "ptx_sm100.cuh"

#include <stdlib.h>
#include <type_traits>

namespace oasr {
namespace {

constexpr int HD = 64;
constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int Q_STAGES = 3;
constexpr int TILE_BYTES = 128 * HD * 2;  // 16 KB
constexpr int P_BYTES = BQ * BKV * 2;     // 32 KB
// K, V, (Q, dO) x Q_STAGES, P, dS
constexpr int DQ_BYTES = BQ * HD * 4;      // 32 KB fp32 staging of dQ_i: two 32-column halves of [128][128B], swizzled
constexpr int BWD_TILES = 2 * TILE_BYTES + Q_STAGES * 2 * TILE_BYTES + 2 * P_BYTES + DQ_BYTES;  // 224 KB
constexpr int BWD_SMEM = BWD_TILES + 256;
constexpr int TMEM_COLS = 512;
constexpr int S_COL = 0, DP_COL = 128, DV_COL = 256, DK_COL = 320, DQ_COL = 384;
// Share of the exponentials of an unmasked tile evaluated on the FMA pipe (exp2_poly2) instead of MUFU.EX2:
// 0 = none, 1 = every fourth pair, 2 = every second pair.
// Measured (profiles/r02_attention_ab_pingpong_poly.txt): 1 = 1.154 ms, 0 = 1.178 ms, 2 = 1.242 ms on the encoder shape.
#ifndef OASR_BWD_POLY
#define OASR_BWD_POLY 1
#endif
constexpr int BWD_THREADS = 512;
// register pool: 8 * 184 (compute) + 4 * 96 (TMA / MMA warpgroup) + 4 * 48 (drain) == 16 * 128
constexpr int COMPUTE_REGS = 184, CONTROL_REGS = 96, DRAIN_REGS = 48;

struct BwdParams {
  const float* lse;   // (B,H,Tq) log2 domain
  const float* delta; // (B,H,Tq) rowsum(dO o O)
  float* dq_accum;    // (B*Tq, H*64) fp32, zero-initialised
  const int32_t* kv_len;
  int B, H, Tq, Tkv;
  int causal;
  float scale, scale_log2;
};

#ifdef OASR_ATTN_TRACE
// Debug build only (python -m olmoasr_b200.build --variant attn_trace -DOASR_ATTN_TRACE): one CTA in the middle of the
// grid stamps clock64() at every phase boundary; tools/trace_attention.py prints the per-phase cycle table.
__device__ unsigned long long g_bwd_trace[4][512];
#define TRACE_DECL(role_)                                                                                     \
  const bool tr_on = blockIdx.x == gridDim.x / 2 && lane == 0;                                                                               \
  const int tr_role = (role_);                                                                                \
  int tr_n = 0;
#define TR(id_)                                                                                               \
  do {                                                                                                        \
    if (tr_on && tr_n < 511) g_bwd_trace[tr_role][++tr_n] = (static_cast<unsigned long long>(id_) << 48) | (clock64() & 0xffffffffffffull); \
  } while (0)
#define TRACE_END()                                                                                           \
  do {                                                                                                        \
    if (tr_on) g_bwd_trace[tr_role][0] = tr_n;                                                                \
  } while (0)
#else
#define TRACE_DECL(role_)
#define TR(id_)
#define TRACE_END()
#endif

// 512 threads: warps 0..7 compute (warp w: TMEM lane quarter w%4, key-column half w/4), warp 8 TMA producer,
// warp 9 MMA issuer + TMEM owner, warps 10-11 idle, warps 12..15 dQ drain (lane quarter w%4).  setmaxnreg gives the
// compute warps 184 registers so that a thread holds its 64 S and 64 dP values at once: both TMEM buffers are released
// right after the load (bar_free) and the MMA warp issues S / dP of the NEXT query tile underneath this tile's exp / dS
// math.
//
// Persistent: a CTA walks over work items (key tile, head, sample) with stride gridDim.x.  Setting a CTA up (barriers,
// TMEM allocation, first K/V/Q loads: ~2900 cycles) and tearing it down (~4200 cycles after the last P / dS store) cost
// 16 % of the encoder shape and 35 % of the cross-attention shape when every item was its own CTA
// (profiles/r02_attention_trace_after_epilogue.txt); now the next item's K / V / Q loads and first S / dP MMAs run
// under the current item's dK / dV epilogue.  mbarrier parities: per query-tile barriers follow g (iterations done by
// this CTA), per item barriers follow w (items with work) or e (epilogues), so every role derives them the same way.
struct Item {
  int kv_tile, h, b, kv0, kv_valid, i_begin, n_iter;
};

__device__ __forceinline__ Item decode_item(const BwdParams& p, int id, int n_kv_tiles, int n_q_tiles) {
  Item w;
  const int rest = id / n_kv_tiles;
  w.kv_tile = (id + rest) % n_kv_tiles;   // rotate the key tile per (head, sample): causal tiles cost 1 .. n_q_tiles iterations
  w.h = rest % p.H;
  w.b = rest / p.H;
  w.kv0 = w.kv_tile * BKV;
  w.kv_valid = p.Tkv;
  if (p.kv_len) w.kv_valid = min(w.kv_valid, max(1, p.kv_len[w.b]));
  w.i_begin = p.causal ? w.kv_tile : 0;
  w.n_iter = (w.kv0 < w.kv_valid) ? n_q_tiles - w.i_begin : 0;   // fully masked key tile: no work, zero gradients
  return w;
}

__global__ void __launch_bounds__(BWD_THREADS, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                     const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmDK,
                     const __grid_constant__ CUtensorMap tmDV, const BwdParams p, const int n_items) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + BWD_TILES);
  uint64_t& bar_kv = bars[0];      // K, V of an item landed                       (parity w)
  uint64_t& bar_sdp = bars[1];     // S, dP ready (MMA commit)                      (parity g)
  uint64_t& bar_free = bars[2];    // S, dP read into registers (8 warp arrivals)   (parity g)
  uint64_t& bar_pds = bars[3];     // P, dS in smem (8 warp arrivals)               (parity g)
  uint64_t& bar_dq = bars[4];      // dV / dK / dQ MMAs of this tile retired        (parity g)
  uint64_t& bar_done = bars[5];    // every MMA of the item retired: dK / dV final, K / V smem free   (parity w)
  uint64_t* bar_q_full = bars + 6;
  uint64_t* bar_q_empty = bars + 6 + Q_STAGES;
  uint64_t& bar_dqfree = bars[6 + 2 * Q_STAGES];    // dQ read out of TMEM (4 drain-warp arrivals)           (parity g)
  uint64_t& bar_accfree = bars[7 + 2 * Q_STAGES];   // dK / dV read out of TMEM (8 warp arrivals)            (parity w)
  uint64_t& bar_epi = bars[8 + 2 * Q_STAGES];       // dK / dV staging tiles read by the TMA engine          (parity e)
  uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(bars + 9 + 2 * Q_STAGES);

  const uint32_t sbase = ptx::smem_u32(smem_raw);
  if ((sbase & 1023u) != 0) {
    if (threadIdx.x == 0) printf("oasr attention_bwd: dynamic smem base %u not 1 KB aligned\n", sbase);
    __trap();
  }
  const uint32_t sK = sbase, sV = sK + TILE_BYTES;
  const uint32_t sQ0 = sV + TILE_BYTES;            // stage s: Q at sQ0 + s*32K, dO right after
  const uint32_t sP = sQ0 + Q_STAGES * 2 * TILE_BYTES, sdS = sP + P_BYTES, sDQ = sdS + P_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv_tiles = (p.Tkv + BKV - 1) / BKV;
  const int n_q_tiles = (p.Tq + BQ - 1) / BQ;
  const int item0 = blockIdx.x, item_step = gridDim.x;

  if (warp == 8 && lane == 0) {
    ptx::tma_prefetch_desc(&tmQ); ptx::tma_prefetch_desc(&tmK); ptx::tma_prefetch_desc(&tmV); ptx::tma_prefetch_desc(&tmdO);
    ptx::tma_prefetch_desc(&tmDQ); ptx::tma_prefetch_desc(&tmDK); ptx::tma_prefetch_desc(&tmDV);
    ptx::mbar_init(ptx::smem_u32(&bar_kv), 1);
    ptx::mbar_init(ptx::smem_u32(&bar_sdp), 1);
    ptx::mbar_init(ptx::smem_u32(&bar_free), 8);
    ptx::mbar_init(ptx::smem_u32(&bar_pds), 8);
    ptx::mbar_init(ptx::smem_u32(&bar_dq), 1);
    ptx::mbar_init(ptx::smem_u32(&bar_done), 1);
    ptx::mbar_init(ptx::smem_u32(&bar_dqfree), 4);
    ptx::mbar_init(ptx::smem_u32(&bar_accfree), 8);
    ptx::mbar_init(ptx::smem_u32(&bar_epi), 1);
    for (int s = 0; s < Q_STAGES; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bar_q_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bar_q_empty[s]), 1);
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

  if (warp >= 12) {
    ptx::setmaxnreg_dec<DRAIN_REGS>();
    // ------------------------------ dQ drain warpgroup ------------------------------
    // dQ of a query tile: TMEM (this warp's 32 lanes x 64 columns) -> swizzled fp32 smem rows (two 32-column halves of
    // [128][128 B]) -> one TMA reduce-add per half into the fp32 scratch.  Rows past Tq carry exact zeros (their P and dS
    // rows are zero), rows past the tensor are clipped by TMA.
    TRACE_DECL(3)
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t t_lane = static_cast<uint32_t>(quarter * 32) << 16;
    const bool issuer = (warp == 12 && lane == 0);
    uint32_t g = 0;
    for (int id = item0; id < n_items; id += item_step) {
      const Item w = decode_item(p, id, n_kv_tiles, n_q_tiles);
      for (int it = 0; it < w.n_iter; ++it, ++g) {
        ptx::mbar_wait(ptx::smem_u32(&bar_dq), g & 1);
        TR(20);
        ptx::tc_fence_after();
        if (issuer) ptx::tma_store_wait_read<0>();   // the previous reduce has finished reading the staging tile
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32];
          ptx::tc_ld_32x32b_x32(tmem + t_lane + DQ_COL + half * 32, v);
          ptx::tc_wait_ld();
          if (half == 1) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bar_dqfree));   // dQ TMEM may be overwritten
          } else {
            asm volatile("bar.sync 1, 128;" ::: "memory");
          }
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const uint32_t addr = sDQ + half * (DQ_BYTES / 2) + r * 128 + ((q4 ^ (r & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(__uint_as_float(v[4 * q4]) * p.scale),
                         "f"(__uint_as_float(v[4 * q4 + 1]) * p.scale), "f"(__uint_as_float(v[4 * q4 + 2]) * p.scale),
                         "f"(__uint_as_float(v[4 * q4 + 3]) * p.scale)
                         : "memory");
          }
        }
        TR(21);
        ptx::fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {
          const int qrow = w.b * p.Tq + (w.i_begin + it) * BQ;
          ptx::tma_reduce_add_2d(&tmDQ, sDQ, w.h * HD, qrow);
          ptx::tma_reduce_add_2d(&tmDQ, sDQ + DQ_BYTES / 2, w.h * HD + 32, qrow);
          ptx::tma_store_commit();
        }
        TR(22);
      }
    }
    if (issuer) ptx::tma_store_wait_read<0>();
    TRACE_END();
  } else if (warp >= 8) {
    ptx::setmaxnreg_dec<CONTROL_REGS>();
    if (warp == 8 && lane == 0) {
      // ------------------------------ TMA producer ------------------------------
      TRACE_DECL(2)
      TR(1);
      int s = 0;
      uint32_t ph = 0, wi = 0;
      for (int id = item0; id < n_items; id += item_step) {
        const Item w = decode_item(p, id, n_kv_tiles, n_q_tiles);
        if (w.n_iter == 0) continue;
        if (wi > 0) ptx::mbar_wait(ptx::smem_u32(&bar_done), (wi - 1) & 1);   // the previous item's MMAs no longer read K / V
        ptx::mbar_arrive_expect_tx(ptx::smem_u32(&bar_kv), 2 * TILE_BYTES);
        ptx::tma_load_2d(sK, &tmK, ptx::smem_u32(&bar_kv), w.h * HD, w.b * p.Tkv + w.kv0);
        ptx::tma_load_2d(sV, &tmV, ptx::smem_u32(&bar_kv), w.h * HD, w.b * p.Tkv + w.kv0);
        for (int it = 0; it < w.n_iter; ++it) {
          ptx::mbar_wait(ptx::smem_u32(&bar_q_empty[s]), ph ^ 1);
          TR(2);
          const uint32_t full = ptx::smem_u32(&bar_q_full[s]);
          ptx::mbar_arrive_expect_tx(full, 2 * TILE_BYTES);
          const int qrow = w.b * p.Tq + (w.i_begin + it) * BQ;
          ptx::tma_load_2d(sQ0 + s * 2 * TILE_BYTES, &tmQ, full, w.h * HD, qrow);
          ptx::tma_load_2d(sQ0 + s * 2 * TILE_BYTES + TILE_BYTES, &tmdO, full, w.h * HD, qrow);
          if (++s == Q_STAGES) { s = 0; ph ^= 1; }
          if (it + 1 == w.n_iter) {
            // K / V of the next item can only be loaded once this item's MMAs have retired; measured 3000+ cycles from
            // that load to its arrival while the epilogue traffic is in flight (profiles/r02_attention_trace_persistent.txt).
            // Pull the tiles into L2 now, one query tile ahead.
            for (int nid = id + item_step; nid < n_items; nid += item_step) {
              const Item nx = decode_item(p, nid, n_kv_tiles, n_q_tiles);
              if (nx.n_iter == 0) continue;
              ptx::tma_prefetch_l2_2d(&tmK, nx.h * HD, nx.b * p.Tkv + nx.kv0);
              ptx::tma_prefetch_l2_2d(&tmV, nx.h * HD, nx.b * p.Tkv + nx.kv0);
              const int qrow = nx.b * p.Tq + nx.i_begin * BQ;
              ptx::tma_prefetch_l2_2d(&tmQ, nx.h * HD, qrow);
              ptx::tma_prefetch_l2_2d(&tmdO, nx.h * HD, qrow);
              break;
            }
          }
        }
        ++wi;
      }
      TRACE_END();
    } else if (warp == 9) {
      // ------------------------------ MMA issuer (whole warp convergent; elect.sync picks the issuing lane) ------------------------------
      TRACE_DECL(1)
      TR(1);
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(128, 128, 0, 0);    // S, dP
      constexpr uint32_t idesc_kv = ptx::umma_idesc_bf16(128, 64, 1, 1);    // dV, dK
      constexpr uint32_t idesc_dq = ptx::umma_idesc_bf16(128, 64, 0, 1);    // dQ
      // descriptor bases: K-major operands (LBO unused = 16, SBO 1024); MN-major A from sP / sdS (LBO = half tile),
      // MN-major B from the [rows][64] tiles (single 64-wide chunk: LBO unused)
      uint32_t hi_k, hi_mnA, hi_mnB, u;
      uint32_t kK_lo, vK_lo, q_lo0, do_lo0, pT_lo, dsT_lo, dsK_lo, doB_lo0, qB_lo0, kB_lo;
      ptx::umma_desc_sw128_lh(sK, 16, 1024, kK_lo, hi_k);
      ptx::umma_desc_sw128_lh(sV, 16, 1024, vK_lo, u);
      ptx::umma_desc_sw128_lh(sQ0, 16, 1024, q_lo0, u);
      ptx::umma_desc_sw128_lh(sQ0 + TILE_BYTES, 16, 1024, do_lo0, u);
      ptx::umma_desc_sw128_lh(sP, P_BYTES / 2, 1024, pT_lo, hi_mnA);
      ptx::umma_desc_sw128_lh(sdS, P_BYTES / 2, 1024, dsT_lo, u);
      ptx::umma_desc_sw128_lh(sdS, 16, 1024, dsK_lo, u);
      ptx::umma_desc_sw128_lh(sQ0 + TILE_BYTES, BQ * 128, 1024, doB_lo0, hi_mnB);
      ptx::umma_desc_sw128_lh(sQ0, BQ * 128, 1024, qB_lo0, u);
      ptx::umma_desc_sw128_lh(sK, BKV * 128, 1024, kB_lo, u);
      constexpr uint32_t STAGE_LO = (2 * TILE_BYTES) >> 4;
      auto issue_s_dp = [&](int stage) {
        const uint32_t ql = q_lo0 + stage * STAGE_LO, dol = do_lo0 + stage * STAGE_LO;
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            ptx::tc_mma_f16_lh(tmem + S_COL, ql + k * 2, hi_k, kK_lo + k * 2, hi_k, idesc_s, k > 0);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            ptx::tc_mma_f16_lh(tmem + DP_COL, dol + k * 2, hi_k, vK_lo + k * 2, hi_k, idesc_s, k > 0);
          ptx::tc_commit(ptx::smem_u32(&bar_sdp));
        }
        __syncwarp();
      };
      int st = 0;
      uint32_t st_ph = 0, g = 0, wi = 0;
      for (int id = item0; id < n_items; id += item_step) {
        const Item w = decode_item(p, id, n_kv_tiles, n_q_tiles);
        if (w.n_iter == 0) continue;
        ptx::mbar_wait(ptx::smem_u32(&bar_kv), wi & 1);
        TR(2);
        ptx::mbar_wait(ptx::smem_u32(&bar_q_full[st]), st_ph);
        if (g > 0) ptx::mbar_wait(ptx::smem_u32(&bar_free), (g - 1) & 1);   // S / dP of the previous item's last tile are in registers
        TR(3);
        ptx::tc_fence_after();
        issue_s_dp(st);
        TR(4);
        for (int it = 0; it < w.n_iter; ++it, ++g) {
          int st_n = st + 1;
          uint32_t ph_n = st_ph;
          if (st_n == Q_STAGES) { st_n = 0; ph_n ^= 1; }
          if (it + 1 < w.n_iter) {   // next tile's S / dP as soon as this tile's have been read into registers
            ptx::mbar_wait(ptx::smem_u32(&bar_q_full[st_n]), ph_n);
            TR(5);
            ptx::mbar_wait(ptx::smem_u32(&bar_free), g & 1);
            TR(6);
            ptx::tc_fence_after();
            issue_s_dp(st_n);
            TR(7);
          }
          const uint32_t dob = doB_lo0 + st * STAGE_LO, qb = qB_lo0 + st * STAGE_LO;
          ptx::mbar_wait(ptx::smem_u32(&bar_pds), g & 1);  // P, dS in smem
          if (it == 0 && wi > 0) ptx::mbar_wait(ptx::smem_u32(&bar_accfree), (wi - 1) & 1);   // previous dK / dV left TMEM
          TR(8);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < BQ / 16; ++k) {  // reduction over the 128 query rows, 16 at a time
              ptx::tc_mma_f16_lh(tmem + DV_COL, pT_lo + k * (2048 >> 4), hi_mnA, dob + k * (2048 >> 4), hi_mnB, idesc_kv,
                                 (it > 0 || k > 0) ? 1u : 0u);
              ptx::tc_mma_f16_lh(tmem + DK_COL, dsT_lo + k * (2048 >> 4), hi_mnA, qb + k * (2048 >> 4), hi_mnB, idesc_kv,
                                 (it > 0 || k > 0) ? 1u : 0u);
            }
          }
          __syncwarp();
          if (g > 0) {   // the drain warps have read dQ of the previous tile out of TMEM
            ptx::mbar_wait(ptx::smem_u32(&bar_dqfree), (g - 1) & 1);
            ptx::tc_fence_after();
          }
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < BKV / 16; ++k)   // dQ_i = dS K : reduction over the 128 keys
              ptx::tc_mma_f16_lh(tmem + DQ_COL, dsK_lo + (k >> 2) * ((P_BYTES / 2) >> 4) + (k & 3) * 2, hi_k,
                                 kB_lo + k * (2048 >> 4), hi_mnB, idesc_dq, k > 0);
            ptx::tc_commit(ptx::smem_u32(&bar_dq));
            ptx::tc_commit(ptx::smem_u32(&bar_q_empty[st]));
            if (it + 1 == w.n_iter) ptx::tc_commit(ptx::smem_u32(&bar_done));
          }
          __syncwarp();
          TR(9);
          st = st_n; st_ph = ph_n;
        }
        ++wi;
      }
      TRACE_END();
    }
  } else {
    // ----------------------------- compute warps -----------------------------
    ptx::setmaxnreg_inc<COMPUTE_REGS>();
    const int quarter = warp & 3;
    const int chalf = warp >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t t_lane = static_cast<uint32_t>(quarter * 32) << 16;
    const float c = p.scale_log2;
    TRACE_DECL(0)
#ifdef OASR_ATTN_TRACE
    const bool tr_keep = tr_on && warp == 0;
#define TRC(id_) do { if (tr_keep) TR(id_); } while (0)
#else
#define TRC(id_)
#endif
    TRC(1);

    // lse / delta of this thread's query row are fetched one iteration ahead -- across item boundaries too (their ~1 us
    // global-load latency used to be the top stall of the compute warps)
    auto load_stats = [&](const Item& w, int it, float& lse_o, float& dlt_o) {
      const int qn = (w.i_begin + it) * BQ + r;
      const bool ok = it < w.n_iter && qn < p.Tq;
      const int64_t at = (static_cast<int64_t>(w.b) * p.H + w.h) * p.Tq + qn;
      lse_o = ok ? __ldg(p.lse + at) : 0.f;
      dlt_o = ok ? __ldg(p.delta + at) : 0.f;
    };
    uint32_t g = 0, wi = 0, e = 0;
    Item w = decode_item(p, item0 < n_items ? item0 : 0, n_kv_tiles, n_q_tiles);
    float lse_nx, dlt_nx;
    load_stats(w, 0, lse_nx, dlt_nx);
    for (int id = item0; id < n_items; id += item_step) {
      for (int it = 0; it < w.n_iter; ++it, ++g) {
        const int q0 = (w.i_begin + it) * BQ;
        const int qi = q0 + r;
        const bool q_ok = qi < p.Tq;
        const float lse = lse_nx, dlt = dlt_nx;
        load_stats(w, it + 1, lse_nx, dlt_nx);
        int limit = w.kv_valid - w.kv0;                     // visible keys of this tile: [0, limit)
        if (p.causal) limit = min(limit, qi - w.kv0 + 1);
        if (!q_ok) limit = 0;
        TRC(10);
        ptx::mbar_wait(ptx::smem_u32(&bar_sdp), g & 1);
        TRC(11);
        ptx::tc_fence_after();
        uint32_t sv[64], dv[64];
        ptx::tc_ld_32x32b_x32(tmem + t_lane + S_COL + chalf * 64, reinterpret_cast<uint32_t (&)[32]>(sv[0]));
        ptx::tc_ld_32x32b_x32(tmem + t_lane + S_COL + chalf * 64 + 32, reinterpret_cast<uint32_t (&)[32]>(sv[32]));
        ptx::tc_ld_32x32b_x32(tmem + t_lane + DP_COL + chalf * 64, reinterpret_cast<uint32_t (&)[32]>(dv[0]));
        ptx::tc_ld_32x32b_x32(tmem + t_lane + DP_COL + chalf * 64 + 32, reinterpret_cast<uint32_t (&)[32]>(dv[32]));
        ptx::tc_wait_ld();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bar_free));   // S / dP TMEM may be overwritten
        TRC(12);
        // All the exp / dS math happens BEFORE waiting for the previous tile's dV/dK/dQ MMAs (which still read sP / sdS):
        // the results wait in registers as packed bf16, so the compute warps never idle behind the tensor pipe.
        const float2 c2 = make_float2(c, c), nl2 = make_float2(-lse, -lse), nd2 = make_float2(-dlt, -dlt);
        uint32_t pp[32], dd[32];   // 64 columns each, packed bf16x2
        // MASKED is decided per warp (warp-uniform branch): interior tiles run without any per-element predicate code
        auto p_ds = [&](auto masked_tag) {
          constexpr bool MASKED = decltype(masked_tag)::value;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int cc = chalf * 2 + hh;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float2 x = __ffma2_rn(make_float2(__uint_as_float(sv[hh * 32 + i]), __uint_as_float(sv[hh * 32 + i + 1])), c2, nl2);
              float2 pe;
              if (!MASKED && ((OASR_BWD_POLY == 2 && ((i >> 1) & 1)) || (OASR_BWD_POLY == 1 && ((i >> 1) & 3) == 3))) pe = exp2_poly2(x);
              else pe = make_float2(fast_exp2(x.x), fast_exp2(x.y));
              if (MASKED) {
                if (cc * 32 + i >= limit) pe.x = 0.f;
                if (cc * 32 + i + 1 >= limit) pe.y = 0.f;
              }
              const float2 dq2 = __fadd2_rn(make_float2(__uint_as_float(dv[hh * 32 + i]), __uint_as_float(dv[hh * 32 + i + 1])), nd2);
              const float2 dsv = __fmul2_rn(pe, dq2);
              pp[hh * 16 + (i >> 1)] = pack_bf16x2(pe.x, pe.y);
              dd[hh * 16 + (i >> 1)] = pack_bf16x2(dsv.x, dsv.y);
            }
          }
        };
        if (__all_sync(0xffffffffu, limit >= (chalf + 1) * 64)) p_ds(std::false_type{}); else p_ds(std::true_type{});
        TRC(13);
        // sP / sdS are free once the previous tile's dV/dK/dQ MMAs have retired and, on an item's first tile, once the TMA
        // engine has read the previous item's dK / dV staging tiles out of them
        if (g > 0) ptx::mbar_wait(ptx::smem_u32(&bar_dq), (g - 1) & 1);
        if (it == 0 && e > 0) ptx::mbar_wait(ptx::smem_u32(&bar_epi), (e - 1) & 1);
        TRC(14);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int cc = chalf * 2 + hh;
          const uint32_t off = (cc >> 1) * (P_BYTES / 2) + r * 128;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const uint32_t a = off + ((((cc & 1) * 4 + q4) ^ (r & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP + a), "r"(pp[hh * 16 + q4 * 4]), "r"(pp[hh * 16 + q4 * 4 + 1]),
                         "r"(pp[hh * 16 + q4 * 4 + 2]), "r"(pp[hh * 16 + q4 * 4 + 3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sdS + a), "r"(dd[hh * 16 + q4 * 4]), "r"(dd[hh * 16 + q4 * 4 + 1]),
                         "r"(dd[hh * 16 + q4 * 4 + 2]), "r"(dd[hh * 16 + q4 * 4 + 3]) : "memory");
          }
        }
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bar_pds));
        TRC(15);
      }
      TRC(30);
      const Item w_next = decode_item(p, id + item_step < n_items ? id + item_step : id, n_kv_tiles, n_q_tiles);
      if (id + item_step < n_items) load_stats(w_next, 0, lse_nx, dlt_nx);

      // ---- dK / dV of the item: key row r of the tile, this warp's 32 of the 64 columns -> swizzled bf16 smem tiles (the
      // P / dS buffers are free now) -> one TMA store each.  (Per-thread 16-byte global stores of 64-byte row pieces took
      // ~3300 cycles and stalled the last dQ drain behind them in the LSU.)  The 3-D tensor maps clip rows past Tkv.
      uint32_t a[32], bq[32];
      if (w.n_iter > 0) {
        ptx::mbar_wait(ptx::smem_u32(&bar_done), wi & 1);
        ptx::tc_fence_after();
        TRC(31);
        ptx::tc_ld_32x32b_x32(tmem + t_lane + DV_COL + chalf * 32, a);
        ptx::tc_ld_32x32b_x32(tmem + t_lane + DK_COL + chalf * 32, bq);
        ptx::tc_wait_ld();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bar_accfree));   // the next item may start accumulating dK / dV
        ++wi;
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) { a[i] = 0u; bq[i] = 0u; }
        // a fully masked item still has to wait for the staging tiles (no P / dS store did it above)
        if (e > 0) ptx::mbar_wait(ptx::smem_u32(&bar_epi), (e - 1) & 1);
      }
#pragma unroll
      for (int q8 = 0; q8 < 4; ++q8) {
        const uint32_t off = r * 128 + (((chalf * 4 + q8) ^ (r & 7)) << 4);
        ptx::st_shared_v4(sP + off, pack_bf16x2(__uint_as_float(a[8 * q8 + 0]), __uint_as_float(a[8 * q8 + 1])),
                          pack_bf16x2(__uint_as_float(a[8 * q8 + 2]), __uint_as_float(a[8 * q8 + 3])),
                          pack_bf16x2(__uint_as_float(a[8 * q8 + 4]), __uint_as_float(a[8 * q8 + 5])),
                          pack_bf16x2(__uint_as_float(a[8 * q8 + 6]), __uint_as_float(a[8 * q8 + 7])));
        ptx::st_shared_v4(sdS + off, pack_bf16x2(__uint_as_float(bq[8 * q8 + 0]) * p.scale, __uint_as_float(bq[8 * q8 + 1]) * p.scale),
                          pack_bf16x2(__uint_as_float(bq[8 * q8 + 2]) * p.scale, __uint_as_float(bq[8 * q8 + 3]) * p.scale),
                          pack_bf16x2(__uint_as_float(bq[8 * q8 + 4]) * p.scale, __uint_as_float(bq[8 * q8 + 5]) * p.scale),
                          pack_bf16x2(__uint_as_float(bq[8 * q8 + 6]) * p.scale, __uint_as_float(bq[8 * q8 + 7]) * p.scale));
      }
      ptx::fence_proxy_async_smem();
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (threadIdx.x == 0) {
        ptx::tma_store_3d(&tmDV, sP, w.h * HD, w.kv0, w.b);
        ptx::tma_store_3d(&tmDK, sdS, w.h * HD, w.kv0, w.b);
        ptx::tma_store_commit();
        ptx::tma_store_wait_read<0>();   // the tiles stay valid until the engine has read them
        ptx::mbar_arrive(ptx::smem_u32(&bar_epi));
      }
      ++e;
      TRC(32);
      w = w_next;
    }
#ifdef OASR_ATTN_TRACE
    if (tr_keep) TRACE_END();
#endif
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem);
  }
}

// delta[b,h,q] = sum_d dO[b,q,h,d] * O[b,q,h,d]   (8 lanes per (row, head), 16-byte loads, 3 shuffle steps)
__global__ void attention_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout, float* __restrict__ delta,
                                       int64_t ldo, int64_t lddo, int B, int H, int Tq) {
  const int64_t item = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 3;  // (row, head) pair
  const int sub = threadIdx.x & 7;
  const int64_t total = static_cast<int64_t>(B) * Tq * H;
  float s = 0.f;
  int h = 0;
  int64_t row = 0;
  if (item < total) {
    h = item % H;
    row = item / H;  // b*Tq + q
    const uint4 a = *reinterpret_cast<const uint4*>(o + row * ldo + h * HD + sub * 8);
    const uint4 g = *reinterpret_cast<const uint4*>(dout + row * lddo + h * HD + sub * 8);
    const uint32_t av[4] = {a.x, a.y, a.z, a.w}, gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 p = unpack_bf16x2(av[j]), q = unpack_bf16x2(gv[j]);
      s += p.x * q.x + p.y * q.y;
    }
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (sub == 0 && item < total) {
    const int b = row / Tq, q = row % Tq;
    delta[(static_cast<int64_t>(b) * H + h) * Tq + q] = s;
  }
}

// dq (bf16, row stride lddq) = bf16(dq_accum (fp32, contiguous rows of `width`))
__global__ void f32_rows_to_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int64_t rows, int width, int64_t ldd) {
  const int vec_per_row = width >> 3;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows * vec_per_row;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / vec_per_row;
    const int v = i % vec_per_row;
    const float4 a = reinterpret_cast<const float4*>(src + row * width)[2 * v];
    const float4 b2 = reinterpret_cast<const float4*>(src + row * width)[2 * v + 1];
    uint4 u;
    u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w);
    u.z = pack_bf16x2(b2.x, b2.y); u.w = pack_bf16x2(b2.z, b2.w);
    reinterpret_cast<uint4*>(dst + row * ldd)[v] = u;
  }
}

}  // namespace
}  // namespace oasr

using namespace oasr;

#ifdef OASR_ATTN_TRACE
extern "C" OASR_API int oasr_debug_bwd_trace(unsigned long long* host_out) {   // 4 x 512 words
  return cudaMemcpyFromSymbol(host_out, g_bwd_trace, sizeof(unsigned long long) * 4 * 512) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int oasr_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                  const void* o, int64_t ldo, const void* dout, int64_t lddo, const float* lse,
                                  float* delta, float* dq_accum, void* dq, int64_t lddq, void* dk, int64_t lddk,
                                  void* dv, int64_t lddv, int64_t B, int64_t H, int64_t Tq, int64_t Tkv,
                                  int64_t head_dim, int causal, const int32_t* kv_len, float scale, void* stream) {
  OASR_REQUIRE(head_dim == HD, "attention_bwd: head_dim %ld unsupported", (long)head_dim);
  OASR_REQUIRE(B > 0 && H > 0 && Tq > 0 && Tkv > 0, "attention_bwd: empty problem");
  OASR_REQUIRE(((ldq | ldk | ldv | ldo | lddo | lddq | lddk | lddv) & 7) == 0, "attention_bwd: strides must be multiples of 8");
  OASR_REQUIRE(!causal || Tq == Tkv, "attention_bwd: causal needs Tq == Tkv");
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tmQ, tmK, tmV, tmdO, tmDQ, tmDK, tmDV;
  int rc;
  if ((rc = make_tmap_2d(&tmQ, q, 2, (uint64_t)(H * HD), (uint64_t)(B * Tq), (uint64_t)ldq * 2, HD, BQ, true))) return rc;
  if ((rc = make_tmap_2d(&tmK, k, 2, (uint64_t)(H * HD), (uint64_t)(B * Tkv), (uint64_t)ldk * 2, HD, BKV, true))) return rc;
  if ((rc = make_tmap_2d(&tmV, v, 2, (uint64_t)(H * HD), (uint64_t)(B * Tkv), (uint64_t)ldv * 2, HD, BKV, true))) return rc;
  if ((rc = make_tmap_2d(&tmdO, dout, 2, (uint64_t)(H * HD), (uint64_t)(B * Tq), (uint64_t)lddo * 2, HD, BQ, true))) return rc;
  if ((rc = make_tmap_2d(&tmDQ, dq_accum, 4, (uint64_t)(H * HD), (uint64_t)(B * Tq), (uint64_t)(H * HD) * 4, 32, BQ, true))) return rc;

  // dK / dV are stored per (batch, key tile): 3-D maps (columns, keys of one sample, batch) so that the 128-row box of the
  // last key tile is clipped at Tkv instead of running into the next sample
  if ((rc = make_tmap_3d(&tmDK, dk, 2, (uint64_t)(H * HD), (uint64_t)Tkv, (uint64_t)B, (uint64_t)lddk * 2, (uint64_t)Tkv * lddk * 2, HD,
                         BKV, 1, true))) return rc;
  if ((rc = make_tmap_3d(&tmDV, dv, 2, (uint64_t)(H * HD), (uint64_t)Tkv, (uint64_t)B, (uint64_t)lddv * 2, (uint64_t)Tkv * lddv * 2, HD,
                         BKV, 1, true))) return rc;

  const int64_t n_items = B * Tq * H;
  attention_delta_kernel<<<(unsigned)ceil_div(n_items * 8, 256), 256, 0, st>>>((const bf16*)o, (const bf16*)dout, delta, ldo, lddo,
                                                                              (int)B, (int)H, (int)Tq);
  OASR_LAUNCH_CHECK();
  OASR_CUDA_OK(cudaMemsetAsync(dq_accum, 0, sizeof(float) * B * Tq * H * HD, st));

  BwdParams p;
  p.lse = lse; p.delta = delta; p.dq_accum = dq_accum; p.kv_len = kv_len;
  p.B = (int)B; p.H = (int)H; p.Tq = (int)Tq; p.Tkv = (int)Tkv; p.causal = causal;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  static bool attr_set = false;
  if (!attr_set) {
    OASR_CUDA_OK(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    attr_set = true;
  }
  // one CTA per SM walking over (key tile, head, sample) items; OASR_BWD_PERSISTENT=0: one CTA per item (A/B)
  const int64_t n_work = ceil_div(Tkv, BKV) * H * B;
  OASR_REQUIRE(n_work < (int64_t(1) << 31), "attention_bwd: too many work items");
  const char* pe = getenv("OASR_BWD_PERSISTENT");   // read on every call: the host layer flips it when collectives share the GPU
  const bool persistent = !(pe && pe[0] == '0');
  int64_t ctas = num_sms();
  if (const char* cap = getenv("OASR_ATTN_MAX_CTAS")) {   // tests: few CTAs, so that small problems walk the multi-item path
    const long v = atol(cap);
    if (v > 0 && v < ctas) ctas = v;
  }
  const unsigned grid = persistent ? (unsigned)(n_work < ctas ? n_work : ctas) : (unsigned)n_work;
  attention_bwd_kernel<<<grid, BWD_THREADS, BWD_SMEM, st>>>(tmQ, tmK, tmV, tmdO, tmDQ, tmDK, tmDV, p, (int)n_work);
  OASR_LAUNCH_CHECK();
  const int64_t rows = B * Tq;
  int64_t blocks = ceil_div(rows * (H * HD / 8), 256);
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  f32_rows_to_bf16_kernel<<<(unsigned)blocks, 256, 0, st>>>(dq_accum, (bf16*)dq, rows, (int)(H * HD), lddq);
  OASR_LAUNCH_CHECK();
  return OASR_OK;
}
