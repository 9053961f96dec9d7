// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma
// (elected-lane issue, fp32 accumulators in TMEM, double-buffered) -> tcgen05.ld epilogue -> staged TMA stores.
//
//   D[M,N] = epi( sum_k A[m,k] B[n,k] )
//
// One CTA per SM, 128 + 32 * EW threads (EW = 8 epilogue warps is the only instantiated value; 16 warps with
// setmaxnreg-rebalanced registers measured no faster):
//   warp 0       TMA producer (warp-convergent loop, elect.sync picks the issuing lane)
//   warp 1       MMA issuer   (same)
//   warp 2       TMEM allocator / deallocator
//   warps 4..    epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 (one output row per thread) and the column
//                group (w-4)/4 of the tile; bf16 outputs go through four 2 KB 64B-swizzled staging slots per warp to
//                TMA bulk stores, residual / pre-activation tiles come in through the same slots by LDGSTS
// CLUSTER = 4 (opt-in, OASR_GEMM_CLUSTER=4): two such pairs stacked along M; CTAs h and h + 2 hold the same half of the
// B tile, each fetches a quarter and multicasts it to both; a slot's empty barrier collects both pairs' commits.
// CLUSTER = 2: two CTAs (an SM pair) compute a 256 x BN tile with ONE tcgen05.mma.cta_group::2 stream issued by the
// leader CTA: each CTA holds its 128 rows of A and its BN/2 rows of B in its own shared memory, so per flop the
// tensor core reads half as much smem and TMA writes half as much (the 1-CTA 128 x 256 tile needs 96 B/clk of operand
// reads plus 96 B/clk of TMA fills against a 128 B/clk shared-memory port; measured 1.30 vs cuBLAS 1.64 PFLOP/s).
// Both CTAs' TMA loads signal the leader's full barrier; the leader's commits are multicast to both CTAs' empty and
// accumulator-full barriers; every epilogue warp of the pair arrives on the leader's accumulator-empty barrier.
//
// Tile = 128 x BN, BK = 64 bf16 (= one 128-byte swizzle row).  Both operand majors are supported
// through the UMMA descriptors, so dgrad (B MN-major) and wgrad (A and B MN-major) need no
// transposes in HBM.  Reference ops replaced: olmoasr/model.py:97-101 (Linear.forward) and its
// autograd, model.py:768-770 (tied logits), model.py:592-593 (Conv1d after im2col).
#include "common.cuh"
#include "ptx_sm100.cuh"

#include <stdio.h>
#include <stdlib.h>

namespace oasr {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;               // 64 bf16 = 128 B = swizzle span
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB

struct GemmParams {
  void* C;
  void* C2;
  const float* bias;
  const void* aux;
  int64_t ldc;
  int64_t ldaux;
  int M, N, K;
  int tiles_m, tiles_n, splits;
  int kblocks_total, kblocks_per_split;
  int epi;
  int raster_m;  // 1: consecutive tiles walk M first (B tile stays hot in L2) -- used when B is the larger operand
  int tma_c;     // 1: bf16 outputs leave through the staging buffers + TMA stores (tmC / tmC2 are valid)
};

template <int BN, int CLUSTER = 1, int EW = 8>
struct Cfg {
  static constexpr int B_STAGE_BYTES = (BN / (CLUSTER >= 2 ? 2 : 1)) * BK * 2;   // per CTA (a pair splits B in two)
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  // bf16 epilogues of the 128/256-wide tiles leave through shared memory: 64 KB of 32-row x 64-byte staging slots
  // (SLOTS per epilogue warp), drained by TMA bulk stores (a per-thread 16-byte store touches 32 lines per warp
  // instruction -- 4096 LSU wavefronts per tile and output, which bounded the K = 1024 GEMMs)
  static constexpr int STG_BYTES = (BN >= 128) ? 65536 : 0;
  static constexpr int SLOTS = 65536 / (EW * 2048);               // 2 KB staging slots per epilogue warp (4 or 2)
  static constexpr int TAIL_BYTES = 2 * BN * 4 + 256;             // bias staging + mbarriers + TMEM slot
  static constexpr int BUDGET = 232448 - TAIL_BYTES - STG_BYTES;
  static constexpr int STAGES = (BUDGET / STAGE_BYTES) > 8 ? 8 : (BUDGET / STAGE_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + TAIL_BYTES;   // all dynamic, 1024-aligned base
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;   // two accumulator buffers
};

__device__ __forceinline__ void store_packed_bf16x32(bf16* dst, const uint32_t (&w)[16], int nvalid, bool vec) {
  if (vec && nvalid == 32) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int q = 0; q < 4; ++q) d4[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
  } else {
    uint16_t* d = reinterpret_cast<uint16_t*>(dst);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < nvalid) d[j] = static_cast<uint16_t>(w[j >> 1] >> (16 * (j & 1)));
  }
}

__device__ __forceinline__ void load_bf16x32(const bf16* src, float (&x)[32], int nvalid, bool vec) {
  if (vec && nvalid == 32) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 u = s4[q];
      float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      x[8 * q + 0] = a.x; x[8 * q + 1] = a.y; x[8 * q + 2] = b.x; x[8 * q + 3] = b.y;
      x[8 * q + 4] = c.x; x[8 * q + 5] = c.y; x[8 * q + 6] = d.x; x[8 * q + 7] = d.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = (j < nvalid) ? __bfloat162float(src[j]) : 0.f;
  }
}

// tile index -> (m block of this CTA, n block, k split).  Default order walks N fastest: the CTAs running together share
// A row panels and keep the (small) weight matrix in L2; raster_m walks M fastest for wide outputs (tied logits).
__device__ __forceinline__ void decode_tile(int tile, const GemmParams& p, int tiles_mc, int cluster, int cta_rank,
                                            int& m_blk, int& n_blk, int& split) {
  int mc;
  if (p.raster_m) {
    mc = tile % tiles_mc;
    n_blk = (tile / tiles_mc) % p.tiles_n;
  } else {
    n_blk = tile % p.tiles_n;
    mc = (tile / p.tiles_n) % tiles_mc;
  }
  split = tile / (p.tiles_n * tiles_mc);
  m_blk = mc * cluster + cta_rank;
}

template <int BN, int A_MN, int B_MN, int CLUSTER, int EW>
__global__ void __launch_bounds__(128 + 32 * EW, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2,
                    const GemmParams p) {
  using C = Cfg<BN, CLUSTER, EW>;
  constexpr int STAGES = C::STAGES;
  constexpr int PAIRW = CLUSTER >= 2 ? 2 : 1;                      // CTAs per MMA (cta_group)
  constexpr int NPAIR = CLUSTER / PAIRW;                          // MMA streams per cluster (CLUSTER = 4: two pairs share B)
  constexpr int CQ = EW / 4;                                      // column groups (4 warps = 128 TMEM lanes each)
  constexpr int ACT = (BN / 32 < CQ) ? BN / 32 : CQ;              // groups that own at least one 32-column chunk
  constexpr int NCH = BN / (32 * ACT);                            // chunks per epilogue warp
  constexpr int SLOTS = C::SLOTS;

  // all shared memory is dynamic so that the 128B-swizzled regions start on a 1024-byte boundary without slack:
  //   [STAGES x (A | B)] [8 warps x 2 x 4 KB output staging] [bias 2 x BN f32] [mbarriers] [TMEM slot]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* tail = smem_raw + STAGES * C::STAGE_BYTES + C::STG_BYTES;
  float (*s_bias)[BN] = reinterpret_cast<float (*)[BN]>(tail);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(tail + 2 * BN * 4);
  uint64_t* bar_empty = bar_full + STAGES;
  uint64_t* bar_tmem_full = bar_empty + STAGES;
  uint64_t* bar_tmem_empty = bar_tmem_full + 2;
  uint32_t& tmem_base_slot = *reinterpret_cast<uint32_t*>(bar_tmem_empty + 2);
  static_assert((2 * STAGES + 4) * 8 + 4 <= 256, "barrier area");

  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&tmA);
    ptx::tma_prefetch_desc(&tmB);
    if (p.tma_c) {
      ptx::tma_prefetch_desc(&tmC);
      if (p.epi == OASR_EPI_BF16_GELU) ptx::tma_prefetch_desc(&tmC2);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bar_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bar_empty[s]), NPAIR);   // one commit per MMA stream that reads this slot's B
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bar_tmem_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bar_tmem_empty[s]), 4 * ACT * PAIRW);  // one arrive per active epilogue warp of the pair
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    if (CLUSTER == 1) {
      ptx::tmem_alloc<C::TMEM_COLS>(ptx::smem_u32(&tmem_base_slot));
      ptx::tmem_relinquish();
    } else {  // the same warp of both CTAs allocates the pair's columns
      ptx::tmem_alloc2<C::TMEM_COLS>(ptx::smem_u32(&tmem_base_slot));
      ptx::tmem_relinquish2();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) ptx::cluster_sync();  // peer barriers are initialised before any remote signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  if (EW == 16) {   // 640 threads start at 96 registers; the control warpgroup hands 32 each to the 16 epilogue warps
    if (warp < 4) ptx::setmaxnreg_dec<64>(); else ptx::setmaxnreg_inc<104>();
  }

  // tile walk: with CLUSTER = 2 the pair (2*pm, 2*pm+1) of M tiles shares n_blk; both CTAs see the same sequence
  const uint32_t cta_rank = (CLUSTER > 1) ? ptx::cluster_ctarank() : 0;
  const int tiles_mc = (p.tiles_m + CLUSTER - 1) / CLUSTER;
  const int num_tiles = tiles_mc * p.tiles_n * p.splits;
  const int tile0 = blockIdx.x / CLUSTER;
  const int tile_step = gridDim.x / CLUSTER;
  constexpr uint16_t kMask = (1u << CLUSTER) - 1;                 // every CTA whose smem slot the commit frees
  const uint16_t pair_mask = static_cast<uint16_t>(0x3u << (cta_rank & ~1u));   // this CTA's MMA pair

  if (warp == 0) {
    // ===================== TMA producer (warp-convergent loop, one elected lane issues) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        int n_blk, m_blk, split;
        decode_tile(tile, p, tiles_mc, CLUSTER, cta_rank, m_blk, n_blk, split);
        const int kb0 = split * p.kblocks_per_split;
        const int kb1 = min(kb0 + p.kblocks_per_split, p.kblocks_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&bar_empty[stage]), phase ^ 1);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + A_STAGE_BYTES;
          if (ptx::elect_one()) {
          if (CLUSTER == 1) {
            const uint32_t full = ptx::smem_u32(&bar_full[stage]);
            ptx::mbar_arrive_expect_tx(full, C::STAGE_BYTES);
            if (A_MN) {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i)
                ptx::tma_load_2d(sa + i * (BK * 128), &tmA, full, m_blk * BM + i * 64, kb * BK);
            } else {
              ptx::tma_load_2d(sa, &tmA, full, kb * BK, m_blk * BM);
            }
            if (B_MN) {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)
                ptx::tma_load_2d(sb + i * (BK * 128), &tmB, full, n_blk * BN + i * 64, kb * BK);
            } else {
              ptx::tma_load_2d(sb, &tmB, full, kb * BK, n_blk * BN);
            }
          } else {
            // both CTAs fill their own smem; every byte is accounted on the pair LEADER's full barrier
            const uint32_t full = ptx::smem_u32(&bar_full[stage]) & ptx::kPeerBitMask;
            if ((cta_rank & 1) == 0) ptx::mbar_arrive_expect_tx(full, PAIRW * C::STAGE_BYTES);
            if (A_MN) {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i)
                ptx::tma_load_2d_2sm(sa + i * (BK * 128), &tmA, full, m_blk * BM + i * 64, kb * BK);
            } else {
              ptx::tma_load_2d_2sm(sa, &tmA, full, kb * BK, m_blk * BM);
            }
            const int h = cta_rank & 1;            // which half of the B tile this CTA's MMA operand is
            if (NPAIR == 1) {
              if (B_MN) {   // this CTA's BN/2 columns of B: BN/128 chunks of 64
#pragma unroll
                for (int i = 0; i < BN / 128; ++i)
                  ptx::tma_load_2d_2sm(sb + i * (BK * 128), &tmB, full, n_blk * BN + (h * (BN / 128) + i) * 64, kb * BK);
              } else {      // this CTA's BN/2 rows of B
                ptx::tma_load_2d_2sm(sb, &tmB, full, kb * BK, n_blk * BN + h * (BN / 2));
              }
            } else {
              // two pairs stacked along M use the same B tile: CTAs h and h + 2 hold identical halves, each fetches
              // one quarter and multicasts it to both (halves the L2 -> SM traffic of B)
              const int pr = cta_rank >> 1;
              const uint16_t mc = static_cast<uint16_t>((1u << h) | (1u << (h + 2)));
              if (B_MN) {   // BN = 256: this half is two 64-column chunks, one per pair
                ptx::tma_load_2d_2sm_mc(sb + pr * (BK * 128), &tmB, full, n_blk * BN + (h * 2 + pr) * 64, kb * BK, mc);
              } else {      // BN/4 rows (8-row swizzle atoms stay aligned: BN/4 * 128 B is a multiple of 1024)
                ptx::tma_load_2d_2sm_mc(sb + pr * (BN / 4) * 128, &tmB, full, kb * BK, n_blk * BN + h * (BN / 2) + pr * (BN / 4), mc);
              }
            }
          }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only when paired) =====================
    // The whole warp runs the loop (all lanes wait on the barriers); elect.sync picks the issuing lane so that the
    // descriptor arithmetic and the UTCHMMAs stay on the uniform datapath.
    if ((cta_rank & 1) == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM * PAIRW, BN, A_MN, B_MN);
      // K-major  : 8-row groups 1024 B apart (SBO); LBO unused for swizzled K-major (CUTLASS sets 1)
      // MN-major : 64-element MN chunks BK*128 B apart (LBO); 8-k groups 1024 B apart (SBO)
      constexpr uint32_t a_lbo = A_MN ? BK * 128 : 16, a_sbo = 1024;
      constexpr uint32_t b_lbo = B_MN ? BK * 128 : 16, b_sbo = 1024;
      constexpr uint32_t a_kstep = (A_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;  // descriptor-lo units per UMMA_K advance
      constexpr uint32_t b_kstep = (B_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
      uint32_t a_lo0, a_hi, b_lo0, b_hi;
      ptx::umma_desc_sw128_lh(smem_base, a_lbo, a_sbo, a_lo0, a_hi);
      ptx::umma_desc_sw128_lh(smem_base + A_STAGE_BYTES, b_lbo, b_sbo, b_lo0, b_hi);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int split = tile / (p.tiles_n * tiles_mc);
        const int kb0 = split * p.kblocks_per_split;
        const int kb1 = min(kb0 + p.kblocks_per_split, p.kblocks_total);
        ptx::mbar_wait(ptx::smem_u32(&bar_tmem_empty[acc]), acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&bar_full[stage]), phase);
          ptx::tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (C::STAGE_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + stage * (C::STAGE_BYTES >> 4);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              if (CLUSTER == 1) ptx::tc_mma_f16_lh(d_tmem, a_lo + k * a_kstep, a_hi, b_lo + k * b_kstep, b_hi, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              else ptx::tc_mma2_f16_lh(d_tmem, a_lo + k * a_kstep, a_hi, b_lo + k * b_kstep, b_hi, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            // frees the smem slot (in both CTAs of a pair) when the MMAs retire
            if (CLUSTER == 1) ptx::tc_commit(ptx::smem_u32(&bar_empty[stage]));
            else ptx::tc_commit2_mc(ptx::smem_u32(&bar_empty[stage]), kMask);
            if (kb + 1 == kb1) {   // accumulator complete (each CTA's epilogue drains its own 128 rows)
              if (CLUSTER == 1) ptx::tc_commit(ptx::smem_u32(&bar_tmem_full[acc]));
              else ptx::tc_commit2_mc(ptx::smem_u32(&bar_tmem_full[acc]), pair_mask);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4 && ((warp - 4) >> 2) < ACT) {
    // ===================== epilogue =====================
    const int q = warp & 3;            // TMEM lane quarter this warp may touch
    const int chalf = (warp - 4) >> 2; // which group of the tile's columns this warp drains
    const int etid = threadIdx.x - 128;
    const int row_in_tile = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool c_is_f32 = (p.epi == OASR_EPI_F32 || p.epi == OASR_EPI_F32_ATOMIC_ADD);
    const bool vec_c = c_is_f32 ? ((p.ldc & 3) == 0) : ((p.ldc & 7) == 0);
    const bool vec_aux = (p.ldaux & 7) == 0;
    const bool use_aux = (p.epi == OASR_EPI_BF16_RESIDUAL || p.epi == OASR_EPI_BF16_GELU_BWD);
    const bool use_tma = (BN >= 128) && p.tma_c != 0;
    // this warp's staging area: four 2 KB slots, each one 32-row x 32-column bf16 chunk (64-byte rows, 64B swizzle).
    // A chunk is handed to the TMA engine as soon as it is staged, and a slot is refilled two (GELU: o1 in slots 0/1,
    // o2 in slots 2/3) or four chunks later, so the drain latency of the bulk store is never waited for.
    const uint32_t stg = smem_base + STAGES * C::STAGE_BYTES + (warp - 4) * (SLOTS * 2048);
    uint32_t slotc = 0;   // chunks staged so far by this warp
    // residual / pre-activation chunks come in through the same slots: coalesced 16-byte LDGSTS (8 rows x 64 B per
    // warp instruction instead of 32 partial lines), then each thread reads its own row back
    const bool aux_stage = use_tma && use_aux && vec_aux && (p.N & 7) == 0 && (reinterpret_cast<uintptr_t>(p.aux) & 15) == 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      int n_blk, m_blk, split_unused;
      decode_tile(tile, p, tiles_mc, CLUSTER, cta_rank, m_blk, n_blk, split_unused);
      const int row = m_blk * BM + row_in_tile;
      const bool row_ok = row < p.M;
      if (p.bias != nullptr) {  // stage this tile's bias once (rounded to bf16 like bias.to(x.dtype) unless fp32 output)
        for (int i = etid; i < BN; i += 128 * ACT) {
          const int col = n_blk * BN + i;
          const float b = col < p.N ? __ldg(p.bias + col) : 0.f;
          s_bias[acc][i] = c_is_f32 ? b : bf16_round(b);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(128 * ACT) : "memory");  // active epilogue warps only
      }
      // aux (residual / pre-activation) rows of this thread, fetched before the accumulator is ready so that the
      // ~1 us global-load latency hides behind the MMAs (it used to be the top stall of the residual epilogue)
      auto issue_aux = [&](int ci, uint32_t buf) {   // aux columns of chunk ci of this warp -> slot `buf` (swizzled)
        const int u = lane & 3, r0 = lane >> 2;
        const int colu = n_blk * BN + (chalf * NCH + ci) * 32 + u * 8;
        const bool col_ok = colu + 8 <= p.N;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ri = r0 + 8 * i;
          const int grow = m_blk * BM + q * 32 + ri;
          const bool ok = col_ok && grow < p.M;
          const bf16* src = reinterpret_cast<const bf16*>(p.aux) + (ok ? static_cast<int64_t>(grow) * p.ldaux + colu : 0);
          ptx::cp_async_16(buf + ri * 64 + ((u ^ ((ri >> 1) & 3)) << 4), src, ok ? 16u : 0u);
        }
        ptx::cp_async_commit();
      };
      constexpr int PRE = (SLOTS == 2) ? NCH : 1;   // chunks whose aux is requested before the accumulator is ready
      static_assert(PRE <= SLOTS, "aux prefetch depth");
      if (aux_stage) {   // aux travels while the accumulator is still being computed
        if (lane == 0) ptx::tma_store_wait_read<SLOTS - PRE>();   // slot of chunk k is free once store k - SLOTS was read
        __syncwarp();
#pragma unroll
        for (int j = 0; j < PRE; ++j)
          if (n_blk * BN + (chalf * NCH + j) * 32 < p.N) issue_aux(j, stg + ((slotc + j) % SLOTS) * 2048u);
      } else if (use_aux && row_ok) {   // direct path: pull this thread's aux bytes towards L2
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const int col = n_blk * BN + (chalf * NCH + ci) * 32;
          if (col < p.N)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const bf16*>(p.aux) + static_cast<int64_t>(row) * p.ldaux + col));
        }
      }
      ptx::mbar_wait(ptx::smem_u32(&bar_tmem_full[acc]), acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int ci = 0; ci < NCH; ++ci) {
        const int c = chalf * NCH + ci;
        const int col = n_blk * BN + c * 32;
        const int nvalid = min(32, p.N - col);
        if (use_tma && nvalid <= 0) break;   // this chunk and everything right of it is outside N
        const bool aux_here = aux_stage || (use_aux && vec_aux && row_ok && (col + 32 <= p.N));
        uint4 auxc[4];
        if (aux_here && !aux_stage) {   // issued before the TMEM load so both latencies overlap
          const uint4* ap4 = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.aux) + static_cast<int64_t>(row) * p.ldaux + col);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) auxc[q4] = ap4[q4];
        }
        uint32_t r[32];
        __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the divergent `continue`
        ptx::tc_ld_32x32b_x32(t_row + c * 32, r);
        if (aux_stage) {
          ptx::cp_async_wait_all();   // this chunk's LDGSTS have landed (all lanes' copies: wait, then warp barrier)
          __syncwarp();
          const uint32_t ab = stg + (slotc & 3u) * 2048u + lane * 64;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) ptx::ld_shared_v4(ab + ((q4 ^ ((lane >> 1) & 3)) << 4), auxc[q4]);
          if (ci + 1 < NCH && col + 32 < p.N) {   // next chunk's aux into the next slot (its last store was 3 groups ago)
            if (lane == 0) ptx::tma_store_wait_read<2>();
            __syncwarp();
            issue_aux(ci + 1, stg + ((slotc + 1) & 3u) * 2048u);
          }
        }
        ptx::tc_wait_ld();
        if (!use_tma && (!row_ok || nvalid <= 0)) continue;
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(r[j]);
        if (p.bias != nullptr) {
          const float4* b4 = reinterpret_cast<const float4*>(&s_bias[acc][c * 32]);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = b4[j];
            const float2 lo = __fadd2_rn(make_float2(x[4 * j], x[4 * j + 1]), make_float2(b.x, b.y));
            const float2 hi = __fadd2_rn(make_float2(x[4 * j + 2], x[4 * j + 3]), make_float2(b.z, b.w));
            x[4 * j] = lo.x; x[4 * j + 1] = lo.y; x[4 * j + 2] = hi.x; x[4 * j + 3] = hi.y;
          }
        }
        const int64_t off = static_cast<int64_t>(row) * p.ldc + col;
        if (c_is_f32) {
          float* dst = reinterpret_cast<float*>(p.C) + off;
          if (p.epi == OASR_EPI_F32) {
            if (vec_c && nvalid == 32) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                reinterpret_cast<float4*>(dst)[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nvalid) dst[j] = x[j];
            }
          } else {
            if (vec_c && nvalid == 32) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                             ::"l"(dst + 4 * j), "f"(x[4 * j]), "f"(x[4 * j + 1]), "f"(x[4 * j + 2]),
                               "f"(x[4 * j + 3])
                             : "memory");
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nvalid) atomicAdd(dst + j, x[j]);
            }
          }
          continue;
        }
        // ---- bf16 epilogues: every variant ends in packed words o1 (and o2 = GELU output), two columns per word
        uint32_t o1[16], o2[16];
        const int nload = row_ok ? max(nvalid, 0) : 0;   // aux elements the scalar fallback may touch
        switch (p.epi) {
          case OASR_EPI_BF16_GELU: {   // h = bf16(acc + b), g = bf16(gelu(h))
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              o1[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
              const float2 g2 = gelu_erf2(unpack_bf16x2(o1[j]));
              o2[j] = pack_bf16x2(g2.x, g2.y);
            }
            break;
          }
          case OASR_EPI_BF16_RESIDUAL: {   // C = bf16(aux + bf16(acc + b))
            if (!aux_here) {
              float a[32];
              load_bf16x32(reinterpret_cast<const bf16*>(p.aux) + static_cast<int64_t>(row_ok ? row : 0) * p.ldaux + col, a, nload, false);
#pragma unroll
              for (int j = 0; j < 16; ++j) (&auxc[j >> 2].x)[j & 3] = pack_bf16x2(a[2 * j], a[2 * j + 1]);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 y = unpack_bf16x2(pack_bf16x2(x[2 * j], x[2 * j + 1]));
              const float2 o = __fadd2_rn(unpack_bf16x2((&auxc[j >> 2].x)[j & 3]), y);
              o1[j] = pack_bf16x2(o.x, o.y);
            }
            break;
          }
          case OASR_EPI_BF16_GELU_BWD: {   // C = bf16(bf16(acc) * gelu'(aux))
            if (!aux_here) {
              float a[32];
              load_bf16x32(reinterpret_cast<const bf16*>(p.aux) + static_cast<int64_t>(row_ok ? row : 0) * p.ldaux + col, a, nload, false);
#pragma unroll
              for (int j = 0; j < 16; ++j) (&auxc[j >> 2].x)[j & 3] = pack_bf16x2(a[2 * j], a[2 * j + 1]);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 dg = unpack_bf16x2(pack_bf16x2(x[2 * j], x[2 * j + 1]));   // bf16(acc): the matmul-backward output
              const float2 gr = gelu_erf_grad2(unpack_bf16x2((&auxc[j >> 2].x)[j & 3]));
              const float2 o = __fmul2_rn(dg, gr);
              o1[j] = pack_bf16x2(o.x, o.y);
            }
            break;
          }
          default: {
#pragma unroll
            for (int j = 0; j < 16; ++j) o1[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
            break;
          }
        }
        if (use_tma) {
          const bool gelu = p.epi == OASR_EPI_BF16_GELU;
          if (!aux_stage) {   // the slot(s) about to be refilled must have been read out by the engine
            if (lane == 0) {
              if (gelu) ptx::tma_store_wait_read<1>(); else ptx::tma_store_wait_read<3>();
            }
            __syncwarp();
          }
          const uint32_t b1 = stg + (gelu ? (slotc & 1u) : (slotc & 3u)) * 2048u;
          const uint32_t b2 = b1 + 4096u;
          const uint32_t rowa = lane * 64;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {   // 64B swizzle: 16-byte unit index ^ ((row >> 1) & 3)
            const uint32_t sw = ((q4 ^ ((lane >> 1) & 3)) << 4) + rowa;
            ptx::st_shared_v4(b1 + sw, o1[4 * q4], o1[4 * q4 + 1], o1[4 * q4 + 2], o1[4 * q4 + 3]);
            if (gelu) ptx::st_shared_v4(b2 + sw, o2[4 * q4], o2[4 * q4 + 1], o2[4 * q4 + 2], o2[4 * q4 + 3]);
          }
          ptx::fence_proxy_async_smem();   // chunk complete: hand it to the TMA engine (rows >= M / columns >= N are clipped)
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tmC, b1, col, m_blk * BM + q * 32);
            if (gelu) ptx::tma_store_2d(&tmC2, b2, col, m_blk * BM + q * 32);
            ptx::tma_store_commit();
          }
          ++slotc;
        } else {
          store_packed_bf16x32(reinterpret_cast<bf16*>(p.C) + off, o1, nvalid, vec_c);
          if (p.epi == OASR_EPI_BF16_GELU) store_packed_bf16x32(reinterpret_cast<bf16*>(p.C2) + off, o2, nvalid, vec_c);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CLUSTER == 1) ptx::mbar_arrive(ptx::smem_u32(&bar_tmem_empty[acc]));
        else ptx::mbar_arrive_cluster(ptx::smem_u32(&bar_tmem_empty[acc]) & ptx::kPeerBitMask);  // the leader's barrier
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (use_tma && lane == 0) ptx::tma_store_wait<0>();   // shared memory stays valid until the engine has drained it
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) ptx::cluster_sync();  // no CTA exits while its peer can still multicast into it
  if (warp == 2) {
    ptx::tc_fence_after();
    if (CLUSTER == 1) ptx::tmem_dealloc<C::TMEM_COLS>(tmem_base);
    else ptx::tmem_dealloc2<C::TMEM_COLS>(tmem_base);
  }
}

template <int BN, int A_MN, int B_MN, int CLUSTER, int EW>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmC2,
           const GemmParams& p, cudaStream_t st) {
  using C = Cfg<BN, CLUSTER, EW>;
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN, CLUSTER, EW>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    OASR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int tiles = ((p.tiles_m + CLUSTER - 1) / CLUSTER) * p.tiles_n * p.splits;  // cluster tiles
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(128 + 32 * EW);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent grid = the clusters that are co-resident (a GPC whose SM count is not a multiple of the cluster size
  // cannot host one on its leftover SMs, so this can be fewer than num_sms / CLUSTER)
  static int max_clusters = 0;  // per instantiation
  if (max_clusters == 0) {
    int n = 0;
    cfg.gridDim = dim3(num_sms() / CLUSTER * CLUSTER);
    if (CLUSTER > 1 && cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0)
      max_clusters = n < num_sms() / CLUSTER ? n : num_sms() / CLUSTER;
    else
      max_clusters = num_sms() / CLUSTER;
    (void)cudaGetLastError();
    if (getenv("OASR_GEMM_VERBOSE"))
      fprintf(stderr, "oasr gemm<BN=%d,cluster=%d>: %d co-resident clusters (occupancy query %d), %d stages\n", BN, CLUSTER,
              max_clusters, n, C::STAGES);
  }
  int use_clusters = gemm_sm_budget() / CLUSTER;
  if (use_clusters > max_clusters) use_clusters = max_clusters;
  if (use_clusters < 1) use_clusters = 1;
  const int grid = (tiles < use_clusters ? tiles : use_clusters) * CLUSTER;
  cfg.gridDim = dim3(grid);
  OASR_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmC2, p));
  return OASR_OK;
}

template <int BN, int CLUSTER, int EW>
int dispatch_major(int a_mn, int b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                   const CUtensorMap& tmC2, const GemmParams& p, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch<BN, 0, 0, CLUSTER, EW>(tmA, tmB, tmC, tmC2, p, st);
  if (!a_mn && b_mn) return launch<BN, 0, 1, CLUSTER, EW>(tmA, tmB, tmC, tmC2, p, st);
  if (a_mn && b_mn) return launch<BN, 1, 1, CLUSTER, EW>(tmA, tmB, tmC, tmC2, p, st);
  return launch<BN, 1, 0, CLUSTER, EW>(tmA, tmB, tmC, tmC2, p, st);
}

}  // namespace
}  // namespace oasr

using namespace oasr;

extern "C" int oasr_gemm_bf16(const void* A, int64_t lda, int a_layout, const void* B, int64_t ldb,
                              int b_layout, void* Cout, int64_t ldc, void* C2, const float* bias,
                              const void* aux, int64_t ldaux, int64_t M, int64_t N, int64_t K,
                              int epilogue, int split_k, int block_n, void* stream) {
  OASR_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%ld N=%ld K=%ld", (long)M, (long)N, (long)K);
  OASR_REQUIRE(A && B && Cout, "gemm: null operand");
  OASR_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
               "gemm: A/B must be 16-byte aligned");
  OASR_REQUIRE((lda & 7) == 0 && (ldb & 7) == 0, "gemm: lda/ldb must be multiples of 8 elements (TMA)");
  OASR_REQUIRE(epilogue >= OASR_EPI_BF16 && epilogue <= OASR_EPI_F32_ATOMIC_ADD, "gemm: bad epilogue %d", epilogue);
  OASR_REQUIRE(a_layout == OASR_K_MAJOR || a_layout == OASR_MN_MAJOR, "gemm: bad a_layout");
  OASR_REQUIRE(b_layout == OASR_K_MAJOR || b_layout == OASR_MN_MAJOR, "gemm: bad b_layout");
  if (epilogue == OASR_EPI_BF16_GELU) OASR_REQUIRE(C2 != nullptr, "gemm: GELU epilogue needs C2");
  if (epilogue == OASR_EPI_BF16_RESIDUAL || epilogue == OASR_EPI_BF16_GELU_BWD)
    OASR_REQUIRE(aux != nullptr, "gemm: epilogue %d needs aux", epilogue);
  if (split_k < 1) split_k = 1;
  OASR_REQUIRE(split_k == 1 || epilogue == OASR_EPI_F32_ATOMIC_ADD, "gemm: split_k needs the atomic-add epilogue");
  if (block_n == 0) block_n = (N >= 256) ? 256 : (N > 64 ? 128 : 64);
  OASR_REQUIRE(block_n == 64 || block_n == 128 || block_n == 256, "gemm: block_n must be 64/128/256");

  GemmParams p;
  p.C = Cout; p.C2 = C2; p.bias = bias; p.aux = aux; p.ldc = ldc; p.ldaux = ldaux;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.tiles_m = (int)ceil_div(M, BM);
  p.tiles_n = (int)ceil_div(N, block_n);
  p.kblocks_total = (int)ceil_div(K, BK);
  if (split_k > p.kblocks_total) split_k = p.kblocks_total;
  p.kblocks_per_split = (int)ceil_div(p.kblocks_total, split_k);
  p.splits = (int)ceil_div(p.kblocks_total, p.kblocks_per_split);
  p.epi = epilogue;
  p.raster_m = (N > M) ? 1 : 0;

  CUtensorMap tmA, tmB;
  int rc;
  if (a_layout == OASR_K_MAJOR)
    rc = make_tmap_2d(&tmA, A, 2, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, BK, BM, true);
  else
    rc = make_tmap_2d(&tmA, A, 2, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 64, BK, true);
  if (rc) return rc;
  // cluster shape: 4 = two CTA pairs stacked along M that share (multicast) the B tile, 2 = one CTA pair, 1 = single CTA.
  // OASR_GEMM_CLUSTER caps it (read per call so that tools/ab_step.py can alternate in-process).  Default 2: the
  // 4-CTA shape moves 25 % fewer operand bytes from L2 and is ~9 % faster per SM, but only 33 such clusters are
  // co-resident on a B200 (132 of 148 SMs; GPCs with an odd TPC count strand one TPC each), and the whole training
  // step measured 203.4 ms against 201.9 ms for pairs (profiles/r01_ab_gemm_cluster.txt).
  const int env_cluster = [] { const char* e = getenv("OASR_GEMM_CLUSTER"); return e ? atoi(e) : 2; }();
  int cluster = 1;
  if (env_cluster >= 2 && p.tiles_m >= 2 && block_n >= 128) cluster = 2;
  if (env_cluster >= 4 && p.tiles_m >= 4 && block_n == 256) cluster = 4;
  if (b_layout == OASR_K_MAJOR)  // a CTA fetches its share of the B tile's rows: all / half (pair) / quarter (two pairs)
    rc = make_tmap_2d(&tmB, B, 2, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, BK, block_n / cluster, true);
  else
    rc = make_tmap_2d(&tmB, B, 2, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, 64, BK, true);
  if (rc) return rc;

  // bf16 outputs of the 128/256-wide tiles are written by TMA from the epilogue's staging buffers (box 32 x 32, 64B
  // swizzle; rows >= M / columns >= N are clipped by the engine).  Needs a 16-byte aligned C with ldc % 8 == 0.
  CUtensorMap tmC = {}, tmC2 = {};
  static const int env_tma_c = [] { const char* e = getenv("OASR_GEMM_TMA_STORE"); return e ? atoi(e) : 1; }();
  const bool bf16_out = epilogue <= OASR_EPI_BF16_GELU_BWD;
  p.tma_c = (env_tma_c && bf16_out && block_n >= 128 && (ldc & 7) == 0 && (reinterpret_cast<uintptr_t>(Cout) & 15) == 0 &&
             (epilogue != OASR_EPI_BF16_GELU || (reinterpret_cast<uintptr_t>(C2) & 15) == 0)) ? 1 : 0;
  if (p.tma_c) {
    rc = make_tmap_2d_sw(&tmC, Cout, 2, (uint64_t)N, (uint64_t)M, (uint64_t)ldc * 2, 32, 32, 64);
    if (rc) return rc;
    if (epilogue == OASR_EPI_BF16_GELU) {
      rc = make_tmap_2d_sw(&tmC2, C2, 2, (uint64_t)N, (uint64_t)M, (uint64_t)ldc * 2, 32, 32, 64);
      if (rc) return rc;
    }
  }

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define OASR_GEMM_DISPATCH(BNv, CLv) return dispatch_major<BNv, CLv, 8>(a_layout, b_layout, tmA, tmB, tmC, tmC2, p, st)
  switch (block_n) {
    case 256:
      if (cluster == 4) OASR_GEMM_DISPATCH(256, 4);
      if (cluster == 2) OASR_GEMM_DISPATCH(256, 2);
      OASR_GEMM_DISPATCH(256, 1);
    case 128:
      if (cluster == 2) OASR_GEMM_DISPATCH(128, 2);
      OASR_GEMM_DISPATCH(128, 1);
    default:
      OASR_GEMM_DISPATCH(64, 1);
  }
#undef OASR_GEMM_DISPATCH
}
