// Flash-style attention forward for head_dim 64 on tcgen05 / TMEM (sm_100a).
//
//   O = softmax(Q K^T * scale + mask) V        per (batch, head); bf16 in/out, fp32 softmax
//
// Replaces F.scaled_dot_product_attention in MultiHeadAttention.forward (olmoasr/model.py:331-340):
//   encoder self-attention (no mask), decoder self-attention (causal + per-sample key length, derived
//   from the dense additive mask the reference passes, model.py:740-743) and cross-attention (no mask).
//
// CTA = 128 query rows x one (b, h); loops over 128-key tiles.  192 threads:
//   warp 0      TMA producer (Q once, K/V double-buffered)      warp 1   tcgen05.mma issuer + TMEM owner
//   warps 2..5  softmax: thread r owns score row r (TMEM lane r) -> no cross-thread reductions at all
// TMEM: S (128 cols fp32) + PV (64 cols fp32) -> 256-column allocation, two CTAs per SM so one CTA's
// softmax overlaps the other's MMAs.  P is written to smem as bf16 in the 128B-swizzled K-major layout
// and fed back as the A operand of the P*V MMA; V is consumed MN-major straight from its TMA tile.
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace oasr {
namespace {

constexpr int HD = 64;
constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int TILE_BYTES = 128 * HD * 2;     // 16 KB: a [128 rows][64 bf16] swizzled tile
constexpr int P_BYTES = BQ * BKV * 2;        // 32 KB: two 64-key halves of [128][128B]
constexpr int ATT_TILES = TILE_BYTES /*Q*/ + 2 * 2 * TILE_BYTES /*K,V x2*/ + P_BYTES;  // 112 KB
constexpr int ATT_SMEM = ATT_TILES + 128;  // + barriers; 2 CTAs/SM => no static smem, no alignment slack
constexpr int TMEM_COLS = 256;
constexpr int S_COL = 0, PV_COL = 128;

struct AttnParams {
  bf16* o;
  float* lse;        // (B, H, Tq) log2-domain log-sum-exp, nullable
  const int32_t* kv_len;  // (B,) valid keys per sample, nullable
  int64_t ldo;
  int B, H, Tq, Tkv;
  int causal;
  float scale_log2;  // scale * log2(e)
};

__global__ void __launch_bounds__(192, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + ATT_TILES);
  uint64_t& bar_q = bars[0];
  uint64_t& bar_s = bars[1];
  uint64_t& bar_p = bars[2];
  uint64_t& bar_pv = bars[3];
  uint64_t* bar_kv_full = bars + 4;
  uint64_t* bar_kv_empty = bars + 6;
  uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(bars + 8);

  const uint32_t sbase = ptx::smem_u32(smem_raw);
  if ((sbase & 1023u) != 0) {  // swizzled tiles need 1 KB alignment; the declaration above should guarantee it
    if (threadIdx.x == 0) printf("oasr attention: dynamic smem base %u not 1 KB aligned\n", sbase);
    __trap();
  }
  const uint32_t sQ = sbase;
  const uint32_t sK0 = sQ + TILE_BYTES;              // stage s: K at sK0 + s*32K, V right after K
  const uint32_t sP = sK0 + 4 * TILE_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = q_tile * BQ;

  int kv_valid = p.Tkv;
  if (p.kv_len) kv_valid = min(kv_valid, max(1, p.kv_len[b]));
  int kv_end = kv_valid;
  if (p.causal) kv_end = min(kv_end, q0 + BQ);
  const int n_kv = (kv_end + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&tmQ); ptx::tma_prefetch_desc(&tmK); ptx::tma_prefetch_desc(&tmV);
    ptx::mbar_init(ptx::smem_u32(&bar_q), 1);
    ptx::mbar_init(ptx::smem_u32(&bar_s), 1);
    ptx::mbar_init(ptx::smem_u32(&bar_p), 4);
    ptx::mbar_init(ptx::smem_u32(&bar_pv), 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bar_kv_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bar_kv_empty[s]), 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc<TMEM_COLS>(ptx::smem_u32(&tmem_slot));
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int qrow = b * p.Tq + q0;
      ptx::mbar_arrive_expect_tx(ptx::smem_u32(&bar_q), TILE_BYTES);
      ptx::tma_load_2d(sQ, &tmQ, ptx::smem_u32(&bar_q), h * HD, qrow);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        ptx::mbar_wait(ptx::smem_u32(&bar_kv_empty[s]), ((j >> 1) & 1) ^ 1);
        const uint32_t full = ptx::smem_u32(&bar_kv_full[s]);
        ptx::mbar_arrive_expect_tx(full, 2 * TILE_BYTES);
        const int krow = b * p.Tkv + j * BKV;
        ptx::tma_load_2d(sK0 + s * 2 * TILE_BYTES, &tmK, full, h * HD, krow);
        ptx::tma_load_2d(sK0 + s * 2 * TILE_BYTES + TILE_BYTES, &tmV, full, h * HD, krow);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = ptx::umma_idesc_bf16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_pv = ptx::umma_idesc_bf16(BQ, HD, 0, 1);
      auto issue_qk = [&](int j) {
        const uint32_t sK = sK0 + (j & 1) * 2 * TILE_BYTES;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          ptx::tc_mma_f16(tmem + S_COL, ptx::umma_smem_desc_sw128(sQ + k * 32, 16, 1024),
                          ptx::umma_smem_desc_sw128(sK + k * 32, 16, 1024), idesc_qk, k > 0);
        ptx::tc_commit(ptx::smem_u32(&bar_s));
      };
      ptx::mbar_wait(ptx::smem_u32(&bar_q), 0);
      ptx::mbar_wait(ptx::smem_u32(&bar_kv_full[0]), 0);
      ptx::tc_fence_after();
      issue_qk(0);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        ptx::mbar_wait(ptx::smem_u32(&bar_p), j & 1);   // P(j) in smem, S and PV TMEM regions free
        ptx::tc_fence_after();
        const uint32_t sV = sK0 + s * 2 * TILE_BYTES + TILE_BYTES;
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)
          ptx::tc_mma_f16(tmem + PV_COL,
                          ptx::umma_smem_desc_sw128(sP + (k >> 2) * (P_BYTES / 2) + (k & 3) * 32, 16, 1024),
                          ptx::umma_smem_desc_sw128(sV + k * 2048, BKV * 128, 1024), idesc_pv, k > 0);
        ptx::tc_commit(ptx::smem_u32(&bar_pv));
        ptx::tc_commit(ptx::smem_u32(&bar_kv_empty[s]));
        if (j + 1 < n_kv) {
          ptx::mbar_wait(ptx::smem_u32(&bar_kv_full[(j + 1) & 1]), ((j + 1) >> 1) & 1);
          ptx::tc_fence_after();
          issue_qk(j + 1);
        }
      }
    }
  } else {
    // ----------------------------- softmax / output warps -----------------------------
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;               // row within the tile == TMEM lane
    const int qi = q0 + r;                           // query index within the sequence
    const uint32_t t_lane = static_cast<uint32_t>(quarter * 32) << 16;
    const float c = p.scale_log2;
    float m = -INFINITY, l = 0.f;
    float o[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) o[i] = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      ptx::mbar_wait(ptx::smem_u32(&bar_s), j & 1);
      ptx::tc_fence_after();
      const int k0 = j * BKV;
      int limit = kv_valid - k0;                     // keys [0, limit) of this tile are visible
      if (p.causal) limit = min(limit, qi - k0 + 1);
      const bool need_mask = limit < BKV;
      // pass 1: row maximum
      float mx = -INFINITY;
#pragma unroll 1
      for (int cc = 0; cc < BKV / 32; ++cc) {
        uint32_t v[32];
        ptx::tc_ld_32x32b_x32(tmem + t_lane + S_COL + cc * 32, v);
        ptx::tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = __uint_as_float(v[i]);
          mx = fmaxf(mx, (need_mask && cc * 32 + i >= limit) ? -INFINITY : s);
        }
      }
      const float m_new = fmaxf(m, mx);
      const float m_off = (m_new == -INFINITY) ? 0.f : m_new * c;
      const float alpha = (m == -INFINITY) ? 0.f : fast_exp2(m * c - m_off);
      // pass 2: probabilities -> bf16 -> swizzled smem (A operand of P*V)
      float rs = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < BKV / 32; ++cc) {
        uint32_t v[32];
        ptx::tc_ld_32x32b_x32(tmem + t_lane + S_COL + cc * 32, v);
        ptx::tc_wait_ld();
        float pr[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float e = fast_exp2(__uint_as_float(v[i]) * c - m_off);
          pr[i] = (need_mask && cc * 32 + i >= limit) ? 0.f : e;
          rs += pr[i];
        }
        const uint32_t half_base = sP + (cc >> 1) * (P_BYTES / 2) + r * 128;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int chunk = (cc & 1) * 4 + q4;       // 16-byte chunk index within the 128-byte row
          const uint32_t addr = half_base + ((chunk ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
                       "r"(pack_bf16x2(pr[8 * q4 + 0], pr[8 * q4 + 1])), "r"(pack_bf16x2(pr[8 * q4 + 2], pr[8 * q4 + 3])),
                       "r"(pack_bf16x2(pr[8 * q4 + 4], pr[8 * q4 + 5])), "r"(pack_bf16x2(pr[8 * q4 + 6], pr[8 * q4 + 7]))
                       : "memory");
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bar_p));
      l = l * alpha + rs;
      m = m_new;
      // accumulate O with this tile's P*V
      ptx::mbar_wait(ptx::smem_u32(&bar_pv), j & 1);
      ptx::tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < HD / 32; ++cc) {
        uint32_t v[32];
        ptx::tc_ld_32x32b_x32(tmem + t_lane + PV_COL + cc * 32, v);
        ptx::tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[cc * 32 + i] = o[cc * 32 + i] * alpha + __uint_as_float(v[i]);
      }
    }
    if (qi < p.Tq) {
      const float inv = 1.f / l;
      bf16* dst = p.o + (static_cast<int64_t>(b) * p.Tq + qi) * p.ldo + h * HD;
#pragma unroll
      for (int q8 = 0; q8 < HD / 8; ++q8) {
        uint4 u;
        u.x = pack_bf16x2(o[8 * q8 + 0] * inv, o[8 * q8 + 1] * inv);
        u.y = pack_bf16x2(o[8 * q8 + 2] * inv, o[8 * q8 + 3] * inv);
        u.z = pack_bf16x2(o[8 * q8 + 4] * inv, o[8 * q8 + 5] * inv);
        u.w = pack_bf16x2(o[8 * q8 + 6] * inv, o[8 * q8 + 7] * inv);
        reinterpret_cast<uint4*>(dst)[q8] = u;
      }
      if (p.lse) p.lse[(static_cast<int64_t>(b) * p.H + h) * p.Tq + qi] = m * c + log2f(l);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem);
  }
}

}  // namespace
}  // namespace oasr

using namespace oasr;

extern "C" int oasr_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                  void* o, int64_t ldo, float* lse, int64_t B, int64_t H, int64_t Tq, int64_t Tkv,
                                  int64_t head_dim, int causal, const int32_t* kv_len, float scale, void* stream) {
  OASR_REQUIRE(head_dim == HD, "attention: head_dim %ld unsupported (every OLMoASR variant uses 64)", (long)head_dim);
  OASR_REQUIRE(B > 0 && H > 0 && Tq > 0 && Tkv > 0, "attention: empty problem");
  OASR_REQUIRE((ldq & 7) == 0 && (ldk & 7) == 0 && (ldv & 7) == 0 && (ldo & 7) == 0, "attention: strides must be multiples of 8");
  OASR_REQUIRE(!causal || Tq == Tkv, "attention: causal needs Tq == Tkv");
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_tmap_2d(&tmQ, q, 2, (uint64_t)(H * HD), (uint64_t)(B * Tq), (uint64_t)ldq * 2, HD, BQ, true))) return rc;
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
  dim3 grid((unsigned)ceil_div(Tq, BQ), (unsigned)H, (unsigned)B);
  attention_fwd_kernel<<<grid, 192, ATT_SMEM, (cudaStream_t)stream>>>(tmQ, tmK, tmV, p);
  OASR_LAUNCH_CHECK();
  return OASR_OK;
}
