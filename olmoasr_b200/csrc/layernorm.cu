// LayerNorm forward / backward, bf16 in/out with fp32 statistics (one warp per row, 16-byte loads,
// shuffle reductions).  HBM-bound: 4 B/element forward, 6-8 B/element backward.
//
// Reference: LayerNorm.forward = F.layer_norm(x.float(), eps=1e-5).type(x.dtype)  (olmoasr/model.py:25-39);
// autograd of the same.  The backward also folds in the residual-branch gradient so that
// dx_total = bf16(d_residual + bf16(dx_ln)) matches autograd's bf16 accumulation order.
#include "common.cuh"

namespace oasr {
namespace {

constexpr int LN_MAX_VEC = 5;  // d <= 5 * 32 * 8 = 1280; kernels are instantiated per ceil(d / 256) to bound registers

template <bool kWriteStats, int NV>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                     bf16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                     int64_t rows, int d, float eps) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nvec = d >> 3;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); row < rows;
       row += static_cast<int64_t>(gridDim.x) * warps_per_block) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * d);
    float v[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        const uint4 u = xr[vi];
        float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y), p2 = unpack_bf16x2(u.z), p3 = unpack_bf16x2(u.w);
        v[i][0] = p0.x; v[i][1] = p0.y; v[i][2] = p1.x; v[i][3] = p1.y;
        v[i][4] = p2.x; v[i][5] = p2.y; v[i][6] = p3.x; v[i][7] = p3.y;
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[i][j];
      }
    }
    const float mean = warp_sum(sum) / d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + 32 * i < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float c = v[i][j] - mean; sq += c * c; }
      }
    const float rstd = rsqrtf(warp_sum(sq) / d + eps);
    if (kWriteStats && lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
    uint4* yr = reinterpret_cast<uint4*>(y + row * d);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w) + 2 * vi), w1 = __ldg(reinterpret_cast<const float4*>(w) + 2 * vi + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(b) + 2 * vi), b1 = __ldg(reinterpret_cast<const float4*>(b) + 2 * vi + 1);
        const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * ww[j] + bb[j];
        uint4 u;
        u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
        u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
        yr[vi] = u;
      }
    }
  }
}

// Backward.  One warp per row computes dx; the parameter gradients (column sums of dy*xhat and dy) go through a
// shared-memory staging tile: each group of 8 rows (one per warp) is written to smem, then thread t folds the 8 rows of
// its own columns into two or three private accumulators.  This keeps the kernel at ~100 registers (the previous
// all-in-registers version needed 214 and ran at one 8-warp block per SM, 23 % of the HBM roofline).
template <int NV>
__global__ void __launch_bounds__(256, 2)
layernorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ w,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     const bf16* __restrict__ dres, bf16* __restrict__ dx, float* __restrict__ dw,
                     float* __restrict__ db, int64_t rows, int d) {
  extern __shared__ __align__(16) float stage[];  // [2][8][d]: dy*xhat, dy
  float* s_a = stage;
  float* s_b = stage + 8 * d;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nvec = d >> 3;
  constexpr int MAXC = (NV * 256 + 255) / 256;  // columns per thread in the fold (d <= NV * 256)
  float acc_w[MAXC], acc_b[MAXC];
#pragma unroll
  for (int i = 0; i < MAXC; ++i) { acc_w[i] = 0.f; acc_b[i] = 0.f; }

  const int64_t n_groups = (rows + 7) / 8;
  for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int64_t row = grp * 8 + warp;
    const bool row_ok = row < rows;
    float xh[NV][8], g[NV][8];
    float s1 = 0.f, s2 = 0.f, rs = 0.f;
    if (row_ok) {
      const uint4* xr = reinterpret_cast<const uint4*>(x + row * d);
      const uint4* gr = reinterpret_cast<const uint4*>(dy + row * d);
      const float mu = mean[row];
      rs = rstd[row];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nvec) {
          const uint4 ux = xr[vi], ug = gr[vi];
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(w) + 2 * vi), w1 = __ldg(reinterpret_cast<const float4*>(w) + 2 * vi + 1);
          const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          float2 a0 = unpack_bf16x2(ux.x), a1 = unpack_bf16x2(ux.y), a2 = unpack_bf16x2(ux.z), a3 = unpack_bf16x2(ux.w);
          float2 c0 = unpack_bf16x2(ug.x), c1 = unpack_bf16x2(ug.y), c2 = unpack_bf16x2(ug.z), c3 = unpack_bf16x2(ug.w);
          const float xv[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
          const float gv[8] = {c0.x, c0.y, c1.x, c1.y, c2.x, c2.y, c3.x, c3.y};
          float pa[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            xh[i][j] = (xv[j] - mu) * rs;
            pa[j] = gv[j] * xh[i][j];
            g[i][j] = gv[j] * ww[j];
            s1 += g[i][j];
            s2 += g[i][j] * xh[i][j];
          }
          float4* da = reinterpret_cast<float4*>(s_a + warp * d + vi * 8);
          float4* dbp = reinterpret_cast<float4*>(s_b + warp * d + vi * 8);
          da[0] = make_float4(pa[0], pa[1], pa[2], pa[3]);
          da[1] = make_float4(pa[4], pa[5], pa[6], pa[7]);
          dbp[0] = make_float4(gv[0], gv[1], gv[2], gv[3]);
          dbp[1] = make_float4(gv[4], gv[5], gv[6], gv[7]);
        }
      }
    } else {
      for (int c = lane; c < d; c += 32) { s_a[warp * d + c] = 0.f; s_b[warp * d + c] = 0.f; }
    }
    s1 = warp_sum(s1) / d;
    s2 = warp_sum(s2) / d;
    if (row_ok) {
      uint4* dxr = reinterpret_cast<uint4*>(dx + row * d);
      const uint4* rr = dres ? reinterpret_cast<const uint4*>(dres + row * d) : nullptr;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nvec) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = rs * (g[i][j] - s1 - xh[i][j] * s2);
          if (rr) {
            const uint4 ur = rr[vi];
            float2 r0 = unpack_bf16x2(ur.x), r1 = unpack_bf16x2(ur.y), r2 = unpack_bf16x2(ur.z), r3 = unpack_bf16x2(ur.w);
            const float rv[8] = {r0.x, r0.y, r1.x, r1.y, r2.x, r2.y, r3.x, r3.y};
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = rv[j] + bf16_round(o[j]);
          }
          uint4 u;
          u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
          u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
          dxr[vi] = u;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = threadIdx.x + i * 256;
      if (c < d) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) { a += s_a[r * d + c]; b += s_b[r * d + c]; }
        acc_w[i] += a; acc_b[i] += b;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < d) { atomicAdd(dw + c, acc_w[i]); atomicAdd(db + c, acc_b[i]); }
  }
}

}  // namespace
}  // namespace oasr

using namespace oasr;

extern "C" int oasr_layernorm_fwd(const void* x, const float* weight, const float* bias, void* y, float* mean,
                                  float* rstd, int64_t rows, int64_t d, float eps, void* stream) {
  OASR_REQUIRE(rows > 0 && d > 0, "layernorm: empty input");
  OASR_REQUIRE((d & 7) == 0 && d <= LN_MAX_VEC * 256, "layernorm: d=%ld must be a multiple of 8 and <= %d", (long)d, LN_MAX_VEC * 256);
  OASR_REQUIRE((mean == nullptr) == (rstd == nullptr), "layernorm: mean and rstd go together");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int wpb = 8;
  int64_t blocks = ceil_div(rows, wpb);
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
#define OASR_LN_FWD(NVv)                                                                                          \
  do {                                                                                                           \
    if (mean)                                                                                                    \
      layernorm_fwd_kernel<true, NVv><<<(int)blocks, wpb * 32, 0, st>>>((const bf16*)x, weight, bias, (bf16*)y, mean, rstd, rows, (int)d, eps); \
    else                                                                                                         \
      layernorm_fwd_kernel<false, NVv><<<(int)blocks, wpb * 32, 0, st>>>((const bf16*)x, weight, bias, (bf16*)y, nullptr, nullptr, rows, (int)d, eps); \
  } while (0)
  switch ((int)ceil_div(d, 256)) {
    case 1: case 2: OASR_LN_FWD(2); break;
    case 3: OASR_LN_FWD(3); break;
    case 4: OASR_LN_FWD(4); break;
    default: OASR_LN_FWD(5); break;
  }
#undef OASR_LN_FWD
  OASR_LAUNCH_CHECK();
  return OASR_OK;
}

extern "C" int oasr_layernorm_bwd(const void* dy, const void* x, const float* weight, const float* mean,
                                  const float* rstd, const void* dresidual, void* dx, float* dweight,
                                  float* dbias, int64_t rows, int64_t d, void* stream) {
  OASR_REQUIRE(rows > 0 && d > 0, "layernorm_bwd: empty input");
  OASR_REQUIRE((d & 7) == 0 && d <= LN_MAX_VEC * 256, "layernorm_bwd: unsupported d=%ld", (long)d);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int wpb = 8;
  int64_t blocks = ceil_div(rows, wpb);
  const int64_t cap = static_cast<int64_t>(num_sms()) * 2;  // 2 resident blocks per SM; few blocks => few global atomics
  if (blocks > cap) blocks = cap;
  const size_t smem = 2 * 8 * d * sizeof(float);
#define OASR_LN_BWD(NVv)                                                                                              \
  do {                                                                                                                 \
    static bool attr_set = false;                                                                                      \
    if (!attr_set) {                                                                                                   \
      OASR_CUDA_OK(cudaFuncSetAttribute(layernorm_bwd_kernel<NVv>, cudaFuncAttributeMaxDynamicSharedMemorySize, NVv * 256 * 64)); \
      attr_set = true;                                                                                                 \
    }                                                                                                                  \
    layernorm_bwd_kernel<NVv><<<(int)blocks, wpb * 32, smem, st>>>(                                                    \
        (const bf16*)dy, (const bf16*)x, weight, mean, rstd, (const bf16*)dresidual, (bf16*)dx, dweight, dbias, rows, (int)d); \
  } while (0)
  switch ((int)ceil_div(d, 256)) {
    case 1: case 2: OASR_LN_BWD(2); break;
    case 3: OASR_LN_BWD(3); break;
    case 4: OASR_LN_BWD(4); break;
    default: OASR_LN_BWD(5); break;
  }
#undef OASR_LN_BWD
  OASR_LAUNCH_CHECK();
  return OASR_OK;
}
