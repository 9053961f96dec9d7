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

// One warp per row; per-lane dw/db partials live in registers across the rows a warp visits, are
// combined across the block's warps in shared memory and flushed with one atomicAdd per column.
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ w,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     const bf16* __restrict__ dres, bf16* __restrict__ dx, float* __restrict__ dw,
                     float* __restrict__ db, int64_t rows, int d) {
  extern __shared__ float red[];  // [2][d]
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nvec = d >> 3;
  float aw[NV][8], ab[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { aw[i][j] = 0.f; ab[i][j] = 0.f; }
  for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) red[i] = 0.f;
  __syncthreads();

  for (int64_t row = static_cast<int64_t>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); row < rows;
       row += static_cast<int64_t>(gridDim.x) * warps_per_block) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * d);
    const uint4* gr = reinterpret_cast<const uint4*>(dy + row * d);
    const float mu = mean[row], rs = rstd[row];
    float xh[NV][8], g[NV][8];
    float s1 = 0.f, s2 = 0.f;
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
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = (xv[j] - mu) * rs;
          ab[i][j] += gv[j];
          aw[i][j] += gv[j] * xh[i][j];
          g[i][j] = gv[j] * ww[j];
          s1 += g[i][j];
          s2 += g[i][j] * xh[i][j];
        }
      }
    }
    s1 = warp_sum(s1) / d;
    s2 = warp_sum(s2) / d;
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
  // block-level reduction of the parameter gradients
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&red[vi * 8 + j], aw[i][j]);
        atomicAdd(&red[d + vi * 8 + j], ab[i][j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    atomicAdd(dw + i, red[i]);
    atomicAdd(db + i, red[d + i]);
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
  const int64_t cap = static_cast<int64_t>(num_sms()) * 4;  // few blocks => few global atomics, long register reuse
  if (blocks > cap) blocks = cap;
#define OASR_LN_BWD(NVv)                                                                      \
  layernorm_bwd_kernel<NVv><<<(int)blocks, wpb * 32, 2 * d * sizeof(float), st>>>(           \
      (const bf16*)dy, (const bf16*)x, weight, mean, rstd, (const bf16*)dresidual, (bf16*)dx, dweight, dbias, rows, (int)d)
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
