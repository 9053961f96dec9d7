// Fused multi-tensor optimizer step: global grad-norm, clip, (un)scale, AdamW -- two launches for the whole model.
//
// Replaces scaler.unscale_(optimizer) -> clip_grad_norm_(model.parameters(), 1.0) -> scaler.step(optimizer)
// (scripts/training/train_timestamps.py:1508-1522; AdamW lr 1.5e-3 betas (0.9, 0.98) eps 1e-6 wd 0.1, :2110-2113).
// HBM-bound: 4 B/param for the norm pass, 28 B/param for the update (read p, g, m, v; write p, m, v).
#include "common.cuh"

namespace oasr {
namespace {

constexpr int CHUNK = 16384;  // elements per block

struct TensorRec {  // one row of the device-side table (5 x int64)
  float* p;
  const float* g;
  float* m;
  float* v;
  int64_t numel;
};

__global__ void __launch_bounds__(256)
grad_sqnorm_kernel(const TensorRec* __restrict__ recs, const int2* __restrict__ chunks, float* __restrict__ out) {
  const int2 ch = chunks[blockIdx.x];
  const TensorRec r = recs[ch.x];
  const int64_t start = static_cast<int64_t>(ch.y) * CHUNK;
  const int64_t end = min(r.numel, start + CHUNK);
  const float* g = r.g + start;
  const int n = static_cast<int>(end - start);
  float s = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    for (int i = threadIdx.x * 4; i + 4 <= n; i += blockDim.x * 4) {
      const float4 v = *reinterpret_cast<const float4*>(g + i);
      s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int i = (n & ~3) + threadIdx.x; i < n; i += blockDim.x) s += g[i] * g[i];
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += g[i] * g[i];
  }
  s = warp_sum(s);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += part[i];
    atomicAdd(out, t);
  }
}

// norm_sq: sum of squares of the (still scaled) gradients.  Skips the whole update when it is not finite
// (GradScaler semantics) and reports that through found_inf.
__global__ void __launch_bounds__(256)
adamw_kernel(const TensorRec* __restrict__ recs, const int2* __restrict__ chunks, const float* __restrict__ norm_sq,
             float* __restrict__ found_inf, float inv_scale, float max_norm, float lr, float beta1, float beta2,
             float eps, float weight_decay, float bc1, float bc2_sqrt) {
  const float nsq = *norm_sq;
  if (!isfinite(nsq)) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *found_inf = 1.f;
    return;
  }
  const float total_norm = sqrtf(nsq) * inv_scale;
  float coef = inv_scale;
  if (max_norm > 0.f) coef *= fminf(1.f, max_norm / (total_norm + 1e-6f));  // clip_grad_norm_
  const int2 ch = chunks[blockIdx.x];
  const TensorRec r = recs[ch.x];
  const int64_t start = static_cast<int64_t>(ch.y) * CHUNK;
  const int n = static_cast<int>(min(r.numel, start + CHUNK) - start);
  float* p = r.p + start;
  const float* g = r.g + start;
  float* m = r.m + start;
  float* v = r.v + start;
  const float step_size = lr / bc1;
  const float decay = 1.f - lr * weight_decay;
  int i0 = 0;
  if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
        reinterpret_cast<uintptr_t>(v)) & 15) == 0) {
    const int n4 = n & ~3;
    for (int i = threadIdx.x * 4; i < n4; i += blockDim.x * 4) {   // 16-byte accesses: 7 vector transactions per 4 params
      float4 pv = *reinterpret_cast<float4*>(p + i);
      const float4 gv = *reinterpret_cast<const float4*>(g + i);
      float4 mv = *reinterpret_cast<float4*>(m + i);
      float4 vv = *reinterpret_cast<float4*>(v + i);
      float* pp = &pv.x; const float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float gi = gp[j] * coef;
        mp[j] = beta1 * mp[j] + (1.f - beta1) * gi;
        vp[j] = beta2 * vp[j] + (1.f - beta2) * gi * gi;
        pp[j] = pp[j] * decay - step_size * (mp[j] / (sqrtf(vp[j]) / bc2_sqrt + eps));
      }
      *reinterpret_cast<float4*>(p + i) = pv;
      *reinterpret_cast<float4*>(m + i) = mv;
      *reinterpret_cast<float4*>(v + i) = vv;
    }
    i0 = n4;
  }
  for (int i = i0 + threadIdx.x; i < n; i += blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] * decay - step_size * (mi / denom);
    m[i] = mi;
    v[i] = vi;
  }
}

}  // namespace
}  // namespace oasr

using namespace oasr;

extern "C" int oasr_optim_chunk_elems(void) { return CHUNK; }

extern "C" int oasr_grad_sqnorm(const void* recs, const void* chunks, int64_t n_chunks, float* out, void* stream) {
  OASR_REQUIRE(n_chunks > 0, "grad_sqnorm: no chunks");
  OASR_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float), (cudaStream_t)stream));
  grad_sqnorm_kernel<<<(unsigned)n_chunks, 256, 0, (cudaStream_t)stream>>>((const TensorRec*)recs, (const int2*)chunks, out);
  OASR_LAUNCH_CHECK();
  return OASR_OK;
}

extern "C" int oasr_adamw_step(const void* recs, const void* chunks, int64_t n_chunks, const float* norm_sq, float* found_inf,
                               float inv_scale, float max_norm, float lr, float beta1, float beta2, float eps,
                               float weight_decay, int64_t step, void* stream) {
  OASR_REQUIRE(n_chunks > 0 && step >= 1, "adamw: bad arguments");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  adamw_kernel<<<(unsigned)n_chunks, 256, 0, (cudaStream_t)stream>>>((const TensorRec*)recs, (const int2*)chunks, norm_sq, found_inf,
                                                                    inv_scale, max_norm, lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt);
  OASR_LAUNCH_CHECK();
  return OASR_OK;
}
