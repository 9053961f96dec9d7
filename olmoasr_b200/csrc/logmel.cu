// Batched log-mel spectrogram on the GPU (fp32): reflect-padded framing -> Hann -> 400-point real DFT
// -> power -> 80-band slaney mel filterbank -> log10 / clamp -> per-clip dynamic-range floor -> (x+4)/4.
//
// Replaces whisper.audio.log_mel_spectrogram, which the reference runs per sample on CPU DataLoader
// workers (scripts/training/train_timestamps.py:196-214; eval.py:157-162; olmoasr/transcribe.py:148).
// The floor uses EACH clip's own maximum, i.e. per-sample semantics of the reference datasets.
//
// Kernel 1 (logmel_frames_kernel): one block = 16 consecutive frames of one clip.  The windowed frame is
// folded (e[n] = xw[n] + xw[400-n], o[n] = xw[n] - xw[400-n]) so the 201-bin DFT costs 199 cos + 199 sin
// MACs per bin instead of 800; thread k owns bin k for all 16 frames (frames broadcast from smem as
// float4).  Writes log10(max(mel, 1e-10)) and folds the clip maximum with an ordered-int atomicMax.
// Kernel 2 (logmel_finalize_kernel): y = (max(x, clipmax - 8) + 4) / 4, vectorised.
#include "common.cuh"

namespace oasr {
namespace {

constexpr int N_FFT = 400;
constexpr int HOP = 160;
constexpr int N_BINS = 201;
constexpr int FPB = 16;  // frames per block
constexpr int SPAN = (FPB - 1) * HOP + N_FFT;  // 2800 samples

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

template <typename InT>
__device__ __forceinline__ float load_sample(const InT* w, int64_t i);
template <>
__device__ __forceinline__ float load_sample<float>(const float* w, int64_t i) { return w[i]; }
template <>
__device__ __forceinline__ float load_sample<int16_t>(const int16_t* w, int64_t i) {
  return static_cast<float>(w[i]) * (1.0f / 32768.0f);  // train_timestamps.py:196
}

template <typename InT>
__global__ void __launch_bounds__(256)
logmel_frames_kernel(const InT* __restrict__ wave, const float* __restrict__ window,   // [400]
                     const float* __restrict__ cos_tab, const float* __restrict__ sin_tab,  // [400]
                     const float* __restrict__ filt,  // [n_mels][201]
                     const int* __restrict__ klo, const int* __restrict__ khi,  // [n_mels]
                     float* __restrict__ out, float* __restrict__ clip_max, int n_samples, int n_frames, int n_mels) {
  // s_pow aliases s_x: the raw samples are dead once the folded frames are built (48 KB static limit)
  constexpr int XP = (FPB * (N_BINS + 1) > SPAN) ? FPB * (N_BINS + 1) : SPAN;
  __shared__ float s_xp[XP];
  __shared__ __align__(16) float s_e[N_BINS][FPB];
  __shared__ __align__(16) float s_o[N_BINS][FPB];
  __shared__ float s_cos[N_FFT], s_sin[N_FFT];
  float* s_x = s_xp;
  float (*s_pow)[N_BINS + 1] = reinterpret_cast<float (*)[N_BINS + 1]>(s_xp);

  const int clip = blockIdx.y;
  const int f0 = blockIdx.x * FPB;
  const InT* w = wave + static_cast<int64_t>(clip) * n_samples;
  const int tid = threadIdx.x;

  for (int i = tid; i < N_FFT; i += blockDim.x) { s_cos[i] = cos_tab[i]; s_sin[i] = sin_tab[i]; }
  // torch.stft(center=True, pad_mode="reflect"): padded[i] = x[reflect(i - 200)]
  for (int i = tid; i < SPAN; i += blockDim.x) {
    int src = f0 * HOP + i - N_FFT / 2;
    if (src < 0) src = -src;
    if (src >= n_samples) src = 2 * (n_samples - 1) - src;
    s_x[i] = (src >= 0 && src < n_samples) ? load_sample<InT>(w, src) : 0.f;
  }
  __syncthreads();
  for (int i = tid; i < N_BINS * FPB; i += blockDim.x) {
    const int n = i / FPB, f = i % FPB;
    const float* fr = s_x + f * HOP;
    float e, o;
    if (n == 0) { e = fr[0] * window[0]; o = 0.f; }
    else if (n == 200) { e = fr[200] * window[200]; o = 0.f; }
    else {
      const float a = fr[n] * window[n], b = fr[N_FFT - n] * window[N_FFT - n];
      e = a + b; o = a - b;
    }
    s_e[n][f] = e; s_o[n][f] = o;
  }
  __syncthreads();

  if (tid < N_BINS) {
    const int k = tid;
    float re[FPB], im[FPB];
    const float sign = (k & 1) ? -1.f : 1.f;  // cos(pi k)
#pragma unroll
    for (int f = 0; f < FPB; ++f) { re[f] = s_e[0][f] + sign * s_e[200][f]; im[f] = 0.f; }
    int idx = 0;
    for (int n = 1; n < 200; ++n) {
      idx += k;
      if (idx >= N_FFT) idx -= N_FFT;
      const float c = s_cos[idx], s = s_sin[idx];
      const float4* e4 = reinterpret_cast<const float4*>(s_e[n]);
      const float4* o4 = reinterpret_cast<const float4*>(s_o[n]);
#pragma unroll
      for (int q = 0; q < FPB / 4; ++q) {
        const float4 ev = e4[q], ov = o4[q];
        re[4 * q + 0] += ev.x * c; re[4 * q + 1] += ev.y * c; re[4 * q + 2] += ev.z * c; re[4 * q + 3] += ev.w * c;
        im[4 * q + 0] += ov.x * s; im[4 * q + 1] += ov.y * s; im[4 * q + 2] += ov.z * s; im[4 * q + 3] += ov.w * s;
      }
    }
#pragma unroll
    for (int f = 0; f < FPB; ++f) s_pow[f][k] = re[f] * re[f] + im[f] * im[f];
  }
  __syncthreads();

  float local_max = -INFINITY;
  for (int i = tid; i < n_mels * FPB; i += blockDim.x) {
    const int m = i / FPB, f = i % FPB;
    if (f0 + f >= n_frames) continue;
    float acc = 0.f;
    const float* fm = filt + m * N_BINS;
    for (int k = klo[m]; k < khi[m]; ++k) acc += __ldg(fm + k) * s_pow[f][k];
    const float v = log10f(fmaxf(acc, 1e-10f));
    out[(static_cast<int64_t>(clip) * n_mels + m) * n_frames + f0 + f] = v;
    local_max = fmaxf(local_max, v);
  }
  local_max = warp_max(local_max);
  if ((tid & 31) == 0 && local_max > -INFINITY) atomic_max_float(clip_max + clip, local_max);
}

__global__ void logmel_init_max_kernel(float* clip_max, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) clip_max[i] = -INFINITY;
}

__global__ void logmel_finalize_kernel(float* __restrict__ x, const float* __restrict__ clip_max, int64_t per_clip, int64_t total) {
  for (int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x * 4) {
    const float floor_v = clip_max[i / per_clip] - 8.0f;
    float4 v = *reinterpret_cast<float4*>(x + i);
    v.x = (fmaxf(v.x, floor_v) + 4.0f) * 0.25f; v.y = (fmaxf(v.y, floor_v) + 4.0f) * 0.25f;
    v.z = (fmaxf(v.z, floor_v) + 4.0f) * 0.25f; v.w = (fmaxf(v.w, floor_v) + 4.0f) * 0.25f;
    *reinterpret_cast<float4*>(x + i) = v;
  }
}

}  // namespace
}  // namespace oasr

using namespace oasr;

extern "C" int oasr_logmel(const void* wave, int in_is_int16, const float* window, const float* cos_tab,
                           const float* sin_tab, const float* filters, const int* klo, const int* khi, float* out,
                           float* clip_max, int64_t batch, int64_t n_samples, int64_t n_mels, void* stream) {
  OASR_REQUIRE(batch > 0 && n_samples >= N_FFT, "logmel: need at least %d samples", N_FFT);
  OASR_REQUIRE(n_samples % HOP == 0, "logmel: n_samples must be a multiple of %d", HOP);
  OASR_REQUIRE(n_mels > 0 && n_mels <= 128, "logmel: n_mels out of range");
  const int n_frames = (int)(n_samples / HOP);  // STFT yields n/160 + 1 frames; upstream drops the last one
  OASR_REQUIRE((n_frames & 3) == 0, "logmel: frame count must be a multiple of 4");
  cudaStream_t st = (cudaStream_t)stream;
  logmel_init_max_kernel<<<(unsigned)ceil_div(batch, 128), 128, 0, st>>>(clip_max, (int)batch);
  dim3 grid((unsigned)ceil_div(n_frames, FPB), (unsigned)batch);
  if (in_is_int16)
    logmel_frames_kernel<int16_t><<<grid, 256, 0, st>>>((const int16_t*)wave, window, cos_tab, sin_tab, filters, klo, khi, out,
                                                        clip_max, (int)n_samples, n_frames, (int)n_mels);
  else
    logmel_frames_kernel<float><<<grid, 256, 0, st>>>((const float*)wave, window, cos_tab, sin_tab, filters, klo, khi, out,
                                                      clip_max, (int)n_samples, n_frames, (int)n_mels);
  OASR_LAUNCH_CHECK();
  const int64_t total = batch * n_mels * n_frames;
  int64_t blocks = ceil_div(total / 4, 256);
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  logmel_finalize_kernel<<<(unsigned)blocks, 256, 0, st>>>(out, clip_max, n_mels * n_frames, total);
  OASR_LAUNCH_CHECK();
  return OASR_OK;
}
