// Batched log-mel spectrogram on the GPU (fp32): reflect-padded framing -> Hann -> 400-point real DFT
// -> power -> 80-band slaney mel filterbank -> log10 / clamp -> per-clip dynamic-range floor -> (x+4)/4.
//
// Replaces whisper.audio.log_mel_spectrogram, which the reference runs per sample on CPU DataLoader
// workers (scripts/training/train_timestamps.py:196-214; eval.py:157-162; olmoasr/transcribe.py:148).
// The floor uses EACH clip's own maximum, i.e. per-sample semantics of the reference datasets.
//
// Kernel 1 (logmel_frames_kernel): one block = 16 consecutive frames of one clip = 8 frame PAIRS.  Two real frames
// a, b ride one complex 400-point FFT (z = a + i b;  A[k] = (Z[k] + conj Z[400-k]) / 2,  B[k] = (Z[k] - conj Z[400-k]) / 2i),
// and the FFT is Cooley-Tukey 400 = 20 x 20 with the 20-point transforms (4 x 5, constant twiddles) held entirely in
// registers: thread (pair, n2) transforms column n2 over n1, multiplies by W_400^(n2 k1) and parks the result in shared
// memory; thread (pair, k1) transforms row k1 over n2.  Two shared-memory round trips and ~8 kflop per frame instead of
// the 160 k MACs of a direct 201-bin DFT (the r01 kernel: FMA-bound at 2 % of its HBM roofline).
// Then power -> sparse mel filterbank -> log10(max(mel, 1e-10)), clip maximum folded with an ordered-int atomicMax.
// Kernel 2 (logmel_finalize_kernel): y = (max(x, clipmax - 8) + 4) / 4, vectorised.
// Algorithmic HBM bytes: 4 B in (2 B for int16) + 80 / 160 x 4 B out per sample = 2.88 MB per 30 s clip (fp32 input).
#include "common.cuh"

namespace oasr {
namespace {

constexpr int N_FFT = 400;
constexpr int HOP = 160;
constexpr int N_BINS = 201;
constexpr int FPB = 16;                         // frames per block
constexpr int PAIRS = FPB / 2;
constexpr int SPAN = (FPB - 1) * HOP + N_FFT;   // 2800 samples
constexpr int YLD = 21;                         // padded row of the 20 x 20 intermediate (conflict-free 8-byte accesses)

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

__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__host__ __device__ __forceinline__ float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }   // -i a
__host__ __device__ __forceinline__ float2 mul_pos_i(float2 a) { return make_float2(-a.y, a.x); }   // +i a

// forward 4-point DFT: X[c] = sum_a v[a] e^(-2 pi i a c / 4)
__host__ __device__ __forceinline__ void dft4(float2& v0, float2& v1, float2& v2, float2& v3) {
  const float2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), a3 = mul_neg_i(csub(v1, v3));
  v0 = cadd(a0, a2); v1 = cadd(a1, a3); v2 = csub(a0, a2); v3 = csub(a1, a3);
}
// forward 5-point DFT
__host__ __device__ __forceinline__ void dft5(float2& v0, float2& v1, float2& v2, float2& v3, float2& v4) {
  constexpr float C1 = 0.30901699437494745f, C2 = -0.80901699437494734f, S1 = 0.95105651629515353f, S2 = 0.58778525229247314f;
  const float2 t1 = cadd(v1, v4), t2 = cadd(v2, v3), t3 = csub(v1, v4), t4 = csub(v2, v3);
  const float2 m1 = make_float2(v0.x + C1 * t1.x + C2 * t2.x, v0.y + C1 * t1.y + C2 * t2.y);
  const float2 m2 = make_float2(v0.x + C2 * t1.x + C1 * t2.x, v0.y + C2 * t1.y + C1 * t2.y);
  const float2 n1 = make_float2(S1 * t3.x + S2 * t4.x, S1 * t3.y + S2 * t4.y);
  const float2 n2 = make_float2(S2 * t3.x - S1 * t4.x, S2 * t3.y - S1 * t4.y);
  v0 = make_float2(v0.x + t1.x + t2.x, v0.y + t1.y + t2.y);
  v1 = cadd(m1, mul_neg_i(n1)); v4 = cadd(m1, mul_pos_i(n1));
  v2 = cadd(m2, mul_neg_i(n2)); v3 = cadd(m2, mul_pos_i(n2));
}
// e^(-2 pi i j / 20), j = 0..12 (all the products b c of the 4 x 5 split)
__host__ __device__ __forceinline__ float2 w20(int j) {
  constexpr float C[13] = {1.f, 0.95105651629515353f, 0.80901699437494745f, 0.58778525229247314f, 0.30901699437494745f, 0.f,
                           -0.30901699437494745f, -0.58778525229247314f, -0.80901699437494745f, -0.95105651629515353f, -1.f,
                           -0.95105651629515353f, -0.80901699437494745f};
  constexpr float S[13] = {0.f, 0.30901699437494745f, 0.58778525229247314f, 0.80901699437494745f, 0.95105651629515353f, 1.f,
                           0.95105651629515353f, 0.80901699437494745f, 0.58778525229247314f, 0.30901699437494745f, 0.f,
                           -0.30901699437494745f, -0.58778525229247314f};
  return make_float2(C[j], -S[j]);
}
// in-register forward 20-point DFT: input x[n], n = 5 a + b; output X[k], k = c + 4 e, returned in natural order in v[]
__host__ __device__ __forceinline__ void dft20(float2 (&v)[20]) {
#pragma unroll
  for (int b = 0; b < 5; ++b) {                    // 4-point transforms over a (stride 5), then the W_20^(b c) twiddles
    dft4(v[b], v[5 + b], v[10 + b], v[15 + b]);    // v[5 c + b] = T[c][b]
#pragma unroll
    for (int c = 1; c < 4; ++c)
      if (b > 0) v[5 * c + b] = cmul(v[5 * c + b], w20(b * c));
  }
  float2 out[20];
#pragma unroll
  for (int c = 0; c < 4; ++c) {                    // 5-point transforms over b
    dft5(v[5 * c], v[5 * c + 1], v[5 * c + 2], v[5 * c + 3], v[5 * c + 4]);
#pragma unroll
    for (int e = 0; e < 5; ++e) out[c + 4 * e] = v[5 * c + e];
  }
#pragma unroll
  for (int k = 0; k < 20; ++k) v[k] = out[k];
}

template <typename InT>
__global__ void __launch_bounds__(256)
logmel_frames_kernel(const InT* __restrict__ wave, const float* __restrict__ window,   // [400]
                     const float* __restrict__ cos_tab, const float* __restrict__ sin_tab,  // [400]: cos / sin(2 pi m / 400)
                     const float* __restrict__ filt,  // [n_mels][201]
                     const int* __restrict__ klo, const int* __restrict__ khi,  // [n_mels]
                     float* __restrict__ out, float* __restrict__ clip_max, int64_t row_stride, int n_valid, int n_frames,
                     int n_frames_valid, int n_mels) {
  // s_xp: raw samples while framing, then the power spectra (the samples are dead by then)
  constexpr int XP = (FPB * (N_BINS + 1) > SPAN) ? FPB * (N_BINS + 1) : SPAN;
  __shared__ float s_xp[XP];
  __shared__ float2 s_y[PAIRS * 20 * YLD];          // pass-1 output [pair][k1][n2], later the spectrum [pair][k] (aliased)
  __shared__ float2 s_tw[N_FFT];                    // e^(-2 pi i m / 400)
  __shared__ float s_win[N_FFT];
  float* s_x = s_xp;
  float (*s_pow)[N_BINS + 1] = reinterpret_cast<float (*)[N_BINS + 1]>(s_xp);
  float2* s_z = s_y;                                // [pair][400] natural order (8 * 400 <= 8 * 20 * 21)

  const int clip = blockIdx.y;
  const int f0 = blockIdx.x * FPB;
  const InT* w = wave + static_cast<int64_t>(clip) * row_stride;
  const int tid = threadIdx.x;

  for (int i = tid; i < N_FFT; i += blockDim.x) { s_tw[i] = make_float2(cos_tab[i], -sin_tab[i]); s_win[i] = window[i]; }
  // torch.stft(center=True, pad_mode="reflect"): padded[i] = x[reflect(i - 200)]
  for (int i = tid; i < SPAN; i += blockDim.x) {
    int src = f0 * HOP + i - N_FFT / 2;
    if (src < 0) src = -src;
    if (src >= n_valid) src = 2 * (n_valid - 1) - src;
    s_x[i] = (src >= 0 && src < n_valid) ? load_sample<InT>(w, src) : 0.f;
  }
  __syncthreads();

  const int pair = tid / 20, col = tid % 20;
  const bool fft_thread = tid < PAIRS * 20;
  float2 v[20];
  if (fft_thread) {   // pass 1: column n2 = col over n1, z[n] = (frame 2 pair)[n] + i (frame 2 pair + 1)[n], windowed
    const float* xa = s_x + (2 * pair) * HOP;
    const float* xb = xa + HOP;
#pragma unroll
    for (int n1 = 0; n1 < 20; ++n1) {
      const int n = 20 * n1 + col;
      const float wn = s_win[n];
      v[n1] = make_float2(xa[n] * wn, xb[n] * wn);
    }
    dft20(v);
    int idx = 0;      // (col * k1) mod 400
#pragma unroll
    for (int k1 = 0; k1 < 20; ++k1) {
      s_y[(pair * 20 + k1) * YLD + col] = cmul(v[k1], s_tw[idx]);
      idx += col;
      if (idx >= N_FFT) idx -= N_FFT;
    }
  }
  __syncthreads();
  if (fft_thread) {   // pass 2: row k1 = col over n2
#pragma unroll
    for (int n2 = 0; n2 < 20; ++n2) v[n2] = s_y[(pair * 20 + col) * YLD + n2];
    dft20(v);
  }
  __syncthreads();    // every read of s_y is done: its storage becomes the natural-order spectrum
  if (fft_thread) {
#pragma unroll
    for (int k2 = 0; k2 < 20; ++k2) s_z[pair * N_FFT + col + 20 * k2] = v[k2];
  }
  __syncthreads();
  // |A[k]|^2 and |B[k]|^2 from Z[k] and Z[400 - k]
  for (int i = tid; i < PAIRS * N_BINS; i += blockDim.x) {
    const int p = i / N_BINS, k = i % N_BINS;
    const float2 z = s_z[p * N_FFT + k];
    const float2 zc = s_z[p * N_FFT + (k == 0 ? 0 : N_FFT - k)];
    const float ar = 0.5f * (z.x + zc.x), ai = 0.5f * (z.y - zc.y);     // A = (Z[k] + conj Z[N-k]) / 2
    const float br = 0.5f * (z.y + zc.y), bi = 0.5f * (zc.x - z.x);     // B = (Z[k] - conj Z[N-k]) / (2 i)
    s_pow[2 * p][k] = ar * ar + ai * ai;
    s_pow[2 * p + 1][k] = br * br + bi * bi;
  }
  __syncthreads();

  float local_max = -INFINITY;
  for (int i = tid; i < n_mels * FPB; i += blockDim.x) {
    const int m = i / FPB, f = i % FPB;
    if (f0 + f >= n_frames) continue;
    float v_out = 0.f;
    if (f0 + f < n_frames_valid) {
      float acc = 0.f;
      const float* fm = filt + m * N_BINS;
      for (int k = klo[m]; k < khi[m]; ++k) acc += __ldg(fm + k) * s_pow[f][k];
      v_out = log10f(fmaxf(acc, 1e-10f));
      local_max = fmaxf(local_max, v_out);
    } else {
      v_out = -INFINITY;   // frames past the recording (row padding): never the maximum, floored by the finalize pass
    }
    out[(static_cast<int64_t>(clip) * n_mels + m) * n_frames + f0 + f] = v_out;
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
                           float* clip_max, int64_t batch, int64_t n_samples, int64_t n_mels, int64_t n_valid, void* stream) {
  if (n_valid <= 0) n_valid = n_samples;
  OASR_REQUIRE(batch > 0 && n_valid >= N_FFT && n_valid <= n_samples, "logmel: need %d <= n_valid (%ld) <= n_samples (%ld)", N_FFT,
               (long)n_valid, (long)n_samples);
  OASR_REQUIRE(n_samples % HOP == 0, "logmel: n_samples must be a multiple of %d", HOP);
  OASR_REQUIRE(n_mels > 0 && n_mels <= 128, "logmel: n_mels out of range");
  const int n_frames = (int)(n_samples / HOP);  // STFT yields n/160 + 1 frames; upstream drops the last one
  const int n_frames_valid = (int)(n_valid / HOP);
  OASR_REQUIRE((n_frames & 3) == 0, "logmel: frame count must be a multiple of 4");
  cudaStream_t st = (cudaStream_t)stream;
  logmel_init_max_kernel<<<(unsigned)ceil_div(batch, 128), 128, 0, st>>>(clip_max, (int)batch);
  dim3 grid((unsigned)ceil_div(n_frames, FPB), (unsigned)batch);
  if (in_is_int16)
    logmel_frames_kernel<int16_t><<<grid, 256, 0, st>>>((const int16_t*)wave, window, cos_tab, sin_tab, filters, klo, khi, out,
                                                        clip_max, n_samples, (int)n_valid, n_frames, n_frames_valid, (int)n_mels);
  else
    logmel_frames_kernel<float><<<grid, 256, 0, st>>>((const float*)wave, window, cos_tab, sin_tab, filters, klo, khi, out,
                                                      clip_max, n_samples, (int)n_valid, n_frames, n_frames_valid, (int)n_mels);
  OASR_LAUNCH_CHECK();
  const int64_t total = batch * n_mels * n_frames;
  int64_t blocks = ceil_div(total / 4, 256);
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  logmel_finalize_kernel<<<(unsigned)blocks, 256, 0, st>>>(out, clip_max, n_mels * n_frames, total);
  OASR_LAUNCH_CHECK();
  return OASR_OK;
}
