// Host-side check of the in-register 20-point transform and the 20 x 20 decomposition used by logmel_frames_kernel:
// the very same __host__ __device__ butterflies, driven by loops instead of threads, against a direct O(N^2) DFT in double.
//   nvcc -std=c++17 --expt-relaxed-constexpr -o /tmp/fft400 tools/micro/fft400_host_check.cu olmoasr_b200/csrc/common.cu && /tmp/fft400
#include "../../olmoasr_b200/csrc/logmel.cu"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace oasr;
int main() {
  const int N = 400;
  std::vector<float2> z(N), Y(20 * 20), X(N);
  srand(1);
  for (auto& v : z) v = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
  for (int n2 = 0; n2 < 20; ++n2) {
    float2 v[20];
    for (int n1 = 0; n1 < 20; ++n1) v[n1] = z[20 * n1 + n2];
    dft20(v);
    for (int k1 = 0; k1 < 20; ++k1) {
      const double ang = -2.0 * M_PI * ((n2 * k1) % N) / N;
      Y[k1 * 20 + n2] = cmul(v[k1], make_float2((float)cos(ang), (float)sin(ang)));
    }
  }
  for (int k1 = 0; k1 < 20; ++k1) {
    float2 v[20];
    for (int n2 = 0; n2 < 20; ++n2) v[n2] = Y[k1 * 20 + n2];
    dft20(v);
    for (int k2 = 0; k2 < 20; ++k2) X[k1 + 20 * k2] = v[k2];
  }
  double worst = 0, scale = 0;
  for (int k = 0; k < N; ++k) {
    double re = 0, im = 0;
    for (int n = 0; n < N; ++n) {
      const double ang = -2.0 * M_PI * ((long)k * n % N) / N;
      re += z[n].x * cos(ang) - z[n].y * sin(ang);
      im += z[n].x * sin(ang) + z[n].y * cos(ang);
    }
    worst = fmax(worst, fmax(fabs(re - X[k].x), fabs(im - X[k].y)));
    scale = fmax(scale, fmax(fabs(re), fabs(im)));
  }
  printf("fft400 max abs err %.3e (max |X| %.3f)\n", worst, scale);
  return worst < 2e-5 * scale ? 0 : 1;
}
