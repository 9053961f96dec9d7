// Microbenchmark: tcgen05.ld throughput (TMEM -> registers) per SM for several shapes / warp counts.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld_32x32b_x32(uint32_t a, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
    : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
      "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
    : "r"(a) : "memory");
}
// 16 lanes x 256 bit, x8 repeats = 64 columns, 32 registers per thread
__device__ __forceinline__ void ld_16x256b_x8(uint32_t a, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
    : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
      "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
    : "r"(a) : "memory");
}
// 16 lanes x 128 bit, x16 = 64 columns, 32 registers
__device__ __forceinline__ void ld_16x128b_x16(uint32_t a, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
    : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
      "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
    : "r"(a) : "memory");
}

template <int SHAPE, int PIPE>
__global__ void k(unsigned long long* out, uint32_t* sink, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (PIPE) {   // four loads in flight (128 registers), one wait: what the attention kernels do per tile
      uint32_t r0[32], r1[32], r2[32], r3[32];
      const uint32_t base = tm + ((warp >> 2) & 1) * 128;
      ld_32x32b_x32(base, r0); ld_32x32b_x32(base + 32, r1); ld_32x32b_x32(base + 64, r2); ld_32x32b_x32(base + 96, r3);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= r0[i] ^ r1[i] ^ r2[i] ^ r3[i];
      continue;
    }
    uint32_t r[32];
#pragma unroll
    for (int c = 0; c < 4; ++c) {   // 4 x (32 lanes x 32 columns x 4 B) = 16 KB per warp per iteration
      if (SHAPE == 0) ld_32x32b_x32(tm + ((warp >> 2) & 1) * 128 + c * 32, r);
      else if (SHAPE == 1) { ld_16x256b_x8(tm + ((warp >> 2) & 1) * 128 + c * 32 + 0, r); }
      else { ld_16x128b_x16(tm + ((warp >> 2) & 1) * 128 + c * 32, r); }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= r[i];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

int main() {
  unsigned long long* out; uint32_t* sink;
  cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
  const int iters = 2000;
  const char* names[3] = {"32x32b.x32 (4 KB/instr)", "16x256b.x8 (2 KB/instr, 16 lanes)", "16x128b.x16 (2 KB/instr, 16 lanes)"};
  for (int shape = 0; shape < 3; ++shape)
    for (int warps = 1; warps <= 16; warps *= 2) {
      if (shape == 0) k<0, 0><<<148, warps * 32>>>(out, sink, iters);
      if (shape == 1) k<1, 0><<<148, warps * 32>>>(out, sink, iters);
      if (shape == 2) k<2, 0><<<148, warps * 32>>>(out, sink, iters);
      cudaError_t e = cudaDeviceSynchronize();
      unsigned long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      const double bytes_per_instr = (shape == 0) ? 4096.0 : 2048.0;
      const double bytes = (double)warps * iters * 4 * bytes_per_instr;
      printf("%-36s warps=%2d: %8.1f cycles/iter  -> %6.1f B/clk/SM  (%s)\n", names[shape], warps, (double)h[0] / iters, bytes / (double)h[0],
             cudaGetErrorString(e));
    }
  // warps 1, 2, 4 sit in different lane quarters (sub-partitions); 8 = two per quarter; 16 = four per quarter
  for (int warps = 1; warps <= 16; warps *= 2) {
    k<0, 1><<<148, warps * 32>>>(out, sink, iters);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    const double bytes = (double)warps * iters * 4 * 4096.0;
    printf("%-36s warps=%2d: %8.1f cycles/iter  -> %6.1f B/clk/SM  (%s)\n", "32x32b.x32, 4 in flight", warps, (double)h[0] / iters,
           bytes / (double)h[0], cudaGetErrorString(e));
  }
  return 0;
}
