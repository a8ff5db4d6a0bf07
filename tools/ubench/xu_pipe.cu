// Micro-benchmark: which pipe do MUFU.EX2, F2FP (fp32x2 -> bf16x2), PRMT, FFMA2, FADD2, FMNMX3 occupy on sm_100a?
// One CTA of 256 threads (2 warps per SM sub-partition), each warp runs N independent-chain iterations; cycles per warp-instr.
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#define ITER 512
template <int MODE>
__global__ void k(float* out, long long* cyc, float seed) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = seed + threadIdx.x * 1e-3f + i;
  unsigned acc = 0;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0 || MODE == 2 || MODE == 3 || MODE == 5) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 4) asm volatile("add.f32 %0, %0, 1.0;" : "+f"(x[i]));
    }
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      if (MODE == 1 || MODE == 2) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i + 1]), "f"(x[i])); acc ^= r; }
      if (MODE == 3) { unsigned r; asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(__float_as_uint(x[i])), "r"(__float_as_uint(x[i + 1]))); acc ^= r; }
      if (MODE == 5) { float r; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(x[i]), "f"(x[i + 1]), "f"(seed)); acc ^= __float_as_uint(r); }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += x[i];
  out[threadIdx.x] = s + acc;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
  const char* names[] = {"8xMUFU", "4xF2FP", "8xMUFU+4xF2FP", "8xMUFU+4xPRMT", "8xFADD", "8xMUFU+4xFMNMX3"};
  for (int threads = 128; threads <= 512; threads *= 2)
    for (int m = 0; m < 6; ++m) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        switch (m) {
          case 0: k<0><<<1, threads>>>(out, cyc, 0.5f); break;
          case 1: k<1><<<1, threads>>>(out, cyc, 0.5f); break;
          case 2: k<2><<<1, threads>>>(out, cyc, 0.5f); break;
          case 3: k<3><<<1, threads>>>(out, cyc, 0.5f); break;
          case 4: k<4><<<1, threads>>>(out, cyc, 0.5f); break;
          case 5: k<5><<<1, threads>>>(out, cyc, 0.5f); break;
        }
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      printf("threads=%d (%d warps/SMSP) %-18s cycles/iter = %.1f\n", threads, threads / 128, names[m], (double)h / ITER);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
