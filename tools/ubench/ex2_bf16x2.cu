// Micro-benchmark: MUFU.EX2 on packed bf16x2 / f16x2 operands vs f32 (cycles per warp-instruction, 4 warps per sub-partition).
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 512
template <int MODE>
__global__ void k(unsigned* out, long long* cyc, unsigned seed) {
  unsigned x[8];
  for (int i = 0; i < 8; ++i) x[i] = seed + threadIdx.x * 3 + i;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(*reinterpret_cast<float*>(&x[i])));
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x[i]));
      if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
    }
  }
  long long t1 = clock64();
  unsigned s = 0;
  for (int i = 0; i < 8; ++i) s ^= x[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void acc(float* out) {  // accuracy of the bf16x2 form on [-16, 0]
  const float x = -16.0f * (threadIdx.x + blockIdx.x * blockDim.x) / (256.0f * 64);
  unsigned short xb = __float_as_uint(x) >> 16;  // truncation is fine for the probe: compare against exp2 of the SAME bf16 input
  unsigned in = xb | (unsigned(xb) << 16), r;
  asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(r) : "r"(in));
  const float got = __uint_as_float(r << 16), ref = exp2f(__uint_as_float(unsigned(xb) << 16));
  out[threadIdx.x + blockIdx.x * blockDim.x] = fabsf(got - ref) / ref;
}
int main() {
  unsigned* out; long long* cyc; float* e;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64); cudaMalloc(&e, 256 * 64 * 4);
  const char* names[] = {"f32", "bf16x2", "f16x2"};
  for (int m = 0; m < 3; ++m) {
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) {
      if (m == 0) k<0><<<1, 512>>>(out, cyc, 0x3c003c00u);
      if (m == 1) k<1><<<1, 512>>>(out, cyc, 0x3c003c00u);
      if (m == 2) k<2><<<1, 512>>>(out, cyc, 0x3c003c00u);
      cudaDeviceSynchronize();
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    }
    printf("ex2 %-7s: %.2f cycles per warp-instruction at 4 warps/SMSP\n", names[m], (double)h / ITER / 32);
  }
  acc<<<64, 256>>>(e);
  static float he[256 * 64];
  cudaMemcpy(he, e, sizeof(he), cudaMemcpyDeviceToHost);
  float mx = 0; double av = 0;
  for (int i = 0; i < 256 * 64; ++i) { mx = he[i] > mx ? he[i] : mx; av += he[i]; }
  printf("bf16x2 ex2 relative error vs exp2f(same bf16 input): max %.4g mean %.4g\n", mx, av / (256 * 64));
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
