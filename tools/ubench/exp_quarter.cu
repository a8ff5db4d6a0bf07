// Micro-benchmark of the soft-max inner loop (exp_half x2 = one 32-score quarter) without TMEM traffic:
// cycles per quarter for 1, 2, 4 warps per SM sub-partition and POLY8 = 0, 1, 2.
#include <cstdio>
#include "../../x2i_b200/csrc/attn_sm100.cuh"
using namespace x2i;
#define ITER 256
template <int POLY8>
__global__ void k(const uint32_t* in, uint32_t* out, long long* cyc, float sc) {
  uint32_t x[32];
  for (int i = 0; i < 32; ++i) x[i] = in[threadIdx.x * 32 + i];
  uint64_t sum2[4] = {0, 0, 0, 0};
  uint32_t pk[16];
  uint32_t acc = 0;
  const uint64_t sc2 = pack_f32x2(sc, sc);
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#ifdef CHAINED
    const float m = __uint_as_float(acc & 0x3f000000u);  // loop-carried: measures the LATENCY of one quarter
#else
    const float m = 0.25f * (it & 3);                    // independent iterations: the THROUGHPUT of back-to-back quarters
#endif
    const uint64_t nm2 = pack_f32x2(-m, -m);
    exp_half<POLY8, 0, 0>(x, sc2, nm2, sum2, pk);
    exp_half<POLY8, 0, 1>(x, sc2, nm2, sum2, pk);
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= pk[i];
  }
  long long t1 = clock64();
  float s0, s1;
  unpack_f32x2(add_f32x2(add_f32x2(sum2[0], sum2[1]), add_f32x2(sum2[2], sum2[3])), s0, s1);
  out[threadIdx.x] = acc + __float_as_uint(s0 + s1);
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  uint32_t *in, *out; long long* cyc;
  cudaMalloc(&in, 1024 * 32 * 4); cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
  cudaMemset(in, 0x3c, 1024 * 32 * 4);
  for (int threads = 128; threads <= 512; threads *= 2)
    for (int m = 0; m < 3; ++m) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (m == 0) k<0><<<1, threads>>>(in, out, cyc, 0.1f);
        if (m == 1) k<1><<<1, threads>>>(in, out, cyc, 0.1f);
        if (m == 2) k<2><<<1, threads>>>(in, out, cyc, 0.1f);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      printf("warps/SMSP=%d POLY8=%d cycles per 32-score quarter per warp-slot = %.1f (MUFU floor %d)\n", threads / 128, m, (double)h / ITER,
             (threads / 128) * (32 - 4 * m) * 8);  // POLY8 counts eighths of the pairs
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
