// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16 -> fp32, M = 128, K = 16) on sm_100a as a function of N, of where A lives
// (shared memory descriptor / TMEM) and of the B operand's major-ness, with and without TMA-like shared-memory write traffic.
// One CTA per SM; warp 0 issues `reps` rounds of `per_round` back-to-back MMAs, commits, waits; clock64 around the whole thing.
// Answers for the attention backward: is a 128 x 64 x 16 product half the cost of a 128 x 128 x 16 one, and is the SS form bound
// by the shared-memory operand reads (A 4 KB + B N*32 B per MMA against 128 B/clk)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I x2i_b200/csrc -o tools/ubench/mma_rate tools/ubench/mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
using namespace x2i;

struct Cfg {
  int n;        // MMA N
  int a_tmem;   // 1: A from TMEM
  int b_mn;     // 1: B MN-major
  int stores;   // 1: the other 4 warps stream 16-byte shared-memory stores meanwhile (stand-in for TMA fills)
  int mixed;    // 1: the backward's tile mix: 16 x SS N=64 then 8 x TS N=128 MN-major
};

template <int N_, int A_TMEM, int B_MN, int MIXED>
__global__ void __launch_bounds__(160, 1) mma_rate_kernel(Cfg c, int reps, int per_round, long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    stop = 0;
  }
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 65536);
    const uint64_t adesc = make_smem_desc_sw128(a_base, 16, 1024);
    const uint64_t bdesc_k = make_smem_desc_sw128(b_base, 16, 1024);
    const uint64_t bdesc_mn = make_smem_desc_sw128(b_base, 8192, 1024);
    constexpr uint32_t idesc = make_idesc_bf16(128, N_, 0, B_MN);
    constexpr uint32_t idesc64 = make_idesc_bf16(128, 64, 0, 0), idesc128mn = make_idesc_bf16(128, 128, 0, 1);
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int r = 0; r < reps; ++r) {
      if (MIXED == 1) {
        for (int t = 0; t < 4; ++t) {
#pragma unroll
        for (int i = 0; i < 16; ++i) umma_ss_w(tmem + (i >> 3) * 64, adesc + ((i & 3) * 2 + ((i >> 2) & 1) * 1024), bdesc_k + ((i & 3) * 2 + ((i >> 2) & 1) * 512), idesc64, i & 7);
#pragma unroll
        for (int i = 0; i < 8; ++i) umma_ts_w(tmem + 256 + (i & 1) * 128, tmem + (i & 3) * 8, bdesc_mn + (i & 3) * 128, idesc128mn, 1);
        }
      } else {
        for (int o = 0; o < per_round; o += 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const uint32_t d = tmem + 256 + ((i >> 3) & 1) * (N_ > 128 ? 0 : 128);
            const uint64_t bd = B_MN ? bdesc_mn + (i & 3) * 128 : bdesc_k + ((i & 3) * 2 + ((i >> 2) & 1) * (N_ * 8));
            if (A_TMEM) umma_ts_w(d, tmem + (i & 7) * 8, bd, idesc, i & 7);
            else umma_ss_w(d, adesc + ((i & 3) * 2 + ((i >> 2) & 1) * 1024), bd, idesc, i & 7);
          }
        }
      }
      umma_commit_w(&bar);
      mbar_wait(&bar, ph);
      ph ^= 1;
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) {
      cyc[blockIdx.x] = t1 - t0;
      stop = 1;
    }
  } else if (c.stores) {
    uint4* dst = reinterpret_cast<uint4*>(smem + 98304) + (threadIdx.x - 32);
    uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 16; ++i) dst[i * 128] = v;  // 128 threads x 16 B = 2 KB per sweep step, conflict-free
      v.x++;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int N_, int A, int B, int M>
static void go(Cfg c, int reps, int per_round, long long* cyc) {
  cudaFuncSetAttribute(mma_rate_kernel<N_, A, B, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  mma_rate_kernel<N_, A, B, M><<<148, 160, 200 * 1024>>>(c, reps, per_round, cyc);
}
static void launch(Cfg c, int reps, int per_round, long long* cyc) {
  if (c.mixed) return go<64, 0, 0, 1>(c, reps, per_round, cyc);
#define CASE(nn, a, b) if (c.n == nn && c.a_tmem == a && c.b_mn == b) return go<nn, a, b, 0>(c, reps, per_round, cyc);
  CASE(256, 0, 0) CASE(128, 0, 0) CASE(64, 0, 0) CASE(32, 0, 0) CASE(256, 1, 0) CASE(128, 1, 0) CASE(64, 1, 0)
  CASE(32, 1, 0) CASE(16, 1, 0) CASE(16, 0, 0) CASE(128, 0, 1) CASE(128, 1, 1) CASE(64, 1, 1) CASE(64, 0, 1) CASE(256, 0, 1) CASE(256, 1, 1)
  printf("no instantiation\n");
  exit(1);
}
int main() {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  struct Row { const char* name; Cfg c; } rows[] = {
      {"SS  N=256 B K-major", {256, 0, 0, 0, 0}},   {"SS  N=128 B K-major", {128, 0, 0, 0, 0}},
      {"SS  N=64  B K-major", {64, 0, 0, 0, 0}},    {"SS  N=32  B K-major", {32, 0, 0, 0, 0}},
      {"TS  N=256 B K-major", {256, 1, 0, 0, 0}},   {"TS  N=128 B K-major", {128, 1, 0, 0, 0}},
      {"TS  N=64  B K-major", {64, 1, 0, 0, 0}},    {"TS  N=32  B K-major", {32, 1, 0, 0, 0}},   {"TS  N=16  B K-major", {16, 1, 0, 0, 0}},   {"SS  N=16  B K-major", {16, 0, 0, 0, 0}},    {"SS  N=128 B MN-major", {128, 0, 1, 0, 0}},
      {"TS  N=128 B MN-major", {128, 1, 1, 0, 0}}, {"SS  N=64  B MN-major", {64, 0, 1, 0, 0}}, {"SS  N=256 B MN-major", {256, 0, 1, 0, 0}}, {"TS  N=256 B MN-major", {256, 1, 1, 0, 0}},  {"TS  N=64  B MN-major", {64, 1, 1, 0, 0}},
      {"SS  N=128 + smem stores", {128, 0, 0, 1, 0}}, {"SS  N=64  + smem stores", {64, 0, 0, 1, 0}},
      {"TS  N=128 MN + smem stores", {128, 1, 1, 1, 0}},
      {"bwd tile mix (16 SS N=64 + 8 TS N=128 MN)", {64, 0, 0, 0, 1}},
      {"bwd tile mix + smem stores", {64, 0, 0, 1, 1}},
  };
  const int reps = 50, per_round = 128;
  for (auto& r : rows) {
    long long h[148];
    for (int rep = 0; rep < 2; ++rep) {
      launch(r.c, reps, per_round, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", r.name, cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148.0 * reps;
    if (r.c.mixed) printf("%-44s cycles per 128x64 tile = %.0f (floor by N/2 cycles per MMA: 1024)\n", r.name, avg / 4);
    else printf("%-44s cycles per MMA = %.1f (one commit + wait per %d MMAs included)\n", r.name, avg / per_round, per_round);
  }
  return 0;
}
