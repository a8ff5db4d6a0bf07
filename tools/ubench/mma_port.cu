// Micro-benchmark: does the tensor pipe keep its rate when the attention forward's operand mix runs next to its K / V fills?
// Per "key step" of the forward (two 128-query tiles): 16 SS products Q K^T (N = 128, A and B from shared memory: 8 KB per 64-cycle
// instruction = the whole 128 B/clk port) + 16 TS products P V (A from TMEM, B MN-major from shared memory: 4 KB), and 64 KB of K / V
// arriving in shared memory (cp.async.bulk: the TMA write path).  One CTA per SM; warp 0 issues the MMAs, warp 1 streams the fills, paced
// to `kb_per_step` KB per step (0 = none).  Prints tensor-pipe cycles per step (2048 at the instruction rate).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I x2i_b200/csrc -o tools/ubench/mma_port tools/ubench/mma_port.cu
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
using namespace x2i;

__global__ void __launch_bounds__(64, 1) mma_port_kernel(const uint8_t* __restrict__ src, int steps, int kb_per_step, int ts_only, long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_step, bar_ld, bar_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar_step, 1);
    mbar_init(&bar_ld, 1);
    mbar_init(&bar_done, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 128 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
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
    const uint32_t q_base = smem_u32(smem), k_base = smem_u32(smem + 65536), v_base = smem_u32(smem + 98304);
    const uint64_t qd = make_smem_desc_sw128(q_base, 16, 1024), kd = make_smem_desc_sw128(k_base, 16, 1024);
    const uint64_t vd = make_smem_desc_sw128(v_base, 16384, 1024);
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0), idesc_o = make_idesc_bf16(128, 128, 0, 1);
    long long t0 = clock64();
    for (int s = 0; s < steps; ++s) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // P.V of tile i (TS), then S of tile i (SS) -- the forward's issue order
          umma_ts_w(tmem + 256 + i * 128, tmem + i * 128 + kk * 8, vd + ((kk * 2048) >> 4), idesc_o, 1);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = ((kk >> 2) * 16384 + (kk & 3) * 32) >> 4;
          if (ts_only) umma_ts_w(tmem + i * 128, tmem + 64 + kk * 8, kd + off, idesc_s, kk != 0);
          else umma_ss_w(tmem + i * 128, qd + ((i * 32768) >> 4) + off, kd + off, idesc_s, kk != 0);
        }
      }
      umma_commit_w(&bar_step);
      if (s >= 2) mbar_wait(&bar_step, (s - 2) & 1);  // at most two steps of MMAs in flight
    }
    umma_commit_w(&bar_done);
    mbar_wait(&bar_done, 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x] = t1 - t0;
  } else if (kb_per_step > 0 && (threadIdx.x & 31) == 0) {
    // fills: kb_per_step KB per MMA step in 16 KB bulk copies into a separate 64 KB region (no data dependence with the MMAs)
    const uint8_t* g = src + static_cast<size_t>(blockIdx.x) * (1 << 20);
    uint8_t* dst = smem + 131072;
    uint32_t ph = 0;
    for (int s = 0; s < steps; ++s) {
      mbar_expect_tx(&bar_ld, kb_per_step * 1024);
      for (int c = 0; c < kb_per_step / 16; ++c)
        bulk_load_1d(dst + (c & 3) * 16384, g + ((s * 4 + c) & 63) * 16384, 16384, &bar_ld);
      mbar_wait(&bar_ld, ph);
      ph ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int main() {
  long long* cyc;
  uint8_t* src;
  cudaMalloc(&cyc, 148 * 8);
  cudaMalloc(&src, 148ull << 20);
  cudaMemset(src, 0, 148ull << 20);
  cudaFuncSetAttribute(mma_port_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int steps = 400;
  struct { const char* name; int kb; int ts; } rows[] = {
      {"forward mix (16 SS N=128 + 16 TS N=128 per step), no fills", 0, 0},
      {"forward mix + 64 KB of bulk fills per step (the K / V tiles)", 64, 0},
      {"forward mix + 128 KB of bulk fills per step", 128, 0},
      {"all-TS mix (Q from TMEM) , no fills", 0, 1},
      {"all-TS mix (Q from TMEM) + 64 KB of bulk fills per step", 64, 1},
  };
  for (auto& r : rows) {
    long long h[148];
    for (int rep = 0; rep < 2; ++rep) {
      mma_port_kernel<<<148, 64, 200 * 1024>>>(src, steps, r.kb, r.ts, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", r.name, cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    printf("%-66s cycles per key step = %.0f (instruction rate: 2048)\n", r.name, avg / 148.0 / steps);
  }
  return 0;
}
