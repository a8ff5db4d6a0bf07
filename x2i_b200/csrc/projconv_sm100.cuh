// The projector's layer-mixing convolution on the tensor pipe (sm_100a):  Conv2d(C -> 1, 5 x 5, padding 2) over the [S, H] plane of the
// all-layer MLLM hidden states x[B, C, S, H] (utils/proj.py:66-70 `self.conv(x)`; called at infer/inference_qwenvl.py:179 and
// train/train_qwenvl.py:575), followed by the LayerNorm over H (a second, tiny kernel: the convolution's fp32 plane is 4 MB).
//
// The stencil costs 25 MAC per input element (925 per output at C = 37): on the FP32 pipe that is 0.12 ms for one 512-token prompt,
// 10 % of the HBM roofline of its 78 MB input (proj_mix_ln_kernel in rowwise.cuh, still used for the two mean modes and odd shapes).
// Here the taps along H become a banded-Toeplitz B operand:
//
//   out[s, h0 + j] = sum_c sum_ds  sum_i x[c, s + ds - 2, w0 + i] * T_{c,ds}[i, j],   T_{c,ds}[i, j] = w[c, ds, i - j]  (0 <= i - j <= 4)
//
// with a 64-column window w0 = h0 - 2 that yields up to 60 output columns: per (channel, ds) one 128 x 64 x 64 product, K = C * 5 * 64 in
// total.  A TMA box must start on a 16-byte boundary in the innermost dimension (a 4-byte-aligned start raises an illegal-instruction
// error on B200: measured), so the windows start at 56 t - 8 and a tile yields the 56 columns [56 t - 6, 56 t + 50).
//   * A operand: ONE TMA box of 132 rows x 64 columns per channel (3-D map [H, S, B * C]; rows and columns outside the plane are the
//     convolution's zero padding = TMA's out-of-bounds fill).  The five ds shifts read the SAME box at row offsets 0..4: a descriptor
//     start address of +ds * 128 B (rows of a 128B-swizzled K-major tile are 128 B apart and 8-row groups are contiguous), so the
//     input is fetched from L2 once, not five times.  (Measured: the field must stay 0 here -- the hardware applies the 128-byte swizzle
//     to absolute shared-memory address bits, exactly as the TMA write did; setting base-offset = ds gives wrong products.)
//   * B operand: never in global memory.  A T tile has five non-zero diagonals; the four otherwise idle epilogue warps write the 320 bf16
//     weights of each tile straight into a pre-zeroed, 128B-swizzled shared-memory tile (generic proxy -> fence.proxy.async -> mbarrier).
//     The projector runs in bf16 (inference_qwenvl.py:91, train_qwenvl.py:399), so bf16 taps are the reference's own weights.
//   * accumulator: 64 TMEM columns; epilogue adds the conv bias and stores the 56 valid fp32 columns of each row.
// One CTA per (128-row tile, 56-column tile): 4 x 37 = 148 CTAs for one 512-token prompt of the 3B model -- one wave on 148 SMs.
// The A ring is 6 channels deep (the kernel streams its input once from HBM: ~12 MB must be in flight), the T ring 2.
#pragma once
#include "common.cuh"
#include "rowwise.cuh"

namespace x2i {

constexpr int PC_THREADS = 192;              // warp 0: TMA, warp 1: MMA, warps 2..5: B-tile builders, then epilogue
constexpr int PC_NA = 6;                     // A ring: one 132 x 64 box per input channel
constexpr int PC_NT = 2;                     // T ring: the five Toeplitz tiles of a channel
constexpr int PC_A_ROWS = 132;
constexpr int PC_A_BYTES = 17 * 1024;        // 132 rows x 128 B = 16896, padded to a multiple of 1024
constexpr int PC_T_BYTES = 8192;             // 64 x 64 bf16
constexpr int PC_T_OFF = PC_NA * PC_A_BYTES;
constexpr int PC_BAR_OFF = PC_T_OFF + PC_NT * 5 * PC_T_BYTES;
constexpr int PC_SMEM_BYTES = PC_BAR_OFF + 1024 /*align*/ + 256 /*barriers*/ + 4096 /*weights: C * 25 floats, C <= 40*/;
constexpr int PC_NV = 56;                    // output columns per tile: [56 t - 6, 56 t + 50)
constexpr int PC_HSHIFT = 6;

struct ProjConvParams {
  int B, C, S, H;
  int n_htiles;
  float bias;
  const float* w;   // [C, 5, 5] fp32 (bf16-representable: the parameter is bf16)
  float* out;       // [B * S, H] fp32: conv output + bias
};

__global__ void __launch_bounds__(PC_THREADS, 1) proj_conv_tc_kernel(const __grid_constant__ CUtensorMap tma_x, const ProjConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + PC_BAR_OFF);
  uint64_t* a_empty = a_full + PC_NA;
  uint64_t* t_full = a_empty + PC_NA;
  uint64_t* t_empty = t_full + PC_NT;
  uint64_t* tfull_bar = t_empty + PC_NT;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  float* wsm = reinterpret_cast<float*>(smem + PC_BAR_OFF + 256);

  const int warp = uniform_warp_id();
  const int lane = threadIdx.x & 31;
  const int ht = blockIdx.x % p.n_htiles, rt = blockIdx.x / p.n_htiles;
  const int rows_per_b = p.S / 128;
  const int b = rt / rows_per_b, s0 = (rt - b * rows_per_b) * 128;
  const int h0 = ht * PC_NV - PC_HSHIFT;  // first output column (even); window = [h0 - 2, h0 + 62), h0 - 2 a multiple of 8

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_x);
    for (int i = 0; i < PC_NA; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < PC_NT; ++i) {
      mbar_init(&t_full[i], 4);  // one arrive per builder warp
      mbar_init(&t_empty[i], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  // zero every T tile once: only the five diagonals are rewritten per channel
  for (int i = threadIdx.x; i < PC_NT * 5 * PC_T_BYTES / 16; i += PC_THREADS) reinterpret_cast<uint4*>(smem + PC_T_OFF)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < p.C * 25; i += PC_THREADS) wsm[i] = p.w[i];
  fence_proxy_async();  // the zeros are read by the tensor pipe (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer: one 132 x 64 box per channel
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < p.C; ++c) {
        mbar_wait(&a_empty[stage], phase ^ 1);
        mbar_expect_tx(&a_full[stage], PC_A_ROWS * 128);
        tma_load_3d(smem + stage * PC_A_BYTES, &tma_x, &a_full[stage], h0 - 2, s0 - 2, b * p.C + c);
        if (++stage == PC_NA) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    int sa = 0, st = 0;
    uint32_t pa = 0, pt = 0;
    for (int c = 0; c < p.C; ++c) {
      mbar_wait(&a_full[sa], pa);
      mbar_wait(&t_full[st], pt);
      tc_fence_after();
      const uint32_t a_base = smem_u32(smem + sa * PC_A_BYTES);
      const uint32_t t_base = smem_u32(smem + PC_T_OFF + st * 5 * PC_T_BYTES);
#pragma unroll
      for (int ds = 0; ds < 5; ++ds) {
        // rows ds .. ds + 127 of the box: start address + ds * 128 B (the swizzle follows absolute address bits: base-offset stays 0)
        const uint64_t adesc = make_smem_desc_sw128(a_base + ds * 128, 16, 1024);
        const uint64_t bdesc = make_smem_desc_sw128(t_base + ds * PC_T_BYTES, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss_w(tmem_base, adesc + ((k * 32) >> 4), bdesc + ((k * 32) >> 4), idesc, (c | ds | k) != 0 ? 1u : 0u);
      }
      umma_commit_w(&a_empty[sa]);
      umma_commit_w(&t_empty[st]);
      if (++sa == PC_NA) { sa = 0; pa ^= 1; }
      if (++st == PC_NT) { st = 0; pt ^= 1; }
    }
    umma_commit_w(tfull_bar);
  } else {
    // ---------------------------------------------------------------- warps 2..5: build the Toeplitz tiles, then drain the accumulator
    const int t = threadIdx.x - 64;  // 0..127
    // tile ds, output column j (row of the K-major tile), tap dh: element (j, i = j + dh) = w[c, ds, dh].  A thread owns the same 13 of
    // the 5 x 64 x 5 elements for every channel: their shared-memory offsets and tap indices are computed once.
    constexpr int PER = (5 * 64 * 5 + 127) / 128;
    int eoff[PER], etap[PER];
#pragma unroll
    for (int n = 0; n < PER; ++n) {
      const int e = t + n * 128;
      const int ds = e / 320, r = e - ds * 320;
      const int j = r / 5, dh = r - j * 5;
      const int i = j + dh;
      const bool ok = e < 5 * 64 * 5 && i < 64;
      eoff[n] = ok ? ds * PC_T_BYTES + j * 128 + ((((i >> 3) ^ (j & 7)) << 4) | ((i & 7) << 1)) : -1;
      etap[n] = ds * 5 + dh;
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int c = 0; c < p.C; ++c) {
      mbar_wait(&t_empty[stage], phase ^ 1);
      uint8_t* tb = smem + PC_T_OFF + stage * 5 * PC_T_BYTES;
#pragma unroll
      for (int n = 0; n < PER; ++n)
        if (eoff[n] >= 0) *reinterpret_cast<__nv_bfloat16*>(tb + eoff[n]) = __float2bfloat16(wsm[c * 25 + etap[n]]);
      fence_proxy_async();  // generic-proxy stores -> visible to the tensor pipe's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_full[stage]);
      if (++stage == PC_NT) { stage = 0; phase ^= 1; }
    }
    // epilogue: thread = output row
    const int quad = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const int s = s0 + quad * 32 + lane;
    float* orow = p.out + (static_cast<long long>(b) * p.S + s) * p.H + h0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t v[32];
      tmem_ld32(tmem_base + lane_off + half * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int j = half * 32 + q * 2;
        const int h = h0 + j;  // h0, j, H even: a pair of columns is inside or outside the plane as a whole
        if (j < PC_NV && h >= 0 && h < p.H)
          *reinterpret_cast<float2*>(orow + j) = make_float2(__uint_as_float(v[q * 2]) + p.bias, __uint_as_float(v[q * 2 + 1]) + p.bias);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

// y[r, :] = LayerNorm(v[r, :]) * gamma + beta (bf16) from the fp32 plane; optionally xm = bf16(v) (the pre-LayerNorm mix the projector's
// backward needs).  One 256-thread CTA per row (B * S rows: 512 CTAs for one prompt), MAXQ float4 per thread (H <= 1024 * MAXQ).
template <int MAXQ>
__global__ void __launch_bounds__(256) ln_rows_f32_kernel(const float* __restrict__ v, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float eps, __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ xm, int rows, int H) {
  __shared__ float red[8];
  const int row = blockIdx.x;
  const int nq = H >> 2;  // float4 chunks per row
  const float4* vr = reinterpret_cast<const float4*>(v + static_cast<long long>(row) * H);
  float4 r[MAXQ];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXQ; ++i) {
    const int c = i * 256 + threadIdx.x;
    r[i] = c < nq ? vr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (r[i].x + r[i].y) + (r[i].z + r[i].w);
  }
  const float mean = block_sum_rt(s, red, 8) / H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXQ; ++i) {
    if (i * 256 + threadIdx.x < nq) {
      const float a = r[i].x - mean, b2 = r[i].y - mean, c2 = r[i].z - mean, d = r[i].w - mean;
      q += a * a + b2 * b2 + c2 * c2 + d * d;
    }
  }
  const float rstd = rsqrtf(block_sum_rt(q, red, 8) / H + eps);
  uint2* yr = reinterpret_cast<uint2*>(y + static_cast<long long>(row) * H);
  uint2* xr = xm ? reinterpret_cast<uint2*>(xm + static_cast<long long>(row) * H) : nullptr;
#pragma unroll
  for (int i = 0; i < MAXQ; ++i) {
    const int c = i * 256 + threadIdx.x;
    if (c < nq) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c), be = __ldg(reinterpret_cast<const float4*>(beta) + c);
      uint2 o;
      o.x = pack_bf16x2((r[i].x - mean) * rstd * g.x + be.x, (r[i].y - mean) * rstd * g.y + be.y);
      o.y = pack_bf16x2((r[i].z - mean) * rstd * g.z + be.z, (r[i].w - mean) * rstd * g.w + be.w);
      yr[c] = o;
      if (xr) {
        uint2 mm;
        mm.x = pack_bf16x2(r[i].x, r[i].y);
        mm.y = pack_bf16x2(r[i].z, r[i].w);
        xr[c] = mm;
      }
    }
  }
}

}  // namespace x2i
