// Implicit-GEMM 2-D convolution on tcgen05 for sm_100a (NHWC activations, bf16, fp32 accumulation in TMEM) -- the
// ControlNeXt nets of the LightControl editing branch (lightcontrol/lightcontrol_flux.py:575-749: 3x3 / 1x1 / 2x2 convs,
// stride 1 or 2), which the reference runs through cuDNN.
//
//   out[n, y, x, co] = epilogue( sum_{ky,kx,ci} in[n, y*s + ky - pad, x*s + kx - pad, ci] * w[co, ky, kx, ci] )
//
// There is no im2col buffer: the persistent GEMM of gemm_sm100.cuh is kept (TMA -> 128B-swizzled smem ring -> tcgen05.mma
// -> double-buffered TMEM accumulators -> fused epilogue) and only the A-operand load changes.  An M tile is an 8 x 16 patch
// of output pixels of one image; for k-block (tap, 64-channel chunk) the producer issues ONE tensor-map load whose box is
// the patch shifted by the tap -- [64 ch, 16 x, 8 y, 1 n] of the NHWC tensor -- which lands in shared memory exactly as a
// 128-row x 64-column K-major tile; rows that fall outside the image are zero-filled by TMA (= the conv's zero padding).
// Stride-2 convs read a parity view of the same tensor, [2C, W/2, 2, H/2, N] (x parity folded into the channel dim), where a
// tap has a fixed (row parity, column parity), so the box is again dense.  Weights are pre-packed [Cout, KH*KW*Cin].
#pragma once
#include "gemm_sm100.cuh"
#include "rowwise.cuh"

namespace x2i {

struct ConvParams {
  GemmParams g;  // g.M = Nimg * Ho * Wo (output pixels), g.N = Cout, g.K = KH * KW * Cin; g.rows_per_batch = Ho * Wo
  int Nimg, Ho, Wo, Cin, KH, KW, stride, pad;
  int tiles_x, tiles_y;  // 16-wide / (8 * MT)-high output tiles per image
  int imgs_per_group;    // > 0: image n uses weight set n / imgs_per_group (B tensor map is 3-D [K, Cout, groups], bias [groups, Cout])
};

constexpr int CONV_TW = 16, CONV_TH = 8;

// MT = vertically adjacent 8x16 pixel patches per CTA tile.  With MT > 1 ONE tensor-map box [64 ch, 16 x, 8*MT y] lands as MT
// consecutive 128-row K-major A tiles, the weight tile of the k-block is staged once and feeds MT MMAs (MT accumulators of BN
// columns each, all double-buffered: 2 * MT * BN <= 512 TMEM columns).  Narrow convs (Cout 64 / 128) are bound by the shared-
// memory port, not the tensor pipe -- an M=128 x N=128 x K=64 MMA reads 32 KB in 256 cycles, N=64 reads 24 KB in 128 -- so sharing
// the B tile cuts the bytes per MMA from 32 to 24 KB (BN = 128, MT = 2) and from 24 to 18 KB (BN = 64, MT = 4).
template <int BN, int MT>
struct ConvCfg {
  static constexpr int A_BYTES = MT * GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (224 * 1024 / STAGE_BYTES) > 8 ? 8 : (224 * 1024 / STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * MT * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS >= 32, "TMEM budget");
};

template <int BN, int EPI, int MT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
conv2d_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const ConvParams cp) {
  using Cfg = ConvCfg<BN, MT>;
  constexpr int NS = Cfg::STAGES;
  constexpr int TH = CONV_TH * MT;  // tile height in output pixels
  const GemmParams& p = cp.g;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NS * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + NS;
  uint64_t* tfull_bar = empty_bar + NS;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = uniform_warp_id();
  const int lane = threadIdx.x & 31;
  const int tiles_img = cp.tiles_x * cp.tiles_y;  // tiles_y counts TH-high tiles
  const int num_m = cp.Nimg * tiles_img;
  const int num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int cchunks = cp.Cin / GEMM_BK;
  const int num_kb = cp.KH * cp.KW * cchunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int i = 0; i < NS; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile % num_m, n_blk = tile / num_m;
        const int img = m_blk / tiles_img, t = m_blk - img * tiles_img;
        const int y0 = (t / cp.tiles_x) * TH, x0 = (t % cp.tiles_x) * CONV_TW;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / cchunks, cc = kb - tap * cchunks;
          const int ky = tap / cp.KW, kx = tap - ky * cp.KW;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (cp.stride == 1) {
            tma_load_4d(sa, &tma_a, &full_bar[stage], cc * GEMM_BK, x0 + kx - cp.pad, y0 + ky - cp.pad, img);
          } else {  // stride 2: input column 2x + kx - pad = 2 * (x + (vx >> 1)) + (vx & 1)
            const int vx = kx - cp.pad, vy = ky - cp.pad;
            tma_load_5d(sa, &tma_a, &full_bar[stage], (vx & 1) * cp.Cin + cc * GEMM_BK, x0 + (vx >> 1), vy & 1, y0 + (vy >> 1), img);
          }
          if (cp.imgs_per_group > 0)
            tma_load_3d(sb, &tma_b, &full_bar[stage], kb * GEMM_BK, n_blk * BN, img / cp.imgs_per_group);
          else
            tma_load_2d(sb, &tma_b, &full_bar[stage], kb * GEMM_BK, n_blk * BN);
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (whole warp, elected lane issues)
    constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * (MT * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint64_t adesc = make_smem_desc_sw128(a_base, 16, 1024);
        const uint64_t bdesc = make_smem_desc_sw128(a_base + Cfg::A_BYTES, 16, 1024);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            umma_ss_w(d_tmem + mt * BN, adesc + ((mt * GEMM_BM * GEMM_BK * 2 + k * 32) >> 4), bdesc + ((k * 32) >> 4), idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit_w(&empty_bar[stage]);
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
      umma_commit_w(&tfull_bar[as]);
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (2..5): one output pixel per thread and M tile
    const int quad = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile % num_m, n_blk = tile / num_m;
      const int img = m_blk / tiles_img, t = m_blk - img * tiles_img;
      const int r = quad * 32 + lane;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int y = (t / cp.tiles_x) * TH + mt * CONV_TH + r / CONV_TW, x = (t % cp.tiles_x) * CONV_TW + r % CONV_TW;
        const int m = (y < cp.Ho && x < cp.Wo) ? (img * cp.Ho + y) * cp.Wo + x : p.M;  // p.M = "row out of range"
        gemm_epilogue_tile<BN, EPI>(p, tmem_base + as * (MT * BN) + mt * BN + lane_off, m, n_blk * BN,
                                    cp.imgs_per_group > 0 ? static_cast<long long>(img / cp.imgs_per_group) * p.N : 0);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// First ControlNeXt conv: Conv2d(3 -> 64, 3x3, stride 2, pad 1) on the NCHW hint image (lightcontrol_flux.py:594),
// output NHWC.  K = 27 is far too small for the tensor cores: direct form, one output pixel per thread, 64 channels.
// blockIdx.y = weight set g (19 ControlNeXt nets share one hint): w [G][64][27], bias [G][64], out [G, N, H/2, W/2, 64].
__global__ void __launch_bounds__(128) conv_first_kernel(const __nv_bfloat16* __restrict__ x /* [N,3,H,W] */,
                                                         const float* __restrict__ w /* [64][27] (co, ci, ky, kx) */,
                                                         const float* __restrict__ bias, __nv_bfloat16* __restrict__ out /* [N,H/2,W/2,64] */,
                                                         int Nimg, int H, int W) {
  __shared__ float ws[27 * 64 + 64];
  w += static_cast<long long>(blockIdx.y) * 64 * 27;
  bias += blockIdx.y * 64;
  out += static_cast<long long>(blockIdx.y) * Nimg * (H / 2) * (W / 2) * 64;
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) ws[(i % 27) * 64 + i / 27] = w[i];  // [tap][co]
  for (int i = threadIdx.x; i < 64; i += blockDim.x) ws[27 * 64 + i] = bias[i];
  __syncthreads();
  const int Ho = H / 2, Wo = W / 2;
  const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pix >= static_cast<long long>(Nimg) * Ho * Wo) return;
  const int xo = static_cast<int>(pix % Wo), yo = static_cast<int>((pix / Wo) % Ho), n = static_cast<int>(pix / (static_cast<long long>(Wo) * Ho));
  float v[27];
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yi = 2 * yo + ky - 1, xi = 2 * xo + kx - 1;
        v[ci * 9 + ky * 3 + kx] = (yi >= 0 && yi < H && xi >= 0 && xi < W)
                                      ? __bfloat162float(x[((static_cast<long long>(n) * 3 + ci) * H + yi) * W + xi]) : 0.f;
      }
  uint4* o4 = reinterpret_cast<uint4*>(out + pix * 64);
#pragma unroll 1
  for (int c8 = 0; c8 < 8; ++c8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = ws[27 * 64 + c8 * 8 + j];
#pragma unroll
    for (int t = 0; t < 27; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[t] * ws[t * 64 + c8 * 8 + j];
    o4[c8] = pack8(acc);
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm on NHWC bf16 (nn.GroupNorm of ControlNeXt: 2 / 4 / 8 groups): deterministic two-stage statistics + one
// fused apply pass  y = act((x - mean) * rstd * gamma + beta) (+ residual).  act: 0 none, 1 ReLU, 2 SiLU.
// pixels per statistics slab: at least 256, grown so that an image has at most ~1024 slabs (keeps the final reduction short)
__host__ __device__ inline int gn_pix_per_cta(int HW) { const int p = (HW / 1024 + 255) / 256 * 256; return p < 256 ? 256 : p; }
// SUB == 2: groups of 4 channels (the VAE's GroupNorm(32, 128)): a thread's 8-channel chunk covers two groups.
template <int SUB>
__global__ void __launch_bounds__(256) gn_stats_partial_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ part /* [N, nsplit, G, 2] */,
                                                               int HW, int C, int G, int nsplit) {
  __shared__ float red[2 * SUB][256];
  const int tpp = C >> 3;            // threads per pixel (16-byte chunks): 8, 16 or 32
  const int ppi = 256 / tpp;         // pixels per CTA iteration
  const int cg = threadIdx.x % tpp, pl = threadIdx.x / tpp;
  const int split = blockIdx.x, n = blockIdx.y;
  const int ppc = gn_pix_per_cta(HW);
  const int p0 = split * ppc, p1 = min(p0 + ppc, HW);
  float s = 0.f, ss = 0.f, s_hi = 0.f, ss_hi = 0.f;
  const __nv_bfloat16* xb = x + static_cast<long long>(n) * HW * C + cg * 8;
  for (int pp = p0 + pl; pp < p1; pp += 4 * ppi) {  // 4 independent 16-byte loads in flight
    uint4 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = (pp + u * ppi < p1) ? ld_stream(xb + static_cast<long long>(pp + u * ppi) * C) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      unpack8(q[u], f);
      if constexpr (SUB == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { s += f[j]; ss += f[j] * f[j]; }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) { s += f[j]; ss += f[j] * f[j]; s_hi += f[j + 4]; ss_hi += f[j + 4] * f[j + 4]; }
      }
    }
  }
  red[0][threadIdx.x] = s;
  red[1][threadIdx.x] = ss;
  if constexpr (SUB == 2) {
    red[2][threadIdx.x] = s_hi;
    red[3][threadIdx.x] = ss_hi;
  }
  __syncthreads();
  // fold the pixel dimension (stride stays a multiple of tpp, so a thread keeps its channel chunk); fixed order
  for (int stride = 128; stride >= tpp; stride >>= 1) {
    if (static_cast<int>(threadIdx.x) < stride) {
#pragma unroll
      for (int q = 0; q < 2 * SUB; ++q) red[q][threadIdx.x] += red[q][threadIdx.x + stride];
    }
    __syncthreads();
  }
  if constexpr (SUB == 2) {
    if (static_cast<int>(threadIdx.x) < G) {  // group g = channel chunk g / 2, half g & 1
      float* o = part + ((static_cast<long long>(n) * nsplit + split) * G + threadIdx.x) * 2;
      o[0] = red[2 * (threadIdx.x & 1)][threadIdx.x >> 1];
      o[1] = red[2 * (threadIdx.x & 1) + 1][threadIdx.x >> 1];
    }
    return;
  }
  if (static_cast<int>(threadIdx.x) < G) {
    const int tpg = tpp / G;  // channel chunks per group
    float a = 0.f, b = 0.f;
    for (int t = 0; t < tpg; ++t) { a += red[0][threadIdx.x * tpg + t]; b += red[1][threadIdx.x * tpg + t]; }
    float* o = part + ((static_cast<long long>(n) * nsplit + split) * G + threadIdx.x) * 2;
    o[0] = a; o[1] = b;
  }
}
// one 128-thread CTA per (image, group): fp64 accumulation of the per-slab partials in a fixed order (strided per thread, shuffle
// tree per warp, 4 warp sums added in order) -- deterministic, and short enough (<= 8 partials per thread) not to show up
// between the two streaming passes
__global__ void __launch_bounds__(128) gn_stats_final_kernel(const float* __restrict__ part, float2* __restrict__ stats /* [N, G] (mean, rstd) */,
                                                             int G, int nsplit, double count, float eps, int total) {
  __shared__ double red[2][4];
  const int i = blockIdx.x;  // n * G + g
  if (i >= total) return;
  const int n = i / G, g = i - n * G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double a = 0.0, b = 0.0;
  for (int sidx = threadIdx.x; sidx < nsplit; sidx += 128) {
    const float2 v = *reinterpret_cast<const float2*>(part + ((static_cast<long long>(n) * nsplit + sidx) * G + g) * 2);
    a += v.x; b += v.y;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, off);
    b += __shfl_xor_sync(0xffffffffu, b, off);
  }
  if (lane == 0) { red[0][warp] = a; red[1][warp] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = (red[0][0] + red[0][1]) + (red[0][2] + red[0][3]);
    b = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]);
    const double mean = a / count;
    const double var = fmax(b / count - mean * mean, 0.0);
    stats[i] = make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps))));
  }
}
// Apply pass.  grid = (slabs of 256 * U chunks, image); a CTA-wide stride is a multiple of the chunks per pixel, so a thread
// keeps its 8 channels for all its chunks: gamma/beta/statistics are folded ONCE into a per-channel (scale, shift) pair and the
// inner loop is load -> 8 FMA (+ activation) -> store, with 32-bit indexing.
__device__ __forceinline__ float silu_fast(float x) {  // x * sigmoid(x), sigmoid(x) = 0.5 * tanh(0.5 x) + 0.5: one MUFU op
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return x * fmaf(0.5f, t, 0.5f);
}
template <int SUB>
__global__ void __launch_bounds__(256) gn_apply_kernel(const __nv_bfloat16* __restrict__ x, const float2* __restrict__ stats,
                                                       const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                                                       const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ y, int chunks_per_img,
                                                       int C, int G, int act, int imgs_per_set /* 0: one gamma/beta for all images */) {
  constexpr int U = 8;  // 16-byte chunks per thread, a CTA-wide stride apart (coalesced), loads issued in two batches of 4
  const int tpp = C >> 3;
  const int cg = threadIdx.x % tpp;
  const int n = blockIdx.y;
  float sc[8], sh[8];
  {
    float2 st, st_hi;
    if constexpr (SUB == 1) {
      st = stats[n * G + cg / (tpp / G)];
      st_hi = st;
    } else {
      st = stats[n * G + 2 * cg];
      st_hi = stats[n * G + 2 * cg + 1];
    }
    float ga[8], be[8];
    const int pset = imgs_per_set > 0 ? (n / imgs_per_set) * tpp : 0;
    unpack8(__ldg(reinterpret_cast<const uint4*>(gamma) + pset + cg), ga);
    unpack8(__ldg(reinterpret_cast<const uint4*>(beta) + pset + cg), be);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float mean = j < 4 ? st.x : st_hi.x, rstd = j < 4 ? st.y : st_hi.y;
      sc[j] = rstd * ga[j];
      sh[j] = be[j] - mean * sc[j];
    }
  }
  const long long img_off = static_cast<long long>(n) * chunks_per_img;
  const uint4* xin = reinterpret_cast<const uint4*>(x) + img_off;
  const uint4* rin = residual != nullptr ? reinterpret_cast<const uint4*>(residual) + img_off : nullptr;
  uint4* yout = reinterpret_cast<uint4*>(y) + img_off;
  const int base = blockIdx.x * (256 * U) + threadIdx.x;
#pragma unroll
  for (int h = 0; h < U / 4; ++h) {
    uint4 q[4], r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + (h * 4 + u) * 256;
      if (i < chunks_per_img) {
        q[u] = ld_stream(xin + i);
        if (rin != nullptr) r[u] = ld_stream(rin + i);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + (h * 4 + u) * 256;
      if (i >= chunks_per_img) continue;
      float f[8];
      unpack8(q[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j], sh[j]);
      if (act == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
      } else if (act == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = silu_fast(f[j]);
      }
      if (rin != nullptr) {
        float rr[8];
        unpack8(r[u], rr);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += rr[j];
      }
      yout[i] = pack8(f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Nearest-neighbour 2x upsampling on NHWC bf16 (Upsample2D of the VAE decoder, F.interpolate(scale_factor=2, mode="nearest")):
// out[n, y, x, :] = in[n, y / 2, x / 2, :].  One 16-byte chunk per thread, writes coalesced.
__global__ void __launch_bounds__(256) upsample2x_nhwc_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long long total8, int H, int W,
                                                              int c8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c = static_cast<int>(i % c8);
  long long pix = i / c8;
  const int xo = static_cast<int>(pix % (2 * W));
  pix /= 2 * W;
  const int yo = static_cast<int>(pix % (2 * H));
  const long long n = pix / (2 * H);
  out[i] = __ldg(in + ((n * H + (yo >> 1)) * W + (xo >> 1)) * c8 + c);
}

// ------------------------------------------------------------------------------------------------
// Row soft-max of fp32 scores -> bf16 probabilities (the single-head, d = 512 self-attention of the VAE mid block, whose
// scores come from a GEMM with fp32 output): P[r, :] = softmax(S[r, :]).  One CTA per row, the row lives in registers
// (cols <= 256 * 4 * SM_VEC).
constexpr int SM_VEC = 16;  // float4 loads per thread -> up to 16384 columns
// bias (nullable): fp32 [bias_rows, ldb] added to the scores before the soft-max, row r uses bias row r % bias_rows (T5's relative position
// bias is shared by every sequence of the batch).
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ S, long long lds, __nv_bfloat16* __restrict__ P, long long ldp,
                                                           int cols, const float* __restrict__ bias = nullptr, long long ldb = 0, int bias_rows = 1) {
  __shared__ float red[8];
  const float* s = S + static_cast<long long>(blockIdx.x) * lds;
  const float* brow = bias != nullptr ? bias + static_cast<long long>(blockIdx.x % bias_rows) * ldb : nullptr;
  __nv_bfloat16* o = P + static_cast<long long>(blockIdx.x) * ldp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 v[SM_VEC];
  float mx = -INFINITY;
#pragma unroll
  for (int u = 0; u < SM_VEC; ++u) {
    const int c = (u * 256 + threadIdx.x) * 4;
    if (c < cols) {
      v[u] = *reinterpret_cast<const float4*>(s + c);
      if (brow != nullptr) {
        const float4 bb = *reinterpret_cast<const float4*>(brow + c);
        v[u].x += bb.x; v[u].y += bb.y; v[u].z += bb.z; v[u].w += bb.w;
      }
      mx = fmaxf(fmaxf(mx, fmaxf(v[u].x, v[u].y)), fmaxf(v[u].z, v[u].w));
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  const float l2e = 1.4426950408889634f;
#pragma unroll
  for (int u = 0; u < SM_VEC; ++u) {
    const int c = (u * 256 + threadIdx.x) * 4;
    if (c < cols) {
      v[u].x = fast_exp2((v[u].x - mx) * l2e);
      v[u].y = fast_exp2((v[u].y - mx) * l2e);
      v[u].z = fast_exp2((v[u].z - mx) * l2e);
      v[u].w = fast_exp2((v[u].w - mx) * l2e);
      sum += (v[u].x + v[u].y) + (v[u].z + v[u].w);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += red[w];  // fixed order: deterministic
  const float inv = 1.0f / sum;
#pragma unroll
  for (int u = 0; u < SM_VEC; ++u) {
    const int c = (u * 256 + threadIdx.x) * 4;
    if (c < cols) {
      uint2 w2;
      w2.x = pack_bf16x2(v[u].x * inv, v[u].y * inv);
      w2.y = pack_bf16x2(v[u].z * inv, v[u].w * inv);
      *reinterpret_cast<uint2*>(o + c) = w2;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm backward on NHWC bf16 (building block of the LightControl trainer, SURVEY.md 8(f) N4: the ControlNeXt nets are the
// trainable part of lightcontrol/train_lightcontrol.py:672-775).  y = act(x^ * gamma + beta), x^ = (x - mean) * rstd per
// (image, group):   g = dy * act'(x^ gamma + beta);  dgamma_c = sum g x^;  dbeta_c = sum g;
//                   dx = rstd * (g gamma_c - S1/m - x^ S2/m),  S1 = sum_grp g gamma_c,  S2 = sum_grp g gamma_c x^,  m = |group| * HW.
// Both group sums follow from the per-channel sums A_c = sum g, B_c = sum g x^, so the reduction pass only keeps those.
// Deterministic: slab partials -> fixed-order fp64 reduction.  Groups of >= 8 channels (all ControlNeXt GroupNorms).
__device__ __forceinline__ float act_grad(float z, int act) {  // act 0 none, 1 ReLU, 2 SiLU
  if (act == 1) return z > 0.f ? 1.f : 0.f;
  if (act == 2) return dsilu_f(z);
  return 1.f;
}
__global__ void __launch_bounds__(256) gn_bwd_partial_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                             const float2* __restrict__ stats, const __nv_bfloat16* __restrict__ gamma,
                                                             const __nv_bfloat16* __restrict__ beta, float* __restrict__ part /* [N, nsplit, C, 2] */,
                                                             int HW, int C, int G, int nsplit, int act) {
  __shared__ float red[16][256];
  const int tpp = C >> 3, ppi = 256 / tpp;
  const int cg = threadIdx.x % tpp, pl = threadIdx.x / tpp;
  const int split = blockIdx.x, n = blockIdx.y;
  const int ppc = gn_pix_per_cta(HW);
  const int p0 = split * ppc, p1 = min(p0 + ppc, HW);
  const float2 st = stats[n * G + cg / (tpp / G)];
  float ga[8], be[8], sa[8], sb[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(gamma) + cg), ga);
  unpack8(__ldg(reinterpret_cast<const uint4*>(beta) + cg), be);
#pragma unroll
  for (int j = 0; j < 8; ++j) sa[j] = sb[j] = 0.f;
  const long long base = static_cast<long long>(n) * HW * C + cg * 8;
  for (int pp = p0 + pl; pp < p1; pp += ppi) {
    float fx[8], fd[8];
    unpack8(ld_stream(x + base + static_cast<long long>(pp) * C), fx);
    unpack8(ld_stream(dy + base + static_cast<long long>(pp) * C), fd);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (fx[j] - st.x) * st.y;
      const float g = fd[j] * act_grad(fmaf(xh, ga[j], be[j]), act);
      sa[j] += g;
      sb[j] += g * xh;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[j][threadIdx.x] = sa[j]; red[8 + j][threadIdx.x] = sb[j]; }
  __syncthreads();
  for (int stride = 128; stride >= tpp; stride >>= 1) {  // fold the pixel dimension, fixed order
    if (static_cast<int>(threadIdx.x) < stride) {
#pragma unroll
      for (int q = 0; q < 16; ++q) red[q][threadIdx.x] += red[q][threadIdx.x + stride];
    }
    __syncthreads();
  }
  if (static_cast<int>(threadIdx.x) < tpp) {
    float* o = part + ((static_cast<long long>(n) * nsplit + split) * C + threadIdx.x * 8) * 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[2 * j] = red[j][threadIdx.x]; o[2 * j + 1] = red[8 + j][threadIdx.x]; }
  }
}
// one CTA per (image, group): lanes = 32 consecutive channels (a warp's load is one 256-byte row segment of the partials), warps = split
// ranges; per-channel sums over the slabs in fp64 -- each warp adds its range in order, the warps' partials are added in warp order
// (deterministic) -> chan[n, c] = (A_c, B_c); the group sums gsum[n, g] = (S1/m, S2/m) are then added in channel order by one thread.
// (History: the first version walked the channels one after the other with two block barriers each, 42 us per launch on the 2-group
// layers; a warp per channel with lane-strided 8-byte loads still took 27 us on a 1024-slab layer: 4 CTAs of latency-bound loads.)
constexpr int GN_FIN_WARPS = 16;
__global__ void __launch_bounds__(GN_FIN_WARPS * 32) gn_bwd_final_kernel(const float* __restrict__ part, const __nv_bfloat16* __restrict__ gamma,
                                                                        float2* __restrict__ chan, float2* __restrict__ gsum, int C, int G, int nsplit,
                                                                        double m) {
  extern __shared__ double gn_fin[];  // [GN_FIN_WARPS][32][2] warp partials, then [cpg][2] = gamma_c * A_c, gamma_c * B_c
  double* wpart = gn_fin;
  double* gprod = gn_fin + GN_FIN_WARPS * 64;
  const int n = blockIdx.x / G, g = blockIdx.x - n * G;
  const int cpg = C / G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = (nsplit + GN_FIN_WARPS - 1) / GN_FIN_WARPS;
  const int s0 = warp * per, s1 = min(s0 + per, nsplit);
  for (int c0 = 0; c0 < cpg; c0 += 32) {
    const int cc = c0 + lane;
    const int c = g * cpg + cc;
    double a = 0.0, b = 0.0;
    if (cc < cpg) {
      const float2* src = reinterpret_cast<const float2*>(part) + (static_cast<long long>(n) * nsplit + s0) * C + c;
#pragma unroll 8
      for (int sidx = s0; sidx < s1; ++sidx, src += C) {
        const float2 v = *src;
        a += v.x; b += v.y;
      }
    }
    wpart[(warp * 32 + lane) * 2] = a;
    wpart[(warp * 32 + lane) * 2 + 1] = b;
    __syncthreads();
    if (warp == 0 && cc < cpg) {
      a = b = 0.0;
      for (int w = 0; w < GN_FIN_WARPS; ++w) { a += wpart[(w * 32 + lane) * 2]; b += wpart[(w * 32 + lane) * 2 + 1]; }
      chan[static_cast<long long>(n) * C + c] = make_float2(static_cast<float>(a), static_cast<float>(b));
      const double gm = static_cast<double>(__bfloat162float(gamma[c]));
      gprod[2 * cc] = gm * a;
      gprod[2 * cc + 1] = gm * b;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double s1g = 0.0, s2g = 0.0;
    for (int cc = 0; cc < cpg; ++cc) { s1g += gprod[2 * cc]; s2g += gprod[2 * cc + 1]; }
    gsum[blockIdx.x] = make_float2(static_cast<float>(s1g / m), static_cast<float>(s2g / m));
  }
}
// dgamma_c (+)= sum_n B_c[n], dbeta_c (+)= sum_n A_c[n]   (fp32 outputs: parameter gradients)
__global__ void gn_bwd_param_kernel(const float2* __restrict__ chan, float* __restrict__ dgamma, float* __restrict__ dbeta, int Nimg, int C,
                                    int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double a = 0.0, b = 0.0;
  for (int n = 0; n < Nimg; ++n) {
    const float2 v = chan[static_cast<long long>(n) * C + c];
    a += v.x; b += v.y;
  }
  dbeta[c] = (accumulate ? dbeta[c] : 0.f) + static_cast<float>(a);
  dgamma[c] = (accumulate ? dgamma[c] : 0.f) + static_cast<float>(b);
}
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                           const float2* __restrict__ stats, const float2* __restrict__ gsum,
                                                           const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                                                           __nv_bfloat16* __restrict__ dx, int chunks_per_img, int C, int G, int act) {
  constexpr int U = 4;
  const int tpp = C >> 3;
  const int cg = threadIdx.x % tpp;
  const int n = blockIdx.y;
  const float2 st = stats[n * G + cg / (tpp / G)];
  const float2 gs = gsum[n * G + cg / (tpp / G)];
  float ga[8], be[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(gamma) + cg), ga);
  unpack8(__ldg(reinterpret_cast<const uint4*>(beta) + cg), be);
  const long long img_off = static_cast<long long>(n) * chunks_per_img;
  const uint4* xin = reinterpret_cast<const uint4*>(x) + img_off;
  const uint4* din = reinterpret_cast<const uint4*>(dy) + img_off;
  uint4* out = reinterpret_cast<uint4*>(dx) + img_off;
  const int base = blockIdx.x * (256 * U) + threadIdx.x;
  uint4 qx[U], qd[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int i = base + u * 256;
    if (i < chunks_per_img) { qx[u] = ld_stream(xin + i); qd[u] = ld_stream(din + i); }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int i = base + u * 256;
    if (i >= chunks_per_img) continue;
    float fx[8], fd[8];
    unpack8(qx[u], fx);
    unpack8(qd[u], fd);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (fx[j] - st.x) * st.y;
      const float g = fd[j] * act_grad(fmaf(xh, ga[j], be[j]), act);
      fx[j] = st.y * (g * ga[j] - gs.x - xh * gs.y);
    }
    out[i] = pack8(fx);
  }
}

// ------------------------------------------------------------------------------------------------
// im2col for the convolution WEIGHT gradient (first, explicit form: dW = dY^T cols through the MN-major wgrad GEMM; the implicit
// split-K form is the next step): cols[(n, yo, xo), (ky, kx, ci)] = x[n, yo*s + ky - pad, xo*s + kx - pad, ci]
// (zero outside the image).  One 16-byte chunk (8 channels) per thread, writes coalesced.
__global__ void __launch_bounds__(256) im2col_nhwc_kernel(const uint4* __restrict__ x, uint4* __restrict__ cols, long long total8, int H, int W,
                                                          int Ho, int Wo, int c8, int KH, int KW, int stride, int pad) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c = static_cast<int>(i % c8);
  long long r = i / c8;
  const int tap = static_cast<int>(r % (KH * KW));
  r /= KH * KW;
  const int xo = static_cast<int>(r % Wo);
  r /= Wo;
  const int yo = static_cast<int>(r % Ho);
  const long long n = r / Ho;
  const int yi = yo * stride + tap / KW - pad, xi = xo * stride + tap % KW - pad;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (yi >= 0 && yi < H && xi >= 0 && xi < W) v = __ldg(x + ((n * H + yi) * W + xi) * c8 + c);
  cols[i] = v;
}

// dx = dy where y > 0 else 0 (ReLU fused into a conv epilogue: mask by the saved output); out = silu(x) (tiny time-embedding rows)
__global__ void __launch_bounds__(256) relu_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y, uint4* __restrict__ dx, long long total8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  float a[8], b[8];
  unpack8(ld_stream(dy + i), a);
  unpack8(ld_stream(y + i), b);
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = b[j] > 0.f ? a[j] : 0.f;
  dx[i] = pack8(a);
}
__global__ void __launch_bounds__(256) silu_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, long long total8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  float a[8];
  unpack8(__ldg(x + i), a);
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = silu_f(a[j]);
  out[i] = pack8(a);
}

}  // namespace x2i
