// HBM-streaming row-wise kernels of the X2I hot path (sm_100a): AdaLN LayerNorm+modulate, the skinny
// modulation / time-embedding GEMV, sinusoidal timestep projection, RoPE table, Euler step, the fused
// attention-distillation KL loss (forward + backward) and the projector's layer-mixing 5x5 conv + LayerNorm.
// All loads/stores are 16-byte vectorised and coalesced; reductions are warp shuffles + one smem hop.
#pragma once
#include "common.cuh"

namespace x2i {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
// same as unpack8, but opaque to common-subexpression elimination: a kernel that keeps a row as packed bf16 and unpacks it once per
// pass must not have the compiler keep the first pass's fp32 copy alive (that is the register footprint it is avoiding)
__device__ __forceinline__ void unpack8_again(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t lo, hi;
    asm volatile("shl.b32 %0, %2, 16;\n\tand.b32 %1, %2, 0xffff0000;" : "=r"(lo), "=r"(hi) : "r"(w[i]));
    f[2 * i] = __uint_as_float(lo);
    f[2 * i + 1] = __uint_as_float(hi);
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
__device__ __forceinline__ uint4 ld_stream(const void* p) {  // read-once data: bypass L1
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// Block-wide sum of NV values per thread; blockDim.x = 32 * NW.  Result broadcast to all threads.
template <int NV, int NW>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red /* [NV * NW] */) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();  // protect `red` against the previous use
  if (l == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[i * NW + w] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NW; ++j) s += red[i * NW + j];  // fixed order -> deterministic
    v[i] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// y[r,:] = LayerNorm(x[r,:]) * (1 + scale[b,:]) + shift[b,:]     (no affine, eps inside sqrt), b = r / rows_per_batch
// One warp per row, row kept in registers (D <= 32*8*MAXC).  AdaLayerNormZero / ZeroSingle / Continuous, norm2.
// AFFINE: y = LN(x) * scale + shift (nn.LayerNorm weight / bias) instead of the AdaLN form LN(x) * (1 + scale) + shift.
// The row stays in registers as PACKED bf16 (48 registers at D = 3072) and is unpacked in each of the three passes (sum, centred
// squares, output): ~64 registers per thread instead of ~128, i.e. 4 CTAs per SM -- all 576 CTAs of a 4608-row call are resident at once
// (one wave instead of 1.95), and the shared scale / shift rows are requested BEFORE the reductions so their L2 latency hides behind them.
// Two row segments in one launch (the image and the text stream of a double block live in separate buffers with their own modulation
// rows): rows [0, rows) belong to segment 0, rows [rows, rows + s1.rows) to `s1`.  s1.rows == 0: a plain single-segment call.
struct LnSeg {
  const __nv_bfloat16* x;
  const __nv_bfloat16* scale;
  const __nv_bfloat16* shift;
  __nv_bfloat16* y;
  long long ldx, ldy, mod_stride;
  int rows, rows_per_batch;
};
template <int MAXC, bool AFFINE = false>
__global__ void __launch_bounds__(256, MAXC <= 12 ? 4 : 2) ln_modulate_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                          const __nv_bfloat16* __restrict__ scale,
                                                          const __nv_bfloat16* __restrict__ shift, long long mod_stride,
                                                          __nv_bfloat16* __restrict__ y, long long ldy, int rows, int D,
                                                          int rows_per_batch, float eps, const LnSeg s1) {
  griddep_launch();  // PDL (common.cuh)
  griddep_wait();
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows + s1.rows) return;
  if (row >= rows) {  // warp-uniform: second segment
    row -= rows;
    x = s1.x; ldx = s1.ldx; scale = s1.scale; shift = s1.shift; mod_stride = s1.mod_stride; y = s1.y; ldy = s1.ldy;
    rows_per_batch = s1.rows_per_batch;
  }
  const int lane = threadIdx.x & 31;
  const int nchunk = D >> 3;  // 16-byte chunks per row
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<long long>(row) * ldx);
  uint4 raw[MAXC];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = i * 32 + lane;
    raw[i] = c < nchunk ? xr[c] : make_uint4(0, 0, 0, 0);
  }
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    float v[8];
    unpack8(raw[i], v);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];  // chunks beyond the row are zero
  }
  const float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    if (i * 32 + lane < nchunk) {
      float v[8];
      unpack8_again(raw[i], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[j] - mean;
        q += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / D + eps);
  const int b = row / rows_per_batch;
  const uint4* sc = reinterpret_cast<const uint4*>(scale + static_cast<long long>(b) * mod_stride);
  const uint4* sh = reinterpret_cast<const uint4*>(shift + static_cast<long long>(b) * mod_stride);
  uint4* yr = reinterpret_cast<uint4*>(y + static_cast<long long>(row) * ldy);
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = i * 32 + lane;
    if (c < nchunk) {
      float v[8], a[8], h[8], o[8];
      unpack8_again(raw[i], v);
      unpack8(__ldg(sc + c), a);
      unpack8(__ldg(sh + c), h);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[j] - mean) * rstd * (AFFINE ? a[j] : 1.0f + a[j]) + h[j];
      yr[c] = pack8(o);
    }
    if ((i & 1) == 1) asm volatile("" ::: "memory");  // keep at most two chunks' scale / shift loads in flight (register budget)
  }
}

// ------------------------------------------------------------------------------------------------
// out[b, n] (+)= bias[n] + sum_k act(x[b,k]) * W[n,k]      B <= 8 rows, weights streamed once from HBM.
// One warp per output column group; activations staged once per CTA in shared memory as fp32.
// act_in: 0 identity, 1 SiLU.   Used for every AdaLN modulation linear of a step in ONE launch (weights are
// concatenated at load) and for the timestep / guidance / pooled-text embedding MLPs.
template <int MAXB>
__global__ void __launch_bounds__(256) skinny_linear_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                            const __nv_bfloat16* __restrict__ W, long long ldw,
                                                            const __nv_bfloat16* __restrict__ bias,
                                                            __nv_bfloat16* __restrict__ out, long long ldo, int B, int N,
                                                            int K, int act_in, int accumulate) {
  extern __shared__ float xs[];  // [B][K]
  for (int i = threadIdx.x; i < B * K; i += blockDim.x) {
    const int b = i / K, k = i - b * K;
    float t = __bfloat162float(x[static_cast<long long>(b) * ldx + k]);
    if (act_in == 1) t = t / (1.0f + __expf(-t));
    xs[i] = t;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int nchunk = K >> 3;
  for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += warps) {
    const uint4* wr = reinterpret_cast<const uint4*>(W + static_cast<long long>(n) * ldw);
    float acc[MAXB];
#pragma unroll
    for (int b = 0; b < MAXB; ++b) acc[b] = 0.f;
    for (int c0 = 0; c0 < nchunk; c0 += 128) {  // 4 x 16B loads in flight per lane
      uint4 u[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + i * 32 + lane;
        u[i] = (c < nchunk) ? ld_stream(wr + c) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + i * 32 + lane;
        if (c < nchunk) {
          float w[8];
          unpack8(u[i], w);
#pragma unroll
          for (int b = 0; b < MAXB; ++b) {
            if (b < B) {
              const float4* xp = reinterpret_cast<const float4*>(xs + b * K + c * 8);
              const float4 x0 = xp[0], x1 = xp[1];
              acc[b] += w[0] * x0.x + w[1] * x0.y + w[2] * x0.z + w[3] * x0.w + w[4] * x1.x + w[5] * x1.y +
                        w[6] * x1.z + w[7] * x1.w;
            }
          }
        }
      }
    }
#pragma unroll
    for (int b = 0; b < MAXB; ++b) {
      if (b < B) {
        const float s = warp_sum(acc[b]);
        if (lane == 0) {
          float r = s + (bias != nullptr ? __bfloat162float(bias[n]) : 0.f);
          __nv_bfloat16* o = out + static_cast<long long>(b) * ldo + n;
          if (accumulate) r = __bfloat162float(__float2bfloat16(r)) + __bfloat162float(*o);
          *o = __float2bfloat16(r);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0): out[b] = [cos(t f_i) | sin(t f_i)], f_i = exp(-ln(1e4) i/half)
__global__ void sinusoid_kernel(const float* __restrict__ t, __nv_bfloat16* __restrict__ out, int B, int dim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim >> 1;
  if (i >= B * half) return;
  const int b = i / half, j = i - b * half;
  const float f = expf(-9.210340371976184f * static_cast<float>(j) / static_cast<float>(half));
  const float a = t[b] * f;
  out[static_cast<long long>(b) * dim + j] = __float2bfloat16(cosf(a));
  out[static_cast<long long>(b) * dim + half + j] = __float2bfloat16(sinf(a));
}

// ------------------------------------------------------------------------------------------------
// FluxPosEmbed: ids[L,3] (fp32 integer-valued) -> cos[L,128], sin[L,128] fp32 (pair-repeated) and the compact
// interleaved table rope[L,64] = (cos, sin) consumed by the QKV epilogue.  Angles formed in float64 like the reference.
__global__ void rope_table_kernel(const float* __restrict__ ids, int L, int a0, int a1, int a2, double theta,
                                  float* __restrict__ cos_out, float* __restrict__ sin_out, float2* __restrict__ rope) {
  const int half = (a0 + a1 + a2) >> 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L * half) return;
  const int l = i / half, p = i - l * half;  // p = pair index
  int axis, j, d;
  if (2 * p < a0) { axis = 0; j = p; d = a0; }
  else if (2 * p < a0 + a1) { axis = 1; j = p - a0 / 2; d = a1; }
  else { axis = 2; j = p - (a0 + a1) / 2; d = a2; }
  const double freq = 1.0 / pow(theta, static_cast<double>(2 * j) / static_cast<double>(d));
  const double ang = static_cast<double>(ids[l * 3 + axis]) * freq;
  const float c = static_cast<float>(cos(ang)), s = static_cast<float>(sin(ang));
  const long long o = static_cast<long long>(l) * (2 * half) + 2 * p;
  if (cos_out) { cos_out[o] = c; cos_out[o + 1] = c; }
  if (sin_out) { sin_out[o] = s; sin_out[o + 1] = s; }
  if (rope) rope[static_cast<long long>(l) * half + p] = make_float2(c, s);
}

// ------------------------------------------------------------------------------------------------
// FlowMatchEulerDiscreteScheduler.step: x <- bf16(float(x) + (sigma_next - sigma) * v)
__global__ void euler_step_kernel(__nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ v, float dsigma,
                                  long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float a[8], b[8];
  unpack8(reinterpret_cast<const uint4*>(x)[i], a);
  unpack8(reinterpret_cast<const uint4*>(v)[i], b);
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] += dsigma * b[j];
  reinterpret_cast<uint4*>(x)[i] = pack8(a);
}

// ------------------------------------------------------------------------------------------------
// Resampler key input (minicpm/resampler.py:156-177): out[b,l,:] = x[b,l,:] + pos[(l / w_b), (l % w_b), :] for
// l < h_b * w_b (the 2-D sincos table sliced to the image's patch grid and flattened), else x[b,l,:] (zero padding).
// tgt_sizes: int32 [B,2] = (h_b, w_b); pos: bf16 [max_h, max_w, D].  Index work is exact.
__global__ void add_pos2d_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ pos,
                                 const int* __restrict__ tgt_sizes, __nv_bfloat16* __restrict__ out, int B, int L, int D,
                                 int max_w) {
  const int nchunk = D >> 3;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * L * nchunk) return;
  const int c = static_cast<int>(i % nchunk);
  const long long row = i / nchunk;
  const int b = static_cast<int>(row / L), l = static_cast<int>(row - static_cast<long long>(b) * L);
  const int th = tgt_sizes[2 * b], tw = tgt_sizes[2 * b + 1];
  float a[8];
  unpack8(reinterpret_cast<const uint4*>(x + row * D)[c], a);
  if (l < th * tw) {
    const int ph = l / tw, pw = l - ph * tw;
    float q[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(pos + (static_cast<long long>(ph) * max_w + pw) * D) + c), q);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += q[j];
  }
  reinterpret_cast<uint4*>(out + row * D)[c] = pack8(a);
}

// ------------------------------------------------------------------------------------------------
// x[r,:] += gate[r / rows_per_batch, :] * y[r,:]   (un-fused form of the AdaLN gate + residual, used only when a
// plug-in attention processor returns the un-gated tensor; the default path fuses this into the GEMM epilogue)
__global__ void gate_residual_kernel(__nv_bfloat16* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ y,
                                     long long ldy, const __nv_bfloat16* __restrict__ gate, long long gate_stride,
                                     int rows, int D, int rows_per_batch) {
  const int nchunk = D >> 3;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(rows) * nchunk) return;
  const int r = static_cast<int>(i / nchunk), c = static_cast<int>(i - static_cast<long long>(r) * nchunk);
  float a[8], b[8], g[8];
  uint4* xp = reinterpret_cast<uint4*>(x + r * ldx) + c;
  unpack8(*xp, a);
  unpack8(reinterpret_cast<const uint4*>(y + r * ldy)[c], b);
  unpack8(__ldg(reinterpret_cast<const uint4*>(gate + (r / rows_per_batch) * gate_stride) + c), g);
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] += g[j] * b[j];
  *xp = pack8(a);
}

// ------------------------------------------------------------------------------------------------
// Attention-distillation loss (train_qwenvl.py:58-61, :601-620), one 128-thread CTA per row of D elements.
//   z = (x - mean) / (1e-7 + std_unbiased);  p = softmax(z / T);  row_kl = sum_j ps_j (log ps_j - log pt_j)
// |z|/T <= sqrt(D-1)/T  (= 18.5 for D=3072, T=3) so exp() needs no max subtraction in fp32.
// Forward writes row_kl[row]; a second deterministic kernel sums rows per layer.
// Backward recomputes the row statistics and writes d loss / d student in bf16.
constexpr int KD_THREADS = 128;
constexpr int KD_STAGES = 4;  // rows in flight per CTA (shared-memory ring of the persistent KD kernel)
__device__ __forceinline__ float sum_f32x2(uint64_t v) {
  float a, b;
  unpack_f32x2(v, a, b);
  return a + b;
}
// All element-wise math runs on packed fp32x2 registers (FADD2 / FMUL2 / FFMA2): ~10 instructions per (teacher,
// student) element pair instead of ~16, which is what keeps this kernel HBM-bound rather than issue-bound (2 MUFU.EX2 per
// pair remain the next limiter: 384 clk per 3072-wide row per SM against 530 clk of HBM time).
template <int MAXC, bool BWD>
__global__ void __launch_bounds__(KD_THREADS, 4) kd_row_kernel(const __nv_bfloat16* __restrict__ teacher,
                                                            const __nv_bfloat16* __restrict__ student, int D,
                                                            float inv_T, float* __restrict__ row_kl,
                                                            const float* __restrict__ row_scale /* BWD: per-row upstream */,
                                                            __nv_bfloat16* __restrict__ grad, long long rows) {
  // Persistent over rows with a shared-memory ring filled by bulk async copies (cp.async.bulk, the 1-D TMA path): thread 0 keeps
  // KD_STAGES rows (teacher + student, 2 x D x 2 B each) in flight per CTA, completion on an mbarrier per stage; the four warps copy a
  // landed row into registers, and the stage is refilled right after the first block-wide reduction (whose barrier orders every
  // thread's reads before the refill).  HBM requests no longer depend on registers or on the three dependent reduction phases of a row:
  // one-row-per-CTA ran at 2.8 TB/s on a 4608-row layer, register prefetch of one row at 3.0 TB/s (profiles/r02_rowwise_bandwidth.md).
  extern __shared__ __align__(16) uint8_t kd_smem[];
  __shared__ float red[3 * (KD_THREADS / 32)];
  __shared__ __align__(8) uint64_t full_bar[KD_STAGES];
  const int nchunk = D >> 3;
  const uint32_t row_bytes = static_cast<uint32_t>(D) * 2u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < KD_STAGES; ++i) mbar_init(&full_bar[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](long long r, int stage) {  // thread 0 only
    uint8_t* dst = kd_smem + static_cast<size_t>(stage) * 2 * row_bytes;
    mbar_expect_tx(&full_bar[stage], 2 * row_bytes);
    bulk_load_1d(dst, teacher + r * D, row_bytes, &full_bar[stage]);
    bulk_load_1d(dst + row_bytes, student + r * D, row_bytes, &full_bar[stage]);
  };
  if (threadIdx.x == 0) {
#pragma unroll 1
    for (int i = 0; i < KD_STAGES; ++i) {
      const long long r = static_cast<long long>(blockIdx.x) + static_cast<long long>(i) * gridDim.x;
      if (r < rows) issue(r, i);
    }
  }
  int it = 0;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x, ++it) {
  const int stage = it % KD_STAGES;
  uint64_t t[MAXC][4], s[MAXC][4];  // fp32x2 pairs
  uint4 tq[MAXC], sq[MAXC];
  mbar_wait(&full_bar[stage], (it / KD_STAGES) & 1);
  {
    const uint4* st_t = reinterpret_cast<const uint4*>(kd_smem + static_cast<size_t>(stage) * 2 * row_bytes);
    const uint4* st_s = reinterpret_cast<const uint4*>(kd_smem + static_cast<size_t>(stage) * 2 * row_bytes + row_bytes);
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = i * KD_THREADS + threadIdx.x;
      if (c < nchunk) {
        tq[i] = st_t[c];
        sq[i] = st_s[c];
      }
    }
  }
  uint64_t at = 0ull, as = 0ull;
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    if (i * KD_THREADS + threadIdx.x < nchunk) {
      t[i][0] = bf16x2_to_f32x2(tq[i].x); t[i][1] = bf16x2_to_f32x2(tq[i].y);
      t[i][2] = bf16x2_to_f32x2(tq[i].z); t[i][3] = bf16x2_to_f32x2(tq[i].w);
      s[i][0] = bf16x2_to_f32x2(sq[i].x); s[i][1] = bf16x2_to_f32x2(sq[i].y);
      s[i][2] = bf16x2_to_f32x2(sq[i].z); s[i][3] = bf16x2_to_f32x2(sq[i].w);
#pragma unroll
      for (int j = 0; j < 4; ++j) { at = add_f32x2(at, t[i][j]); as = add_f32x2(as, s[i][j]); }
    }
  }
  float r2[2] = {sum_f32x2(at), sum_f32x2(as)};
  block_sum<2, KD_THREADS / 32>(r2, red);
  if (threadIdx.x == 0) {  // every thread has read this stage (it passed the barriers of block_sum): refill it
    const long long r = row + static_cast<long long>(KD_STAGES) * gridDim.x;
    if (r < rows) issue(r, stage);
  }
  const float mt = r2[0] / D, ms = r2[1] / D;
  const uint64_t nmt2 = pack_f32x2(-mt, -mt), nms2 = pack_f32x2(-ms, -ms);
  at = as = 0ull;
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    if (i * KD_THREADS + threadIdx.x < nchunk) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        t[i][j] = add_f32x2(t[i][j], nmt2);
        s[i][j] = add_f32x2(s[i][j], nms2);
        at = fma_f32x2(t[i][j], t[i][j], at);
        as = fma_f32x2(s[i][j], s[i][j], as);
      }
    }
  }
  r2[0] = sum_f32x2(at); r2[1] = sum_f32x2(as);
  block_sum<2, KD_THREADS / 32>(r2, red);
  const float sd_t = sqrtf(r2[0] / (D - 1)), sd_s = sqrtf(r2[1] / (D - 1));
  const float kt = inv_T / (1e-7f + sd_t), ks = inv_T / (1e-7f + sd_s);
  // logits u = z/T, evaluated in the exp2 domain: e = 2^(u log2e)
  constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  const uint64_t kt2 = pack_f32x2(kt * LOG2E, kt * LOG2E), ks2 = pack_f32x2(ks * LOG2E, ks * LOG2E);
  const uint64_t nkt2 = pack_f32x2(-kt * LOG2E, -kt * LOG2E);
  uint64_t aet = 0ull, aes = 0ull, aeg = 0ull;  // sum e_t, sum e_s, sum e_s * gap   (gap = (u_s - u_t) log2e)
  uint64_t es[BWD ? MAXC : 1][4];                // BWD keeps e_s (saves a second MUFU pass)
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    if (i * KD_THREADS + threadIdx.x < nchunk) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t us2 = mul_f32x2(s[i][j], ks2);
        const uint64_t ut2 = mul_f32x2(t[i][j], kt2);
        const uint64_t gap2 = fma_f32x2(t[i][j], nkt2, us2);
        float a0, a1, b0, b1;
        unpack_f32x2(us2, a0, a1);
        unpack_f32x2(ut2, b0, b1);
        const uint64_t es2 = pack_f32x2(fast_exp2(a0), fast_exp2(a1));
        const uint64_t et2 = pack_f32x2(fast_exp2(b0), fast_exp2(b1));
        aet = add_f32x2(aet, et2);
        aes = add_f32x2(aes, es2);
        aeg = fma_f32x2(es2, gap2, aeg);
        t[i][j] = gap2;  // keep the logit gap (log2 units); s stays centred
        if constexpr (BWD) es[i][j] = es2;
      }
    }
  }
  float r3[3] = {sum_f32x2(aet), sum_f32x2(aes), sum_f32x2(aeg)};
  block_sum<3, KD_THREADS / 32>(r3, red);
  const float lse_t = __logf(r3[0]), lse_s = __logf(r3[1]);
  const float kl = LN2 * r3[2] / r3[1] - lse_s + lse_t;
  if constexpr (!BWD) {
    if (threadIdx.x == 0) row_kl[row] = kl;
  } else {
    // g_j = dKL/du_j = ps_j (a_j - kl), a_j = (us_j - ut_j) - lse_s + lse_t ; then back through normalize().
    const float inv_zs = 1.0f / r3[1];
    const float shift = lse_t - lse_s - kl;
    const uint64_t izs2 = pack_f32x2(inv_zs, inv_zs), ln2_2 = pack_f32x2(LN2, LN2), shift2 = pack_f32x2(shift, shift);
    uint64_t ag = 0ull, ags = 0ull;  // sum g, sum g * (s - mean)
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      if (i * KD_THREADS + threadIdx.x < nchunk) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t ps2 = mul_f32x2(es[i][j], izs2);
          const uint64_t g2 = mul_f32x2(ps2, fma_f32x2(t[i][j], ln2_2, shift2));
          t[i][j] = g2;
          ag = add_f32x2(ag, g2);
          ags = fma_f32x2(g2, s[i][j], ags);
        }
      }
    }
    float r[2] = {sum_f32x2(ag), sum_f32x2(ags)};
    block_sum<2, KD_THREADS / 32>(r, red);
    const float up = row_scale[row];
    const float gmean = r[0] / D;
    const bool dead = (up == 0.f);  // layer skipped by the inf/nan guard: exactly zero gradient, never 0 * NaN
    // du_j/ds_k = ks (delta_jk - 1/D) - (s_j-mean)(s_k-mean) * inv_T / ((eps+sd)^2 (D-1) sd)
    const float c2 = r[1] * inv_T / ((1e-7f + sd_s) * (1e-7f + sd_s) * (D - 1) * fmaxf(sd_s, 1e-30f));
    const uint64_t ngm2 = pack_f32x2(-gmean, -gmean), a2 = pack_f32x2(up * ks, up * ks), b2 = pack_f32x2(-up * c2, -up * c2);
    uint4* gr = reinterpret_cast<uint4*>(grad + row * D);
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = i * KD_THREADS + threadIdx.x;
      if (c < nchunk) {
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t v2 = fma_f32x2(add_f32x2(t[i][j], ngm2), a2, mul_f32x2(s[i][j], b2));
          float v0, v1;
          unpack_f32x2(v2, v0, v1);
          o[j] = dead ? 0u : pack_bf16x2(v0, v1);
        }
        gr[c] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  }  // row loop
}

// Deterministic reduction, stage 1: seg_sum[s] = sum of row_kl over the segment's rows.  A segment is a contiguous
// run of rows belonging to one (layer, batch element) -- e.g. the reference's stacked [B, n_layers, L, D] tensors have
// B * n_layers segments (train_qwenvl.py:590-592).
__global__ void __launch_bounds__(256) kd_segment_reduce_kernel(const float* __restrict__ row_kl,
                                                                const long long* __restrict__ seg_row_start /* [n+1] */,
                                                                double* __restrict__ seg_sum) {
  __shared__ double red[256];
  const int sidx = blockIdx.x;
  const long long r0 = seg_row_start[sidx], r1 = seg_row_start[sidx + 1];
  double acc = 0.0;
  for (long long r = r0 + threadIdx.x; r < r1; r += 256) acc += static_cast<double>(row_kl[r]);
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) seg_sum[sidx] = red[0];
}
// stage 2: layer_term[l] = (sum of the layer's segments) / batch; loss = sum of FINITE layer terms (the reference
// skips inf/nan terms, train_qwenvl.py:606-609); valid[l] = 0/1.  Fixed summation order.
__global__ void kd_finalize_kernel(const double* __restrict__ seg_sum, const int* __restrict__ seg_layer, int n_seg,
                                   int n_layers, float inv_batch, float* __restrict__ layer_term,
                                   float* __restrict__ loss, int* __restrict__ valid) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float total = 0.f;
    for (int l = 0; l < n_layers; ++l) {
      double acc = 0.0;
      for (int sidx = 0; sidx < n_seg; ++sidx)
        if (seg_layer[sidx] == l) acc += seg_sum[sidx];
      const float v = static_cast<float>(acc * inv_batch);
      const int ok = isfinite(v) ? 1 : 0;
      layer_term[l] = v;
      valid[l] = ok;
      if (ok) total += v;
    }
    *loss = total;
  }
}
// per-row upstream for the backward: row_scale[r] = dloss * valid[layer(segment(r))] / batch
__global__ void kd_row_scale_kernel(const long long* __restrict__ seg_row_start, const int* __restrict__ seg_layer,
                                    const int* __restrict__ valid, const float* __restrict__ dloss, float inv_batch,
                                    float* __restrict__ row_scale) {
  const int sidx = blockIdx.y;
  const long long r = seg_row_start[sidx] + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r < seg_row_start[sidx + 1]) row_scale[r] = valid[seg_layer[sidx]] ? (*dloss) * inv_batch : 0.f;
}

// ------------------------------------------------------------------------------------------------
// Projector front end (utils/proj.py:62-72 + MLP3.layernorm :29):
//   y[b,s,:] = LayerNorm_H( mix(x[b,:,s,:]) ) * gamma + beta
//   mode 0: Conv2d(C -> 1, 5x5, pad 2) over the (S, H) plane + bias
//   mode 1: mean_c(cha_scale[c] * x)        mode 2: mean_c(x)
// Register-blocked stencil: a CTA owns PROJ_R consecutive output rows of one batch element, thread i owns 8 consecutive
// h.  Every input row s0-2 .. s0+R+1 of every channel is loaded ONCE per CTA (16 B + two 4 B halos per thread, no shared
// memory, no barriers in the main loop) and feeds all output rows it overlaps: 40 FMAs per (input row, output row).
// The op is FP32-FMA bound, not HBM bound: 25 MAC per input element = 25 FLOP/B.
__device__ __forceinline__ float block_sum_rt(float v, float* red, int nwarps) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  v = warp_sum(v);
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float s = 0.f;
  for (int j = 0; j < nwarps; ++j) s += red[j];  // fixed order
  return s;
}
// PROJ_R = 2 when B * S / 2 CTAs fill the machine (training batches), 1 otherwise: a single 512-token prompt gives only 256 CTAs of 8 warps at
// R = 2 -- 1.7 waves on 148 SMs at ~14 warps per SM, which left this FMA-bound stencil latency-bound (profiles/r02_prof_train_ncu_full.csv).
template <int PROJ_R>
__global__ void __launch_bounds__(512) proj_mix_ln_kernel(const __nv_bfloat16* __restrict__ x, int mode,
                                                          const float* __restrict__ w /* [C,5,5] | [C] */,
                                                          float conv_bias, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float eps,
                                                          __nv_bfloat16* __restrict__ y, int B, int C, int S, int H,
                                                          __nv_bfloat16* __restrict__ xm /* optional: pre-LN mix (training) */) {
  extern __shared__ float sm[];
  float* wsm = sm;                           // C*25 (mode 0) | C (mode 1)
  float* red = sm + ((C * 25 + 3) & ~3);     // one float per warp
  const int tiles = (S + PROJ_R - 1) / PROJ_R;
  const int b = blockIdx.x / tiles, s0 = (blockIdx.x - b * tiles) * PROJ_R;
  const int nwarps = blockDim.x >> 5;
  const int h0 = threadIdx.x * 8;
  const bool active = h0 < H;
  const int nw = mode == 0 ? C * 25 : (mode == 1 ? C : 0);
  for (int i = threadIdx.x; i < nw; i += blockDim.x) wsm[i] = w[i];
  __syncthreads();
  float acc[PROJ_R][8];
#pragma unroll
  for (int o = 0; o < PROJ_R; ++o)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[o][j] = 0.f;
  if (active) {
    if (mode == 0) {
      for (int c = 0; c < C; ++c) {
        float wk[25];
#pragma unroll
        for (int i = 0; i < 25; ++i) wk[i] = wsm[c * 25 + i];
        const __nv_bfloat16* xc = x + (static_cast<long long>(b) * C + c) * S * H;
        // all rows of this channel are requested before any is consumed: 3 x (R+4) loads in flight per thread
        uint4 raw[PROJ_R + 4];
        uint32_t lft[PROJ_R + 4], rgt[PROJ_R + 4];
#pragma unroll
        for (int ri = 0; ri < PROJ_R + 4; ++ri) {
          const int r = s0 + ri - 2;
          const bool ok = r >= 0 && r < S;  // zero padding in S (CTA-uniform)
          const __nv_bfloat16* xr = xc + static_cast<long long>(ok ? r : 0) * H + h0;
          raw[ri] = ok ? ld_stream(xr) : make_uint4(0, 0, 0, 0);
          lft[ri] = (ok && h0 > 0) ? __ldg(reinterpret_cast<const uint32_t*>(xr - 2)) : 0u;  // zero padding in H
          rgt[ri] = (ok && h0 + 8 < H) ? __ldg(reinterpret_cast<const uint32_t*>(xr + 8)) : 0u;
        }
#pragma unroll
        for (int ri = 0; ri < PROJ_R + 4; ++ri) {
          float v[12];
          unpack8(raw[ri], v + 2);
          v[0] = bf16_lo(lft[ri]); v[1] = bf16_hi(lft[ri]); v[10] = bf16_lo(rgt[ri]); v[11] = bf16_hi(rgt[ri]);
#pragma unroll
          for (int o = 0; o < PROJ_R; ++o) {
            const int dy = ri - o;  // = r - (s0 + o) + 2
            if (dy < 0 || dy > 4) continue;  // compile-time after unrolling
#pragma unroll
            for (int j = 0; j < 8; ++j)
              acc[o][j] += wk[dy * 5 + 0] * v[j] + wk[dy * 5 + 1] * v[j + 1] + wk[dy * 5 + 2] * v[j + 2] +
                           wk[dy * 5 + 3] * v[j + 3] + wk[dy * 5 + 4] * v[j + 4];
          }
        }
      }
    } else {
      for (int c = 0; c < C; ++c) {
        const float wc = mode == 1 ? wsm[c] : 1.0f;
#pragma unroll
        for (int o = 0; o < PROJ_R; ++o) {
          if (s0 + o >= S) continue;
          float f[8];
          unpack8(ld_stream(x + ((static_cast<long long>(b) * C + c) * S + s0 + o) * H + h0), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[o][j] += wc * f[j];
        }
      }
    }
  }
  const float post_mul = mode == 0 ? 1.0f : 1.0f / C;
  const float post_add = mode == 0 ? conv_bias : 0.f;
#pragma unroll
  for (int o = 0; o < PROJ_R; ++o) {
    if (s0 + o >= S) break;  // CTA-uniform
    float part = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[o][j] = active ? acc[o][j] * post_mul + post_add : 0.f;
      part += acc[o][j];
    }
    if (xm != nullptr && active) *reinterpret_cast<uint4*>(xm + (static_cast<long long>(b) * S + s0 + o) * H + h0) = pack8(acc[o]);
    const float mean = block_sum_rt(part, red, nwarps) / H;
    part = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[o][j] = active ? acc[o][j] - mean : 0.f;
      part += acc[o][j] * acc[o][j];
    }
    const float rstd = rsqrtf(block_sum_rt(part, red, nwarps) / H + eps);
    if (active) {
      float ov[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) ov[j] = acc[o][j] * rstd * gamma[h0 + j] + beta[h0 + j];
      *reinterpret_cast<uint4*>(y + (static_cast<long long>(b) * S + s0 + o) * H + h0) = pack8(ov);
    }
  }
}

// pooled[b, n] = mean_s y[b, s, n]   (MLP3: torch.mean(fc(x2), 1)); y bf16 [B, S, N] -> fp32/bf16 [B, N]
__global__ void mean_over_s_kernel(const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ out, int B, int S,
                                   int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N) return;
  const int b = i / N, n = i - b * N;
  float acc = 0.f;
  for (int s = 0; s < S; ++s) acc += __bfloat162float(y[(static_cast<long long>(b) * S + s) * N + n]);
  out[i] = __float2bfloat16(acc / S);
}

}  // namespace x2i
