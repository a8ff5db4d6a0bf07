// Row-wise (HBM-bound) backward kernels of the attention-distillation step: everything between the dgrad GEMMs and the
// attention backward of a FLUX block, plus the reductions that carry gradient into the AdaLN modulation (-> temb ->
// pooled projection -> projector).  The reference gets these from autograd over ~35 eager ops per block
// (train/train_qwenvl.py:625 through lightcontrol_flux.py:82-104, :159-204); here each is one fused pass.
// All column reductions are two-stage with a fixed summation order (deterministic).
#pragma once
#include "rowwise.cuh"

namespace x2i {

// ------------------------------------------------------------------------------------------------
// dy[r,:] = gate[b,:] * dx[r,:] (+ addend[r,:])       backward of  x' = x + gate * y  towards y (b = r / rows_per_batch);
// addend = the KD-loss gradient arriving at the hooked tensor y (train_qwenvl.py:186-214).
__global__ void gate_bwd_kernel(const __nv_bfloat16* __restrict__ dx, long long lddx, const __nv_bfloat16* __restrict__ gate,
                                long long gate_stride, const __nv_bfloat16* __restrict__ addend, long long ldadd,
                                __nv_bfloat16* __restrict__ dy, long long lddy, int rows, int D, int rows_per_batch) {
  const int nchunk = D >> 3;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(rows) * nchunk) return;
  const int r = static_cast<int>(i / nchunk), c = static_cast<int>(i - static_cast<long long>(r) * nchunk);
  float a[8], g[8];
  unpack8(reinterpret_cast<const uint4*>(dx + r * lddx)[c], a);
  unpack8(__ldg(reinterpret_cast<const uint4*>(gate + (r / rows_per_batch) * gate_stride) + c), g);
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] *= g[j];
  if (addend != nullptr) {
    float e[8];
    unpack8(reinterpret_cast<const uint4*>(addend + r * ldadd)[c], e);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += e[j];
  }
  reinterpret_cast<uint4*>(dy + r * lddy)[c] = pack8(a);
}

// ------------------------------------------------------------------------------------------------
// Backward of y = LN(x) * (1 + scale[b]) + shift[b]  (AFFINE: y = LN(x) * gamma + beta) towards x:
//   g = dn * (1 + scale);  dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) + dres
// Also writes stats[r] = (mean, rstd) for the column reductions (dscale/dshift).
// One 128-thread CTA per row (row in registers, MAXC 16-byte chunks per thread): small register footprint -> many rows in
// flight per SM, which is what an HBM-bound kernel with 3 reads + 1 write per element needs.
constexpr int LNB_THREADS = 128;
template <int MAXC, bool AFFINE>
__global__ void __launch_bounds__(LNB_THREADS) ln_mod_bwd_kernel(const __nv_bfloat16* __restrict__ dn, long long lddn,
                                                                 const __nv_bfloat16* __restrict__ x, long long ldx,
                                                                 const __nv_bfloat16* __restrict__ scale, long long mod_stride,
                                                                 const __nv_bfloat16* __restrict__ dres, long long ldr,
                                                                 __nv_bfloat16* __restrict__ dx, long long lddx,
                                                                 float2* __restrict__ stats, int rows, int D, int rows_per_batch,
                                                                 float eps) {
  __shared__ float red[2 * (LNB_THREADS / 32)];
  const int row = blockIdx.x;
  const int nchunk = D >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<long long>(row) * ldx);
  const uint4* gr = reinterpret_cast<const uint4*>(dn + static_cast<long long>(row) * lddn);
  const uint4* sc = reinterpret_cast<const uint4*>(scale + static_cast<long long>(row / rows_per_batch) * mod_stride);
  const uint4* rr = dres != nullptr ? reinterpret_cast<const uint4*>(dres + static_cast<long long>(row) * ldr) : nullptr;
  uint4 qx[MAXC], qg[MAXC], qr[MAXC];
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {  // all loads first
    const int c = i * LNB_THREADS + threadIdx.x;
    if (c < nchunk) {
      qx[i] = ld_stream(xr + c);
      qg[i] = ld_stream(gr + c);
      qr[i] = rr != nullptr ? ld_stream(rr + c) : make_uint4(0, 0, 0, 0);
    }
  }
  float v[MAXC][8], g[MAXC][8];
  float s2[2] = {0.f, 0.f};
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = i * LNB_THREADS + threadIdx.x;
    if (c < nchunk) {
      unpack8(qx[i], v[i]);
      unpack8(qg[i], g[i]);
      float a[8];
      unpack8(__ldg(sc + c), a);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s2[0] += v[i][j];
        g[i][j] *= AFFINE ? a[j] : 1.0f + a[j];
        s2[1] += g[i][j];
      }
    }
  }
  block_sum<2, LNB_THREADS / 32>(s2, red);
  const float mean = s2[0] / D, mg = s2[1] / D;
  float q1[1] = {0.f};
#pragma unroll
  for (int i = 0; i < MAXC; ++i)
    if (i * LNB_THREADS + threadIdx.x < nchunk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[i][j] -= mean;
        q1[0] += v[i][j] * v[i][j];
      }
    }
  block_sum<1, LNB_THREADS / 32>(q1, red);
  const float rstd = rsqrtf(q1[0] / D + eps);
  float sgx[1] = {0.f};
#pragma unroll
  for (int i = 0; i < MAXC; ++i)
    if (i * LNB_THREADS + threadIdx.x < nchunk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[i][j] *= rstd;  // xhat
        sgx[0] += g[i][j] * v[i][j];
      }
    }
  block_sum<1, LNB_THREADS / 32>(sgx, red);
  const float mgx = sgx[0] / D;
  if (threadIdx.x == 0 && stats != nullptr) stats[row] = make_float2(mean, rstd);
  uint4* outr = reinterpret_cast<uint4*>(dx + static_cast<long long>(row) * lddx);
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = i * LNB_THREADS + threadIdx.x;
    if (c < nchunk) {
      float o[8], e[8];
      unpack8(qr[i], e);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rstd * (g[i][j] - mg - v[i][j] * mgx) + e[j];
      outr[c] = pack8(o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Column reductions over the rows of each batch element (stage 1): a CTA owns 1024 columns x COLSUM_ROWS rows.
//   part0[b, split, col] = sum_r A[r, col]                      (if part0)
//   part1[b, split, col] = sum_r A[r, col] * Bv[r, col]         (if part1);  Bv = B, or (B - mean_r) * rstd_r with stats
// Used for dshift / dscale of every AdaLN (A = dn, B = LN input), dgate (A = dx, B = the un-gated branch output),
// the affine LayerNorm's dbeta / dgamma and bias gradients.
constexpr int COLSUM_ROWS = 32;
// Narrow matrices (D / 8 < 128 sixteen-byte chunks per row: the conv bias / GroupNorm / time-row sums of the ControlNeXt nets, C = 64-256
// channels over up to 262 144 pixel rows) fold `groups` = 128 / chunks row groups into one CTA: thread (g, c) owns chunk c of rows
// t0 + g, t0 + g + groups, ...; a split then covers COLSUM_ROWS * groups rows and the group sums are added in a fixed order through
// shared memory.  With one chunk column per thread a 128-channel tensor ran 16 live threads per CTA and 8192 splits, whose serial
// second stage took 25 us per launch (profiles/r01_lightcontrol_train_launch_list_summary.md).
__host__ __device__ inline int colsum_groups(int D) {
  const int chunks = D >> 3;
  int g = 1;
  while (g * 2 * chunks <= 128 && g < 16) g *= 2;
  return g;
}
__global__ void __launch_bounds__(128) colsum_partial_kernel(const __nv_bfloat16* __restrict__ A, long long lda,
                                                             const __nv_bfloat16* __restrict__ Bm, long long ldb,
                                                             const float2* __restrict__ stats, float* __restrict__ part0,
                                                             float* __restrict__ part1, int rows_per_batch, int D, int nsplit, int groups) {
  __shared__ float red[2][128][8];
  const int chunks = D >> 3;
  int c, g;
  if (groups > 1) {
    g = threadIdx.x / chunks;
    c = threadIdx.x - g * chunks;
  } else {
    g = 0;
    c = blockIdx.x * 128 + threadIdx.x;  // 16-byte chunk index
  }
  const bool live = c < chunks && g < groups;
  if (groups == 1 && !live) return;
  const int split = blockIdx.y, b = blockIdx.z;
  const int span = COLSUM_ROWS * groups;
  const int t0 = split * span, t1 = min(t0 + span, rows_per_batch);
  const bool two = part1 != nullptr;
  float s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s0[j] = s1[j] = 0.f;
  constexpr int U = 8;  // rows in flight per thread (2 x U independent 16-byte loads)
  if (live) {
    for (int t = t0 + g; t < t1; t += U * groups) {
      uint4 ra[U], rb[U];
      float2 st[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int tt = t + u * groups;
        const long long r = static_cast<long long>(b) * rows_per_batch + (tt < t1 ? tt : t);
        ra[u] = ld_stream(A + r * lda + c * 8);
        if (two) rb[u] = ld_stream(Bm + r * ldb + c * 8);
        st[u] = (two && stats != nullptr) ? stats[r] : make_float2(0.f, 1.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (t + u * groups < t1) {
          float a[8];
          unpack8(ra[u], a);
#pragma unroll
          for (int j = 0; j < 8; ++j) s0[j] += a[j];
          if (two) {
            float v[8];
            unpack8(rb[u], v);
#pragma unroll
            for (int j = 0; j < 8; ++j) s1[j] += a[j] * ((v[j] - st[u].x) * st[u].y);
          }
        }
      }
    }
  }
  if (groups > 1) {  // add the row groups of this CTA in a fixed order
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[0][threadIdx.x][j] = s0[j];
      red[1][threadIdx.x][j] = s1[j];
    }
    __syncthreads();
    if (g != 0 || !live) return;
#pragma unroll
    for (int j = 0; j < 8; ++j) s0[j] = s1[j] = 0.f;
    for (int gg = 0; gg < groups; ++gg)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s0[j] += red[0][gg * chunks + c][j];
        s1[j] += red[1][gg * chunks + c][j];
      }
  }
  const long long o = (static_cast<long long>(b) * nsplit + split) * D + c * 8;
  if (part0 != nullptr) {
    *reinterpret_cast<float4*>(part0 + o) = make_float4(s0[0], s0[1], s0[2], s0[3]);
    *reinterpret_cast<float4*>(part0 + o + 4) = make_float4(s0[4], s0[5], s0[6], s0[7]);
  }
  if (part1 != nullptr) {
    *reinterpret_cast<float4*>(part1 + o) = make_float4(s1[0], s1[1], s1[2], s1[3]);
    *reinterpret_cast<float4*>(part1 + o + 4) = make_float4(s1[4], s1[5], s1[6], s1[7]);
  }
}
// stage 2: out[b * ldo + col] (+)= sum over splits.  A CTA owns 32 columns: its blockDim.x / 32 thread groups (8 for wide matrices, 32
// for narrow ones whose grid would otherwise be 2-4 CTAs of long serial loops) each sum every G-th partial (coalesced 128-byte rows),
// then the group sums are added in a fixed order (deterministic).
__global__ void __launch_bounds__(1024) colsum_final_kernel(const float* __restrict__ part, float* __restrict__ out, long long ldo,
                                                            int D, int nsplit, int accumulate) {
  __shared__ float red[32][32];
  const int cl = threadIdx.x & 31, grp = threadIdx.x >> 5, ngrp = blockDim.x >> 5;
  const int col = blockIdx.x * 32 + cl;
  const int b = blockIdx.y;
  float s = 0.f;
  if (col < D)
    for (int i = grp; i < nsplit; i += ngrp) s += part[(static_cast<long long>(b) * nsplit + i) * D + col];
  red[grp][cl] = s;
  __syncthreads();
  if (grp == 0 && col < D) {
    float t = 0.f;
    for (int j = 0; j < ngrp; ++j) t += red[j][cl];
    float* o = out + static_cast<long long>(b) * ldo + col;
    *o = accumulate ? *o + t : t;
  }
}

// ------------------------------------------------------------------------------------------------
// Backward of the QKV epilogue (per-head RMSNorm(128) * w, then RoPE) + head-major -> token-major transposition:
//   dq, dk, dv [B, H, L_total, 128] (grads w.r.t. the post-RoPE q, k and v the attention consumed)
//   qk_pre [M, ldqk]: pre-norm q | k of this stream's tokens (saved by the forward epilogue)
//   out [M, ldo]: [dq_pre | dk_pre | dv] (the A operand of the QKV dgrad GEMM)
// One half-warp per (token, head): 8 elements per lane.
__global__ void __launch_bounds__(256) qk_norm_rope_bwd_kernel(const __nv_bfloat16* __restrict__ dq, const __nv_bfloat16* __restrict__ dk,
                                                               const __nv_bfloat16* __restrict__ dv,
                                                               const __nv_bfloat16* __restrict__ qk_pre, long long ldqk,
                                                               const __nv_bfloat16* __restrict__ rms_q,
                                                               const __nv_bfloat16* __restrict__ rms_k,
                                                               const float2* __restrict__ rope, __nv_bfloat16* __restrict__ out,
                                                               long long ldo, int M, int H, int rows_per_batch, int row_offset,
                                                               int L_total, float eps) {
  const long long unit = static_cast<long long>(blockIdx.x) * 16 + (threadIdx.x >> 4);
  const bool live = unit < static_cast<long long>(M) * H;
  const long long u = live ? unit : 0;
  const int m = static_cast<int>(u / H), h = static_cast<int>(u - static_cast<long long>(m) * H);
  const int l16 = threadIdx.x & 15;
  const int b = m / rows_per_batch, pos = row_offset + (m - b * rows_per_batch);
  const int D = H * 128;
  const long long src = ((static_cast<long long>(b) * H + h) * L_total + pos) * 128 + l16 * 8;
  __nv_bfloat16* orow = out + static_cast<long long>(m) * ldo + h * 128 + l16 * 8;
#pragma unroll
  for (int sec = 0; sec < 2; ++sec) {
    float dy[8], t[8], w[8];
    unpack8(*reinterpret_cast<const uint4*>((sec == 0 ? dq : dk) + src), dy);
    if (rope != nullptr) {
      const float4* rp = reinterpret_cast<const float4*>(rope + static_cast<long long>(pos) * 64 + l16 * 4);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float4 cs = __ldg(rp + j);  // (cos0, sin0, cos1, sin1)
        const float a0 = dy[4 * j], a1 = dy[4 * j + 1], a2 = dy[4 * j + 2], a3 = dy[4 * j + 3];
        dy[4 * j] = a0 * cs.x + a1 * cs.y;
        dy[4 * j + 1] = a1 * cs.x - a0 * cs.y;
        dy[4 * j + 2] = a2 * cs.z + a3 * cs.w;
        dy[4 * j + 3] = a3 * cs.z - a2 * cs.w;
      }
    }
    unpack8(*reinterpret_cast<const uint4*>(qk_pre + static_cast<long long>(m) * ldqk + sec * D + h * 128 + l16 * 8), t);
    unpack8(__ldg(reinterpret_cast<const uint4*>((sec == 0 ? rms_q : rms_k) + l16 * 8)), w);
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += t[j] * t[j];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rinv = rsqrtf(ss * (1.0f / 128.0f) + eps);
    float gx = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dy[j] *= w[j];     // g
      t[j] *= rinv;      // xhat
      gx += dy[j] * t[j];
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) gx += __shfl_xor_sync(0xffffffffu, gx, o);
    gx *= (1.0f / 128.0f);
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = rinv * (dy[j] - t[j] * gx);
    if (live) *reinterpret_cast<uint4*>(orow + sec * D) = pack8(r);
  }
  if (live) *reinterpret_cast<uint4*>(orow + 2 * D) = *reinterpret_cast<const uint4*>(dv + src);
}

// ------------------------------------------------------------------------------------------------
// Prologue of the attention backward: token-major dO (+ optional addend, the KD gradient of a hooked single-block
// attention output) and O  ->  head-major dO [B, H, L, 128] and delta[b, h, pos] = sum_d dO * O  ([B*H, Lpad], 0 in the pad).
// Token-major sources are split like the forward's outputs: rows pos < split live in p0 (row (b*split + pos) * ld0), the
// rest in p1 (row (b*(L-split) + pos-split) * ld1).
struct TokSrc {
  const __nv_bfloat16* p0;
  long long ld0;
  const __nv_bfloat16* p1;
  long long ld1;
};
__device__ __forceinline__ const __nv_bfloat16* tok_ptr(const TokSrc& s, int b, int pos, int split, int L) {
  return pos < split ? s.p0 + (static_cast<long long>(b) * split + pos) * s.ld0
                     : s.p1 + (static_cast<long long>(b) * (L - split) + (pos - split)) * s.ld1;
}
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const TokSrc dO, const TokSrc O, const TokSrc add,
                                                            __nv_bfloat16* __restrict__ do_hm, float* __restrict__ delta, int B,
                                                            int H, int L, int Lpad, int split) {
  const long long unit = static_cast<long long>(blockIdx.x) * 16 + (threadIdx.x >> 4);
  const bool live = unit < static_cast<long long>(B) * Lpad * H;
  const long long u = live ? unit : 0;
  const int h = static_cast<int>(u % H);
  const long long bp = u / H;
  const int pos = static_cast<int>(bp % Lpad), b = static_cast<int>(bp / Lpad);
  const int l16 = threadIdx.x & 15;
  const bool real = pos < L;
  float g[8], o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) g[j] = o[j] = 0.f;
  if (real) {
    const int col = h * 128 + l16 * 8;
    unpack8(*reinterpret_cast<const uint4*>(tok_ptr(dO, b, pos, split, L) + col), g);
    unpack8(*reinterpret_cast<const uint4*>(tok_ptr(O, b, pos, split, L) + col), o);
    if (add.p0 != nullptr || add.p1 != nullptr) {
      float e[8];
      unpack8(*reinterpret_cast<const uint4*>(tok_ptr(add, b, pos, split, L) + col), e);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += e[j];
    }
  }
  const uint4 packed = pack8(g);
  // delta from the bf16-rounded dO the MMAs will consume
  float gr[8];
  unpack8(packed, gr);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += gr[j] * o[j];
#pragma unroll
  for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (live) {
    if (real) *reinterpret_cast<uint4*>(do_hm + ((static_cast<long long>(b) * H + h) * L + pos) * 128 + l16 * 8) = packed;
    if (l16 == 0) delta[(static_cast<long long>(b) * H + h) * Lpad + pos] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// Transposed skinny linear: out[b, k] = act'(pre[b, k]) * sum_n g[b, n] * W[n, k]      (B <= 8 rows per launch)
// -- the backward of every AdaLN modulation linear of a step in one pass over the concatenated weights (6.5 GB, read
// once), and of the time/text embedding MLPs.  Stage 1: CTA `s` reduces its slab of n rows, thread t owns columns
// [8t, 8t+8) (W rows are read fully coalesced); stage 2 sums the slabs in a fixed order.  g fp32, W bf16, out fp32.
template <int MAXB>
__global__ void __launch_bounds__(512) skinny_linear_t_kernel(const float* __restrict__ g, long long ldg,
                                                               const __nv_bfloat16* __restrict__ W, long long ldw,
                                                               float* __restrict__ part /* [nslab, B, K] */, int B, int N, int K,
                                                               int rows_per_slab) {
  const int c = threadIdx.x;
  if (c * 8 >= K) return;
  const int n0 = blockIdx.x * rows_per_slab, n1 = min(n0 + rows_per_slab, N);
  float acc[MAXB][8];
#pragma unroll
  for (int b = 0; b < MAXB; ++b)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[b][j] = 0.f;
  constexpr int U = 8;  // W rows in flight per thread
  const __nv_bfloat16* wp = W + static_cast<long long>(n0) * ldw + c * 8;
  int n = n0;
  for (; n + U <= n1; n += U, wp += U * ldw) {  // full groups: U independent 16-byte loads, then the math
    uint4 wq[U];
#pragma unroll
    for (int u = 0; u < U; ++u) wq[u] = __ldcs(reinterpret_cast<const uint4*>(wp + u * ldw));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float w[8];
      unpack8(wq[u], w);
#pragma unroll
      for (int b = 0; b < MAXB; ++b) {
        if (b < B) {
          const float gv = __ldg(g + static_cast<long long>(b) * ldg + n + u);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[b][j] += gv * w[j];
        }
      }
    }
  }
  for (; n < n1; ++n, wp += ldw) {
    float w[8];
    unpack8(__ldcs(reinterpret_cast<const uint4*>(wp)), w);
#pragma unroll
    for (int b = 0; b < MAXB; ++b) {
      if (b < B) {
        const float gv = __ldg(g + static_cast<long long>(b) * ldg + n);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[b][j] += gv * w[j];
      }
    }
  }
#pragma unroll
  for (int b = 0; b < MAXB; ++b) {
    if (b < B) {
      float* o = part + (static_cast<long long>(blockIdx.x) * B + b) * K + c * 8;
      *reinterpret_cast<float4*>(o) = make_float4(acc[b][0], acc[b][1], acc[b][2], acc[b][3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(acc[b][4], acc[b][5], acc[b][6], acc[b][7]);
    }
  }
}
// dact: 0 none, 1 SiLU'(pre).  Same 32-column x 8-group layout as colsum_final_kernel.
__global__ void __launch_bounds__(256) skinny_linear_t_final_kernel(const float* __restrict__ part, const __nv_bfloat16* __restrict__ pre,
                                                                    long long ldpre, float* __restrict__ out, long long ldo, int B, int K,
                                                                    int nslab, int dact, int accumulate) {
  __shared__ float red[8][32];
  const int cl = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + cl;
  const int b = blockIdx.y;
  float s = 0.f;
  if (k < K)
    for (int i = grp; i < nslab; i += 8) s += part[(static_cast<long long>(i) * B + b) * K + k];
  red[grp][cl] = s;
  __syncthreads();
  if (grp == 0 && k < K) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][cl];
    if (dact == 1) t *= dsilu_f(__bfloat162float(pre[static_cast<long long>(b) * ldpre + k]));
    float* o = out + static_cast<long long>(b) * ldo + k;
    *o = accumulate ? *o + t : t;
  }
}

// ------------------------------------------------------------------------------------------------
// Backward of pooled = mean_s y[b, s, :]  (utils/proj.py:32):  dy[b, s, :] = dpooled[b, :] / S
__global__ void mean_over_s_bwd_kernel(const __nv_bfloat16* __restrict__ dpooled, __nv_bfloat16* __restrict__ dy, int B, int S, int N) {
  const int nchunk = N >> 3;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * S * nchunk) return;
  const int c = static_cast<int>(i % nchunk);
  const int b = static_cast<int>(i / (static_cast<long long>(S) * nchunk));
  float a[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(dpooled + static_cast<long long>(b) * N) + c), a);
  const float inv = 1.0f / S;
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] *= inv;
  reinterpret_cast<uint4*>(dy)[i] = pack8(a);
}

// ------------------------------------------------------------------------------------------------
// Weight gradient of the projector's layer-mixing front end (utils/proj.py:62-72):
//   CONV : dW[c, ky, kx] = sum_{b,s,h} g[b,s,h] * x[b, c, s+ky-2, h+kx-2]     (Conv2d(C -> 1, 5x5, pad 2))
//   !CONV: dW[c]         = sum_{b,s,h} g[b,s,h] * x[b, c, s, h]               (cha_scale; caller divides by C)
// g = gradient w.r.t. the mixed [B,S,H] plane.  Same register-blocked stencil as the forward: a CTA owns PROJB_R rows of
// one (b, c) plane, thread i owns 8 consecutive h; every x row is loaded once and meets the 5 g rows it overlaps.
// Stage 1 writes part[(b * tiles + tile) * C + c][25]; stage 2 sums b and tiles in a fixed order.
constexpr int PROJB_R = 4;
template <bool CONV>
__global__ void __launch_bounds__(512) proj_mix_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ g,
                                                             float* __restrict__ part, int B, int C, int S, int H) {
  __shared__ float red[16][25];
  const int tiles = gridDim.x;
  const int tile = blockIdx.x, c = blockIdx.y, b = blockIdx.z;
  const int s0 = tile * PROJB_R;
  const int h0 = threadIdx.x * 8;
  const bool active = h0 < H;
  constexpr int NT = CONV ? 25 : 1;
  float acc[NT];
#pragma unroll
  for (int i = 0; i < NT; ++i) acc[i] = 0.f;
  if (active) {
    const __nv_bfloat16* xc = x + (static_cast<long long>(b) * C + c) * S * H;
    const __nv_bfloat16* gb = g + static_cast<long long>(b) * S * H;
    float gv[PROJB_R][8];
#pragma unroll
    for (int o = 0; o < PROJB_R; ++o) {
      if (s0 + o < S) {
        unpack8(*reinterpret_cast<const uint4*>(gb + static_cast<long long>(s0 + o) * H + h0), gv[o]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) gv[o][j] = 0.f;
      }
    }
    if constexpr (CONV) {
#pragma unroll
      for (int ri = 0; ri < PROJB_R + 4; ++ri) {
        const int r = s0 + ri - 2;
        if (r < 0 || r >= S) continue;  // zero padding (CTA-uniform)
        const __nv_bfloat16* xr = xc + static_cast<long long>(r) * H + h0;
        float v[12];
        unpack8(ld_stream(xr), v + 2);
        const uint32_t lft = h0 > 0 ? __ldg(reinterpret_cast<const uint32_t*>(xr - 2)) : 0u;
        const uint32_t rgt = h0 + 8 < H ? __ldg(reinterpret_cast<const uint32_t*>(xr + 8)) : 0u;
        v[0] = bf16_lo(lft); v[1] = bf16_hi(lft); v[10] = bf16_lo(rgt); v[11] = bf16_hi(rgt);
#pragma unroll
        for (int o = 0; o < PROJB_R; ++o) {
          const int ky = ri - o;  // x row r = (s0 + o) + ky - 2
          if (ky < 0 || ky > 4) continue;
#pragma unroll
          for (int kx = 0; kx < 5; ++kx)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[ky * 5 + kx] += gv[o][j] * v[j + kx];
        }
      }
    } else {
#pragma unroll
      for (int o = 0; o < PROJB_R; ++o) {
        if (s0 + o >= S) continue;
        float v[8];
        unpack8(ld_stream(xc + static_cast<long long>(s0 + o) * H + h0), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[0] += gv[o][j] * v[j];
      }
    }
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const float sres = warp_sum(acc[i]);
    if (l == 0) red[w][i] = sres;
  }
  __syncthreads();
  if (threadIdx.x < NT) {
    float sres = 0.f;
    for (int j = 0; j < nw; ++j) sres += red[j][threadIdx.x];
    part[((static_cast<long long>(b) * tiles + tile) * C + c) * NT + threadIdx.x] = sres;
  }
}
__global__ void proj_mix_wgrad_final_kernel(const float* __restrict__ part, float* __restrict__ out, int n_parts, int n_out, float mul) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (c, tap)
  if (i >= n_out) return;
  float sres = 0.f;
  for (int p = 0; p < n_parts; ++p) sres += part[static_cast<long long>(p) * n_out + i];
  out[i] = sres * mul;
}

// fp32 -> bf16 row copy (gradients leave the fp32 reduction buffers as bf16 tensors)
__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16(in[i]);
}

}  // namespace x2i
