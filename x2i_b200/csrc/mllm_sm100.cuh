// Row-wise kernels of the MLLM prefill that feeds the alignment projector: the decoder stack of the Qwen2.5-VL
// text model run ONCE over the padded prompt with every layer's hidden state kept (infer/inference_qwenvl.py:121-132,:176-179;
// train/train_qwenvl.py:773-775).  The contractions run on the tcgen05 GEMMs (EPI_BIAS, EPI_SWIGLU, EPI_GATE_RESIDUAL) and the
// causal / grouped-query form of the fused attention kernel; these are the HBM-bound pieces in between.
#pragma once
#include "rowwise.cuh"

namespace x2i {

// out[r, :] = table[ids[r], :]   (nn.Embedding; index work, bit-exact).  One warp per row, 16-byte vectors.
// Rows are addressed as (b, s): out + b * out_batch_stride + s * ldo, so the embeddings land directly in layer slot 0 of the
// projector's [B, C, S, H] input.
__global__ void __launch_bounds__(256) gather_rows_kernel(const long long* __restrict__ ids, const __nv_bfloat16* __restrict__ table,
                                                          long long ldt, int vocab, __nv_bfloat16* __restrict__ out, long long ldo,
                                                          long long out_batch_stride, int rows, int rows_per_batch, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  long long id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);  // an out-of-range id would read outside the table: clamp (the reference raises)
  const int b = row / rows_per_batch, s = row - b * rows_per_batch;
  const uint4* src = reinterpret_cast<const uint4*>(table + id * ldt);
  uint4* dst = reinterpret_cast<uint4*>(out + b * out_batch_stride + static_cast<long long>(s) * ldo);
  for (int c = lane; c < (D >> 3); c += 32) dst[c] = __ldg(src + c);
}

// y[r, :] = w * bf16( x[r, :] * rsqrt(mean(x^2) + eps) )   -- Qwen2RMSNorm: variance in fp32, the normalised row rounded to the input
// dtype BEFORE the weight multiply.  One warp per row, row in registers (D <= 256 * MAXC).  x rows are addressed as (b, s) with a
// batch stride (they live in a layer slot of the [B, C, S, H] capture buffer); y is a dense [rows, D] GEMM operand.
template <int MAXC>
__global__ void __launch_bounds__(256) rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, long long x_batch_stride,
                                                      const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ y, long long ldy,
                                                      long long y_batch_stride, int rows, int rows_per_batch, int D, float eps) {
  griddep_launch();  // PDL (common.cuh)
  griddep_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nchunk = D >> 3;
  const int b = row / rows_per_batch, s = row - b * rows_per_batch;
  const uint4* xr = reinterpret_cast<const uint4*>(x + b * x_batch_stride + static_cast<long long>(s) * ldx);
  float v[MAXC][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = i * 32 + lane;
    if (c < nchunk) {
      unpack8(xr[c], v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += v[i][j] * v[i][j];
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / D + eps);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  uint4* yr = reinterpret_cast<uint4*>(y + b * y_batch_stride + static_cast<long long>(s) * ldy);
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = i * 32 + lane;
    if (c < nchunk) {
      float ww[8], o[8];
      unpack8(__ldg(wr + c), ww);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = ww[j] * __bfloat162float(__float2bfloat16_rn(v[i][j] * rstd));
      yr[c] = pack8(o);
    }
  }
}

// Split the fused QKV projection of a grouped-query decoder layer into head-major q / k / v and apply the rotate-half RoPE of
// Qwen2 (apply_multimodal_rotary_pos_emb with text-only positions: the three M-RoPE sections carry the same position, which is plain
// 1-D RoPE):   out[d] = x[d] cos(p f_d) - x[d + 64] sin(p f_d),   out[d + 64] = x[d + 64] cos(p f_d) + x[d] sin(p f_d),  d < 64.
// qkv [rows, ld] token-major = [q (H*128) | k (Hkv*128) | v (Hkv*128)]; q -> [B, H, S, 128], k, v -> [B, Hkv, S, 128];
// pos int32 [rows] (position of every token: cumsum(attention_mask) - 1, padded tokens 1); inv_freq fp32 [64].
// One warp per token; lane l owns frequencies l and l + 32 (its two sin / cos pairs serve all H + Hkv heads).
__global__ void __launch_bounds__(256) rope_half_split_kernel(const __nv_bfloat16* __restrict__ qkv, long long ld, const int* __restrict__ pos,
                                                              const float* __restrict__ inv_freq, __nv_bfloat16* __restrict__ q,
                                                              __nv_bfloat16* __restrict__ k, __nv_bfloat16* __restrict__ v, int rows,
                                                              int S, int H, int Hkv) {
  griddep_launch();  // PDL (common.cuh)
  griddep_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int b = row / S, s = row - b * S;
  const float p = static_cast<float>(pos[row]);
  float c0, s0, c1, s1;
  sincosf(p * inv_freq[lane], &s0, &c0);
  sincosf(p * inv_freq[lane + 32], &s1, &c1);
  const __nv_bfloat16* xr = qkv + static_cast<long long>(row) * ld;
  for (int hh = 0; hh < H + Hkv; ++hh) {
    const __nv_bfloat16* x = xr + hh * 128;
    __nv_bfloat16* dst = hh < H ? q + ((static_cast<long long>(b) * H + hh) * S + s) * 128
                                : k + ((static_cast<long long>(b) * Hkv + (hh - H)) * S + s) * 128;
    const float a0 = __bfloat162float(x[lane]), a1 = __bfloat162float(x[lane + 32]);
    const float b0 = __bfloat162float(x[lane + 64]), b1 = __bfloat162float(x[lane + 96]);
    dst[lane] = __float2bfloat16_rn(a0 * c0 - b0 * s0);
    dst[lane + 32] = __float2bfloat16_rn(a1 * c1 - b1 * s1);
    dst[lane + 64] = __float2bfloat16_rn(b0 * c0 + a0 * s0);
    dst[lane + 96] = __float2bfloat16_rn(b1 * c1 + a1 * s1);
  }
  const uint2* vs = reinterpret_cast<const uint2*>(xr + (H + Hkv) * 128);
  for (int hh = 0; hh < Hkv; ++hh) {
    uint2* dst = reinterpret_cast<uint2*>(v + ((static_cast<long long>(b) * Hkv + hh) * S + s) * 128);
    dst[lane] = vs[hh * 32 + lane];
  }
}

}  // namespace x2i
