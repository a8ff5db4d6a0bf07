// Persistent, warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   acc[M,N] = A[M,K] (bf16, K contiguous) x W[N,K]^T (bf16, K contiguous -- torch Linear layout)
//
// One CTA per SM, 192 threads: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one lane),
// warps 2..5 = epilogue (TMEM -> registers -> fused math -> global).  A/B tiles are staged by TMA
// into a multi-stage 128B-swizzled shared-memory ring; the fp32 accumulator lives in TMEM and is
// double-buffered (2 x BN columns) so the epilogue of tile i overlaps the main loop of tile i+1.
//
// Epilogues (the reference ops they replace are cited in include/x2i_b200.h):
//   EPI_BIAS            C = acc + bias; optional second output aux = gelu_erf(C)
//   EPI_BIAS_GELU_TANH  C = gelu_tanh(acc + bias)                        FeedForward net.0 / proj_mlp
//   EPI_BIAS_GELU_ERF   C = gelu_erf(acc + bias)                         projector MLP3
//   EPI_GATE_RESIDUAL   C = residual + gate[b,:] * (acc + bias); aux = acc + bias (optional, KD hook tensor)
//   EPI_QKV             per 128-column head tile: q,k -> +bias, RMSNorm(128)*w, RoPE, head-major store;
//                       v -> +bias, head-major store; columns >= 3D -> gelu_tanh -> mlp buffer (single block);
//                       training: also the pre-norm q,k (token-major) and the pre-GELU mlp values the backward needs
//   EPI_DACT            backward of a Linear followed by GELU: C = addend + acc * act'(pre) for columns >= n_split,
//                       C = addend + acc below it (dgrad GEMMs of the distillation step)
//
//   EPI_SWIGLU          gated MLP of a decoder LM (Qwen2MLP: down(silu(gate(x)) * up(x))): W holds gate and up rows interleaved in
//                       blocks of 128, so every 256-column accumulator tile is [gate(128) | up(128)] of the same 128 outputs:
//                       C[:, n/2 ..] = silu(acc_gate + b) * (acc_up + b)           (BN = 256 only)
//   EPI_CONV            implicit-GEMM convolution (conv_sm100.cuh): C = relu?(acc + bias + rowvec[b, :]) + residual
//
// Operand forms: B_MN takes B as [K, N] (N contiguous) and A_MN takes A as [K, M] (M contiguous) -- the "MN-major"
// UMMA operands.  dgrad (dX = dY W) uses B_MN on the weights exactly as nn.Linear stores them; wgrad (dW = dY^T X) uses
// both, so no tensor is ever transposed in memory.
#pragma once
#include "common.cuh"

namespace x2i {

enum { EPI_BIAS = 0, EPI_BIAS_GELU_TANH = 1, EPI_BIAS_GELU_ERF = 2, EPI_GATE_RESIDUAL = 3, EPI_QKV = 4, EPI_DACT = 5, EPI_CONV = 6, EPI_SWIGLU = 7 };

struct GemmParams {
  int M, N, K;
  const __nv_bfloat16* bias;  // [N] or null
  __nv_bfloat16* C;           // [M, ldc]
  long long ldc;
  // EPI_GATE_RESIDUAL
  const __nv_bfloat16* residual;  // [M, ldr] (may alias C)
  long long ldr;
  const __nv_bfloat16* gate;  // gate[(m / rows_per_batch) * gate_stride + n]
  long long gate_stride;
  int rows_per_batch;
  __nv_bfloat16* aux;  // optional un-gated output [M, ldaux]
  long long ldaux;
  // EPI_QKV
  __nv_bfloat16 *q, *k, *v;           // [B, H, L_total, 128]
  const __nv_bfloat16 *rms_q, *rms_k;  // [128]
  const float2* rope;                  // [L_total, 64] (cos, sin) per rotated pair, or null
  int L_total, row_offset, heads;
  float eps;
  __nv_bfloat16* mlp;  // [M, ldmlp]
  long long ldmlp;
  // training-mode saves of EPI_QKV (nullable): pre-norm q|k token-major [M, ldqk] and pre-GELU mlp [M, ldmlp_pre]
  __nv_bfloat16* qk_pre;
  long long ldqk;
  __nv_bfloat16* mlp_pre;
  long long ldmlp_pre;
  // EPI_BIAS: aux_act selects the activation of the second output (0/2 gelu_erf, 1 gelu_tanh)
  int aux_act;
  // EPI_BIAS with an fp32 destination (attention scores of the VAE mid block): C32 = alpha * (acc + bias); C is not written
  float* c32;
  long long ldc32;
  float alpha;
  // split-K (single-CTA kernel only; the weight-gradient GEMMs of narrow layers have a handful of output tiles and a huge K): the
  // k-blocks are dealt to `ksplit` tile groups, group s stores its raw fp32 accumulator to c32 + s * M * ldc32 (no epilogue math);
  // splitk_reduce_kernel adds the partials in a fixed order.  0 / 1 = off.
  int ksplit;
  // EPI_DACT: pre-activation [M, ldpre] (column n - n_split), derivative kind dact (1 tanh-GELU, 2 erf-GELU), addend = residual
  const __nv_bfloat16* pre;
  long long ldpre;
  int n_split, dact;
  // EPI_CONV: per-image channel vector added before the activation (ResnetBlock2D time embedding), ReLU flag; residual/ldr
  const __nv_bfloat16* rowvec;
  long long rowvec_stride;
  int relu;
  // B_CONV (implicit convolution weight gradient): B = the im2col matrix of an NHWC tensor, never materialised -- the producer loads
  // 64-pixel x 64-channel boxes of the tensor itself, shifted by the tap of the column block (zero padding = TMA's OOB fill)
  int cv_cin, cv_kw, cv_pad, cv_wo, cv_howo;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;  // 2 accumulator stages (power of two for BN in {64,128,256})
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// 64 B of one input row per thread (read-only data): two 256-bit loads when 32-byte aligned (full-sector requests, see store_bf16x32)
__device__ __forceinline__ void load_bf16x32(const __nv_bfloat16* p, float (&out)[32]) {
  if ((reinterpret_cast<uintptr_t>(p) & 31) == 0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      uint32_t w[8];
      ld_global_nc_256(p + 16 * i, w);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        out[16 * i + 2 * j] = bf16_lo(w[j]);
        out[16 * i + 2 * j + 1] = bf16_hi(w[j]);
      }
    }
    return;
  }
  const uint4* p4 = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u = __ldg(p4 + i);
    out[8 * i + 0] = bf16_lo(u.x); out[8 * i + 1] = bf16_hi(u.x);
    out[8 * i + 2] = bf16_lo(u.y); out[8 * i + 3] = bf16_hi(u.y);
    out[8 * i + 4] = bf16_lo(u.z); out[8 * i + 5] = bf16_hi(u.z);
    out[8 * i + 6] = bf16_lo(u.w); out[8 * i + 7] = bf16_hi(u.w);
  }
}
// 64 B of one output row per thread.  Two 256-bit stores (sm_100: STG.E.256) when the row segment is 32-byte aligned, so that every request is
// a FULL 32-byte sector: with four 128-bit stores each warp instruction carries 32 half-sector writes to 32 different rows, and that
// request stream competes with the TMA loads of the shared-memory-bound mainloop (tools/gemm_probe.py: an epilogue that only reads TMEM costs
// nothing, one that only stores costs the plain epilogue's 8 %; profiles/r02_gemm_epilogue_probe.md).
__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* p, const float (&v)[32]) {
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
  if ((reinterpret_cast<uintptr_t>(p) & 31) == 0) {
    st_global_256(p, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
    st_global_256(p + 16, w[8], w[9], w[10], w[11], w[12], w[13], w[14], w[15]);
  } else {
    uint4* p4 = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) p4[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  }
}

// Fused epilogue of one 128 x BN accumulator tile: thread `m` owns output row m; t_acc is the TMEM address of the tile
// (lane quadrant of the calling warp already applied).
template <int BN, int EPI>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmParams& p, const uint32_t t_acc, const int m, const int n_tile0,
                                                   const long long bias_off = 0 /* EPI_CONV with per-group weights */,
                                                   const int ks = 0 /* split-K group */
#ifdef X2I_EPI_STAGE
                                                   , uint8_t* stage = nullptr /* experiment builds: 4 KB of shared memory per epilogue warp */
#endif
                                                   ) {
  const bool row_ok = m < p.M;
  if (p.ksplit > 1) {  // raw fp32 partial of this k-group
    float* dst = p.c32 + (static_cast<long long>(ks) * p.M + m) * p.ldc32 + n_tile0;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      if (n_tile0 + c * 32 >= p.N) break;
      uint32_t r[32];
      tmem_ld32(t_acc + c * 32, r);
      tmem_ld_wait();
      if (row_ok) {
        float4* d4 = reinterpret_cast<float4*>(dst + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          d4[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
      }
    }
    return;
  }

  if constexpr (EPI == EPI_QKV) {
    const int D = p.heads * 128;
    const int b = row_ok ? m / p.rows_per_batch : 0;
    const int tok = row_ok ? m - b * p.rows_per_batch : 0;
    const int pos = p.row_offset + tok;
#pragma unroll 1
    for (int hh = 0; hh < BN / 128; ++hh) {
      const int n_h = n_tile0 + hh * 128;
      if (n_h >= p.N) break;
      const int sec = n_h / D;  // 0 q, 1 k, 2 v, >=3 mlp
      uint32_t r[32];
      float x[32], bb[32];
      const __nv_bfloat16* wnorm = sec == 0 ? p.rms_q : p.rms_k;
      if (sec < 2 && wnorm != nullptr) {
        float ss = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          tmem_ld32(t_acc + hh * 128 + c * 32, r);
          load_bf16x32(p.bias + n_h + c * 32, bb);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            x[j] = __uint_as_float(r[j]) + bb[j];
            ss += x[j] * x[j];
          }
          if (p.qk_pre != nullptr && row_ok) store_bf16x32(p.qk_pre + static_cast<long long>(m) * p.ldqk + n_h + c * 32, x);
        }
        const float rinv = rsqrtf(ss * (1.0f / 128.0f) + p.eps);
        const int h = (n_h - sec * D) >> 7;
        __nv_bfloat16* dst = (sec == 0 ? p.q : p.k) +
                             ((static_cast<long long>(b) * p.heads + h) * p.L_total + pos) * 128;
        const __nv_bfloat16* w = sec == 0 ? p.rms_q : p.rms_k;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          tmem_ld32(t_acc + hh * 128 + c * 32, r);
          load_bf16x32(p.bias + n_h + c * 32, bb);
          float ww[32];
          load_bf16x32(w + c * 32, ww);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = (__uint_as_float(r[j]) + bb[j]) * rinv * ww[j];
          if (p.rope != nullptr && row_ok) {
            const float4* rp = reinterpret_cast<const float4*>(p.rope + static_cast<long long>(pos) * 64 + c * 16);
            float4 csv[8];  // 128 B of this row's table: four 256-bit loads when the table is 32-byte aligned
            if ((reinterpret_cast<uintptr_t>(rp) & 31) == 0) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint32_t w[8];
                ld_global_nc_256(rp + 2 * j, w);
                csv[2 * j] = make_float4(__uint_as_float(w[0]), __uint_as_float(w[1]), __uint_as_float(w[2]), __uint_as_float(w[3]));
                csv[2 * j + 1] = make_float4(__uint_as_float(w[4]), __uint_as_float(w[5]), __uint_as_float(w[6]), __uint_as_float(w[7]));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) csv[j] = __ldg(rp + j);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 cs = csv[j];  // (cos0, sin0, cos1, sin1)
              const float a0 = x[4 * j + 0], a1 = x[4 * j + 1], a2 = x[4 * j + 2], a3 = x[4 * j + 3];
              x[4 * j + 0] = a0 * cs.x - a1 * cs.y;
              x[4 * j + 1] = a1 * cs.x + a0 * cs.y;
              x[4 * j + 2] = a2 * cs.z - a3 * cs.w;
              x[4 * j + 3] = a3 * cs.z + a2 * cs.w;
            }
          }
          if (row_ok) store_bf16x32(dst + c * 32, x);
        }
      } else if (sec <= 2) {  // v, or a q / k section without RMSNorm+RoPE (plain head-major projection)
        const int h = (n_h - sec * D) >> 7;
        __nv_bfloat16* base = sec == 0 ? p.q : (sec == 1 ? p.k : p.v);
        __nv_bfloat16* dst = base + ((static_cast<long long>(b) * p.heads + h) * p.L_total + pos) * 128;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          tmem_ld32(t_acc + hh * 128 + c * 32, r);
          load_bf16x32(p.bias + n_h + c * 32, bb);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(r[j]) + bb[j];
          if (row_ok) store_bf16x32(dst + c * 32, x);
        }
      } else {
        __nv_bfloat16* dst = p.mlp + static_cast<long long>(m) * p.ldmlp + (n_h - 3 * D);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          tmem_ld32(t_acc + hh * 128 + c * 32, r);
          load_bf16x32(p.bias + n_h + c * 32, bb);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(r[j]) + bb[j];
          if (p.mlp_pre != nullptr && row_ok)
            store_bf16x32(p.mlp_pre + static_cast<long long>(m) * p.ldmlp_pre + (n_h - 3 * D) + c * 32, x);
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = gelu_tanh_f(x[j]);
          if (row_ok) store_bf16x32(dst + c * 32, x);
        }
      }
    }
  } else if constexpr (EPI == EPI_SWIGLU) {
#pragma unroll 1
    for (int c = 0; c < (BN == 256 ? 4 : 0); ++c) {  // only launched with 256-column tiles: [gate(128) | up(128)]
      if (n_tile0 >= p.N) break;
      uint32_t rg[32], ru[32];
      float g[32], u[32];
      tmem_ld32(t_acc + c * 32, rg);
      tmem_ld32(t_acc + (BN == 256 ? 128 : 0) + c * 32, ru);
      if (p.bias != nullptr) {
        load_bf16x32(p.bias + n_tile0 + c * 32, g);
        load_bf16x32(p.bias + n_tile0 + 128 + c * 32, u);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) g[j] = u[j] = 0.f;
      }
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float gv = g[j] + __uint_as_float(rg[j]);
        g[j] = (p.aux_act == 1 ? gelu_tanh_f(gv) : silu_f(gv)) * (u[j] + __uint_as_float(ru[j]));  // GEGLU (T5 gated-gelu) | SwiGLU
      }
      if (row_ok) store_bf16x32(p.C + static_cast<long long>(m) * p.ldc + (n_tile0 >> 1) + c * 32, g);
    }
  } else {
    const __nv_bfloat16* gate_row = nullptr;
    if constexpr (EPI == EPI_GATE_RESIDUAL) {
      const int b = row_ok ? m / p.rows_per_batch : 0;
      if (p.gate != nullptr) gate_row = p.gate + static_cast<long long>(b) * p.gate_stride;
    }
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int n0 = n_tile0 + c * 32;
      if (n0 >= p.N) break;
      uint32_t r[32];
      float x[32];
      tmem_ld32(t_acc + c * 32, r);
      if (p.bias != nullptr) {
        load_bf16x32(p.bias + bias_off + n0, x);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = 0.f;
      }
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] += __uint_as_float(r[j]);
      if constexpr (EPI == EPI_BIAS_GELU_TANH) {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = gelu_tanh_f(x[j]);
      } else if constexpr (EPI == EPI_BIAS_GELU_ERF) {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = gelu_erf_f(x[j]);
      }
      if (row_ok) {
        if constexpr (EPI == EPI_BIAS) {
          if (p.c32 != nullptr) {
            float4* d4 = reinterpret_cast<float4*>(p.c32 + static_cast<long long>(m) * p.ldc32 + n0);
#pragma unroll
            for (int j = 0; j < 8; ++j) d4[j] = make_float4(x[4 * j] * p.alpha, x[4 * j + 1] * p.alpha, x[4 * j + 2] * p.alpha, x[4 * j + 3] * p.alpha);
            continue;
          }
          if (p.aux != nullptr) {  // second output: gelu_erf(C) (projector: MLP3.fc consumes GELU(x2), utils/proj.py:31)
            float g[32];
            if (p.aux_act == 1) {
#pragma unroll
              for (int j = 0; j < 32; ++j) g[j] = gelu_tanh_f(x[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) g[j] = gelu_erf_f(x[j]);
            }
            store_bf16x32(p.aux + static_cast<long long>(m) * p.ldaux + n0, g);
          }
        }
        if constexpr (EPI == EPI_DACT) {
          if (p.pre != nullptr && n0 >= p.n_split) {
            float pr[32];
            load_bf16x32(p.pre + static_cast<long long>(m) * p.ldpre + (n0 - p.n_split), pr);
            if (p.dact == 1) {
#pragma unroll
              for (int j = 0; j < 32; ++j) x[j] *= dgelu_tanh_f(pr[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) x[j] *= dgelu_erf_f(pr[j]);
            }
          }
          if (p.residual != nullptr) {
            float res[32];
            load_bf16x32(p.residual + static_cast<long long>(m) * p.ldr + n0, res);
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] += res[j];
          }
        }
        if constexpr (EPI == EPI_CONV) {
          if (p.rowvec != nullptr) {
            float rv[32];
            load_bf16x32(p.rowvec + static_cast<long long>(m / p.rows_per_batch) * p.rowvec_stride + n0, rv);
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] += rv[j];
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
          }
          if (p.residual != nullptr) {
            float res[32];
            load_bf16x32(p.residual + static_cast<long long>(m) * p.ldr + n0, res);
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] += res[j];
          }
        }
        if constexpr (EPI == EPI_GATE_RESIDUAL) {
          if (p.aux != nullptr) store_bf16x32(p.aux + static_cast<long long>(m) * p.ldaux + n0, x);
          float g[32], res[32];
          if (gate_row != nullptr) {
            load_bf16x32(gate_row + n0, g);
          } else {  // plain residual connection (decoder LM layers): gate == 1
#pragma unroll
            for (int j = 0; j < 32; ++j) g[j] = 1.0f;
          }
          load_bf16x32(p.residual + static_cast<long long>(m) * p.ldr + n0, res);
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = res[j] + g[j] * x[j];
        }
#ifdef X2I_EPI_STAGE
        if (stage == nullptr) store_bf16x32(p.C + static_cast<long long>(m) * p.ldc + n0, x);
#else
        store_bf16x32(p.C + static_cast<long long>(m) * p.ldc + n0, x);
#endif
      }
#ifdef X2I_EPI_STAGE
      // Experiment (opt-in build, profiles/r02_gemm_epilogue_probe.md): the warp's 32 rows x 64 columns go through a 4 KB shared-memory
      // transpose (16-byte units, unit' = unit ^ (row & 7): conflict-free both ways) so that each 256-bit store instruction writes 8 rows x
      // one full 128-byte line instead of 32 rows x one 32-byte sector -- 4x fewer requests for the same bytes.
      if (stage != nullptr) {
        const int lane = threadIdx.x & 31;
        const uint32_t sbase = smem_u32(stage);
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(x[2 * i], x[2 * i + 1]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + (lane * 8 + (((c & 1) * 4 + u) ^ (lane & 7))) * 16), "r"(w[4 * u]),
                       "r"(w[4 * u + 1]), "r"(w[4 * u + 2]), "r"(w[4 * u + 3])
                       : "memory");
        if (c & 1) {
          __syncwarp();
          const int m0 = m - lane;                 // first row of this warp
          const int ncol = n0 - 32;                // first column of the 64-column pair
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int r = it * 8 + (lane >> 2), q = lane & 3;
            uint4 a, b;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(sbase + (r * 8 + ((2 * q) ^ (r & 7))) * 16) : "memory");
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(sbase + (r * 8 + ((2 * q + 1) ^ (r & 7))) * 16) : "memory");
            if (m0 + r < p.M)
              st_global_256(p.C + static_cast<long long>(m0 + r) * p.ldc + ncol + q * 16, a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w);
          }
          __syncwarp();
        }
      }
#endif
    }
  }
}

// B_CONV = 1 (stride 1: 4-D map [C, W, H, N]) / 2 (stride 2: 5-D parity view [2C, W/2, 2, H/2, N]): with A_MN + B_MN this is the weight
// gradient of a convolution, dW[Cout, (ky, kx, Cin)] = sum over output pixels of dY[pix, Cout] x X[pix * stride + tap - pad, Cin];
// a k-block is 64 consecutive output pixels (whole rows of one image: the host checks Wo % 64 == 0, or 64 % Wo == 0 with
// Ho * Wo % 64 == 0 and a box of 64 / Wo rows), the column block selects (tap, 64 input channels).
template <int BN, int EPI, bool B_MN, bool A_MN = false, int B_CONV = 0>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int NS = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NS * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + NS;
  uint64_t* tfull_bar = empty_bar + NS;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = uniform_warp_id();
  const int lane = threadIdx.x & 31;
  const int num_m = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int num_n = (p.N + BN - 1) / BN;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int ksplit = p.ksplit > 1 ? p.ksplit : 1;
  const int kb_per = (num_kb + ksplit - 1) / ksplit;  // host guarantees every group is non-empty
  const int num_mn = num_m * num_n;
  const int num_tiles = num_mn * ksplit;

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
  griddep_launch();  // PDL (common.cuh): prologue above overlaps the predecessor's tail; no global access before the wait
  griddep_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int ks = tile / num_mn, mn = tile - ks * num_mn;
        const int m_blk = mn % num_m, n_blk = mn / num_m;
        const int kb0 = ks * kb_per, kb1 = min(kb0 + kb_per, num_kb);
        // B_CONV: per tile, the (tap, channel block) of each 64-column group and the first output pixel of the k range; the single
        // producer thread must not divide per k-block (a k-block is only 192-256 tensor-pipe cycles)
        [[maybe_unused]] int cv_c[BN / 64], cv_dx[BN / 64], cv_dy[BN / 64], cv_py[BN / 64];
        [[maybe_unused]] int cv_img = 0, cv_yo = 0, cv_xo = 0, cv_rows = 1, cv_ho = 1;
        if constexpr (B_CONV != 0) {
#pragma unroll
          for (int g = 0; g < BN / 64; ++g) {
            const int col = n_blk * BN + g * 64;
            const int tap = col / p.cv_cin, c0 = col - tap * p.cv_cin;
            const int ky = tap / p.cv_kw, kx = tap - ky * p.cv_kw;
            if constexpr (B_CONV == 1) {
              cv_c[g] = c0; cv_dx[g] = kx - p.cv_pad; cv_dy[g] = ky - p.cv_pad; cv_py[g] = 0;
            } else {  // input pixel = 2 * output pixel + t, t in {-1, 0, 1}: parity t & 1, half-resolution offset (t - parity) / 2
              const int tx = kx - p.cv_pad, ty = ky - p.cv_pad;
              const int px = tx & 1, py = ty & 1;
              cv_c[g] = px * p.cv_cin + c0; cv_dx[g] = (tx - px) >> 1; cv_dy[g] = (ty - py) >> 1; cv_py[g] = py;
            }
          }
          const int k0 = kb0 * GEMM_BK;
          cv_img = k0 / p.cv_howo;
          const int rem = k0 - cv_img * p.cv_howo;
          cv_yo = rem / p.cv_wo;
          cv_xo = rem - cv_yo * p.cv_wo;
          cv_rows = p.cv_wo >= GEMM_BK ? 1 : GEMM_BK / p.cv_wo;
          cv_ho = p.cv_howo / p.cv_wo;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if constexpr (!A_MN) {
            tma_load_2d(sa, &tma_a, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
          } else {
#pragma unroll
            for (int g = 0; g < GEMM_BM / 64; ++g)  // A stored [K, M]: boxes of 64 (m) x 64 (k rows)
              tma_load_2d(sa + g * 8192, &tma_a, &full_bar[stage], m_blk * GEMM_BM + g * 64, kb * GEMM_BK);
          }
          if constexpr (B_CONV != 0) {
#pragma unroll
            for (int g = 0; g < BN / 64; ++g) {
              if constexpr (B_CONV == 1) tma_load_4d(sb + g * 8192, &tma_b, &full_bar[stage], cv_c[g], cv_xo + cv_dx[g], cv_yo + cv_dy[g], cv_img);
              else tma_load_5d(sb + g * 8192, &tma_b, &full_bar[stage], cv_c[g], cv_xo + cv_dx[g], cv_py[g], cv_yo + cv_dy[g], cv_img);
            }
            cv_xo += GEMM_BK;  // next 64 output pixels: whole rows of one image (host-checked geometry), no division on this path
            if (cv_xo >= p.cv_wo) {
              cv_xo = 0;
              cv_yo += cv_rows;
              if (cv_yo >= cv_ho) { cv_yo = 0; ++cv_img; }
            }
          } else if constexpr (!B_MN) {
            tma_load_2d(sb, &tma_b, &full_bar[stage], kb * GEMM_BK, n_blk * BN);
          } else {
#pragma unroll
            for (int g = 0; g < BN / 64; ++g)  // B stored [K, N]: boxes of 64 (n) x 64 (k rows)
              tma_load_2d(sb + g * 8192, &tma_b, &full_bar[stage], n_blk * BN + g * 64, kb * GEMM_BK);
          }
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (whole warp, elected lane issues)
    {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        const int ks = tile / num_mn;
        const int kb0 = ks * kb_per, kb1 = min(kb0 + kb_per, num_kb);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_base = a_base + Cfg::A_BYTES;
          const uint64_t adesc = A_MN ? make_smem_desc_sw128(a_base, 8192, 1024) : make_smem_desc_sw128(a_base, 16, 1024);
          const uint64_t bdesc = B_MN ? make_smem_desc_sw128(b_base, 8192, 1024) : make_smem_desc_sw128(b_base, 16, 1024);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            umma_ss_w(d_tmem, adesc + ((A_MN ? k * 2048 : k * 32) >> 4), bdesc + ((B_MN ? k * 2048 : k * 32) >> 4), idesc,
                      (kb != kb0 || k != 0) ? 1u : 0u);
          umma_commit_w(&empty_bar[stage]);
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        umma_commit_w(&tfull_bar[as]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (2..5)
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int ks = tile / num_mn, mn = tile - ks * num_mn;
      const int m_blk = mn % num_m, n_blk = mn / num_m;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + as * BN + lane_off;
      const int m = m_blk * GEMM_BM + quad * 32 + lane;
      gemm_epilogue_tile<BN, EPI>(p, t_acc, m, n_blk * BN, 0, ks);
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

// out[m, n] (+)= sum_s part[s, m, n] in the fixed order s = 0, 1, ...; 8 columns per thread.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, __nv_bfloat16* __restrict__ out, long long ldo, int M,
                                                            int N, int S, int accumulate) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int n8 = N / 8;
  if (i >= static_cast<long long>(M) * n8) return;
  const int m = static_cast<int>(i / n8), c = static_cast<int>(i % n8) * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int s = 0; s < S; ++s) {
    const float4* p4 = reinterpret_cast<const float4*>(part + (static_cast<long long>(s) * M + m) * N + c);
    const float4 a = p4[0], b = p4[1];
    acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
    acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
  }
  uint4* o = reinterpret_cast<uint4*>(out + static_cast<long long>(m) * ldo + c);
  if (accumulate) {
    const uint4 u = *o;
    acc[0] += bf16_lo(u.x); acc[1] += bf16_hi(u.x); acc[2] += bf16_lo(u.y); acc[3] += bf16_hi(u.y);
    acc[4] += bf16_lo(u.z); acc[5] += bf16_hi(u.z); acc[6] += bf16_lo(u.w); acc[7] += bf16_hi(u.w);
  }
  uint4 r;
  r.x = pack_bf16x2(acc[0], acc[1]); r.y = pack_bf16x2(acc[2], acc[3]); r.z = pack_bf16x2(acc[4], acc[5]); r.w = pack_bf16x2(acc[6], acc[7]);
  *o = r;
}

}  // namespace x2i
