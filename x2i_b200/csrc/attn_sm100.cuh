// Fused MMDiT joint attention forward for sm_100a (head_dim 128, non-causal, no mask):
//     O = softmax(Q K^T / sqrt(128)) V      Q,K,V: [B, H, L, 128] bf16 (text tokens first, then image tokens)
// replaces F.scaled_dot_product_attention inside FluxAttnProcessor2_0.
//
// One CTA per (batch, head, 256 query rows).  320 threads:
//   warp 0      TMA producer: Q (2 tiles of 128 rows) once, then K_j / V_j tiles of 128 keys through a 4-slot ring
//   warp 1      tcgen05.mma issuer (one lane): S_i = Q_i K_j^T (SS, both K-major) and O_i += P_i V_j (A = P from TMEM,
//               B = V as MN-major smem operand), ping-ponging the two query tiles so that the softmax of one tile
//               overlaps the MMAs of the other
//   warps 2..5  softmax warpgroup for query tile 0  (one query row per thread)
//   warps 6..9  softmax warpgroup for query tile 1
// TMEM (512 columns): S0 | S1 | O0 | O1, 128 fp32 columns each; P_i (bf16) aliases the first 64 columns of S_i.
// Online softmax in the exp2 domain with lazy rescaling of O (only when the running max grows by > 2^8).
#pragma once
#include "common.cuh"

namespace x2i {

struct AttnParams {
  int B, H, L;   // L = number of query rows per (batch, head)
  int Lkv;       // number of key/value rows per (batch, head) (== L for self-attention)
  const int* kv_len;  // optional device array [B]: valid keys of batch b (key-padding mask); null = Lkv
  float scale_log2;  // log2(e) / sqrt(head_dim)
  __nv_bfloat16* out0;  // rows with pos <  split: out0[(b*split + pos) * ld0 + h*128 + d]
  long long ld0;
  int split;
  __nv_bfloat16* out1;  // rows with pos >= split: out1[(b*(L-split) + pos-split) * ld1 + h*128 + d]
  long long ld1;
  float* lse;  // optional (training): lse[(b*H + h) * Lpad + pos] = log2-domain log-sum-exp of row pos (+inf for pos >= L)
  int Lpad;    // row stride of lse (L rounded up to 128)
  // decoder-LM prefill form (MLLM text encoders, x2i_causal_attention): causal mask, grouped-query K/V heads, left padding
  int causal;           // 1: key j is visible to query row i only if j <= i
  int Hkv;              // number of K/V heads (0 or == H: one per query head); query head h reads K/V head h / (H / Hkv)
  const int* kv_start;  // optional device array [B]: first valid key of batch b (left padding); rows with no visible key output 0
  long long* trace;  // diagnostics only (DBG == 1 instantiation): clock64 stamps of CTA (0,0,0)
};
#define ATT_STAMP(slot) do { if constexpr (DBG == 1) { if (trace != nullptr && lane == 0) trace[slot] = clock64(); } } while (0)

constexpr int ATT_THREADS = 320;
#ifndef X2I_ATT_POLY8  // tools/jobs/gpu_job_r04i.sh builds alternative libraries with 1 / 3 (sweep of the lagged form)
#define X2I_ATT_POLY8 2
#endif
constexpr int ATT_DEFAULT_POLY8 = X2I_ATT_POLY8;  // production instantiation (see capi.cu)
constexpr int ATT_TILE_BYTES = 128 * 128 * 2;  // 32 KB: two 128B-swizzled boxes of [128 rows x 64 cols]
constexpr int ATT_KV_SLOTS = 4;
constexpr int ATT_SMEM_BYTES = 2 * ATT_TILE_BYTES + ATT_KV_SLOTS * ATT_TILE_BYTES + 1024 + 256;

// exp2(x * sc - m) of one half (HALF = 0 / 1) of a 32-score quarter -> 8 of its 16 packed bf16x2 words + packed partial sums
// GUARD (lagged-max steps): pguard collects the max ARGUMENT of the polynomial lanes -- above 127 their exponent insertion wraps around
// silently, where MUFU.EX2 returns +inf (which the row sum shows).
template <int POLY8, int DBG, int HALF, bool GUARD = false>
__device__ __forceinline__ void exp_half(const uint32_t (&x)[32], uint64_t sc2, uint64_t nm2, uint64_t (&sum2)[4], uint32_t (&pk)[16],
                                         float* pguard = nullptr) {
#pragma unroll
  for (int k = 8 * HALF; k < 8 * HALF + 8; ++k) {
    const uint64_t a2 = fma_f32x2(pack_f32x2(__uint_as_float(x[2 * k]), __uint_as_float(x[2 * k + 1])), sc2, nm2);
    float a0, a1, p0, p1;
    unpack_f32x2(a2, a0, a1);
    if constexpr (DBG == 2) {  // timing experiment: no exponentials at all
      p0 = a0;
      p1 = a1;
    } else if ((k & 7) < POLY8) {  // POLY8 of every 8 pairs: exp2 on the FMA pipe instead of MUFU
      if constexpr (GUARD) *pguard = max3f(*pguard, a0, a1);
      poly_exp2_x2(a0, a1, p0, p1);
    } else {
      p0 = fast_exp2(a0);
      p1 = fast_exp2(a1);
    }
    sum2[k & 3] = add_f32x2(sum2[k & 3], pack_f32x2(p0, p1));
    pk[k] = pack_bf16x2(p0, p1);
  }
}

// One online-softmax step of one query row over a 128-key tile (executed by the 128 threads of a softmax warpgroup).
// MASKED is instantiated only for the last, ragged key tile so the common path carries no masking instructions.
// PAIR: the P hand-off barriers live in the leader CTA of a cta_group::2 pair (attn2_sm100.cuh) and are arrived on through mapa.
__device__ __forceinline__ void mbar_arrive_rank0(uint64_t* bar_local) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar_local))
      : "memory");
}
template <int POLY8, bool MASKED, int DBG, bool PAIR = false>
__device__ __forceinline__ void softmax_step_core(uint32_t t_s, uint32_t t_o, uint64_t* s_full_i, uint64_t* p_full_i, uint32_t parity,
                                                  bool first, int valid, float sc, float& m_run, float& l_run, int lane, long long* trace,
                                                  int lo) {
  // MASKED: keys with tile index outside [lo, valid) are excluded (ragged tail, left padding, causal diagonal).  parity: phase of
  // s_full (one completion per key step of the CTA); first: first key step of this query row (no running max yet).
  constexpr int P8 = MASKED ? 0 : POLY8;  // masked tiles keep every exponential on MUFU: ex2(-inf) is exactly +0 (the polynomial
                                          // clamps at 2^-125), so a row with no visible key ends with l == 0 and outputs 0
  ATT_STAMP(0);
  if constexpr (DBG == 4) {
    while (!mbar_test(s_full_i, parity)) {
    }
  } else {
    mbar_wait(s_full_i, parity);
  }
  tc_fence_after();
  ATT_STAMP(1);
  uint32_t r[4][32];
#pragma unroll
  for (int c = 0; c < 4; ++c) tmem_ld32(t_s + c * 32, r[c]);
  tmem_ld_wait();
  ATT_STAMP(2);
  if constexpr (MASKED) {  // keys >= valid were zero-filled by TMA: exclude them
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (c * 32 + k >= valid || c * 32 + k < lo) r[c][k] = 0xff800000u;  // -inf
  }
  float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // 4 independent FMNMX3 chains
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < 32; k += 2)
      mx4[(k >> 1) & 3] = max3f(mx4[(k >> 1) & 3], __uint_as_float(r[c][k]), __uint_as_float(r[c][k + 1]));
  float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
  if constexpr (DBG == 3) mx = 4.0f / sc;  // timing experiment: no max phase
  const float m_cand = fmaxf(m_run, mx * sc);
  float alpha = 1.0f;
  bool rescale = false;
  if (first) {
    m_run = m_cand;
  } else if (m_cand - m_run > 8.0f) {  // lazy rescaling: tolerate a stale max up to 2^8
    alpha = fast_exp2(m_run - m_cand);
    m_run = m_cand;
    rescale = true;
  }
  // O_i may be touched here without waiting on a barrier: S_i(j) was issued after PV_i(j-1) and tcgen05.commit
  // tracks completion of ALL earlier MMAs of the issuing thread, so s_full(j) implies PV_i(j-1) has retired.
  if (!first && __any_sync(0xffffffffu, rescale)) {
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      tmem_ld32(t_o + c * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
      tmem_st32(t_o + c * 32, o);
    }
  }
  ATT_STAMP(3);
  const float m_eff = (MASKED && m_run == -INFINITY) ? 0.f : m_run;  // no visible key so far: avoid (-inf) - (-inf)
  const uint64_t sc2 = pack_f32x2(sc, sc), nm2 = pack_f32x2(-m_eff, -m_eff);
  uint64_t sum2[4] = {0ull, 0ull, 0ull, 0ull};  // packed (FADD2) partial row sums, 4 independent chains
  // Second pass, quarter by quarter (32 keys each), software-pipelined by hand because the TMEM store -> wait::st -> fence ->
  // mbarrier arrive sequence costs ~120 cycles in which the MUFU pipe would idle:
  //   * quarter 0 is still in registers from the max pass; quarter c+1 is re-read from TMEM in the MIDDLE of quarter c, after
  //     the store of P quarter c-1 in program order, so ptxas cannot hoist its exponentials above that store (with all 128
  //     scores live in registers it sank every store to the end of the step and serialised the PV MMAs behind the soft-max);
  //   * the completion wait + hand-off of quarter c-1 is deferred to the middle of quarter c, when its store has long landed.
  // P quarter c (columns [16c,16c+16)) only overwrites S columns of quarters <= c/2, which have been consumed.
  uint32_t rq[2][32];
  uint32_t pk[16];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t(&x)[32] = c == 0 ? r[0] : rq[c & 1];
    if (c > 0) {
      tmem_ld_wait();
      if constexpr (MASKED) {
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (c * 32 + k >= valid || c * 32 + k < lo) rq[c & 1][k] = 0xff800000u;
      }
    }
    exp_half<P8, DBG, 0>(x, sc2, nm2, sum2, pk);
    if (c > 0) {
      tmem_st_wait();  // store of quarter c-1, issued half a quarter ago
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_rank0(&p_full_i[c - 1]); else mbar_arrive(&p_full_i[c - 1]);
      }
      ATT_STAMP(3 + c);
    }
    if (c < 3) tmem_ld32(t_s + (c + 1) * 32, rq[(c + 1) & 1]);
    exp_half<P8, DBG, 1>(x, sc2, nm2, sum2, pk);
    tmem_st16(t_s + c * 16, pk);  // P_i aliases S_i columns [0, 64)
  }
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) {
    if constexpr (PAIR) mbar_arrive_rank0(&p_full_i[3]); else mbar_arrive(&p_full_i[3]);
  }
  ATT_STAMP(7);
  float s0, s1;
  unpack_f32x2(add_f32x2(add_f32x2(sum2[0], sum2[1]), add_f32x2(sum2[2], sum2[3])), s0, s1);
  l_run = l_run * alpha + (s0 + s1);
}
// Lagged form of a key step (persistent kernel, X2I_ATTN_LAG; unmasked, not the first step of a row).  The max pass over the 128
// scores (~270-390 cycles of every ~2800-cycle step, all of it on the S -> soft-max -> PV -> S chain that bounds the kernel,
// profiles/r02_attn_clock64_trace.txt) disappears: the exponentials of step j start right away against the reference m_run the
// row already has, and the reference for step j + 1 comes from the step's own ROW SUM, which is computed anyway:
//     max_k p_k <= sum_k p_k <= 128 max_k p_k,   p_k = 2^(s_k - m_run)
// so m_run + log2(sum) bounds the tile's max from above within 7 (log2 units).  The reference need not be the true max -- O, l and
// m_run stay consistent at every step and the final O / l does not depend on it -- it only has to keep p away from overflow:
// when a step's sum exceeds 2^16 the next step first rescales O and l (the lazy rescale, one step late).
// Overflow guard: a tile that exceeds the reference by more than 2^96 (66 nats above EVERY earlier key of the row) shows as a sum
// above 2^96 or inf / nan; the polynomial lanes, whose exponent insertion would wrap silently beyond 2^127, are covered by the
// max of their arguments (16 FMNMX3 per step).  Either sets the CTA's redo flag and the kernel re-runs that CTA's items with
// the classic step (attn_persist_sm100.cuh).
// What did NOT work (profiles/r02_attn_probe_lagged.md): comparing each quarter's max before its exponentials (exact p <= 2^8
// bound, one uniform branch per quarter) was 4 % slower than the classic step -- ptxas cannot overlap a quarter's pack / store tail
// with the next quarter's exponentials across a basic-block boundary; forming the tile's true max in the shadow of the MUFU stream
// (64 FMNMX3 per step) kept only +2.7 % of the +6.8 % that the missing max pass is worth.
constexpr float ATT_LAG_RESCALE = 16.0f;  // log2 of the step sum above which the next step rescales
constexpr float ATT_LAG_LIMIT = 96.0f;    // log2 of the step sum above which the step counts as overflowed (p <= 2^96: P.V partial sums stay below
                                          // 2^96 x 128 keys x |v| -- far from the fp32 range for any bf16 V a network produces)
template <int POLY8>
__device__ __forceinline__ void softmax_step_lagged(uint32_t t_s, uint32_t t_o, uint64_t* s_full_i, uint64_t* p_full_i, uint32_t parity,
                                                    float sc, float& m_run, float& l_run, float& m_next, int lane, int* redo_flag) {
  // lazy rescale against the reference the PREVIOUS step proposed (m_next); alpha does not depend on this step's scores
  const bool rescale = m_next - m_run > ATT_LAG_RESCALE;
  const float alpha = rescale ? fast_exp2(m_run - m_next) : 1.0f;
  if (rescale) m_run = m_next;
  mbar_wait(s_full_i, parity);
  tc_fence_after();
  uint32_t rq[2][32];
  uint32_t pk[16];
  tmem_ld32(t_s, rq[0]);
  // O_i may be touched without a barrier: s_full(j) implies P.V_i(j-1) has retired (see softmax_step_core)
  if (__any_sync(0xffffffffu, rescale)) {
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      tmem_ld32(t_o + c * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
      tmem_st32(t_o + c * 32, o);
    }
  }
  const uint64_t sc2 = pack_f32x2(sc, sc), nm2 = pack_f32x2(-m_run, -m_run);
  uint64_t sum2[4] = {0ull, 0ull, 0ull, 0ull};
  float pguard = -INFINITY;
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t(&x)[32] = rq[c & 1];
    if (c < 3) tmem_ld32(t_s + (c + 1) * 32, rq[(c + 1) & 1]);
    exp_half<POLY8, 0, 0, true>(x, sc2, nm2, sum2, pk, &pguard);
    if (c > 0) {
      tmem_st_wait();  // store of quarter c-1, issued half a quarter ago
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full_i[c - 1]);
    }
    if (c < 3) tmem_ld_wait();
    exp_half<POLY8, 0, 1, true>(x, sc2, nm2, sum2, pk, &pguard);
    tmem_st16(t_s + c * 16, pk);  // P_i aliases S_i columns [0, 64)
  }
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(&p_full_i[3]);
  float s0, s1;
  unpack_f32x2(add_f32x2(add_f32x2(sum2[0], sum2[1]), add_f32x2(sum2[2], sum2[3])), s0, s1);
  const float step_sum = s0 + s1;
  l_run = l_run * alpha + step_sum;
  const float lg = __log2f(step_sum);  // -inf for an all-underflow step: no rescale
  m_next = m_run + lg;
  if (!(lg <= ATT_LAG_LIMIT) || pguard > 126.0f) *redo_flag = 1;  // p may have overflowed: this CTA's items are re-run with the classic step
}
// one-item kernels: the key-step counter j of the CTA gives both the barrier parity and "first step"
template <int POLY8, bool MASKED, int DBG, bool PAIR = false>
__device__ __forceinline__ void softmax_step(uint32_t t_s, uint32_t t_o, uint64_t* s_full_i, uint64_t* p_full_i, int j, int valid,
                                             float sc, float& m_run, float& l_run, int lane, long long* trace, int lo = 0) {
  softmax_step_core<POLY8, MASKED, DBG, PAIR>(t_s, t_o, s_full_i, p_full_i, static_cast<uint32_t>(j & 1), j == 0, valid, sc, m_run, l_run, lane,
                                              trace, lo);
}
// persistent kernel (attn_persist_sm100.cuh): parity from the CTA-wide step counter, `first` from the item-local one
template <int POLY8, bool MASKED>
__device__ __forceinline__ void softmax_step_p(uint32_t t_s, uint32_t t_o, uint64_t* s_full_i, uint64_t* p_full_i, uint32_t parity, bool first,
                                               int valid, float sc, float& m_run, float& l_run, int lane, int lo) {
  softmax_step_core<POLY8, MASKED, 0, false>(t_s, t_o, s_full_i, p_full_i, parity, first, valid, sc, m_run, l_run, lane, nullptr, lo);
}

// Epilogue of one soft-max thread (one query row): O_i / l -> bf16 -> global, token-major [.., H*128] so the out-projection GEMM
// reads it as its A operand; optional log-sum-exp for the backward.
__device__ __forceinline__ void attn_epilogue(const AttnParams& p, uint64_t* o_full_i, uint32_t t_o, int pos, int b, int h, int bh,
                                              float m_run, float l_run, uint32_t parity = 0) {
  mbar_wait(o_full_i, parity);  // committed once per work item, after the last PV MMA of this tile
  tc_fence_after();
  const float inv_l = l_run > 0.f ? 1.0f / l_run : 0.f;  // no visible key (padded query row of a causal prefill): output 0
  const bool ok = pos < p.L;
  if (p.lse != nullptr && pos < p.Lpad)
    p.lse[static_cast<long long>(bh) * p.Lpad + pos] = ok ? m_run + log2f(l_run) : INFINITY;
  __nv_bfloat16* dst;
  if (pos < p.split)
    dst = p.out0 + (static_cast<long long>(b) * p.split + pos) * p.ld0 + h * 128;
  else
    dst = p.out1 + (static_cast<long long>(b) * (p.L - p.split) + (pos - p.split)) * p.ld1 + h * 128;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t o[32];
    tmem_ld32(t_o + c * 32, o);
    tmem_ld_wait();
    if (ok) st_bf16x32_scaled(dst + c * 32, o, inv_l);
  }
}

// POLY8: how many of every 8 element PAIRS of the softmax use the FMA-pipe poly_exp2_x2 instead of MUFU.EX2.
// DBG: 0 product; 1 clock64 trace; 2 / 3 timing experiments (wrong results); 4 non-suspending barrier polls (same results).
// LM: decoder-LM prefill form (causal mask, left padding, grouped-query K/V heads).  A separate instantiation: the production MMDiT
// instantiation sits at the 168-register limit of a 320-thread CTA, and the extra live state of the general mask spills it.
template <int POLY8, int DBG = 0, bool LM = false>
__global__ void __launch_bounds__(ATT_THREADS, 1)
mmdit_attention_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                           const __grid_constant__ CUtensorMap tma_v, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sq = smem;                         // Q0 | Q1
  uint8_t* skv = smem + 2 * ATT_TILE_BYTES;   // ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(skv + ATT_KV_SLOTS * ATT_TILE_BYTES);
  uint64_t* q_full = bars;              // 1
  uint64_t* kv_full = bars + 1;         // 4
  uint64_t* kv_empty = bars + 5;        // 4
  uint64_t* s_full = bars + 9;          // 2
  uint64_t* p_full = bars + 11;         // 2 tiles x 4 quarters (32 keys each): PV MMAs start as soon as a quarter of P lands
  uint64_t* o_full = bars + 19;         // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

  const int warp = uniform_warp_id();
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int bh = b * p.H + h;
  const int kv_valid = p.kv_len != nullptr ? min(max(p.kv_len[b], 1), p.Lkv) : p.Lkv;
  // key-tile window [j0, j0 + n_kv) of this CTA: all tiles, minus the tiles left of the first valid key (left padding) and right of
  // the causal diagonal of its last query row.  At least one tile is always processed (a CTA of padded rows outputs zeros).
  const int kv_lo = (LM && p.kv_start != nullptr) ? min(max(p.kv_start[b], 0), kv_valid - 1) : 0;
  const int j0 = LM ? (kv_lo >> 7) : 0;
  int j_end = (kv_valid + 127) / 128;
  if (LM && p.causal) j_end = min(j_end, (min(q0 + 255, p.L - 1) >> 7) + 1);
  const int n_kv = LM ? max(j_end - j0, 1) : j_end;
  const int bh_kv = (LM && p.Hkv > 0 && p.Hkv != p.H) ? b * p.Hkv + h / (p.H / p.Hkv) : bh;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < ATT_KV_SLOTS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      for (int c = 0; c < 4; ++c) mbar_init(&p_full[i * 4 + c], 4);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * ATT_TILE_BYTES);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int g = 0; g < 2; ++g)
          tma_load_3d(sq + i * ATT_TILE_BYTES + g * 16384, &tma_q, q_full, g * 64, q0 + i * 128, bh);
      for (int seq = 0; seq < 2 * n_kv; ++seq) {
        const int slot = seq & (ATT_KV_SLOTS - 1);
        const uint32_t ph = (seq / ATT_KV_SLOTS) & 1;
        mbar_wait(&kv_empty[slot], ph ^ 1);
        mbar_expect_tx(&kv_full[slot], ATT_TILE_BYTES);
        const CUtensorMap* map = (seq & 1) ? &tma_v : &tma_k;
        const int j = j0 + (seq >> 1);
        uint8_t* dst = skv + slot * ATT_TILE_BYTES;
        tma_load_3d(dst, map, &kv_full[slot], 0, j * 128, bh_kv);
        tma_load_3d(dst + 16384, map, &kv_full[slot], 64, j * 128, bh_kv);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (whole warp, elected lane issues)
    {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 128, 0, 1);
      const uint32_t q_base = smem_u32(sq);
      const uint32_t kv_base = smem_u32(skv);
      // descriptors are built once per tile; per MMA only the start-address field moves (a 64-bit add of a constant)
      const uint64_t qdesc = make_smem_desc_sw128(q_base, 16, 1024);
      auto issue_s = [&](int i, int slot) {  // S_i = Q_i K^T
        const uint64_t ad = qdesc + ((i * ATT_TILE_BYTES) >> 4);
        const uint64_t bd = make_smem_desc_sw128(kv_base + slot * ATT_TILE_BYTES, 16, 1024);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = ((kk >> 2) * 16384 + (kk & 3) * 32) >> 4;
          umma_ss_w(tmem_base + i * 128, ad + off, bd + off, idesc_s, kk != 0);
        }
      };
      auto issue_pv = [&](int i, int slot, bool acc, uint32_t ph) {  // O_i += P_i V, quarter by quarter as P lands
        const uint64_t bd = make_smem_desc_sw128(kv_base + slot * ATT_TILE_BYTES, 16384, 1024);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if constexpr (DBG == 4) {
            while (!mbar_test(&p_full[i * 4 + c], ph)) {
            }
          } else {
            mbar_wait(&p_full[i * 4 + c], ph);
          }
          tc_fence_after();
#pragma unroll
          for (int kk = 2 * c; kk < 2 * c + 2; ++kk)
            umma_ts_w(tmem_base + 256 + i * 128, tmem_base + i * 128 + kk * 8, bd + ((kk * 2048) >> 4), idesc_o, (acc || kk != 0) ? 1u : 0u);
        }
      };
      auto kv_wait = [&](int seq) { mbar_wait(&kv_full[seq & (ATT_KV_SLOTS - 1)], (seq / ATT_KV_SLOTS) & 1); };

      mbar_wait(q_full, 0);
      kv_wait(0);
      tc_fence_after();
      issue_s(0, 0);
      umma_commit_w(&s_full[0]);
      issue_s(1, 0);
      umma_commit_w(&s_full[1]);
      umma_commit_w(&kv_empty[0]);
      long long* const trace_base = (DBG == 1 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? p.trace : nullptr;
      for (int j = 0; j < n_kv; ++j) {
        long long* const trace = (trace_base != nullptr && j < 16) ? trace_base + 512 + j * 8 : nullptr;
        const int vseq = 2 * j + 1, kseq = 2 * j + 2;
        const int vslot = vseq & (ATT_KV_SLOTS - 1), kslot = kseq & (ATT_KV_SLOTS - 1);
        const bool more = (j + 1 < n_kv);
        ATT_STAMP(0);
        kv_wait(vseq);
        ATT_STAMP(1);
        issue_pv(0, vslot, j > 0, j & 1);
        ATT_STAMP(2);
        if (!more) umma_commit_w(&o_full[0]);
        if (more) {
          kv_wait(kseq);
          tc_fence_after();
          issue_s(0, kslot);
          umma_commit_w(&s_full[0]);
        }
        ATT_STAMP(3);
        issue_pv(1, vslot, j > 0, j & 1);
        ATT_STAMP(4);
        if (!more) umma_commit_w(&o_full[1]);
        umma_commit_w(&kv_empty[vslot]);
        if (more) {
          issue_s(1, kslot);
          umma_commit_w(&s_full[1]);
          umma_commit_w(&kv_empty[kslot]);
        }
        ATT_STAMP(5);
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax warpgroups
    const int i = (warp - 2) >> 2;  // query tile 0 / 1
    const int quad = warp & 3;      // TMEM lane quadrant accessible to this warp
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_s = tmem_base + i * 128 + lane_off;
    const uint32_t t_o = tmem_base + 256 + i * 128 + lane_off;
    const int pos = q0 + i * 128 + quad * 32 + lane;
    float m_run = -INFINITY;  // running max, log2 domain (already multiplied by scale_log2)
    float l_run = 0.f;
    const float sc = p.scale_log2;

    const bool traced = DBG == 1 && quad == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && p.trace != nullptr;
    if constexpr (!LM) {
      // MMDiT joint attention / cross attention (the hot path): every key tile but a ragged last one is unmasked -- a tight loop over
      // the mask-free instantiation (a per-step mask decision in this loop costs ~6 % of the kernel, profiles/r02_attn_probe.md)
      const bool ragged = (kv_valid & 127) != 0;
      const int n_full = ragged ? n_kv - 1 : n_kv;
      for (int j = 0; j < n_full; ++j)
        softmax_step<POLY8, false, DBG>(t_s, t_o, &s_full[i], &p_full[i * 4], j, 128, sc, m_run, l_run, lane,
                                        (traced && j < 16) ? p.trace + (i * 16 + j) * 8 : nullptr);
      if (ragged)
        softmax_step<POLY8, true, DBG>(t_s, t_o, &s_full[i], &p_full[i * 4], n_kv - 1, kv_valid - (n_kv - 1) * 128, sc, m_run, l_run, lane,
                                       nullptr);
    } else {
      // decoder-LM prefill: left padding and the causal diagonal make the mask a per-row, per-tile property
      for (int it = 0; it < n_kv; ++it) {
        const int j = j0 + it;
        int hi = min(128, kv_valid - j * 128);                 // keys of this tile with index in [lo, hi) are visible to this row
        const int lo = max(0, kv_lo - j * 128);
        if (p.causal) hi = min(hi, pos - j * 128 + 1);
        if (__any_sync(0xffffffffu, hi < 128 || lo > 0))
          softmax_step<POLY8, true, DBG>(t_s, t_o, &s_full[i], &p_full[i * 4], it, hi, sc, m_run, l_run, lane, nullptr, lo);
        else
          softmax_step<POLY8, false, DBG>(t_s, t_o, &s_full[i], &p_full[i * 4], it, 128, sc, m_run, l_run, lane, nullptr);
      }
    }
    attn_epilogue(p, &o_full[i], t_o, pos, b, h, bh, m_run, l_run);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace x2i
