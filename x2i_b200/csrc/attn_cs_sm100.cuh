// Column-split form of the persistent fused attention forward (attn_persist_sm100.cuh), MMDiT case (no causal mask, no key padding).
//
// In the persistent kernel soft-max warpgroup i owns query tile i: one thread walks all 128 scores of its row, and the
// S_i(j) -> soft-max -> P.V_i(j) -> S_i(j+1) chain of a tile (512 + ~1870 + hand-off cycles per 128 keys) IS the period of the key loop:
// ~2790 cycles against 2048 of tensor work, with each soft-max warp alone on its sub-partition's MUFU pipe at ~70 % of its rate
// (profiles/r02_attn_clock64_trace.txt, tools/ubench/exp_quarter.cu).  Here BOTH warpgroups work on the SAME tile: warpgroup h owns the
// 64-key half h of every score tile (thread = query row, as before), so a tile's soft-max takes half as long, the two tiles are processed
// one after the other, and while the eight warps are busy with tile 1 the tensor pipe has P.V_0(j) and S_0(j+1) to itself: the chain of a
// tile (S 512 + soft-max ~1000) now fits inside the period (2 x ~1000 of soft-max = the 2048 cycles of tensor work).
//
// What the split costs: the row max is the max of two threads' halves -- exchanged through shared memory behind a 64-thread named
// barrier per tile and key step (the two warps that own the same TMEM lane quadrant); that barrier also orders "both halves of S have
// been read" before either warpgroup's P overwrites S columns.  P half h lives in the first 32 columns of ITS OWN 64 S columns (so a
// warpgroup never writes where the other still reads), the P.V products take their A operand from there.  The rare O rescale (lazy, when
// the running max grows by more than 2^8) is split by column halves and needs a second barrier on that path only.  Row sums are kept
// per half and combined once per item.
#pragma once
#include "attn_persist_sm100.cuh"

namespace x2i {

constexpr int ATTCS_XCH_BYTES = 2 * 2 * 128 * 4 + 2 * 128 * 4;  // max exchange (2 slots x 2 halves x 128 rows) + row-sum exchange (2 tiles)
constexpr int ATTCS_SMEM_BYTES = ATT_SMEM_BYTES + 4096;

// keeps ptxas from hoisting the exponentials of the second quarter above the tcgen05.st of the first (asm volatile statements keep
// their program order; the values cannot be used before this statement)
__device__ __forceinline__ void reg_fence32(uint32_t (&r)[32]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                    "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                    "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                    "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// One online-softmax step of one query row over the 64-key half `h` of a 128-key tile.  t_s / t_o: TMEM addresses of the tile's S / O
// (lane quadrant applied).  xch: [2 slots][2 halves][128 rows] floats; xcnt: exchanges done by this thread so far (slot = parity).
template <int POLY8, bool MASKED>
__device__ __forceinline__ void softmax_step_cs(uint32_t t_s, uint32_t t_o, uint64_t* s_full_i, uint64_t* p_full_i, uint32_t parity, bool first,
                                                int valid, float sc, float& m_run, float& l_half, int lane, int h, int quad, float* xch,
                                                int& xcnt) {
  constexpr int P8 = MASKED ? 0 : POLY8;
  mbar_wait(s_full_i, parity);
  tc_fence_after();
  const uint32_t ts_h = t_s + h * 64;
  const int row = quad * 32 + lane;
  uint32_t ra[32], rb[32];
  tmem_ld32(ts_h, ra);
  tmem_ld32(ts_h + 32, rb);
  tmem_ld_wait();
  if constexpr (MASKED) {  // keys >= valid were zero-filled by TMA: exclude them
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      if (h * 64 + k >= valid) ra[k] = 0xff800000u;
      if (h * 64 + 32 + k >= valid) rb[k] = 0xff800000u;
    }
  }
  float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int k = 0; k < 32; k += 2) {
    mx4[(k >> 1) & 3] = max3f(mx4[(k >> 1) & 3], __uint_as_float(ra[k]), __uint_as_float(ra[k + 1]));
    mx4[((k >> 1) + 2) & 3] = max3f(mx4[((k >> 1) + 2) & 3], __uint_as_float(rb[k]), __uint_as_float(rb[k + 1]));
  }
  float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
  // row max = max of the two halves; the barrier also means: both warpgroups hold their S columns in registers, P may overwrite them
  float* slot = xch + (xcnt & 1) * 256;
  slot[h * 128 + row] = mx;
  named_bar_sync(1 + quad, 64);
  mx = fmaxf(mx, slot[(1 - h) * 128 + row]);
  ++xcnt;
  const float m_cand = fmaxf(m_run, mx * sc);
  float alpha = 1.0f;
  bool rescale = false;
  if (first) {
    m_run = m_cand;
  } else if (m_cand - m_run > 8.0f) {  // lazy rescaling: tolerate a stale max up to 2^8 (both threads of a row decide alike)
    alpha = fast_exp2(m_run - m_cand);
    m_run = m_cand;
    rescale = true;
  }
  // O_i may be touched without waiting on a barrier: s_full(j) implies P.V_i(j-1) has retired (see softmax_step_core).  Each warpgroup
  // rescales its 64 columns; no P of this step may be handed over before BOTH have finished (a P.V product updates all 128 columns).
  if (!first && __any_sync(0xffffffffu, rescale)) {
#pragma unroll 1
    for (int c = 2 * h; c < 2 * h + 2; ++c) {
      uint32_t o[32];
      tmem_ld32(t_o + c * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
      tmem_st32(t_o + c * 32, o);
    }
    tmem_st_wait();
    tc_fence_before();
    named_bar_sync(1 + quad, 64);
    tc_fence_after();
  }
  const float m_eff = (MASKED && m_run == -INFINITY) ? 0.f : m_run;
  const uint64_t sc2 = pack_f32x2(sc, sc), nm2 = pack_f32x2(-m_eff, -m_eff);
  uint64_t sum2[4] = {0ull, 0ull, 0ull, 0ull};
  uint32_t pk[16];
  // quarter 2h (keys 64h .. 64h+31) -> P columns [0,16) of this warpgroup's own S columns
  exp_half<P8, 0, 0>(ra, sc2, nm2, sum2, pk);
  exp_half<P8, 0, 1>(ra, sc2, nm2, sum2, pk);
  tmem_st16(ts_h, pk);
  reg_fence32(rb);
  // quarter 2h+1; the hand-off of the first quarter is issued in the middle, when its store has landed
  exp_half<P8, 0, 0>(rb, sc2, nm2, sum2, pk);
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(&p_full_i[2 * h]);
  exp_half<P8, 0, 1>(rb, sc2, nm2, sum2, pk);
  tmem_st16(ts_h + 16, pk);
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(&p_full_i[2 * h + 1]);
  float s0, s1;
  unpack_f32x2(add_f32x2(add_f32x2(sum2[0], sum2[1]), add_f32x2(sum2[2], sum2[3])), s0, s1);
  l_half = l_half * alpha + (s0 + s1);
}

template <int POLY8>
__global__ void __launch_bounds__(ATT_THREADS, 1)
mmdit_attention_fwd_persistent_cs_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                                         const __grid_constant__ CUtensorMap tma_v, const AttnParams p, const int n_qblk, const int n_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sq = smem;                         // Q0 | Q1
  uint8_t* skv = smem + 2 * ATT_TILE_BYTES;   // ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(skv + ATT_KV_SLOTS * ATT_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* q_empty = bars + 1;
  uint64_t* kv_full = bars + 2;
  uint64_t* kv_empty = kv_full + ATT_KV_SLOTS;
  uint64_t* s_full = kv_empty + ATT_KV_SLOTS;    // 2
  uint64_t* p_full = s_full + 2;                 // 2 tiles x 4 quarters; quarters 2h, 2h+1 come from warpgroup h
  uint64_t* o_full = p_full + 8;                 // 2
  uint64_t* o_empty = o_full + 2;                // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);
  float* xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [2][2][128] max exchange
  float* lx = xch + 512;                                                            // [2 tiles][128] row-sum exchange

  const int warp = uniform_warp_id();
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < ATT_KV_SLOTS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      for (int c = 0; c < 4; ++c) mbar_init(&p_full[i * 4 + c], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 4);
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
  griddep_launch();  // PDL (common.cuh)
  griddep_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (as in the persistent kernel)
    if (lane == 0) {
      int seq = 0;
      int n = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++n) {
        const AttnItem t = attn_item<false>(p, w, n_qblk);
        mbar_wait(q_empty, (n & 1) ^ 1);
        mbar_expect_tx(q_full, 2 * ATT_TILE_BYTES);
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int g = 0; g < 2; ++g)
            tma_load_3d(sq + i * ATT_TILE_BYTES + g * 16384, &tma_q, q_full, g * 64, t.q0 + i * 128, t.bh);
        for (int s = 0; s < 2 * t.n_kv; ++s, ++seq) {
          const int slot = seq & (ATT_KV_SLOTS - 1);
          const uint32_t ph = (seq / ATT_KV_SLOTS) & 1;
          mbar_wait(&kv_empty[slot], ph ^ 1);
          mbar_expect_tx(&kv_full[slot], ATT_TILE_BYTES);
          const CUtensorMap* map = (s & 1) ? &tma_v : &tma_k;
          const int j = s >> 1;
          uint8_t* dst = skv + slot * ATT_TILE_BYTES;
          tma_load_3d(dst, map, &kv_full[slot], 0, j * 128, t.bh_kv);
          tma_load_3d(dst + 16384, map, &kv_full[slot], 64, j * 128, t.bh_kv);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer: the persistent kernel's sequence; only the A operand of
    // the P.V products moves (P quarter q sits at S column (q >> 1) * 64 + (q & 1) * 16)
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 128, 0, 1);
    const uint32_t q_base = smem_u32(sq);
    const uint32_t kv_base = smem_u32(skv);
    const uint64_t qdesc = make_smem_desc_sw128(q_base, 16, 1024);
    auto issue_s = [&](int i, int slot) {
      const uint64_t ad = qdesc + ((i * ATT_TILE_BYTES) >> 4);
      const uint64_t bd = make_smem_desc_sw128(kv_base + slot * ATT_TILE_BYTES, 16, 1024);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t off = ((kk >> 2) * 16384 + (kk & 3) * 32) >> 4;
        umma_ss_w(tmem_base + i * 128, ad + off, bd + off, idesc_s, kk != 0);
      }
    };
    auto issue_pv = [&](int i, int slot, bool acc, uint32_t ph) {
      const uint64_t bd = make_smem_desc_sw128(kv_base + slot * ATT_TILE_BYTES, 16384, 1024);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        mbar_wait(&p_full[i * 4 + c], ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 2 * c; kk < 2 * c + 2; ++kk)
          umma_ts_w(tmem_base + 256 + i * 128, tmem_base + i * 128 + (kk >> 2) * 64 + (kk & 3) * 8, bd + ((kk * 2048) >> 4), idesc_o,
                    (acc || kk != 0) ? 1u : 0u);
      }
    };
    auto kv_wait = [&](int s) { mbar_wait(&kv_full[s & (ATT_KV_SLOTS - 1)], (s / ATT_KV_SLOTS) & 1); };

    int seq = 0;
    int it = 0;
    int n = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++n) {
      const AttnItem t = attn_item<false>(p, w, n_qblk);
      mbar_wait(q_full, n & 1);
      kv_wait(seq);
      tc_fence_after();
      issue_s(0, seq & (ATT_KV_SLOTS - 1));
      umma_commit_w(&s_full[0]);
      issue_s(1, seq & (ATT_KV_SLOTS - 1));
      umma_commit_w(&s_full[1]);
      umma_commit_w(&kv_empty[seq & (ATT_KV_SLOTS - 1)]);
      if (t.n_kv == 1) umma_commit_w(q_empty);
      for (int j = 0; j < t.n_kv; ++j, ++it) {
        const int vseq = seq + 2 * j + 1, kseq = seq + 2 * j + 2;
        const int vslot = vseq & (ATT_KV_SLOTS - 1), kslot = kseq & (ATT_KV_SLOTS - 1);
        const bool more = (j + 1 < t.n_kv);
        kv_wait(vseq);
        if (j == 0) {
          mbar_wait(&o_empty[0], (n & 1) ^ 1);
          tc_fence_after();
        }
        issue_pv(0, vslot, j > 0, it & 1);
        if (!more) umma_commit_w(&o_full[0]);
        if (more) {
          kv_wait(kseq);
          tc_fence_after();
          issue_s(0, kslot);
          umma_commit_w(&s_full[0]);
        }
        if (j == 0) {
          mbar_wait(&o_empty[1], (n & 1) ^ 1);
          tc_fence_after();
        }
        issue_pv(1, vslot, j > 0, it & 1);
        if (!more) umma_commit_w(&o_full[1]);
        umma_commit_w(&kv_empty[vslot]);
        if (more) {
          issue_s(1, kslot);
          umma_commit_w(&s_full[1]);
          umma_commit_w(&kv_empty[kslot]);
          if (j + 2 == t.n_kv) umma_commit_w(q_empty);
        }
      }
      seq += 2 * t.n_kv;
    }
  } else {
    // ---------------------------------------------------------------- soft-max warpgroups: warpgroup h = key half h of BOTH query tiles
    const int h = (warp - 2) >> 2;
    const int quad = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const int row = quad * 32 + lane;
    const float sc = p.scale_log2;
    int it = 0;
    int n = 0;
    int xcnt = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++n) {
      const AttnItem t = attn_item<false>(p, w, n_qblk);
      float m_run[2] = {-INFINITY, -INFINITY};
      float l_half[2] = {0.f, 0.f};
      const bool ragged = (t.kv_valid & 127) != 0;
      const int n_full = ragged ? t.n_kv - 1 : t.n_kv;
      for (int jj = 0; jj < n_full; ++jj, ++it) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
          softmax_step_cs<POLY8, false>(tmem_base + i * 128 + lane_off, tmem_base + 256 + i * 128 + lane_off, &s_full[i], &p_full[i * 4], it & 1,
                                        jj == 0, 128, sc, m_run[i], l_half[i], lane, h, quad, xch, xcnt);
      }
      if (ragged) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
          softmax_step_cs<POLY8, true>(tmem_base + i * 128 + lane_off, tmem_base + 256 + i * 128 + lane_off, &s_full[i], &p_full[i * 4], it & 1,
                                       n_full == 0, t.kv_valid - n_full * 128, sc, m_run[i], l_half[i], lane, h, quad, xch, xcnt);
        ++it;
      }
      // warpgroup h finishes tile h: the other half's row sum comes through shared memory
      lx[(1 - h) * 128 + row] = l_half[1 - h];
      named_bar_sync(1 + quad, 64);
      const float l_run = l_half[h] + lx[h * 128 + row];
      const int pos = t.q0 + h * 128 + row;
      attn_epilogue(p, &o_full[h], tmem_base + 256 + h * 128 + lane_off, pos, t.b, t.h, t.bh, m_run[h], l_run, n & 1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[h]);
      named_bar_sync(1 + quad, 64);  // lx is rewritten at the end of the next item: both partners have read it
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace x2i
