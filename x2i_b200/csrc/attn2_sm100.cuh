// CTA-pair (cta_group::2) form of the fused MMDiT joint attention forward (attn_sm100.cuh): two CTAs on the SMs of one TPC
// share every K / V tile.
//
// Why: per 128-key step the single-CTA kernel moves 256 KB through one SM's shared memory (128 KB of Q.K^T operand reads, 64 KB
// of V operand reads, 64 KB of TMA fills) in the 2048 cycles the tensor pipe needs -- 125 of the 128 B/clk the port has -- and
// every CTA pulls the whole K and V of its (batch, head) through L2 (432 CTAs x 2.36 MB = 1 GB per launch).  In the pair form a
// CTA stages only ITS HALF of each K tile (64 keys) and of each V tile (64 head-dim columns); tcgen05.mma.cta_group::2 (M = 256:
// 128 query rows in each CTA, N = 128) reads both halves.  Per SM: 160 KB per step (64 + 32 + 32 + 32) and half the L2 -> SM bytes,
// and the freed shared memory makes the K/V ring 8 half-tiles deep (4 key steps of prefetch instead of 2).
//
// One pair per (batch, head, 512 query rows); CTA rank r owns query rows [512 p + 256 r, +256) as two 128-row tiles, exactly the
// work of one single-CTA block, so the soft-max warpgroups, the TMEM plan (S0 | S1 | O0 | O1, P aliasing S) and the epilogue are
// those of attn_sm100.cuh.  What changes is the plumbing:
//   kv_full[s], q_full      (leader CTA)  armed by the leader's producer with the bytes of BOTH CTAs; both CTAs' TMA loads signal it
//   kv_empty[s], s_full[i], o_full[i]     tcgen05.commit.cta_group::2 ... multicast: arrives in both CTAs
//   p_full[i][quarter]      (leader CTA)  8 arrivals: 4 soft-max warps x 2 CTAs (the peer arrives remotely through mapa)
// Only the leader's warp 1 issues MMAs; the peer's warp 1 idles.
#pragma once
#include "attn_sm100.cuh"
#include "gemm2_sm100.cuh"

namespace x2i {

constexpr int ATT2_SLOTS = 8;                    // ring of half tiles (16 KB): K half = 64 keys x 128 d, V half = 128 keys x 64 d
constexpr int ATT2_HALF_BYTES = ATT_TILE_BYTES / 2;
constexpr int ATT2_SMEM_BYTES = 2 * ATT_TILE_BYTES + ATT2_SLOTS * ATT2_HALF_BYTES + 1024 + 512;

__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* leader_bar_local, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(leader_bar_local) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_ts_pair_w(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// tma_k: box [64 d, 64 keys, 1]; tma_q / tma_v: box [64 d, 128 rows, 1] (capi.cu).
template <int POLY8>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ATT_THREADS, 1)
mmdit_attention_fwd2_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                            const __grid_constant__ CUtensorMap tma_v, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sq = smem;                         // Q0 | Q1 (this CTA's 2 x 128 query rows)
  uint8_t* skv = smem + 2 * ATT_TILE_BYTES;   // ring of half tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(skv + ATT2_SLOTS * ATT2_HALF_BYTES);
  uint64_t* q_full = bars;                       // 1   (leader's copy is used)
  uint64_t* kv_full = bars + 1;                  // 8   (leader's copy is used)
  uint64_t* kv_empty = kv_full + ATT2_SLOTS;     // 8   (each CTA)
  uint64_t* s_full = kv_empty + ATT2_SLOTS;      // 2   (each CTA)
  uint64_t* p_full = s_full + 2;                 // 8   (leader's copy is used)
  uint64_t* o_full = p_full + 8;                 // 2   (each CTA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = uniform_warp_id();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int q0 = blockIdx.x * 256;  // blockIdx.x = 2 * pair + rank
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int bh = b * p.H + h;
  const int kv_valid = p.kv_len != nullptr ? min(max(p.kv_len[b], 1), p.Lkv) : p.Lkv;
  const int n_kv = (kv_valid + 127) / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < ATT2_SLOTS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      for (int c = 0; c < 4; ++c) mbar_init(&p_full[i * 4 + c], 8);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of BOTH CTAs initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (one per CTA: own Q, own halves of K / V)
    if (lane == 0) {
      if (rank == 0) mbar_expect_tx(q_full, 4 * ATT_TILE_BYTES);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int g = 0; g < 2; ++g)
          tma_load_3d_pair(sq + i * ATT_TILE_BYTES + g * 16384, &tma_q, q_full, g * 64, q0 + i * 128, bh);
      for (int seq = 0; seq < 2 * n_kv; ++seq) {
        const int slot = seq & (ATT2_SLOTS - 1);
        const uint32_t ph = (seq / ATT2_SLOTS) & 1;
        mbar_wait(&kv_empty[slot], ph ^ 1);
        if (rank == 0) mbar_expect_tx(&kv_full[slot], 2 * ATT2_HALF_BYTES);
        const int j = seq >> 1;
        uint8_t* dst = skv + slot * ATT2_HALF_BYTES;
        if ((seq & 1) == 0) {  // K half: keys [128 j + 64 rank, +64), all 128 d as two K-major boxes of 64 rows
          tma_load_3d_pair(dst, &tma_k, &kv_full[slot], 0, j * 128 + static_cast<int>(rank) * 64, bh);
          tma_load_3d_pair(dst + 8192, &tma_k, &kv_full[slot], 64, j * 128 + static_cast<int>(rank) * 64, bh);
        } else {               // V half: all 128 keys, d columns [64 rank, +64) as one MN-major box
          tma_load_3d_pair(dst, &tma_v, &kv_full[slot], static_cast<int>(rank) * 64, j * 128, bh);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (leader CTA only; whole warp, elected lane issues)
    if (rank == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(256, 128, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(256, 128, 0, 1);
      const uint32_t q_base = smem_u32(sq);
      const uint32_t kv_base = smem_u32(skv);
      const uint64_t qdesc = make_smem_desc_sw128(q_base, 16, 1024);
      auto issue_s = [&](int i, int slot) {  // S_i = Q_i K^T over both CTAs' query rows
        const uint64_t ad = qdesc + ((i * ATT_TILE_BYTES) >> 4);
        const uint64_t bd = make_smem_desc_sw128(kv_base + slot * ATT2_HALF_BYTES, 16, 1024);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off_a = ((kk >> 2) * 16384 + (kk & 3) * 32) >> 4;
          const uint32_t off_b = ((kk >> 2) * 8192 + (kk & 3) * 32) >> 4;
          umma_ss_pair_w(tmem_base + i * 128, ad + off_a, bd + off_b, idesc_s, kk != 0);
        }
      };
      auto issue_pv = [&](int i, int slot, bool acc, uint32_t ph) {  // O_i += P_i V, quarter by quarter as P lands in BOTH CTAs
        const uint64_t bd = make_smem_desc_sw128(kv_base + slot * ATT2_HALF_BYTES, 16384, 1024);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          mbar_wait(&p_full[i * 4 + c], ph);
          tc_fence_after();
#pragma unroll
          for (int kk = 2 * c; kk < 2 * c + 2; ++kk)
            umma_ts_pair_w(tmem_base + 256 + i * 128, tmem_base + i * 128 + kk * 8, bd + ((kk * 2048) >> 4), idesc_o, (acc || kk != 0) ? 1u : 0u);
        }
      };
      auto kv_wait = [&](int seq) { mbar_wait(&kv_full[seq & (ATT2_SLOTS - 1)], (seq / ATT2_SLOTS) & 1); };

      mbar_wait(q_full, 0);
      kv_wait(0);
      tc_fence_after();
      issue_s(0, 0);
      umma_commit_pair_w(&s_full[0]);
      issue_s(1, 0);
      umma_commit_pair_w(&s_full[1]);
      umma_commit_pair_w(&kv_empty[0]);
      for (int j = 0; j < n_kv; ++j) {
        const int vseq = 2 * j + 1, kseq = 2 * j + 2;
        const int vslot = vseq & (ATT2_SLOTS - 1), kslot = kseq & (ATT2_SLOTS - 1);
        const bool more = (j + 1 < n_kv);
        kv_wait(vseq);
        issue_pv(0, vslot, j > 0, j & 1);
        if (!more) umma_commit_pair_w(&o_full[0]);
        if (more) {
          kv_wait(kseq);
          tc_fence_after();
          issue_s(0, kslot);
          umma_commit_pair_w(&s_full[0]);
        }
        issue_pv(1, vslot, j > 0, j & 1);
        if (!more) umma_commit_pair_w(&o_full[1]);
        umma_commit_pair_w(&kv_empty[vslot]);
        if (more) {
          issue_s(1, kslot);
          umma_commit_pair_w(&s_full[1]);
          umma_commit_pair_w(&kv_empty[kslot]);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax warpgroups (both CTAs, own 2 x 128 query rows)
    const int i = (warp - 2) >> 2;
    const int quad = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_s = tmem_base + i * 128 + lane_off;
    const uint32_t t_o = tmem_base + 256 + i * 128 + lane_off;
    const int pos = q0 + i * 128 + quad * 32 + lane;
    float m_run = -INFINITY;
    float l_run = 0.f;
    const float sc = p.scale_log2;

    for (int j = 0; j < n_kv; ++j) {
      const int hi = min(128, kv_valid - j * 128);
      if (hi < 128)
        softmax_step<POLY8, true, 0, true>(t_s, t_o, &s_full[i], &p_full[i * 4], j, hi, sc, m_run, l_run, lane, nullptr);
      else
        softmax_step<POLY8, false, 0, true>(t_s, t_o, &s_full[i], &p_full[i * 4], j, 128, sc, m_run, l_run, lane, nullptr);
    }
    attn_epilogue(p, &o_full[i], t_o, pos, b, h, bh, m_run, l_run);
  }

  tc_fence_before();
  cluster_sync_all();  // both CTAs done with TMEM; no remote arrival or multicast commit still in flight
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace x2i
