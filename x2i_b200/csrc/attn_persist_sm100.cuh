// Persistent form of the fused attention forward (attn_sm100.cuh): one CTA per SM loops over (batch, head, 256-query block) work
// items, so the per-item prologue (TMEM allocation, barrier init, Q / first K load latency, the first two S products: ~5300 cycles
// of a ~100 k cycle item in the clock64 trace, profiles/r02_attn_clock64_trace.txt) and the drain of the epilogue are overlapped with
// neighbouring items instead of being paid 432 times:
//   * the TMA producer runs ahead: as soon as the MMA warp has issued the LAST S product of item n (q_empty) it loads Q of item n+1,
//     and the K/V ring simply continues into the next item's tiles;
//   * the MMA warp issues S0 / S1 of item n+1 right behind the last P.V of item n, i.e. while the soft-max warpgroups are still
//     normalising and storing O of item n; the first P.V of item n+1 waits for that drain (o_empty);
//   * TMEM, barriers and tensor-map prefetches are set up once per CTA.
// Work items are ordered query-block fastest, so the CTAs in flight at any time share a handful of (batch, head) K/V tensors in L2.
// Soft-max code, TMEM plan and epilogue are exactly those of the one-item kernel (bit-identical results).
#pragma once
#include "attn_sm100.cuh"

namespace x2i {

constexpr int ATTP_BARRIERS = 1 + 1 + 2 * ATT_KV_SLOTS + 2 + 8 + 2 + 2;  // q_full, q_empty, kv_full/empty, s_full, p_full, o_full, o_empty

struct AttnItem {
  int q0, h, b, bh, bh_kv, kv_valid, kv_lo, j0, n_kv;
};
template <bool LM>
__device__ __forceinline__ AttnItem attn_item(const AttnParams& p, int w, int n_qblk) {
  AttnItem t;
  const int qb = w % n_qblk;
  t.bh = w / n_qblk;
  t.b = t.bh / p.H;
  t.h = t.bh - t.b * p.H;
  t.q0 = qb * 256;
  t.kv_valid = p.kv_len != nullptr ? min(max(p.kv_len[t.b], 1), p.Lkv) : p.Lkv;
  t.kv_lo = (LM && p.kv_start != nullptr) ? min(max(p.kv_start[t.b], 0), t.kv_valid - 1) : 0;
  t.j0 = LM ? (t.kv_lo >> 7) : 0;
  int j_end = (t.kv_valid + 127) / 128;
  if (LM && p.causal) j_end = min(j_end, (min(t.q0 + 255, p.L - 1) >> 7) + 1);
  t.n_kv = LM ? max(j_end - t.j0, 1) : j_end;
  t.bh_kv = (LM && p.Hkv > 0 && p.Hkv != p.H) ? t.b * p.Hkv + t.h / (p.H / p.Hkv) : t.bh;
  return t;
}

// The three warp roles over this CTA's work items.  LAGGED: lagged soft-max steps (softmax_step_lagged, attn_sm100.cuh) for every unmasked
// key step but the first of a row.
template <int POLY8, bool LM, bool LAGGED>
__device__ __forceinline__ void attn_persist_roles(const CUtensorMap& tma_q, const CUtensorMap& tma_k, const CUtensorMap& tma_v, const AttnParams& p,
                                                   const int n_qblk, const int n_items, uint8_t* sq, uint8_t* skv, uint64_t* bars,
                                                   const uint32_t tmem_base, int* redo_flag, const int warp, const int lane) {
  uint64_t* q_full = bars;
  uint64_t* q_empty = bars + 1;
  uint64_t* kv_full = bars + 2;
  uint64_t* kv_empty = kv_full + ATT_KV_SLOTS;
  uint64_t* s_full = kv_empty + ATT_KV_SLOTS;
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = p_full + 8;
  uint64_t* o_empty = o_full + 2;
  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int seq = 0;  // K / V tiles loaded so far (ring position), across items
      int n = 0;    // items processed by this CTA
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++n) {
        const AttnItem t = attn_item<LM>(p, w, n_qblk);
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
          const int j = t.j0 + (s >> 1);
          uint8_t* dst = skv + slot * ATT_TILE_BYTES;
          tma_load_3d(dst, map, &kv_full[slot], 0, j * 128, t.bh_kv);
          tma_load_3d(dst + 16384, map, &kv_full[slot], 64, j * 128, t.bh_kv);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (whole warp, elected lane issues)
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 128, 0, 1);
    const uint32_t q_base = smem_u32(sq);
    const uint32_t kv_base = smem_u32(skv);
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
        mbar_wait(&p_full[i * 4 + c], ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 2 * c; kk < 2 * c + 2; ++kk)
          umma_ts_w(tmem_base + 256 + i * 128, tmem_base + i * 128 + kk * 8, bd + ((kk * 2048) >> 4), idesc_o, (acc || kk != 0) ? 1u : 0u);
      }
    };
    auto kv_wait = [&](int s) { mbar_wait(&kv_full[s & (ATT_KV_SLOTS - 1)], (s / ATT_KV_SLOTS) & 1); };

    int seq = 0;   // ring position of the K tile of the current key step
    int it = 0;    // key steps processed so far, across items (parity of s_full / p_full)
    int n = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++n) {
      const AttnItem t = attn_item<LM>(p, w, n_qblk);
      // S0, S1 of the item's first key step: behind the last P.V of the previous item, overlapping its epilogue
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
        if (j == 0) {  // the first P.V overwrites O_0: the previous item's epilogue must have read it out
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
          if (j + 2 == t.n_kv) umma_commit_w(q_empty);  // that was the item's last S product: Q may be replaced
        }
      }
      seq += 2 * t.n_kv;
    }
  } else {
    // ---------------------------------------------------------------- softmax warpgroups
    const int i = (warp - 2) >> 2;  // query tile 0 / 1
    const int quad = warp & 3;      // TMEM lane quadrant accessible to this warp
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_s = tmem_base + i * 128 + lane_off;
    const uint32_t t_o = tmem_base + 256 + i * 128 + lane_off;
    const float sc = p.scale_log2;
    int it = 0;
    int n = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++n) {
      const AttnItem t = attn_item<LM>(p, w, n_qblk);
      const int pos = t.q0 + i * 128 + quad * 32 + lane;
      float m_run = -INFINITY;
      float l_run = 0.f;
      if constexpr (!LM) {  // hot path: mask-free tight loop, a ragged last tile apart (see attn_sm100.cuh)
        const bool ragged = (t.kv_valid & 127) != 0;
        const int n_full = ragged ? t.n_kv - 1 : t.n_kv;
        if constexpr (LAGGED) {
          float m_next = -INFINITY;
          if (n_full > 0) {
            softmax_step_p<POLY8, false>(t_s, t_o, &s_full[i], &p_full[i * 4], it & 1, true, 128, sc, m_run, l_run, lane, 0);
            ++it;
          }
          for (int jj = 1; jj < n_full; ++jj, ++it)
            softmax_step_lagged<POLY8>(t_s, t_o, &s_full[i], &p_full[i * 4], it & 1, sc, m_run, l_run, m_next, lane, redo_flag);
        } else {
          for (int jj = 0; jj < n_full; ++jj, ++it)
            softmax_step_p<POLY8, false>(t_s, t_o, &s_full[i], &p_full[i * 4], it & 1, jj == 0, 128, sc, m_run, l_run, lane, 0);
        }
        if (ragged) {
          softmax_step_p<POLY8, true>(t_s, t_o, &s_full[i], &p_full[i * 4], it & 1, n_full == 0, t.kv_valid - n_full * 128, sc, m_run, l_run,
                                      lane, 0);
          ++it;
        }
      } else {
        for (int jj = 0; jj < t.n_kv; ++jj, ++it) {
          const int j = t.j0 + jj;
          int hi = min(128, t.kv_valid - j * 128);
          const int lo = max(0, t.kv_lo - j * 128);
          if (p.causal) hi = min(hi, pos - j * 128 + 1);
          if (__any_sync(0xffffffffu, hi < 128 || lo > 0))
            softmax_step_p<POLY8, true>(t_s, t_o, &s_full[i], &p_full[i * 4], it & 1, jj == 0, hi, sc, m_run, l_run, lane, lo);
          else
            softmax_step_p<POLY8, false>(t_s, t_o, &s_full[i], &p_full[i * 4], it & 1, jj == 0, 128, sc, m_run, l_run, lane, 0);
        }
      }
      attn_epilogue(p, &o_full[i], t_o, pos, t.b, t.h, t.bh, m_run, l_run, n & 1);
      tc_fence_before();  // the tcgen05.ld of O are complete (wait::ld inside the epilogue): hand O_i back to the MMA warp
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[i]);
    }
  }

}

// LAG: lagged soft-max steps in the main pass.  A step whose exponentials may have overflowed sets the CTA's redo flag; the CTA then drains,
// re-initialises its barriers and runs its items once more with the classic step (a second, straight-line copy of the role code: a loop
// around the roles cost 13 % -- profiles/r02_attn_probe_lagged.md), overwriting the outputs of the first pass.
template <int POLY8, bool LM = false, bool LAG = false>
__global__ void __launch_bounds__(ATT_THREADS, 1)
mmdit_attention_fwd_persistent_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                                      const __grid_constant__ CUtensorMap tma_v, const AttnParams p, const int n_qblk, const int n_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sq = smem;                         // Q0 | Q1
  uint8_t* skv = smem + 2 * ATT_TILE_BYTES;   // ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(skv + ATT_KV_SLOTS * ATT_TILE_BYTES);
  uint64_t* q_full = bars;                       // 1
  uint64_t* q_empty = bars + 1;                  // 1: all S products of the item have been issued and retired -> Q may be overwritten
  uint64_t* kv_full = bars + 2;                  // ATT_KV_SLOTS
  uint64_t* kv_empty = kv_full + ATT_KV_SLOTS;   // ATT_KV_SLOTS
  uint64_t* s_full = kv_empty + ATT_KV_SLOTS;    // 2
  uint64_t* p_full = s_full + 2;                 // 2 x 4
  uint64_t* o_full = p_full + 8;                 // 2
  uint64_t* o_empty = o_full + 2;                // 2: the soft-max warpgroup has read O_i of the previous item out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);
  int* redo_flag = reinterpret_cast<int*>(tmem_slot + 1);       // LAG: some step of this CTA may have overflowed
  uint64_t* drain_bar = reinterpret_cast<uint64_t*>(tmem_slot + 2);  // LAG: all tcgen05 work of the first pass has retired
  static_assert(ATTP_BARRIERS * 8 + 16 <= 256, "barrier block");

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
    mbar_init(drain_bar, 1);
    *redo_flag = 0;
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_launch();  // PDL (common.cuh): prologue overlaps the predecessor's tail; global memory only after griddep_wait()
  griddep_wait();
  const uint32_t tmem_base = *tmem_slot;

  attn_persist_roles<POLY8, LM, LAG>(tma_q, tma_k, tma_v, p, n_qblk, n_items, sq, skv, bars, tmem_base, redo_flag, warp, lane);
  if constexpr (LAG) {
    tc_fence_before();
    __syncthreads();  // every role is through its items: the flag is final
    tc_fence_after();
    if (*redo_flag != 0) {  // rare; uniform over the CTA
      if (warp == 1) {      // no commit of the first pass may still be in flight when the barriers are re-initialised
        umma_commit_w(drain_bar);
        mbar_wait(drain_bar, 0);
      }
      __syncthreads();
      if (warp == 0 && lane == 0) {
        for (int i = 0; i < ATTP_BARRIERS; ++i) mbar_inval(&bars[i]);
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
      __syncthreads();
      // second, straight-line copy of the role code; its inputs are re-derived here so that nothing of it extends a live range of the main pass
      uint8_t* smem2 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
      uint64_t* bars2 = reinterpret_cast<uint64_t*>(smem2 + (2 + ATT_KV_SLOTS) * ATT_TILE_BYTES);
      const uint32_t tmem_base2 = *reinterpret_cast<volatile uint32_t*>(bars2 + ATTP_BARRIERS);
      attn_persist_roles<POLY8, LM, false>(tma_q, tma_k, tma_v, p, n_qblk, n_items, smem2, smem2 + 2 * ATT_TILE_BYTES, bars2, tmem_base2, redo_flag,
                                       uniform_warp_id(), static_cast<int>(threadIdx.x & 31));
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
