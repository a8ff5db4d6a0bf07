// CTA-pair (cta_group::2) variant of the tcgen05 GEMM: two CTAs on the SMs of one TPC cooperate on a 256 x 256 tile.
//
// Each CTA stages ITS 128 rows of A and ITS 128-row half of the W tile (the pair's tensor cores read both halves), so per
// SM the TMA write traffic into shared memory and the UMMA operand reads are both 2/3 of the single-CTA 128x256 kernel
// -- the single-CTA kernel asks ~192 B/clk of a 128 B/clk shared-memory port.  The leader CTA (cluster rank 0) issues
// tcgen05.mma.cta_group::2 (M=256, N=256, K=16); accumulators live in both CTAs' TMEM (128 lanes x 256 columns each,
// double-buffered); each CTA's epilogue warps drain their own 128 rows with the same fused epilogues as gemm_sm100.cuh.
//
// Synchronisation (per CTA pair):
//   full_bar[s]   (leader CTA): armed by the leader's producer with the bytes of BOTH CTAs; both CTAs' TMA loads signal it
//   empty_bar[s]  (each CTA)  : tcgen05.commit.cta_group::2 ... multicast -> both producers may refill stage s
//   tfull_bar[a]  (each CTA)  : multicast commit after the last k-block -> both epilogues start
//   tempty_bar[a] (leader CTA): 8 arrivals (4 epilogue warps x 2 CTAs; the peer arrives remotely via mapa)
#pragma once
#include "gemm_sm100.cuh"

namespace x2i {

constexpr int GEMM2_STAGES = 6;
constexpr int GEMM2_STAGE_BYTES = 2 * 128 * GEMM_BK * 2;  // A half (16 KB) + W half (16 KB) per CTA
#ifdef X2I_EPI_STAGE  // experiment build: + 4 KB per epilogue warp for the store transpose (gemm_epilogue_tile)
constexpr int GEMM2_SMEM_BYTES = GEMM2_STAGES * GEMM2_STAGE_BYTES + 1024 + 256 + 4 * 4096;
#else
constexpr int GEMM2_SMEM_BYTES = GEMM2_STAGES * GEMM2_STAGE_BYTES + 1024 + 256;
#endif
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared-window address -> the pair's even CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* leader_bar_local, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(leader_bar_local) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_ss_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {  // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
// whole-warp forms (see common.cuh): all 32 lanes call, one elected lane issues
__device__ __forceinline__ void umma_ss_pair_w(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair_w(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}\n" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar_local) {  // arrive on the leader CTA's copy of `bar`
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar_local))
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// Up to two problems (same epilogue kind; e.g. the image and text streams of a double block, which have different
// weights and M) share one persistent launch: tiles [0, tiles0) belong to problem 0, the rest to problem 1.  This fills
// the partial waves a 512-row text GEMM (24-96 pair tiles on 74 SM pairs) would otherwise leave idle.
struct GemmGroup {
  GemmParams p[2];
  int tiles0;     // pair tiles of problem 0
  int num_tiles;  // pair tiles of both problems
};

struct TileRef {
  int prob, m2, n_blk;
};
__device__ __forceinline__ TileRef locate_tile(const GemmGroup& g, int tile) {
  TileRef t;
  t.prob = tile >= g.tiles0 ? 1 : 0;
  const int local = tile - (t.prob ? g.tiles0 : 0);
  const int num_m2 = (g.p[t.prob].M + 255) / 256;
  t.m2 = local % num_m2;
  t.n_blk = local / num_m2;
  return t;
}

// B_MN: W given as [K, N] (N contiguous) -- each CTA stages its 128-column half as two 64 x 64 boxes (dgrad GEMMs).
template <int EPI, bool B_MN = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a0, const __grid_constant__ CUtensorMap tma_b0,
                     const __grid_constant__ CUtensorMap tma_a1, const __grid_constant__ CUtensorMap tma_b1,
                     const __grid_constant__ GemmGroup g) {
  constexpr int NS = GEMM2_STAGES;
  constexpr int BN = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NS * GEMM2_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + NS;
  uint64_t* tfull_bar = empty_bar + NS;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = uniform_warp_id();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_tiles = g.num_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a0);
    tma_prefetch_desc(&tma_b0);
    if (g.num_tiles > g.tiles0) {
      tma_prefetch_desc(&tma_a1);
      tma_prefetch_desc(&tma_b1);
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);
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
  griddep_launch();  // PDL: the next kernel's prologue may overlap this kernel's tail ...
  griddep_wait();    // ... and this kernel touches global memory only after its predecessors are complete

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (one per CTA)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const TileRef t = locate_tile(g, tile);
        const CUtensorMap* ma = t.prob ? &tma_a1 : &tma_a0;
        const CUtensorMap* mb = t.prob ? &tma_b1 : &tma_b0;
        const int num_kb = (g.p[t.prob].K + GEMM_BK - 1) / GEMM_BK;
        const int row_a = t.m2 * 256 + rank * 128, row_b = t.n_blk * BN + rank * 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * GEMM2_STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * GEMM2_STAGE_BYTES);
          tma_load_2d_pair(sa, ma, &full_bar[stage], kb * GEMM_BK, row_a);
          if constexpr (!B_MN) {
            tma_load_2d_pair(sa + 16384, mb, &full_bar[stage], kb * GEMM_BK, row_b);
          } else {
            tma_load_2d_pair(sa + 16384, mb, &full_bar[stage], row_b, kb * GEMM_BK);
            tma_load_2d_pair(sa + 16384 + 8192, mb, &full_bar[stage], row_b + 64, kb * GEMM_BK);
          }
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0) {  // whole warp runs the loop, the elected lane issues (uniform-register descriptors)
      constexpr uint32_t idesc = make_idesc_bf16(256, BN, 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        const int num_kb = (g.p[tile >= g.tiles0 ? 1 : 0].K + GEMM_BK - 1) / GEMM_BK;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * GEMM2_STAGE_BYTES);
          const uint64_t adesc = make_smem_desc_sw128(a_base, 16, 1024);
          const uint64_t bdesc = B_MN ? make_smem_desc_sw128(a_base + 16384, 8192, 1024) : make_smem_desc_sw128(a_base + 16384, 16, 1024);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            umma_ss_pair_w(d_tmem, adesc + ((k * 32) >> 4), bdesc + ((B_MN ? k * 2048 : k * 32) >> 4), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_pair_w(&empty_bar[stage]);
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair_w(&tfull_bar[as]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (2..5) of each CTA: own 128 rows
    const int quad = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      const TileRef t = locate_tile(g, tile);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + as * BN + lane_off;
      const int m = t.m2 * 256 + rank * 128 + quad * 32 + lane;
#if defined(X2I_GEMM_EPI_TMEM_ONLY)  // timing experiment (tools/jobs/gpu_job_r04m.sh): the epilogue only READS the accumulator tile
      {
        uint32_t acc = 0;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(t_acc + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc ^= r[j];
        }
        if (acc == 0x12345678u && m < g.p[t.prob].M) g.p[t.prob].C[m] = __float2bfloat16(1.0f);
      }
#elif defined(X2I_GEMM_EPI_STORE_ONLY)  // timing experiment: the epilogue only WRITES (zeros), no TMEM read
      {
        const GemmParams& gp = g.p[t.prob];
        if (m < gp.M && gp.C != nullptr) {
#pragma unroll 1
          for (int c = 0; c < BN / 32; ++c) {
            const int n0 = t.n_blk * BN + c * 32;
            if (n0 >= gp.N) break;
            uint4* d4 = reinterpret_cast<uint4*>(gp.C + static_cast<long long>(m) * gp.ldc + n0);
#pragma unroll
            for (int j = 0; j < 4; ++j) d4[j] = make_uint4(0, 0, 0, 0);
          }
        }
      }
#elif !defined(X2I_GEMM_SKIP_EPI)  // X2I_GEMM_SKIP_EPI: timing experiment (tools/jobs/gpu_job_r04l.sh): mainloop only, no results
#ifdef X2I_EPI_STAGE
      {
        const GemmParams& gp = g.p[t.prob];  // staged stores need 32-byte aligned rows and whole 64-column pairs
        const bool ok = (reinterpret_cast<uintptr_t>(gp.C) & 31) == 0 && (gp.ldc & 15) == 0 && gp.C != nullptr;
        gemm_epilogue_tile<BN, EPI>(gp, t_acc, m, t.n_blk * BN, 0, 0, ok ? smem + NS * GEMM2_STAGE_BYTES + 256 + (warp - 2) * 4096 : nullptr);
      }
#else
      gemm_epilogue_tile<BN, EPI>(g.p[t.prob], t_acc, m, t.n_blk * BN);
#endif
#endif
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  cluster_sync_all();  // both CTAs done with TMEM / no more remote arrivals in flight
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace x2i
