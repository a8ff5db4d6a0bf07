// Backward of the fused MMDiT joint attention for sm_100a (head_dim 128, non-causal, no mask) -- the student pass of
// the attention-distillation step back-propagates through 57 of these (train/train_qwenvl.py:625).
//
//   P = softmax(Q K^T / sqrt(128)),  dV = P^T dO,  dP = dO V^T,  dS = P o (dP - delta),  delta_i = sum_d dO_id O_id
//   dQ = dS K / sqrt(128),  dK = dS^T Q / sqrt(128)
//
// Two launches of ONE kernel template, no atomics, bit-reproducible:
//   KV = true : a CTA owns 128 keys   (X0 = K_j, X1 = V_j resident), streams 64-query tiles (Y0 = Q, Y1 = dO):
//               T1 = S^T = X0 Y0^T, T2 = dP^T = X1 Y1^T; dK_j += dS~ Y0, dV_j += P~ Y1
//   KV = false: a CTA owns 128 queries (X0 = Q_i, X1 = dO_i resident), streams 64-key tiles  (Y0 = K, Y1 = V):
//               T1 = S = X0 Y0^T,   T2 = dP = X1 Y1^T;   dQ_i += dS~ Y0
// so both modes issue the same MMAs: two SS products into a double-buffered TMEM region (2 x (64 + 64) columns), the
// soft-max warpgroup turns them into bf16 P~ / dS~ IN TMEM (aliasing T1 / T2), and the accumulating products read them
// as the A operand from TMEM (TS form) with the streamed tile as an MN-major B operand.  The MMAs of stream tile y+2
// and the accumulation of tile y overlap the soft-max math of tile y+1.
//
// Measured and rejected (tools/attn_bwd_sweep.sh, X2I_ATTN_EXPERIMENTS build): removing ALL soft-max math changes the run time by
// < 13 % (dK/dV launch) / 0 % (dQ launch), so the kernel is bound by its MMA stream, not by MUFU or the hand-off; moving the owner
// tiles into TMEM (TS-form T products, three single-accumulator launches) made it SLOWER (0.69 -> 0.86 ms): the 128 x 64 x 16 T
// products take as long as 128 x 128 x 16 ones whether A comes from shared memory or TMEM, i.e. N = 64 runs the tensor pipe at
// half rate.  The next step is a 128-wide streamed tile (needs a TMEM plan with single-buffered T2), not more soft-max tuning.
//
// Round 2 (tools/ubench/mma_rate.cu, profiles/r02_mma_rate.txt): a 128 x 64 x 16 product costs 32 cycles with A in TMEM but 48 with A in
// shared memory -- the SS form re-reads the 4 KB A slice per instruction and the 128 B/clk shared-memory port, not the tensor pipe, paces it
// (SS N = 128: 64 cycles = 8 KB at exactly 128 B/clk).  So in the dQ launch (one accumulator, TMEM has room) the owner tiles Q_i / dO_i are
// written ONCE into TMEM as bf16 A operands by the soft-max warps and both T products run in the TS form; the dK/dV launch needs
// 2 x 128 accumulator + 2 x 128 T columns and keeps its owners in shared memory.
//
// 320 threads: warp 0 TMA producer, warp 1 MMA issuer (one lane), warps 2..5 / 6..9 two soft-max + epilogue
// warpgroups (one owner row per thread; warpgroup g owns T buffer g).  lse / delta are [B*H, Lpad] fp32 (log2 domain, +inf / 0 in the padding) written by the forward kernel and
// by attn_bwd_prep_kernel.  All tensors are head-major [B*H, L, 128] bf16.
#pragma once
#include "common.cuh"

namespace x2i {

struct AttnBwdParams {
  int B, H, L, Lpad;
  float scale_log2;  // log2(e) / sqrt(128)
  float scale;       // 1 / sqrt(128)
  const float* lse;
  const float* delta;
  __nv_bfloat16* out0;  // KV: dK, else dQ
  __nv_bfloat16* out1;  // KV: dV
  const __nv_bfloat16* x0g;  // dQ launch: the owner tensors (Q, dO) as plain pointers -- their rows go to TMEM through registers
  const __nv_bfloat16* x1g;
  int dbg;  // X2I_ATTN_EXPERIMENTS builds only: 1 = soft-max warps skip all math (timing floor of the MMA / TMA pipeline, wrong results)
};

constexpr int ABW_THREADS = 320;
constexpr int ABW_STAGES_KV = 4;  // dK/dV launch: the owner tiles take 64 KB of shared memory
constexpr int ABW_STAGES_Q = 6;   // dQ launch: owners live in TMEM, the whole shared memory is the stream ring (more bytes in flight from L2)
constexpr int ABW_OWNER_BYTES = 2 * 32768;  // X0 | X1, 128 x 128 bf16 each (two 128B-swizzled column halves of 16 KB)
constexpr int ABW_STAGE_BYTES = 2 * 16384;  // Y0 | Y1, 64 x 128 bf16 each (two column halves of 8 KB)
constexpr int ABW_STATS_BYTES = 512;        // per stage: lse[64] | delta[64] of the streamed queries (KV mode)
constexpr int ABW_SMEM_KV = ABW_OWNER_BYTES + ABW_STAGES_KV * (ABW_STAGE_BYTES + ABW_STATS_BYTES) + 256 + 1024;
constexpr int ABW_SMEM_Q = ABW_STAGES_Q * (ABW_STAGE_BYTES + ABW_STATS_BYTES) + 256 + 1024;
constexpr int ABW_SMEM_BYTES = ABW_SMEM_KV > ABW_SMEM_Q ? ABW_SMEM_KV : ABW_SMEM_Q;

// multicast forms (cluster of 2 CTAs with adjacent owner tiles of one head): a box lands at the same offset in both CTAs and signals the
// barrier at the same offset in both; the commit releases a ring stage in both CTAs
__device__ __forceinline__ void tma_load_3d_mc2(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc2_w(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}\n" ::"r"(smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ uint32_t abw_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void abw_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// MC: the two CTAs of a cluster own adjacent 128-row tiles of the same head and share every streamed tile -- each issues the TMA boxes
// of ITS 64-column half with .multicast::cluster, so the L2 -> SM traffic of a launch halves (2.0 -> 1.0 GB at 24 x 4608: every CTA streams
// the whole head).  A ring stage is refilled when BOTH CTAs' MMAs have consumed it (multicast commit, 2 arrivals); nothing else crosses
// the CTA boundary, in particular nothing on the T -> soft-max -> accumulate chain.
template <bool KV, bool MC>
__device__ __forceinline__ void attention_bwd_body(const CUtensorMap& tma_x0, const CUtensorMap& tma_x1, const CUtensorMap& tma_y0,
                                                   const CUtensorMap& tma_y1, const AttnBwdParams& p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NST = KV ? ABW_STAGES_KV : ABW_STAGES_Q;
  uint8_t* sx = smem;
  uint8_t* sy = smem + (KV ? ABW_OWNER_BYTES : 0);
  float* sstat = reinterpret_cast<float*>(sy + NST * ABW_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sstat) + NST * ABW_STATS_BYTES);
  uint64_t* x_full = bars;         // 1
  uint64_t* t_full = bars + 1;     // 2: T1 of a buffer is complete
  uint64_t* t2_full = bars + 3;    // 2: T2 of a buffer is complete, so the exponentials start while T2 is still running
  uint64_t* pd_full = bars + 5;    // 2 buffers x 2 halves (32 streamed rows each): the accumulating MMAs start per half
  uint64_t* acc_full = bars + 9;   // 1
  uint64_t* x_tmem = bars + 10;    // dQ launch: the 8 soft-max warps have stored the owner rows into TMEM
  uint64_t* y_full = bars + 11;    // NST
  uint64_t* y_empty = bars + 11 + NST;  // NST
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11 + 2 * NST);
  auto stg = [](int y) { return y % NST; };
  auto sph = [](int y) { return static_cast<uint32_t>((y / NST) & 1); };

  const int warp = uniform_warp_id();
  const int lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * 128;
  const int bh = blockIdx.z * p.H + blockIdx.y;
  const int n_y = (p.L + 63) / 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_x0);
    tma_prefetch_desc(&tma_x1);
    tma_prefetch_desc(&tma_y0);
    tma_prefetch_desc(&tma_y1);
    mbar_init(x_full, 1);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&y_full[i], 1);
      mbar_init(&y_empty[i], MC ? 2 : 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t2_full[i], 1);
      mbar_init(&pd_full[2 * i], 4);
      mbar_init(&pd_full[2 * i + 1], 4);
    }
    mbar_init(acc_full, 1);
    mbar_init(x_tmem, 8);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) abw_cluster_sync();  // the peer's barriers exist before any multicast box or commit can reach them
  tc_fence_after();
  if constexpr (!MC) {
    griddep_launch();  // PDL (common.cuh): the dQ launch's prologue overlaps the dK/dV launch's tail, and so on down the stream
    griddep_wait();
  }
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = MC ? abw_cluster_ctarank() : 0u;
  // TMEM columns, dK/dV launch: [0,128) T buffer 0 (T1 | T2), [128,256) T buffer 1, [256,384) acc0, [384,512) acc1
  //               dQ launch   : [0,64) Q_i bf16, [64,128) dO_i bf16 (A operands), [128,256) T buffer 0, [256,384) T buffer 1, [384,512) acc0
  constexpr uint32_t TOFF = KV ? 0u : 128u;
  constexpr uint32_t ACC0 = KV ? 256u : 384u;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      if constexpr (KV) {
        mbar_expect_tx(x_full, ABW_OWNER_BYTES);
#pragma unroll
        for (int op = 0; op < 2; ++op)
#pragma unroll
          for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int rb = 0; rb < 2; ++rb)
              tma_load_3d(sx + op * 32768 + g * 16384 + rb * 8192, op ? &tma_x1 : &tma_x0, x_full, g * 64, r0 + rb * 64, bh);
      }
      for (int y = 0; y < n_y; ++y) {
        const int stage = stg(y);
        const uint32_t ph = sph(y);
        mbar_wait(&y_empty[stage], ph ^ 1);
        mbar_expect_tx(&y_full[stage], ABW_STAGE_BYTES + (KV ? ABW_STATS_BYTES : 0));
        uint8_t* dst = sy + stage * ABW_STAGE_BYTES;
#pragma unroll
        for (int op = 0; op < 2; ++op) {
          if constexpr (MC) {
            const int g = static_cast<int>(crank);
            tma_load_3d_mc2(dst + op * 16384 + g * 8192, op ? &tma_y1 : &tma_y0, &y_full[stage], g * 64, y * 64, bh);
          } else {
#pragma unroll
            for (int g = 0; g < 2; ++g)
              tma_load_3d(dst + op * 16384 + g * 8192, op ? &tma_y1 : &tma_y0, &y_full[stage], g * 64, y * 64, bh);
          }
        }
        if constexpr (KV) {
          float* st = sstat + stage * (ABW_STATS_BYTES / 4);
          const long long off = static_cast<long long>(bh) * p.Lpad + y * 64;
          bulk_load_1d(st, p.lse + off, 256, &y_full[stage]);
          bulk_load_1d(st + 64, p.delta + off, 256, &y_full[stage]);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (whole warp, elected lane issues)
    {
      constexpr uint32_t idesc_t = make_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_acc = make_idesc_bf16(128, 128, 0, 1);
      const uint32_t x_base = smem_u32(sx);
      const uint32_t y_base = smem_u32(sy);
      // descriptors are built once; per MMA only the 14-bit start-address field moves (a 64-bit add of a constant)
      const uint64_t xdesc = make_smem_desc_sw128(x_base, 16, 1024);
      auto issue_t = [&](int y, int buf) {  // T1 = X0 Y0^T, T2 = X1 Y1^T  (128 x 64 each, K = 128); one commit per product
        const uint64_t ydesc = make_smem_desc_sw128(y_base + stg(y) * ABW_STAGE_BYTES, 16, 1024);
#pragma unroll
        for (int op = 0; op < 2; ++op) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t ox = op * 32768 + (kk >> 2) * 16384 + (kk & 3) * 32;
            const uint32_t oy = op * 16384 + (kk >> 2) * 8192 + (kk & 3) * 32;
            if constexpr (KV) umma_ss_w(tmem_base + buf * 128 + op * 64, xdesc + (ox >> 4), ydesc + (oy >> 4), idesc_t, kk != 0);
            else umma_ts_w(tmem_base + TOFF + buf * 128 + op * 64, tmem_base + op * 64 + kk * 8, ydesc + (oy >> 4), idesc_t, kk != 0);
          }
          umma_commit_w(op ? &t2_full[buf] : &t_full[buf]);
        }
      };
      auto issue_acc = [&](int y, int buf, int half) {  // acc0 += dS~ Y0 ; KV: acc1 += P~ Y1   (A from TMEM, 32 streamed rows per half)
        const uint64_t ydesc = make_smem_desc_sw128(y_base + stg(y) * ABW_STAGE_BYTES, 8192, 1024);
        const uint32_t acc = y > 0 ? 1u : 0u;
        // bf16 P~ / dS~ of half h sit in the first 16 columns of that half's 32 fp32 columns of T1 / T2 (written by warpgroup h)
        const uint32_t a_base = tmem_base + TOFF + buf * 128 + half * 32;
#pragma unroll
        for (int kk = 2 * half; kk < 2 * half + 2; ++kk) {
          umma_ts_w(tmem_base + ACC0, a_base + 64 + (kk & 1) * 8, ydesc + ((kk * 2048) >> 4), idesc_acc, kk > 0 ? 1u : acc);
          if constexpr (KV)
            umma_ts_w(tmem_base + 384, a_base + (kk & 1) * 8, ydesc + ((16384 + kk * 2048) >> 4), idesc_acc, kk > 0 ? 1u : acc);
        }
      };
      auto y_wait = [&](int y) { mbar_wait(&y_full[stg(y)], sph(y)); };

      mbar_wait(KV ? x_full : x_tmem, 0);
      y_wait(0);
      tc_fence_after();
      issue_t(0, 0);
      if (n_y > 1) {
        y_wait(1);
        tc_fence_after();
        issue_t(1, 1);
      }
      for (int y = 0; y < n_y; ++y) {
        const int buf = y & 1;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          mbar_wait(&pd_full[2 * buf + half], (y >> 1) & 1);
          tc_fence_after();
          issue_acc(y, buf, half);
        }
        if constexpr (MC) umma_commit_mc2_w(&y_empty[stg(y)]);
        else umma_commit_w(&y_empty[stg(y)]);
        if (y + 2 < n_y) {
          y_wait(y + 2);
          tc_fence_after();
          issue_t(y + 2, buf);
        }
      }
      umma_commit_w(acc_full);
    }
  } else {
    // ---------------------------------------------------------------- two soft-max warpgroups (one owner row per thread);
    // warpgroup g owns the stream tiles with y % 2 == g, i.e. T buffer g, so every SM sub-partition always has two
    // soft-max warps to interleave (TMEM-load and MUFU latency of one hides behind the other).
    const int wg = (warp - 2) >> 2;
    const int quad = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const int grow = r0 + quad * 32 + lane;  // owner row (key in KV mode, query otherwise)
    const float sc = p.scale_log2;
    float my_lse = 0.f, my_delta = 0.f;
    if constexpr (!KV) {
      if (grow < p.Lpad) {
        my_lse = p.lse[static_cast<long long>(bh) * p.Lpad + grow];
        my_delta = p.delta[static_cast<long long>(bh) * p.Lpad + grow];
      } else {
        my_lse = __int_as_float(0x7f800000);  // padding tile of a cluster pair: P~ = 0, nothing is stored
      }
    }
    if constexpr (!KV) {
      // owner rows -> TMEM as the bf16 A operand of the T products: lane = row, 32-bit column c holds elements (2c, 2c+1);
      // warpgroup 0 stores Q_i into columns [0,64), warpgroup 1 dO_i into [64,128)
      const uint4* src = reinterpret_cast<const uint4*>((wg == 0 ? p.x0g : p.x1g) + (static_cast<long long>(bh) * p.L + grow) * 128);
      const bool in = grow < p.L;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 u = in ? __ldg(src + h * 8 + i) : make_uint4(0, 0, 0, 0);
          w[4 * i] = u.x; w[4 * i + 1] = u.y; w[4 * i + 2] = u.z; w[4 * i + 3] = u.w;
        }
        tmem_st32(tmem_base + lane_off + wg * 64 + h * 32, w);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(x_tmem);
    }
    // Both warpgroups work on EVERY stream tile: warpgroup h owns the 32-column half h of T1 / T2 (streamed rows 32h .. 32h+31), so a
    // tile's soft-max latency -- which sits on the T -> soft-max -> accumulate -> T(y+2) chain of a buffer -- is half of what one
    // warpgroup per tile needs, and the exponentials of T1 start while the tensor pipe is still producing T2.
    for (int y = 0; y < n_y; ++y) {
      const int buf = y & 1;
      const uint32_t ph = (y >> 1) & 1;
      const uint32_t t1 = tmem_base + TOFF + buf * 128 + wg * 32 + lane_off, t2 = t1 + 64;
      const float* st = sstat + stg(y) * (ABW_STATS_BYTES / 4) + wg * 32;
      const int valid = p.L - y * 64 - wg * 32;  // streamed rows beyond L: zero-filled by TMA
      mbar_wait(&t_full[buf], ph);
      tc_fence_after();
#ifdef X2I_ATTN_EXPERIMENTS
      if (p.dbg == 1) {
        mbar_wait(&t2_full[buf], ph);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pd_full[2 * buf + wg]);
        continue;
      }
#endif
      uint32_t a[32], d[32];
      tmem_ld32(t1, a);
      if constexpr (KV) mbar_wait(&y_full[stg(y)], sph(y));  // lse / delta of this tile landed
      tmem_ld_wait();
      const uint64_t sc2 = pack_f32x2(sc, sc);
      uint32_t pk[16], dk[16];
      // ---- P~ = exp2(T1 * scale_log2 - lse)
#pragma unroll
      for (int k = 0; k < 32; k += 4) {
        uint64_t nl2[2];  // (-lse, -lse) pairs of the 4 streamed rows k..k+3
        if constexpr (KV) {
          const float4 lv = *reinterpret_cast<const float4*>(st + k);
          nl2[0] = pack_f32x2(-lv.x, -lv.y); nl2[1] = pack_f32x2(-lv.z, -lv.w);
        } else {
          nl2[0] = nl2[1] = pack_f32x2(-my_lse, -my_lse);
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int kk = k + 2 * e;
          float x0, x1;
          unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(a[kk]), __uint_as_float(a[kk + 1])), sc2, nl2[e]), x0, x1);
          float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
          if constexpr (!KV) {
            if (kk >= valid) p0 = 0.f;  // padded keys
            if (kk + 1 >= valid) p1 = 0.f;
          }
          a[kk] = __float_as_uint(p0);
          a[kk + 1] = __float_as_uint(p1);
          pk[kk >> 1] = pack_bf16x2(p0, p1);
        }
      }
      // ---- dS~ = P~ o (T2 - delta)
      mbar_wait(&t2_full[buf], ph);
      tc_fence_after();
      tmem_ld32(t2, d);
      tmem_ld_wait();
      // P~ / dS~ alias the first 16 columns of this warpgroup's own 32 columns of T1 / T2, all of which it has loaded by now
      if constexpr (KV) tmem_st16(t1, pk);
#pragma unroll
      for (int k = 0; k < 32; k += 4) {
        uint64_t ndl2[2];
        if constexpr (KV) {
          const float4 dv = *reinterpret_cast<const float4*>(st + 64 + k);
          ndl2[0] = pack_f32x2(-dv.x, -dv.y); ndl2[1] = pack_f32x2(-dv.z, -dv.w);
        } else {
          ndl2[0] = ndl2[1] = pack_f32x2(-my_delta, -my_delta);
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int kk = k + 2 * e;
          const uint64_t dd = add_f32x2(pack_f32x2(__uint_as_float(d[kk]), __uint_as_float(d[kk + 1])), ndl2[e]);
          float s0, s1;
          unpack_f32x2(mul_f32x2(pack_f32x2(__uint_as_float(a[kk]), __uint_as_float(a[kk + 1])), dd), s0, s1);
          dk[kk >> 1] = pack_bf16x2(s0, s1);
        }
      }
      tmem_st16(t2, dk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pd_full[2 * buf + wg]);
    }
    // ---- epilogue: accumulators -> bf16, head-major.  KV: warpgroup 0 drains dK, warpgroup 1 dV; else each takes 64 columns of dQ
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const bool ok = grow < p.L;
    const long long orow = (static_cast<long long>(bh) * p.L + grow) * 128;
    const int which = KV ? wg : 0;
    const int c_lo = KV ? 0 : wg * 2, c_hi = KV ? 4 : wg * 2 + 2;
    const float mul = which == 0 ? p.scale : 1.0f;
    __nv_bfloat16* dst = (which == 0 ? p.out0 : p.out1) + orow;
#pragma unroll 1
    for (int c = c_lo; c < c_hi; ++c) {
      uint32_t o[32];
      tmem_ld32(tmem_base + ACC0 + which * 128 + lane_off + c * 32, o);
      tmem_ld_wait();
      if (ok) st_bf16x32_scaled(dst + c * 32, o, mul);
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (MC) abw_cluster_sync();  // no multicast box or commit of the peer is still on its way into this CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <bool KV>
__global__ void __launch_bounds__(ABW_THREADS, 1)
mmdit_attention_bwd_kernel(const __grid_constant__ CUtensorMap tma_x0, const __grid_constant__ CUtensorMap tma_x1,
                           const __grid_constant__ CUtensorMap tma_y0, const __grid_constant__ CUtensorMap tma_y1,
                           const AttnBwdParams p) {
  attention_bwd_body<KV, false>(tma_x0, tma_x1, tma_y0, tma_y1, p);
}
template <bool KV>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ABW_THREADS, 1)
mmdit_attention_bwd_mc_kernel(const __grid_constant__ CUtensorMap tma_x0, const __grid_constant__ CUtensorMap tma_x1,
                              const __grid_constant__ CUtensorMap tma_y0, const __grid_constant__ CUtensorMap tma_y1,
                              const AttnBwdParams p) {
  attention_bwd_body<KV, true>(tma_x0, tma_x1, tma_y0, tma_y1, p);
}

// ---------------------------------------------------------------------------------------------------------------------------------------
// dK/dV launch, all products in the TS form (round 2, second design).  The 64-query form above keeps K_j / V_j in shared memory because
// 2 x 128 accumulator + 2 x 128 T columns leave no TMEM for them, and its SS-form T products are paced by the shared-memory port (48
// instead of 32 cycles).  With 32-query stream tiles a T buffer is 64 columns, so  [0,128) K_j | V_j as bf16 A operands, [128,256) two T
// buffers (T1 | T2 of 32 columns each), [256,512) dK | dV  fits: T products run as TS N = 32 (18 cycles measured, tools/ubench/mma_rate.cu),
// the accumulations as TS N = 128 with K = 32 (one k16 step per warpgroup half).  Per 64 queries: 2 x (16 x 18 + 4 x 65) = 1100 cycles
// against 1306 for the SS form.  Same soft-max code path: warpgroup h owns the 16-column half h of T1 / T2 of EVERY tile.
constexpr int ABK_YT = 32;
constexpr int ABK_STAGES = 8;
constexpr int ABK_STAGE_BYTES = 2 * ABK_YT * 256;   // Q | dO tile: 32 rows x 128 bf16 each, as two 64-column halves of 4 KB
constexpr int ABK_STATS_BYTES = 2 * ABK_YT * 4;     // lse[32] | delta[32]
constexpr int ABK_SMEM_BYTES = ABK_STAGES * (ABK_STAGE_BYTES + ABK_STATS_BYTES) + 256 + 1024;

__global__ void __launch_bounds__(ABW_THREADS, 1)
mmdit_attention_bwd_kv32_kernel(const __grid_constant__ CUtensorMap tma_y0 /* Q, 64 x 32 boxes */,
                                const __grid_constant__ CUtensorMap tma_y1 /* dO */, const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NST = ABK_STAGES;
  uint8_t* sy = smem;
  float* sstat = reinterpret_cast<float*>(sy + NST * ABK_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sstat) + NST * ABK_STATS_BYTES);
  uint64_t* t_full = bars;         // 2
  uint64_t* t2_full = bars + 2;    // 2
  uint64_t* pd_full = bars + 4;    // 2 buffers x 2 halves (16 streamed rows each)
  uint64_t* acc_full = bars + 8;   // 1
  uint64_t* x_tmem = bars + 9;     // 1: the 8 soft-max warps have stored K_j / V_j into TMEM
  uint64_t* y_full = bars + 10;    // NST
  uint64_t* y_empty = bars + 10 + NST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10 + 2 * NST);
  static_assert((10 + 2 * NST) * 8 + 4 <= 256, "barrier block");
  auto stg = [](int y) { return y % NST; };
  auto sph = [](int y) { return static_cast<uint32_t>((y / NST) & 1); };

  const int warp = uniform_warp_id();
  const int lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * 128;
  const int bh = blockIdx.z * p.H + blockIdx.y;
  const int n_y = (p.L + ABK_YT - 1) / ABK_YT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_y0);
    tma_prefetch_desc(&tma_y1);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&y_full[i], 1);
      mbar_init(&y_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t2_full[i], 1);
      mbar_init(&pd_full[2 * i], 4);
      mbar_init(&pd_full[2 * i + 1], 4);
    }
    mbar_init(acc_full, 1);
    mbar_init(x_tmem, 8);
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
  constexpr uint32_t TOFF = 128u, TBUF = 2u * ABK_YT, ACC0 = 256u, ACC1 = 384u;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      for (int y = 0; y < n_y; ++y) {
        const int stage = stg(y);
        mbar_wait(&y_empty[stage], sph(y) ^ 1);
        mbar_expect_tx(&y_full[stage], ABK_STAGE_BYTES + ABK_STATS_BYTES);
        uint8_t* dst = sy + stage * ABK_STAGE_BYTES;
#pragma unroll
        for (int op = 0; op < 2; ++op)
#pragma unroll
          for (int g = 0; g < 2; ++g)
            tma_load_3d(dst + op * (ABK_YT * 256) + g * (ABK_YT * 128), op ? &tma_y1 : &tma_y0, &y_full[stage], g * 64, y * ABK_YT, bh);
        float* st = sstat + stage * (ABK_STATS_BYTES / 4);
        const long long off = static_cast<long long>(bh) * p.Lpad + y * ABK_YT;
        bulk_load_1d(st, p.lse + off, ABK_YT * 4, &y_full[stage]);
        bulk_load_1d(st + ABK_YT, p.delta + off, ABK_YT * 4, &y_full[stage]);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (whole warp, elected lane issues)
    constexpr uint32_t idesc_t = make_idesc_bf16(128, ABK_YT, 0, 0);
    constexpr uint32_t idesc_acc = make_idesc_bf16(128, 128, 0, 1);
    const uint32_t y_base = smem_u32(sy);
    auto issue_t = [&](int y, int buf) {  // T1 = K_j Q^T, T2 = V_j dO^T  (128 x 32 each, K = 128), A from TMEM; one commit per product
      const uint64_t ydesc = make_smem_desc_sw128(y_base + stg(y) * ABK_STAGE_BYTES, 16, 1024);
#pragma unroll
      for (int op = 0; op < 2; ++op) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t oy = op * (ABK_YT * 256) + (kk >> 2) * (ABK_YT * 128) + (kk & 3) * 32;
          umma_ts_w(tmem_base + TOFF + buf * TBUF + op * ABK_YT, tmem_base + op * 64 + kk * 8, ydesc + (oy >> 4), idesc_t, kk != 0);
        }
        umma_commit_w(op ? &t2_full[buf] : &t_full[buf]);
      }
    };
    auto issue_acc = [&](int y, int buf, int half) {  // dK_j += dS~ Q, dV_j += P~ dO over the 16 streamed rows of this half (one k16 step)
      const uint64_t ydesc = make_smem_desc_sw128(y_base + stg(y) * ABK_STAGE_BYTES, ABK_YT * 128, 1024);
      const uint32_t acc = (y > 0 || half > 0) ? 1u : 0u;
      const uint32_t a_base = tmem_base + TOFF + buf * TBUF + half * (ABK_YT / 2);  // bf16 results sit in the first 8 columns of the half
      umma_ts_w(tmem_base + ACC0, a_base + ABK_YT, ydesc + ((half * 2048) >> 4), idesc_acc, acc);
      umma_ts_w(tmem_base + ACC1, a_base, ydesc + ((ABK_YT * 256 + half * 2048) >> 4), idesc_acc, acc);
    };
    auto y_wait = [&](int y) { mbar_wait(&y_full[stg(y)], sph(y)); };

    mbar_wait(x_tmem, 0);
    y_wait(0);
    tc_fence_after();
    issue_t(0, 0);
    if (n_y > 1) {
      y_wait(1);
      tc_fence_after();
      issue_t(1, 1);
    }
    for (int y = 0; y < n_y; ++y) {
      const int buf = y & 1;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        mbar_wait(&pd_full[2 * buf + half], (y >> 1) & 1);
        tc_fence_after();
        issue_acc(y, buf, half);
      }
      umma_commit_w(&y_empty[stg(y)]);
      if (y + 2 < n_y) {
        y_wait(y + 2);
        tc_fence_after();
        issue_t(y + 2, buf);
      }
    }
    umma_commit_w(acc_full);
  } else {
    // ---------------------------------------------------------------- two soft-max warpgroups: warpgroup h owns columns [16h, 16h + 16)
    const int wg = (warp - 2) >> 2;
    const int quad = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const int grow = r0 + quad * 32 + lane;  // owner row = key
    const float sc = p.scale_log2;
    {
      // K_j (warpgroup 0) / V_j (warpgroup 1) rows -> TMEM as the bf16 A operands of the T products
      const uint4* src = reinterpret_cast<const uint4*>((wg == 0 ? p.x0g : p.x1g) + (static_cast<long long>(bh) * p.L + grow) * 128);
      const bool in = grow < p.L;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 u = in ? __ldg(src + h * 8 + i) : make_uint4(0, 0, 0, 0);
          w[4 * i] = u.x; w[4 * i + 1] = u.y; w[4 * i + 2] = u.z; w[4 * i + 3] = u.w;
        }
        tmem_st32(tmem_base + lane_off + wg * 64 + h * 32, w);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(x_tmem);
    }
    constexpr int HC = ABK_YT / 2;  // columns (streamed rows) per warpgroup and tile
    for (int y = 0; y < n_y; ++y) {
      const int buf = y & 1;
      const uint32_t ph = (y >> 1) & 1;
      const uint32_t t1 = tmem_base + TOFF + buf * TBUF + wg * HC + lane_off, t2 = t1 + ABK_YT;
      const float* st = sstat + stg(y) * (ABK_STATS_BYTES / 4) + wg * HC;
      mbar_wait(&t_full[buf], ph);
      tc_fence_after();
      uint32_t a[HC], d[HC];
      tmem_ld16(t1, a);
      mbar_wait(&y_full[stg(y)], sph(y));  // lse / delta of this tile landed
      tmem_ld_wait();
      const uint64_t sc2 = pack_f32x2(sc, sc);
      uint32_t pk[HC / 2], dk[HC / 2];
#pragma unroll
      for (int k = 0; k < HC; k += 4) {
        const float4 lv = *reinterpret_cast<const float4*>(st + k);
        const uint64_t nl2[2] = {pack_f32x2(-lv.x, -lv.y), pack_f32x2(-lv.z, -lv.w)};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int kk = k + 2 * e;
          float x0, x1;
          unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(a[kk]), __uint_as_float(a[kk + 1])), sc2, nl2[e]), x0, x1);
          const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);  // streamed rows beyond L carry lse = +inf: P~ = 0
          a[kk] = __float_as_uint(p0);
          a[kk + 1] = __float_as_uint(p1);
          pk[kk >> 1] = pack_bf16x2(p0, p1);
        }
      }
      mbar_wait(&t2_full[buf], ph);
      tc_fence_after();
      tmem_ld16(t2, d);
      tmem_ld_wait();
      tmem_st8(t1, pk);  // P~ / dS~ alias the first 8 columns of this warpgroup's own 16 columns of T1 / T2
#pragma unroll
      for (int k = 0; k < HC; k += 4) {
        const float4 dv = *reinterpret_cast<const float4*>(st + ABK_YT + k);
        const uint64_t ndl2[2] = {pack_f32x2(-dv.x, -dv.y), pack_f32x2(-dv.z, -dv.w)};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int kk = k + 2 * e;
          const uint64_t dd = add_f32x2(pack_f32x2(__uint_as_float(d[kk]), __uint_as_float(d[kk + 1])), ndl2[e]);
          float s0, s1;
          unpack_f32x2(mul_f32x2(pack_f32x2(__uint_as_float(a[kk]), __uint_as_float(a[kk + 1])), dd), s0, s1);
          dk[kk >> 1] = pack_bf16x2(s0, s1);
        }
      }
      tmem_st8(t2, dk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pd_full[2 * buf + wg]);
    }
    // ---- epilogue: warpgroup 0 drains dK (scaled), warpgroup 1 dV
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const bool ok = grow < p.L;
    const float mul = wg == 0 ? p.scale : 1.0f;
    __nv_bfloat16* dst = (wg == 0 ? p.out0 : p.out1) + (static_cast<long long>(bh) * p.L + grow) * 128;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      tmem_ld32(tmem_base + (wg == 0 ? ACC0 : ACC1) + lane_off + c * 32, o);
      tmem_ld_wait();
      if (ok) st_bf16x32_scaled(dst + c * 32, o, mul);
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
