// K1p — the generic implicit-GEMM convolution (conv_igemm.cuh, streamed weights) on a CTA PAIR: tcgen05.mma.cta_group::2,
// M = 256 (one 128-pixel im2col tile per CTA), N = block_n.
//
// Why: K1 streams an A tile (16 KB) AND a whole [block_n x 64] weight tile (24-32 KB) per k-block into every SM; on the
// wide layers (conv4_x 256 -> 576 / 576 -> 256, conv5_x) that is 95-105 B/clk/SM of L2 -> shared-memory traffic against
// 384-512 clocks of MMAs per k-block, and the layers sit at 85-93 % of the tensor peak.  With cta_group::2 each CTA
// loads its own A tile and HALF of the weight tile's rows (the tensor cores of both SMs read both halves), so the
// per-SM fill drops by 30-40 % for the same MMA work.
//
// Pair protocol (conv_slab_pair.cuh): both CTAs' TMA loads (im2col 5-D for A, tiled 2-D for B) count on the LEADER's
// full barrier (.cta_group::2 form of cp.async.bulk.tensor); the leader's MMA warp issues the M = 256 MMAs and commits
// with .multicast::cluster onto both CTAs' empty / acc_full barriers; the epilogue warps of both CTAs hand their
// accumulator back on the leader's acc_empty.  Work item = (tile pair, N tile); CTA `rank` owns M tile 2*pair + rank.
// Warp roles per CTA (384 threads): warp0 A producer, warp1 MMA issuer (leader only), warp2 TMEM allocator, warp3 B
// producer, warps 4-11 epilogue (shared with K1: epilogue.cuh).
// Replaces the same cuDNN convolution calls as K1 (reference model/R2Plus1.py:27-38,67-71, net.py:40-51).
#pragma once
#include "conv_igemm.cuh"
#include "conv_slab_pair.cuh"
#include "pdl.cuh"

namespace fvt {

namespace pair {
__device__ __forceinline__ void tma_load_im2col_5d_2sm(uint32_t dst, const void* tmap, uint32_t bar_cluster_addr, int c, int w,
                                                       int h, int d, int n, uint16_t ow, uint16_t oh, uint16_t od) {
  asm volatile(
      "cp.async.bulk.tensor.5d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], {%8, %9, %10};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c), "r"(w), "r"(h), "r"(d), "r"(n),
        "h"(ow), "h"(oh), "h"(od)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const void* tmap, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
}  // namespace pair

// p.num_m_tiles = M tiles of 128 pixels; p.k_splits must be 1 and p.b_stationary 0.
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                       const ConvKernelParams p) {
  fvt_pdl_entry();
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = pair::ctarank();
  const bool leader = rank == 0;
  const int stages = p.stages;
  const int n_half = p.block_n >> 1;
  const int b_half_bytes = n_half * kBlockK * 2;
  const int taps = p.kt * p.kh * p.kw;
  const int k_blocks = taps * p.cin_blocks;
  const int stage_bytes = kATileBytes + ((b_half_bytes + 1023) & ~1023);

  uint8_t* smem_tiles = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_tiles + stages * stage_bytes);
  uint64_t* full_bar = bars;                       // [stages] leader: A and B of both CTAs have landed
  uint64_t* empty_bar = bars + kMaxStages;         // [stages] multicast commit
  uint64_t* acc_full_bar = bars + 2 * kMaxStages;  // [2] multicast commit
  uint64_t* acc_empty_bar = acc_full_bar + 2;      // [2] leader: 16 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty_bar + 2);
  // [2][kMaxCout] scale, shift — or [4 quadrants][2][n_pad] per-CTA statistics partials (training forward), as in K1
  float* affine_smem = reinterpret_cast<float*>(tmem_slot + 4);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_x);
    ptx::prefetch_tensormap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 2);        // the leader's A and B producers arrive (with the expected bytes of both CTAs)
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&acc_full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&acc_empty_bar[s]), 16);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) pair::tmem_alloc2(ptx::smem_u32(tmem_slot), 512);
  if (p.scale != nullptr) {
    const int padded = p.num_n_tiles * p.block_n;
    for (int i = threadIdx.x; i < padded; i += kConvThreads) {
      affine_smem[i] = i < p.cout_store ? __ldg(p.scale + i) : 0.f;
      affine_smem[kMaxCout + i] = i < p.cout_store ? __ldg(p.shift + i) : 0.f;
    }
  }
  const bool bnbwd = (p.flags & kConvBnBwd) != 0;
  const bool acc_stats = (p.flags & kConvStats) != 0 && (p.scale == nullptr || bnbwd);
  const int n_pad = p.num_n_tiles * p.block_n;
  float* stat_base = bnbwd ? affine_smem + 2 * kMaxCout : affine_smem;
  if (acc_stats)
    for (int i = threadIdx.x; i < 8 * n_pad; i += kConvThreads) stat_base[i] = 0.f;
  ptx::tc_fence_before();
  __syncthreads();
  pair::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_pairs = (p.num_m_tiles + 1) / 2;
  const int num_items = num_pairs * p.num_n_tiles;
  const int item0 = static_cast<int>(pair::cluster_id_x());
  const int item_step = static_cast<int>(pair::nclusters_x());

  if (warp == 0) {
    // ===================================================== A producer: own 128-pixel im2col tile, one load per k-block
    int stage = 0;
    uint32_t phase = 0;
    for (int item = item0; item < num_items; item += item_step) {
      const int m_blk = 2 * (item / p.num_n_tiles) + static_cast<int>(rank);      // >= num_m_tiles: dummy tile, coordinates past N -> zero fill
      int m0 = m_blk * kBlockM;
      const int ow = m0 % p.wo;  m0 /= p.wo;
      const int oh = m0 % p.ho;  m0 /= p.ho;
      const int ot = m0 % p.to;
      const int on = m0 / p.to;
      const int cw = ow * p.sw - p.pw, ch = oh * p.sh - p.ph, cd = ot * p.st - p.pt;
      int cb = 0, dw = 0, dh = 0, dt = 0;
      for (int kb = 0; kb < k_blocks; ++kb) {
        ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1);
        const uint32_t fb = pair::map_to_rank(ptx::smem_u32(&full_bar[stage]), 0);
        if (ptx::elect_one()) {
          if (leader) ptx::mbar_arrive_expect_tx(ptx::smem_u32(&full_bar[stage]), 2 * kATileBytes);
          pair::tma_load_im2col_5d_2sm(ptx::smem_u32(smem_tiles + stage * stage_bytes), &tmap_x, fb, cb * kBlockK, cw, ch, cd, on,
                                       static_cast<uint16_t>(dw), static_cast<uint16_t>(dh), static_cast<uint16_t>(dt));
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
        if (++cb == p.cin_blocks) {
          cb = 0;
          if (++dw == p.kw) { dw = 0; if (++dh == p.kh) { dh = 0; ++dt; } }
        }
      }
    }
  } else if (warp == 3) {
    // ===================================================== B producer: this CTA's half of the weight tile's rows
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t b_dst0 = ptx::smem_u32(smem_tiles) + kATileBytes;
    for (int item = item0; item < num_items; item += item_step) {
      const int n0 = (item % p.num_n_tiles) * p.block_n + static_cast<int>(rank) * n_half;
      int cb = 0, tap = 0, kcoord = 0;
      for (int kb = 0; kb < k_blocks; ++kb) {
        ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1);
        const uint32_t fb = pair::map_to_rank(ptx::smem_u32(&full_bar[stage]), 0);
        if (ptx::elect_one()) {
          if (leader) ptx::mbar_arrive_expect_tx(ptx::smem_u32(&full_bar[stage]), 2 * b_half_bytes);
          pair::tma_load_2d_2sm(b_dst0 + stage * stage_bytes, &tmap_w, fb, kcoord, n0);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
        kcoord += kBlockK;
        if (++cb == p.cin_blocks) { cb = 0; ++tap; kcoord = tap * p.k_per_tap; }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================================================== MMA issuer (leader CTA): M = 256 over the pair
    const uint32_t idesc = ptx::make_idesc_bf16(256, p.block_n, 0, 0);
    const uint32_t tiles_u32 = ptx::smem_u32(smem_tiles);
    const uint32_t full_u32 = ptx::smem_u32(full_bar), empty_u32 = ptx::smem_u32(empty_bar);
    const uint64_t a_desc_s0 = ptx::make_sw128_desc(tiles_u32, 16, 1024);
    const uint32_t stage_step = static_cast<uint32_t>(stage_bytes) >> 4;
    const uint32_t b_in_stage = static_cast<uint32_t>(kATileBytes >> 4);
    const int last_k16 = p.cin_k16 - (p.cin_blocks - 1) * (kBlockK / 16);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    uint64_t a_desc = a_desc_s0;
    for (int item = item0; item < num_items; item += item_step) {
      pair::wait_cluster(ptx::smem_u32(&acc_empty_bar[acc]), acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      int cb = 0;
      for (int kb = 0; kb < k_blocks; ++kb) {
        const int k16 = (cb == p.cin_blocks - 1) ? last_k16 : kBlockK / 16;
        pair::wait_cluster(full_u32 + stage * 8, phase);
        ptx::tc_fence_after();
        const uint64_t b_desc = a_desc + b_in_stage;
        if (ptx::elect_one()) {
          pair::umma2_bf16_ss(d_tmem, a_desc, b_desc, idesc, kb != 0);
          if (k16 > 1) pair::umma2_bf16_ss(d_tmem, a_desc + 2, b_desc + 2, idesc, 1);
          if (k16 > 2) pair::umma2_bf16_ss(d_tmem, a_desc + 4, b_desc + 4, idesc, 1);
          if (k16 > 3) pair::umma2_bf16_ss(d_tmem, a_desc + 6, b_desc + 6, idesc, 1);
          pair::umma2_commit_both(empty_u32 + stage * 8);
        }
        __syncwarp();
        if (++cb == p.cin_blocks) cb = 0;
        a_desc += stage_step;
        if (++stage == stages) { stage = 0; phase ^= 1; a_desc = a_desc_s0; }
      }
      if (ptx::elect_one()) pair::umma2_commit_both(ptx::smem_u32(&acc_full_bar[acc]));
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue (own 128 accumulator rows)
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int et = threadIdx.x - 128;
    const bool has_affine = p.scale != nullptr;
    const uint32_t acc_empty_leader0 = pair::map_to_rank(ptx::smem_u32(&acc_empty_bar[0]), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = item0; item < num_items; item += item_step) {
      const int m_blk = 2 * (item / p.num_n_tiles) + static_cast<int>(rank);
      const int n0 = (item % p.num_n_tiles) * p.block_n;
      const long long row = static_cast<long long>(m_blk) * kBlockM + q * 32 + lane;
      const bool row_ok = row < p.m_total;
      EpilogueArgs ea;
      ea.block_n = p.block_n; ea.cout_store = p.cout_store; ea.flags = acc_stats ? p.flags : (p.flags & ~kConvStats);
      ea.scale_smem = has_affine ? affine_smem : nullptr; ea.shift_smem = affine_smem + kMaxCout;
      ea.residual = p.residual; ea.y = p.y;
      ea.stat_smem = stat_base + q * 2 * n_pad + n0; ea.stat_stride = n_pad;   // statistics only without a folded affine (acc_stats)
      ea.stat_mask = stat_mask_below(static_cast<long long>(m_blk) * kBlockM + q * 32, lane, p.m_total);
      epilogue_prefetch_residual(ea, n0, row_ok ? row : -1ll, grp);
      ptx::mbar_wait(ptx::smem_u32(&acc_full_bar[acc]), acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
      epilogue_chunks(ea, taddr, n0, row_ok ? row : -1ll, grp, lane);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) pair::remote_arrive(acc_empty_leader0 + acc * 8);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (acc_stats && item0 < num_items) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      flush_quadrant_stats(stat_base, n_pad, p.cout_store, p.stats, et, kEpilogueThreads);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  pair::cluster_sync_all();                    // the peer's shared memory and TMEM stay alive until every MMA has retired
  if (warp == 2) {
    ptx::tc_fence_after();
    pair::tmem_dealloc2(tmem_base, 512);
  }
}

}  // namespace fvt
