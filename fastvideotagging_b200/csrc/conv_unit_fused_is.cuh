// K2f (input-stationary form) — the whole (2+1)D unit of conv2_x in ONE kernel, temporal conv as ONE N = 192 MMA chain.
//
// Same contract, tiling, pair protocol and work split as conv_unit_fused.cuh (read that header first).  What changes is
// the temporal convolution.  A tcgen05.mma instruction issued by one thread has a cadence of max(64, N/2) clocks
// whatever its N (profiles/r01_mma_rate_shifted_desc.log; ncu on the output-stationary kernel: tensor pipe active 61 %
// of the time although the issue thread never idles), so the output-stationary form — 27 MMAs of N = 64 per frame —
// pays 1728 clocks for 864 clocks of tensor work.  Here every converted mid frame P[f] is multiplied ONCE by all three
// temporal taps side by side:
//     [ D[f-1] | D[f] | D[f+1] ]  +=  P[f] x [ Wt[2] | Wt[1] | Wt[0] ]          9 MMAs, M = 256, N = 192, K = 16
// The three output accumulators live in three fixed TMEM slots (frame o in slot o mod 3), so the tap that belongs in a
// slot rotates with f.  The filter is therefore stored as FIVE row blocks [W2 W1 W0 W2 W1] per 64-channel K block and
// the B descriptor starts at block (4 - f mod 3) mod 3: every cyclic rotation is a contiguous 3-block window.  With
// cta_group::2 CTA `rank` supplies N rows [96 rank, 96 rank + 96), so a CTA holds output channels [32 rank, 32 rank + 32)
// of every tap and accumulator column 96 h + 32 slot + c holds channel 32 h + c of that slot's frame.
// One MMA has one accumulate flag for all its columns: the MMAs always accumulate, and the output warps ZERO a slot
// (tcgen05.st) right after they have read a finished frame out of it — the slot's next frame starts from zero.
//
// Per frame the issue thread now spends 36 x 72 + 9 x 96 = 3456 clocks (all of it tensor work) instead of 4320.  The
// spatial accumulator S is handed back as soon as it sits in the convert warps' registers (s_empty), because only
// 864 clocks of queued temporal MMAs are left to hide the convert behind.
// TMEM columns: S [0, mid) | P0 P1 [mid, 2 mid) | D [320, 512).
// Warp roles per CTA (768 threads): warp0 slab producer, warp1 MMA issuer (leader) / relay (peer), warp2 TMEM allocator,
// warp3 filter producer, warps 4-15 convert (32 rows x mid/3 columns each), warps 16-23 output (32 rows x 32 channels).
// Replaces the Conv3D / BatchNorm / Activation / add chain of reference model/R2Plus1.py:27-38,59-62,76-81.
#pragma once
#include "conv_unit_fused.cuh"
#include "pdl.cuh"

namespace fvt {

// 4-D tiled load whose completion bytes are counted on a barrier of EITHER CTA of the pair (.cta_group::2): both CTAs' slabs
// signal the leader's barrier directly, so the MMA warp waits once per frame and no relay warp is needed.
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const void* tmap, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

constexpr int kUnitIsThreads = 768;
constexpr int kUnitIsDCol0 = 320;          // first TMEM column of the 192-column accumulator window

__global__ void __launch_bounds__(kUnitIsThreads, 1)
unit2p1_fused_is_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_ws,
                        const __grid_constant__ CUtensorMap tmap_wt, const UnitFusedParams p) {
  fvt_pdl_entry();
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
  const uint32_t rank = pair::ctarank();
  const bool leader = rank == 0;

  constexpr int kBtBlocks = 5, kBtBlockBytes = 32 * 128;
  const int kTaps = p.kh * p.kw;
  const int n_mid_half = p.n_mid >> 1;
  const int bs_slab = n_mid_half * 128;                                // one spatial tap, this CTA's filter rows
  const int bt_cb_bytes = kBtBlocks * kBtBlockBytes;                    // one 64-channel K block: [W2 W1 W0 W2 W1] x 32 rows
  uint8_t* smem_bs = smem;                                             // [9][n_mid_half x 64]
  uint8_t* smem_bt = smem_bs + kTaps * bs_slab;                        // [mid_blocks][5][32 x 64]
  uint8_t* smem_a = smem_bt + p.mid_blocks * bt_cb_bytes;               // [stages][slot]
  uint8_t* aux = smem_a + p.stages * p.slab_slot_bytes;
  uint64_t* slab_full = reinterpret_cast<uint64_t*>(aux);             // [kUnitMaxStages] leader: both CTAs' slabs have landed
  uint64_t* peer_full = slab_full + kUnitMaxStages;                    // [kUnitMaxStages] (unused)
  uint64_t* slab_empty = peer_full + kUnitMaxStages;                   // [kUnitMaxStages] multicast commit
  uint64_t* b_full = slab_empty + kUnitMaxStages;                      // [1] local filter halves landed
  uint64_t* peer_b_full = b_full + 1;                                  // [1] leader: the peer's filter halves landed
  uint64_t* s_full = peer_b_full + 1;                                  // [1] multicast commit: spatial accumulator complete
  uint64_t* s_empty = s_full + 1;                                      // [1] leader: 24 convert warps hold S in registers
  uint64_t* p_full = s_empty + 1;                                      // [2] leader: 24 convert warps wrote P (frame parity)
  uint64_t* d_full = p_full + 2;                                       // [1] multicast commit: temporal step complete
  uint64_t* d_empty = d_full + 1;                                      // [1] leader: 16 output warps drained + zeroed their slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 1);
  float* aff_mid = reinterpret_cast<float*>(tmem_slot + 4);            // scale[n_mid], shift[n_mid]
  float* aff_out = aff_mid + 2 * p.n_mid;                              // scale[n_out], shift[n_out]

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_x);
    ptx::prefetch_tensormap(&tmap_ws);
    ptx::prefetch_tensormap(&tmap_wt);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&slab_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&peer_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&slab_empty[s]), 1);
    }
    ptx::mbar_init(ptx::smem_u32(b_full), 1);
    ptx::mbar_init(ptx::smem_u32(peer_b_full), 1);
    ptx::mbar_init(ptx::smem_u32(s_full), 1);
    ptx::mbar_init(ptx::smem_u32(s_empty), 24);
    ptx::mbar_init(ptx::smem_u32(&p_full[0]), 24);
    ptx::mbar_init(ptx::smem_u32(&p_full[1]), 24);
    ptx::mbar_init(ptx::smem_u32(d_full), 1);
    ptx::mbar_init(ptx::smem_u32(d_empty), 16);
    ptx::fence_mbar_init();
  }
  if (warp == 2) pair::tmem_alloc2(ptx::smem_u32(tmem_slot), 512);
  for (int i = threadIdx.x; i < p.n_mid; i += kUnitIsThreads) {
    aff_mid[i] = __ldg(p.scale_mid + i);
    aff_mid[p.n_mid + i] = __ldg(p.shift_mid + i);
  }
  for (int i = threadIdx.x; i < p.n_out; i += kUnitIsThreads) {
    aff_out[i] = __ldg(p.scale_out + i);
    aff_out[p.n_out + i] = __ldg(p.shift_out + i);
  }
  ptx::tc_fence_before();
  __syncthreads();
  pair::cluster_sync_all();                    // barriers of both CTAs are initialised before anything arrives remotely
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t p_cols = static_cast<uint32_t>(p.n_mid) >> 1;        // TMEM columns of one bf16 mid frame
  const uint32_t p_col0 = static_cast<uint32_t>(p.n_mid);

  const bool has_work = unit::Segments(p.total_steps, p.t).g < unit::Segments(p.total_steps, p.t).g1;
  int u, tb, te, fs0, fs1;

  if (warp == 0) {
    // ===================================================== input slab producer: own row tile, frames [fs0, fs1) of each segment
    int stage = 0;
    uint32_t phase = 0;
    for (unit::Segments sg(p.total_steps, p.t); sg.next(u, tb, te, fs0, fs1);) {
      const int clip = u / p.pairs_per_frame;
      const int tile = 2 * (u - clip * p.pairs_per_frame) + static_cast<int>(rank);
      const int h0 = tile * p.r_out;            // a dummy tile (odd tiles_per_frame) starts beyond H: its rows are never stored
      for (int t = fs0; t < fs1; ++t) {
        ptx::mbar_wait(ptx::smem_u32(&slab_empty[stage]), phase ^ 1);
        const uint32_t fb = pair::map_to_rank(ptx::smem_u32(&slab_full[stage]), 0);      // the LEADER's barrier counts both slabs
        if (ptx::elect_one()) {
          if (leader) ptx::mbar_arrive_expect_tx(ptx::smem_u32(&slab_full[stage]), 2 * p.slab_tx_bytes);
          tma_load_4d_pair(ptx::smem_u32(smem_a + stage * p.slab_slot_bytes), &tmap_x, fb, 0, -p.pw, h0 - p.ph, clip * p.t + t);
          if (t + p.stages < fs1) tma_prefetch_4d(&tmap_x, 0, -p.pw, h0 - p.ph, clip * p.t + t + p.stages);     // warm L2 for the load after next
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ===================================================== filter producer: this CTA's halves of both filters, once
    if (has_work) {
      const uint32_t bb = ptx::smem_u32(b_full);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(bb, kTaps * bs_slab + p.mid_blocks * bt_cb_bytes);
        for (int tap = 0; tap < kTaps; ++tap)
          ptx::tma_load_2d(ptx::smem_u32(smem_bs + tap * bs_slab), &tmap_ws, bb, tap * 64, static_cast<int>(rank) * n_mid_half);
        for (int cb = 0; cb < p.mid_blocks; ++cb)
          for (int blk = 0; blk < kBtBlocks; ++blk) {
            const int tap = blk < 3 ? 2 - blk : 5 - blk;                   // [W2 W1 W0 W2 W1]
            // a partial last K block also fetches channels of the next tap (or zero fill): never multiplied
            ptx::tma_load_2d(ptx::smem_u32(smem_bt + cb * bt_cb_bytes + blk * kBtBlockBytes), &tmap_wt, bb,
                             tap * p.n_mid + cb * 64, static_cast<int>(rank) * 32);
          }
      }
      __syncwarp();
    }
  } else if (warp == 1 && !leader) {
    // ===================================================== relay (peer CTA): forward the filter's TMA completion to the leader
    if (has_work) {
      ptx::mbar_wait(ptx::smem_u32(b_full), 0);
      if (ptx::elect_one()) pair::remote_arrive(pair::map_to_rank(ptx::smem_u32(peer_b_full), 0));
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA): M = 256 over the pair
    const uint32_t idesc_s = ptx::make_idesc_bf16(256, p.n_mid, 0, 0);
    const uint32_t idesc_t = ptx::make_idesc_bf16(256, 192, 0, 0);
    const uint64_t bs_desc0 = ptx::make_sw128_desc(ptx::smem_u32(smem_bs), 16, 1024);
    const uint64_t bt_desc0 = ptx::make_sw128_desc(ptx::smem_u32(smem_bt), 16, 1024);
    const uint32_t bs_step = static_cast<uint32_t>(bs_slab) >> 4;
    const uint32_t bt_cb_step = static_cast<uint32_t>(bt_cb_bytes) >> 4;
    const uint32_t bt_rot_step = static_cast<uint32_t>(kBtBlockBytes) >> 4;
    const uint32_t a_row_step = static_cast<uint32_t>(p.wp) * 8u;      // one padded image row, in 16-byte units
    const uint32_t d_tmem = tmem_base + kUnitIsDCol0;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t gf = 0;                                                   // spatial frames issued so far
    uint32_t pw = 0;                                                   // p_full phases consumed (= mid frames known converted)
    uint32_t ks = 0;                                                   // temporal steps issued so far
    if (has_work) {
      ptx::mbar_wait(ptx::smem_u32(b_full), 0);
      pair::wait_cluster(ptx::smem_u32(peer_b_full), 0);
    }
    for (unit::Segments sg(p.total_steps, p.t); sg.next(u, tb, te, fs0, fs1);) {
      for (int tp = fs0; tp <= fs1; ++tp) {
        if (tp < fs1) {
          // ---- spatial conv of frame tp into S, as soon as the previous frame's S sits in the convert warps' registers
          pair::wait_cluster(ptx::smem_u32(&slab_full[stage]), phase);
          if (gf > 0) pair::wait_cluster(ptx::smem_u32(s_empty), (gf - 1u) & 1u);
          ++gf;
          ptx::tc_fence_after();
          const uint64_t a_desc0 = ptx::make_sw128_desc(ptx::smem_u32(smem_a + stage * p.slab_slot_bytes), 16, 1024);
          if (ptx::elect_one()) {
            uint32_t acc_flag = 0;
            uint64_t b_desc = bs_desc0;
            uint64_t a_row = a_desc0;
            for (int dh = 0; dh < p.kh; ++dh, a_row += a_row_step) {
              uint64_t a_tap = a_row;
              for (int dw = 0; dw < p.kw; ++dw, a_tap += 8, b_desc += bs_step) {
                pair::umma2_bf16_ss(tmem_base, a_tap, b_desc, idesc_s, acc_flag);
                acc_flag = 1;
                pair::umma2_bf16_ss(tmem_base, a_tap + 2, b_desc + 2, idesc_s, 1);
                pair::umma2_bf16_ss(tmem_base, a_tap + 4, b_desc + 4, idesc_s, 1);
                pair::umma2_bf16_ss(tmem_base, a_tap + 6, b_desc + 6, idesc_s, 1);
              }
            }
            pair::umma2_commit_both(ptx::smem_u32(s_full));
            pair::umma2_commit_both(ptx::smem_u32(&slab_empty[stage]));
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (tp > fs0) {
          // ---- temporal step of mid frame f: all three taps at once into the accumulator window
          const int f = tp - 1;
          const uint32_t need = gf - (tp < fs1 ? 1u : 0u);              // every mid frame but the one just issued
          while (pw < need) { pair::wait_cluster(ptx::smem_u32(&p_full[pw & 1u]), (pw >> 1) & 1u); ++pw; }
          pair::wait_cluster(ptx::smem_u32(d_empty), ks & 1u);          // the slot of frame f+1 has been drained and zeroed
          ++ks;
          ptx::tc_fence_after();
          const uint32_t rot = static_cast<uint32_t>((4 - f % 3) % 3);
          if (ptx::elect_one()) {
            uint32_t a_t = tmem_base + p_col0 + static_cast<uint32_t>(f & 1) * p_cols;
            uint64_t b_t = bt_desc0 + rot * bt_rot_step;
            int k16 = p.mid_k16;
            for (int cb = 0; cb < p.mid_blocks; ++cb, a_t += 32, b_t += bt_cb_step, k16 -= 4) {
              unit::umma2_bf16_ts(d_tmem, a_t, b_t, idesc_t, 1);
              if (k16 > 1) unit::umma2_bf16_ts(d_tmem, a_t + 8, b_t + 2, idesc_t, 1);
              if (k16 > 2) unit::umma2_bf16_ts(d_tmem, a_t + 16, b_t + 4, idesc_t, 1);
              if (k16 > 3) unit::umma2_bf16_ts(d_tmem, a_t + 24, b_t + 6, idesc_t, 1);
            }
            pair::umma2_commit_both(ptx::smem_u32(d_full));
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4 && warp < 16) {
    // ===================================================== convert: S -> registers (S released) -> BN -> ReLU -> bf16 -> P[t & 1]
    const int q = warp & 3;
    const int g = (warp - 4) >> 2;               // column group 0..2
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const int cg = p.n_mid / 3;                  // S columns per group (multiple of 16, <= 48)
    const int sc0 = g * cg;
    const uint32_t s_empty_leader = pair::map_to_rank(ptx::smem_u32(s_empty), 0);
    const uint32_t p_full_leader = pair::map_to_rank(ptx::smem_u32(p_full), 0);      // [2]: + 8 * (frame & 1)
    const uint32_t s_addr = tmem_base + lane_base + sc0;
    uint32_t s_phase = 0, cf = 0;
    for (unit::Segments sg(p.total_steps, p.t); sg.next(u, tb, te, fs0, fs1);) {
      for (int t = fs0; t < fs1; ++t) {
        ptx::mbar_wait(ptx::smem_u32(s_full), s_phase);
        s_phase ^= 1;
        ptx::tc_fence_after();
        uint32_t v[3][16];
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (16 * i < cg) ptx::tmem_ld_32x32b_x16(s_addr + 16 * i, v[i]);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) pair::remote_arrive(s_empty_leader);         // S is in registers: the next frame's spatial MMAs may start
        const uint32_t p_addr = tmem_base + lane_base + p_col0 + static_cast<uint32_t>(t & 1) * p_cols + (sc0 >> 1);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          if (16 * i < cg) {
            const float4* sc4 = reinterpret_cast<const float4*>(aff_mid + sc0 + 16 * i);
            const float4* sh4 = reinterpret_cast<const float4*>(aff_mid + p.n_mid + sc0 + 16 * i);
            uint32_t o8[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 a = sc4[j], b = sh4[j];
              const float f0 = fmaxf(fmaf(__uint_as_float(v[i][4 * j + 0]), a.x, b.x), 0.f);
              const float f1 = fmaxf(fmaf(__uint_as_float(v[i][4 * j + 1]), a.y, b.y), 0.f);
              const float f2 = fmaxf(fmaf(__uint_as_float(v[i][4 * j + 2]), a.z, b.z), 0.f);
              const float f3 = fmaxf(fmaf(__uint_as_float(v[i][4 * j + 3]), a.w, b.w), 0.f);
              o8[2 * j] = pack_bf16x2(f0, f1);
              o8[2 * j + 1] = pack_bf16x2(f2, f3);
            }
            unit::tmem_st_32x32b_x8(p_addr + 8 * i, o8);
          }
        }
        unit::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) pair::remote_arrive(p_full_leader + 8u * (cf & 1u));
        ++cf;
      }
    }
  } else if (warp >= 16) {
    // ===================================================== output: finished slot -> registers, slot zeroed and released, then
    //                                                       BN (+ residual) -> ReLU -> Y; 32 rows x 32 channels per warp
    const int q = warp & 3;
    const int hf = (warp - 16) >> 2;             // channel half: output channels [32 hf, 32 hf + 32)
    const bool has_res = (p.flags & kConvResidual) != 0;
    const bool no_data = (p.flags & kDbgNoEpilogue) != 0, no_store = (p.flags & kDbgNoStore) != 0;
    const int r = q * 32 + lane;                 // GEMM row = padded position inside the tile
    const int hl = r / p.wp, wl = r - hl * p.wp;
    const uint32_t d_empty_leader = pair::map_to_rank(ptx::smem_u32(d_empty), 0);
    const uint32_t d_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kUnitIsDCol0 + hf * 96;     // + 32 * slot
    uint32_t zero8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) zero8[i] = 0u;
    auto zero_slot = [&](int slot) {
#pragma unroll
      for (int i = 0; i < 4; ++i) unit::tmem_st_32x32b_x8(d_addr + 32 * slot + 8 * i, zero8);
    };
    auto release = [&]() {
      unit::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) pair::remote_arrive(d_empty_leader);
    };
    // BN (+ residual) -> ReLU -> 64 bytes of Y from 32 accumulator columns in registers
    auto finish = [&](uint32_t (&d)[2][16], const uint32_t (&rr)[2][8], size_t off) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float4* sc4 = reinterpret_cast<const float4*>(aff_out + 32 * hf + 16 * c);
        const float4* sh4 = reinterpret_cast<const float4*>(aff_out + p.n_out + 32 * hf + 16 * c);
        uint32_t o8[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 a = sc4[i], b = sh4[i];
          float f0 = fmaf(__uint_as_float(d[c][4 * i + 0]), a.x, b.x);
          float f1 = fmaf(__uint_as_float(d[c][4 * i + 1]), a.y, b.y);
          float f2 = fmaf(__uint_as_float(d[c][4 * i + 2]), a.z, b.z);
          float f3 = fmaf(__uint_as_float(d[c][4 * i + 3]), a.w, b.w);
          if (has_res) {
            f0 += bf16_lo(rr[c][2 * i]);     f1 += bf16_hi(rr[c][2 * i]);
            f2 += bf16_lo(rr[c][2 * i + 1]); f3 += bf16_hi(rr[c][2 * i + 1]);
          }
          o8[2 * i] = pack_bf16x2(fmaxf(f0, 0.f), fmaxf(f1, 0.f));
          o8[2 * i + 1] = pack_bf16x2(fmaxf(f2, 0.f), fmaxf(f3, 0.f));
        }
        if (!no_store) ptx::st_global_256(p.y + off + 16 * c, o8);
      }
    };
    if (has_work) {                              // all three slots start from zero
      zero_slot(0); zero_slot(1); zero_slot(2);
      release();
    }
    uint32_t ks = 0;
    for (unit::Segments sg(p.total_steps, p.t); sg.next(u, tb, te, fs0, fs1);) {
      const int clip = u / p.pairs_per_frame;
      const int tile = 2 * (u - clip * p.pairs_per_frame) + static_cast<int>(rank);
      const int h0 = tile * p.r_out;
      const bool ok = tile < p.tiles_per_frame && hl < p.r_out && wl < p.w && (h0 + hl) < p.h;
      const size_t row0 = ok ? ((static_cast<size_t>(clip) * p.t * p.h + h0 + hl) * p.w + wl) * p.n_out + 32 * hf : 0;    // frame 0
      const size_t frame_pitch = static_cast<size_t>(p.h) * p.w * p.n_out;
      for (int f = fs0; f < fs1; ++f, ++ks) {
        const bool last = f == fs1 - 1;
        const int o1 = f - 1;                                          // finished by this step
        // u1 / u2 are warp-uniform (the tcgen05 loads are warp-collective); `ok` is this thread's row
        const bool u1 = o1 >= tb && o1 < te && !no_data;
        const bool u2 = last && f < te && !no_data;                    // the clip's last frame has no successor: finished too
        const size_t off1 = row0 + static_cast<size_t>(u1 ? o1 : 0) * frame_pitch;
        uint32_t rr[2][8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rr[0][i] = rr[1][i] = 0u;
        if (has_res && u1 && ok) {                                     // in flight before the wait
          ptx::ld_global_nc_256(p.residual + off1, rr[0]);
          ptx::ld_global_nc_256(p.residual + off1 + 16, rr[1]);
        }
        ptx::mbar_wait(ptx::smem_u32(d_full), ks & 1u);
        ptx::tc_fence_after();
        const int s1 = (o1 + 3) % 3;
        uint32_t d[2][16];
        if (u1) {
          ptx::tmem_ld_32x32b_x16(d_addr + 32 * s1, d[0]);
          ptx::tmem_ld_32x32b_x16(d_addr + 32 * s1 + 16, d[1]);
          ptx::tmem_ld_wait();
        }
        if (!last) {
          zero_slot(s1);
          release();                                                   // the next temporal step may be issued
          if (u1 && ok) finish(d, rr, off1);
        } else {
          if (u1 && ok) finish(d, rr, off1);
          if (u2) {
            const size_t off2 = row0 + static_cast<size_t>(f) * frame_pitch;
            if (has_res && ok) {
              ptx::ld_global_nc_256(p.residual + off2, rr[0]);
              ptx::ld_global_nc_256(p.residual + off2 + 16, rr[1]);
            }
            const int s2 = f % 3;
            ptx::tmem_ld_32x32b_x16(d_addr + 32 * s2, d[0]);
            ptx::tmem_ld_32x32b_x16(d_addr + 32 * s2 + 16, d[1]);
            ptx::tmem_ld_wait();
            if (ok) finish(d, rr, off2);
          }
          zero_slot(0); zero_slot(1); zero_slot(2);                    // the next segment starts from zero
          release();
        }
      }
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
