// K2f — the whole (2+1)D unit of conv2_x in ONE kernel (inference):
//   1x3x3 conv (64 -> mid) -> folded BN -> ReLU -> 3x1x1 conv (mid -> 64) -> folded BN (-> + residual) -> ReLU
// (reference model/R2Plus1.py:19-40 inside R3DBlock :42-82; symbol twin net.py:31-52).  The `mid` tensor (144 channels,
// 1.39 GB at batch 48 — written once and re-read once by the two-launch form, 60 % of the unit's HBM traffic) never
// leaves the SM, and the 288-byte-row stores that pace the stand-alone spatial kernel disappear.
//
// A CTA PAIR (tcgen05.mma.cta_group::2, M = 256: one 128-row spatial tile per CTA) walks its two row tiles through the
// T frames of one clip.  Per frame t:
//   spatial   S[128 x mid] (TMEM, fp32)  = 9 shifted views of the frame's input slab (conv_slab.cuh) x spatial filter
//   convert   8 warps per CTA: tcgen05.ld S -> scale/shift -> ReLU -> bf16 pairs -> tcgen05.st into P[t % 4]
//             (TMEM, mid/2 columns: the A-operand layout of a 16-bit K-major matrix — row = lane, two K values per column)
//   temporal  D (TMEM, fp32) for output frame t-2 = P[t-3] x Wt[0] + P[t-2] x Wt[1] + P[t-1] x Wt[2] with the A operand
//             read FROM TMEM (tcgen05.mma [d], [a_tmem], b_desc).  It trails the spatial conv by two frames so that all
//             of its operands are converted BEFORE frame t's spatial MMAs are issued: the tensor pipe has 27 queued
//             temporal MMAs to run while the convert warps turn S(t) into P[t % 4], and S is free again by the time
//             they retire.  Frames outside the clip are zero padding = skipped taps.
//   output    8 warps per CTA: D -> registers, accumulator handed back at once, then scale/shift (+ residual) -> ReLU
//             -> 64-byte row segments of Y; the drain overlaps the next frame's spatial MMAs.
// Both filters are stationary, half of their N rows per CTA (83 KB + 36 KB), next to three input-slab stages.
// TMEM columns: S [0, mid) | P0..P3 [mid, 3*mid) | D [448, 512).
//
// Pair protocol as in conv_slab_pair.cuh (own slab per CTA, relay warp, multicast commits); additionally the convert
// and output warps of BOTH CTAs arrive on the leader's p_full / d_empty barriers.
// Work split: the clusters share the num_units * T output frames evenly; a cluster that starts or ends inside a unit
// recomputes one halo frame of the spatial conv on that side (< 1 % extra work, no tail wave).
// Warp roles per CTA (640 threads): warp0 slab producer, warp1 MMA issuer (leader) / relay (peer), warp2 TMEM allocator,
// warp3 filter producer, warps 4-11 convert, warps 12-19 output (32 rows x 32 columns each).
#pragma once
#include "ptx.cuh"
#include "epilogue.cuh"
#include "conv_slab.cuh"
#include "conv_slab_pair.cuh"
#include "pdl.cuh"

namespace fvt {

constexpr int kUnitThreads = 640;
constexpr int kUnitMaxStages = 4;
constexpr int kUnitDCol0 = 448;            // first TMEM column of the output accumulator (64 columns)
constexpr int kUnitPSlots = 4;             // converted mid frames resident in TMEM

struct UnitFusedParams {
  int clips, t;              // clips, frames per clip
  int h, w, wp;              // image extent, padded row pitch W + 2 pw
  int kh, kw, ph, pw;        // spatial filter extent and its 'same' padding (the output-stationary kernel handles 3 x 3 only)
  int r_out, r_in;           // output rows per tile, slab rows loaded per tile
  int tiles_per_frame;       // row tiles per frame
  int pairs_per_frame;       // ceil(tiles_per_frame / 2): tile 2*pair + rank belongs to CTA `rank`
  int num_units;             // clips * pairs_per_frame
  long long total_steps;     // num_units * t output frames, shared evenly by the clusters
  int slab_slot_bytes, slab_tx_bytes, stages;
  int n_mid, n_out;          // stored channels of mid (multiple of 16, <= 144) and of the output (64)
  int mid_blocks, mid_k16;   // 64-channel blocks / 16-channel MMA steps of mid
  int flags;                 // kConvResidual (output epilogue); ReLU is always applied to mid and to the output
  const float* scale_mid;
  const float* shift_mid;
  const float* scale_out;
  const float* shift_out;
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
};

namespace unit {
// D[tmem] (+)= A[tmem] * B[smem] over the CTA pair: A rows = TMEM lanes of each CTA, two bf16 K values per column
__device__ __forceinline__ void umma2_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// registers -> TMEM: this warp's 32 lanes x 8 consecutive 32-bit columns (one row per thread)
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// A cluster's share [g, g1) of the global output-frame sequence (unit-major), cut into per-unit segments.
struct Segments {
  long long g, g1;
  int t;
  __device__ __forceinline__ Segments(long long total, int frames_per_unit) : t(frames_per_unit) {
    const long long c = pair::cluster_id_x(), n = pair::nclusters_x();
    g = total * c / n;
    g1 = total * (c + 1) / n;
  }
  // next segment: unit u, output frames [tb, te), spatial (mid) frames [fs0, fs1) = the outputs' frames plus the halo inside the clip
  __device__ __forceinline__ bool next(int& u, int& tb, int& te, int& fs0, int& fs1) {
    if (g >= g1) return false;
    u = static_cast<int>(g / t);
    const long long base = static_cast<long long>(u) * t;
    tb = static_cast<int>(g - base);
    const long long end = g1 < base + t ? g1 : base + t;
    te = static_cast<int>(end - base);
    g = end;
    fs0 = tb > 0 ? tb - 1 : 0;
    fs1 = te < t ? te + 1 : t;
    return true;
  }
};
}  // namespace unit

__global__ void __launch_bounds__(kUnitThreads, 1)
unit2p1_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_ws,
                     const __grid_constant__ CUtensorMap tmap_wt, const UnitFusedParams p) {
  fvt_pdl_entry();
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
  const uint32_t rank = pair::ctarank();
  const bool leader = rank == 0;

  constexpr int kTaps = 9, kKt = 3;
  const int n_mid_half = p.n_mid >> 1, n_out_half = p.n_out >> 1;
  const int bs_slab = n_mid_half * 128;                                // one spatial tap, this CTA's filter rows
  const int bt_slab = n_out_half * 128;                                // one (temporal tap, 64-channel block)
  uint8_t* smem_bs = smem;                                             // [9][n_mid_half x 64]
  uint8_t* smem_bt = smem_bs + kTaps * bs_slab;                        // [3][mid_blocks][n_out_half x 64]
  uint8_t* smem_a = smem_bt + ((kKt * p.mid_blocks * bt_slab + 1023) & ~1023);      // [stages][slot]
  uint8_t* aux = smem_a + p.stages * p.slab_slot_bytes;
  uint64_t* slab_full = reinterpret_cast<uint64_t*>(aux);             // [kUnitMaxStages] local TMA completion
  uint64_t* peer_full = slab_full + kUnitMaxStages;                    // [kUnitMaxStages] leader: the peer's slab has landed
  uint64_t* slab_empty = peer_full + kUnitMaxStages;                   // [kUnitMaxStages] multicast commit
  uint64_t* b_full = slab_empty + kUnitMaxStages;                      // [1] local filter halves landed
  uint64_t* peer_b_full = b_full + 1;                                  // [1] leader: the peer's filter halves landed
  uint64_t* s_full = peer_b_full + 1;                                  // [1] multicast commit: spatial accumulator complete
  uint64_t* p_full = s_full + 1;                                       // [1] leader: 16 convert warps (S drained, P written)
  uint64_t* d_full = p_full + 1;                                       // [1] multicast commit: output accumulator complete
  uint64_t* d_empty = d_full + 1;                                      // [1] leader: 16 output warps
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
    ptx::mbar_init(ptx::smem_u32(p_full), 16);
    ptx::mbar_init(ptx::smem_u32(d_full), 1);
    ptx::mbar_init(ptx::smem_u32(d_empty), 16);
    ptx::fence_mbar_init();
  }
  if (warp == 2) pair::tmem_alloc2(ptx::smem_u32(tmem_slot), 512);
  for (int i = threadIdx.x; i < p.n_mid; i += kUnitThreads) {
    aff_mid[i] = __ldg(p.scale_mid + i);
    aff_mid[p.n_mid + i] = __ldg(p.shift_mid + i);
  }
  for (int i = threadIdx.x; i < p.n_out; i += kUnitThreads) {
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
    // ===================================================== input slab producer: own row tile, every frame of the clip
    int stage = 0;
    uint32_t phase = 0;
    for (unit::Segments sg(p.total_steps, p.t); sg.next(u, tb, te, fs0, fs1);) {
      const int clip = u / p.pairs_per_frame;
      const int tile = 2 * (u - clip * p.pairs_per_frame) + static_cast<int>(rank);
      const int h0 = tile * p.r_out;            // a dummy tile (odd tiles_per_frame) starts beyond H: its rows are never stored
      for (int t = fs0; t < fs1; ++t) {
        ptx::mbar_wait(ptx::smem_u32(&slab_empty[stage]), phase ^ 1);
        const uint32_t fb = ptx::smem_u32(&slab_full[stage]);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(fb, p.slab_tx_bytes);
          tma_load_4d(ptx::smem_u32(smem_a + stage * p.slab_slot_bytes), &tmap_x, fb, 0, -1, h0 - 1, clip * p.t + t);
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
        ptx::mbar_arrive_expect_tx(bb, kTaps * bs_slab + kKt * p.mid_blocks * bt_slab);
        for (int tap = 0; tap < kTaps; ++tap)
          ptx::tma_load_2d(ptx::smem_u32(smem_bs + tap * bs_slab), &tmap_ws, bb, tap * 64, static_cast<int>(rank) * n_mid_half);
        int j = 0;
        for (int dt = 0; dt < kKt; ++dt)
          for (int cb = 0; cb < p.mid_blocks; ++cb, ++j)       // a partial last block also fetches channels of the next tap: never multiplied
            ptx::tma_load_2d(ptx::smem_u32(smem_bt + j * bt_slab), &tmap_wt, bb, dt * p.n_mid + cb * 64,
                             static_cast<int>(rank) * n_out_half);
      }
      __syncwarp();
    }
  } else if (warp == 1 && !leader) {
    // ===================================================== relay (peer CTA): forward local TMA completions to the leader
    if (has_work) {
      ptx::mbar_wait(ptx::smem_u32(b_full), 0);
      if (ptx::elect_one()) pair::remote_arrive(pair::map_to_rank(ptx::smem_u32(peer_b_full), 0));
      __syncwarp();
    }
    int stage = 0;
    uint32_t phase = 0;
    for (unit::Segments sg(p.total_steps, p.t); sg.next(u, tb, te, fs0, fs1);) {
      for (int t = fs0; t < fs1; ++t) {
        ptx::mbar_wait(ptx::smem_u32(&slab_full[stage]), phase);
        if (ptx::elect_one()) pair::remote_arrive(pair::map_to_rank(ptx::smem_u32(&peer_full[stage]), 0));
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA): M = 256 over the pair
    const uint32_t idesc_s = ptx::make_idesc_bf16(256, p.n_mid, 0, 0);
    const uint32_t idesc_t = ptx::make_idesc_bf16(256, p.n_out, 0, 0);
    const uint64_t bs_desc0 = ptx::make_sw128_desc(ptx::smem_u32(smem_bs), 16, 1024);
    const uint64_t bt_desc0 = ptx::make_sw128_desc(ptx::smem_u32(smem_bt), 16, 1024);
    const uint32_t bs_step = static_cast<uint32_t>(bs_slab) >> 4;
    const uint32_t bt_step = static_cast<uint32_t>(bt_slab) >> 4;
    const uint32_t bt_tap_step = bt_step * static_cast<uint32_t>(p.mid_blocks);
    const uint32_t a_row_step = static_cast<uint32_t>(p.wp) * 8u;      // one padded image row, in 16-byte units
    int stage = 0;
    uint32_t phase = 0, p_phase = 0;
    uint32_t go = 0;                                                   // output frames issued so far
    if (has_work) {
      ptx::mbar_wait(ptx::smem_u32(b_full), 0);
      pair::wait_cluster(ptx::smem_u32(peer_b_full), 0);
    }
    const uint32_t d_tmem = tmem_base + kUnitDCol0;
    for (unit::Segments sg(p.total_steps, p.t); sg.next(u, tb, te, fs0, fs1);) {
      for (int tp = fs0; tp < fs1 + 2; ++tp) {
        if (tp < fs1) {
          // ---- spatial conv of frame tp into S (S is free: the p_full wait of the previous frame covered its drain)
          ptx::mbar_wait(ptx::smem_u32(&slab_full[stage]), phase);
          pair::wait_cluster(ptx::smem_u32(&peer_full[stage]), phase);
          ptx::tc_fence_after();
          const uint64_t a_desc0 = ptx::make_sw128_desc(ptx::smem_u32(smem_a + stage * p.slab_slot_bytes), 16, 1024);
          if (ptx::elect_one()) {
            uint32_t acc_flag = 0;
            uint64_t b_desc = bs_desc0;
            uint64_t a_row = a_desc0;
            for (int dh = 0; dh < 3; ++dh, a_row += a_row_step) {
              uint64_t a_tap = a_row;
#pragma unroll
              for (int dw = 0; dw < 3; ++dw, a_tap += 8, b_desc += bs_step) {
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
        const int o = tp - 2;                                          // output frame whose three mid frames are all converted
        if (o >= tb && o < te) {
          pair::wait_cluster(ptx::smem_u32(d_empty), (go & 1u) ^ 1u);   // the output warps of both CTAs drained frame o-1
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            uint32_t acc_flag = 0;
            uint64_t b_tap = bt_desc0;
#pragma unroll
            for (int dt = 0; dt < 3; ++dt, b_tap += bt_tap_step) {
              const int f = o - 1 + dt;
              if (f < 0 || f >= p.t) continue;                          // temporal zero padding
              uint32_t a_t = tmem_base + p_col0 + static_cast<uint32_t>(f & (kUnitPSlots - 1)) * p_cols;
              uint64_t b_t = b_tap;
              int k16 = p.mid_k16;
              for (int cb = 0; cb < p.mid_blocks; ++cb, a_t += 32, b_t += bt_step, k16 -= 4) {
                unit::umma2_bf16_ts(d_tmem, a_t, b_t, idesc_t, acc_flag);
                acc_flag = 1;
                if (k16 > 1) unit::umma2_bf16_ts(d_tmem, a_t + 8, b_t + 2, idesc_t, 1);
                if (k16 > 2) unit::umma2_bf16_ts(d_tmem, a_t + 16, b_t + 4, idesc_t, 1);
                if (k16 > 3) unit::umma2_bf16_ts(d_tmem, a_t + 24, b_t + 6, idesc_t, 1);
              }
            }
            pair::umma2_commit_both(ptx::smem_u32(d_full));
          }
          __syncwarp();
          ++go;
        }
        if (tp < fs1) {
          pair::wait_cluster(ptx::smem_u32(p_full), p_phase);           // frame tp: S drained, P[tp % 4] written (both CTAs)
          p_phase ^= 1;
          ptx::tc_fence_after();
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================================================== convert: S (fp32) -> BN -> ReLU -> bf16 -> P[t % 4], own 128 rows
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t p_full_leader = pair::map_to_rank(ptx::smem_u32(p_full), 0);
    const int n_chunks = p.n_mid >> 4;
    uint32_t s_phase = 0;
    for (unit::Segments sg(p.total_steps, p.t); sg.next(u, tb, te, fs0, fs1);) {
      for (int t = fs0; t < fs1; ++t) {
        ptx::mbar_wait(ptx::smem_u32(s_full), s_phase);
        s_phase ^= 1;
        ptx::tc_fence_after();
        const uint32_t s_addr = tmem_base + lane_base;
        const uint32_t p_addr = tmem_base + lane_base + p_col0 + static_cast<uint32_t>(t & (kUnitPSlots - 1)) * p_cols;
        uint32_t v[16], vn[16];
        int ci = grp;
        if (ci < n_chunks) ptx::tmem_ld_32x32b_x16(s_addr + ci * 16, vn);
        for (; ci < n_chunks; ci += 2) {
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = vn[i];
          if (ci + 2 < n_chunks) ptx::tmem_ld_32x32b_x16(s_addr + (ci + 2) * 16, vn);
          const float4* sc4 = reinterpret_cast<const float4*>(aff_mid + ci * 16);
          const float4* sh4 = reinterpret_cast<const float4*>(aff_mid + p.n_mid + ci * 16);
          float f[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 a = sc4[i], b = sh4[i];
            f[4 * i + 0] = fmaxf(fmaf(__uint_as_float(v[4 * i + 0]), a.x, b.x), 0.f);
            f[4 * i + 1] = fmaxf(fmaf(__uint_as_float(v[4 * i + 1]), a.y, b.y), 0.f);
            f[4 * i + 2] = fmaxf(fmaf(__uint_as_float(v[4 * i + 2]), a.z, b.z), 0.f);
            f[4 * i + 3] = fmaxf(fmaf(__uint_as_float(v[4 * i + 3]), a.w, b.w), 0.f);
          }
          uint32_t o8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o8[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
          unit::tmem_st_32x32b_x8(p_addr + ci * 8, o8);
        }
        unit::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) pair::remote_arrive(p_full_leader);
      }
    }
  } else if (warp >= 12) {
    // ===================================================== output: D -> BN (+ residual) -> ReLU -> Y; 32 rows x 32 columns per warp
    const int q = warp & 3;
    const int c0 = ((warp - 12) >> 2) * 32;      // this warp's first output channel
    const bool has_res = (p.flags & kConvResidual) != 0;
    const bool no_data = (p.flags & kDbgNoEpilogue) != 0, no_store = (p.flags & kDbgNoStore) != 0;
    const int r = q * 32 + lane;                 // GEMM row = padded position inside the tile
    const int hl = r / p.wp, wl = r - hl * p.wp;
    const uint32_t d_empty_leader = pair::map_to_rank(ptx::smem_u32(d_empty), 0);
    const uint32_t taddr = tmem_base + kUnitDCol0 + c0 + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t go = 0;
    for (unit::Segments sg(p.total_steps, p.t); sg.next(u, tb, te, fs0, fs1);) {
      const int clip = u / p.pairs_per_frame;
      const int tile = 2 * (u - clip * p.pairs_per_frame) + static_cast<int>(rank);
      const int h0 = tile * p.r_out;
      const bool ok = tile < p.tiles_per_frame && hl < p.r_out && wl < p.w && (h0 + hl) < p.h;
      for (int t = tb; t < te; ++t, ++go) {
        const size_t off = ok ? (((static_cast<size_t>(clip) * p.t + t) * p.h + h0 + hl) * p.w + wl) * p.n_out + c0 : 0;
        uint32_t rr[2][8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rr[0][i] = rr[1][i] = 0u;
        if (has_res && ok && !no_data) {           // the residual does not depend on the accumulator: in flight before the wait
          ptx::ld_global_nc_256(p.residual + off, rr[0]);
          ptx::ld_global_nc_256(p.residual + off + 16, rr[1]);
        }
        ptx::mbar_wait(ptx::smem_u32(d_full), go & 1u);
        ptx::tc_fence_after();
        uint32_t v[2][16];
        if (!no_data) {
          ptx::tmem_ld_32x32b_x16(taddr, v[0]);
          ptx::tmem_ld_32x32b_x16(taddr + 16, v[1]);
          ptx::tmem_ld_wait();
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) pair::remote_arrive(d_empty_leader);       // the accumulator is in registers: hand it back before the math
        if (no_data || !ok) continue;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float4* sc4 = reinterpret_cast<const float4*>(aff_out + c0 + 16 * c);
          const float4* sh4 = reinterpret_cast<const float4*>(aff_out + p.n_out + c0 + 16 * c);
          float f[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 a = sc4[i], b = sh4[i];
            f[4 * i + 0] = fmaf(__uint_as_float(v[c][4 * i + 0]), a.x, b.x);
            f[4 * i + 1] = fmaf(__uint_as_float(v[c][4 * i + 1]), a.y, b.y);
            f[4 * i + 2] = fmaf(__uint_as_float(v[c][4 * i + 2]), a.z, b.z);
            f[4 * i + 3] = fmaf(__uint_as_float(v[c][4 * i + 3]), a.w, b.w);
          }
          if (has_res) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              f[2 * i] += bf16_lo(rr[c][i]);
              f[2 * i + 1] += bf16_hi(rr[c][i]);
            }
          }
          uint32_t o8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o8[i] = pack_bf16x2(fmaxf(f[2 * i], 0.f), fmaxf(f[2 * i + 1], 0.f));
          if (!no_store) ptx::st_global_256(p.y + off + 16 * c, o8);
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
