// K1s — "slab" variant of the implicit-GEMM forward kernel for stride-1 spatial convolutions (1 x kh x kw, the 1x3x3
// convs that carry 74 % of the R(2+1)D FLOPs).
//
// K1 fetches one im2col tile per filter tap, i.e. reads every input pixel kh*kw times from L2; with the small channel
// counts of conv2_x/conv3_x that makes the layer L2->SM-bandwidth-bound (ncu: 11.8 GB of L2 reads for 2.0 GB of HBM
// traffic).  Here each CTA loads a *slab* of R_in whole padded image rows ONCE with a tiled TMA box
// [64 ch, W+2*pw, R_in rows] (the out-of-bounds halo is zero-filled by TMA), and every filter tap (dh, dw) is the SAME
// slab addressed through a shared-memory descriptor whose start is shifted by (dh*(W+2pw) + dw) rows: with 128-byte
// swizzling the UMMA address generator swizzles on absolute smem address bits, so any 128-byte-aligned start inside a
// TMA-written slab is a valid K-major operand (verified in tools/experiments/shifted_desc.cu).
//
// GEMM rows of a tile = R_out consecutive padded image rows (positions m' = hl*Wp + w', Wp = W + 2pw); the 2pw junk
// columns per row and the rows beyond R_out*Wp are computed but never stored.  Weights: when all kh*kw*cin_blocks
// [n_tile x 64] slabs fit next to two input slabs they are loaded once and stay resident for the CTA's lifetime
// ("stationary"); otherwise they stream through a ring.
//
// Warp roles (384 threads): warp0 slab producer, warp1 MMA issuer, warp2 TMEM allocator, warp3 weight producer,
// warps 4-11 epilogue (shared with K1: epilogue.cuh).
#pragma once
#include "ptx.cuh"
#include "epilogue.cuh"
#include "det_sum.cuh"
#include "pdl.cuh"

namespace fvt {

constexpr int kSlabThreads = 384;          // 4 control warps + 8 epilogue warps
constexpr int kSlabThreadsWide = 640;      // 4 control warps + 16 epilogue warps (narrow-N layers are paced by the epilogue)
constexpr int kSlabMaxStages = 4;
constexpr int kSlabMaxBRing = 24;

struct SlabParams {
  int frames;              // N*T
  int h, w;                // image extent (output extent is the same: stride 1, 'same' padding)
  int wp;                  // padded row pitch W + 2*pw
  int ph, pw, kh, kw;
  int r_out, r_in;         // output rows per tile, slab rows loaded per tile
  int tiles_per_frame;
  int cin_blocks, cin_k16; // 64-channel blocks / 16-channel MMA steps of the input
  int k_per_tap;           // K elements per filter tap in the packed weights (= stored Cin)
  int n_tile, num_n_tiles;
  int slab_slot_bytes;     // per 64-channel block, multiple of 1024
  int slab_tx_bytes;       // bytes one slab TMA load delivers (wp * r_in * 128)
  int stages;              // slab stages
  int b_ring;              // weight ring slots
  int b_stationary;
  int box_rows;            // image rows per TMA box (r_in: one box per slab block; smaller: several boxes per block)
  int prefetch_dist;       // > 0: L2-prefetch the slab of the tile `prefetch_dist` iterations ahead
  int cout_store, flags;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
  unsigned long long* stats;   // [2][cout_store] exact accumulators (det_sum.cuh)
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// L2 prefetch of a 4-D box (no shared-memory destination, no barrier): warms L2 for a slab this CTA loads a few tiles later.
__device__ __forceinline__ void tma_prefetch_4d(const void* tmap, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

template <int kEpiWarps>
__global__ void __launch_bounds__(128 + 32 * kEpiWarps, 1)
conv_slab_fwd_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                     const SlabParams p) {
  fvt_pdl_entry();
  constexpr int kThreads = 128 + 32 * kEpiWarps;
  constexpr int kEpiThreads = 32 * kEpiWarps;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();      // swizzle atoms need a 1024-byte aligned carve-up

  const int taps = p.kh * p.kw;
  const int b_slab_bytes = p.n_tile * 128;
  const int stage_bytes = p.cin_blocks * p.slab_slot_bytes;
  uint8_t* smem_b = smem;                                              // [b_ring][n_tile x 64]
  uint8_t* smem_a = smem + p.b_ring * b_slab_bytes;                    // [stages][cin_blocks][slot]
  uint8_t* aux = smem_a + p.stages * stage_bytes;
  uint64_t* slab_full = reinterpret_cast<uint64_t*>(aux);             // [kSlabMaxStages]
  uint64_t* slab_empty = slab_full + kSlabMaxStages;
  uint64_t* b_full = slab_empty + kSlabMaxStages;                      // [kSlabMaxBRing]
  uint64_t* b_empty = b_full + kSlabMaxBRing;
  uint64_t* acc_full = b_empty + kSlabMaxBRing;                        // [2]
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  // scale[n_total], shift[n_total] — or, with kConvStats (training forward, no folded affine), the per-CTA statistics
  // partials [4 quadrants][2][n_total] (the host sizes the region accordingly)
  float* affine_smem = reinterpret_cast<float*>(tmem_slot + 4);
  const int n_total = p.n_tile * p.num_n_tiles;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_x);
    ptx::prefetch_tensormap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&slab_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&slab_empty[s]), 1);
    }
    for (int s = 0; s < p.b_ring; ++s) {
      ptx::mbar_init(ptx::smem_u32(&b_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&b_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&acc_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&acc_empty[s]), kEpiWarps);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512);
    ptx::tmem_relinquish();
  }
  if (p.scale != nullptr) {
    for (int i = threadIdx.x; i < n_total; i += kThreads) {
      affine_smem[i] = i < p.cout_store ? __ldg(p.scale + i) : 0.f;
      affine_smem[n_total + i] = i < p.cout_store ? __ldg(p.shift + i) : 0.f;
    }
  }
  // training forward (statistics, no folded affine): sums of all tiles of this CTA accumulate in the unused affine area,
  // one [2][n_total] block per TMEM lane quadrant, and are flushed to global memory once after the tile loop
  // (kConvBnBwd needs the staged scale/shift AND the partials: they then follow the affine area)
  const bool bnbwd = (p.flags & kConvBnBwd) != 0;
  const bool acc_stats = (p.flags & kConvStats) != 0 && (p.scale == nullptr || bnbwd);
  float* stat_base = bnbwd ? affine_smem + 2 * n_total : affine_smem;
  if (acc_stats)
    for (int i = threadIdx.x; i < 8 * n_total; i += kThreads) stat_base[i] = 0.f;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m_tiles = p.frames * p.tiles_per_frame;
  const int b_per_ntile = taps * p.cin_blocks;

  if (warp == 0) {
    // ===================================================== input slab producer (warp-uniform, elected lane issues)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int mt = blockIdx.x; mt < num_m_tiles; mt += gridDim.x) {
        const int frame = mt / p.tiles_per_frame;
        const int h0 = (mt - frame * p.tiles_per_frame) * p.r_out;
        ptx::mbar_wait(ptx::smem_u32(&slab_empty[stage]), phase ^ 1);
        const uint32_t fb = ptx::smem_u32(&slab_full[stage]);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(fb, p.cin_blocks * p.slab_tx_bytes);
          for (int cb = 0; cb < p.cin_blocks; ++cb)
            for (int r = 0; r < p.r_in; r += p.box_rows)
              tma_load_4d(ptx::smem_u32(smem_a + stage * stage_bytes + cb * p.slab_slot_bytes + r * p.wp * 128), &tmap_x,
                          fb, cb * 64, -p.pw, h0 - p.ph + r, frame);
          if (p.prefetch_dist > 0) {
            const int mt2 = mt + p.prefetch_dist * gridDim.x;
            if (mt2 < num_m_tiles) {
              const int frame2 = mt2 / p.tiles_per_frame;
              const int h2 = (mt2 - frame2 * p.tiles_per_frame) * p.r_out;
              for (int cb = 0; cb < p.cin_blocks; ++cb)
                for (int r = 0; r < p.r_in; r += p.box_rows)
                  tma_prefetch_4d(&tmap_x, cb * 64, -p.pw, h2 - p.ph + r, frame2);
            }
          }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ===================================================== weight producer
    {
      int slot = 0;
      uint32_t phase = 0;
      bool first = true;
      for (int mt = blockIdx.x; mt < num_m_tiles; mt += gridDim.x) {
        if (p.b_stationary && !first) break;
        // lean loop (single issuing thread): no divisions, the K coordinate of the weight slab advances incrementally
        const uint32_t bfull_u32 = ptx::smem_u32(b_full), bempty_u32 = ptx::smem_u32(b_empty);
        const uint32_t smem_b_u32 = ptx::smem_u32(smem_b);
        for (int nt = 0; nt < p.num_n_tiles; ++nt) {
          const int n0 = nt * p.n_tile;
          int k_tap = 0;
          for (int tap = 0; tap < taps; ++tap, k_tap += p.k_per_tap) {
            for (int cb = 0; cb < p.cin_blocks; ++cb) {
              ptx::mbar_wait(bempty_u32 + slot * 8, phase ^ 1);
              if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(bfull_u32 + slot * 8, b_slab_bytes);
                ptx::tma_load_2d(smem_b_u32 + slot * b_slab_bytes, &tmap_w, bfull_u32 + slot * 8, k_tap + cb * 64, n0);
              }
              __syncwarp();
              if (++slot == p.b_ring) { slot = 0; phase ^= 1; }
            }
          }
        }
        first = false;
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (warp-uniform loop, elected lane issues)
    {
      const uint32_t idesc = ptx::make_idesc_bf16(128, p.n_tile, 0, 0);
      int stage = 0, slot = 0, acc = 0;
      uint32_t phase = 0, bphase = 0, acc_phase = 0;
      bool first = true;
      for (int mt = blockIdx.x; mt < num_m_tiles; mt += gridDim.x) {
        ptx::mbar_wait(ptx::smem_u32(&slab_full[stage]), phase);
        ptx::tc_fence_after();
        const uint32_t a_base = ptx::smem_u32(smem_a + stage * stage_bytes);
        if (p.b_stationary) {
          // ---- lean issue path: the whole filter is resident, so a tile's MMAs need no barrier in between and are
          // issued as ONE straight run by the elected lane (descriptor = base + precomputed offset).  The issue loop
          // runs on a single thread: with ~100 instructions of bookkeeping per filter tap it, not the tensor pipe,
          // paced the kernel (ncu: 480 clk per 4 MMAs against a 288 clk floor).
          if (first) {
            for (int j = 0; j < b_per_ntile; ++j) ptx::mbar_wait(ptx::smem_u32(&b_full[j]), 0);
            first = false;
          }
          ptx::mbar_wait(ptx::smem_u32(&acc_empty[acc]), acc_phase ^ 1);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * 256;
          const uint64_t a_desc0 = ptx::make_sw128_desc(a_base, 16, 1024);
          const uint64_t b_desc0 = ptx::make_sw128_desc(ptx::smem_u32(smem_b), 16, 1024);
          const uint32_t a_cb_step = static_cast<uint32_t>(p.slab_slot_bytes) >> 4;
          const uint32_t b_step = static_cast<uint32_t>(b_slab_bytes) >> 4;
          const uint32_t a_row_step = static_cast<uint32_t>(p.wp) * 8u;        // one padded image row, in 16-byte units
          if (ptx::elect_one()) {
            uint32_t acc_flag = 0;
            uint64_t b_desc = b_desc0;
            uint64_t a_row = a_desc0;
            for (int dh = 0; dh < p.kh; ++dh, a_row += a_row_step) {
              uint64_t a_tap = a_row;
              for (int dw = 0; dw < p.kw; ++dw, a_tap += 8) {
                uint64_t a_desc = a_tap;
                int k16 = p.cin_k16;
                for (int cb = 0; cb < p.cin_blocks; ++cb, a_desc += a_cb_step, b_desc += b_step, k16 -= 4) {
                  ptx::umma_bf16_ss(d_tmem, a_desc, b_desc, idesc, acc_flag);
                  acc_flag = 1;
                  if (k16 > 1) ptx::umma_bf16_ss(d_tmem, a_desc + 2, b_desc + 2, idesc, 1);
                  if (k16 > 2) ptx::umma_bf16_ss(d_tmem, a_desc + 4, b_desc + 4, idesc, 1);
                  if (k16 > 3) ptx::umma_bf16_ss(d_tmem, a_desc + 6, b_desc + 6, idesc, 1);
                }
              }
            }
            ptx::umma_commit(ptx::smem_u32(&acc_full[acc]));
          }
          __syncwarp();
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        } else {
          // ---- streamed filter: one barrier per [n_tile x 64] weight slab; same lean bookkeeping (descriptors advance
          // by precomputed steps, barrier addresses are base + slot * 8)
          const uint64_t a_desc0 = ptx::make_sw128_desc(a_base, 16, 1024);
          const uint64_t b_desc_s0 = ptx::make_sw128_desc(ptx::smem_u32(smem_b), 16, 1024);
          const uint32_t a_cb_step = static_cast<uint32_t>(p.slab_slot_bytes) >> 4;
          const uint32_t b_step = static_cast<uint32_t>(b_slab_bytes) >> 4;
          const uint32_t a_row_step = static_cast<uint32_t>(p.wp) * 8u;
          const uint32_t bfull_u32 = ptx::smem_u32(b_full), bempty_u32 = ptx::smem_u32(b_empty);
          for (int nt = 0; nt < p.num_n_tiles; ++nt) {
            ptx::mbar_wait(ptx::smem_u32(&acc_empty[acc]), acc_phase ^ 1);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 256;
            uint32_t acc_flag = 0;
            uint64_t a_row = a_desc0;
            for (int dh = 0; dh < p.kh; ++dh, a_row += a_row_step) {
              uint64_t a_tap = a_row;
              for (int dw = 0; dw < p.kw; ++dw, a_tap += 8) {
                uint64_t a_desc = a_tap;
                int k16 = p.cin_k16;
                for (int cb = 0; cb < p.cin_blocks; ++cb, a_desc += a_cb_step, k16 -= 4) {
                  ptx::mbar_wait(bfull_u32 + slot * 8, bphase);
                  ptx::tc_fence_after();
                  const uint64_t b_desc = b_desc_s0 + static_cast<uint32_t>(slot) * b_step;
                  if (ptx::elect_one()) {
                    ptx::umma_bf16_ss(d_tmem, a_desc, b_desc, idesc, acc_flag);
                    if (k16 > 1) ptx::umma_bf16_ss(d_tmem, a_desc + 2, b_desc + 2, idesc, 1);
                    if (k16 > 2) ptx::umma_bf16_ss(d_tmem, a_desc + 4, b_desc + 4, idesc, 1);
                    if (k16 > 3) ptx::umma_bf16_ss(d_tmem, a_desc + 6, b_desc + 6, idesc, 1);
                    ptx::umma_commit(bempty_u32 + slot * 8);
                  }
                  __syncwarp();
                  acc_flag = 1;
                  if (++slot == p.b_ring) { slot = 0; bphase ^= 1; }
                }
              }
            }
            if (ptx::elect_one()) ptx::umma_commit(ptx::smem_u32(&acc_full[acc]));
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          }
        }
        if (ptx::elect_one()) ptx::umma_commit(ptx::smem_u32(&slab_empty[stage]));
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int et = threadIdx.x - 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    EpilogueArgs ea;
    ea.ngrp = kEpiWarps / 4;
    ea.block_n = p.n_tile; ea.cout_store = p.cout_store; ea.flags = acc_stats ? p.flags : (p.flags & ~kConvStats);
    ea.scale_smem = p.scale != nullptr ? affine_smem : nullptr; ea.shift_smem = affine_smem + n_total;
    ea.residual = p.residual; ea.y = p.y; ea.stat_smem = affine_smem; ea.stat_stride = n_total;
    const int r = q * 32 + lane;                 // GEMM row = padded position m' inside the tile
    const int hl = r / p.wp, wl = r - hl * p.wp;
    // rows whose statistics this thread gathers (16x256b fragment: q*32 + lane/4 + 8j): tile-invariant (hl, wl)
    int shl[4];
    uint32_t smask_static = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rj = q * 32 + (lane >> 2) + 8 * j;
      shl[j] = rj / p.wp;
      const int swl = rj - shl[j] * p.wp;
      if (shl[j] < p.r_out && swl < p.w) smask_static |= 1u << j;
    }
    for (int mt = blockIdx.x; mt < num_m_tiles; mt += gridDim.x) {
      const int frame = mt / p.tiles_per_frame;
      const int h0 = (mt - frame * p.tiles_per_frame) * p.r_out;
      const bool ok = hl < p.r_out && wl < p.w && (h0 + hl) < p.h;
      {
        uint32_t m = smask_static;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (h0 + shl[j] >= p.h) m &= ~(1u << j);
        ea.stat_mask = m;
      }
      const long long out_row = ok ? (static_cast<long long>(frame) * p.h + h0 + hl) * p.w + wl : -1ll;
      for (int nt = 0; nt < p.num_n_tiles; ++nt) {
        const int n0 = nt * p.n_tile;
        ea.stat_smem = stat_base + q * 2 * n_total + n0;
        epilogue_prefetch_residual(ea, n0, out_row, grp);
        ptx::mbar_wait(ptx::smem_u32(&acc_full[acc]), acc_phase);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
        epilogue_chunks(ea, taddr, n0, out_row, grp, lane);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&acc_empty[acc]));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    if (acc_stats && static_cast<int>(blockIdx.x) < num_m_tiles) {
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      flush_quadrant_stats(stat_base, n_total, p.cout_store, p.stats, et, kEpiThreads);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace fvt
