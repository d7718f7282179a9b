// K1i — input-stationary temporal convolution (kt x 1 x 1, stride 1) for filters that fit in shared memory.
//
//   out[t] = sum_dt in[t + dt - pt] * W[dt]        <=>        in[f] contributes to out[f - dt + pt] through W[dt]
//
// K1 (im2col) fetches every input pixel kt times; the frame-ring kernel (K1t) fetches it once but must keep kt frames
// resident, which leaves no pipeline depth for the 144-channel conv2_x layers.  Here the INPUT frame is the unit of
// work: a CTA walks one 128-position block of the H*W plane through time, loads each frame's [128 x Cin] block once into
// an ordinary multi-stage pipeline (a stage is released as soon as that frame's MMAs are issued), and multiplies it by
// all kt taps, each tap accumulating into the TMEM accumulator of the output frame it belongs to.  Up to 8 output
// accumulators (S = 512 / N columns) rotate through TMEM: an output frame is complete after its last contributing
// input frame and is then drained by the epilogue warps while later frames compute.
//
// Warp roles (384 threads) and epilogue as K1.  The issue loops (one thread each) do one barrier round trip per frame:
// kt * Cin/16 MMAs per iteration.
// Replaces cuDNN convolution calls for Conv3D(k=(3,1,1)) at reference model/R2Plus1.py:34-38,107-111, net.py:49-51,131.
#pragma once
#include "ptx.cuh"
#include "epilogue.cuh"
#include "det_sum.cuh"
#include "conv_slab.cuh"
#include "pdl.cuh"

namespace fvt {

constexpr int kTisThreads = 384;
constexpr int kTisMaxStages = 6;
constexpr int kTisMaxAcc = 8;

struct TemporalIsParams {
  int n, t, hw;
  int blocks_per_frame;
  int t_chunk, chunks_per_clip;
  int num_items;
  int kt, pt;
  int cin_blocks, cin_k16, k_per_tap;
  int n_tile;                 // accumulator columns per output frame (stored Cout rounded to the N tile)
  int acc_slots;              // S: output accumulators rotating through TMEM
  int stages;
  int cout_store, flags;
  int tma_store;              // 1: output tiles leave through shared memory + TMA (n_tile == cout_store == 64)
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
  unsigned long long* stats;   // [2][cout_store] exact accumulators (det_sum.cuh)
};

constexpr int kTisOutTileBytes = 128 * 128;     // [128 positions x 64 channels] bf16 staging tile (TMA-store epilogue)

__global__ void __launch_bounds__(kTisThreads, 1)
conv_temporal_is_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                        const __grid_constant__ CUtensorMap tmap_y, const TemporalIsParams p) {
  fvt_pdl_entry();
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();

  const int CB = p.cin_blocks;
  const int b_tile_bytes = p.n_tile * 128;
  const int w_bytes = p.kt * CB * b_tile_bytes;
  const int stage_bytes = CB * (128 * 128);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + w_bytes;
  uint8_t* smem_o = smem_a + p.stages * stage_bytes;           // [2][kTisOutTileBytes] when p.tma_store
  uint8_t* aux = smem_o + (p.tma_store ? 2 * kTisOutTileBytes : 0);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);       // [kTisMaxStages]
  uint64_t* empty_bar = full_bar + kTisMaxStages;
  uint64_t* acc_full = empty_bar + kTisMaxStages;              // [kTisMaxAcc]
  uint64_t* acc_empty = acc_full + kTisMaxAcc;
  uint64_t* w_full = acc_empty + kTisMaxAcc;                   // [1] + 1 pad
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 2);
  float* affine_smem = reinterpret_cast<float*>(tmem_slot + 4);   // scale[n_tile], shift[n_tile]
  float* stat_smem = affine_smem + 2 * p.n_tile;                   // [4 quadrants][2][n_tile] per-CTA partials

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_x);
    ptx::prefetch_tensormap(&tmap_w);
    if (p.tma_store) ptx::prefetch_tensormap(&tmap_y);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < p.acc_slots; ++s) {
      ptx::mbar_init(ptx::smem_u32(&acc_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&acc_empty[s]), 8);
    }
    ptx::mbar_init(ptx::smem_u32(w_full), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512);
    ptx::tmem_relinquish();
  }
  if (p.scale != nullptr) {
    for (int i = threadIdx.x; i < p.n_tile; i += kTisThreads) {
      affine_smem[i] = i < p.cout_store ? __ldg(p.scale + i) : 0.f;
      affine_smem[p.n_tile + i] = i < p.cout_store ? __ldg(p.shift + i) : 0.f;
    }
  }
  // per-channel statistics accumulate in shared memory over ALL tiles of this CTA; one flush after the tile loop
  if (p.flags & kConvStats)
    for (int i = threadIdx.x; i < 8 * p.n_tile; i += blockDim.x) stat_smem[i] = 0.f;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int frames_per_item = p.t_chunk + p.kt - 1;

  if (warp == 0) {
    // ===================================================== producer: filter once, then one stage per input frame
    if (blockIdx.x < p.num_items) {
      const uint32_t wb = ptx::smem_u32(w_full);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(wb, w_bytes);
        for (int dt = 0; dt < p.kt; ++dt)
          for (int cb = 0; cb < CB; ++cb)
            ptx::tma_load_2d(ptx::smem_u32(smem_w + (dt * CB + cb) * b_tile_bytes), &tmap_w, wb, dt * p.k_per_tap + cb * 64, 0);
      }
      __syncwarp();
    }
    const uint32_t full_u32 = ptx::smem_u32(full_bar), empty_u32 = ptx::smem_u32(empty_bar);
    const uint32_t a_u32 = ptx::smem_u32(smem_a);
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int b = item % p.blocks_per_frame;
      const int rest = item / p.blocks_per_frame;
      const int chunk = rest % p.chunks_per_clip;
      const int n = rest / p.chunks_per_clip;
      const int f0 = chunk * p.t_chunk - p.pt;
      for (int fi = 0; fi < frames_per_item; ++fi) {
        ptx::mbar_wait(empty_u32 + stage * 8, phase ^ 1u);
        if (ptx::elect_one()) {
          const uint32_t fb = full_u32 + stage * 8;
          ptx::mbar_arrive_expect_tx(fb, stage_bytes);
          uint32_t dst = a_u32 + stage * stage_bytes;
          for (int cb = 0; cb < CB; ++cb, dst += 128 * 128) tma_load_4d(dst, &tmap_x, fb, cb * 64, b * 128, f0 + fi, n);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (warp-uniform control flow, elected lane issues)
    const uint32_t idesc = ptx::make_idesc_bf16(128, p.n_tile, 0, 0);
    if (blockIdx.x < p.num_items) ptx::mbar_wait(ptx::smem_u32(w_full), 0);
    const uint32_t full_u32 = ptx::smem_u32(full_bar), empty_u32 = ptx::smem_u32(empty_bar);
    const uint32_t accf_u32 = ptx::smem_u32(acc_full), acce_u32 = ptx::smem_u32(acc_empty);
    const uint64_t a_desc_s0 = ptx::make_sw128_desc(ptx::smem_u32(smem_a), 16, 1024);
    const uint64_t w_desc0 = ptx::make_sw128_desc(ptx::smem_u32(smem_w), 16, 1024);
    const uint32_t stage_step = static_cast<uint32_t>(stage_bytes) >> 4;
    const uint32_t blk_step = (128 * 128) >> 4;
    const uint32_t w_step = static_cast<uint32_t>(b_tile_bytes) >> 4;
    const int S = p.acc_slots;
    int stage = 0;
    uint32_t phase = 0;
    uint64_t a_stage = a_desc_s0;
    // slot_f / par_f: accumulator slot (and use parity) that output frame "lo = fi" of the current item maps to; it
    // advances with EVERY input frame (also the kt-1 trailing halo frames, which open no output) so that the older
    // outputs fed through taps dt = 1..kt-1 are always the dt preceding slots; it is rewound by kt-1 at the item's end.
    int slot_f = 0;
    uint32_t par_f = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      for (int fi = 0; fi < frames_per_item; ++fi) {
        if (fi < p.t_chunk) {
          // this frame opens output fi (tap 0 overwrites the accumulator): the slot's previous tenant must be drained
          ptx::mbar_wait(acce_u32 + slot_f * 8, par_f ^ 1u);
        }
        ptx::mbar_wait(full_u32 + stage * 8, phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          int slot = slot_f;
          uint64_t w_tap = w_desc0;
          for (int dt = 0; dt < p.kt; ++dt, w_tap += static_cast<uint32_t>(CB) * w_step) {
            const int lo = fi - dt;                            // local output frame fed through tap dt
            if (lo >= 0 && lo < p.t_chunk) {
              const uint32_t d_tmem = tmem_base + slot * p.n_tile;
              uint64_t a_d = a_stage, w_d = w_tap;
              int k16 = p.cin_k16;
              uint32_t flag = dt != 0;
              for (int cb = 0; cb < CB; ++cb, a_d += blk_step, w_d += w_step, k16 -= 4) {
                ptx::umma_bf16_ss(d_tmem, a_d, w_d, idesc, flag);
                flag = 1;
                if (k16 > 1) ptx::umma_bf16_ss(d_tmem, a_d + 2, w_d + 2, idesc, 1);
                if (k16 > 2) ptx::umma_bf16_ss(d_tmem, a_d + 4, w_d + 4, idesc, 1);
                if (k16 > 3) ptx::umma_bf16_ss(d_tmem, a_d + 6, w_d + 6, idesc, 1);
              }
              if (dt == p.kt - 1) ptx::umma_commit(accf_u32 + slot * 8);   // last contribution: output frame complete
            }
            if (--slot < 0) slot = S - 1;                      // the next older output lives one slot back
          }
          ptx::umma_commit(empty_u32 + stage * 8);
        }
        __syncwarp();
        a_stage += stage_step;
        if (++stage == p.stages) { stage = 0; phase ^= 1u; a_stage = a_desc_s0; }
        if (++slot_f == S) { slot_f = 0; par_f ^= 1u; }
      }
      for (int r = 0; r < p.kt - 1; ++r) {                     // rewind: the next item's first output follows this item's last
        if (--slot_f < 0) { slot_f = S - 1; par_f ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int et = threadIdx.x - 128;
    int slot = 0;
    uint32_t par = 0;
    EpilogueArgs ea;
    ea.block_n = p.n_tile; ea.cout_store = p.cout_store; ea.flags = p.flags;
    ea.scale_smem = p.scale != nullptr ? affine_smem : nullptr; ea.shift_smem = affine_smem + p.n_tile;
    ea.residual = p.residual; ea.y = p.y; ea.stat_smem = stat_smem + q * 2 * p.n_tile; ea.stat_stride = p.n_tile;
    const int r = q * 32 + lane;
    int obuf = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int b = item % p.blocks_per_frame;
      const int rest = item / p.blocks_per_frame;
      const int chunk = rest % p.chunks_per_clip;
      const int n = rest / p.chunks_per_clip;
      const int pos = b * 128 + r;
      ea.stat_mask = stat_mask_below(static_cast<long long>(b) * 128 + q * 32, lane, p.hw);
      for (int lo = 0; lo < p.t_chunk; ++lo) {
        const int tt = chunk * p.t_chunk + lo;
        const long long out_row = pos < p.hw ? (static_cast<long long>(n) * p.t + tt) * p.hw + pos : -1ll;
        if (p.tma_store) {
          // the TMA store that read this staging buffer two tiles ago must have finished reading it
          if (et == 0) ptx::tma_store_wait_read<1>();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          ea.stage_smem = ptx::smem_u32(smem_o + obuf * kTisOutTileBytes);
          ea.stage_row = r;
        }
        epilogue_prefetch_residual(ea, 0, out_row, grp);
        ptx::mbar_wait(ptx::smem_u32(&acc_full[slot]), par);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + slot * p.n_tile + (static_cast<uint32_t>(q * 32) << 16);
        epilogue_chunks(ea, taddr, 0, out_row, grp, lane);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&acc_empty[slot]));
        if (++slot == p.acc_slots) { slot = 0; par ^= 1u; }
        if (p.tma_store) {
          ptx::fence_proxy_async_smem();                 // generic-proxy writes -> visible to the TMA (async proxy)
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (et == 0) {
            ptx::tma_store_4d(&tmap_y, ea.stage_smem, 0, b * 128, tt, n);     // positions >= H*W are clipped by the map
            ptx::tma_store_commit();
          }
          obuf ^= 1;
        }
      }
    }
    if (p.tma_store && et == 0) ptx::tma_store_wait<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if ((p.flags & kConvStats) && static_cast<int>(blockIdx.x) < p.num_items) {
    flush_quadrant_stats(stat_smem, p.n_tile, p.cout_store, p.stats, static_cast<int>(threadIdx.x), static_cast<int>(blockDim.x), 0,
                         p.n_tile < p.cout_store ? p.n_tile : p.cout_store);
  }
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace fvt
