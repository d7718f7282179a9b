// K1t — "frame ring" variant of the implicit-GEMM forward kernel for stride-1 temporal convolutions (kt x 1 x 1, the
// 3x1x1 halves of every (2+1)D unit; also their data gradients, which are again kt x 1 x 1 convolutions).
//
// K1 fetches one im2col tile per filter tap, i.e. it reads every input pixel kt times from L2.  For the temporal convs
// of conv2_x (144 -> 64 channels, 4.8 M pixels at batch 48) that makes the kernel L2-request-bound: a 288-byte pixel is
// not 128-byte aligned, so each 128-byte box row costs two requests and ncu shows the request path (l1tex2xbar /
// lts tag lookups) ~70 % busy at 2.3 TB/s of DRAM traffic.  The taps of a temporal conv are the SAME 128 pixels of
// consecutive frames, so here a CTA walks one 128-pixel block of the (H*W) plane through time:
//
//   * every [128 px x 64 ch] block of every frame is loaded ONCE (tiled TMA over {C, H*W, T, N}; frames outside
//     [0, T) and pixels beyond H*W are zero-filled by TMA) into a ring of 16 KiB slots;
//   * output frame t = sum over taps dt of  frame(t - pt + dt) x W[dt]: the MMAs address the ring slots of the kt
//     resident frames; a slot is released as soon as the last output frame that needs it has been issued
//     (loop order: channel block outer, tap inner, so slots free up — and refill — while the tile is still computing);
//   * the whole filter (kt * cin_blocks tiles of [Cout x 64]) stays resident in shared memory.
//
// Warp roles (384 threads) and epilogue are those of K1 (conv_igemm.cuh / epilogue.cuh).
// Replaces cuDNN convolution calls for Conv3D(k=(3,1,1)) at reference model/R2Plus1.py:34-38,107-111, net.py:49-51,131.
#pragma once
#include "ptx.cuh"
#include "epilogue.cuh"
#include "det_sum.cuh"
#include "conv_slab.cuh"
#include "pdl.cuh"

namespace fvt {

constexpr int kRingThreads = 384;
constexpr int kRingMaxSlots = 16;
constexpr int kRingBlockBytes = 128 * 128;      // [128 px x 64 ch] bf16

struct FrameRingParams {
  int n, t, hw;
  int blocks_per_frame;       // ceil(hw / 128)
  int t_chunk, chunks_per_clip;
  int num_items;              // n * chunks_per_clip * blocks_per_frame
  int kt, pt;
  int cin_blocks, cin_k16, k_per_tap;
  int n_tile;                 // stored Cout rounded to the N tile (single N tile)
  int slots;
  int prefetch_frames;        // L2 prefetch distance in frames (0 = off)
  int cout_store, flags;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
  unsigned long long* stats;   // [2][cout_store] exact accumulators (det_sum.cuh)
};

__global__ void __launch_bounds__(kRingThreads, 1)
conv_frame_ring_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                       const FrameRingParams p) {
  fvt_pdl_entry();
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();

  const int CB = p.cin_blocks;
  const int b_tile_bytes = p.n_tile * 128;
  const int w_bytes = p.kt * CB * b_tile_bytes;
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + w_bytes;
  uint8_t* aux = smem_a + p.slots * kRingBlockBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);       // [kRingMaxSlots]
  uint64_t* empty_bar = full_bar + kRingMaxSlots;
  uint64_t* acc_full = empty_bar + kRingMaxSlots;              // [2]
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* w_full = acc_empty + 2;                            // [1] (+1 pad keeps 16-byte alignment below)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 2);
  float* affine_smem = reinterpret_cast<float*>(tmem_slot + 4);   // scale[n_tile], shift[n_tile]
  float* stat_smem = affine_smem + 2 * p.n_tile;                   // [4 quadrants][2][n_tile] per-CTA partials

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_x);
    ptx::prefetch_tensormap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.slots; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
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
    for (int i = threadIdx.x; i < p.n_tile; i += kRingThreads) {
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
  const int blocks_per_item = frames_per_item * CB;

  if (warp == 0) {
    // ===================================================== producer: filter once, then frame blocks in ring order
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
    // ring position of the next block to load (slot index and the parity of its use count), advanced incrementally:
    // the issue loops of this kernel run on one thread each, so they must stay free of integer divisions
    uint32_t slot = 0, par = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int b = item % p.blocks_per_frame;
      const int rest = item / p.blocks_per_frame;
      const int chunk = rest % p.chunks_per_clip;
      const int n = rest / p.chunks_per_clip;
      const int f0 = chunk * p.t_chunk - p.pt;
      for (int fi = 0; fi < frames_per_item; ++fi) {
        for (int cb = 0; cb < CB; ++cb) {
          ptx::mbar_wait(ptx::smem_u32(&empty_bar[slot]), par ^ 1u);
          const uint32_t fb = ptx::smem_u32(&full_bar[slot]);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(fb, kRingBlockBytes);
            tma_load_4d(ptx::smem_u32(smem_a + slot * kRingBlockBytes), &tmap_x, fb, cb * 64, b * 128, f0 + fi, n);
            // The ring holds only ~kt frames, too shallow to hide DRAM latency: warm L2 a few frames ahead (same item)
            const int f2 = f0 + fi + p.prefetch_frames;
            if (p.prefetch_frames > 0 && fi + p.prefetch_frames < frames_per_item && f2 >= 0 && f2 < p.t)
              tma_prefetch_4d(&tmap_x, cb * 64, b * 128, f2, n);
          }
          __syncwarp();
          if (++slot == static_cast<uint32_t>(p.slots)) { slot = 0; par ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (warp-uniform control flow, elected lane issues)
    const uint32_t idesc = ptx::make_idesc_bf16(128, p.n_tile, 0, 0);
    if (blockIdx.x < p.num_items) ptx::mbar_wait(ptx::smem_u32(w_full), 0);
    const uint32_t S = static_cast<uint32_t>(p.slots);
    const uint32_t smem_a_u32 = ptx::smem_u32(smem_a), smem_w_u32 = ptx::smem_u32(smem_w);
    uint32_t s0 = 0, par0 = 0;                       // slot / use-parity of block (output frame lo, cb = 0)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      for (int lo = 0; lo < p.t_chunk; ++lo) {
        ptx::mbar_wait(ptx::smem_u32(&acc_empty[acc]), acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        const bool last = lo == p.t_chunk - 1;
        for (int cb = 0; cb < CB; ++cb) {
          int k16 = p.cin_k16 - cb * 4;
          if (k16 > 4) k16 = 4;
          uint32_t slot = s0 + cb, par = par0;       // block (lo + dt, cb), dt = 0 .. kt-1: CB slots apart
          if (slot >= S) { slot -= S; par ^= 1u; }
          // wait for the blocks used here for the first time (all kt of them at the first output frame of an item,
          // afterwards only the newest frame's), then issue the group's MMAs as one straight run: the issue loop runs
          // on a single thread and must stay lean (see conv_slab.cuh)
          {
            uint32_t ws = slot, wp_ = par;
            for (int dt = 0; dt < p.kt; ++dt) {
              if (lo == 0 || dt == p.kt - 1) ptx::mbar_wait(ptx::smem_u32(&full_bar[ws]), wp_);
              ws += CB;
              if (ws >= S) { ws -= S; wp_ ^= 1u; }
            }
            ptx::tc_fence_after();
          }
          const uint64_t a_desc0 = ptx::make_sw128_desc(smem_a_u32, 16, 1024);
          uint64_t b_desc = ptx::make_sw128_desc(smem_w_u32 + cb * b_tile_bytes, 16, 1024);
          const uint32_t b_step = static_cast<uint32_t>(CB * b_tile_bytes) >> 4;
          if (ptx::elect_one()) {
            uint32_t sl = slot;
            for (int dt = 0; dt < p.kt; ++dt, b_desc += b_step) {
              const uint64_t a_desc = a_desc0 + sl * (kRingBlockBytes >> 4);
              ptx::umma_bf16_ss(d_tmem, a_desc, b_desc, idesc, (cb | dt) != 0);
              if (k16 > 1) ptx::umma_bf16_ss(d_tmem, a_desc + 2, b_desc + 2, idesc, 1);
              if (k16 > 2) ptx::umma_bf16_ss(d_tmem, a_desc + 4, b_desc + 4, idesc, 1);
              if (k16 > 3) ptx::umma_bf16_ss(d_tmem, a_desc + 6, b_desc + 6, idesc, 1);
              // last output frame of the item: no later frame needs any of these blocks, release each after its MMAs
              if (last) ptx::umma_commit(ptx::smem_u32(&empty_bar[sl]));
              sl += CB;
              if (sl >= S) sl -= S;
            }
            // otherwise only the oldest frame's block of this channel group is finished: release its slot
            if (!last) ptx::umma_commit(ptx::smem_u32(&empty_bar[slot]));
          }
          __syncwarp();
        }
        if (ptx::elect_one()) ptx::umma_commit(ptx::smem_u32(&acc_full[acc]));
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        s0 += CB;
        if (s0 >= S) { s0 -= S; par0 ^= 1u; }
      }
      // skip the kt-1 halo frames at the end of the item
      s0 += (p.kt - 1) * CB;
      if (s0 >= S) { s0 -= S; par0 ^= 1u; }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int et = threadIdx.x - 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    EpilogueArgs ea;
    ea.block_n = p.n_tile; ea.cout_store = p.cout_store; ea.flags = p.flags;
    ea.scale_smem = p.scale != nullptr ? affine_smem : nullptr; ea.shift_smem = affine_smem + p.n_tile;
    ea.residual = p.residual; ea.y = p.y; ea.stat_smem = stat_smem + q * 2 * p.n_tile; ea.stat_stride = p.n_tile;
    const int r = q * 32 + lane;
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
        epilogue_prefetch_residual(ea, 0, out_row, grp);
        ptx::mbar_wait(ptx::smem_u32(&acc_full[acc]), acc_phase);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
        epilogue_chunks(ea, taddr, 0, out_row, grp, lane);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&acc_empty[acc]));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
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
