// K3s — "slab" weight gradient for stride-1 'same' spatial convolutions (1 x kh x kw), the layers that carry two thirds
// of the weight-gradient FLOPs of R(2+1)D:
//
//     dW[co, ci, tap] += sum over output positions p of  dY[p, co] * X[p + shift(tap), ci]
//
// K3 (conv_wgrad.cuh) fetches one im2col slab of X per filter tap and owns one 128-row accumulator, so it re-reads X
// kh*kw times and dY once per M tile from L2 (measured 7x off the tensor roofline on conv2_x).  Here a CTA
//   * loads, per pixel tile, each X slab ONCE as R_in whole zero-padded image rows (tiled TMA, halo by OOB fill — the
//     same geometry as the forward slab kernel, conv_slab.cuh) plus the matching dY tile in padded-row indexing
//     (box [64 ch, W+2pw, R_out]: the 2pw junk columns and rows beyond H come back as zeros, so they add nothing);
//   * treats both as MN-major UMMA operands (GEMM-K = positions).  A filter tap is the SAME slab addressed through a
//     descriptor shifted by (dh*(W+2pw)+dw) rows, and ONE M=128 instruction covers TWO (tap, 64-channel) groups: the
//     descriptor's leading-dimension byte offset is simply the distance between the two shifted views
//     (tools/experiments/mn_stack.cu verifies this on hardware);
//   * keeps up to 512 TMEM columns of fp32 accumulators (several 128-row M tiles x one N tile) for its whole pixel
//     range, so X and dY are read once per CTA and the epilogue runs once: plain stores of the fp32 tile into dW (one
//     pixel split) or into this split's slice of the workspace (reduced in split order afterwards: deterministic).
//
// Warp roles (192 threads): warp0 TMA producer, warp1 MMA issuer (+ TMEM alloc), warps 2-5 epilogue.
// Replaces cuDNN backward-filter for the Conv3D(1,3,3) layers at reference model/R2Plus1.py:27-31, net.py:40-42.
#pragma once
#include "ptx.cuh"
#include "conv_slab.cuh"
#include "pdl.cuh"

namespace fvt {

constexpr int kWgsThreads = 192;
constexpr int kWgsMaxStages = 4;
constexpr int kWgsMaxMt = 5;            // 128-row accumulator tiles per CTA

struct WgradSlabParams {
  int frames, h, w, wp;
  int ph, pw, kh, kw;
  int r_out, r_in, tiles_per_frame, num_tiles;
  int ksteps;                 // 16-position MMA steps per pixel tile (= ceil(r_out*wp / 16))
  int taps, cin_blocks, groups;   // groups = cin_blocks * taps, ordered cb-major: g = cb*taps + tap
  int mt_per_cta, m_chunks;
  int n_tile, n_tiles, n_blocks, acc_stride;
  int ncb_max;                // X slabs per stage
  int slab_slot_bytes, slab_tx_bytes, dy_tx_bytes;
  int stage_bytes, stages;
  int splits, tiles_per_split;
  int cin_real, cout_real;
  // temporal mode (kt x 1 x 1 convs): a pixel tile = 128 positions of one output frame; a CTA's M chunk = (tap, range of
  // 64-channel blocks); the tap's input frame blocks are plain [128 x 64] TMA boxes (no shifted views: shift = 0)
  int temporal;
  int hw, t_frames, blocks_per_frame, kt, pt, chunks_per_tap;
  // where the input of filter tap k sits relative to an output tile: (k - pt) * tap_frames frames and (k - pt) * tap_pos
  // positions further.  Per-frame tiles: (1, 0), maps {C, H*W, T, N}.  Flattened tiles (a tile = 128 consecutive positions of
  // a clip's T*H*W, no padding at frame ends): (0, H*W), maps {C, T*H*W, 1, N} — positions outside the clip come back as
  // zeros, which is exactly the temporal zero padding.
  int tap_frames, tap_pos;
  int t_stride;                // temporal stride of the convolution (per-frame tiles only): input frame = t * t_stride + ...
  // taps_on_n (temporal, per-frame tiles, stride 1, <= 64 output channels: conv2_x 144 -> 64, the stem's 45 -> 64): the
  // filter taps move from the M side to the N side.  A tile loads the INPUT frame block once (all channel blocks) and the
  // kt OUTPUT-gradient frame blocks t+pt, t+pt-1, .. that pair with it, side by side as one N = kt*64 operand (the
  // descriptor's leading-dimension offset steps from block to block): dW[tap][ci, co] += X[t]^T dY[t + pt - tap].
  // One N = 192 instruction instead of three N = 64 ones (a tcgen05.mma costs max(64, N/2) clocks), and X — the large
  // operand — is read once instead of once per tap.
  int taps_on_n;
  float* dw;
  int w_ohwi;                  // dw layout (O, taps, I) instead of (O, I, taps)
  int dbg_no_store;            // experiments only (fvt_set_option("wgrad_no_store")): epilogue reads TMEM, stores nothing
  // Split reduction: with splits > 1 every pixel split STORES its partial gradient into its own dW-shaped slice
  // ws[split][...] of the caller's workspace (full 128-byte lines, no read-modify-write in L2) and one reduce pass adds
  // the slices in split order and overwrites dw; with a single split the tile is stored straight into dw.  No atomics.
  float* ws;                   // nullptr: one split, store into dw
  long long ws_split_stride;   // elements of one slice (= cout_real * cin_real * taps)
};

// One work item (pixel split, M chunk, N tile) of one layer.  `global_maps`: the tensor maps live in global memory (a
// group table written by the host) instead of the kernel's parameter space.
__device__ __forceinline__ void wgrad_slab_body(const CUtensorMap* tmap_x_p, const CUtensorMap* tmap_dy_p, const WgradSlabParams& p_in,
                                                int item, uint8_t* smem, bool global_maps) {
  // A private copy: in the grouped kernel the parameters sit in shared memory behind a generic reference, and every global
  // store of the epilogue (which may alias anything, as far as the compiler knows) forced the fields it uses to be re-read
  // — a ~50-clock dependent chain per store (measured: conv5_x group 275 us, 85 us with the stores compiled out).
  const WgradSlabParams p = p_in;
  const CUtensorMap& tmap_x = *tmap_x_p;
  const CUtensorMap& tmap_dy = *tmap_dy_p;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();

  uint8_t* aux = smem + p.stages * p.stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);          // [kWgsMaxStages]
  uint64_t* empty_bar = full_bar + kWgsMaxStages;
  uint64_t* acc_bar = empty_bar + kWgsMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  // ---- work item: (pixel split, M chunk, N tile)
  const int nt = item % p.n_tiles;     item /= p.n_tiles;
  const int chunk = item % p.m_chunks; item /= p.m_chunks;
  const int split = item;
  // spatial: groups g = cb*taps + tap over the whole filter;  temporal: this CTA's tap is fixed, groups = channel blocks
  const int tap_c = (p.temporal && !p.taps_on_n) ? chunk / p.chunks_per_tap : 0;      // the tap this CTA's chunk index encodes
  const int tap_t = p.taps_on_n ? p.pt : tap_c;                                        // the tap whose input frame it loads
  const int n_groups = p.temporal ? p.cin_blocks : p.groups;
  const int g_lo = (p.temporal ? chunk - tap_c * p.chunks_per_tap : chunk) * 2 * p.mt_per_cta;
  int g_hi = g_lo + 2 * p.mt_per_cta;
  if (g_hi > n_groups) g_hi = n_groups;
  const int mt_count = (g_hi - g_lo + 1) >> 1;
  const int cb_lo = p.temporal ? g_lo : g_lo / p.taps;
  const int cb_hi = p.temporal ? g_hi - 1 : (g_hi - 1) / p.taps;
  const int ncb = cb_hi - cb_lo + 1;
  const int tile0 = split * p.tiles_per_split;
  int tile1 = tile0 + p.tiles_per_split;
  if (tile1 > p.num_tiles) tile1 = p.num_tiles;

  // Zero the operand stages once: TMA only ever writes the boxes, so the slab tails (rows past R_in*Wp that shifted
  // taps reach) and the dY rows past R_out*Wp stay zero and contribute nothing (stale NaN bit patterns would).
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = p.stages * p.stage_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += kWgsThreads) z[i] = make_uint4(0, 0, 0, 0);
    ptx::fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    if (global_maps) {            // descriptors written by the host into global memory: make them visible to the TMA unit
      ptx::fence_tensormap_acquire(&tmap_x);
      ptx::fence_tensormap_acquire(&tmap_dy);
    }
    ptx::prefetch_tensormap(&tmap_x);
    ptx::prefetch_tensormap(&tmap_dy);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    ptx::mbar_init(ptx::smem_u32(acc_bar), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int dy_off = p.ncb_max * p.slab_slot_bytes;      // dY blocks follow the X slabs inside a stage

  // (Measured: issuing the dY boxes from a second warp — as the generic kernel K3 now does — changes nothing here: the
  // boxes are 15-30 KB, the loop is not bound by its issuing thread.)
  if (warp == 0) {
    // ===================================================== producer
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx = ncb * p.slab_tx_bytes + p.n_blocks * p.dy_tx_bytes;
    for (int tile = tile0; tile < tile1; ++tile) {
      const int frame = tile / p.tiles_per_frame;
      const int h0 = (tile - frame * p.tiles_per_frame) * p.r_out;
      ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1);
      const uint32_t fb = ptx::smem_u32(&full_bar[stage]);
      const uint32_t base = ptx::smem_u32(smem + stage * p.stage_bytes);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(fb, tx);
        if (!p.temporal) {
          for (int c = 0; c < ncb; ++c)
            tma_load_4d(base + c * p.slab_slot_bytes, &tmap_x, fb, (cb_lo + c) * 64, -p.pw, h0 - p.ph, frame);
          for (int j = 0; j < p.n_blocks; ++j)
            tma_load_4d(base + dy_off + j * (128 * 128), &tmap_dy, fb, nt * p.n_tile + j * 64, 0, h0, frame);
        } else {
          // tile -> (clip n, output frame t, 128-position block b); maps are {C, H*W, T, N}; frames outside [0, T) and
          // positions beyond H*W come back as zeros
          const int b = tile % p.blocks_per_frame;
          const int nt_ = tile / p.blocks_per_frame;
          const int t = nt_ % p.t_frames, n = nt_ / p.t_frames;
          for (int c = 0; c < ncb; ++c)
            tma_load_4d(base + c * p.slab_slot_bytes, &tmap_x, fb, (cb_lo + c) * 64, b * 128 + (tap_t - p.pt) * p.tap_pos,
                        t * p.t_stride + (tap_t - p.pt) * p.tap_frames, n);
          if (p.taps_on_n) {
            for (int j = 0; j < p.n_blocks; ++j)          // N block j = filter tap j: the output frame that input frame t feeds through it
              tma_load_4d(base + dy_off + j * (128 * 128), &tmap_dy, fb, 0, b * 128, t + p.pt - j, n);
          } else {
            for (int j = 0; j < p.n_blocks; ++j)
              tma_load_4d(base + dy_off + j * (128 * 128), &tmap_dy, fb, nt * p.n_tile + j * 64, b * 128, t, n);
          }
        }
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (warp-uniform control flow, elected lane issues)
    const uint32_t idesc = ptx::make_idesc_bf16(128, p.n_tile, 1, 1);
    // per M tile: byte offset of its first group inside a stage and the distance to its second group
    uint32_t a_off[kWgsMaxMt], a_lbo[kWgsMaxMt];
#pragma unroll
    for (int i = 0; i < kWgsMaxMt; ++i) {
      a_off[i] = 0; a_lbo[i] = 0;
      if (i < mt_count) {
        const int ga = g_lo + 2 * i;
        const int gb = ga + 1 < g_hi ? ga + 1 : ga;          // odd tail: second half duplicates the first, ignored later
        uint32_t oa, ob;
        if (p.temporal) {
          oa = (ga - cb_lo) * p.slab_slot_bytes;
          ob = (gb - cb_lo) * p.slab_slot_bytes;
        } else {
          const int cba = ga / p.taps, tapa = ga - cba * p.taps;
          const int cbb = gb / p.taps, tapb = gb - cbb * p.taps;
          oa = (cba - cb_lo) * p.slab_slot_bytes + ((tapa / p.kw) * p.wp + tapa % p.kw) * 128;
          ob = (cbb - cb_lo) * p.slab_slot_bytes + ((tapb / p.kw) * p.wp + tapb % p.kw) * 128;
        }
        a_off[i] = oa;
        a_lbo[i] = ob - oa;
      }
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile0; tile < tile1; ++tile) {
      ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
      ptx::tc_fence_after();
      const uint32_t base = ptx::smem_u32(smem + stage * p.stage_bytes);
      const uint64_t b_desc = ptx::make_sw128_desc(base + dy_off, 128 * 128, 1024);
      if (ptx::elect_one()) {
#pragma unroll
        for (int i = 0; i < kWgsMaxMt; ++i) {
          if (i < mt_count) {
            const uint64_t a_desc = ptx::make_sw128_desc(base + a_off[i], a_lbo[i], 1024);
            const uint32_t d_tmem = tmem_base + i * p.acc_stride;
            for (int ks = 0; ks < p.ksteps; ++ks)      // 16 positions = 2048 B along K -> +128 in the (addr >> 4) field
              ptx::umma_bf16_ss(d_tmem, a_desc + 128 * ks, b_desc + 128 * ks, idesc, (tile > tile0 || ks > 0) ? 1u : 0u);
          }
        }
        ptx::umma_commit(ptx::smem_u32(&empty_bar[stage]));
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    if (ptx::elect_one()) ptx::umma_commit(ptx::smem_u32(acc_bar));
    __syncwarp();
  } else if (tile1 > tile0) {
    // ===================================================== epilogue (warps 2-5 own TMEM lane quadrants warp % 4)
    const int q = warp & 3;
    ptx::mbar_wait(ptx::smem_u32(acc_bar), 0);
    ptx::tc_fence_after();
    const int r = q * 32 + lane;
    for (int i = 0; i < mt_count; ++i) {
      const int g = g_lo + 2 * i + (r >> 6);
      const int cb = p.temporal ? g : g / p.taps;
      const int tap = p.temporal ? tap_t : g - cb * p.taps;
      const int ci = cb * 64 + (r & 63);
      const bool row_ok = g < g_hi && ci < p.cin_real;
      const uint32_t taddr = tmem_base + i * p.acc_stride + (static_cast<uint32_t>(q * 32) << 16);
      for (int c = 0; c < p.n_tile; c += 16) {
        uint32_t v[16];
        ptx::tmem_ld_32x32b_x16(taddr + c, v);
        ptx::tmem_ld_wait();
        if (row_ok && !p.dbg_no_store) {
          float* dst = p.ws != nullptr ? p.ws + static_cast<size_t>(split) * p.ws_split_stride : p.dw;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = nt * p.n_tile + c + j;
            const int co = p.taps_on_n ? (col & 63) : col;
            const int tp = p.taps_on_n ? (col >> 6) : tap;
            if (co < p.cout_real) {
              const size_t idx = p.w_ohwi ? (static_cast<size_t>(co) * p.taps + tp) * p.cin_real + ci
                                          : (static_cast<size_t>(co) * p.cin_real + ci) * p.taps + tp;
              dst[idx] = __uint_as_float(v[j]);                               // a warp writes 32 consecutive floats (OHWI)
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

__global__ void __launch_bounds__(kWgsThreads, 1)
conv_wgrad_slab_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                       const WgradSlabParams p) {
  fvt_pdl_entry();
  extern __shared__ __align__(1024) uint8_t smem[];
  wgrad_slab_body(&tmap_x, &tmap_dy, p, blockIdx.x, smem, false);
}

// ---------------------------------------------------------------------------------------------------------------------
// Grouped launch: the weight gradients of SEVERAL layers in one grid.  The small-feature-map stages (conv4_x / conv5_x
// at a few clips per GPU: 784 - 6272 output positions per layer against 0.3 - 5 M weights) cannot fill 148 SMs per layer
// without cutting the pixel range into many splits, and every split re-reads both operands and adds a dW-sized slice to
// the reduction; launched one by one each layer also pays its own launch, pipeline fill, epilogue and reduce pass
// (measured 35 - 55 us per layer against 2 - 12 us of tensor work).  A weight gradient feeds nothing but the optimiser,
// so the training plan defers the layers of a stage and runs them together: the (M chunk, N tile) items of all layers
// fill the machine with few or no pixel splits.  One table entry per layer (its tensor maps + parameters), one
// (entry, item) pair per CTA, longest items first.
struct __align__(128) WgradGroupEntry {
  CUtensorMap tmx, tmdy;
  WgradSlabParams p;
};

__global__ void __launch_bounds__(kWgsThreads, 1)
conv_wgrad_group_kernel(const WgradGroupEntry* __restrict__ entries, const int2* __restrict__ cta_map) {
  fvt_pdl_entry();
  extern __shared__ __align__(1024) uint8_t smem[];
  const int2 m = cta_map[blockIdx.x];
  const WgradGroupEntry* e = entries + m.x;
  // the parameters move into the CTA's auxiliary area (behind the barriers): every role reads them many times
  const int stages = e->p.stages, stage_bytes = e->p.stage_bytes;
  WgradSlabParams* sp = reinterpret_cast<WgradSlabParams*>(smem + stages * stage_bytes + 256);
  static_assert(sizeof(WgradSlabParams) <= 768 && sizeof(WgradSlabParams) % 4 == 0, "parameter copy must fit the aux area");
  for (int i = threadIdx.x; i < static_cast<int>(sizeof(WgradSlabParams) / 4); i += kWgsThreads)
    reinterpret_cast<int*>(sp)[i] = reinterpret_cast<const int*>(&e->p)[i];
  __syncthreads();
  wgrad_slab_body(&e->tmx, &e->tmdy, *sp, m.y, smem, true);
}

// Slice reduction of every split layer of a group in one launch: block -> (entry, 4096-float chunk of its dW);
// dw[i] = sum over k < splits of ws[k][i] in split order (deterministic).
constexpr int kWgrChunk = 4096;
__global__ void __launch_bounds__(256)
wgrad_group_reduce_kernel(const WgradGroupEntry* __restrict__ entries, const int2* __restrict__ red_map) {
  fvt_pdl_entry();
  const int2 m = red_map[blockIdx.x];
  const WgradSlabParams& p = entries[m.x].p;
  const long long elems = p.ws_split_stride;
  const long long lo = static_cast<long long>(m.y) * kWgrChunk;
  long long hi = lo + kWgrChunk;
  if (hi > elems) hi = elems;
  const float* ws = p.ws;
  float* dw = p.dw;
  const int splits = p.splits;
  if ((elems & 3) == 0 && (reinterpret_cast<uintptr_t>(dw) & 15) == 0 && (reinterpret_cast<uintptr_t>(ws) & 15) == 0) {
    const long long n4 = elems >> 2;
    const float4* w4 = reinterpret_cast<const float4*>(ws);
    float4* d4 = reinterpret_cast<float4*>(dw);
    for (long long i = (lo >> 2) + threadIdx.x; i < (hi >> 2); i += 256) {
      float4 a = __ldg(w4 + i);
      for (int k = 1; k < splits; ++k) {
        const float4 b = __ldg(w4 + k * n4 + i);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      d4[i] = a;
    }
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += 256) {
      float a = __ldg(ws + i);
      for (int k = 1; k < splits; ++k) a += __ldg(ws + k * elems + i);
      dw[i] = a;
    }
  }
}

}  // namespace fvt
