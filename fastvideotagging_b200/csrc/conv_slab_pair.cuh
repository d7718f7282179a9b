// K1s2 — the slab convolution (conv_slab.cuh, stationary filter) on a CTA PAIR: tcgen05.mma.cta_group::2, M = 256.
//
// Why: the conv2_x 1x3x3 layers (64 -> 144, the largest single share of the network) keep their whole filter in shared
// memory (9 taps x 144 rows x 128 B = 166 KB), which leaves one SM no room for anything but two input stages: the
// epilogue has to store 32-byte row segments from registers (32 different 128-byte lines per warp instruction), and
// that request stream, not the tensor pipe, paces the kernel (main loop 555 us, with the epilogue 737 us at batch 48).
// With cta_group::2 each CTA of the pair holds HALF of the filter's N rows (83 KB) and its own 128-row input slab; one
// M = 256 instruction issued by the leader CTA drives both SMs' tensor cores, and each CTA's TMEM receives its 128 rows
// x all N columns.  The freed 83 KB pay for two [R_out x W x Cout] staging tiles, so the output leaves as ONE TMA store
// per tile (dense rows: the junk columns of the padded-row GEMM are simply never staged).
//
// Pair protocol (verified in isolation by tools/experiments/cta_pair_mma.cu):
//   * both CTAs allocate TMEM with tcgen05.alloc.cta_group::2; teardown after a cluster barrier;
//   * every CTA loads its own slab (local mbarrier) — the peer's relay warp forwards "my slab has landed" to the
//     leader's peer_full barrier with a remote mbarrier.arrive (mapa), the same for the filter halves;
//   * the leader's MMA warp waits for both, issues the M = 256 MMAs and commits with .multicast::cluster onto BOTH
//     CTAs' slab_empty / acc_full barriers (same shared-memory offsets);
//   * both CTAs' epilogue warps drain their own TMEM rows and arrive (locally / remotely) on the LEADER's acc_empty.
//
// Two generalisations make the pair the default schedule for the layers whose filter does NOT fit one SM but does fit
// two (conv3_x 1x3x3 128 -> 288: 663 KB streamed per 128-row tile by K1s, the smem fill rate paces it; conv2_x 1x3x3
// data gradient 144 -> 64: 221 KB):
//   * N tiles: a cluster keeps ONE N tile (n_tile columns) of the filter for its lifetime, cluster c works on N tile
//     c mod n_tiles and on every (clusters / n_tiles)-th tile pair; the input tile is read once per N tile (L2);
//   * column chunks: a tile may be r_out rows x w_tile columns of a wide image, which is how the stride-1 kt x 1 x 1
//     TEMPORAL convs whose filter fits two SMs (conv3_x 288 -> 128 and its data gradient) run here: image rows = frames,
//     image columns = the H*W positions, 16 frames x 8 positions per tile, zero frames by TMA out-of-bounds fill;
//   * the input ring is per 64-channel BLOCK, not per tile: the MMAs run channel-block-major, a block slot is released
//     as soon as its nine taps are issued, so two slots overlap loads and MMAs even when only one tile's slab fits.
// Both CTAs' slab loads count on the LEADER's barrier (cp.async.bulk.tensor .cta_group::2), so the MMA warp waits once
// per block and only the one-time filter load needs the relay.
//
// Warp roles per CTA (384 threads): warp0 slab producer, warp1 MMA issuer (leader) / relay (peer), warp2 TMEM
// allocator, warp3 filter producer, warps 4-11 epilogue.
// Replaces cuDNN convolution calls for Conv3D(k=(1,3,3)) at reference model/R2Plus1.py:27-31,100-104, net.py:40-42.
#pragma once
#include "ptx.cuh"
#include "epilogue.cuh"
#include "det_sum.cuh"
#include "conv_slab.cuh"
#include "pdl.cuh"

namespace fvt {

constexpr int kPairThreads = 384;
constexpr int kPairMaxStages = 6;          // input ring slots (one 64-channel block each)

struct SlabPairParams {
  SlabParams s;               // geometry as in conv_slab.cuh (n_tile = full N, b_ring = taps * cin_blocks)
  int n_half;                 // filter rows per CTA (= n_tile / 2, multiple of 8)
  int num_pairs;              // ceil(num_m_tiles / 2): tile 2*pair + rank belongs to CTA `rank` of the cluster
  int out_tile_bytes;         // r_out * w * cout_store * 2 rounded up to 128
  int tma_store;              // 1: staged TMA store (needs n_tile == cout_store), 0: register stores
  int n_tiles;                // N tiles of n_tile columns; a cluster works on N tile (cluster id mod n_tiles) only
  int w_chunks, w_tile;       // column chunks per row tile and their width (1, s.w for whole-row tiles).  With chunks
                              // (kw = 1, pw = 0, s.wp = w_tile) a tile is r_out rows x w_tile columns: a kt x 1 x 1 temporal
                              // conv runs here as a (kh = kt, kw = 1) conv over the image [H' = T] x [W' = H*W]
};

namespace pair {
__device__ __forceinline__ uint32_t ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t nclusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// Arrive on a barrier of another CTA of the cluster.  Default semantics (.release.cta), as CUTLASS's ClusterBarrier does:
// what crosses CTAs here is tensor memory (ordered by tcgen05.wait / tcgen05.fence) and TMA-written shared memory
// (ordered by the barrier's transaction count), never generic-proxy data.  `.release.cluster` compiles to
// MEMBAR.ALL.GPU + ERRBAR, which waits for every outstanding global store of the warp: ncu attributed 8 % of all stall
// samples of the fused unit to it and it sat on the S-release / accumulator-release critical paths.
__device__ __forceinline__ void remote_arrive(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a LOCAL barrier that remote CTAs (or the tensor core / TMA of the pair) arrive on: the plain CTA-scope wait
// (an .acquire.cluster wait adds an L1 invalidate per success; see remote_arrive for why CTA scope is enough)
__device__ __forceinline__ void wait_cluster(uint32_t bar, uint32_t parity) { ptx::mbar_wait(bar, parity); }
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// all prior MMAs of this thread -> arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(static_cast<uint16_t>(3)) : "memory");
}
// 4-D tiled store without a swizzle (dense staging tile)
__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3) {
  ptx::tma_store_4d(tmap, src, c0, c1, c2, c3);
}
// 4-D tiled load whose completion bytes are counted on a barrier of EITHER CTA of the pair (.cta_group::2)
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const void* tmap, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
}  // namespace pair

__global__ void __launch_bounds__(kPairThreads, 1)
conv_slab_pair_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                      const __grid_constant__ CUtensorMap tmap_y, const SlabPairParams pp) {
  fvt_pdl_entry();
  extern __shared__ __align__(1024) uint8_t smem[];
  const SlabParams& p = pp.s;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
  const uint32_t rank = pair::ctarank();
  const bool leader = rank == 0;

  const int taps = p.kh * p.kw;
  const int b_slab_bytes = pp.n_half * 128;
  const int b_all = taps * p.cin_blocks;
  const int stage_bytes = p.slab_slot_bytes;                           // one ring slot = one 64-channel block of one tile
  uint8_t* smem_b = smem;                                              // [b_all][n_half x 64]
  uint8_t* smem_a = smem + ((b_all * b_slab_bytes + 1023) & ~1023);    // [stages][slot]
  uint8_t* smem_o = smem_a + p.stages * stage_bytes;                   // [2][out_tile_bytes] (tma_store)
  uint8_t* aux = smem_o + (pp.tma_store ? 2 * pp.out_tile_bytes : 0);
  uint64_t* slab_full = reinterpret_cast<uint64_t*>(aux);             // [kPairMaxStages] leader: both CTAs' blocks have landed
  uint64_t* peer_full = slab_full + kPairMaxStages;                    // [kPairMaxStages] (unused)
  uint64_t* slab_empty = peer_full + kPairMaxStages;                   // [kPairMaxStages] multicast commit
  uint64_t* b_full = slab_empty + kPairMaxStages;                      // [1] local filter half landed
  uint64_t* peer_b_full = b_full + 1;                                  // [1] leader: the peer's filter half landed
  uint64_t* acc_full = peer_b_full + 1;                                // [2] multicast commit
  uint64_t* acc_empty = acc_full + 2;                                  // [2] leader: 16 epilogue warps (8 local + 8 remote)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* affine_smem = reinterpret_cast<float*>(tmem_slot + 4);        // scale[n_tile], shift[n_tile] / statistics
  const int n_total = p.n_tile * pp.n_tiles;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_x);
    ptx::prefetch_tensormap(&tmap_w);
    if (pp.tma_store) ptx::prefetch_tensormap(&tmap_y);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&slab_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&peer_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&slab_empty[s]), 1);
    }
    ptx::mbar_init(ptx::smem_u32(b_full), 1);
    ptx::mbar_init(ptx::smem_u32(peer_b_full), 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&acc_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&acc_empty[s]), 16);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) pair::tmem_alloc2(ptx::smem_u32(tmem_slot), 512);
  if (p.scale != nullptr) {
    for (int i = threadIdx.x; i < n_total; i += kPairThreads) {
      affine_smem[i] = i < p.cout_store ? __ldg(p.scale + i) : 0.f;
      affine_smem[n_total + i] = i < p.cout_store ? __ldg(p.shift + i) : 0.f;
    }
  }
  // statistics accumulate in the affine area, [4 quadrants][2][n_total] (sized by the host when FVT_CONV_STATS is set)
  const bool bnbwd = (p.flags & kConvBnBwd) != 0;               // staged scale/shift AND partials: the partials follow the affine area
  const bool acc_stats = (p.flags & kConvStats) != 0 && (p.scale == nullptr || bnbwd);
  float* stat_base = bnbwd ? affine_smem + 2 * n_total : affine_smem;
  if (acc_stats)
    for (int i = threadIdx.x; i < 8 * n_total; i += kPairThreads) stat_base[i] = 0.f;
  ptx::tc_fence_before();
  __syncthreads();
  pair::cluster_sync_all();                    // barriers of both CTAs are initialised before anything arrives remotely
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m_tiles = p.frames * p.tiles_per_frame;
  const int nt = static_cast<int>(pair::cluster_id_x()) % pp.n_tiles;            // this cluster's N tile
  const int n0 = nt * p.n_tile;
  const int pair0 = static_cast<int>(pair::cluster_id_x()) / pp.n_tiles;
  const int pair_step = static_cast<int>(pair::nclusters_x()) / pp.n_tiles;      // the host launches a multiple of n_tiles clusters

  if (warp == 0) {
    // ===================================================== input slab producer (own tile, block by block; out-of-range tiles zero-fill)
    int stage = 0;
    uint32_t phase = 0;
    for (int pr = pair0; pr < pp.num_pairs; pr += pair_step) {
      const int mt = 2 * pr + static_cast<int>(rank);
      const int frame = mt / p.tiles_per_frame;
      const int rem = mt - frame * p.tiles_per_frame;
      const int rt = rem / pp.w_chunks;
      const int h0 = rt * p.r_out, w0 = (rem - rt * pp.w_chunks) * pp.w_tile;
      for (int cb = 0; cb < p.cin_blocks; ++cb) {
        ptx::mbar_wait(ptx::smem_u32(&slab_empty[stage]), phase ^ 1);
        const uint32_t fb = pair::map_to_rank(ptx::smem_u32(&slab_full[stage]), 0);       // the LEADER's barrier counts both blocks
        if (ptx::elect_one()) {
          if (leader) ptx::mbar_arrive_expect_tx(ptx::smem_u32(&slab_full[stage]), 2 * p.slab_tx_bytes);
          pair::tma_load_4d_2sm(ptx::smem_u32(smem_a + stage * stage_bytes), &tmap_x, fb, cb * 64, w0 - p.pw, h0 - p.ph, frame);
          if (p.prefetch_dist > 0) {
            const int mt2 = mt + 2 * p.prefetch_dist * pair_step;
            if (mt2 < num_m_tiles) {
              const int frame2 = mt2 / p.tiles_per_frame;
              const int rem2 = mt2 - frame2 * p.tiles_per_frame;
              const int rt2 = rem2 / pp.w_chunks;
              tma_prefetch_4d(&tmap_x, cb * 64, (rem2 - rt2 * pp.w_chunks) * pp.w_tile - p.pw, rt2 * p.r_out - p.ph, frame2);
            }
          }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ===================================================== filter producer: this CTA's half of the N rows, once
    if (pair0 < pp.num_pairs) {
      const uint32_t bb = ptx::smem_u32(b_full);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(bb, b_all * b_slab_bytes);
        int j = 0;
        for (int tap = 0; tap < taps; ++tap)
          for (int cb = 0; cb < p.cin_blocks; ++cb, ++j)
            ptx::tma_load_2d(ptx::smem_u32(smem_b + j * b_slab_bytes), &tmap_w, bb, tap * p.k_per_tap + cb * 64,
                             n0 + static_cast<int>(rank) * pp.n_half);
      }
      __syncwarp();
    }
  } else if (warp == 1 && !leader) {
    // ===================================================== relay (peer CTA): forward the filter's TMA completion to the leader
    if (pair0 < pp.num_pairs) {
      ptx::mbar_wait(ptx::smem_u32(b_full), 0);
      if (ptx::elect_one()) pair::remote_arrive(pair::map_to_rank(ptx::smem_u32(peer_b_full), 0));
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA): M = 256 over the pair
    const uint32_t idesc = ptx::make_idesc_bf16(256, p.n_tile, 0, 0);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    bool first = true;
    const uint64_t b_desc0 = ptx::make_sw128_desc(ptx::smem_u32(smem_b), 16, 1024);
    const uint32_t b_step = static_cast<uint32_t>(b_slab_bytes) >> 4;
    const uint32_t b_tap_step = b_step * static_cast<uint32_t>(p.cin_blocks);
    const uint32_t a_row_step = static_cast<uint32_t>(p.wp) * 8u;
    for (int pr = pair0; pr < pp.num_pairs; pr += pair_step) {
      if (first) {
        ptx::mbar_wait(ptx::smem_u32(b_full), 0);
        pair::wait_cluster(ptx::smem_u32(peer_b_full), 0);
        first = false;
      }
      pair::wait_cluster(ptx::smem_u32(&acc_empty[acc]), acc_phase ^ 1);
      const uint32_t d_tmem = tmem_base + acc * 256;
      uint32_t acc_flag = 0;
      int k16 = p.cin_k16;
      for (int cb = 0; cb < p.cin_blocks; ++cb, k16 -= 4) {
        // ---- channel-block-major: the nine taps of one 64-channel block, then the block's ring slot is free again
        pair::wait_cluster(ptx::smem_u32(&slab_full[stage]), phase);
        ptx::tc_fence_after();
        const uint64_t a_desc0 = ptx::make_sw128_desc(ptx::smem_u32(smem_a + stage * stage_bytes), 16, 1024);
        if (ptx::elect_one()) {
          uint64_t b_desc = b_desc0 + static_cast<uint32_t>(cb) * b_step;
          uint64_t a_row = a_desc0;
          for (int dh = 0; dh < p.kh; ++dh, a_row += a_row_step) {
            uint64_t a_tap = a_row;
            for (int dw = 0; dw < p.kw; ++dw, a_tap += 8, b_desc += b_tap_step) {
              pair::umma2_bf16_ss(d_tmem, a_tap, b_desc, idesc, acc_flag);
              if (k16 > 1) pair::umma2_bf16_ss(d_tmem, a_tap + 2, b_desc + 2, idesc, 1);
              if (k16 > 2) pair::umma2_bf16_ss(d_tmem, a_tap + 4, b_desc + 4, idesc, 1);
              if (k16 > 3) pair::umma2_bf16_ss(d_tmem, a_tap + 6, b_desc + 6, idesc, 1);
              acc_flag = 1;
            }
          }
          if (cb == p.cin_blocks - 1) pair::umma2_commit_both(ptx::smem_u32(&acc_full[acc]));
          pair::umma2_commit_both(ptx::smem_u32(&slab_empty[stage]));
        }
        __syncwarp();
        acc_flag = 1;
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue (own 128 accumulator rows)
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int et = threadIdx.x - 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    EpilogueArgs ea;
    ea.block_n = p.n_tile; ea.cout_store = p.cout_store; ea.flags = acc_stats ? p.flags : (p.flags & ~kConvStats);
    ea.scale_smem = p.scale != nullptr ? affine_smem : nullptr; ea.shift_smem = affine_smem + n_total;
    ea.residual = p.residual; ea.y = p.y; ea.stat_smem = stat_base + q * 2 * n_total + n0; ea.stat_stride = n_total;
    const int r = q * 32 + lane;                 // GEMM row = padded position inside the tile
    const int hl = r / p.wp, wl = r - hl * p.wp;
    int shl[4], swl[4];
    uint32_t smask_static = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rj = q * 32 + (lane >> 2) + 8 * j;
      shl[j] = rj / p.wp;
      swl[j] = rj - shl[j] * p.wp;
      if (shl[j] < p.r_out && swl[j] < pp.w_tile) smask_static |= 1u << j;
    }
    const uint32_t acc_empty_leader0 = pair::map_to_rank(ptx::smem_u32(&acc_empty[0]), 0);
    int obuf = 0;
    for (int pr = pair0; pr < pp.num_pairs; pr += pair_step) {
      const int mt = 2 * pr + static_cast<int>(rank);
      const int frame = mt / p.tiles_per_frame;
      const int rem = mt - frame * p.tiles_per_frame;
      const int rt = rem / pp.w_chunks;
      const int h0 = rt * p.r_out, w0 = (rem - rt * pp.w_chunks) * pp.w_tile;
      const bool in_range = mt < num_m_tiles;
      const bool ok = in_range && hl < p.r_out && wl < pp.w_tile && (h0 + hl) < p.h && (w0 + wl) < p.w;
      const long long out_row = ok ? (static_cast<long long>(frame) * p.h + h0 + hl) * p.w + w0 + wl : -1ll;
      {
        uint32_t m = in_range ? smask_static : 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (h0 + shl[j] >= p.h || w0 + swl[j] >= p.w) m &= ~(1u << j);
        ea.stat_mask = m;
      }
      if (pp.tma_store) {
        if (et == 0) ptx::tma_store_wait_read<1>();      // the store that read this staging buffer two tiles ago is done
        asm volatile("bar.sync 1, 256;" ::: "memory");
        ea.stage_smem = ptx::smem_u32(smem_o + obuf * pp.out_tile_bytes);
        ea.stage_row = (hl < p.r_out && wl < p.w) ? hl * p.w + wl : -1;       // staged store: whole-row tiles only (w_chunks == 1)
        ea.stage_pitch = p.cout_store * 2;
        ea.stage_swizzle = 0;
      }
      epilogue_prefetch_residual(ea, n0, out_row, grp);
      ptx::mbar_wait(ptx::smem_u32(&acc_full[acc]), acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
      epilogue_chunks(ea, taddr, n0, out_row, grp, lane);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) pair::remote_arrive(acc_empty_leader0 + acc * 8);      // the leader owns the accumulator hand-back
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (pp.tma_store) {
        ptx::fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0 && in_range) {
          pair::tma_store_4d(&tmap_y, ea.stage_smem, 0, 0, h0, frame);      // rows beyond H are clipped by the map
          ptx::tma_store_commit();
        }
        obuf ^= 1;
      }
    }
    if (pp.tma_store && et == 0) ptx::tma_store_wait<0>();
    if (acc_stats && pair0 < pp.num_pairs) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      flush_quadrant_stats(stat_base, n_total, p.cout_store, p.stats, et, 256, n0,
                           n0 + p.n_tile < p.cout_store ? n0 + p.n_tile : p.cout_store);
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
