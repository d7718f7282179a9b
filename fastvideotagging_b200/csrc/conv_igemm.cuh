// K1 — implicit-GEMM 3-D convolution forward for sm_100a.
//
//   Y[m, co] = epilogue( sum_{tap, ci} X[pixel(m) + tap, ci] * Wp[co, tap, ci] )
//
// * activations NDHWC bf16; one GEMM row m = one output pixel in (n, t, h, w) order, so a 128-row tile is
//   128 consecutive output pixels and the output tile is a dense [128 x Cout] block of Y.
// * A operand: TMA *im2col* loads (cp.async.bulk.tensor.5d...im2col): one load per (filter tap, 64-channel
//   block) brings the 128 pixels x 64 channels slab, zero-filled at the padding halo, straight into
//   128B-swizzled shared memory in the canonical K-major UMMA layout.
// * B operand: packed weights [Cout_pad, taps*Cin] (K-major), 2-D tiled TMA.
// * MMA: tcgen05.mma cta_group::1 kind::f16, M=128, N=block_n (16..256), K=16 per instruction, fp32
//   accumulators in TMEM (two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1).
// * Warp roles (384 threads): warp0 TMA producer, warp1 MMA issuer, warp2 TMEM allocator, warps4-11 epilogue
//   (two warps per TMEM lane quadrant, alternate 16-column chunks, TMEM/residual loads software-pipelined).
// * Epilogue: tcgen05.ld -> per-channel scale/shift (folded BatchNorm) -> (+ residual) -> ReLU -> bf16 store;
//   optionally per-channel sum / sum-of-squares of the raw conv output for training-mode BatchNorm.
//
// Replaces the cuDNN convolution + BatchNorm + Activation + elemwise_add launches that MXNet issues for
// reference model/R2Plus1.py:27-38,59-62,67-71,81 and net.py:40-51,79-101.
#pragma once
#include "ptx.cuh"
#include "epilogue.cuh"
#include "det_sum.cuh"
#include "pdl.cuh"

namespace fvt {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;            // bf16 elements per 128-byte swizzle row
constexpr int kConvThreads = 384;          // warps 0-3: TMA / MMA / TMEM alloc / spare; warps 4-11: epilogue
constexpr int kEpilogueThreads = 256;
constexpr int kMaxCout = 1536;             // per-channel scale/shift staged in shared memory
constexpr int kMaxStages = 8;
constexpr int kATileBytes = kBlockM * kBlockK * 2;   // 16 KiB

struct ConvKernelParams {
  int m_total;            // N*To*Ho*Wo
  int to, ho, wo;
  int st, sh, sw;
  int pt, ph, pw;
  int kt, kh, kw;
  int cin_k16;            // 16-channel MMA steps per filter tap (= stored Cin / 16)
  int cin_blocks;         // 64-channel TMA blocks per filter tap
  int k_per_tap;          // K elements per tap in the packed weights (= stored Cin)
  int block_n;            // N tile
  int num_m_tiles, num_n_tiles;
  int cout_store;         // channel pitch of Y / residual (elements)
  int flags;
  int stages;
  int k_splits;           // > 1: split-K — item = (tile, split); partial tiles are stored into ws[split] (fp32), no epilogue math
  int kb_per_split;
  float* ws;              // [k_splits][m_total][cout_store] fp32 (split-K only)
  int b_stationary;       // 1: all taps*cin_blocks weight tiles are loaded once and stay in shared memory
  const float* scale;     // [cout_store] or nullptr (identity)
  const float* shift;     // [cout_store] or nullptr
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
  unsigned long long* stats;   // [2][cout_store] exact accumulators (det_sum.cuh): sum, sum of squares; or nullptr
  // Output lattice map (fvt_conv3d_fwd_ex): GEMM row (n, ot, oh, ow) is stored at (n, ot*om_st + om_t0, oh*om_sh + om_h0,
  // ow*om_sw + om_w0) of a [N, om_t, om_h, om_w, cout_store] tensor (residual read at the same place).  om_on = 0: dense.
  int om_on;
  int om_t, om_h, om_w;
  int om_st, om_sh, om_sw;
  int om_t0, om_h0, om_w0;
};

// Row of Y (in pixels) a GEMM row is stored at; < 0 for rows beyond the problem.
__device__ __forceinline__ long long conv_out_row(const ConvKernelParams& p, int row) {
  if (row >= p.m_total) return -1ll;
  if (!p.om_on) return static_cast<long long>(row);
  int m = row;
  const int ow = m % p.wo;  m /= p.wo;
  const int oh = m % p.ho;  m /= p.ho;
  const int ot = m % p.to;
  const int on = m / p.to;
  return ((static_cast<long long>(on) * p.om_t + (ot * p.om_st + p.om_t0)) * p.om_h + (oh * p.om_sh + p.om_h0)) * p.om_w +
         (ow * p.om_sw + p.om_w0);
}

__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_fwd_kernel(const __grid_constant__ CUtensorMap tmap_x,
                      const __grid_constant__ CUtensorMap tmap_w,
                      const ConvKernelParams p) {
  fvt_pdl_entry();
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.stages;
  const int b_tile_bytes = p.block_n * kBlockK * 2;
  const int taps = p.kt * p.kh * p.kw;
  const int k_blocks = taps * p.cin_blocks;
  // stationary weights: [k_blocks][block_n x 64] ahead of the A ring, whose stages then hold the A tile only
  const int b_region_bytes = p.b_stationary ? k_blocks * b_tile_bytes : 0;
  const int stage_bytes = p.b_stationary ? kATileBytes : kATileBytes + b_tile_bytes;

  uint8_t* smem_b = smem;
  uint8_t* smem_tiles = smem + b_region_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_tiles + stages * stage_bytes);
  uint64_t* full_bar = bars;                       // [stages]
  uint64_t* empty_bar = bars + kMaxStages;         // [stages]
  uint64_t* acc_full_bar = bars + 2 * kMaxStages;  // [2]
  uint64_t* acc_empty_bar = acc_full_bar + 2;      // [2]
  uint64_t* b_full_bar = acc_empty_bar + 2;        // [1] stationary weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full_bar + 2);   // keeps the float4 arrays below 16-byte aligned
  // [2][kMaxCout] scale, shift — or, for the training forward (statistics, no folded affine), [4 quadrants][2][n_pad]
  // per-CTA channel partials (the host sizes the region: conv_aux_bytes)
  float* affine_smem = reinterpret_cast<float*>(tmem_slot + 4);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_x);
    ptx::prefetch_tensormap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), p.b_stationary ? 1 : 2);   // streamed weights: A and B producers arrive
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&acc_full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&acc_empty_bar[s]), 8);   // one arrive per epilogue warp
    }
    ptx::mbar_init(ptx::smem_u32(b_full_bar), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512);
    ptx::tmem_relinquish();
  }
  if (p.scale != nullptr) {
    const int padded = p.num_n_tiles * p.block_n;
    for (int i = threadIdx.x; i < padded; i += kConvThreads) {
      affine_smem[i] = i < p.cout_store ? __ldg(p.scale + i) : 0.f;
      affine_smem[kMaxCout + i] = i < p.cout_store ? __ldg(p.shift + i) : 0.f;
    }
  }
  // Training forward (statistics, no folded affine): the per-channel sums of ALL tiles of this CTA accumulate in the
  // otherwise unused affine area, one [2][n_pad] block per TMEM lane quadrant, and reach global memory once, after the
  // tile loop (no per-tile barriers, no floating-point atomics: see det_sum.cuh).
  // (kConvBnBwd — data gradient fused with the consumer BatchNorm's backward sums — needs BOTH the staged scale/shift and
  // the partials: they then follow the affine area)
  const bool bnbwd = (p.flags & kConvBnBwd) != 0;
  const bool acc_stats = (p.flags & kConvStats) != 0 && (p.scale == nullptr || bnbwd) && p.k_splits == 1;
  const int n_pad = p.num_n_tiles * p.block_n;
  float* stat_base = bnbwd ? affine_smem + 2 * kMaxCout : affine_smem;
  if (acc_stats)
    for (int i = threadIdx.x; i < 8 * n_pad; i += kConvThreads) stat_base[i] = 0.f;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.num_m_tiles * p.num_n_tiles * p.k_splits;     // work items (tile x K split)

  if (warp == 0) {
    // ===================================================== TMA producer (warp-uniform loop, elected lane issues)
    {
      if (p.b_stationary && blockIdx.x < num_tiles) {
        const uint32_t bb = ptx::smem_u32(b_full_bar);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(bb, b_region_bytes);
          for (int kb = 0; kb < k_blocks; ++kb) {
            const int tap = kb / p.cin_blocks, cb = kb - tap * p.cin_blocks;
            ptx::tma_load_2d(ptx::smem_u32(smem_b + kb * b_tile_bytes), &tmap_w, bb, tap * p.k_per_tap + cb * kBlockK, 0);
          }
        }
        __syncwarp();
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_tiles; item += gridDim.x) {
        const int tile = item / p.k_splits;
        const int split = item - tile * p.k_splits;
        const int m_blk = tile / p.num_n_tiles;
        const int n_blk = tile - m_blk * p.num_n_tiles;
        int m0 = m_blk * kBlockM;
        const int ow = m0 % p.wo;  m0 /= p.wo;
        const int oh = m0 % p.ho;  m0 /= p.ho;
        const int ot = m0 % p.to;
        const int on = m0 / p.to;
        const int cw = ow * p.sw - p.pw;
        const int ch = oh * p.sh - p.ph;
        const int cd = ot * p.st - p.pt;
        // this item's k-block range [kb0, kb1) and the filter position of kb0
        const int kb0 = split * p.kb_per_split;
        const int kb1 = kb0 + p.kb_per_split < k_blocks ? kb0 + p.kb_per_split : k_blocks;
        int cb = kb0 % p.cin_blocks;
        int tap = kb0 / p.cin_blocks;
        int dw = tap % p.kw, dh = (tap / p.kw) % p.kh, dt = tap / (p.kw * p.kh);
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = ptx::smem_u32(&full_bar[stage]);
          uint8_t* a_dst = smem_tiles + stage * stage_bytes;
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(fb, kATileBytes);
            ptx::tma_load_im2col_5d(ptx::smem_u32(a_dst), &tmap_x, fb, cb * kBlockK, cw, ch, cd, on,
                                    static_cast<uint16_t>(dw), static_cast<uint16_t>(dh), static_cast<uint16_t>(dt));
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
          if (++cb == p.cin_blocks) {
            cb = 0; ++tap;
            if (++dw == p.kw) { dw = 0; if (++dh == p.kh) { dh = 0; ++dt; } }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================================================== weight-tile producer (streamed weights only)
    // A single issuing thread sustains one TMA per ~275 clk (profiles/r01_tma_rate_microbench.log), so with the input
    // and the weight tile on one thread the mid-size layers (N <= 256: MMA work per k-block < 550 clk) were paced by the
    // producer (conv3_x 288->128 temporal: 7900 clk per tile against 3840 clk of MMAs).  The weight tiles therefore come
    // from their own warp; both producers arrive on the stage's full barrier.
    if (!p.b_stationary) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full_u32 = ptx::smem_u32(full_bar), empty_u32 = ptx::smem_u32(empty_bar);
      const uint32_t b_dst0 = ptx::smem_u32(smem_tiles) + kATileBytes;
      for (int item = blockIdx.x; item < num_tiles; item += gridDim.x) {
        const int tile = item / p.k_splits;
        const int split = item - tile * p.k_splits;
        const int n0 = (tile % p.num_n_tiles) * p.block_n;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = kb0 + p.kb_per_split < k_blocks ? kb0 + p.kb_per_split : k_blocks;
        int cb = kb0 % p.cin_blocks;
        int tap = kb0 / p.cin_blocks;
        int kcoord = tap * p.k_per_tap + cb * kBlockK;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(empty_u32 + stage * 8, phase ^ 1);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(full_u32 + stage * 8, b_tile_bytes);
            ptx::tma_load_2d(b_dst0 + stage * stage_bytes, &tmap_w, full_u32 + stage * 8, kcoord, n0);
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
          kcoord += kBlockK;
          if (++cb == p.cin_blocks) { cb = 0; ++tap; kcoord = tap * p.k_per_tap; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    // The whole warp runs the loop with warp-uniform control flow (barrier waits, descriptor arithmetic in uniform
    // registers); one elected lane issues the tcgen05 instructions.  A thread-divergent issuer (if lane == 0 around the
    // loop) makes the compiler wrap every UTCHMMA in an ELECT/BRA.U.ANY serialisation loop and costs ~230 clk per MMA.
    {
      // The issue loop runs on ONE thread: every instruction of bookkeeping per k-block is ~6 clk of serial latency, so
      // descriptors advance by precomputed steps and barrier addresses are plain adds.
      const uint32_t idesc = ptx::make_idesc_bf16(kBlockM, p.block_n, 0, 0);
      const uint32_t tiles_u32 = ptx::smem_u32(smem_tiles);
      const uint32_t full_u32 = ptx::smem_u32(full_bar), empty_u32 = ptx::smem_u32(empty_bar);
      const uint64_t a_desc_s0 = ptx::make_sw128_desc(tiles_u32, 16, 1024);
      const uint32_t stage_step = static_cast<uint32_t>(stage_bytes) >> 4;
      const uint64_t b_stat0 = ptx::make_sw128_desc(ptx::smem_u32(smem_b), 16, 1024);
      const uint32_t b_step = static_cast<uint32_t>(b_tile_bytes) >> 4;
      const uint32_t b_in_stage = p.b_stationary ? 0u : static_cast<uint32_t>(kATileBytes >> 4);
      const int last_k16 = p.cin_k16 - (p.cin_blocks - 1) * (kBlockK / 16);   // MMAs of the last channel block of a tap
      int stage = 0;
      uint32_t phase = 0;
      uint64_t a_desc = a_desc_s0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (p.b_stationary && blockIdx.x < num_tiles) ptx::mbar_wait(ptx::smem_u32(b_full_bar), 0);
      for (int item = blockIdx.x; item < num_tiles; item += gridDim.x) {
        ptx::mbar_wait(ptx::smem_u32(&acc_empty_bar[acc]), acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        const int split = item % p.k_splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = kb0 + p.kb_per_split < k_blocks ? kb0 + p.kb_per_split : k_blocks;
        int cb = kb0 % p.cin_blocks;
        uint64_t b_stat = b_stat0 + static_cast<uint32_t>(kb0) * b_step;
        for (int kb = kb0; kb < kb1; ++kb, b_stat += b_step) {
          const int k16 = (cb == p.cin_blocks - 1) ? last_k16 : kBlockK / 16;
          ptx::mbar_wait(full_u32 + stage * 8, phase);
          ptx::tc_fence_after();
          const uint64_t b_desc = p.b_stationary ? b_stat : a_desc + b_in_stage;
          if (ptx::elect_one()) {
            // +32 bytes (16 bf16) along K inside the swizzle atom == +2 in the (addr >> 4) field
            ptx::umma_bf16_ss(d_tmem, a_desc, b_desc, idesc, kb != kb0);
            if (k16 > 1) ptx::umma_bf16_ss(d_tmem, a_desc + 2, b_desc + 2, idesc, 1);
            if (k16 > 2) ptx::umma_bf16_ss(d_tmem, a_desc + 4, b_desc + 4, idesc, 1);
            if (k16 > 3) ptx::umma_bf16_ss(d_tmem, a_desc + 6, b_desc + 6, idesc, 1);
            ptx::umma_commit(empty_u32 + stage * 8);
          }
          __syncwarp();
          if (++cb == p.cin_blocks) cb = 0;
          a_desc += stage_step;
          if (++stage == stages) { stage = 0; phase ^= 1; a_desc = a_desc_s0; }
        }
        if (ptx::elect_one()) ptx::umma_commit(ptx::smem_u32(&acc_full_bar[acc]));
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue (8 warps: 2 per TMEM lane quadrant)
    // warp w reads TMEM lanes 32*(w%4).. ; the two warps of a quadrant take alternate 16-column chunks.
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;            // 0 or 1
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool has_affine = p.scale != nullptr;
    const int et = threadIdx.x - 128;           // 0..255
    for (int item = blockIdx.x; item < num_tiles; item += gridDim.x) {
      const int tile = item / p.k_splits;
      const int m_blk = tile / p.num_n_tiles;
      const int n_blk = tile - m_blk * p.num_n_tiles;
      const int n0 = n_blk * p.block_n;
      const int row = m_blk * kBlockM + q * 32 + lane;
      const bool row_ok = row < p.m_total;
      if (p.k_splits > 1) {
        // split-K: store this item's fp32 partial tile into its split's slice of the workspace (16-byte stores); the
        // epilogue math runs in splitk_finalize_kernel, which adds the slices in split order
        ptx::mbar_wait(ptx::smem_u32(&acc_full_bar[acc]), acc_phase);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
        const int split = item - tile * p.k_splits;
        float* wrow = p.ws + (static_cast<size_t>(split) * p.m_total + (row_ok ? row : 0)) * p.cout_store + n0;
        for (int c = grp * 16; c < p.block_n; c += 32) {
          uint32_t v[16];
          ptx::tmem_ld_32x32b_x16(taddr + c, v);
          ptx::tmem_ld_wait();
          if (row_ok && n0 + c < p.cout_store) {
            float4* dst = reinterpret_cast<float4*>(wrow + c);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                   __uint_as_float(v[4 * i + 3]));
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&acc_empty_bar[acc]));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      const long long out_row = conv_out_row(p, row);
      {
        EpilogueArgs pre;
        pre.block_n = p.block_n; pre.cout_store = p.cout_store; pre.flags = p.flags; pre.residual = p.residual;
        epilogue_prefetch_residual(pre, n0, out_row, grp);
      }
      ptx::mbar_wait(ptx::smem_u32(&acc_full_bar[acc]), acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
      {
        EpilogueArgs ea;
        ea.block_n = p.block_n; ea.cout_store = p.cout_store; ea.flags = acc_stats ? p.flags : (p.flags & ~kConvStats);
        ea.scale_smem = has_affine ? affine_smem : nullptr; ea.shift_smem = affine_smem + kMaxCout;
        ea.residual = p.residual; ea.y = p.y;
        ea.stat_smem = stat_base + q * 2 * n_pad + n0; ea.stat_stride = n_pad;
        ea.stat_mask = stat_mask_below(static_cast<long long>(m_blk) * kBlockM + q * 32, lane, p.m_total);
        epilogue_chunks(ea, taddr, n0, out_row, grp, lane);
      }
      // release the accumulator stage (all of this warp's TMEM reads have completed: wait::ld above)
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&acc_empty_bar[acc]));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (acc_stats && blockIdx.x < num_tiles) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      flush_quadrant_stats(stat_base, n_pad, p.cout_store, p.stats, et, kEpilogueThreads);
    }
  }

  // teardown
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace fvt
