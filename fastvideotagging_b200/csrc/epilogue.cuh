// Shared epilogue of the tcgen05 convolution kernels: TMEM accumulator tile -> registers -> (per-channel statistics)
// -> scale/shift (folded BatchNorm) -> (+ residual) -> (ReLU) -> bf16 rows of Y.
// One call handles this warp's 32 accumulator rows (TMEM lanes) for the 16-column chunks grp, grp+2, grp+4, ...;
// TMEM and residual loads of chunk i+1 are in flight while chunk i is processed.
#pragma once
#include "ptx.cuh"

namespace fvt {

enum ConvFlags : int {
  kConvRelu = 1,
  kConvResidual = 2,
  kConvStats = 4,
  kConvBnBwd = 16,       // data-gradient epilogue fused with the consumer BatchNorm's backward reduction (see epilogue_chunks_bnbwd)
  kDbgNoStore = 256,     // experiments only (fvt_set_option("debug_flags")): skip the global stores
  kDbgNoEpilogue = 512,  // experiments only: epilogue does the barrier handshakes but touches no data
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

struct EpilogueArgs {
  int block_n;                      // accumulator columns of this tile
  int cout_store;                   // channel pitch of Y / residual
  int flags;                        // ConvFlags
  const float* scale_smem;          // staged per-channel scale (indexed by absolute channel) or nullptr
  const float* shift_smem;
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
  // kConvStats: per-CTA partial sums of THIS warp's TMEM lane quadrant, [2][stat_stride] floats (sum, sum of squares).
  // Every (quadrant, channel) slot is written by exactly one warp (a quadrant's 16-column chunks are dealt to its warps
  // statically), tiles are processed in a fixed order, so the partials are bit-reproducible; the kernels combine the
  // four quadrants in a fixed order and add the result exactly (det_sum.cuh).
  float* stat_smem;
  int stat_stride;
  int ngrp = 2;                     // epilogue warps per TMEM lane quadrant: warp `grp` takes chunks grp, grp+ngrp, ...
  // TMA-store epilogue (N tile == 64 channels == one 128-byte row): instead of 32-byte global stores from every thread
  // (32 different lines per warp instruction: the L2 request path, not HBM, paces the narrow-N layers) the bf16 row
  // segments go into a 128B-swizzled [128 x 128 B] shared-memory tile that one thread hands to the TMA unit.
  uint32_t stage_smem = 0;          // shared-memory address of this tile's staging buffer (0: store to global)
  int stage_row = 0;                // this thread's row inside the staging tile (< 0: padding row, nothing staged)
  int stage_pitch = 128;            // bytes per staged row
  int stage_swizzle = 1;            // 1: 128-byte rows, SWIZZLE_128B (unit j at j ^ (row & 7)); 0: dense rows, no swizzle
  // kConvStats: bit j set <=> row (lane/4) + 8*j of this warp's 32 accumulator rows is a real output pixel
  uint32_t stat_mask = 0xFu;
};

// Row-validity mask for the statistics of this warp's rows base + (lane/4) + 8*j, j = 0..3, against a row limit.
__device__ __forceinline__ uint32_t stat_mask_below(long long base_row, int lane, long long limit) {
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) m |= (base_row + (lane >> 2) + 8 * j < limit) ? (1u << j) : 0u;
  return m;
}

// Issued BEFORE waiting for the accumulator: pulls this thread's residual row segments into L2 so that the residual
// loads of the epilogue proper do not expose DRAM latency (the epilogue of the residual convs was the bottleneck:
// 144 -> 64 temporal conv 385 us without, 563 us with the residual).
__device__ __forceinline__ void epilogue_prefetch_residual(const EpilogueArgs& p, int n0, long long out_row, int grp) {
  if (!(p.flags & kConvResidual) || out_row < 0) return;
  const __nv_bfloat16* rrow = p.residual + static_cast<size_t>(out_row) * p.cout_store + n0;
  for (int c = grp * 16; c < p.block_n && n0 + c < p.cout_store; c += 16 * p.ngrp)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(rrow + c));
}

// taddr: TMEM address of (this warp's first lane, column 0 of the tile); n0: first absolute channel of the tile;
// out_row: row of Y this thread's accumulator row maps to, or < 0 when the row is padding / out of range.
template <bool do_stats>
__device__ __forceinline__ void epilogue_chunks_impl(const EpilogueArgs& p, uint32_t taddr, int n0, long long out_row,
                                                     int grp, int lane) {
  const bool has_affine = p.scale_smem != nullptr;
  const bool has_res = (p.flags & kConvResidual) != 0;
  const bool relu = (p.flags & kConvRelu) != 0;
  const bool row_ok = out_row >= 0;
  const int n_chunks = (p.flags & kDbgNoEpilogue) ? 0 : (p.block_n >> 4);
  float* stat_smem = p.stat_smem;
  __nv_bfloat16* yrow = p.y + static_cast<size_t>(row_ok ? out_row : 0) * p.cout_store;
  const __nv_bfloat16* rrow = has_res ? p.residual + static_cast<size_t>(row_ok ? out_row : 0) * p.cout_store : nullptr;
  // software pipeline: TMEM load + residual load of chunk i+1 are in flight while chunk i is processed
  uint32_t v[16], vn[16];
  uint32_t sa[8], sb[8], san[8], sbn[8];          // statistics fragments (16x256b shape: 4 rows x 4 columns per thread)
  uint32_t rr[8] = {0, 0, 0, 0, 0, 0, 0, 0}, rn[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int ci = grp;
  if (ci < n_chunks) {
    ptx::tmem_ld_32x32b_x16(taddr + ci * 16, vn);
    if (do_stats) {
      ptx::tmem_ld_16x256b_x2(taddr + ci * 16, san);
      ptx::tmem_ld_16x256b_x2(taddr + (16u << 16) + ci * 16, sbn);
    }
    if (has_res && n0 + ci * 16 < p.cout_store) ptx::ld_global_nc_256(rrow + n0 + ci * 16, rn);
  }
  const int ngrp = p.ngrp;
  for (; ci < n_chunks; ci += ngrp) {
    ptx::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = vn[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) rr[i] = rn[i];
    if (do_stats) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { sa[i] = san[i]; sb[i] = sbn[i]; }
    }
    const int c = ci * 16;
    const int ch0 = n0 + c;
    const int cnext = ci + ngrp;
    if (cnext < n_chunks) {
      ptx::tmem_ld_32x32b_x16(taddr + cnext * 16, vn);
      if (do_stats) {
        ptx::tmem_ld_16x256b_x2(taddr + cnext * 16, san);
        ptx::tmem_ld_16x256b_x2(taddr + (16u << 16) + cnext * 16, sbn);
      }
      if (has_res && n0 + cnext * 16 < p.cout_store) ptx::ld_global_nc_256(rrow + n0 + cnext * 16, rn);
    }
    if (ch0 >= p.cout_store) continue;           // N tail (weights zero-padded to a whole tile)
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
    if (do_stats) {
      // Per-channel sum / sum^2 over this warp's 32 rows.  The accumulator chunk is read a second time in the 16x256b
      // shape, where a thread holds FOUR rows (lane/4 + 8j) of four columns (2*(lane%4) + {0, 1, 8, 9}): the rows are
      // added locally and only the 8 row-groups (lane bits 2..4) remain to be combined — 7 shuffles per chunk for both
      // quantities instead of 32 with one row per thread.  Statistics are those of the value actually stored
      // (bf16-rounded); rows beyond the tensor contribute 0 (p.stat_mask).
      float a8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a8[i] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {                 // row j: fragment (sa: rows 0..15, sb: rows 16..31), registers +2 for +8
        const bool live = ((p.stat_mask >> j) & 1u) != 0u;
#pragma unroll
        for (int h = 0; h < 2; ++h) {               // column pair h: registers 4h, 4h+1 (columns 2*(lane%4) + 8h, +1)
          const int reg = 4 * h + 2 * (j & 1);
          const uint32_t* frag = j < 2 ? sa : sb;
          // one packed conversion rounds both columns to bf16 (the stored values); unpack = shift / mask
          const uint32_t pk = pack_bf16x2(__uint_as_float(frag[reg]), __uint_as_float(frag[reg + 1]));
          const float r0 = live ? bf16_lo(pk) : 0.f, r1 = live ? bf16_hi(pk) : 0.f;
          a8[2 * h] += r0;     a8[4 + 2 * h] = fmaf(r0, r0, a8[4 + 2 * h]);
          a8[2 * h + 1] += r1; a8[5 + 2 * h] = fmaf(r1, r1, a8[5 + 2 * h]);
        }
      }
      // recursive halving over lane bits 4, 3, 2: 8 -> 4 -> 2 -> 1 values
      {
        const bool up = (lane & 16) != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float send = up ? a8[i] : a8[i + 4];
          const float keep = up ? a8[i + 4] : a8[i];
          a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
      }
      {
        const bool up = (lane & 8) != 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float send = up ? a8[i] : a8[i + 2];
          const float keep = up ? a8[i + 2] : a8[i];
          a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
      }
      {
        const bool up = (lane & 4) != 0;
        const float send = up ? a8[0] : a8[1];
        const float keep = up ? a8[1] : a8[0];
        a8[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      // lane bit 4: quantity (sum / sum^2); bits 3,2: which of this thread's four columns; bits 1,0: column pair
      const int kcol = ((lane & 8) ? 2 : 0) + ((lane & 4) ? 1 : 0);
      const int col = 2 * (lane & 3) + (kcol & 1) + ((kcol & 2) ? 8 : 0);
      stat_smem[((lane & 16) ? p.stat_stride : 0) + c + col] += a8[0];      // this warp owns the slot: no atomic
    }
    if (has_affine) {
      const float4* sc4 = reinterpret_cast<const float4*>(p.scale_smem + ch0);
      const float4* sh4 = reinterpret_cast<const float4*>(p.shift_smem + ch0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = sc4[i], b = sh4[i];
        f[4 * i + 0] = fmaf(f[4 * i + 0], a.x, b.x);
        f[4 * i + 1] = fmaf(f[4 * i + 1], a.y, b.y);
        f[4 * i + 2] = fmaf(f[4 * i + 2], a.z, b.z);
        f[4 * i + 3] = fmaf(f[4 * i + 3], a.w, b.w);
      }
    }
    if ((row_ok || p.stage_smem != 0u) && !(p.flags & kDbgNoStore)) {
      if (has_res) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          f[2 * i] += bf16_lo(rr[i]);
          f[2 * i + 1] += bf16_hi(rr[i]);
        }
      }
      if (relu) {
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
      }
      uint32_t o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
      if (p.stage_smem != 0u) {
        if (p.stage_row >= 0) {
          // chunk ci covers the 16-byte units 2ci, 2ci+1 of the row; SWIZZLE_128B stores unit j at j ^ (row & 7)
          const uint32_t rowb = p.stage_smem + static_cast<uint32_t>(p.stage_row) * static_cast<uint32_t>(p.stage_pitch);
          const uint32_t sw = p.stage_swizzle ? (static_cast<uint32_t>(p.stage_row) & 7u) : 0u;
          ptx::st_shared_128(rowb + (((2u * ci) ^ sw) << 4), o[0], o[1], o[2], o[3]);
          ptx::st_shared_128(rowb + (((2u * ci + 1u) ^ sw) << 4), o[4], o[5], o[6], o[7]);
        }
      } else {
        ptx::st_global_256(yrow + ch0, o);          // one full 32-byte sector per thread
      }
    }
  }
}

// Data-gradient epilogue fused with the first pass of the BatchNorm backward it feeds (kConvBnBwd).  The convolution
// computes dact = d(loss)/d(activation) of a layer whose activation was relu(raw*scale + shift) (BatchNorm + ReLU on the raw
// conv output `raw`).  Instead of storing dact and letting bn_bwd_reduce read dact and raw again, the epilogue
//   * reads this thread's row of `raw` (through p.residual, like a residual operand),
//   * masks: dz = (raw*scale + shift > 0) ? dact : 0  (scale / shift staged in shared memory like the folded affine),
//   * stores dz (bf16) and accumulates the per-channel sums of the STORED values: sum dz*raw and sum dz (quantities 0 and 1 of
//     the statistics partials; bn_bwd_apply turns them into dgamma = inv_std*(sum dz*raw - mean*sum dz) and dbeta = sum dz).
// One accumulator row per thread, so a column sum is a reduction over the warp's 32 lanes: recursive halving over the 32
// values (16 columns x 2 quantities) leaves lane L with the sum of value L after 16 + 8 + 4 + 2 + 1 = 31 shuffles.
__device__ __forceinline__ void epilogue_chunks_bnbwd(const EpilogueArgs& p, uint32_t taddr, int n0, long long out_row, int grp,
                                                      int lane) {
  const bool row_ok = out_row >= 0;
  const int n_chunks = p.block_n >> 4;
  float* stat_smem = p.stat_smem;
  __nv_bfloat16* yrow = p.y + static_cast<size_t>(row_ok ? out_row : 0) * p.cout_store;
  const __nv_bfloat16* rrow = p.residual + static_cast<size_t>(row_ok ? out_row : 0) * p.cout_store;
  uint32_t v[16], vn[16];
  uint32_t rr[8] = {0, 0, 0, 0, 0, 0, 0, 0}, rn[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int ci = grp;
  if (ci < n_chunks) {
    ptx::tmem_ld_32x32b_x16(taddr + ci * 16, vn);
    if (row_ok && n0 + ci * 16 < p.cout_store) ptx::ld_global_nc_256(rrow + n0 + ci * 16, rn);
  }
  const int ngrp = p.ngrp;
  for (; ci < n_chunks; ci += ngrp) {
    ptx::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = vn[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) rr[i] = rn[i];
    const int c = ci * 16;
    const int ch0 = n0 + c;
    const int cnext = ci + ngrp;
    if (cnext < n_chunks) {
      ptx::tmem_ld_32x32b_x16(taddr + cnext * 16, vn);
      if (row_ok && n0 + cnext * 16 < p.cout_store) ptx::ld_global_nc_256(rrow + n0 + cnext * 16, rn);
    }
    if (ch0 >= p.cout_store) continue;           // N tail (weights zero-padded to a whole tile)
    float vals[32];
    uint32_t o[8];
    {
      const float4* sc4 = reinterpret_cast<const float4*>(p.scale_smem + ch0);
      const float4* sh4 = reinterpret_cast<const float4*>(p.shift_smem + ch0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = sc4[i], b = sh4[i];
        const float as[4] = {a.x, a.y, a.z, a.w}, bs[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          const int e = 4 * i + j;
          const float r0 = bf16_lo(rr[e >> 1]), r1 = bf16_hi(rr[e >> 1]);
          const float g0 = (row_ok && fmaf(r0, as[j], bs[j]) > 0.f) ? __uint_as_float(v[e]) : 0.f;
          const float g1 = (row_ok && fmaf(r1, as[j + 1], bs[j + 1]) > 0.f) ? __uint_as_float(v[e + 1]) : 0.f;
          const uint32_t pk = pack_bf16x2(g0, g1);
          o[e >> 1] = pk;
          const float q0 = bf16_lo(pk), q1 = bf16_hi(pk);        // the stored values
          vals[e] = q0 * r0;      vals[e + 1] = q1 * r1;
          vals[16 + e] = q0;      vals[16 + e + 1] = q1;
        }
      }
    }
    if (row_ok && !(p.flags & kDbgNoStore)) ptx::st_global_256(yrow + ch0, o);
    // recursive halving over lane bits 4..0: lane L ends with the warp's sum of vals[L]
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
      const bool up = (lane & half) != 0;
#pragma unroll
      for (int i = 0; i < half; ++i) {
        const float send = up ? vals[i] : vals[i + half];
        const float keep = up ? vals[i + half] : vals[i];
        vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
      }
    }
    stat_smem[((lane & 16) ? p.stat_stride : 0) + c + (lane & 15)] += vals[0];      // this warp owns the slot: no atomic
  }
}

// Two instantiations: the statistics variant (training forward) carries the extra TMEM fragments and shuffles; the
// inference variant must not pay for them in registers or scheduling.
__device__ __forceinline__ void epilogue_chunks(const EpilogueArgs& p, uint32_t taddr, int n0, long long out_row,
                                                int grp, int lane) {
  if (p.flags & kConvBnBwd) epilogue_chunks_bnbwd(p, taddr, n0, out_row, grp, lane);
  else if (p.flags & kConvStats) epilogue_chunks_impl<true>(p, taddr, n0, out_row, grp, lane);
  else epilogue_chunks_impl<false>(p, taddr, n0, out_row, grp, lane);
}

}  // namespace fvt
