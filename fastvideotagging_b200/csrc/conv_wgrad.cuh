// K3 — convolution weight gradient on tcgen05:   dW[co, ci, tap] += sum_pixels dY[pixel, co] * X[pixel + tap, ci]
//
// The reduction (GEMM-K) dimension is the output-pixel index, and both operands are stored pixel-major
// ([pixels, channels] NDHWC), i.e. "MN-major" in UMMA terms: a [64 pixels x 64 channels] slab brought in by one TMA
// load (128-byte rows, 128B swizzle) is directly a 64(MN) x 64(K) operand block.  The X side uses the same im2col
// tensor map as the forward kernel (filter-tap offsets, zero-filled halo); the dY side uses an im2col map with a
// 1x1x1 window, which is simply "64 consecutive pixels".
//
//   mode 0 (Cin multiple of 64: 1x3x3 / 1x1x1 / stem):  GEMM-M = (tap, ci) in 64-channel groups, two groups per
//           128-row tile; GEMM-N = co tile (<= 256).
//   mode 1 (Cout multiple of 64: 3x1x1 convs):           GEMM-M = co groups; GEMM-N = ci tile of one tap.
//
// One CTA = one (M tile, N tile[, tap], pixel split) work item: TMA producer warp, single-thread MMA issuer,
// fp32 accumulator [128 x n_tile] in TMEM, 4 epilogue warps that add the tile into the fp32 gradient tensor (in the
// reference's (O, I, kT, kH, kW) layout) with atomics — pixel splits of the same tile meet there.
// Replaces the cuDNN backward-filter calls MXNet issues for the Conv3D layers of model/R2Plus1.py / net.py.
#pragma once
#include "ptx.cuh"
#include "pdl.cuh"

namespace fvt {

constexpr int kWgradThreads = 256;      // warp0 + warp3 TMA, warp1 MMA, warp2 TMEM alloc, warps 4-7 epilogue
constexpr int kWgPix = 64;              // pixels per k-block
constexpr int kSlabBytes = kWgPix * 128;   // one [64 px x 64 ch] slab
constexpr int kWgMaxStages = 6;

struct WgradParams {
  int m_total;
  int to, ho, wo;
  int st, sh, sw, pt, ph, pw;
  int kt, kh, kw;
  int mode;
  int m_groups;          // 64-channel groups on the M side
  int m_tiles;           // ceil(m_groups / 2)
  int cin_blocks;        // 64-channel blocks per tap on the X side
  int n_tile, n_tiles;   // N tile width / count (mode 0: over cout; mode 1: over cin of one tap)
  int n_loads;           // ceil(n_tile / 64)
  int taps;
  int splits, kblocks_per_split, kblocks_total;
  int cin_real, cout_real;
  int stages;
  float* dw;
  int w_ohwi;            // dw layout (O, taps, I) instead of (O, I, taps)
  int dbg_no_store;      // experiments only: epilogue reads TMEM, stores nothing
  // split reduction as in conv_wgrad_slab.cuh: splits > 1 store into ws[split][...] (reduced in split order by the host's
  // wgrad_reduce pass), a single split stores straight into dw; no atomics
  float* ws;
  long long ws_split_stride;
};

__global__ void __launch_bounds__(kWgradThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                  const WgradParams p) {
  fvt_pdl_entry();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.stages;
  const int stage_bytes = (2 + p.n_loads) * kSlabBytes;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kWgMaxStages;
  uint64_t* acc_bar = bars + 2 * kWgMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  // ---- decode the work item
  int item = blockIdx.x;
  const int split = item % p.splits;  item /= p.splits;
  const int nt = item % p.n_tiles;    item /= p.n_tiles;
  int tap1 = 0;
  if (p.mode == 1) { tap1 = item % p.taps; item /= p.taps; }
  const int mt = item;
  const int kb0 = split * p.kblocks_per_split;
  int kb1 = kb0 + p.kblocks_per_split;
  if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
  const int g0 = mt * 2;
  const int m_valid_groups = (g0 + 1 < p.m_groups) ? 2 : 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_x);
    ptx::prefetch_tensormap(&tmap_dy);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    ptx::mbar_init(ptx::smem_u32(acc_bar), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 || warp == 3) {
    // Two producer warps: a TMA box costs its issuing thread ~275 clk whatever its size (profiles/r01_tma_rate_microbench.log),
    // and a 64-pixel k-block needs up to six 8 KB boxes — one thread made the loop issue-bound.  Warp 0 posts the stage's
    // byte count and loads the M side, warp 3 loads the N side; both complete on the stage's full barrier.
    {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = (m_valid_groups + p.n_loads) * kSlabBytes;
      for (int kb = kb0; kb < kb1; ++kb) {
        int m0 = kb * kWgPix;
        const int ow = m0 % p.wo;  m0 /= p.wo;
        const int oh = m0 % p.ho;  m0 /= p.ho;
        const int ot = m0 % p.to;
        const int on = m0 / p.to;
        const int cw = ow * p.sw - p.pw, ch = oh * p.sh - p.ph, cd = ot * p.st - p.pt;
        ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1);
        const uint32_t fb = ptx::smem_u32(&full_bar[stage]);
        uint8_t* base = smem + stage * stage_bytes;
        if (ptx::elect_one()) {
          if (warp == 0) {
            ptx::mbar_arrive_expect_tx(fb, tx_bytes);
            // ---- M side
            for (int g = 0; g < m_valid_groups; ++g) {
              const uint32_t dst = ptx::smem_u32(base + g * kSlabBytes);
              if (p.mode == 0) {
                const int grp = g0 + g;
                const int tap = grp / p.cin_blocks, cb = grp - tap * p.cin_blocks;
                const int dw_ = tap % p.kw, dh_ = (tap / p.kw) % p.kh, dt_ = tap / (p.kw * p.kh);
                ptx::tma_load_im2col_5d(dst, &tmap_x, fb, cb * 64, cw, ch, cd, on, (uint16_t)dw_, (uint16_t)dh_, (uint16_t)dt_);
              } else {
                ptx::tma_load_im2col_5d(dst, &tmap_dy, fb, (g0 + g) * 64, ow, oh, ot, on, 0, 0, 0);
              }
            }
          } else {
            // ---- N side
            for (int j = 0; j < p.n_loads; ++j) {
              const uint32_t dst = ptx::smem_u32(base + (2 + j) * kSlabBytes);
              const int c0 = nt * p.n_tile + j * 64;
              if (p.mode == 0) {
                ptx::tma_load_im2col_5d(dst, &tmap_dy, fb, c0, ow, oh, ot, on, 0, 0, 0);
              } else {
                const int dw_ = tap1 % p.kw, dh_ = (tap1 / p.kw) % p.kh, dt_ = tap1 / (p.kw * p.kh);
                ptx::tma_load_im2col_5d(dst, &tmap_x, fb, c0, cw, ch, cd, on, (uint16_t)dw_, (uint16_t)dh_, (uint16_t)dt_);
              }
            }
          }
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc = ptx::make_idesc_bf16(128, p.n_tile, 1, 1);    // both operands MN-major
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
        ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(smem + stage * stage_bytes);
        const uint32_t b_addr = a_addr + 2 * kSlabBytes;
        // MN-major SW128: LBO = distance between 64-channel groups (one slab), SBO = 8 pixel rows = 1024 B
        const uint64_t a_desc = ptx::make_sw128_desc(a_addr, kSlabBytes, 1024);
        const uint64_t b_desc = ptx::make_sw128_desc(b_addr, kSlabBytes, 1024);
        if (ptx::elect_one()) {
          // 16 pixels = 2 swizzle atoms = 2048 B along K  ->  +128 in the (addr >> 4) field
          ptx::umma_bf16_ss(tmem_base, a_desc, b_desc, idesc, kb > kb0 ? 1u : 0u);
          ptx::umma_bf16_ss(tmem_base, a_desc + 128, b_desc + 128, idesc, 1u);
          ptx::umma_bf16_ss(tmem_base, a_desc + 256, b_desc + 256, idesc, 1u);
          ptx::umma_bf16_ss(tmem_base, a_desc + 384, b_desc + 384, idesc, 1u);
          ptx::umma_commit(ptx::smem_u32(&empty_bar[stage]));
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (ptx::elect_one()) ptx::umma_commit(ptx::smem_u32(acc_bar));
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    if (kb1 > kb0) {
      ptx::mbar_wait(ptx::smem_u32(acc_bar), 0);
      ptx::tc_fence_after();
      const int r = q * 32 + lane;                       // row of the 128-row tile
      const int grp = g0 + (r >> 6);
      bool row_ok;
      int ci = 0, tap = 0, co = 0;
      if (p.mode == 0) {
        tap = grp / p.cin_blocks;
        ci = (grp - tap * p.cin_blocks) * 64 + (r & 63);
        row_ok = grp < p.m_groups && ci < p.cin_real;
      } else {
        co = grp * 64 + (r & 63);
        tap = tap1;
        row_ok = grp < p.m_groups && co < p.cout_real;
      }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      for (int c = 0; c < p.n_tile; c += 16) {
        uint32_t v[16];
        ptx::tmem_ld_32x32b_x16(taddr + c, v);
        ptx::tmem_ld_wait();
        const int nbase = nt * p.n_tile + c;
        float* dst0 = p.ws != nullptr ? p.ws + static_cast<size_t>(split) * p.ws_split_stride : p.dw;
        if (row_ok && !p.dbg_no_store && p.mode == 1 && p.w_ohwi && (p.cin_real & 3) == 0 && nbase + 16 <= p.cin_real) {
          // (O, taps, I) layout, input channels along N: this thread's 16 columns are 16 consecutive floats
          float4* dst = reinterpret_cast<float4*>(dst0 + (static_cast<size_t>(co) * p.taps + tap) * p.cin_real + nbase);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                 __uint_as_float(v[4 * i + 3]));
        } else if (row_ok && !p.dbg_no_store) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int nidx = nbase + j;
            if (p.mode == 0) {
              if (nidx < p.cout_real)
                dst0[p.w_ohwi ? (static_cast<size_t>(nidx) * p.taps + tap) * p.cin_real + ci
                              : (static_cast<size_t>(nidx) * p.cin_real + ci) * p.taps + tap] = __uint_as_float(v[j]);
            } else {
              if (nidx < p.cin_real)
                dst0[p.w_ohwi ? (static_cast<size_t>(co) * p.taps + tap) * p.cin_real + nidx
                              : (static_cast<size_t>(co) * p.cin_real + nidx) * p.taps + tap] = __uint_as_float(v[j]);
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace fvt
