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

namespace fvt {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;            // bf16 elements per 128-byte swizzle row
constexpr int kConvThreads = 384;          // warps 0-3: TMA / MMA / TMEM alloc / spare; warps 4-11: epilogue
constexpr int kEpilogueThreads = 256;
constexpr int kMaxCout = 1536;             // per-channel scale/shift staged in shared memory
constexpr int kMaxStages = 8;
constexpr int kATileBytes = kBlockM * kBlockK * 2;   // 16 KiB

enum ConvFlags : int {
  kConvRelu = 1,
  kConvResidual = 2,
  kConvStats = 4,
};

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
  const float* scale;     // [cout_store] or nullptr (identity)
  const float* shift;     // [cout_store] or nullptr
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
  float* stats;           // [2][cout_store]: sum, sum of squares (atomically accumulated) or nullptr
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_fwd_kernel(const __grid_constant__ CUtensorMap tmap_x,
                      const __grid_constant__ CUtensorMap tmap_w,
                      const ConvKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.stages;
  const int b_tile_bytes = p.block_n * kBlockK * 2;
  const int stage_bytes = kATileBytes + b_tile_bytes;

  uint8_t* smem_tiles = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* full_bar = bars;                       // [stages]
  uint64_t* empty_bar = bars + kMaxStages;         // [stages]
  uint64_t* acc_full_bar = bars + 2 * kMaxStages;  // [2]
  uint64_t* acc_empty_bar = acc_full_bar + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty_bar + 2);
  float* stat_smem = reinterpret_cast<float*>(tmem_slot + 4);   // [2][256] per-CTA channel partials
  float* affine_smem = stat_smem + 512;                          // [2][kMaxCout] scale, shift

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_x);
    ptx::prefetch_tensormap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&acc_full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&acc_empty_bar[s]), 8);   // one arrive per epilogue warp
    }
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
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int taps = p.kt * p.kh * p.kw;
  const int k_blocks = taps * p.cin_blocks;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
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
        const int n0 = n_blk * p.block_n;
        for (int dt = 0; dt < p.kt; ++dt) {
          for (int dh = 0; dh < p.kh; ++dh) {
            for (int dw = 0; dw < p.kw; ++dw) {
              const int tap = (dt * p.kh + dh) * p.kw + dw;
              for (int cb = 0; cb < p.cin_blocks; ++cb) {
                ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1);
                const uint32_t fb = ptx::smem_u32(&full_bar[stage]);
                ptx::mbar_arrive_expect_tx(fb, stage_bytes);
                uint8_t* a_dst = smem_tiles + stage * stage_bytes;
                ptx::tma_load_im2col_5d(ptx::smem_u32(a_dst), &tmap_x, fb, cb * kBlockK, cw, ch, cd, on,
                                        static_cast<uint16_t>(dw), static_cast<uint16_t>(dh),
                                        static_cast<uint16_t>(dt));
                ptx::tma_load_2d(ptx::smem_u32(a_dst + kATileBytes), &tmap_w, fb,
                                 tap * p.k_per_tap + cb * kBlockK, n0);
                if (++stage == stages) { stage = 0; phase ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (single thread)
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(kBlockM, p.block_n, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(ptx::smem_u32(&acc_empty_bar[acc]), acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = 0; kb < k_blocks; ++kb) {
          const int cb = kb % p.cin_blocks;
          int k16 = p.cin_k16 - cb * (kBlockK / 16);
          if (k16 > kBlockK / 16) k16 = kBlockK / 16;
          ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem_tiles + stage * stage_bytes);
          const uint64_t a_desc = ptx::make_sw128_desc(a_addr, 16, 1024);
          const uint64_t b_desc = ptx::make_sw128_desc(a_addr + kATileBytes, 16, 1024);
          for (int k = 0; k < k16; ++k) {
            // +32 bytes (16 bf16) along K inside the swizzle atom == +2 in the (addr >> 4) field
            ptx::umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
          ptx::umma_commit(ptx::smem_u32(&empty_bar[stage]));
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(ptx::smem_u32(&acc_full_bar[acc]));
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
    const bool do_stats = (p.flags & kConvStats) != 0;
    const bool has_affine = p.scale != nullptr;
    const bool has_res = (p.flags & kConvResidual) != 0;
    const bool relu = (p.flags & kConvRelu) != 0;
    const int et = threadIdx.x - 128;           // 0..255
    const int n_chunks = p.block_n >> 4;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.num_n_tiles;
      const int n_blk = tile - m_blk * p.num_n_tiles;
      const int n0 = n_blk * p.block_n;
      const int row = m_blk * kBlockM + q * 32 + lane;
      const bool row_ok = row < p.m_total;
      if (do_stats) {
        for (int i = et; i < 512; i += kEpilogueThreads) stat_smem[i] = 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      ptx::mbar_wait(ptx::smem_u32(&acc_full_bar[acc]), acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
      __nv_bfloat16* yrow = p.y + static_cast<size_t>(row_ok ? row : 0) * p.cout_store;
      const __nv_bfloat16* rrow = has_res ? p.residual + static_cast<size_t>(row_ok ? row : 0) * p.cout_store : nullptr;

      // software pipeline: TMEM load + residual load of chunk i+1 are in flight while chunk i is processed
      uint32_t v[16], vn[16];
      uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0, rn0 = r0, rn1 = r0;
      int ci = grp;
      if (ci < n_chunks) {
        ptx::tmem_ld_32x32b_x16(taddr + ci * 16, vn);
        if (has_res && n0 + ci * 16 < p.cout_store) {
          rn0 = __ldg(reinterpret_cast<const uint4*>(rrow + n0 + ci * 16));
          rn1 = __ldg(reinterpret_cast<const uint4*>(rrow + n0 + ci * 16 + 8));
        }
      }
      for (; ci < n_chunks; ci += 2) {
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = vn[i];
        r0 = rn0; r1 = rn1;
        const int c = ci * 16;
        const int ch0 = n0 + c;
        const int cnext = ci + 2;
        if (cnext < n_chunks) {
          ptx::tmem_ld_32x32b_x16(taddr + cnext * 16, vn);
          if (has_res && n0 + cnext * 16 < p.cout_store) {
            rn0 = __ldg(reinterpret_cast<const uint4*>(rrow + n0 + cnext * 16));
            rn1 = __ldg(reinterpret_cast<const uint4*>(rrow + n0 + cnext * 16 + 8));
          }
        }
        if (ch0 >= p.cout_store) continue;           // N tail (weights zero-padded to a whole tile)
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
        if (do_stats) {
          // per-channel sum / sum^2 over this warp's 32 rows: recursive-halving butterfly, 16 values -> 1 per lane pair
          float s1[16], s2[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            // statistics of the value that is actually stored (bf16-rounded); rows beyond M contribute 0
            float r = row_ok ? __bfloat162float(__float2bfloat16_rn(f[i])) : 0.f;
            s1[i] = r; s2[i] = r * r;
          }
#pragma unroll
          for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
            const bool upper = (lane & bit) != 0;
#pragma unroll
            for (int i = 0; i < half; ++i) {
              const float send1 = upper ? s1[i] : s1[i + half];
              const float keep1 = upper ? s1[i + half] : s1[i];
              const float send2 = upper ? s2[i] : s2[i + half];
              const float keep2 = upper ? s2[i + half] : s2[i];
              s1[i] = keep1 + __shfl_xor_sync(0xffffffffu, send1, bit);
              s2[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, bit);
            }
          }
          s1[0] += __shfl_xor_sync(0xffffffffu, s1[0], 1);
          s2[0] += __shfl_xor_sync(0xffffffffu, s2[0], 1);
          if ((lane & 1) == 0) {
            const int chl = ((lane & 16) ? 8 : 0) + ((lane & 8) ? 4 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0);
            atomicAdd(&stat_smem[c + chl], s1[0]);
            atomicAdd(&stat_smem[256 + c + chl], s2[0]);
          }
        }
        if (has_affine) {
          const float4* sc4 = reinterpret_cast<const float4*>(affine_smem + ch0);
          const float4* sh4 = reinterpret_cast<const float4*>(affine_smem + kMaxCout + ch0);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 a = sc4[i], b = sh4[i];
            f[4 * i + 0] = fmaf(f[4 * i + 0], a.x, b.x);
            f[4 * i + 1] = fmaf(f[4 * i + 1], a.y, b.y);
            f[4 * i + 2] = fmaf(f[4 * i + 2], a.z, b.z);
            f[4 * i + 3] = fmaf(f[4 * i + 3], a.w, b.w);
          }
        }
        if (row_ok) {
          if (has_res) {
            const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
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
          uint4 o0, o1;
          o0.x = pack_bf16x2(f[0], f[1]);   o0.y = pack_bf16x2(f[2], f[3]);
          o0.z = pack_bf16x2(f[4], f[5]);   o0.w = pack_bf16x2(f[6], f[7]);
          o1.x = pack_bf16x2(f[8], f[9]);   o1.y = pack_bf16x2(f[10], f[11]);
          o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
          *reinterpret_cast<uint4*>(yrow + ch0) = o0;
          *reinterpret_cast<uint4*>(yrow + ch0 + 8) = o1;
        }
      }
      // release the accumulator stage (all of this warp's TMEM reads have completed: wait::ld above)
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&acc_empty_bar[acc]));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (do_stats) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = et; i < p.block_n; i += kEpilogueThreads) {
          if (n0 + i < p.cout_store) {
            atomicAdd(p.stats + n0 + i, stat_smem[i]);
            atomicAdd(p.stats + p.cout_store + n0 + i, stat_smem[256 + i]);
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
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
