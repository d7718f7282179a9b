// K5-K7, K12 and friends: the HBM-bound passes of a training step over NDHWC bf16 activations [rows, C].
//
//   K5  bn_finalize          per-channel (sum, sum^2) -> mean / biased var -> scale, shift, running-stat update
//   K6  bn_apply             act = relu?( raw*scale + shift  [+ res  |  + res_raw*res_scale + res_shift] )
//   K7a bn_bwd_reduce        dgamma, dbeta = sum_rows( dz * xhat ), sum_rows( dz ),  dz = dact * [mask > 0]
//   K7b bn_bwd_apply         draw = gamma*inv_std * ( dz - dbeta/M - xhat*dgamma/M )   (and optionally dz itself)
//       zero_insert          dY -> dY placed on the stride lattice of the conv input (dgrad of strided convs)
//       pool_fc_bwd          dlogits -> dW_fc, db_fc, d(conv5 output)
//   K12 sgd_momentum_multi   MXNet sgd_mom_update over a list of tensors in one launch
//
// Thread mapping for the [rows, C] passes: one thread owns 8 consecutive channels (one 16-byte vector) of a row;
// consecutive threads walk the channel vectors of a row, then the next row, so every warp access is a run of
// contiguous 16-byte vectors (fully coalesced).  A thread's channel group is fixed across its row loop
// (blockDim.x is a multiple of C/8 ... enforced by the launcher), so per-channel parameters live in registers and the
// reductions accumulate per thread, finish with a shared-memory fold and ONE atomic per channel per CTA.
// MXNet semantics restated: reference model/R2Plus1.py:32,59,62,71 (nn.BatchNorm defaults), net.py:44-45 (eps=1e-3),
// biased variance, running = momentum*running + (1-momentum)*batch.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string.h>

#include "../../include/fvt_b200.h"
#include "host_common.h"
#include "det_sum.cuh"
#include "pdl.cuh"

namespace fvt {

constexpr int kUnroll = 4;      // rows per thread per loop iteration in the [rows, C] passes

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&b);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// ------------------------------------------------------------------------------------------------ K5
__global__ void bn_finalize_kernel(const unsigned long long* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, int c_store, int c_real, double inv_rows, float eps,
                                   float momentum, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ invstd_out) {
  fvt_pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= c_store) return;
  if (c >= c_real) {                      // pad channels: identity-zero so they stay exactly 0
    scale[c] = 0.f; shift[c] = 0.f; mean_out[c] = 0.f; invstd_out[c] = 0.f;
    return;
  }
  const double m = det_read(stats + static_cast<size_t>(c) * kDetLimbs) * inv_rows;
  double v = det_read(stats + static_cast<size_t>(c_store + c) * kDetLimbs) * inv_rows - m * m;
  if (v < 0.0) v = 0.0;
  const float inv_std = static_cast<float>(1.0 / sqrt(v + static_cast<double>(eps)));
  const float g = gamma[c];
  scale[c] = g * inv_std;
  shift[c] = beta[c] - static_cast<float>(m) * g * inv_std;
  mean_out[c] = static_cast<float>(m);
  invstd_out[c] = inv_std;
  if (running_mean != nullptr) {
    running_mean[c] = momentum * running_mean[c] + (1.f - momentum) * static_cast<float>(m);
    running_var[c] = momentum * running_var[c] + (1.f - momentum) * static_cast<float>(v);
  }
}

// ------------------------------------------------------------------------------------------------ K6 / K7
// The [rows, C] passes are HBM-bound, so what matters is bytes in flight per SM.  The first version kept every
// per-channel constant of a thread's 8 channels in registers (92-141 registers per thread -> 1-2 CTAs per SM, 12-24 %
// of the warp slots, 38-50 % of DRAM bandwidth in ncu).  Here the constants sit in shared memory (computed once per
// CTA) and are read two channels at a time inside the compute phase, which brings the kernels to <= 64 registers
// (4 CTAs of 256 threads per SM) with 4 rows x 2-3 16-byte loads in flight per thread.
struct Words { uint32_t w[4]; };
__device__ __forceinline__ Words ld_words(const uint4* p) {
  const uint4 v = __ldg(p);
  Words r; r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
  return r;
}
__device__ __forceinline__ uint4 to_uint4(const Words& a) { return make_uint4(a.w[0], a.w[1], a.w[2], a.w[3]); }
__device__ __forceinline__ float lo_f(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float hi_f(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// res_mode: 0 none, 1 add bf16 tensor `res`, 2 add res*res_scale + res_shift (projection shortcut's own BatchNorm)
// Fused K5: when `fin.stats` is set, every CTA derives (scale, shift) from the accumulated (sum, sum^2) itself — the
// same arithmetic as bn_finalize_kernel — and CTA 0 also publishes scale/shift/mean/inv_std for the backward pass and
// updates the running statistics, so the forward pass needs no separate finalize launch.
struct BnFinalizeArgs {
  const unsigned long long* stats; const float* gamma; const float* beta;
  float* running_mean; float* running_var;
  float* scale_out; float* shift_out; float* mean_out; float* invstd_out;
  int c_real; double inv_rows; float eps, momentum;
};

template <int kResMode>
__global__ void __launch_bounds__(256, 4)
bn_apply_kernel(const uint4* __restrict__ raw, const float* __restrict__ scale, const float* __restrict__ shift,
                const uint4* __restrict__ res, const float* __restrict__ res_scale, const float* __restrict__ res_shift,
                uint4* __restrict__ out, size_t rows, int cvec, int relu, const BnFinalizeArgs fin) {
  fvt_pdl_entry();
  extern __shared__ float cst[];                     // [4][C]: scale, shift, res_scale, res_shift
  const int c_store = cvec * 8;
  for (int ch = threadIdx.x; ch < c_store; ch += blockDim.x) {
    if (fin.stats != nullptr) {
      float sc = 0.f, sh = 0.f, mf = 0.f, inv_std = 0.f, vf = 0.f;
      if (ch < fin.c_real) {                         // pad channels: identity-zero so they stay exactly 0
        const double m = det_read(fin.stats + static_cast<size_t>(ch) * kDetLimbs) * fin.inv_rows;
        double v = det_read(fin.stats + static_cast<size_t>(c_store + ch) * kDetLimbs) * fin.inv_rows - m * m;
        if (v < 0.0) v = 0.0;
        inv_std = static_cast<float>(1.0 / sqrt(v + static_cast<double>(fin.eps)));
        const float g = fin.gamma[ch];
        mf = static_cast<float>(m); vf = static_cast<float>(v);
        sc = g * inv_std;
        sh = fin.beta[ch] - mf * g * inv_std;
      }
      cst[ch] = sc;
      cst[c_store + ch] = sh;
      if (blockIdx.x == 0) {
        fin.scale_out[ch] = sc; fin.shift_out[ch] = sh; fin.mean_out[ch] = mf; fin.invstd_out[ch] = inv_std;
        if (fin.running_mean != nullptr && ch < fin.c_real) {
          fin.running_mean[ch] = fin.momentum * fin.running_mean[ch] + (1.f - fin.momentum) * mf;
          fin.running_var[ch] = fin.momentum * fin.running_var[ch] + (1.f - fin.momentum) * vf;
        }
      }
    } else {
      cst[ch] = scale[ch];
      cst[c_store + ch] = shift[ch];
    }
    if (kResMode == 2) { cst[2 * c_store + ch] = res_scale[ch]; cst[3 * c_store + ch] = res_shift[ch]; }
  }
  __syncthreads();
  const int tpr = blockDim.x / cvec;                 // rows handled per CTA iteration
  const int cv = threadIdx.x % cvec;
  const int rsub = threadIdx.x / cvec;
  if (rsub >= tpr) return;
  const float2* sc2 = reinterpret_cast<const float2*>(cst) + cv * 4;
  const float2* sh2 = reinterpret_cast<const float2*>(cst + c_store) + cv * 4;
  const float2* rs2 = reinterpret_cast<const float2*>(cst + 2 * c_store) + cv * 4;
  const float2* rh2 = reinterpret_cast<const float2*>(cst + 3 * c_store) + cv * 4;
  const float lo_clamp = relu ? 0.f : -INFINITY;
  const size_t rstep = static_cast<size_t>(gridDim.x) * tpr;
  for (size_t r0 = static_cast<size_t>(blockIdx.x) * tpr + rsub; r0 < rows; r0 += rstep * kUnroll) {
    Words vx[kUnroll], vq[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t r = r0 + u * rstep;
      if (r < rows) {
        vx[u] = ld_words(raw + r * cvec + cv);
        if (kResMode) vq[u] = ld_words(res + r * cvec + cv);
      }
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float2 sc = sc2[w], sh = sh2[w];
      float2 rs = make_float2(1.f, 1.f), rh = make_float2(0.f, 0.f);
      if (kResMode == 2) { rs = rs2[w]; rh = rh2[w]; }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        float y0 = fmaf(lo_f(vx[u].w[w]), sc.x, sh.x), y1 = fmaf(hi_f(vx[u].w[w]), sc.y, sh.y);
        if (kResMode) {
          y0 += fmaf(lo_f(vq[u].w[w]), rs.x, rh.x);
          y1 += fmaf(hi_f(vq[u].w[w]), rs.y, rh.y);
        }
        vx[u].w[w] = pack2(fmaxf(y0, lo_clamp), fmaxf(y1, lo_clamp));
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t r = r0 + u * rstep;
      if (r < rows) out[r * cvec + cv] = to_uint4(vx[u]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ K7a
// acc[0..C) += sum dz*(raw - mean)  (dgamma = inv_std * that), acc[C..2C) += sum dz (dbeta) — exact accumulators
// kMask: 0 no mask, 1 mask tensor (dz = dact * [mask > 0]), 2 ReLU mask recomputed from raw: [raw*scale + shift > 0]
template <int kMask>
__global__ void __launch_bounds__(256, 3)
bn_bwd_reduce_kernel(const uint4* __restrict__ raw, const uint4* __restrict__ dact, const uint4* __restrict__ mask,
                     const float* __restrict__ mean,
                     const float* __restrict__ relu_scale, const float* __restrict__ relu_shift,
                     unsigned long long* __restrict__ acc, size_t rows, int cvec, int c_store) {
  fvt_pdl_entry();
  extern __shared__ float sred[];                    // [3][C] constants, then [blockDim.x][16] fold area
  float* cst = sred;
  float* fold = sred + 3 * c_store;
  for (int ch = threadIdx.x; ch < c_store; ch += blockDim.x) {
    cst[ch] = mean[ch];
    if (kMask == 2) { cst[c_store + ch] = relu_scale[ch]; cst[2 * c_store + ch] = relu_shift[ch]; }
  }
  __syncthreads();
  const int tpr = blockDim.x / cvec;
  const int cv = threadIdx.x % cvec;
  const int rsub = threadIdx.x / cvec;
  float dg[8], db[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { dg[i] = 0.f; db[i] = 0.f; }
  if (rsub < tpr) {
    const float2* mu2 = reinterpret_cast<const float2*>(cst) + cv * 4;
    const float2* rs2 = reinterpret_cast<const float2*>(cst + c_store) + cv * 4;
    const float2* rh2 = reinterpret_cast<const float2*>(cst + 2 * c_store) + cv * 4;
    const size_t rstep = static_cast<size_t>(gridDim.x) * tpr;
    for (size_t r0 = static_cast<size_t>(blockIdx.x) * tpr + rsub; r0 < rows; r0 += rstep * kUnroll) {
      Words vx[kUnroll], vg[kUnroll], vm[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const size_t r = r0 + u * rstep;
        if (r < rows) {
          vx[u] = ld_words(raw + r * cvec + cv);
          vg[u] = ld_words(dact + r * cvec + cv);
          if (kMask == 1) vm[u] = ld_words(mask + r * cvec + cv);
        } else {
#pragma unroll
          for (int w = 0; w < 4; ++w) { vx[u].w[w] = 0u; vg[u].w[w] = 0u; vm[u].w[w] = 0u; }   // zero gradient: adds nothing
        }
      }
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float2 mu = mu2[w];
        float2 rs = make_float2(0.f, 0.f), rh = make_float2(1.f, 1.f);
        if (kMask == 2) { rs = rs2[w]; rh = rh2[w]; }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const float x0 = lo_f(vx[u].w[w]), x1 = hi_f(vx[u].w[w]);
          float g0 = lo_f(vg[u].w[w]), g1 = hi_f(vg[u].w[w]);
          if (kMask == 1) {
            g0 = lo_f(vm[u].w[w]) > 0.f ? g0 : 0.f;
            g1 = hi_f(vm[u].w[w]) > 0.f ? g1 : 0.f;
          } else if (kMask == 2) {
            g0 = fmaf(x0, rs.x, rh.x) > 0.f ? g0 : 0.f;
            g1 = fmaf(x1, rs.y, rh.y) > 0.f ? g1 : 0.f;
          }
          db[2 * w] += g0; db[2 * w + 1] += g1;
          dg[2 * w] = fmaf(g0, x0 - mu.x, dg[2 * w]);
          dg[2 * w + 1] = fmaf(g1, x1 - mu.y, dg[2 * w + 1]);
        }
      }
    }
  }
  float* mine = fold + threadIdx.x * 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) { mine[i] = dg[i]; mine[8 + i] = db[i]; }
  __syncthreads();
  // fold the tpr row-subsets of each channel vector in a fixed order, then one exact add per channel per CTA
  for (int o = threadIdx.x; o < cvec * 16; o += blockDim.x) {
    const int v = o / 16, k = o % 16;
    float s = 0.f;
    for (int t = 0; t < tpr; ++t) s += fold[(t * cvec + v) * 16 + k];
    const int ch = v * 8 + (k & 7);
    det_add(acc + static_cast<size_t>((k < 8 ? 0 : c_store) + ch) * kDetLimbs, s);
  }
}

// ------------------------------------------------------------------------------------------------ K7b
// draw = gamma*inv_std*(dz - dbeta/M - xhat*dgamma/M); optionally also writes dz (masked dact) for the shortcut path.
// Also publishes sums = [dgamma | dbeta] as plain floats (CTA 0).
template <int kMask, bool kDz>
__global__ void __launch_bounds__(256, 3)
bn_bwd_apply_kernel(const uint4* __restrict__ raw, const uint4* __restrict__ dact, const uint4* __restrict__ mask,
                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                    const float* __restrict__ relu_scale, const float* __restrict__ relu_shift,
                    const unsigned long long* __restrict__ acc, float* __restrict__ sums, uint4* __restrict__ draw,
                    uint4* __restrict__ dz_out, size_t rows, int cvec, int c_store, int c_real, float inv_rows, int acc_raw) {
  fvt_pdl_entry();
  extern __shared__ float cst[];      // [6][C]: a = gamma*inv_std, a*dbeta/M, a*inv_std*dgamma/M, mean, relu scale, relu shift
  for (int ch = threadIdx.x; ch < c_store; ch += blockDim.x) {
    const float is = invstd[ch];
    const float a = (ch < c_real ? gamma[ch] : 0.f) * is;
    // exact sums -> one rounding each: dgamma = inv_std * sum dz*(raw - mean), dbeta = sum dz.  acc_raw: the first sum was
    // accumulated as sum dz*raw by the fused data-gradient epilogue; the mean term is taken out here, in double
    const double s_dz = det_read(acc + static_cast<size_t>(c_store + ch) * kDetLimbs);
    double s_dzx = det_read(acc + static_cast<size_t>(ch) * kDetLimbs);
    if (acc_raw) s_dzx -= static_cast<double>(mean[ch]) * s_dz;
    const float dgamma = static_cast<float>(s_dzx) * is;
    const float dbeta = static_cast<float>(s_dz);
    if (blockIdx.x == 0) { sums[ch] = dgamma; sums[c_store + ch] = dbeta; }
    cst[ch] = a;
    cst[c_store + ch] = a * dbeta * inv_rows;
    cst[2 * c_store + ch] = a * is * dgamma * inv_rows;
    cst[3 * c_store + ch] = mean[ch];
    if (kMask == 2) { cst[4 * c_store + ch] = relu_scale[ch]; cst[5 * c_store + ch] = relu_shift[ch]; }
  }
  __syncthreads();
  const int tpr = blockDim.x / cvec;
  const int cv = threadIdx.x % cvec;
  const int rsub = threadIdx.x / cvec;
  if (rsub >= tpr) return;
  const float2* a2 = reinterpret_cast<const float2*>(cst) + cv * 4;
  const float2* b2 = reinterpret_cast<const float2*>(cst + c_store) + cv * 4;
  const float2* c2 = reinterpret_cast<const float2*>(cst + 2 * c_store) + cv * 4;
  const float2* mu2 = reinterpret_cast<const float2*>(cst + 3 * c_store) + cv * 4;
  const float2* rs2 = reinterpret_cast<const float2*>(cst + 4 * c_store) + cv * 4;
  const float2* rh2 = reinterpret_cast<const float2*>(cst + 5 * c_store) + cv * 4;
  const size_t rstep = static_cast<size_t>(gridDim.x) * tpr;
  for (size_t r0 = static_cast<size_t>(blockIdx.x) * tpr + rsub; r0 < rows; r0 += rstep * kUnroll) {
    Words vx[kUnroll], vg[kUnroll], vm[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t r = r0 + u * rstep;
      if (r < rows) {
        vx[u] = ld_words(raw + r * cvec + cv);
        vg[u] = ld_words(dact + r * cvec + cv);
        if (kMask == 1) vm[u] = ld_words(mask + r * cvec + cv);
      }
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float2 a = a2[w], ab = b2[w], ac = c2[w], mu = mu2[w];
      float2 rs = make_float2(0.f, 0.f), rh = make_float2(1.f, 1.f);
      if (kMask == 2) { rs = rs2[w]; rh = rh2[w]; }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const float x0 = lo_f(vx[u].w[w]), x1 = hi_f(vx[u].w[w]);
        float g0 = lo_f(vg[u].w[w]), g1 = hi_f(vg[u].w[w]);
        if (kMask == 1) {
          g0 = lo_f(vm[u].w[w]) > 0.f ? g0 : 0.f;
          g1 = hi_f(vm[u].w[w]) > 0.f ? g1 : 0.f;
        } else if (kMask == 2) {
          g0 = fmaf(x0, rs.x, rh.x) > 0.f ? g0 : 0.f;
          g1 = fmaf(x1, rs.y, rh.y) > 0.f ? g1 : 0.f;
        }
        // a*(g - dbeta/M - xhat*dgamma/M) with the per-channel products folded into ab, ac
        const float o0 = fmaf(a.x, g0, -ab.x) - (x0 - mu.x) * ac.x;
        const float o1 = fmaf(a.y, g1, -ab.y) - (x1 - mu.y) * ac.y;
        vx[u].w[w] = pack2(o0, o1);
        if (kDz) vg[u].w[w] = pack2(g0, g1);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t r = r0 + u * rstep;
      if (r < rows) {
        const size_t idx = r * cvec + cv;
        draw[idx] = to_uint4(vx[u]);
        if (kDz) dz_out[idx] = to_uint4(vg[u]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ zero insert
// up[n, to*st, ho*sh, wo*sw, :] = dy[n, to, ho, wo, :], zero elsewhere; up has the conv INPUT's spatial extent.
__global__ void __launch_bounds__(256)
zero_insert_kernel(const uint4* __restrict__ dy, uint4* __restrict__ up, int n, int t, int h, int w, int to, int ho,
                   int wo, int st, int sh, int sw, int cvec) {
  fvt_pdl_entry();
  const size_t total = static_cast<size_t>(n) * t * h * w * cvec;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % cvec);
    size_t r = i / cvec;
    const int iw = static_cast<int>(r % w); r /= w;
    const int ih = static_cast<int>(r % h); r /= h;
    const int it = static_cast<int>(r % t);
    const int in = static_cast<int>(r / t);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iw % sw == 0 && ih % sh == 0 && it % st == 0) {
      const int ow = iw / sw, oh = ih / sh, ot = it / st;
      if (ow < wo && oh < ho && ot < to)
        v = __ldg(dy + (((static_cast<size_t>(in) * to + ot) * ho + oh) * wo + ow) * cvec + cv);
    }
    up[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------ head backward
// dlogits [n, K] fp32, pooled [n, C] fp32 (saved by the forward), w [K, C] fp32:
//   dw[k, c] = sum_n dlogits[n,k]*pooled[n,c];  db[k] = sum_n dlogits[n,k];   (overwritten: grad_req='write')
//   dx[n, p, c] = (sum_k dlogits[n,k]*w[k,c]) / positions      (bf16, broadcast over the pooled positions)
__global__ void pool_fc_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ pooled,
                                   const float* __restrict__ w, int n, int num_class, int c, int positions,
                                   float* __restrict__ dw, float* __restrict__ db, __nv_bfloat16* __restrict__ dx,
                                   int c_store) {
  fvt_pdl_entry();
  // grid.x = n (dx part) + num_class (dw/db part)
  if (blockIdx.x < n) {
    const int in = blockIdx.x;
    const float inv = 1.f / static_cast<float>(positions);
    for (int ch = threadIdx.x; ch < c_store; ch += blockDim.x) {
      float s = 0.f;
      if (ch < c)
        for (int k = 0; k < num_class; ++k) s = fmaf(dlogits[in * num_class + k], w[static_cast<size_t>(k) * c + ch], s);
      const __nv_bfloat16 v = __float2bfloat16_rn(s * inv);
      for (int p = 0; p < positions; ++p) dx[(static_cast<size_t>(in) * positions + p) * c_store + ch] = v;
    }
  } else {
    const int k = blockIdx.x - n;
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
      float s = 0.f;
      for (int in = 0; in < n; ++in) s = fmaf(dlogits[in * num_class + k], pooled[static_cast<size_t>(in) * c + ch], s);
      dw[static_cast<size_t>(k) * c + ch] = s;
    }
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int in = 0; in < n; ++in) s += dlogits[in * num_class + k];
      db[k] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------ K12
// MXNet sgd_mom_update on a list of tensors: g' = rescale*g + wd*w; mom = momentum*mom - lr*g'; w += mom
// (gluon.Trainer 'sgd', train_simple_r3d.py:95-97,124).  One launch for the whole list: the descriptor table lives
// in device memory; CTA b processes chunk b of the concatenated index space.
struct SgdTensor { float* w; const float* g; float* mom; unsigned long long numel; float wd; float lr_mult; };

__global__ void __launch_bounds__(256)
sgd_momentum_multi_kernel(const SgdTensor* __restrict__ tensors, const unsigned int* __restrict__ chunk_tensor,
                          const unsigned int* __restrict__ chunk_offset, float lr, float momentum, float rescale,
                          unsigned int chunk_elems) {
  fvt_pdl_entry();
  const SgdTensor t = tensors[chunk_tensor[blockIdx.x]];
  const unsigned long long start = static_cast<unsigned long long>(chunk_offset[blockIdx.x]) * chunk_elems;
  unsigned long long end = start + chunk_elems;
  if (end > t.numel) end = t.numel;
  const float lrt = lr * t.lr_mult;
  // 16-byte path (the flat buffers keep every slot 16-byte aligned and chunks start at multiples of 65536 elements): four
  // elements per thread and access — the scalar loop below ran at 4.8 TB/s over the 1.27 GB the update moves
  if (((reinterpret_cast<uintptr_t>(t.w) | reinterpret_cast<uintptr_t>(t.g) | reinterpret_cast<uintptr_t>(t.mom)) & 15) == 0 && (start & 3) == 0) {
    const unsigned long long n4 = (end - start) >> 2;
    float4* w4 = reinterpret_cast<float4*>(t.w + start);
    const float4* g4 = reinterpret_cast<const float4*>(t.g + start);
    float4* m4 = reinterpret_cast<float4*>(t.mom + start);
    for (unsigned long long i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 w = w4[i];
      const float4 g = g4[i];
      float4 m = m4[i];
      m.x = momentum * m.x - lrt * fmaf(rescale, g.x, t.wd * w.x);
      m.y = momentum * m.y - lrt * fmaf(rescale, g.y, t.wd * w.y);
      m.z = momentum * m.z - lrt * fmaf(rescale, g.z, t.wd * w.z);
      m.w = momentum * m.w - lrt * fmaf(rescale, g.w, t.wd * w.w);
      w.x += m.x; w.y += m.y; w.z += m.z; w.w += m.w;
      m4[i] = m;
      w4[i] = w;
    }
    for (unsigned long long i = start + (n4 << 2) + threadIdx.x; i < end; i += blockDim.x) {
      const float w = t.w[i];
      const float gp = fmaf(rescale, t.g[i], t.wd * w);
      const float m = momentum * t.mom[i] - lrt * gp;
      t.mom[i] = m;
      t.w[i] = w + m;
    }
    return;
  }
  for (unsigned long long i = start + threadIdx.x; i < end; i += blockDim.x) {
    const float w = t.w[i];
    const float gp = fmaf(rescale, t.g[i], t.wd * w);
    const float m = momentum * t.mom[i] - lrt * gp;
    t.mom[i] = m;
    t.w[i] = w + m;
  }
}

// ------------------------------------------------------------------------------------------------ split-K finalize
// ws [splits][rows, C] fp32 holds the partial tiles of a split-K convolution (conv_igemm.cuh), one slice per split.  This
// pass is the convolution's epilogue: y = bf16( relu?( (sum of the slices in split order)*scale + shift + residual ) ),
// optional per-channel (sum, sum^2) of the bf16-rounded raw output for training BatchNorm (exact accumulators).
__global__ void __launch_bounds__(256)
splitk_finalize_kernel(const float4* __restrict__ ws, int splits, size_t slice_vec4, const float* __restrict__ scale,
                       const float* __restrict__ shift, const uint4* __restrict__ res, uint4* __restrict__ y,
                       unsigned long long* __restrict__ stats, size_t rows, int cvec, int c_store, int relu) {
  fvt_pdl_entry();
  extern __shared__ float sred[];                    // [blockDim.x][16] (statistics only)
  const int tpr = blockDim.x / cvec;
  const int cv = threadIdx.x % cvec;
  const int rsub = threadIdx.x / cvec;
  float sc[8], sh[8], s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc[i] = scale != nullptr ? scale[cv * 8 + i] : 1.f;
    sh[i] = shift != nullptr ? shift[cv * 8 + i] : 0.f;
    s1[i] = 0.f; s2[i] = 0.f;
  }
  if (rsub < tpr) {
    for (size_t r = static_cast<size_t>(blockIdx.x) * tpr + rsub; r < rows; r += static_cast<size_t>(gridDim.x) * tpr) {
      const size_t idx = r * cvec + cv;
      float4 a = __ldg(ws + 2 * idx), b = __ldg(ws + 2 * idx + 1);
      for (int k = 1; k < splits; ++k) {
        const float4 a2 = __ldg(ws + k * slice_vec4 + 2 * idx), b2 = __ldg(ws + k * slice_vec4 + 2 * idx + 1);
        a.x += a2.x; a.y += a2.y; a.z += a2.z; a.w += a2.w;
        b.x += b2.x; b.y += b2.y; b.z += b2.z; b.w += b2.w;
      }
      float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      if (stats != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float q = __bfloat162float(__float2bfloat16_rn(v[i]));
          s1[i] += q; s2[i] = fmaf(q, q, s2[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], sc[i], sh[i]);
      if (res != nullptr) {
        float q[8];
        unpack8(__ldg(res + idx), q);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += q[i];
      }
      if (relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      y[idx] = pack8(v);
    }
  }
  if (stats != nullptr) {
    float* mine = sred + threadIdx.x * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) { mine[i] = s1[i]; mine[8 + i] = s2[i]; }
    __syncthreads();
    for (int o = threadIdx.x; o < cvec * 16; o += blockDim.x) {
      const int v = o / 16, k = o % 16;
      float s = 0.f;
      for (int t = 0; t < tpr; ++t) s += sred[(t * cvec + v) * 16 + k];
      det_add(stats + static_cast<size_t>((k < 8 ? 0 : c_store) + v * 8 + (k & 7)) * kDetLimbs, s);
    }
  }
}

// ------------------------------------------------------------------------------------------------ eval-mode BN folding
// scale = gamma / sqrt(var + eps), shift = beta - mean*scale for every BatchNorm of a network in ONE launch (one CTA per
// layer; pad channels get (0, 0) so they stay exactly zero).  Replaces ~5 eager tensor ops per layer each time an
// inference plan is (re)built after the weights moved.
__global__ void __launch_bounds__(256)
bn_fold_multi_kernel(const fvt_bn_fold_entry* __restrict__ table) {
  const fvt_bn_fold_entry e = table[blockIdx.x];
  for (int c = threadIdx.x; c < e.c_store; c += blockDim.x) {
    float sc = 0.f, sh = 0.f;
    if (c < e.c_real) {
      sc = e.gamma[c] / sqrtf(e.var[c] + e.eps);
      sh = e.beta[c] - e.mean[c] * sc;
    }
    e.scale[c] = sc;
    e.shift[c] = sh;
  }
}

// n plain floats <-> n exact accumulators (fvt_stats_encode / fvt_stats_decode)
__global__ void stats_encode_kernel(const float* __restrict__ v, unsigned long long* __restrict__ acc, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < kDetLimbs; ++k) acc[static_cast<size_t>(i) * kDetLimbs + k] = 0ull;
  det_add(acc + static_cast<size_t>(i) * kDetLimbs, v[i]);
}
__global__ void stats_decode_kernel(const unsigned long long* __restrict__ acc, float* __restrict__ v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = static_cast<float>(det_read(acc + static_cast<size_t>(i) * kDetLimbs));
}

static int rows_launch(size_t rows, int cvec, int* blocks, int* threads) {
  // blockDim multiple of cvec (thread -> fixed channel vector), <= 256 threads when possible
  int tpr = 256 / cvec;
  if (tpr < 1) tpr = 1;
  *threads = tpr * cvec;
  if (*threads > 1024) return -1;
  size_t b = (rows + static_cast<size_t>(tpr) * kUnroll - 1) / (static_cast<size_t>(tpr) * kUnroll);
  const size_t cap = 148 * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  *blocks = static_cast<int>(b);
  return 0;
}

int launch_splitk_finalize(const float* ws, int splits, const float* scale, const float* shift, const void* residual, void* y,
                           unsigned long long* stats, size_t rows, int c_store, int relu, cudaStream_t stream, bool pdl) {
  int blocks, threads;
  if (c_store % 8 || rows_launch(rows, c_store / 8, &blocks, &threads)) return set_error(FVT_ERR_BAD_DESC, "split-K finalize: bad channel count");
  // one row per thread per iteration here: rows_launch sized the grid for kUnroll rows per iteration
  size_t b = (rows + (threads / (c_store / 8)) - 1) / (threads / (c_store / 8));
  if (b > 148 * 8) b = 148 * 8;
  if (stats != nullptr) {
    // every CTA ends with 4 integer reductions per channel into the same accumulators: one row per CTA meant 784 x 2304 x 2
    // of them for a conv5_x layer (32 us for a 3.6 MB tensor) — wide layers get fewer CTAs that walk several rows
    size_t by_channels = 300000 / (4 * static_cast<size_t>(c_store));
    if (by_channels < 74) by_channels = 74;
    if (b > by_channels) b = by_channels;
  }
  fvt::launch(splitk_finalize_kernel, static_cast<int>(b), threads, stats ? threads * 16 * sizeof(float) : 0, stream, 1, pdl,
              reinterpret_cast<const float4*>(ws), splits, rows * static_cast<size_t>(c_store) / 4, scale, shift, (const uint4*)residual,
              (uint4*)y, stats, rows, c_store / 8, c_store, relu);
  return check_launch("splitk_finalize_kernel");
}

}  // namespace fvt

using namespace fvt;

static int launch_bn_apply(const void* raw, const float* scale, const float* shift, const void* res, const float* res_scale,
                           const float* res_shift, void* out, int64_t rows, int c_store, int relu, const BnFinalizeArgs& fin,
                           int blocks, int threads, cudaStream_t stream, bool pdl) {
  const int res_mode = res == nullptr ? 0 : (res_scale ? 2 : 1);
  const size_t smem = sizeof(float) * 4 * c_store;
#define FVT_BN_APPLY(M) fvt::launch(bn_apply_kernel<M>, blocks, threads, smem, stream, 1, pdl, \
      (const uint4*)raw, scale, shift, (const uint4*)res, res_scale, res_shift, (uint4*)out, (size_t)rows, c_store / 8, relu, fin)
  if (res_mode == 0) FVT_BN_APPLY(0); else if (res_mode == 1) FVT_BN_APPLY(1); else FVT_BN_APPLY(2);
#undef FVT_BN_APPLY
  return check_launch("bn_apply_kernel");
}

extern "C" {

int fvt_stats_encode(fvt_handle_t handle, const float* values, void* stats_acc, int32_t n, void* stream) {
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  if (!values || !stats_acc || n <= 0) return set_error(FVT_ERR_BAD_DESC, "bad fvt_stats_encode arguments");
  stats_encode_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(values, (unsigned long long*)stats_acc, n);
  return check_launch("stats_encode_kernel");
}

int fvt_stats_decode(fvt_handle_t handle, const void* stats_acc, float* values, int32_t n, void* stream) {
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  if (!values || !stats_acc || n <= 0) return set_error(FVT_ERR_BAD_DESC, "bad fvt_stats_decode arguments");
  stats_decode_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const unsigned long long*)stats_acc, values, n);
  return check_launch("stats_decode_kernel");
}

int fvt_bn_fold_multi(fvt_handle_t handle, const fvt_bn_fold_entry* table_dev, int32_t n_entries, void* stream) {
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  if (table_dev == nullptr || n_entries <= 0) return set_error(FVT_ERR_BAD_DESC, "empty BatchNorm fold table");
  bn_fold_multi_kernel<<<n_entries, 256, 0, (cudaStream_t)stream>>>(table_dev);
  return check_launch("bn_fold_multi_kernel");
}

int fvt_bn_finalize(fvt_handle_t handle, const void* stats_acc, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, int32_t c_store, int32_t c_real, int64_t rows, float eps, float momentum,
                    float* scale, float* shift, float* mean, float* invstd, void* stream) {
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  if (!stats_acc || !gamma || !beta || !scale || !shift || !mean || !invstd) return set_error(FVT_ERR_BAD_DESC, "null pointer");
  if (c_store <= 0 || c_real <= 0 || c_real > c_store || rows <= 0) return set_error(FVT_ERR_BAD_DESC, "bad bn_finalize extent");
  fvt::launch(bn_finalize_kernel, (c_store + 127) / 128, 128, 0, (cudaStream_t)stream, 1, handle_pdl(handle),
              (const unsigned long long*)stats_acc, gamma, beta, running_mean, running_var, c_store, c_real, 1.0 / static_cast<double>(rows), eps,
              momentum, scale, shift, mean, invstd);
  return check_launch("bn_finalize_kernel");
}

int fvt_bn_apply(fvt_handle_t handle, const void* raw, const float* scale, const float* shift, const void* res,
                 const float* res_scale, const float* res_shift, void* out, int64_t rows, int32_t c_store, int32_t relu,
                 void* stream) {
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  if (!raw || !scale || !shift || !out) return set_error(FVT_ERR_BAD_DESC, "null pointer");
  if (c_store <= 0 || c_store % 8 || rows <= 0) return set_error(FVT_ERR_BAD_DESC, "bad bn_apply extent");
  if ((res_scale == nullptr) != (res_shift == nullptr)) return set_error(FVT_ERR_BAD_DESC, "res_scale/res_shift must come together");
  int blocks, threads;
  if (rows_launch(rows, c_store / 8, &blocks, &threads)) return set_error(FVT_ERR_BAD_DESC, "channel count too large");
  BnFinalizeArgs fin;
  memset(&fin, 0, sizeof(fin));
  return launch_bn_apply(raw, scale, shift, res, res_scale, res_shift, out, rows, c_store, relu, fin, blocks, threads,
                         (cudaStream_t)stream, handle_pdl(handle));
}

int fvt_bn_finalize_apply(fvt_handle_t handle, const void* stats_acc, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, int32_t c_store, int32_t c_real, int64_t rows, float eps, float momentum,
                          float* scale, float* shift, float* mean, float* invstd, const void* raw, const void* res,
                          const float* res_scale, const float* res_shift, void* out, int32_t relu, void* stream) {
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  if (!stats_acc || !gamma || !beta || !scale || !shift || !mean || !invstd || !raw || !out) return set_error(FVT_ERR_BAD_DESC, "null pointer");
  if (c_store <= 0 || c_store % 8 || c_real <= 0 || c_real > c_store || rows <= 0) return set_error(FVT_ERR_BAD_DESC, "bad bn_finalize_apply extent");
  if ((res_scale == nullptr) != (res_shift == nullptr)) return set_error(FVT_ERR_BAD_DESC, "res_scale/res_shift must come together");
  if ((running_mean == nullptr) != (running_var == nullptr)) return set_error(FVT_ERR_BAD_DESC, "running_mean/running_var must come together");
  int blocks, threads;
  if (rows_launch(rows, c_store / 8, &blocks, &threads)) return set_error(FVT_ERR_BAD_DESC, "channel count too large");
  BnFinalizeArgs fin;
  fin.stats = (const unsigned long long*)stats_acc; fin.gamma = gamma; fin.beta = beta; fin.running_mean = running_mean; fin.running_var = running_var;
  fin.scale_out = scale; fin.shift_out = shift; fin.mean_out = mean; fin.invstd_out = invstd;
  fin.c_real = c_real; fin.inv_rows = 1.0 / static_cast<double>(rows); fin.eps = eps; fin.momentum = momentum;
  return launch_bn_apply(raw, nullptr, nullptr, res, res_scale, res_shift, out, rows, c_store, relu, fin, blocks, threads,
                         (cudaStream_t)stream, handle_pdl(handle));
}

int fvt_bn_backward(fvt_handle_t handle, const void* raw, const void* dact, const void* mask, const float* mean,
                    const float* invstd, const float* gamma, const float* relu_scale, const float* relu_shift, float* sums,
                    void* sums_acc, void* draw, void* dz_out, int64_t rows, int32_t c_store, int32_t c_real, int32_t dz_in,
                    void* stream) {
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  if ((relu_scale == nullptr) != (relu_shift == nullptr)) return set_error(FVT_ERR_BAD_DESC, "relu_scale/relu_shift must come together");
  if (mask != nullptr && relu_scale != nullptr) return set_error(FVT_ERR_BAD_DESC, "give either a mask tensor or relu_scale/relu_shift");
  if (!raw || !dact || !mean || !invstd || !gamma || !sums || !sums_acc || !draw) return set_error(FVT_ERR_BAD_DESC, "null pointer");
  if (c_store <= 0 || c_store % 8 || rows <= 0 || c_real > c_store) return set_error(FVT_ERR_BAD_DESC, "bad bn_backward extent");
  if (((uintptr_t)sums_acc) & 7) return set_error(FVT_ERR_MISALIGNED, "sums_acc must be 8-byte aligned");
  if (dz_in && (mask != nullptr || dz_out != nullptr)) return set_error(FVT_ERR_BAD_DESC, "dz_in: dact already is dz (no mask tensor, no dz_out)");
  int blocks, threads;
  if (rows_launch(rows, c_store / 8, &blocks, &threads)) return set_error(FVT_ERR_BAD_DESC, "channel count too large");
  unsigned long long* acc = (unsigned long long*)sums_acc;
  int mask_mode = mask != nullptr ? 1 : (relu_scale != nullptr ? 2 : 0);
  if (!dz_in) {
    cudaMemsetAsync(acc, 0, fvt_stats_bytes(c_store), (cudaStream_t)stream);
    const size_t smem_r = sizeof(float) * (3 * c_store + threads * 16);
    // every CTA ends with one exact add per channel and quantity into the SAME 2*C accumulators: keep the grid at what is
    // resident anyway (3 CTAs per SM) — 1184 CTAs queued 1184 same-address reductions per channel at the L2 (+4.5 us per launch)
    // ... and every CTA ends with 4 integer reductions per channel (2 quantities x 2 limbs): 444 CTAs x 576 channels are a
    // million same-address reductions queued at the L2 slices (measured inside a replayed graph: 27 us for the reduce pass
    // of a 7 MB conv4_x tensor, most of it that queue) — wide layers get fewer, longer CTAs (<= ~300k reductions)
    int rblocks = blocks < 148 * 3 ? blocks : 148 * 3;
    const int by_channels = 300000 / (4 * c_store);
    if (rblocks > by_channels) rblocks = by_channels < 74 ? 74 : by_channels;
#define FVT_BN_RED(M) fvt::launch(bn_bwd_reduce_kernel<M>, rblocks, threads, smem_r, (cudaStream_t)stream, 1, handle_pdl(handle), \
      (const uint4*)raw, (const uint4*)dact, (const uint4*)mask, mean, relu_scale, relu_shift, acc, (size_t)rows, c_store / 8, c_store)
    if (mask_mode == 0) FVT_BN_RED(0); else if (mask_mode == 1) FVT_BN_RED(1); else FVT_BN_RED(2);
#undef FVT_BN_RED
    if (int e = check_launch("bn_bwd_reduce_kernel")) return e;
  } else {
    mask_mode = 0;                                   // the producer already applied the ReLU mask
  }
  const size_t smem_a = sizeof(float) * 6 * c_store;
#define FVT_BN_APP(M, D) fvt::launch(bn_bwd_apply_kernel<M, D>, blocks, threads, smem_a, (cudaStream_t)stream, 1, handle_pdl(handle), \
      (const uint4*)raw, (const uint4*)dact, (const uint4*)mask, mean, invstd, gamma, relu_scale, relu_shift, acc, sums, \
      (uint4*)draw, (uint4*)dz_out, (size_t)rows, c_store / 8, c_store, c_real, 1.0f / static_cast<float>(rows), dz_in == 2 ? 1 : 0)
  if (dz_out != nullptr) {
    if (mask_mode == 0) FVT_BN_APP(0, true); else if (mask_mode == 1) FVT_BN_APP(1, true); else FVT_BN_APP(2, true);
  } else {
    if (mask_mode == 0) FVT_BN_APP(0, false); else if (mask_mode == 1) FVT_BN_APP(1, false); else FVT_BN_APP(2, false);
  }
#undef FVT_BN_APP
  return check_launch("bn_bwd_apply_kernel");
}

int fvt_zero_insert(fvt_handle_t handle, const void* dy, void* up, int32_t n, int32_t t, int32_t h, int32_t w, int32_t to, int32_t ho,
                    int32_t wo, int32_t st_, int32_t sh, int32_t sw, int32_t c_store, void* stream) {
  if (!dy || !up) return set_error(FVT_ERR_BAD_DESC, "null pointer");
  if (c_store <= 0 || c_store % 8) return set_error(FVT_ERR_BAD_DESC, "bad channel count");
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  const size_t total = static_cast<size_t>(n) * t * h * w * (c_store / 8);
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  fvt::launch(zero_insert_kernel, static_cast<int>(blocks), 256, 0, (cudaStream_t)stream, 1, handle_pdl(handle), (const uint4*)dy,
              (uint4*)up, n, t, h, w, to, ho, wo, st_, sh, sw, c_store / 8);
  return check_launch("zero_insert_kernel");
}

int fvt_pool_fc_bwd(fvt_handle_t handle, const float* dlogits, const float* pooled, const float* w, int32_t n, int32_t num_class,
                    int32_t c, int32_t positions, float* dw, float* db, void* dx, int32_t c_store, void* stream) {
  if (!dlogits || !pooled || !w || !dw || !db || !dx) return set_error(FVT_ERR_BAD_DESC, "null pointer");
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  fvt::launch(pool_fc_bwd_kernel, n + num_class, 256, 0, (cudaStream_t)stream, 1, handle_pdl(handle), dlogits, pooled, w, n, num_class,
              c, positions, dw, db, (__nv_bfloat16*)dx, c_store);
  return check_launch("pool_fc_bwd_kernel");
}

int fvt_sgd_momentum_multi(fvt_handle_t handle, const void* tensor_table, const uint32_t* chunk_tensor, const uint32_t* chunk_offset,
                           int32_t num_chunks, uint32_t chunk_elems, float lr, float momentum, float rescale,
                           void* stream) {
  if (!tensor_table || !chunk_tensor || !chunk_offset || num_chunks <= 0) return set_error(FVT_ERR_BAD_DESC, "bad sgd table");
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  fvt::launch(sgd_momentum_multi_kernel, num_chunks, 256, 0, (cudaStream_t)stream, 1, handle_pdl(handle), (const SgdTensor*)tensor_table,
              chunk_tensor, chunk_offset, lr, momentum, rescale, chunk_elems);
  return check_launch("sgd_momentum_multi_kernel");
}

}  // extern "C"
