// fp32 path of the hot path (north-star: "logits within rel 1e-4 for the fp32 path"; BASELINE configs[0] is the
// reference's own fp32 case: R(2+1)D-18 forward, batch 2, 8x112x112).
//
// The tcgen05 kernels store activations and operands in bf16 (fp32 accumulate), which bounds end-to-end agreement with
// an fp32 reference at ~1e-2.  This file is the same operator set in plain fp32 on the CUDA cores — Conv3D with the
// folded-BatchNorm / residual / ReLU epilogue, and AvgPool3D + Dense — so that the layer wiring, padding, stride and
// BatchNorm-folding semantics can be checked against the oracle four orders of magnitude tighter than bf16 allows.
// It is a verification-grade path (a few TFLOP/s), not the product's fast path, and it is inference-only.
//
// Layouts: activations NDHWC fp32 (any channel count), weights (kT, kH, kW, I, O) fp32 — output channels innermost, so a
// warp (32 consecutive output channels of one pixel group) reads weights coalesced and activations as broadcasts.
// Replaces nn.Conv3D / nn.BatchNorm (eval) / Activation / add at reference model/R2Plus1.py:27-38,59-62,67-71,81,100-114.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fvt_b200.h"
#include "host_common.h"

namespace fvt {

constexpr int kF32Pix = 8;          // output pixels per thread
constexpr int kF32Warps = 4;        // warps per CTA: 4 pixel groups x 32 output channels

__global__ void __launch_bounds__(32 * kF32Warps)
conv3d_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
                  const float* __restrict__ shift, const float* __restrict__ res, float* __restrict__ y, int n, int t, int h,
                  int wd, int cin, int cout, int kt, int kh, int kw, int st, int sh, int sw, int pt, int ph, int pw, int to,
                  int ho, int wo, int relu) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int co = blockIdx.y * 32 + lane;
  const long long m_total = static_cast<long long>(n) * to * ho * wo;
  const long long m0 = (static_cast<long long>(blockIdx.x) * kF32Warps + warp) * kF32Pix;
  if (m0 >= m_total) return;
  int on[kF32Pix], ot[kF32Pix], oh[kF32Pix], ow[kF32Pix];
  bool live[kF32Pix];
#pragma unroll
  for (int p = 0; p < kF32Pix; ++p) {
    long long m = m0 + p;
    live[p] = m < m_total;
    if (!live[p]) m = m_total - 1;
    ow[p] = static_cast<int>(m % wo); m /= wo;
    oh[p] = static_cast<int>(m % ho); m /= ho;
    ot[p] = static_cast<int>(m % to);
    on[p] = static_cast<int>(m / to);
  }
  float acc[kF32Pix];
#pragma unroll
  for (int p = 0; p < kF32Pix; ++p) acc[p] = 0.f;
  const bool co_ok = co < cout;
  for (int dt = 0; dt < kt; ++dt)
    for (int dh = 0; dh < kh; ++dh)
      for (int dw = 0; dw < kw; ++dw) {
        const float* xp[kF32Pix];
#pragma unroll
        for (int p = 0; p < kF32Pix; ++p) {
          const int it = ot[p] * st - pt + dt, ih = oh[p] * sh - ph + dh, iw = ow[p] * sw - pw + dw;
          const bool ok = live[p] && it >= 0 && it < t && ih >= 0 && ih < h && iw >= 0 && iw < wd;
          xp[p] = ok ? x + (((static_cast<size_t>(on[p]) * t + it) * h + ih) * wd + iw) * cin : nullptr;
        }
        const float* wp = w + static_cast<size_t>((dt * kh + dh) * kw + dw) * cin * cout + co;
        for (int ci = 0; ci < cin; ++ci) {
          const float wv = co_ok ? __ldg(wp + static_cast<size_t>(ci) * cout) : 0.f;
#pragma unroll
          for (int p = 0; p < kF32Pix; ++p)
            if (xp[p] != nullptr) acc[p] = fmaf(__ldg(xp[p] + ci), wv, acc[p]);     // warp-uniform branch and address
        }
      }
  if (!co_ok) return;
  const float sc = scale != nullptr ? scale[co] : 1.f, sf = shift != nullptr ? shift[co] : 0.f;
#pragma unroll
  for (int p = 0; p < kF32Pix; ++p) {
    if (!live[p]) continue;
    const size_t o = static_cast<size_t>(m0 + p) * cout + co;
    float v = fmaf(acc[p], sc, sf);
    if (res != nullptr) v += res[o];
    if (relu) v = fmaxf(v, 0.f);
    y[o] = v;
  }
}

// mean over `positions` pixels then dense; one CTA per clip (fp32 twin of pool_fc_kernel in aux_kernels.cu)
__global__ void pool_fc_f32_kernel(const float* __restrict__ x, int positions, int c, const float* __restrict__ w,
                                   const float* __restrict__ b, int num_class, float* __restrict__ pooled,
                                   float* __restrict__ logits) {
  extern __shared__ float sp[];
  const int n = blockIdx.x;
  const float* xn = x + static_cast<size_t>(n) * positions * c;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < positions; ++p) s += xn[static_cast<size_t>(p) * c + ch];
    s /= static_cast<float>(positions);
    sp[ch] = s;
    if (pooled != nullptr) pooled[static_cast<size_t>(n) * c + ch] = s;
  }
  __syncthreads();
  if (logits == nullptr) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int k = warp; k < num_class; k += nwarps) {
    float s = 0.f;
    for (int ch = lane; ch < c; ch += 32) s = fmaf(sp[ch], __ldg(w + static_cast<size_t>(k) * c + ch), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) logits[static_cast<size_t>(n) * num_class + k] = s + (b ? b[k] : 0.f);
  }
}

}  // namespace fvt

using namespace fvt;

extern "C" {

int fvt_conv3d_fwd_f32(fvt_handle_t handle, const fvt_conv_desc* d, const float* x, const float* w_thwio, const float* scale, const float* shift,
                       const float* residual, float* y, void* stream) {
  if (d == nullptr || x == nullptr || w_thwio == nullptr || y == nullptr) return set_error(FVT_ERR_BAD_DESC, "null pointer");
  if (d->n <= 0 || d->t <= 0 || d->h <= 0 || d->w <= 0 || d->cin <= 0 || d->cout <= 0 || d->kt < 1 || d->kh < 1 || d->kw < 1 ||
      d->st < 1 || d->sh < 1 || d->sw < 1 || d->pt < 0 || d->ph < 0 || d->pw < 0)
    return set_error(FVT_ERR_BAD_DESC, "bad fp32 conv descriptor");
  if ((d->flags & FVT_CONV_RESIDUAL) && residual == nullptr) return set_error(FVT_ERR_BAD_DESC, "FVT_CONV_RESIDUAL without a residual tensor");
  if (d->flags & FVT_CONV_STATS) return set_error(FVT_ERR_BAD_DESC, "the fp32 path is inference-only (no statistics)");
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  const int to = (d->t + 2 * d->pt - d->kt) / d->st + 1, ho = (d->h + 2 * d->ph - d->kh) / d->sh + 1,
            wo = (d->w + 2 * d->pw - d->kw) / d->sw + 1;
  if (to <= 0 || ho <= 0 || wo <= 0) return set_error(FVT_ERR_BAD_DESC, "filter larger than padded input");
  const long long m_total = static_cast<long long>(d->n) * to * ho * wo;
  const long long per_cta = static_cast<long long>(kF32Warps) * kF32Pix;
  dim3 grid(static_cast<unsigned>((m_total + per_cta - 1) / per_cta), static_cast<unsigned>((d->cout + 31) / 32));
  conv3d_f32_kernel<<<grid, 32 * kF32Warps, 0, (cudaStream_t)stream>>>(
      x, w_thwio, scale, shift, (d->flags & FVT_CONV_RESIDUAL) ? residual : nullptr, y, d->n, d->t, d->h, d->w, d->cin, d->cout,
      d->kt, d->kh, d->kw, d->st, d->sh, d->sw, d->pt, d->ph, d->pw, to, ho, wo, (d->flags & FVT_CONV_RELU) ? 1 : 0);
  return check_launch("conv3d_f32_kernel");
}

int fvt_pool_fc_fwd_f32(fvt_handle_t handle, const float* x, int32_t n, int32_t positions, int32_t c, const float* w, const float* b,
                        int32_t num_class, float* pooled, float* logits, void* stream) {
  if (x == nullptr) return set_error(FVT_ERR_BAD_DESC, "null tensor pointer");
  if (n <= 0 || positions <= 0 || c <= 0) return set_error(FVT_ERR_BAD_DESC, "bad pool/fc extent");
  if (logits != nullptr && (w == nullptr || num_class <= 0)) return set_error(FVT_ERR_BAD_DESC, "logits requested without weights");
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  pool_fc_f32_kernel<<<n, 512, c * sizeof(float), (cudaStream_t)stream>>>(x, positions, c, w, b, num_class, pooled, logits);
  return check_launch("pool_fc_f32_kernel");
}

}  // extern "C"
