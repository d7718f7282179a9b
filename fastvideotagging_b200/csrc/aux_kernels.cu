// HBM-bound helpers around K1: stem input unfold, global-average-pool + dense head.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/fvt_b200.h"
#include "host_common.h"
#include "pdl.cuh"

namespace fvt {

// ---------------------------------------------------------------------------------------------------------
// Stem unfold.  One thread per output pixel (n, t, h, ow): gathers kw_taps x 3 fp32 values along W from the three
// channel planes of the NCDHW clip and writes `cu` bf16 channels (64 B for cu = 32) in one go.
// Reads are coalesced across the warp (adjacent ow -> addresses sw apart, overlapping taps hit L1);
// writes are fully coalesced (cu*2 contiguous bytes per thread, consecutive threads consecutive pixels).
// ---------------------------------------------------------------------------------------------------------
// hpair != 0: the unfolded rows 2*h2 and 2*h2+1 are interleaved per pixel, u2[n,t,h2,ow, (h&1)*CU + k] = u[n,t,h,ow,k]
// (h even), so the stride-2 walk of the stem conv over H becomes a stride-1 walk over h2 with 2*CU channels.
template <int CU>
__global__ void stem_unfold_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ u, int n, int t, int h,
                                   int w, int wo, int kw_taps, int sw, int pw, int hpair) {
  fvt_pdl_entry();
  const size_t total = static_cast<size_t>(n) * t * h * wo;
  const size_t plane = static_cast<size_t>(h) * w;           // one (n, c, t) frame
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ow = static_cast<int>(i % wo);
    size_t r = i / wo;
    const int ih = static_cast<int>(r % h);  r /= h;
    const int it = static_cast<int>(r % t);
    const int in = static_cast<int>(r / t);
    const int w0 = ow * sw - pw;
    __align__(16) __nv_bfloat16 vals[CU];
#pragma unroll
    for (int k = 0; k < CU; ++k) vals[k] = __float2bfloat16_rn(0.f);
    const float* base = x + ((static_cast<size_t>(in) * 3) * t + it) * plane + static_cast<size_t>(ih) * w;
    for (int k = 0; k < kw_taps; ++k) {
      const int iw = w0 + k;
      if (iw >= 0 && iw < w) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          vals[k * 3 + c] = __float2bfloat16_rn(__ldg(base + static_cast<size_t>(c) * t * plane + iw));
        }
      }
    }
    size_t o = i;
    if (hpair) o = ((((static_cast<size_t>(in) * t + it) * (h >> 1) + (ih >> 1)) * wo + ow) << 1) + (ih & 1);
    uint4* dst = reinterpret_cast<uint4*>(u + o * CU);
    const uint4* src = reinterpret_cast<const uint4*>(vals);
#pragma unroll
    for (int k = 0; k < CU / 8; ++k) dst[k] = src[k];
  }
}

// ---------------------------------------------------------------------------------------------------------
// Head: mean over `positions` pixels then dense.  One CTA per clip; thread c reduces channel c (coalesced:
// consecutive threads read consecutive channels of one pixel), then each warp computes classes by a
// shuffle-reduced dot product over the pooled vector in shared memory.
// ---------------------------------------------------------------------------------------------------------
__global__ void pool_fc_kernel(const __nv_bfloat16* __restrict__ x, int positions, int c, int c_real,
                               const float* __restrict__ w, const float* __restrict__ b, int num_class,
                               float* __restrict__ pooled, float* __restrict__ logits) {
  fvt_pdl_entry();
  extern __shared__ float sp[];   // [c]
  const int n = blockIdx.x;
  const __nv_bfloat16* xn = x + static_cast<size_t>(n) * positions * c;
  const float inv = 1.f / static_cast<float>(positions);
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < positions; ++p) s += __bfloat162float(xn[static_cast<size_t>(p) * c + ch]);
    s *= inv;
    sp[ch] = s;
    if (pooled != nullptr && ch < c_real) pooled[static_cast<size_t>(n) * c_real + ch] = s;
  }
  __syncthreads();
  if (logits == nullptr) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int k = warp; k < num_class; k += nwarps) {
    const float* wk = w + static_cast<size_t>(k) * c_real;
    float s = 0.f;
    for (int ch = lane; ch < c_real; ch += 32) s = fmaf(sp[ch], __ldg(wk + ch), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) logits[static_cast<size_t>(n) * num_class + k] = s + (b ? b[k] : 0.f);
  }
}

}  // namespace fvt

using namespace fvt;

extern "C" {

static int stem_unfold_launch(fvt_handle_t handle, const float* x_ncdhw, void* u, int32_t n, int32_t t, int32_t h, int32_t w, int32_t kw_taps,
                              int32_t sw, int32_t pw, int32_t cu, int hpair, void* stream) {
  if (x_ncdhw == nullptr || u == nullptr) return set_error(FVT_ERR_BAD_DESC, "null tensor pointer");
  if (n <= 0 || t <= 0 || h <= 0 || w <= 0 || kw_taps <= 0 || sw <= 0 || pw < 0)
    return set_error(FVT_ERR_BAD_DESC, "bad stem unfold extent");
  if (cu != 32 || kw_taps * 3 > cu) return set_error(FVT_ERR_BAD_DESC, "stem unfold supports cu=32 with 3*kw_taps <= 32");
  if (hpair && (h & 1)) return set_error(FVT_ERR_BAD_DESC, "row-paired stem unfold needs an even height (got %d)", h);
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  const int wo = (w + 2 * pw - kw_taps) / sw + 1;
  const size_t total = static_cast<size_t>(n) * t * h * wo;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  fvt::launch(stem_unfold_kernel<32>, static_cast<int>(blocks), 256, 0, (cudaStream_t)stream, 1, handle_pdl(handle),
              x_ncdhw, (__nv_bfloat16*)u, n, t, h, w, wo, kw_taps, sw, pw, hpair);
  return check_launch("stem_unfold_kernel");
}

int fvt_stem_unfold(fvt_handle_t handle, const float* x_ncdhw, void* u, int32_t n, int32_t t, int32_t h, int32_t w, int32_t kw_taps,
                    int32_t sw, int32_t pw, int32_t cu, void* stream) {
  return stem_unfold_launch(handle, x_ncdhw, u, n, t, h, w, kw_taps, sw, pw, cu, 0, stream);
}

int fvt_stem_unfold_hpair(fvt_handle_t handle, const float* x_ncdhw, void* u, int32_t n, int32_t t, int32_t h, int32_t w, int32_t kw_taps,
                          int32_t sw, int32_t pw, int32_t cu, void* stream) {
  return stem_unfold_launch(handle, x_ncdhw, u, n, t, h, w, kw_taps, sw, pw, cu, 1, stream);
}

int fvt_pool_fc_fwd(fvt_handle_t handle, const void* x, int32_t n, int32_t positions, int32_t c, int32_t c_real, const float* w,
                    const float* b, int32_t num_class, float* pooled, float* logits, void* stream) {
  if (x == nullptr) return set_error(FVT_ERR_BAD_DESC, "null tensor pointer");
  if (n <= 0 || positions <= 0 || c <= 0 || c_real <= 0 || c_real > c) return set_error(FVT_ERR_BAD_DESC, "bad pool/fc extent");
  if (logits != nullptr && (w == nullptr || num_class <= 0)) return set_error(FVT_ERR_BAD_DESC, "logits requested without weights");
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  fvt::launch(pool_fc_kernel, n, 512, c * sizeof(float), (cudaStream_t)stream, 1, handle_pdl(handle), (const __nv_bfloat16*)x, positions,
              c, c_real, w, b, num_class, pooled, logits);
  return check_launch("pool_fc_kernel");
}

}  // extern "C"
