// Rows "next to" the hot path: the step just before it (clip pre-processing) and the step just after
// it (evaluation tail), kept on the device so that neither costs a host pass or a per-batch synchronisation.
//
//   N2  clip_stats_u8 / clip_normalize_u8   decoded uint8 frames (N,T,H,W,3) -> per-batch per-channel mean / std ->
//       (x - mean) / (std + 1e-3) in the reference's NCDHW fp32 layout, optional horizontal flip per clip
//       (videos_reader.py:69-76,93-97; fixed ImageNet statistics variant data/ucf101.py:124-128)
//   N3  softmax_accumulate / argmax_correct  multi-clip softmax averaging + accuracy (validation.py:39-66)
//       topk_iou                             top-k (k <= 4) intersection / union counts (train_simple_r3d.py:169-197)
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/fvt_b200.h"
#include "host_common.h"

namespace fvt {

// ------------------------------------------------------------------------------------------------ N2
// sums[c] = sum x, sums[3 + c] = sum x^2 over all pixels of channel c (exact integer arithmetic)
__global__ void __launch_bounds__(256)
clip_stats_u8_kernel(const uint8_t* __restrict__ clips, size_t pixels, unsigned long long* __restrict__ sums) {
  unsigned long long s[3] = {0, 0, 0}, q[3] = {0, 0, 0};
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < pixels;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const unsigned v = clips[3 * i + c];
      s[c] += v; q[c] += v * v;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    for (int o = 16; o > 0; o >>= 1) {
      s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
      q[c] += __shfl_xor_sync(0xffffffffu, q[c], o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(sums + c, s[c]);
      atomicAdd(sums + 3 + c, q[c]);
    }
  }
}

// One normalised pixel value, the same two roundings wherever it is computed: fma(v, scale, -mean), then * inv_std.
__device__ __forceinline__ float norm_px(unsigned v, float scale, float m, float inv) {
  return __fmul_rn(__fmaf_rn(static_cast<float>(v), scale, -m), inv);
}

// out[n, c, t, h, w] = (in[n, t, h, w', c] * scale - mean[c]) * inv_std[c],  w' = flip[n] ? W-1-w : w
__global__ void __launch_bounds__(256)
clip_normalize_u8_kernel(const uint8_t* __restrict__ clips, const uint8_t* __restrict__ flip, float* __restrict__ out, int n,
                         int t, int h, int w, float scale, float m0, float m1, float m2, float i0, float i1, float i2) {
  const size_t plane = static_cast<size_t>(t) * h * w;
  const size_t total = static_cast<size_t>(n) * plane;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t in_ = i / plane;
    const size_t r = i - in_ * plane;              // (t, h, w) of the OUTPUT pixel
    const int ow = static_cast<int>(r % w);
    const size_t row = r / w;
    const int sw = (flip != nullptr && flip[in_]) ? w - 1 - ow : ow;
    const uint8_t* src = clips + ((in_ * plane + row * w + sw) * 3);
    float* dst = out + in_ * 3 * plane + r;
    dst[0] = norm_px(src[0], scale, m0, i0);
    dst[plane] = norm_px(src[1], scale, m1, i1);
    dst[2 * plane] = norm_px(src[2], scale, m2, i2);
  }
}

// N2 as SURVEY 8f specifies it: decoded uint8 frames -> (crop, flip, per-channel normalise) -> the W-unfolded NDHWC bf16
// STEM INPUT, in one HBM-bound pass (no fp32 NCDHW tensor in between; the host ships 1 byte per value instead of 4):
//   u[n, t, h, ow, kw*3 + c] = norm(frame[n, t, y0 + h, x0 + w'(ow*sw - pw + kw), c])   (zero outside the crop)
// with w'(j) = flip[n] ? W-1-j : j; hpair != 0 interleaves the rows 2*h2, 2*h2+1 per pixel exactly like
// stem_unfold_kernel (aux_kernels.cu), whose output this reproduces bit for bit.  One thread per output pixel: 21 byte
// reads (neighbouring threads overlap: L1), one 64-byte store.
template <int CU>
__global__ void __launch_bounds__(256)
clip_unfold_u8_kernel(const uint8_t* __restrict__ clips, const uint8_t* __restrict__ flip, const int* __restrict__ crop_yx,
                      __nv_bfloat16* __restrict__ u, int n, int t, int hs, int ws, int h, int w, int wo, int kw_taps, int sw,
                      int pw, int hpair, float scale, float m0, float m1, float m2, float i0, float i1, float i2) {
  const size_t total = static_cast<size_t>(n) * t * h * wo;
  const float mm[3] = {m0, m1, m2}, ii[3] = {i0, i1, i2};
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ow = static_cast<int>(i % wo);
    size_t r = i / wo;
    const int ih = static_cast<int>(r % h);  r /= h;
    const int it = static_cast<int>(r % t);
    const int in = static_cast<int>(r / t);
    const int y0 = crop_yx != nullptr ? crop_yx[2 * in] : 0, x0 = crop_yx != nullptr ? crop_yx[2 * in + 1] : 0;
    const bool fl = flip != nullptr && flip[in] != 0;
    const uint8_t* row = clips + ((static_cast<size_t>(in) * t + it) * hs + (y0 + ih)) * static_cast<size_t>(ws) * 3;
    __align__(16) __nv_bfloat16 vals[CU];
#pragma unroll
    for (int k = 0; k < CU; ++k) vals[k] = __float2bfloat16_rn(0.f);
    const int w0 = ow * sw - pw;
    for (int k = 0; k < kw_taps; ++k) {
      const int iw = w0 + k;
      if (iw >= 0 && iw < w) {
        const uint8_t* px = row + static_cast<size_t>(x0 + (fl ? w - 1 - iw : iw)) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) vals[k * 3 + c] = __float2bfloat16_rn(norm_px(px[c], scale, mm[c], ii[c]));
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

// ------------------------------------------------------------------------------------------------ N3
// acc[row, :] += softmax(logits[row, :])   (one warp per row)
__global__ void __launch_bounds__(256)
softmax_accumulate_kernel(const float* __restrict__ logits, float* __restrict__ acc, int rows, int c) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = logits + static_cast<size_t>(row) * c;
  float m = -INFINITY;
  for (int j = lane; j < c; j += 32) m = fmaxf(m, x[j]);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int j = lane; j < c; j += 32) s += expf(x[j] - m);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv = 1.f / s;
  for (int j = lane; j < c; j += 32) acc[static_cast<size_t>(row) * c + j] += expf(x[j] - m) * inv;
}

// correct += [argmax_j acc[row, j] == label[row]]  (first maximum wins, like np.argmax)
__global__ void __launch_bounds__(256)
argmax_correct_kernel(const float* __restrict__ acc, const int* __restrict__ labels, int rows, int c, int* __restrict__ pred,
                      unsigned long long* __restrict__ correct) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = acc + static_cast<size_t>(row) * c;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = lane; j < c; j += 32) {
    const float v = x[j];
    if (v > best) { best = v; bi = j; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) {
    if (pred != nullptr) pred[row] = bi;
    if (labels != nullptr && bi == labels[row]) atomicAdd(correct, 1ull);
  }
}

// Per row: indices of the k_max largest scores in the order of `argsort()[:, ::-1]` (descending; among equal scores the
// LARGER index first, because the reference reverses an ascending stable sort); label set = {j : target > 0.1};
// inter[k-1] += |top_k ∩ labels|, uni[k-1] += |top_k ∪ labels| for k = 1..k_max.
__global__ void __launch_bounds__(256)
topk_iou_kernel(const float* __restrict__ scores, const float* __restrict__ target, int rows, int c, int k_max,
                unsigned long long* __restrict__ inter, unsigned long long* __restrict__ uni) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = scores + static_cast<size_t>(row) * c;
  const float* y = target + static_cast<size_t>(row) * c;
  int n_lab = 0;
  for (int j = lane; j < c; j += 32) n_lab += y[j] > 0.1f ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) n_lab += __shfl_xor_sync(0xffffffffu, n_lab, o);
  int chosen[4] = {-1, -1, -1, -1};
  int hits = 0;
  for (int k = 0; k < k_max; ++k) {
    float best = -INFINITY;
    int bi = -1;
    for (int j = lane; j < c; j += 32) {
      bool taken = false;
      for (int q = 0; q < k; ++q) taken |= chosen[q] == j;
      const float v = x[j];
      if (!taken && (v > best || (v == best && j > bi) || bi < 0)) { best = v; bi = j; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi >= 0 && (bi < 0 || ov > best || (ov == best && oi > bi))) { best = ov; bi = oi; }
    }
    chosen[k] = bi;
    if (bi >= 0 && y[bi] > 0.1f) ++hits;
    if (lane == 0) {
      const int kk = k + 1 < c ? k + 1 : c;           // fewer than k classes: the top-k set is the whole row
      atomicAdd(inter + k, static_cast<unsigned long long>(hits));
      atomicAdd(uni + k, static_cast<unsigned long long>(kk + n_lab - hits));
    }
  }
}

}  // namespace fvt

using namespace fvt;

extern "C" {

int fvt_clip_stats_u8(fvt_handle_t handle, const uint8_t* clips_nthwc, int64_t pixels, uint64_t* sums6, void* stream) {
  if (!clips_nthwc || !sums6 || pixels <= 0) return set_error(FVT_ERR_BAD_DESC, "bad clip_stats arguments");
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  cudaMemsetAsync(sums6, 0, 6 * sizeof(uint64_t), (cudaStream_t)stream);
  size_t blocks = (static_cast<size_t>(pixels) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  clip_stats_u8_kernel<<<static_cast<int>(blocks), 256, 0, (cudaStream_t)stream>>>(clips_nthwc, static_cast<size_t>(pixels),
                                                                                   reinterpret_cast<unsigned long long*>(sums6));
  return check_launch("clip_stats_u8_kernel");
}

int fvt_clip_normalize_u8(fvt_handle_t handle, const uint8_t* clips_nthwc, const uint8_t* flip, float* out_ncdhw, int32_t n, int32_t t, int32_t h,
                          int32_t w, float scale, const float mean[3], const float inv_std[3], void* stream) {
  if (!clips_nthwc || !out_ncdhw || !mean || !inv_std || n <= 0 || t <= 0 || h <= 0 || w <= 0)
    return set_error(FVT_ERR_BAD_DESC, "bad clip_normalize arguments");
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  const size_t total = static_cast<size_t>(n) * t * h * w;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  clip_normalize_u8_kernel<<<static_cast<int>(blocks), 256, 0, (cudaStream_t)stream>>>(
      clips_nthwc, flip, out_ncdhw, n, t, h, w, scale, mean[0], mean[1], mean[2], inv_std[0], inv_std[1], inv_std[2]);
  return check_launch("clip_normalize_u8_kernel");
}

int fvt_clip_unfold_u8(fvt_handle_t handle, const uint8_t* clips_nthwc, const uint8_t* flip, const int32_t* crop_yx, void* u, int32_t n,
                       int32_t t, int32_t hs, int32_t ws, int32_t h, int32_t w, float scale, const float mean[3],
                       const float inv_std[3], int32_t kw_taps, int32_t sw, int32_t pw, int32_t cu, int32_t hpair, void* stream) {
  if (!clips_nthwc || !u || !mean || !inv_std || n <= 0 || t <= 0 || h <= 0 || w <= 0 || hs < h || ws < w)
    return set_error(FVT_ERR_BAD_DESC, "bad clip_unfold arguments (frames %dx%d, crop %dx%d)", hs, ws, h, w);
  if (cu != 32 || kw_taps <= 0 || kw_taps * 3 > cu || sw <= 0 || pw < 0) return set_error(FVT_ERR_BAD_DESC, "clip unfold supports cu=32 with 3*kw_taps <= 32");
  if (hpair && (h & 1)) return set_error(FVT_ERR_BAD_DESC, "row-paired unfold needs an even height (got %d)", h);
  if (crop_yx == nullptr && (hs != h || ws != w)) return set_error(FVT_ERR_BAD_DESC, "frames larger than the crop need crop offsets");
  if (((uintptr_t)u) & 15) return set_error(FVT_ERR_MISALIGNED, "u must be 16-byte aligned");
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  const int wo = (w + 2 * pw - kw_taps) / sw + 1;
  const size_t total = static_cast<size_t>(n) * t * h * wo;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  clip_unfold_u8_kernel<32><<<static_cast<int>(blocks), 256, 0, (cudaStream_t)stream>>>(
      clips_nthwc, flip, crop_yx, (__nv_bfloat16*)u, n, t, hs, ws, h, w, wo, kw_taps, sw, pw, hpair, scale, mean[0], mean[1], mean[2],
      inv_std[0], inv_std[1], inv_std[2]);
  return check_launch("clip_unfold_u8_kernel");
}

int fvt_softmax_accumulate(fvt_handle_t handle, const float* logits, float* acc, int32_t rows, int32_t num_class, void* stream) {
  if (!logits || !acc || rows <= 0 || num_class <= 0) return set_error(FVT_ERR_BAD_DESC, "bad softmax_accumulate arguments");
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  softmax_accumulate_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(logits, acc, rows, num_class);
  return check_launch("softmax_accumulate_kernel");
}

int fvt_argmax_correct(fvt_handle_t handle, const float* acc, const int32_t* labels, int32_t rows, int32_t num_class, int32_t* pred,
                       uint64_t* correct, void* stream) {
  if (!acc || rows <= 0 || num_class <= 0 || (labels && !correct)) return set_error(FVT_ERR_BAD_DESC, "bad argmax_correct arguments");
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  argmax_correct_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(acc, labels, rows, num_class, pred,
                                                                          reinterpret_cast<unsigned long long*>(correct));
  return check_launch("argmax_correct_kernel");
}

int fvt_topk_iou(fvt_handle_t handle, const float* scores, const float* target, int32_t rows, int32_t num_class, int32_t k_max, uint64_t* inter,
                 uint64_t* uni, void* stream) {
  if (!scores || !target || !inter || !uni || rows <= 0 || num_class <= 0 || k_max < 1 || k_max > 4)
    return set_error(FVT_ERR_BAD_DESC, "bad topk_iou arguments (k_max in [1, 4])");
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  topk_iou_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(scores, target, rows, num_class, k_max,
                                                                    reinterpret_cast<unsigned long long*>(inter),
                                                                    reinterpret_cast<unsigned long long*>(uni));
  return check_launch("topk_iou_kernel");
}

}  // extern "C"
