// K9-K11 + softmax head: multi-label ranking losses of reference model/mlc_loss.py and the gluon losses the
// training scripts pick (train_simple_r3d.py:43,70-78), forward + backward in ONE launch each.
//
// All of them are (batch x num_class) problems with num_class <= 1024 (63 / 101 in the reference), so the shape is:
// one CTA per sample row, the row's scores/labels staged in shared memory, each thread owning one class k and
// looping over the C partners of the C x C pairwise term (shared-memory tiled, no atomics, fixed summation order),
// a warp-shuffle + shared-memory block reduction for the row scalar, and a "last CTA" tail (atomic ticket) that
// folds the per-row partials in row order — deterministic — and applies the batch-coupled normalisation.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fvt_b200.h"
#include "host_common.h"
#include "philox.cuh"

namespace fvt {

constexpr int kLossThreads = 256;
constexpr int kMaxClass = 1024;

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  t = red[0];
  __syncthreads();
  return t;
}

// Returns true in exactly one CTA: the last one to arrive.  `ticket` must be zero on entry; it is reset on exit.
__device__ __forceinline__ bool last_block(unsigned int* ticket) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// ------------------------------------------------------------------------------------------------ LSEP
// mode 0: LsepLoss.forward (mlc_loss.py:63-86)  loss = log(1 + sum_b sum_{i in pos_b, j in neg_b} exp(p_bj - p_bi));
//         gradient = autodiff of that expression.
// mode 1: LSEP_funcLoss exactly as written (mlc_loss.py:8-54), including its two quirks:
//         forward indexes the score ROW by the enumerate counter of the positive (line 27-29), and backward is
//         fac * sum_pairs (phot - nhot) * exp(-pred*(phot - nhot)) with fac = -1/loss (lines 36, 51-53).
__global__ void __launch_bounds__(kLossThreads)
lsep_kernel(const float* __restrict__ pred, const float* __restrict__ target, int batch, int C, int mode,
            float* __restrict__ loss, float* __restrict__ grad, float* __restrict__ partial, unsigned int* ticket) {
  __shared__ float sp[kMaxClass];
  __shared__ float st[kMaxClass];
  __shared__ int pos_list[kMaxClass];
  __shared__ float red[32];
  __shared__ int npos_s, nneg_s;
  const int b = blockIdx.x;
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    sp[k] = pred[static_cast<size_t>(b) * C + k];
    st[k] = target[static_cast<size_t>(b) * C + k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int np = 0, nn = 0;
    for (int k = 0; k < C; ++k) {
      if (st[k] > 0.f) pos_list[np++] = k;
      if (mode == 0 ? (st[k] == 0.f) : (st[k] <= 0.f)) ++nn;
    }
    npos_s = np; nneg_s = nn;
  }
  __syncthreads();
  float s_local = 0.f;
  if (mode == 0) {
    for (int k = threadIdx.x; k < C; k += blockDim.x) {
      const float pk = sp[k];
      float g = 0.f;
      if (st[k] > 0.f) {
        for (int j = 0; j < C; ++j)
          if (st[j] == 0.f) g -= expf(sp[j] - pk);
        s_local -= g;                         // each (i=k, j) pair counted once, from the positive side
      } else if (st[k] == 0.f) {
        for (int i = 0; i < C; ++i)
          if (st[i] > 0.f) g += expf(pk - sp[i]);
      }
      grad[static_cast<size_t>(b) * C + k] = g;    // un-normalised; scaled by 1/(1+S) in the tail
    }
  } else {
    // as-written forward: for (q, pj) in enumerate(pos): for nj in neg: exp(pred[q, nj] - pred[q, pj])
    const int np = npos_s;
    for (int q = threadIdx.x; q < np; q += blockDim.x) {
      if (q < batch) {
        const int pj = pos_list[q];
        const float ppj = pred[static_cast<size_t>(q) * C + pj];
        for (int nj = 0; nj < C; ++nj)
          if (st[nj] <= 0.f) s_local += expf(pred[static_cast<size_t>(q) * C + nj] - ppj);
      }
    }
    // as-written backward: k positive: +#neg * exp(-p_k); k negative: -#pos * exp(p_k)  (times fac in the tail)
    for (int k = threadIdx.x; k < C; k += blockDim.x) {
      float g = 0.f;
      if (st[k] > 0.f) g += static_cast<float>(nneg_s) * expf(-sp[k]);
      if (st[k] <= 0.f) g -= static_cast<float>(npos_s) * expf(sp[k]);
      grad[static_cast<size_t>(b) * C + k] = g;
    }
  }
  const float s_row = block_sum(s_local, red);
  if (threadIdx.x == 0) partial[b] = s_row;
  if (last_block(ticket)) {
    float S = 0.f;
    if (threadIdx.x == 0) {
      for (int r = 0; r < batch; ++r) S += partial[r];     // fixed order
      red[0] = S;
    }
    __syncthreads();
    S = red[0];
    const float L = logf(1.f + S);
    const float factor = (mode == 0) ? 1.f / (1.f + S) : -1.f / L;
    const size_t total = static_cast<size_t>(batch) * C;
    for (size_t i = threadIdx.x; i < total; i += blockDim.x) grad[i] *= factor;
    if (threadIdx.x == 0) { loss[0] = L; *ticket = 0u; }
  }
}

// ------------------------------------------------------------------------------------------------ WARP
// Sampling (mlc_loss.py:129-147 / :198-215): for every (b, j) with target == 1 draw uniform negatives of row b until
// one violates (p_neg - p_j >= 0) or `max_trials` draws were made; r_j = floor(max_trials / num_trials);
// L[b,j] = rank_weights[r_j] with rank_weights[k] = H_{k+1} (:117-119).  The reference draws with
// np.random.choice (MT19937, not reproducible on a device); the contract here is the counter-based stream
//   u = philox4x32_10(key = seed, counter = (global_sample_index, j, trial, 0)).x ;  neg = negatives[u % n_neg]
// with trial = 1, 2, ... and `negatives` the ascending list of classes with target == 0, shared bit-for-bit with
// oracle/mlc_loss.py.
// mode 0: WarpLoss   (:151-174)  loss = sum_b sum_i L_bi * sum_j relu(1 + pos_i*neg_j*(p_j - p_i))
// mode 1: WARP_funcLoss (:217-232) loss = sum_b (sum_j L_bj) * sum_c (1 - pos_c p_c + neg_c p_c);
//                                  grad = (sum_j L_bj) * (neg - pos)
__global__ void __launch_bounds__(kLossThreads)
warp_kernel(const float* __restrict__ pred, const float* __restrict__ target, int batch, int C, int label_size,
            int max_trials, int mode, unsigned long long seed, unsigned long long sample_offset,
            const float* __restrict__ L_in, float* __restrict__ L_out, int* __restrict__ trials_out,
            float* __restrict__ loss, float* __restrict__ grad, float* __restrict__ partial, unsigned int* ticket) {
  __shared__ float sp[kMaxClass];
  __shared__ float st[kMaxClass];
  __shared__ float sL[kMaxClass];
  __shared__ int neg_list[kMaxClass];
  __shared__ float red[32];
  __shared__ int nneg_s;
  const int b = blockIdx.x;
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    sp[k] = pred[static_cast<size_t>(b) * C + k];
    st[k] = target[static_cast<size_t>(b) * C + k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int nn = 0;
    for (int k = 0; k < C; ++k)
      if (st[k] == 0.f) neg_list[nn++] = k;
    nneg_s = nn;
  }
  __syncthreads();
  const int nneg = nneg_s;
  for (int j = threadIdx.x; j < C; j += blockDim.x) {
    float Lj = 0.f;
    int trials = 0;
    if (L_in != nullptr) {
      Lj = L_in[static_cast<size_t>(b) * C + j];
    } else if (st[j] == 1.f && nneg == 0) {
      Lj = __int_as_float(0x7fc00000);        // the reference never terminates on a row without negatives: poison
    } else if (st[j] == 1.f) {
      const float pj = sp[j];
      float margin = -1.f;
      while (margin < 0.f && trials < max_trials) {
        ++trials;
        const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>((sample_offset + b) & 0xffffffffull),
                                                 static_cast<uint32_t>(j), static_cast<uint32_t>(trials), 0u),
                                      make_uint2(static_cast<uint32_t>(seed & 0xffffffffull),
                                                 static_cast<uint32_t>(seed >> 32)));
        margin = sp[neg_list[r.x % static_cast<uint32_t>(nneg)]] - pj;
      }
      // rank_weights[r] = H_{r+1}; r = floor(max_trials / trials) <= label_size - 1 is guaranteed by the host check
      const int r_j = max_trials / trials;
      double h = 0.0;                         // python-float accumulation of the reference, then fp32 storage
      for (int i = 1; i <= r_j + 1; ++i) h += 1.0 / static_cast<double>(i);
      Lj = static_cast<float>(h);
    }
    sL[j] = Lj;
    if (L_out != nullptr) L_out[static_cast<size_t>(b) * C + j] = Lj;
    if (trials_out != nullptr) trials_out[static_cast<size_t>(b) * C + j] = trials;
  }
  __syncthreads();
  float l_local = 0.f;
  if (mode == 0) {
    for (int k = threadIdx.x; k < C; k += blockDim.x) {
      const float pk = sp[k];
      float g = 0.f;
      const bool kpos = st[k] > 0.f, kneg = st[k] == 0.f;
      const float Lk = sL[k];
      // row term of class k as the "i" index: L_k * sum_j relu(1 + pos_k*neg_j*(p_j - p_k))
      if (Lk != 0.f) {
        float s = 0.f;
        for (int j = 0; j < C; ++j) {
          const float f = (kpos && st[j] == 0.f) ? 1.f : 0.f;
          const float e = 1.f + f * (sp[j] - pk);
          if (e > 0.f) { s += e; g -= f * Lk; }
        }
        l_local += Lk * s;
      }
      // class k as the "j" index of other positives
      if (kneg) {
        for (int i = 0; i < C; ++i) {
          if (st[i] > 0.f && sL[i] != 0.f && 1.f + (pk - sp[i]) > 0.f) g += sL[i];
        }
      }
      grad[static_cast<size_t>(b) * C + k] = g;
    }
  } else {
    float ls = 0.f, ts = 0.f;
    for (int k = threadIdx.x; k < C; k += blockDim.x) {
      ls += sL[k];
      const float posk = st[k] > 0.f ? 1.f : 0.f, negk = st[k] == 0.f ? 1.f : 0.f;
      ts += 1.f - posk * sp[k] + negk * sp[k];
    }
    const float Lsum = block_sum(ls, red);
    const float Tsum = block_sum(ts, red);
    if (threadIdx.x == 0) l_local = Lsum * Tsum;
    for (int k = threadIdx.x; k < C; k += blockDim.x) {
      const float posk = st[k] > 0.f ? 1.f : 0.f, negk = st[k] == 0.f ? 1.f : 0.f;
      grad[static_cast<size_t>(b) * C + k] = Lsum * (negk - posk);
    }
  }
  const float l_row = block_sum(l_local, red);
  if (threadIdx.x == 0) partial[b] = l_row;
  if (last_block(ticket)) {
    if (threadIdx.x == 0) {
      float S = 0.f;
      for (int r = 0; r < batch; ++r) S += partial[r];
      loss[0] = S;
      *ticket = 0u;
    }
  }
}

// ------------------------------------------------------------------------------------------------ sigmoid BCE
// gluon.loss.SigmoidBinaryCrossEntropyLoss (train_simple_r3d.py:76,237): per-sample mean over classes.
//   from_sigmoid = 0:  relu(x) - x*z + log(1 + exp(-|x|))           grad = (sigmoid(x) - z) / C
//   from_sigmoid = 1: -(log(p + 1e-12)*z + log(1 - p + 1e-12)*(1-z))  grad = -(z/(p+eps) - (1-z)/(1-p+eps)) / C
__global__ void __launch_bounds__(kLossThreads)
bce_kernel(const float* __restrict__ pred, const float* __restrict__ target, int C, int from_sigmoid,
           float* __restrict__ loss, float* __restrict__ grad) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float invc = 1.f / static_cast<float>(C);
  float acc = 0.f;
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    const float x = pred[static_cast<size_t>(b) * C + k];
    const float z = target[static_cast<size_t>(b) * C + k];
    float l, g;
    if (!from_sigmoid) {
      l = fmaxf(x, 0.f) - x * z + log1pf(expf(-fabsf(x)));
      g = 1.f / (1.f + expf(-x)) - z;
    } else {
      const float eps = 1e-12f;
      l = -(logf(x + eps) * z + logf(1.f - x + eps) * (1.f - z));
      g = -(z / (x + eps) - (1.f - z) / (1.f - x + eps));
    }
    acc += l;
    if (grad != nullptr) grad[static_cast<size_t>(b) * C + k] = g * invc;
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) loss[b] = s * invc;
}

// ------------------------------------------------------------------------------------------------ softmax head
// mode 0: gluon.loss.SoftmaxCrossEntropyLoss (train_simple_r3d.py:43): loss_b = -log_softmax(x_b)[label_b],
//         grad = softmax - onehot.        out = loss[batch]
// mode 1: mx.sym.SoftmaxOutput(multi_output, use_ignore, normalization='null') (net.py:167-169): out = softmax
//         probabilities [batch, C]; grad = p - onehot, rows whose label == ignore_label (-1) get zero gradient.
__global__ void __launch_bounds__(kLossThreads)
softmax_kernel(const float* __restrict__ logits, const float* __restrict__ label, int C, int mode,
               float* __restrict__ out, float* __restrict__ grad) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float* x = logits + static_cast<size_t>(b) * C;
  float mx = -INFINITY;
  for (int k = threadIdx.x; k < C; k += blockDim.x) mx = fmaxf(mx, x[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float se = 0.f;
  for (int k = threadIdx.x; k < C; k += blockDim.x) se += expf(x[k] - mx);
  const float sum = block_sum(se, red);
  const float lab = label[b];
  // MXNet's `pick` (SoftmaxCrossEntropyLoss) clips the class index into [0, C-1]; a NaN label picks class 0
  int li = lab == lab ? static_cast<int>(fminf(fmaxf(lab, -1.f), static_cast<float>(C))) : 0;
  if (mode == 0) li = li < 0 ? 0 : (li >= C ? C - 1 : li);
  const bool ignored = (mode == 1) && (lab == -1.f);
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    const float p = expf(x[k] - mx) / sum;
    if (mode == 1) out[static_cast<size_t>(b) * C + k] = p;
    if (grad != nullptr) grad[static_cast<size_t>(b) * C + k] = ignored ? 0.f : (p - (k == li ? 1.f : 0.f));
  }
  if (mode == 0 && threadIdx.x == 0) out[b] = -(x[li] - mx - logf(sum));
}

}  // namespace fvt

using namespace fvt;

static int loss_common_check(fvt_handle_t handle, const void* pred, const void* target, int batch, int C) {
  if (pred == nullptr || target == nullptr) return set_error(FVT_ERR_BAD_DESC, "null tensor pointer");
  if (batch <= 0 || C <= 0 || C > kMaxClass) return set_error(FVT_ERR_BAD_DESC, "batch=%d num_class=%d out of range (num_class <= %d)", batch, C, kMaxClass);
  int st = 0;
  if (handle_device(handle, &st) == nullptr) return st;
  return 0;
}

extern "C" {

size_t fvt_loss_workspace_bytes(int32_t batch) { return (static_cast<size_t>(batch) + 4) * sizeof(float); }

int fvt_lsep_fwd_bwd(fvt_handle_t handle, const float* pred, const float* target, int32_t batch, int32_t num_class, int32_t mode,
                     float* loss, float* grad, void* workspace, void* stream) {
  if (int e = loss_common_check(handle, pred, target, batch, num_class)) return e;
  if (loss == nullptr || grad == nullptr || workspace == nullptr) return set_error(FVT_ERR_BAD_DESC, "null output pointer");
  if (mode != 0 && mode != 1) return set_error(FVT_ERR_BAD_DESC, "lsep mode must be 0 (LsepLoss) or 1 (LSEP_funcLoss as written)");
  float* ws = static_cast<float*>(workspace);
  cudaMemsetAsync(ws + batch, 0, 4 * sizeof(float), (cudaStream_t)stream);
  lsep_kernel<<<batch, kLossThreads, 0, (cudaStream_t)stream>>>(pred, target, batch, num_class, mode, loss, grad, ws,
                                                                reinterpret_cast<unsigned int*>(ws + batch));
  return check_launch("lsep_kernel");
}

int fvt_warp_fwd_bwd(fvt_handle_t handle, const float* pred, const float* target, int32_t batch, int32_t num_class, int32_t label_size,
                     int32_t max_trials, int32_t mode, uint64_t seed, uint64_t sample_offset, const float* rank_in,
                     float* rank_out, int32_t* trials_out, float* loss, float* grad, void* workspace, void* stream) {
  if (int e = loss_common_check(handle, pred, target, batch, num_class)) return e;
  if (loss == nullptr || grad == nullptr || workspace == nullptr) return set_error(FVT_ERR_BAD_DESC, "null output pointer");
  if (mode != 0 && mode != 1) return set_error(FVT_ERR_BAD_DESC, "warp mode must be 0 (WarpLoss) or 1 (WARP_funcLoss)");
  if (max_trials < 1) return set_error(FVT_ERR_BAD_DESC, "max_trials must be >= 1");
  // rank_weights has label_size entries and is indexed by floor(max_trials / trials) <= max_trials (mlc_loss.py:144-147)
  if (rank_in == nullptr && max_trials > label_size - 1)
    return set_error(FVT_ERR_BAD_DESC, "max_trials=%d would index rank_weights[%d] of a %d-entry table (IndexError in the reference)",
                     max_trials, max_trials, label_size);
  float* ws = static_cast<float*>(workspace);
  cudaMemsetAsync(ws + batch, 0, 4 * sizeof(float), (cudaStream_t)stream);
  warp_kernel<<<batch, kLossThreads, 0, (cudaStream_t)stream>>>(pred, target, batch, num_class, label_size, max_trials, mode,
                                                                seed, sample_offset, rank_in, rank_out, trials_out, loss,
                                                                grad, ws, reinterpret_cast<unsigned int*>(ws + batch));
  return check_launch("warp_kernel");
}

int fvt_bce_fwd_bwd(fvt_handle_t handle, const float* pred, const float* target, int32_t batch, int32_t num_class, int32_t from_sigmoid,
                    float* loss, float* grad, void* stream) {
  if (int e = loss_common_check(handle, pred, target, batch, num_class)) return e;
  if (loss == nullptr) return set_error(FVT_ERR_BAD_DESC, "null output pointer");
  bce_kernel<<<batch, kLossThreads, 0, (cudaStream_t)stream>>>(pred, target, num_class, from_sigmoid, loss, grad);
  return check_launch("bce_kernel");
}

int fvt_softmax_fwd_bwd(fvt_handle_t handle, const float* logits, const float* label, int32_t batch, int32_t num_class, int32_t mode,
                        float* out, float* grad, void* stream) {
  if (int e = loss_common_check(handle, logits, label, batch, num_class)) return e;
  if (out == nullptr) return set_error(FVT_ERR_BAD_DESC, "null output pointer");
  if (mode != 0 && mode != 1) return set_error(FVT_ERR_BAD_DESC, "softmax mode must be 0 (SoftmaxCrossEntropyLoss) or 1 (SoftmaxOutput)");
  softmax_kernel<<<batch, kLossThreads, 0, (cudaStream_t)stream>>>(logits, label, num_class, mode, out, grad);
  return check_launch("softmax_kernel");
}

// Raw Philox4x32-10 block, exposed so tests can pin the device stream to the published known-answer vectors.
int fvt_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]) {
  const uint4 r = philox4x32_10_host(make_uint4(counter[0], counter[1], counter[2], counter[3]), make_uint2(key[0], key[1]));
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
  return 0;
}

}  // extern "C"
