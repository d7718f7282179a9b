// C-ABI host side of libfvt_b200.so: descriptor validation, TMA tensor-map encoding, kernel launches.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <unordered_map>

#include "../../include/fvt_b200.h"
#include "conv_igemm.cuh"
#include "conv_wgrad.cuh"
#include "conv_slab.cuh"
#include "host_common.h"

namespace fvt {

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ------------------------------------------------------------------------------------------------ driver entry points
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct DeviceInfo {
  bool checked = false;
  int status = 0;
  int sm_count = 0;
  int driver_version = 0;
  EncodeTiledFn encode_tiled = nullptr;
  EncodeIm2colFn encode_im2col = nullptr;
};
static DeviceInfo g_dev[16];
static int g_disable_slab = 0;   // fvt_set_option("disable_slab", 1): force the generic im2col kernel (A/B runs, tests)
static std::mutex g_mu;

static int resolve_driver(DeviceInfo& di) {
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr) return set_error(FVT_ERR_DRIVER, "cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
  di.encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  fn = nullptr;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr) return set_error(FVT_ERR_DRIVER, "cuTensorMapEncodeIm2col not available: %s", cudaGetErrorString(e));
  di.encode_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
  return 0;
}

const DeviceInfo* device_info(int device, int* status) {
  if (device < 0 || device >= 16) { *status = set_error(FVT_ERR_BAD_DESC, "device index %d out of range", device); return nullptr; }
  std::lock_guard<std::mutex> lk(g_mu);
  DeviceInfo& di = g_dev[device];
  if (!di.checked) {
    di.checked = true;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
      di.status = set_error(FVT_ERR_CUDA, "cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
    } else if (prop.major != 10) {
      di.status = set_error(FVT_ERR_UNSUPPORTED_ARCH, "device %d is sm_%d%d; this library is sm_100a only (no fallback)",
                            device, prop.major, prop.minor);
    } else {
      di.sm_count = prop.multiProcessorCount;
      cudaDriverGetVersion(&di.driver_version);
      di.status = resolve_driver(di);
    }
  }
  *status = di.status;
  if (di.status != 0 && g_err[0] == 0) set_error(di.status, "device %d unusable (status %d)", device, di.status);
  return di.status == 0 ? &di : nullptr;
}

const DeviceInfo* current_device_info(int* status) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { *status = set_error(FVT_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e)); return nullptr; }
  return device_info(dev, status);
}
int sm_count_of(const DeviceInfo* di) { return di->sm_count; }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FVT_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------------------------------------------------------ conv helpers
static int validate_conv(const fvt_conv_desc* d) {
  if (d == nullptr) return set_error(FVT_ERR_BAD_DESC, "null conv descriptor");
  if (d->n <= 0 || d->t <= 0 || d->h <= 0 || d->w <= 0) return set_error(FVT_ERR_BAD_DESC, "non-positive input extent");
  if (d->cin <= 0 || d->cin % 16) return set_error(FVT_ERR_BAD_DESC, "cin=%d must be a positive multiple of 16", d->cin);
  if (d->cout <= 0 || d->cout % 16) return set_error(FVT_ERR_BAD_DESC, "cout=%d must be a positive multiple of 16", d->cout);
  if (d->cout > kMaxCout - 256) return set_error(FVT_ERR_BAD_DESC, "cout=%d exceeds the supported maximum %d", d->cout, kMaxCout - 256);
  if (d->kt < 1 || d->kh < 1 || d->kw < 1 || d->kt > 16 || d->kh > 16 || d->kw > 16) return set_error(FVT_ERR_BAD_DESC, "filter extent out of range");
  if (d->st < 1 || d->sh < 1 || d->sw < 1 || d->st > 8 || d->sh > 8 || d->sw > 8) return set_error(FVT_ERR_BAD_DESC, "stride must be in [1, 8]");
  if (d->pt < 0 || d->ph < 0 || d->pw < 0 || d->pt > 15 || d->ph > 15 || d->pw > 15) return set_error(FVT_ERR_BAD_DESC, "padding must be in [0, 15]");
  if (d->pt - (d->kt - 1) < -16 || d->ph - (d->kh - 1) < -16 || d->pw - (d->kw - 1) < -16) return set_error(FVT_ERR_BAD_DESC, "filter/padding outside the im2col corner range");
  if (d->t + 2 * d->pt < d->kt || d->h + 2 * d->ph < d->kh || d->w + 2 * d->pw < d->kw) return set_error(FVT_ERR_BAD_DESC, "filter larger than padded input");
  if (d->block_n != 0 && (d->block_n % 16 || d->block_n < 16 || d->block_n > 256)) return set_error(FVT_ERR_BAD_DESC, "block_n=%d must be a multiple of 16 in [16, 256]", d->block_n);
  return 0;
}

static void conv_out_shape(const fvt_conv_desc* d, int* to, int* ho, int* wo) {
  *to = (d->t + 2 * d->pt - d->kt) / d->st + 1;
  *ho = (d->h + 2 * d->ph - d->kh) / d->sh + 1;
  *wo = (d->w + 2 * d->pw - d->kw) / d->sw + 1;
}

static int pick_block_n(const fvt_conv_desc* d) {
  if (d->block_n) return d->block_n;
  const int c = d->cout;
  const int nt0 = (c + 255) / 256;
  int best_bn = 0, best_total = 1 << 30;
  for (int nt = nt0; nt <= nt0 + 2; ++nt) {
    int bn = ((c + nt - 1) / nt + 15) / 16 * 16;
    if (bn > 256) continue;
    int total = bn * ((c + bn - 1) / bn);
    if (total < best_total) { best_total = total; best_bn = bn; }
  }
  return best_bn;
}

static int weight_rows(const fvt_conv_desc* d, int bn) { return (d->cout + bn - 1) / bn * bn; }

// ------------------------------------------------------------------------------------------------ weight packing
// dgrad = 1: pack the weights of the data-gradient convolution instead (input/output channels swapped, taps
// reversed): out[ci][taps-1-tap][co] = w[co][ci][tap]; here rows index ci and the K axis runs over (tap, co).
__global__ void pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int rows, int taps,
                                   int cin_store, int cout_real, int cin_real, int dgrad) {
  const size_t total = static_cast<size_t>(rows) * taps * cin_store;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % cin_store);
    const size_t r = i / cin_store;
    const int tap = static_cast<int>(r % taps);
    const int o = static_cast<int>(r / taps);
    float v = 0.f;
    if (!dgrad) {
      if (o < cout_real && ci < cin_real) v = w[(static_cast<size_t>(o) * cin_real + ci) * taps + tap];
    } else {
      // this conv: out channel o = forward ci, in channel ci = forward co; forward dims are (cin_real, cout_real)
      if (o < cout_real && ci < cin_real) v = w[(static_cast<size_t>(ci) * cout_real + o) * taps + (taps - 1 - tap)];
    }
    out[i] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------------ tensor maps
static int encode_x_map(const DeviceInfo* di, const fvt_conv_desc* d, const void* x, CUtensorMap* map) {
  const cuuint64_t dims[5] = {(cuuint64_t)d->cin, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->t, (cuuint64_t)d->n};
  const cuuint64_t strides[4] = {(cuuint64_t)d->cin * 2, (cuuint64_t)d->cin * 2 * d->w,
                                 (cuuint64_t)d->cin * 2 * d->w * d->h, (cuuint64_t)d->cin * 2 * d->w * d->h * d->t};
  const int lower[3] = {-d->pw, -d->ph, -d->pt};
  const int upper[3] = {d->pw - (d->kw - 1), d->ph - (d->kh - 1), d->pt - (d->kt - 1)};
  const cuuint32_t estr[5] = {1, (cuuint32_t)d->sw, (cuuint32_t)d->sh, (cuuint32_t)d->st, 1};
  CUresult r = di->encode_im2col(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, lower,
                                 upper, kBlockK, kBlockM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(FVT_ERR_DRIVER, "cuTensorMapEncodeIm2col failed (CUresult %d)", (int)r);
  // Driver quirk (<= 13.1): im2col maps over tensors smaller than 128 KiB come back with a bit set that makes
  // the load fault; clear it (same workaround CUTLASS applies).
  const size_t bytes = (size_t)d->cin * 2 * d->w * d->h * d->t * d->n;
  if (di->driver_version <= 13010 && bytes < 131072) reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
  return 0;
}

static int encode_w_map(const DeviceInfo* di, const void* w, int k_total, int rows, int bn, CUtensorMap* map) {
  const cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
  const cuuint32_t box[2] = {kBlockK, (cuuint32_t)bn};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = di->encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(FVT_ERR_DRIVER, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return 0;
}

}  // namespace fvt

using namespace fvt;

extern "C" {

int fvt_version(void) { return 101; }

int fvt_set_option(const char* name, int value) {
  if (name != nullptr && strcmp(name, "disable_slab") == 0) { g_disable_slab = value; return 0; }
  return set_error(FVT_ERR_BAD_DESC, "unknown option");
}

const char* fvt_last_error(void) { return g_err; }

int fvt_device_check(int device) {
  int st = 0;
  device_info(device, &st);
  return st;
}

int fvt_conv3d_out_shape(const fvt_conv_desc* d, int32_t* to, int32_t* ho, int32_t* wo) {
  if (int e = validate_conv(d)) return e;
  int a, b, c;
  conv_out_shape(d, &a, &b, &c);
  if (to) *to = a;
  if (ho) *ho = b;
  if (wo) *wo = c;
  return 0;
}

int fvt_conv3d_block_n(const fvt_conv_desc* d) {
  if (int e = validate_conv(d)) return e;
  return pick_block_n(d);
}

size_t fvt_conv3d_packed_weight_elems(const fvt_conv_desc* d) {
  if (validate_conv(d)) return 0;
  const int bn = pick_block_n(d);
  return (size_t)weight_rows(d, bn) * d->kt * d->kh * d->kw * d->cin;
}

int fvt_pack_conv_weight(const fvt_conv_desc* d, const float* w_oidhw, int32_t cout_real, int32_t cin_real,
                         void* w_packed, void* stream) {
  if (int e = validate_conv(d)) return e;
  if (cout_real <= 0 || cout_real > d->cout || cin_real <= 0 || cin_real > d->cin)
    return set_error(FVT_ERR_BAD_DESC, "real filter counts (%d, %d) exceed stored (%d, %d)", cout_real, cin_real, d->cout, d->cin);
  if (w_oidhw == nullptr || w_packed == nullptr) return set_error(FVT_ERR_BAD_DESC, "null weight pointer");
  const int bn = pick_block_n(d);
  const int rows = weight_rows(d, bn);
  const int taps = d->kt * d->kh * d->kw;
  const size_t total = (size_t)rows * taps * d->cin;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  pack_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w_oidhw, (__nv_bfloat16*)w_packed, rows, taps, d->cin,
                                                              cout_real, cin_real, 0);
  return check_launch("pack_weight_kernel");
}

int fvt_pack_conv_weight_dgrad(const fvt_conv_desc* d, const float* w_oidhw, int32_t fwd_cout_real, int32_t fwd_cin_real,
                               void* w_packed, void* stream) {
  // `d` describes the data-gradient convolution: d->cin = stored forward Cout, d->cout = stored forward Cin.
  if (int e = validate_conv(d)) return e;
  if (fwd_cout_real <= 0 || fwd_cout_real > d->cin || fwd_cin_real <= 0 || fwd_cin_real > d->cout)
    return set_error(FVT_ERR_BAD_DESC, "forward filter counts (%d, %d) exceed the dgrad descriptor (%d, %d)", fwd_cout_real, fwd_cin_real, d->cin, d->cout);
  if (w_oidhw == nullptr || w_packed == nullptr) return set_error(FVT_ERR_BAD_DESC, "null weight pointer");
  const int bn = pick_block_n(d);
  const int rows = weight_rows(d, bn);
  const int taps = d->kt * d->kh * d->kw;
  const size_t total = (size_t)rows * taps * d->cin;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  // kernel view: rows o = forward ci (< fwd_cin_real), K channel ci = forward co (< fwd_cout_real)
  pack_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w_oidhw, (__nv_bfloat16*)w_packed, rows, taps, d->cin,
                                                              fwd_cin_real, fwd_cout_real, 1);
  return check_launch("pack_weight_kernel(dgrad)");
}

int fvt_conv3d_fwd(const fvt_conv_desc* d, const void* x, const void* w_packed, const float* scale,
                   const float* shift, const void* residual, void* y, float* stats, void* stream) {
  if (int e = validate_conv(d)) return e;
  if (x == nullptr || w_packed == nullptr || y == nullptr) return set_error(FVT_ERR_BAD_DESC, "null tensor pointer");
  if ((scale == nullptr) != (shift == nullptr)) return set_error(FVT_ERR_BAD_DESC, "scale and shift must be given together");
  if ((d->flags & FVT_CONV_RESIDUAL) && residual == nullptr) return set_error(FVT_ERR_BAD_DESC, "FVT_CONV_RESIDUAL without a residual tensor");
  if ((d->flags & FVT_CONV_STATS) && stats == nullptr) return set_error(FVT_ERR_BAD_DESC, "FVT_CONV_STATS without a stats buffer");
  if (((uintptr_t)x | (uintptr_t)w_packed | (uintptr_t)y | (uintptr_t)residual) & 15)
    return set_error(FVT_ERR_MISALIGNED, "tensor pointers must be 16-byte aligned");
  int st = 0;
  const DeviceInfo* di = current_device_info(&st);
  if (di == nullptr) return st;

  int to, ho, wo;
  conv_out_shape(d, &to, &ho, &wo);
  const int bn = pick_block_n(d);
  const int rows = weight_rows(d, bn);
  const int taps = d->kt * d->kh * d->kw;

  // ---- K1s: stride-1 'same' spatial convs with <= 128 input channels load each input row once (conv_slab.cuh)
  if (!g_disable_slab && d->kt == 1 && d->st == 1 && d->sh == 1 && d->sw == 1 && d->pt == 0 && (d->kh > 1 || d->kw > 1) &&
      2 * d->ph == d->kh - 1 && 2 * d->pw == d->kw - 1 && d->cin % 64 == 0 && d->cin <= 128 && d->w + 2 * d->pw <= 128) {
    SlabParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.frames = d->n * d->t; sp.h = d->h; sp.w = d->w; sp.wp = d->w + 2 * d->pw;
    sp.ph = d->ph; sp.pw = d->pw; sp.kh = d->kh; sp.kw = d->kw;
    sp.r_out = 128 / sp.wp;
    if (sp.r_out > d->h) sp.r_out = d->h;
    sp.r_in = sp.r_out + d->kh - 1;
    sp.tiles_per_frame = (d->h + sp.r_out - 1) / sp.r_out;
    const double useful = (double)d->h * d->w / ((double)sp.tiles_per_frame * 128.0);
    sp.cin_blocks = d->cin / 64; sp.cin_k16 = d->cin / 16;
    sp.n_tile = bn; sp.num_n_tiles = rows / bn;
    const int slot_rows = (128 + (d->kh - 1) * sp.wp + d->kw - 1 + 7) / 8 * 8;
    sp.slab_slot_bytes = slot_rows * 128;
    sp.slab_tx_bytes = sp.wp * sp.r_in * 128;
    const int stage_bytes = sp.cin_blocks * sp.slab_slot_bytes;
    const int b_slab = bn * 128;
    const int b_all = taps * sp.cin_blocks;
    const bool want_stats = (d->flags & FVT_CONV_STATS) != 0;
    const int aux = (512 + 2 * rows * 4 + (want_stats ? 2 * bn * 4 : 0) + 255) / 256 * 256;
    const int kSmemMax = 227 * 1024;
    bool ok = useful >= 0.6 && sp.r_in * sp.wp <= slot_rows && sp.r_in <= 256;
    if (ok) {
      if (sp.num_n_tiles == 1 && b_all <= kSlabMaxBRing && b_all * b_slab + 2 * stage_bytes + aux <= kSmemMax) {
        sp.b_stationary = 1;
        sp.b_ring = b_all;
        sp.stages = (kSmemMax - aux - b_all * b_slab) / stage_bytes;
      } else {
        sp.b_stationary = 0;
        sp.stages = 2;
        sp.b_ring = (kSmemMax - aux - 2 * stage_bytes) / b_slab;
        if (sp.b_ring > kSlabMaxBRing) sp.b_ring = kSlabMaxBRing;
        if (sp.b_ring < 4) ok = false;
      }
      if (sp.stages > kSlabMaxStages) sp.stages = kSlabMaxStages;
    }
    if (ok) {
      sp.cout_store = d->cout; sp.flags = d->flags;
      sp.scale = scale; sp.shift = shift; sp.residual = (const __nv_bfloat16*)residual;
      sp.y = (__nv_bfloat16*)y; sp.stats = stats;
      const int smem_bytes = sp.b_ring * b_slab + sp.stages * stage_bytes + aux;
      CUtensorMap tmx, tmw;
      const cuuint64_t dims[4] = {(cuuint64_t)d->cin, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)sp.frames};
      const cuuint64_t strides[3] = {(cuuint64_t)d->cin * 2, (cuuint64_t)d->cin * 2 * d->w, (cuuint64_t)d->cin * 2 * d->w * d->h};
      const cuuint32_t box[4] = {64, (cuuint32_t)sp.wp, (cuuint32_t)sp.r_in, 1};
      const cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = di->encode_tiled(&tmx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return set_error(FVT_ERR_DRIVER, "cuTensorMapEncodeTiled(slab x) failed (CUresult %d)", (int)r);
      if (int e = encode_w_map(di, w_packed, taps * d->cin, rows, bn, &tmw)) return e;
      static bool attr_set_s[16] = {false};
      int dev = 0;
      cudaGetDevice(&dev);
      if (!attr_set_s[dev]) {
        cudaError_t e = cudaFuncSetAttribute(conv_slab_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
        if (e != cudaSuccess) return set_error(FVT_ERR_CUDA, "cudaFuncSetAttribute(conv_slab_fwd_kernel): %s", cudaGetErrorString(e));
        attr_set_s[dev] = true;
      }
      const int m_tiles = sp.frames * sp.tiles_per_frame;
      const int grid = m_tiles < di->sm_count ? m_tiles : di->sm_count;
      conv_slab_fwd_kernel<<<grid, kSlabThreads, smem_bytes, (cudaStream_t)stream>>>(tmx, tmw, sp);
      return check_launch("conv_slab_fwd_kernel");
    }
  }

  ConvKernelParams p;
  memset(&p, 0, sizeof(p));
  p.m_total = d->n * to * ho * wo;
  p.to = to; p.ho = ho; p.wo = wo;
  p.st = d->st; p.sh = d->sh; p.sw = d->sw;
  p.pt = d->pt; p.ph = d->ph; p.pw = d->pw;
  p.kt = d->kt; p.kh = d->kh; p.kw = d->kw;
  p.cin_k16 = d->cin / 16;
  p.cin_blocks = (d->cin + kBlockK - 1) / kBlockK;
  p.k_per_tap = d->cin;
  p.block_n = bn;
  p.num_m_tiles = (p.m_total + kBlockM - 1) / kBlockM;
  p.num_n_tiles = rows / bn;
  p.cout_store = d->cout;
  p.flags = d->flags;
  p.scale = scale; p.shift = shift;
  p.residual = (const __nv_bfloat16*)residual;
  p.y = (__nv_bfloat16*)y;
  p.stats = stats;

  const int stage_bytes = kATileBytes + bn * kBlockK * 2;
  const int kAuxBytes = 4096 + 2 * kMaxCout * 4;   // barriers + stats partials + staged scale/shift
  const int budget = 227 * 1024 - 1024 - kAuxBytes;
  int stages = budget / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return set_error(FVT_ERR_BAD_DESC, "tile does not fit shared memory");
  p.stages = stages;
  const int smem_bytes = 1024 + stages * stage_bytes + kAuxBytes;

  CUtensorMap tmx, tmw;
  if (int e = encode_x_map(di, d, x, &tmx)) return e;
  if (int e = encode_w_map(di, w_packed, taps * d->cin, rows, bn, &tmw)) return e;

  static bool attr_set[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_error(FVT_ERR_CUDA, "cudaFuncSetAttribute(conv_igemm_fwd_kernel): %s", cudaGetErrorString(e));
    attr_set[dev] = true;
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < di->sm_count ? tiles : di->sm_count;
  conv_igemm_fwd_kernel<<<grid, kConvThreads, smem_bytes, (cudaStream_t)stream>>>(tmx, tmw, p);
  return check_launch("conv_igemm_fwd_kernel");
}


int fvt_conv3d_wgrad(const fvt_conv_desc* d, const void* x, const void* dy, float* dw, int32_t cout_real,
                     int32_t cin_real, void* stream) {
  if (int e = validate_conv(d)) return e;
  if (x == nullptr || dy == nullptr || dw == nullptr) return set_error(FVT_ERR_BAD_DESC, "null tensor pointer");
  if (cout_real <= 0 || cout_real > d->cout || cin_real <= 0 || cin_real > d->cin)
    return set_error(FVT_ERR_BAD_DESC, "real filter counts exceed stored");
  if (((uintptr_t)x | (uintptr_t)dy) & 15) return set_error(FVT_ERR_MISALIGNED, "tensor pointers must be 16-byte aligned");
  int st = 0;
  const DeviceInfo* di = current_device_info(&st);
  if (di == nullptr) return st;
  int to, ho, wo;
  conv_out_shape(d, &to, &ho, &wo);

  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.m_total = d->n * to * ho * wo;
  p.to = to; p.ho = ho; p.wo = wo;
  p.st = d->st; p.sh = d->sh; p.sw = d->sw;
  p.pt = d->pt; p.ph = d->ph; p.pw = d->pw;
  p.kt = d->kt; p.kh = d->kh; p.kw = d->kw;
  p.taps = d->kt * d->kh * d->kw;
  p.cin_blocks = (d->cin + 63) / 64;
  p.cin_real = cin_real; p.cout_real = cout_real;
  p.dw = dw;
  fvt_conv_desc tmp = *d;
  tmp.block_n = 0;
  if (d->cin % 64 != 0 && d->cout % 64 == 0) {
    p.mode = 1;
    p.m_groups = d->cout / 64;
    tmp.cout = d->cin;                                  // N runs over the input channels of one tap
  } else {
    p.mode = 0;
    p.m_groups = p.taps * p.cin_blocks;
  }
  p.n_tile = pick_block_n(&tmp);
  p.n_tiles = (tmp.cout + p.n_tile - 1) / p.n_tile;
  p.n_loads = (p.n_tile + 63) / 64;
  p.m_tiles = (p.m_groups + 1) / 2;
  p.kblocks_total = (p.m_total + kWgPix - 1) / kWgPix;
  const int items = p.m_tiles * p.n_tiles * (p.mode == 1 ? p.taps : 1);
  int splits = (2 * di->sm_count + items - 1) / items;
  if (splits > p.kblocks_total) splits = p.kblocks_total;
  if (splits < 1) splits = 1;
  p.kblocks_per_split = (p.kblocks_total + splits - 1) / splits;
  p.splits = (p.kblocks_total + p.kblocks_per_split - 1) / p.kblocks_per_split;
  const int stage_bytes = (2 + p.n_loads) * kSlabBytes;
  int stages = (227 * 1024 - 2048) / stage_bytes;
  if (stages > kWgMaxStages) stages = kWgMaxStages;
  p.stages = stages;
  const int smem_bytes = 1024 + stages * stage_bytes + 1024;

  // tensor maps: X as in the forward pass but 64 pixels per load; dY as a 1x1x1 "im2col" over the output tensor
  CUtensorMap tmx, tmdy;
  {
    const cuuint64_t dims[5] = {(cuuint64_t)d->cin, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->t, (cuuint64_t)d->n};
    const cuuint64_t strides[4] = {(cuuint64_t)d->cin * 2, (cuuint64_t)d->cin * 2 * d->w, (cuuint64_t)d->cin * 2 * d->w * d->h,
                                   (cuuint64_t)d->cin * 2 * d->w * d->h * d->t};
    const int lower[3] = {-d->pw, -d->ph, -d->pt};
    const int upper[3] = {d->pw - (d->kw - 1), d->ph - (d->kh - 1), d->pt - (d->kt - 1)};
    const cuuint32_t estr[5] = {1, (cuuint32_t)d->sw, (cuuint32_t)d->sh, (cuuint32_t)d->st, 1};
    CUresult r = di->encode_im2col(&tmx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, lower, upper,
                                   64, kWgPix, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(FVT_ERR_DRIVER, "cuTensorMapEncodeIm2col(x, wgrad) failed (CUresult %d)", (int)r);
    if (di->driver_version <= 13010 && (size_t)d->cin * 2 * d->w * d->h * d->t * d->n < 131072) reinterpret_cast<uint64_t*>(&tmx)[1] &= ~(1ull << 21);
  }
  {
    const cuuint64_t dims[5] = {(cuuint64_t)d->cout, (cuuint64_t)wo, (cuuint64_t)ho, (cuuint64_t)to, (cuuint64_t)d->n};
    const cuuint64_t strides[4] = {(cuuint64_t)d->cout * 2, (cuuint64_t)d->cout * 2 * wo, (cuuint64_t)d->cout * 2 * wo * ho,
                                   (cuuint64_t)d->cout * 2 * wo * ho * to};
    const int zero3[3] = {0, 0, 0};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = di->encode_im2col(&tmdy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dy), dims, strides, zero3, zero3,
                                   64, kWgPix, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(FVT_ERR_DRIVER, "cuTensorMapEncodeIm2col(dy, wgrad) failed (CUresult %d)", (int)r);
    if (di->driver_version <= 13010 && (size_t)d->cout * 2 * wo * ho * to * d->n < 131072) reinterpret_cast<uint64_t*>(&tmdy)[1] &= ~(1ull << 21);
  }
  static bool attr_set[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_error(FVT_ERR_CUDA, "cudaFuncSetAttribute(conv_wgrad_kernel): %s", cudaGetErrorString(e));
    attr_set[dev] = true;
  }
  const int grid = items * p.splits;
  conv_wgrad_kernel<<<grid, kWgradThreads, smem_bytes, (cudaStream_t)stream>>>(tmx, tmdy, p);
  return check_launch("conv_wgrad_kernel");
}

}  // extern "C"
