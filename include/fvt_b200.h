/*
 * fvt_b200.h — C ABI of the B200-native R(2+1)D hot path (libfvt_b200.so).
 *
 * The reference (bruceyang2012/FastVideoTagging) has no operator code of its own: every arithmetic step of
 * model/R2Plus1.py, net.py and model/mlc_loss.py is an MXNet operator (cuDNN/cuBLAS/mshadow).  This header is the
 * replacement boundary for exactly those operator calls; each entry point cites the reference call sites it
 * stands in for.  Plain C: raw device pointers, sizes and a cudaStream_t (passed as void*).
 *
 * Conventions
 *   - activations: NDHWC bf16, channel count stored padded to a multiple of 16 ("stored channels"); pad
 *     channels hold zeros.
 *   - every launching function takes an fvt_handle_t first: one handle per (host thread, device), the analogue of the
 *     reference's one executor per context (train.py:54, mx.module.Module(net, context=[...])).  The handle owns the tuning
 *     switches only; it holds no buffers and no pointers of the caller.
 *   - every function is stream-ordered, re-entrant and allocates nothing; the caller owns all buffers, workspaces included,
 *     and passes them per call (fvt_conv3d_workspace_bytes says how much a call would like).
 *   - deterministic: no floating-point atomics anywhere — split reductions go through workspace slices added in a fixed
 *     order, per-channel statistics through exact integer accumulators (fvt_stats_bytes); the same inputs give the same
 *     bits on every run.
 *   - return value: 0 on success, negative fvt_status otherwise; fvt_last_error() gives a thread-local message.
 *   - no CPU fallback: on a device that is not sm_100 fvt_create returns FVT_ERR_UNSUPPORTED_ARCH.
 */
#ifndef FVT_B200_H_
#define FVT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum fvt_status {
  FVT_OK = 0,
  FVT_ERR_BAD_DESC = -1,
  FVT_ERR_UNSUPPORTED_ARCH = -2,
  FVT_ERR_MISALIGNED = -3,
  FVT_ERR_WORKSPACE = -4,
  FVT_ERR_CUDA = -5,
  FVT_ERR_DRIVER = -6,
  FVT_ERR_BAD_HANDLE = -7
} fvt_status;

typedef struct fvt_handle_s* fvt_handle_t;

/* epilogue / mode flags for fvt_conv_desc.flags */
#define FVT_CONV_RELU 1      /* y = max(y, 0)                       (Activation 'relu', R2Plus1.py:33,60,81)   */
#define FVT_CONV_RESIDUAL 2  /* y += residual before the ReLU       (nd.relu(y+x), R2Plus1.py:81; net.py:100)   */
#define FVT_CONV_STATS 4     /* accumulate per-channel sum, sum^2 of the raw conv output (training BatchNorm) */
#define FVT_CONV_BN_BWD 16   /* fvt_conv3d_fwd as a DATA GRADIENT fused with the first pass of the BatchNorm backward it feeds.  The
                              * convolution's output is dact = d(loss)/d(act) of a layer with act = relu(raw*scale + shift);
                              * with this flag (plus FVT_CONV_STATS | FVT_CONV_RESIDUAL) `residual` is that layer's RAW conv
                              * output and scale/shift its forward BatchNorm constants: the epilogue stores
                              * dz = dact * [raw*scale + shift > 0] instead of dact and accumulates stats_acc =
                              * [sum dz*raw (cout), sum dz (cout)] of the stored values.  fvt_bn_backward(dz_in = 2) finishes
                              * the BatchNorm backward with one pass instead of two.                                     */
#define FVT_CONV_W_OHWI 8    /* fvt_pack_conv_weight[_dgrad] / fvt_conv3d_wgrad only: the fp32 weight (gradient) tensor is
                              * laid out (O, kT, kH, kW, I) instead of the reference's (O, I, kT, kH, kW).  With input
                              * channels innermost a warp of the weight-gradient epilogue adds 32 consecutive floats
                              * (one coalesced reduction) instead of 32 scattered ones.                           */

/* One 3-D convolution (cross-correlation, no bias, dilation 1) — the parameters of nn.Conv3D / mx.sym.Convolution
 * at R2Plus1.py:27-31,34-38,67-70,100-111 and net.py:40-42,49-51,95-96,122-131. */
typedef struct fvt_conv_desc {
  int32_t n, t, h, w;     /* input extent                                                   */
  int32_t cin;            /* stored input channels (multiple of 16)                          */
  int32_t cout;           /* stored output channels (multiple of 16)                         */
  int32_t kt, kh, kw;     /* filter                                                         */
  int32_t st, sh, sw;     /* stride                                                         */
  int32_t pt, ph, pw;     /* symmetric zero padding                                         */
  int32_t flags;          /* FVT_CONV_*                                                     */
  int32_t block_n;        /* N tile of the implicit GEMM; 0 = library default               */
} fvt_conv_desc;

/* ---- library ------------------------------------------------------------------------------------------- */
int fvt_version(void);
const char* fvt_last_error(void);
/* 0 if `device` is an sm_100 part and the driver exposes the tensor-map encoders, else a negative status. */
int fvt_device_check(int device);
/* Handle for `device` (must be sm_100, else FVT_ERR_UNSUPPORTED_ARCH).  Use it only while `device` is the calling thread's
 * current CUDA device (FVT_ERR_BAD_HANDLE otherwise); one handle per (thread, device); destroy it with fvt_destroy. */
int fvt_create(fvt_handle_t* handle, int device);
int fvt_destroy(fvt_handle_t handle);
/* Tuning/debug switches of ONE handle (A/B runs and tests; defaults are the measured best).  "disable_slab" = 1 routes
 * every convolution through the generic im2col kernel (K1); "disable_frame_ring" / "disable_temporal_is" = 1 do the same
 * for the temporal kernels (K1t / K1i) only; "disable_b_stationary" = 1 makes K1 stream its weights;
 * "disable_wgrad_slab" = 1 routes every weight gradient through the im2col kernel (K3); "disable_split_k" = 1 keeps
 * small-M convolutions single-pass; "slab_prefetch" / "slab_box_rows" / "ring_prefetch" / "slab_single_stage" /
 * "slab_epi_warps" / "debug_flags" / "wgrad_no_store" are load-path and epilogue experiments (tools/gpu_*_ab.py);
 * "slab_pair_auto" = 0 keeps the layers whose filter fits two SMs but not one off the CTA-pair kernel (2: pair even for
 * small problems), "igemm_pair" = 0 keeps the generic im2col convolution on single CTAs (2: pairs even for small
 * problems), "slab_pair" = 1|2 forces the pair kernel for single-SM-stationary layers (1: staged TMA store),
 * "unit_input_stationary" = 0 selects the output-stationary form of the fused (2+1)D unit, "disable_tis_tma_store" = 0
 * turns K1i's TMA-store epilogue on, "disable_dgrad_direct" = 1 sends strided data gradients through fvt_zero_insert. */
int fvt_set_option(fvt_handle_t handle, const char* name, int value);
int fvt_get_option(fvt_handle_t handle, const char* name, int* value);

/* ---- exact per-channel accumulators ----------------------------------------------------------------------------
 * Training-mode BatchNorm sums (sum x, sum x^2; sum dz*xhat, sum dz) are accumulated across CTAs EXACTLY in 128-bit
 * fixed point (four 64-bit limbs per value, integer atomics: order-independent, so bit-reproducible).  A statistics
 * buffer for c_store channels is [2][c_store] accumulators = fvt_stats_bytes(c_store) bytes, 8-byte aligned, zeroed by
 * the caller (cudaMemsetAsync) before the producing launch.  fvt_stats_encode / fvt_stats_decode convert n plain floats
 * to / from n accumulators (tests, interop). */
size_t fvt_stats_bytes(int32_t c_store);
int fvt_stats_encode(fvt_handle_t handle, const float* values, void* stats_acc, int32_t n, void* stream);
int fvt_stats_decode(fvt_handle_t handle, const void* stats_acc, float* values, int32_t n, void* stream);

/* ---- convolution (K1) ----------------------------------------------------------------------------------- */
/* Output extent floor((x + 2p - k)/s) + 1 per axis (MXNet convention). */
int fvt_conv3d_out_shape(const fvt_conv_desc* d, int32_t* to, int32_t* ho, int32_t* wo);
/* N tile the library will use for this descriptor (multiple of 16, <= 256). */
int fvt_conv3d_block_n(const fvt_conv_desc* d);
/* Number of bf16 elements of the packed weight buffer: rows(cout rounded up to whole N tiles) x kt*kh*kw*cin. */
size_t fvt_conv3d_packed_weight_elems(const fvt_conv_desc* d);
/* Pack fp32 weights in the reference layout (O, I, kT, kH, kW) (device pointer, cout_real x cin_real filters)
 * into the K-major bf16 layout K1 consumes; rows/channels beyond the real counts are zero. */
int fvt_pack_conv_weight(fvt_handle_t handle, const fvt_conv_desc* d, const float* w_oidhw, int32_t cout_real, int32_t cin_real,
                         void* w_packed, void* stream);
/* y = epilogue(conv(x, w)):  acc*scale[c] + shift[c] (if scale != NULL)  (+ residual)  (ReLU)  -> bf16.
 * stats_acc (FVT_CONV_STATS, training forward: scale/shift must be NULL): [2][cout] exact accumulators
 * (fvt_stats_bytes(cout)), sum then sum of squares of the bf16-rounded raw conv output — zero them first.
 * workspace (optional, may be NULL): caller-owned scratch, contents irrelevant on entry.  With room for >= 2 fp32
 * [M, cout] slices it lets small-M convolutions (fewer output tiles than SMs) split their reduction over several CTAs
 * (split-K: every split stores its partial tile into its own slice, one finalize pass adds the slices in split order);
 * without it every convolution runs single-pass.  fvt_conv3d_workspace_bytes(handle, d, FVT_OP_FWD, 0, 0) is the size
 * the call would like. */
int fvt_conv3d_fwd(fvt_handle_t handle, const fvt_conv_desc* d, const void* x, const void* w_packed, const float* scale,
                   const float* shift, const void* residual, void* y, void* stats_acc, void* workspace,
                   size_t workspace_bytes, void* stream);
/* The same convolution with (a) per-axis HIGH padding pad_hi[t,h,w] (d->pt/ph/pw stay the LOW padding) and (b) an output
 * lattice: output pixel (ot, oh, ow) is stored at (ot*out_stride[0] + out_offset[0], oh*.., ow*..) of a tensor with extent
 * out_extent[t,h,w] (cout stored channels; `residual`, if flagged, is read at the same place).  Output extent per axis =
 * floor((x + p_lo + pad_hi - k)/s) + 1.  Always runs on the generic kernel (K1); no statistics, no split-K.
 *
 * This is the building block of the data gradient of a STRIDED convolution (the nine stride-2 convs of the net; cuDNN
 * backward-data in the reference).  For stride s, filter k, padding p per axis, dX positions i = s*j + par (parity class
 * par) only receive the taps kk with (par + p - kk) % s == 0, from dY[j + e], e = (par + p - kk)/s >= 0.  So every class
 * is a STRIDE-1 convolution of dY with the sub-filter {e -> tap par + p - s*e} (fvt_pack_entry kind 2: tap_a = par + p,
 * tap_s = s), low padding 0, high padding = class extent - dY extent + sub-filter extent - 1, written to the lattice
 * (stride s, offset par) of dX.  s_t*s_h*s_w launches on dY replace "zero-insert dY to the input extent, then a full
 * convolution": a 1x3x3/s2 layer needs 9 filter-tap GEMMs instead of 36 and no staging tensor.  Classes no tap reaches
 * (1x1x1/s2 shortcuts) stay zero: clear dX first.  engine.py (DgradPlan) composes the calls. */
typedef struct fvt_conv_ext {
  int32_t pad_hi[3];
  int32_t out_extent[3];
  int32_t out_stride[3];
  int32_t out_offset[3];
} fvt_conv_ext;
int fvt_conv3d_fwd_ex(fvt_handle_t handle, const fvt_conv_desc* d, const fvt_conv_ext* ext, const void* x, const void* w_packed,
                      const float* scale, const float* shift, const void* residual, void* y, void* stream);
/* Workspace bytes a call with this descriptor would like (0: none).  op = FVT_OP_FWD: fvt_conv3d_fwd (also the data
 * gradient, which is a forward convolution); FVT_OP_WGRAD: fvt_conv3d_wgrad (cout_real / cin_real as there; <= 0: the
 * stored counts).  A smaller workspace is legal: the library uses as many splits as fit. */
#define FVT_OP_FWD 0
#define FVT_OP_WGRAD 1
size_t fvt_conv3d_workspace_bytes(fvt_handle_t handle, const fvt_conv_desc* d, int32_t op, int32_t cout_real, int32_t cin_real);

/* ---- fused (2+1)D unit (K2f, inference) ------------------------------------------------------------------------
 * y = relu( bn_out( conv3x1x1( relu( bn_mid( conv1x3x3(x) ) ) ) ) [+ residual] ) in ONE launch: the factorised unit
 * get_spatial_temporal_conv (model/R2Plus1.py:19-40, net.py:31-52) plus the BatchNorm/ReLU/add that R3DBlock wraps
 * around it (model/R2Plus1.py:59-62,76-81), eval mode.  d_spatial: 1x3x3, stride 1, pad (0,1,1), cin = 64, cout = mid
 * (stored, <= 144); d_temporal: 3x1x1, stride 1, pad (1,0,0), cin = mid, cout = 64, same N/T/H/W; FVT_CONV_RESIDUAL in
 * d_temporal->flags adds `residual` before the final ReLU.  Weights are the buffers fvt_pack_conv_weight makes for the
 * two descriptors; scale/shift are the folded BatchNorm constants (mid: d_spatial->cout floats, out: 64 floats).
 * The mid tensor stays in tensor memory.  fvt_unit2p1_supported returns 1 when the pair of descriptors is eligible on
 * the current device, 0 when the caller must use two fvt_conv3d_fwd calls. */
int fvt_unit2p1_supported(fvt_handle_t handle, const fvt_conv_desc* d_spatial, const fvt_conv_desc* d_temporal);
int fvt_unit2p1_fwd(fvt_handle_t handle, const fvt_conv_desc* d_spatial, const fvt_conv_desc* d_temporal, const void* x,
                    const void* w_spatial_packed, const float* scale_mid, const float* shift_mid,
                    const void* w_temporal_packed, const float* scale_out, const float* shift_out,
                    const void* residual, void* y, void* stream);

/* ---- stem input transform --------------------------------------------------------------------------------- */
/* Clips in the reference layout NCDHW fp32 (data/data.py:46-47) with 3 channels -> W-unfolded NDHWC bf16
 * u[n,t,h,ow, kw*3+ci] = x[n,ci,t,h, ow*sw - pw + kw] (zero outside), channels >= 3*kw_taps are zero.
 * The 1x7x7/s(1,2,2) stem conv (R2Plus1.py:100-104, net.py:122-123) then runs on K1 as a (1,7,1)/s(1,2,1)
 * conv over u with cin = cu. */
int fvt_stem_unfold(fvt_handle_t handle, const float* x_ncdhw, void* u, int32_t n, int32_t t, int32_t h, int32_t w, int32_t kw_taps,
                    int32_t sw, int32_t pw, int32_t cu, void* stream);
/* Row-paired variant (h even): u2[n,t,h/2,ow, (h&1)*cu + kw*3+ci] — rows 2*h2 and 2*h2+1 side by side in 2*cu channels.
 * The stem's stride-2 walk over H is then a STRIDE-1 (1,5,1) conv over h2 with pad (0,2,0) (w2[o, par*cu+kw*3+ci, kh2] =
 * w[o, ci, kh = 2*kh2+par-1, kw], zero where kh is outside 0..6), which the slab kernel runs reading every input row
 * once instead of 7 times (same reference call sites as fvt_stem_unfold). */
int fvt_stem_unfold_hpair(fvt_handle_t handle, const float* x_ncdhw, void* u, int32_t n, int32_t t, int32_t h, int32_t w, int32_t kw_taps,
                          int32_t sw, int32_t pw, int32_t cu, void* stream);

/* ---- fp32 path (inference only) --------------------------------------------------------------------------------
 * The same convolution + folded-BatchNorm / residual / ReLU epilogue in plain fp32 on the CUDA cores: activations NDHWC
 * fp32 with the REAL channel counts (no padding rule), weights (kT, kH, kW, I, O) fp32.  desc.flags: FVT_CONV_RELU,
 * FVT_CONV_RESIDUAL.  It exists to check the layer semantics against an fp32 reference at rel 1e-4 (north-star "fp32
 * path", BASELINE configs[0]: the reference's own fp32 CPU-runnable case); the bf16 tcgen05 kernels are the fast path. */
int fvt_conv3d_fwd_f32(fvt_handle_t handle, const fvt_conv_desc* desc, const float* x, const float* w_thwio, const float* scale,
                       const float* shift, const float* residual, float* y, void* stream);
int fvt_pool_fc_fwd_f32(fvt_handle_t handle, const float* x, int32_t n, int32_t positions, int32_t c, const float* w, const float* b,
                        int32_t num_class, float* pooled, float* logits, void* stream);

/* ---- head: global average pool + dense (A5) ------------------------------------------------------------------ */
/* x: [n, positions, c] bf16 (NDHWC with T*H*W flattened); pooled (optional out): [n, c] fp32;
 * logits[n, k] = sum_c pooled[n,c] * w[k,c] + b[k]   (AvgPool3D + Dense, R2Plus1.py:168-171,243-245). */
int fvt_pool_fc_fwd(fvt_handle_t handle, const void* x, int32_t n, int32_t positions, int32_t c, int32_t c_real, const float* w,
                    const float* b, int32_t num_class, float* pooled, float* logits, void* stream);

/* ---- training: convolution gradients (K2, K3) ----------------------------------------------------------------- */
/* Data gradient of a STRIDE-1 convolution is itself a convolution (K1) of dY with the channel-transposed, tap-reversed
 * filter and padding k-1-p.  `d` describes that convolution (d->cin = stored forward Cout, d->cout = stored forward
 * Cin); this packs the forward weights (O, I, kT, kH, kW) for it.  Strided convolutions go through fvt_zero_insert
 * first.  (cuDNN backward-data in the reference.) */
int fvt_pack_conv_weight_dgrad(fvt_handle_t handle, const fvt_conv_desc* d, const float* w_oidhw, int32_t fwd_cout_real,
                               int32_t fwd_cin_real, void* w_packed, void* stream);
/* All operand copies of a training step in ONE launch (the step re-packs 69 + 68 conv weights after every optimiser
 * update).  table_dev: DEVICE array of n_entries fvt_pack_entry, sorted by block0 with block0 = running sum of nblocks
 * and nblocks = fvt_pack_entry_blocks(kind, taps, k_store, rows); total_blocks = their sum.  Sources are fp32 masters in
 * the (O, kT, kH, kW, I) layout (FVT_CONV_W_OHWI).  kind 0 = the layout fvt_pack_conv_weight makes for a descriptor d
 * (rows = packed rows = elems / (taps*d->cin), k_store = d->cin, cout_real / cin_real as there); kind 1 = the layout
 * fvt_pack_conv_weight_dgrad makes for the data-gradient descriptor dd (rows = packed rows of dd, k_store = dd->cin =
 * stored forward Cout, cout_real / cin_real = the FORWARD filter counts); kind 2 = one parity sub-filter (below). */
typedef struct fvt_pack_entry {
  const float* w;
  void* out;
  int32_t kind, taps, k_store, rows, cout_real, cin_real;
  uint32_t block0, nblocks;
  /* kind 2 only (sub-filter of a strided convolution's data gradient, see fvt_conv3d_fwd_ex): the packed filter has
   * taps = sub[0]*sub[1]*sub[2] taps u = (ut, uh, uw); tap u copies source tap (tap_a[a] - tap_s[a]*u[a]) per axis
   * a = t, h, w of the src_k[0] x src_k[1] x src_k[2] forward filter:  out[r][u][k] = w[k][tap(u)][r]. */
  int32_t sub[3], src_k[3], tap_a[3], tap_s[3];
} fvt_pack_entry;
uint32_t fvt_pack_entry_blocks(int32_t kind, int32_t taps, int32_t k_store, int32_t rows);
int fvt_pack_conv_weights_multi(fvt_handle_t handle, const fvt_pack_entry* table_dev, int32_t n_entries, uint32_t total_blocks,
                                void* stream);
/* dw[(co*cin_real + ci)*taps + tap] = sum over output pixels of dy[pixel, co] * x[pixel + tap, ci]  — fp32, the
 * reference's (O, I, kT, kH, kW) layout (FVT_CONV_W_OHWI in d->flags: (O, kT, kH, kW, I)), OVERWRITTEN (MXNet
 * grad_req='write').  `d` is the FORWARD descriptor; x: stored input activation, dy: gradient w.r.t. the raw conv
 * output, both NDHWC bf16.  (cuDNN backward-filter.)
 * workspace (optional): caller-owned scratch, contents irrelevant.  The reduction over output pixels is split over CTAs;
 * every split stores its partial dW into its own dW-shaped slice of the workspace and one pass adds the slices in split
 * order into dw (deterministic, no atomics).  Without a workspace (or one smaller than two slices) the launch uses a
 * single split per dW tile: correct, slower on large-M layers.  fvt_conv3d_workspace_bytes(.., FVT_OP_WGRAD, ..). */
int fvt_conv3d_wgrad(fvt_handle_t handle, const fvt_conv_desc* d, const void* x, const void* dy, float* dw, int32_t cout_real,
                     int32_t cin_real, void* workspace, size_t workspace_bytes, void* stream);
/* Weight gradients of SEVERAL layers in one launch (the analogue of cuDNN backward-filter called once per Conv3D of a
 * residual stage, reference model/R2Plus1.py:73-82 under autograd, train_simple_r3d.py:118-125).  A weight gradient feeds
 * nothing but the optimiser, so a training loop may defer the layers of a stage and run them together: the small-feature-
 * map stages (conv4_x / conv5_x at a few clips per GPU) cannot fill the machine one layer at a time.
 * fvt_conv3d_wgrad_group_plan: descs[l] / x[l] / dy[l] / dw[l] / cout_real[l] / cin_real[l] as for fvt_conv3d_wgrad
 * (FVT_CONV_W_OHWI per descriptor).  in_group[l] (out) = 1 when layer l is part of the grouped launch, 0 when the caller
 * must run fvt_conv3d_wgrad for it (strided and 1x1x1 layers).  *table_bytes / *workspace_bytes_needed (out): sizes of the
 * launch table and of the slice workspace the group wants (0: no layer is split).  With host_table == NULL only the
 * sizes and in_group are computed (x / dy / dw may be NULL).  Otherwise host_table (>= *table_bytes) is filled; the
 * caller copies it to 128-byte aligned device memory once and passes both copies to fvt_conv3d_wgrad_group_run.  The
 * table bakes in the tensor pointers and the workspace pointer: re-plan when any of them changes.  Deterministic
 * (pixel splits meet through workspace slices added in split order); every dw is overwritten. */
int fvt_conv3d_wgrad_group_plan(fvt_handle_t handle, int32_t n, const fvt_conv_desc* descs, const void* const* x, const void* const* dy,
                                float* const* dw, const int32_t* cout_real, const int32_t* cin_real, void* workspace,
                                size_t workspace_bytes, void* host_table, size_t host_table_bytes, size_t* table_bytes,
                                size_t* workspace_bytes_needed, int32_t* in_group);
int fvt_conv3d_wgrad_group_run(fvt_handle_t handle, const void* host_table, const void* device_table, void* stream);
/* up[n, to*st, ho*sh, wo*sw, :] = dy[n, to, ho, wo, :], zero elsewhere (up has the conv input's T,H,W). */
int fvt_zero_insert(fvt_handle_t handle, const void* dy, void* up, int32_t n, int32_t t, int32_t h, int32_t w, int32_t to, int32_t ho,
                    int32_t wo, int32_t st, int32_t sh, int32_t sw, int32_t c_store, void* stream);

/* ---- training: BatchNorm (K5-K7), MXNet semantics (A4) ---------------------------------------------------------- */
/* Eval-mode BatchNorm folded into the producing convolution's epilogue, for every layer of a network in one launch:
 * scale[c] = gamma[c]/sqrt(var[c] + eps), shift[c] = beta[c] - mean[c]*scale[c] for c < c_real, (0, 0) for the pad channels
 * up to c_store (they stay exactly zero).  table_dev: DEVICE array of n_entries fvt_bn_fold_entry (device pointers).
 * (nn.BatchNorm in inference mode, model/R2Plus1.py:32,59,62,71,105,112; running statistics in the MXNet convention.) */
typedef struct fvt_bn_fold_entry {
  const float* gamma;
  const float* beta;
  const float* mean;
  const float* var;
  float* scale;
  float* shift;
  int32_t c_real, c_store;
  float eps;
  int32_t reserved;
} fvt_bn_fold_entry;
int fvt_bn_fold_multi(fvt_handle_t handle, const fvt_bn_fold_entry* table_dev, int32_t n_entries, void* stream);
/* stats_acc = [sum(c_store), sum^2(c_store)] exact accumulators from fvt_conv3d_fwd(FVT_CONV_STATS) over `rows` pixels ->
 * mean, inv_std = 1/sqrt(biased_var + eps), scale = gamma*inv_std, shift = beta - mean*scale;
 * running = momentum*running + (1-momentum)*batch (biased variance) when running_mean != NULL. */
int fvt_bn_finalize(fvt_handle_t handle, const void* stats_acc, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, int32_t c_store, int32_t c_real, int64_t rows, float eps, float momentum,
                    float* scale, float* shift, float* mean, float* invstd, void* stream);
/* out = relu?( raw*scale + shift [+ res | + res*res_scale + res_shift] ), [rows, c_store] bf16. */
int fvt_bn_apply(fvt_handle_t handle, const void* raw, const float* scale, const float* shift, const void* res,
                 const float* res_scale, const float* res_shift, void* out, int64_t rows, int32_t c_store, int32_t relu,
                 void* stream);
/* fvt_bn_finalize followed by fvt_bn_apply in ONE launch (the training forward's per-layer pair): every CTA derives
 * scale/shift from `stats_acc` itself; scale/shift/mean/invstd and the running statistics are written once, as by
 * fvt_bn_finalize.  Replaces nn.BatchNorm + Activation('relu') at reference model/R2Plus1.py:32-33,59-60,62,81. */
int fvt_bn_finalize_apply(fvt_handle_t handle, const void* stats_acc, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, int32_t c_store, int32_t c_real, int64_t rows, float eps, float momentum,
                          float* scale, float* shift, float* mean, float* invstd, const void* raw, const void* res,
                          const float* res_scale, const float* res_shift, void* out, int32_t relu, void* stream);
/* dz = dact * [mask > 0]  (mask tensor given: the ReLU sits after a residual add, R2Plus1.py:81),
 *    = dact * [raw*relu_scale + relu_shift > 0]  (relu_scale/relu_shift = the forward scale/shift of this BatchNorm:
 *      the ReLU directly follows it, R2Plus1.py:33,60 — the mask is recomputed from raw, saving one tensor read),
 *    = dact  (neither given);
 * sums = [dgamma(c_store), dbeta(c_store)] (overwritten);  sums_acc: scratch of fvt_stats_bytes(c_store) bytes (the exact
 * accumulators behind `sums`; zeroed inside, contents irrelevant on entry);
 * draw = gamma*inv_std*(dz - dbeta/rows - xhat*dgamma/rows);  dz_out (optional) receives dz.
 * dz_in = 2: `dact` already IS dz and sums_acc already holds [sum dz*raw, sum dz] — both produced by the data gradient
 * convolution that wrote dact (fvt_conv3d_fwd with FVT_CONV_BN_BWD) — so only the apply pass runs
 * (dgamma = inv_std*(sum dz*raw - mean*sum dz), formed in double from the exact sums).  dz_in = 1: the same with
 * sums_acc = [sum dz*(raw-mean), sum dz]. */
int fvt_bn_backward(fvt_handle_t handle, const void* raw, const void* dact, const void* mask, const float* mean,
                    const float* invstd, const float* gamma, const float* relu_scale, const float* relu_shift, float* sums,
                    void* sums_acc, void* draw, void* dz_out, int64_t rows, int32_t c_store, int32_t c_real, int32_t dz_in,
                    void* stream);

/* ---- training: head backward, optimiser ----------------------------------------------------------------------- */
/* dw[k,c] = sum_n dlogits[n,k]*pooled[n,c]; db[k] = sum_n dlogits[n,k] (overwritten); dx[n,p,c] = (dlogits[n,:] . w[:,c]) / positions. */
int fvt_pool_fc_bwd(fvt_handle_t handle, const float* dlogits, const float* pooled, const float* w, int32_t n, int32_t num_class,
                    int32_t c, int32_t positions, float* dw, float* db, void* dx, int32_t c_store, void* stream);
/* MXNet sgd_mom_update over a tensor list in one launch (gluon.Trainer 'sgd', train_simple_r3d.py:95-97,124):
 * g' = rescale*g + wd*w; mom = momentum*mom - lr*lr_mult*g'; w += mom.
 * tensor_table: device array of {float* w; const float* g; float* mom; uint64 numel; float wd; float lr_mult};
 * chunk_tensor/chunk_offset: device arrays, one entry per CTA: tensor index and chunk index inside it. */
int fvt_sgd_momentum_multi(fvt_handle_t handle, const void* tensor_table, const uint32_t* chunk_tensor, const uint32_t* chunk_offset,
                           int32_t num_chunks, uint32_t chunk_elems, float lr, float momentum, float rescale,
                           void* stream);

/* ---- losses (K9-K11): forward and backward in one launch, gradients are d(loss)/d(pred) for head-gradient 1 ------ */
/* Scratch for the loss kernels: (batch + 4) floats. */
size_t fvt_loss_workspace_bytes(int32_t batch);
/* LSEP.  mode 0 = LsepLoss.forward (model/mlc_loss.py:63-86) + its autodiff gradient;
 *        mode 1 = LSEP_funcLoss exactly as written, row-index quirk and -1/loss backward included (:8-54).
 * pred/target: [batch, num_class] fp32; loss: float[1]; grad: [batch, num_class]. */
int fvt_lsep_fwd_bwd(fvt_handle_t handle, const float* pred, const float* target, int32_t batch, int32_t num_class, int32_t mode,
                     float* loss, float* grad, void* workspace, void* stream);
/* WARP.  mode 0 = WarpLoss.forward (:122-174), mode 1 = WARP_funcLoss (:177-233).
 * Negative sampling (np.random.choice in the reference, :137,207) is replaced by the counter-based stream
 * philox4x32_10(key=seed, counter=(sample_offset+row, class j, trial, 0)).x % n_neg over the ascending negative list.
 * rank_in != NULL skips sampling and uses the given rank weights L[batch, num_class];
 * rank_out / trials_out (optional) receive L and the number of draws per positive. */
int fvt_warp_fwd_bwd(fvt_handle_t handle, const float* pred, const float* target, int32_t batch, int32_t num_class, int32_t label_size,
                     int32_t max_trials, int32_t mode, uint64_t seed, uint64_t sample_offset, const float* rank_in,
                     float* rank_out, int32_t* trials_out, float* loss, float* grad, void* workspace, void* stream);
/* gluon SigmoidBinaryCrossEntropyLoss (train_simple_r3d.py:76,237): loss[batch] = mean over classes;
 * grad (optional) = d(sum_b loss_b)/d(pred). */
int fvt_bce_fwd_bwd(fvt_handle_t handle, const float* pred, const float* target, int32_t batch, int32_t num_class, int32_t from_sigmoid,
                    float* loss, float* grad, void* stream);
/* mode 0: gluon SoftmaxCrossEntropyLoss with sparse labels (train_simple_r3d.py:43): out = loss[batch];
 * mode 1: mx.sym.SoftmaxOutput (net.py:167-169): out = probabilities [batch, num_class], label -1 ignored.
 * label: float[batch] class indices; grad (optional) = softmax - onehot (un-normalised, as MXNet). */
int fvt_softmax_fwd_bwd(fvt_handle_t handle, const float* logits, const float* label, int32_t batch, int32_t num_class, int32_t mode,
                        float* out, float* grad, void* stream);
/* ---- rows next to the hot path ----------------------------------------------------------------- */
/* N2, clip pre-processing (videos_reader.py:69-76,93-97; data/ucf101.py:124-128).
 * clips_nthwc: decoded uint8 frames (N, T, H, W, 3).  fvt_clip_stats_u8: sums6 = [sum x (3 channels), sum x^2 (3)] over all
 * `pixels` = N*T*H*W pixels (exact integers; overwritten).  fvt_clip_normalize_u8:
 * out[n, c, t, h, w] = (clips[n, t, h, w', c]*scale - mean[c]) * inv_std[c] in the reference's NCDHW fp32 layout, with
 * w' = W-1-w for clips whose flip[n] != 0 (flip may be NULL).  mean / inv_std are HOST arrays of 3 floats. */
int fvt_clip_stats_u8(fvt_handle_t handle, const uint8_t* clips_nthwc, int64_t pixels, uint64_t* sums6, void* stream);
int fvt_clip_normalize_u8(fvt_handle_t handle, const uint8_t* clips_nthwc, const uint8_t* flip, float* out_ncdhw, int32_t n, int32_t t, int32_t h,
                          int32_t w, float scale, const float mean[3], const float inv_std[3], void* stream);
/* N2 fused into the stem's input transform: decoded uint8 frames (N, T, Hs, Ws, 3) -> crop (h x w at crop_yx[n] = (y0, x0);
 * NULL: frames are already h x w) -> flip -> (v*scale - mean[c]) * inv_std[c] -> the W-unfolded NDHWC bf16 stem input of
 * fvt_stem_unfold (hpair = 0) / fvt_stem_unfold_hpair (hpair = 1), bit for bit what fvt_clip_normalize_u8 followed by
 * those produces, in ONE pass and without the fp32 NCDHW tensor.  crop_yx is a DEVICE int32[n][2]; mean / inv_std are
 * HOST arrays.  (videos_reader.py:44-49,58,69-76,93-97; data/ucf101.py:124-128.) */
int fvt_clip_unfold_u8(fvt_handle_t handle, const uint8_t* clips_nthwc, const uint8_t* flip, const int32_t* crop_yx, void* u, int32_t n,
                       int32_t t, int32_t hs, int32_t ws, int32_t h, int32_t w, float scale, const float mean[3],
                       const float inv_std[3], int32_t kw_taps, int32_t sw, int32_t pw, int32_t cu, int32_t hpair, void* stream);
/* N3, evaluation tail.  acc[rows, C] += softmax(logits[rows, C]) (validation.py:49-51);
 * pred[row] = argmax acc[row] (first maximum), *correct += number of rows with pred == labels (validation.py:61-63). */
int fvt_softmax_accumulate(fvt_handle_t handle, const float* logits, float* acc, int32_t rows, int32_t num_class, void* stream);
int fvt_argmax_correct(fvt_handle_t handle, const float* acc, const int32_t* labels, int32_t rows, int32_t num_class, int32_t* pred,
                       uint64_t* correct, void* stream);
/* Top-k IoU counts (train_simple_r3d.py:169-193): per row the k largest scores in `argsort()[:, ::-1]` order (ties: larger
 * index first), labels = {j : target > 0.1}; inter[k-1] += |top_k & labels|, uni[k-1] += |top_k | labels|, k = 1..k_max <= 4.
 * The caller zeroes inter / uni and adds the reference's 1e-4 offsets when forming the ratio. */
int fvt_topk_iou(fvt_handle_t handle, const float* scores, const float* target, int32_t rows, int32_t num_class, int32_t k_max, uint64_t* inter,
                 uint64_t* uni, void* stream);
/* Host-side Philox4x32-10 block (same code the device uses) for known-answer tests. */
int fvt_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* FVT_B200_H_ */
