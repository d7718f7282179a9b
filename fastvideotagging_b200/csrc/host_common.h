// Host-side helpers shared by the translation units of libfvt_b200.so.
#pragma once
#include <cuda_runtime.h>

#include "../../include/fvt_b200.h"

namespace fvt {

struct DeviceInfo;

// Records a thread-local error message (returned by fvt_last_error) and returns `code`.
int set_error(int code, const char* fmt, ...);
// cudaGetLastError() -> fvt status.
int check_launch(const char* what);
// sm_100 check + driver entry points for the current device; nullptr (and *status < 0) when unusable.
const DeviceInfo* current_device_info(int* status);
// Validates `h` (fvt_create) and that its device is the calling thread's current device; nullptr (and *status < 0) otherwise.
const DeviceInfo* handle_device(fvt_handle_t h, int* status);
const DeviceInfo* device_info(int device, int* status);
// The handle's "pdl" option (programmatic dependent launch, pdl.cuh).
bool handle_pdl(fvt_handle_t h);
int sm_count_of(const DeviceInfo* di);
// Epilogue pass of a split-K convolution (bn_kernels.cu): sum of the `splits` fp32 slices ws[split][rows][c_store], in split
// order -> y bf16 (+ exact per-channel statistics of the bf16-rounded raw output).
int launch_splitk_finalize(const float* ws, int splits, const float* scale, const float* shift, const void* residual, void* y,
                           unsigned long long* stats, size_t rows, int c_store, int relu, cudaStream_t stream, bool pdl);

}  // namespace fvt
