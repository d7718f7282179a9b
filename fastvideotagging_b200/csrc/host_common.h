// Host-side helpers shared by the translation units of libfvt_b200.so.
#pragma once
#include <cuda_runtime.h>

namespace fvt {

struct DeviceInfo;

// Records a thread-local error message (returned by fvt_last_error) and returns `code`.
int set_error(int code, const char* fmt, ...);
// cudaGetLastError() -> fvt status.
int check_launch(const char* what);
// sm_100 check + driver entry points for the current device; nullptr (and *status < 0) when unusable.
const DeviceInfo* current_device_info(int* status);
const DeviceInfo* device_info(int device, int* status);
int sm_count_of(const DeviceInfo* di);
// Epilogue pass of a split-K convolution (bn_kernels.cu): ws fp32 -> y bf16 (+ stats), leaves ws zeroed.
int launch_splitk_finalize(float* ws, const float* scale, const float* shift, const void* residual, void* y, float* stats,
                           size_t rows, int c_store, int relu, cudaStream_t stream);

}  // namespace fvt
