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

}  // namespace fvt
