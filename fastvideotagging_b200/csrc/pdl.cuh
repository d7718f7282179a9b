// Programmatic dependent launch (PDL): the launches of a training / inference step are hundreds of short, dependent kernels
// (conv4_x / conv5_x at a few clips per GPU: 5-10 us each), and the gap between two dependent kernels — the grid of the
// second is only scheduled once the first has drained — is a visible part of the step.  Every kernel of the hot path
// starts with fvt_pdl_entry(): `griddepcontrol.launch_dependents` lets the NEXT kernel of the stream be scheduled as soon as
// all CTAs of this one are running (its CTAs become resident on free SMs), and `griddepcontrol.wait` then blocks until the
// PREVIOUS kernel of the stream has completed and its writes are visible — before this kernel reads or writes any global
// memory, so the data dependencies are exactly those of an ordinary stream.  Both instructions are no-ops for a launch
// without the attribute.  Host side: fvt::launch() adds cudaLaunchAttributeProgrammaticStreamSerialization when the handle
// option "pdl" is on — ONLY for kernels that call fvt_pdl_entry() (a kernel without the wait must never get the attribute).
// Works inside CUDA graph capture (programmatic dependency edges).
#pragma once
#include <cuda_runtime.h>

namespace fvt {

__device__ __forceinline__ void fvt_pdl_entry() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, int cluster, bool pdl,
                          Args&&... args) {
  cudaLaunchConfig_t cfg;
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(static_cast<unsigned>(block));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = static_cast<unsigned>(cluster); attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace fvt
