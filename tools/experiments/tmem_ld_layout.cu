// Experiment: which (TMEM lane, column) does each register of tcgen05.ld.16x256b.x2 hold?
// TMEM is filled with tcgen05.st.32x32b.x16 (thread t -> lane t, columns 0..15, value = lane*100 + column), then read
// back with the 16x256b shape at lane offsets 0 and 16; every thread prints its registers decoded as (lane, column).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include "../../fastvideotagging_b200/csrc/ptx.cuh"
using namespace fvt;

__global__ void __launch_bounds__(128, 1) k(int* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&slot), 32); ptx::tmem_relinquish(); }
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    uint32_t v[16];
    for (int c = 0; c < 16; ++c) v[c] = lane * 100 + c;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
        ::"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(tm)
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    for (int half = 0; half < 2; ++half) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(tm + (static_cast<uint32_t>(half * 16) << 16))
                   : "memory");
      ptx::tmem_ld_wait();
      for (int i = 0; i < 8; ++i) out[(half * 32 + lane) * 8 + i] = static_cast<int>(r[i]);
    }
  }
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 32); }
}

int main() {
  int* d; cudaMalloc(&d, 64 * 8 * sizeof(int));
  k<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  int h[64 * 8]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int half = 0; half < 2; ++half)
    for (int t = 0; t < 32; ++t) {
      printf("half %d thread %2d:", half, t);
      for (int i = 0; i < 8; ++i) printf(" r%d=(l%2d,c%2d)", i, h[(half * 32 + t) * 8 + i] / 100, h[(half * 32 + t) * 8 + i] % 100);
      printf("\n");
    }
  return 0;
}
