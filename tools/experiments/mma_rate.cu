// Microbenchmark: issue rate of tcgen05.mma cta_group::1 kind::f16 M=128 for several N, back-to-back on one
// accumulator vs alternating two accumulators, operands in (uninitialised) 128B-swizzled smem.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include "../../fastvideotagging_b200/csrc/ptx.cuh"
using namespace fvt;

__global__ void __launch_bounds__(128, 1) k(int n, int iters, int two_acc, int kadv, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
  if (warp == 1) { ptx::tmem_alloc(ptx::smem_u32(&slot), 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, n, 0, 0);
    const uint64_t ad = ptx::make_sw128_desc(ptx::smem_u32(smem), 16, 1024);
    const uint64_t bd = ptx::make_sw128_desc(ptx::smem_u32(smem) + 32768, 16, 1024);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tm + ((two_acc && (i & 1)) ? 256 : 0);
      const int kk = kadv ? (i & 3) : 0;
      ptx::umma_bf16_ss(d, ad + 2 * kk, bd + 2 * kk, idesc, 1);
    }
    long long t1 = clock64();
    ptx::umma_commit(ptx::smem_u32(&bar));
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  for (int two = 0; two < 2; ++two)
    for (int n : {64, 128, 144, 192, 256}) {
      k<<<1, 128, 100 * 1024>>>(n, iters, two, 1, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("N=%3d two_acc=%d: issue %.1f clk/MMA, complete %.1f clk/MMA (ideal %d)\n", n, two, (double)h[0] / iters, (double)h[1] / iters, n / 2);
    }
  // all SMs busy at once (power / shared limits?)
  for (int n : {144, 256}) {
    k<<<148, 128, 100 * 1024>>>(n, iters, 0, 1, d);
    cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("148 CTAs N=%3d: complete %.1f clk/MMA (ideal %d)\n", n, (double)h[1] / iters, n / 2);
  }
  return 0;
}
