// Microbenchmark 2: completion rate of tcgen05.mma M=128 when the A descriptor starts at a row that is not a multiple
// of the 8-row swizzle atom (the "shifted tap" views used by conv_slab.cuh / conv_wgrad_slab.cuh), for K-major and
// MN-major operands, several N.  Operands live in zero... constant-filled 128B-swizzled smem.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include "../../fastvideotagging_b200/csrc/ptx.cuh"
using namespace fvt;

__global__ void __launch_bounds__(128, 1) k(int n, int iters, int mn_major, int shift_rows, int lbo_rows, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
  if (warp == 1) { ptx::tmem_alloc(ptx::smem_u32(&slot), 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, n, mn_major, mn_major);
    const uint32_t a_addr = ptx::smem_u32(smem) + shift_rows * 128;
    const uint32_t b_addr = ptx::smem_u32(smem) + 65536;
    const uint64_t ad = mn_major ? ptx::make_sw128_desc(a_addr, lbo_rows * 128, 1024) : ptx::make_sw128_desc(a_addr, 16, 1024);
    const uint64_t bd = mn_major ? ptx::make_sw128_desc(b_addr, 16384, 1024) : ptx::make_sw128_desc(b_addr, 16, 1024);
    const int kstep = mn_major ? 128 : 2;
    const int kmod = mn_major ? 7 : 3;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int kk = i & kmod;
      ptx::umma_bf16_ss(tm + ((i >> 3) & 1) * 256, ad + kstep * kk, bd + kstep * kk, idesc, 1);
    }
    ptx::umma_commit(ptx::smem_u32(&bar));
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
    long long t2 = clock64();
    out[0] = t2 - t0;
  }
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4000;
  for (int mn = 0; mn < 2; ++mn)
    for (int n : {64, 80, 96, 144, 256})
      for (int shift : {0, 1, 3, 8, 58, 59}) {
        const int lbos[3] = {64, 1, 57};
        for (int li = 0; li < (mn ? 3 : 1); ++li) {
          for (int grid : {1, 148}) {
            k<<<grid, 128, 200 * 1024>>>(n, iters, mn, shift, lbos[li], d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
            long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%s N=%3d shift=%2d lbo_rows=%2d grid=%3d: %.1f clk/MMA (floor %d)\n", mn ? "MN-major" : "K-major ", n, shift,
                   mn ? lbos[li] : 0, grid, (double)h[0] / iters, n / 2);
          }
        }
      }
  return 0;
}
