// Microbenchmark: sustained time per TMA 2-D box load [rows x 128 B] (SWIZZLE_128B) into shared memory, one issuing
// thread per CTA, `depth` loads in flight, for several row pitches (contiguous / aligned / unaligned) and box heights.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../fastvideotagging_b200/csrc/ptx.cuh"
using namespace fvt;

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tm, int box_rows, int depth, int iters, int total_rows, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[16];
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) ptx::mbar_init(ptx::smem_u32(&bar[i]), 1); ptx::fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x < 32) {
    // warp-uniform loop, elected lane issues (a divergent single-thread issuer gets wrapped in serialisation loops)
    const int box_bytes = box_rows * 128;
    const int nbox = total_rows / box_rows;
    int row = (blockIdx.x * 977) % nbox;
    int s = 0; uint32_t par = 0;
    const uint32_t bar0 = ptx::smem_u32(&bar[0]), sm0 = ptx::smem_u32(smem);
    long long t0 = clock64();
    for (int i = 0; i < iters + depth; ++i) {
      if (i >= depth) ptx::mbar_wait(bar0 + s * 8, par);
      if (i < iters) {
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(bar0 + s * 8, box_bytes);
          ptx::tma_load_2d(sm0 + s * box_bytes, &tm, bar0 + s * 8, 0, row * box_rows);
        }
        __syncwarp();
        row += 37; if (row >= nbox) row -= nbox;
      }
      if (++s == depth) { s = 0; if (i >= depth) par ^= 1; }
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const size_t bytes = 64ull << 20;          // 64 MiB: L2 resident after warm-up
  uint8_t* d; cudaMalloc(&d, bytes + 4096); cudaMemset(d, 0, bytes + 4096);
  long long* out; cudaMalloc(&out, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int pitches[] = {128, 256, 288, 2304, 2592};
  for (int pitch : pitches) {
    const int total_rows = (int)(bytes / pitch);
    for (int box_rows : {64, 128, 232}) {
      CUtensorMap tm;
      cuuint64_t dims[2] = {64, (cuuint64_t)total_rows}, str[1] = {(cuuint64_t)pitch};
      cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, es[2] = {1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
      for (int depth : {1, 2, 4, 6}) {
        if (depth * box_rows * 128 > 190 * 1024) continue;
        for (int grid : {1, 148}) {
          const int iters = 400;
          k<<<grid, 128, 200 * 1024>>>(tm, box_rows, depth, iters, total_rows, out);   // warm L2
          k<<<grid, 128, 200 * 1024>>>(tm, box_rows, depth, iters, total_rows, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
          long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
          printf("pitch %4d box %3d rows depth %d grid %3d: %7.1f clk/box  %5.2f clk/row  %5.1f B/clk/SM\n", pitch, box_rows, depth, grid,
                 (double)h / iters, (double)h / iters / box_rows, box_rows * 128.0 * iters / h);
        }
      }
    }
  }
  return 0;
}
