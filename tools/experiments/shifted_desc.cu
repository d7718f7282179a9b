// Experiment: can a K-major SW128 UMMA descriptor start at an arbitrary 128-byte row (not 1024-byte aligned) inside a
// TMA-written slab, and does it need the descriptor's base_offset field?  D[128 x 64] = A_shift[128 x 64] * B[64 x 64]^T
// where A_shift = rows r .. r+127 of a 272-row slab.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../fastvideotagging_b200/csrc/ptx.cuh"
using namespace fvt;

__global__ void __launch_bounds__(128, 1)
k(const __grid_constant__ CUtensorMap tma, const __grid_constant__ CUtensorMap tmb, float* out, int shift, int use_base_offset) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                 // 272 rows * 128 B = 34816
  uint8_t* sb = smem + 36864;         // 64 rows * 128 B
  uint64_t* bar = (uint64_t*)(smem + 36864 + 8192);
  uint64_t* bar2 = bar + 1;
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(bar), 1); ptx::mbar_init(ptx::smem_u32(bar2), 1); ptx::fence_mbar_init(); }
  if (warp == 1) { ptx::tmem_alloc(ptx::smem_u32(slot), 64); ptx::tmem_relinquish(); }
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(ptx::smem_u32(bar), 272 * 128 + 64 * 128);
    // two loads of <=256 rows each (box limit 256)
    ptx::tma_load_2d(ptx::smem_u32(sa), &tma, ptx::smem_u32(bar), 0, 0);
    ptx::tma_load_2d(ptx::smem_u32(sa + 136 * 128), &tma, ptx::smem_u32(bar), 0, 136);
    ptx::tma_load_2d(ptx::smem_u32(sb), &tmb, ptx::smem_u32(bar), 0, 0);
    ptx::mbar_wait(ptx::smem_u32(bar), 0);
    ptx::tc_fence_after();
    const uint32_t a_addr = ptx::smem_u32(sa) + shift * 128;
    uint64_t ad = ptx::make_sw128_desc(a_addr, 16, 1024);
    if (use_base_offset) ad |= (uint64_t)((a_addr >> 7) & 7) << 49;
    const uint64_t bd = ptx::make_sw128_desc(ptx::smem_u32(sb), 16, 1024);
    const uint32_t idesc = ptx::make_idesc_bf16(128, 64, 0, 0);
    for (int kk = 0; kk < 4; ++kk) ptx::umma_bf16_ss(tm, ad + 2 * kk, bd + 2 * kk, idesc, kk > 0);
    ptx::umma_commit(ptx::smem_u32(bar2));
  }
  __syncthreads();
  ptx::mbar_wait(ptx::smem_u32(bar2), 0);
  ptx::tc_fence_after();
  for (int c = 0; c < 64; c += 16) {
    uint32_t v[16];
    ptx::tmem_ld_32x32b_x16(tm + ((uint32_t)(warp * 32) << 16) + c, v);
    ptx::tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 64 + c + i] = __uint_as_float(v[i]);
  }
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 64); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  const int RA = 272, RB = 64, K = 64;
  std::vector<__nv_bfloat16> ha(RA * K), hb(RB * K);
  std::vector<float> fa(RA * K), fb(RB * K);
  srand(1);
  for (int i = 0; i < RA * K; ++i) { float v = (rand() % 17 - 8) / 8.f; ha[i] = __float2bfloat16(v); fa[i] = __bfloat162float(ha[i]); }
  for (int i = 0; i < RB * K; ++i) { float v = (rand() % 13 - 6) / 8.f; hb[i] = __float2bfloat16(v); fb[i] = __bfloat162float(hb[i]); }
  __nv_bfloat16 *da, *db; float* dout;
  cudaMalloc(&da, RA * K * 2); cudaMalloc(&db, RB * K * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(da, ha.data(), RA * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(db, hb.data(), RB * K * 2, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  CUtensorMap tma, tmb;
  cuuint64_t dimsa[2] = {64, (cuuint64_t)RA}, dimsb[2] = {64, (cuuint64_t)RB}, str[1] = {128};
  cuuint32_t boxa[2] = {64, 136}, boxb[2] = {64, 64}, es[2] = {1, 1};
  CUresult r1 = enc(&tma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, dimsa, str, boxa, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&tmb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, dimsb, str, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d %d\n", (int)r1, (int)r2);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> ho(128 * 64);
  for (int ubo = 0; ubo < 2; ++ubo) {
    for (int shift : {0, 1, 2, 3, 5, 7, 8, 9, 58, 59, 117, 118, 136, 137, 140}) {
      cudaMemset(dout, 0, 128 * 64 * 4);
      k<<<1, 128, 64 * 1024>>>(tma, tmb, dout, shift, ubo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("base_offset=%d shift=%d CUDA error %s\n", ubo, shift, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(ho.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
        double s = 0; for (int kk = 0; kk < K; ++kk) s += fa[(m + shift) * K + kk] * fb[n * K + kk];
        double e2 = fabs(s - ho[m * 64 + n]); if (e2 > maxerr) maxerr = e2;
      }
      printf("base_offset=%d shift=%3d max_err=%.5f %s\n", ubo, shift, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
    }
  }
  return 0;
}
