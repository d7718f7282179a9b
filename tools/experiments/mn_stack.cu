// Experiment: MN-major SW128 A operand whose two 64-channel MN atoms are the SAME TMA-written pixel slab seen through two
// different filter-tap shifts (LBO = byte distance between the taps), i.e. one M=128 UMMA computes the weight-gradient
// rows of two taps at once:   D[g*64 + c, n] = sum_p X[p + s_g][c] * dY[p][n],  g in {0,1}.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../fastvideotagging_b200/csrc/ptx.cuh"
using namespace fvt;

__global__ void __launch_bounds__(128, 1)
k(const __grid_constant__ CUtensorMap tma, const __grid_constant__ CUtensorMap tmb, float* out, int s1, int s2) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                 // 272 rows * 128 B
  uint8_t* sb = smem + 36864;         // 128 rows * 128 B
  uint64_t* bar = (uint64_t*)(smem + 36864 + 16384);
  uint64_t* bar2 = bar + 1;
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(bar), 1); ptx::mbar_init(ptx::smem_u32(bar2), 1); ptx::fence_mbar_init(); }
  if (warp == 1) { ptx::tmem_alloc(ptx::smem_u32(slot), 64); ptx::tmem_relinquish(); }
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(ptx::smem_u32(bar), 272 * 128 + 128 * 128);
    ptx::tma_load_2d(ptx::smem_u32(sa), &tma, ptx::smem_u32(bar), 0, 0);
    ptx::tma_load_2d(ptx::smem_u32(sa + 136 * 128), &tma, ptx::smem_u32(bar), 0, 136);
    ptx::tma_load_2d(ptx::smem_u32(sb), &tmb, ptx::smem_u32(bar), 0, 0);
    ptx::mbar_wait(ptx::smem_u32(bar), 0);
    ptx::tc_fence_after();
    const uint32_t a_addr = ptx::smem_u32(sa) + s1 * 128;
    const uint64_t ad = ptx::make_sw128_desc(a_addr, (uint32_t)(s2 - s1) * 128u, 1024);
    const uint64_t bd = ptx::make_sw128_desc(ptx::smem_u32(sb), 8192, 1024);
    const uint32_t idesc = ptx::make_idesc_bf16(128, 64, 1, 1);
    for (int kk = 0; kk < 8; ++kk) ptx::umma_bf16_ss(tm, ad + 128 * kk, bd + 128 * kk, idesc, kk > 0);
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
  const int RA = 272, RB = 128, K = 64;
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
  cuuint32_t boxa[2] = {64, 136}, boxb[2] = {64, 128}, es[2] = {1, 1};
  CUresult r1 = enc(&tma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, dimsa, str, boxa, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&tmb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, dimsb, str, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d %d\n", (int)r1, (int)r2);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> ho(128 * 64);
  const int pairs[][2] = {{0, 64}, {0, 8}, {0, 1}, {1, 2}, {2, 58}, {59, 60}, {3, 118}, {0, 0}, {5, 5}, {7, 16}, {60, 116}, {117, 118}};
  int bad = 0;
  for (auto& pr : pairs) {
    const int s1 = pr[0], s2 = pr[1];
    cudaMemset(dout, 0, 128 * 64 * 4);
    k<<<1, 128, 64 * 1024>>>(tma, tmb, dout, s1, s2);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("s=(%d,%d) CUDA error %s\n", s1, s2, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(ho.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
      const int sh = m < 64 ? s1 : s2;
      double s = 0; for (int p = 0; p < 128; ++p) s += fa[(p + sh) * K + (m & 63)] * fb[p * K + n];
      double e2 = fabs(s - ho[m * 64 + n]); if (e2 > maxerr) maxerr = e2;
    }
    printf("taps shift=(%3d,%3d) max_err=%.5f %s\n", s1, s2, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
    if (maxerr >= 1e-3) ++bad;
  }
  return bad ? 2 : 0;
}
