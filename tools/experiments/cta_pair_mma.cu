// Experiment: one tcgen05.mma.cta_group::2 tile (M = 256 over a CTA pair, N = 64..256, K = 64) with operands written to
// 128B-swizzled shared memory by ordinary stores, checked against a CPU product.  Establishes, on this hardware and
// toolchain, the exact forms of: cluster launch, tcgen05.alloc/dealloc.cta_group::2 (issued by one warp in EACH CTA),
// the leader-only MMA issue, the multicast commit onto both CTAs' mbarriers, and the operand split (each CTA holds its
// own 128 rows of A and HALF of B's N rows; each CTA's TMEM receives its 128 rows x all N columns).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../fastvideotagging_b200/csrc/ptx.cuh"
using namespace fvt;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ D, int n) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  uint8_t* a_tile = smem;                      // [128 rows][128 B]
  uint8_t* b_tile = smem + 128 * 128;          // [n/2 rows][128 B]
  const int nb = n / 2;
  // K-major, 128B swizzle: 16-byte unit j of row r lives at unit j ^ (r & 7)
  for (int i = threadIdx.x; i < 128 * 8; i += 128) {
    const int r = i >> 3, j = i & 7;
    *reinterpret_cast<uint4*>(a_tile + r * 128 + ((j ^ (r & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(A + (static_cast<size_t>(rank) * 128 + r) * 64 + j * 8);
  }
  for (int i = threadIdx.x; i < nb * 8; i += 128) {
    const int r = i >> 3, j = i & 7;
    *reinterpret_cast<uint4*>(b_tile + r * 128 + ((j ^ (r & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(B + (static_cast<size_t>(rank) * nb + r) * 64 + j * 8);
  }
  ptx::fence_proxy_async_smem();
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(&slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tm = slot;
  if (rank == 0 && warp == 1) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_bf16(256, n, 0, 0);
      const uint64_t ad = ptx::make_sw128_desc(ptx::smem_u32(a_tile), 16, 1024);
      const uint64_t bd = ptx::make_sw128_desc(ptx::smem_u32(b_tile), 16, 1024);
      for (int ks = 0; ks < 4; ++ks) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
            ::"r"(tm), "l"(ad + 2 * ks), "l"(bd + 2 * ks), "r"(idesc), "r"(ks > 0 ? 1u : 0u) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(ptx::smem_u32(&bar)), "h"(static_cast<uint16_t>(3)) : "memory");
    }
    __syncwarp();
  }
  ptx::mbar_wait(ptx::smem_u32(&bar), 0);
  ptx::tc_fence_after();
  for (int c = 0; c < n; c += 16) {
    uint32_t v[16];
    ptx::tmem_ld_32x32b_x16(tm + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    ptx::tmem_ld_wait();
    float* dst = D + (static_cast<size_t>(rank) * 128 + warp * 32 + lane) * n + c;
    for (int i = 0; i < 16; ++i) dst[i] = __uint_as_float(v[i]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    ptx::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u) : "memory");
  }
}

int main() {
  for (int n : {64, 144, 256}) {
    std::vector<__nv_bfloat16> hA(256 * 64), hB(n * 64);
    std::vector<float> fA(256 * 64), fB(n * 64);
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { float v = (rand() % 17 - 8) / 8.0f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { float v = (rand() % 13 - 6) / 4.0f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 256 * n * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, 256 * n * 4);
    const int smem = 128 * 128 + 128 * 128 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<<<2, 128, smem>>>(dA, dB, dD, n);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("n=%d CUDA error: %s\n", n, cudaGetErrorString(e)); return 1; }
    std::vector<float> hD(256 * n);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0; int bad = 0;
    for (int m = 0; m < 256; ++m)
      for (int j = 0; j < n; ++j) {
        double ref = 0;
        for (int kk = 0; kk < 64; ++kk) ref += (double)fA[m * 64 + kk] * fB[j * 64 + kk];
        double err = fabs(ref - hD[m * n + j]);
        if (!(err <= 1e-3)) { if (bad < 5) printf("  mismatch m=%d n=%d ref=%f got=%f\n", m, j, ref, hD[m * n + j]); ++bad; }
        if (err > worst) worst = err;
      }
    printf("cta_group::2 M=256 N=%d K=64: %s (max |err| %.3g, %d mismatches)\n", n, bad ? "MISMATCH" : "OK", worst, bad);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  return 0;
}
