// Philox4x32-10 counter-based RNG (Salmon et al., SC'11; Random123 constants).  One 128-bit counter + 64-bit key
// -> four uint32.  The WARP negative-sampling contract (oracle/mlc_loss.py: philox4x32_10) uses exactly this
// function on both sides, so sampled ranks are bit-identical between the CUDA kernel and the CPU oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fvt {

__host__ __device__ __forceinline__ void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
  const uint64_t p = static_cast<uint64_t>(a) * static_cast<uint64_t>(b);
  hi = static_cast<uint32_t>(p >> 32);
  lo = static_cast<uint32_t>(p);
}

__host__ __device__ __forceinline__ uint4 philox4x32_10_impl(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    philox_mulhilo(M0, c.x, hi0, lo0);
    philox_mulhilo(M1, c.z, hi1, lo1);
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) { return philox4x32_10_impl(c, k); }
inline uint4 philox4x32_10_host(uint4 c, uint2 k) { return philox4x32_10_impl(c, k); }

}  // namespace fvt
