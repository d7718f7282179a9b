// Exact, order-independent accumulation of fp32 partial sums across CTAs.
//
// Training-mode BatchNorm needs per-channel sums over the whole activation (sum x, sum x^2 in the forward pass, sum dz,
// sum dz*xhat in the backward pass).  Every CTA reduces its share in a fixed order and then has to meet the other CTAs
// in global memory.  fp32 atomics do that in arrival order, so the low bits of the result change from run to run, and
// through bf16 rounding flips and the chaotic amplification of the 69-layer backward pass the weight gradients of two
// identical steps differed by 2 % (round 1).  Here a partial is added EXACTLY instead: its 24-bit mantissa is placed
// into a 128-bit fixed-point window made of four signed 64-bit limbs that carry 32 payload bits each (limb k weighs
// 2^(32k - 64)); the two limbs a mantissa straddles receive one integer atomic each.  Integer addition is associative,
// so the limbs — and the double they are read back as — do not depend on the order in which CTAs arrive.
//
//   representable: |v| < 2^55 (larger / non-finite values poison the accumulator: it reads back NaN);
//   resolution:    2^-64 absolute (bits below are truncated, identically on every run);
//   capacity:      2^21 additions per accumulator before a limb could leave the exact range of a double.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fvt {

constexpr int kDetLimbs = 4;                 // unsigned long long per accumulator (32 bytes)

__device__ __forceinline__ void det_add(unsigned long long* acc, float v) {
  const uint32_t u = __float_as_uint(v);
  uint32_t ex = (u >> 23) & 0xffu;
  uint32_t man = u & 0x7fffffu;
  if (ex == 0u) {
    if (man == 0u) return;                   // +-0 adds nothing
    ex = 1u;                                 // subnormal: no hidden bit
  } else {
    man |= 0x800000u;
  }
  // value = man * 2^(ex - 150); bit position of the mantissa's LSB inside the window (LSB weight 2^-64)
  int b = static_cast<int>(ex) - 150 + 64;
  if (ex == 0xffu || b > 95) {               // inf / nan / |v| >= 2^55
    atomicAdd(acc + 3, 1ull << 62);
    return;
  }
  if (b < 0) {
    man = b <= -24 ? 0u : (man >> (-b));
    b = 0;
    if (man == 0u) return;
  }
  const int k = b >> 5, s = b & 31;
  const unsigned long long wide = static_cast<unsigned long long>(man) << s;      // < 2^55
  unsigned long long lo = wide & 0xffffffffull, hi = wide >> 32;
  if (u >> 31) { lo = 0ull - lo; hi = 0ull - hi; }                                // two's complement: signed limbs
  if (lo != 0ull) atomicAdd(acc + k, lo);
  if (hi != 0ull) atomicAdd(acc + k + 1, hi);
}

__device__ __forceinline__ double det_read(const unsigned long long* acc) {
  const long long l3 = static_cast<long long>(acc[3]);
  if (l3 >= (1ll << 61) || l3 <= -(1ll << 61)) return __longlong_as_double(0x7ff8000000000000ll);
  double r = static_cast<double>(l3) * 4294967296.0;                              // 2^32
  r += static_cast<double>(static_cast<long long>(acc[2]));
  r += static_cast<double>(static_cast<long long>(acc[1])) * (1.0 / 4294967296.0);
  r += static_cast<double>(static_cast<long long>(acc[0])) * (1.0 / 4294967296.0 / 4294967296.0);
  return r;
}

// Flush of a convolution CTA's per-channel statistics: `part` holds [4 quadrants][2][n_pad] fp32 partials (see
// epilogue.cuh); the quadrants are combined in a fixed order and the CTA's total for the channels [ch_lo, ch_hi) is
// added exactly into stats[quantity * cout_store + channel].  Called by `nthreads` threads (tid = 0 .. nthreads-1)
// after a barrier.
__device__ __forceinline__ void flush_quadrant_stats(const float* part, int n_pad, int cout_store, unsigned long long* stats,
                                                     int tid, int nthreads, int ch_lo, int ch_hi) {
  const int span = ch_hi - ch_lo;
  for (int i = tid; i < 2 * span; i += nthreads) {
    const int qty = i >= span ? 1 : 0;
    const int ch = ch_lo + i - qty * span;
    const float* s = part + qty * n_pad + ch;
    const float v = (s[0] + s[2 * n_pad]) + (s[4 * n_pad] + s[6 * n_pad]);
    det_add(stats + (static_cast<size_t>(qty) * cout_store + ch) * kDetLimbs, v);
  }
}
__device__ __forceinline__ void flush_quadrant_stats(const float* part, int n_pad, int cout_store, unsigned long long* stats,
                                                     int tid, int nthreads) {
  flush_quadrant_stats(part, n_pad, cout_store, stats, tid, nthreads, 0, cout_store);
}

}  // namespace fvt
