// The "P16" activation layout of the eval-mode tensor-core CNN: every fp32 activation x is stored ALREADY SPLIT into the
// fp16 pair the tensor cores consume,  16 x = hi + lo  (hi = fp16(16 x), lo = fp16(16 x - hi); ~22 bits of x survive,
// |x| < 4094), as
//        [N][C8 = ceil(C/8)][part: hi, lo][H][W][8 halfs]          (4 bytes per element, like fp32)
// so a pixel x 8-channel chunk of one part is one 16-byte unit = one row of a K-major, un-swizzled UMMA core matrix: a TMA
// box of the tensor lands in shared memory as an MMA-ready operand (no split pass, no generic-proxy writes), and the
// producing layer's epilogue pays for the split once instead of every consumer's CTA paying for it per tile.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace gfr_p16 {

constexpr float X_SCALE = 16.0f;           // activations are multiplied by 2^4 before the split
constexpr float X_INV = 1.0f / 16.0f;
constexpr float X_LIMIT = 4094.0f;         // |x| * 16 must stay below fp16's 65504

// 8 fp32 values -> the two 16-byte units (hi, lo)
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float a = v[2 * k] * X_SCALE, b = v[2 * k + 1] * X_SCALE;
    const __half2 hh = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(a - hf.x, b - hf.y);
    h[k] = *reinterpret_cast<const uint32_t*>(&hh);
    l[k] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// the two 16-byte units -> 8 fp32 values
__device__ __forceinline__ void join8(const uint4& hi, const uint4& lo, float (&v)[8]) {
  const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h[k]));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&l[k]));
    v[2 * k] = (a.x + b.x) * X_INV;
    v[2 * k + 1] = (a.y + b.y) * X_INV;
  }
}

// the same two helpers in the x16 domain of the stored pair (v16 = 16 x = hi + lo): no scaling on either side
__device__ __forceinline__ void split8_x16(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __half2 hh = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(v[2 * k] - hf.x, v[2 * k + 1] - hf.y);
    h[k] = *reinterpret_cast<const uint32_t*>(&hh);
    l[k] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void join8_x16(const uint4& hi, const uint4& lo, float (&v)[8]) {
  const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h[k]));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&l[k]));
    v[2 * k] = a.x + b.x;
    v[2 * k + 1] = a.y + b.y;
  }
}

// element offset (in halfs) of the hi unit of (n, chunk c8, y, x); the lo unit is plane_halfs = H*W*8 further
__device__ __forceinline__ size_t unit_offset(int n, int groups, int c8, int H, int W, int y, int x) {
  return ((((size_t)n * groups + c8) * 2) * H + y) * (size_t)W * 8 + (size_t)x * 8;
}

}  // namespace gfr_p16
