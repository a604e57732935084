// Hard-shadow ray-march (forward) for sm_100a.
//
// Replaces the per-sample python loop of the reference, TRAIN:374-515 / TEST1:351-496
// (TRAIN = train_raytracing_relighting_CelebAHQ_DSSIM_8x.py, TEST1 = test_relight_single_image.py).
//
// Mapping: one THREAD per pixel-ray, a CTA owns a 32x8 pixel tile and all of its threads walk the
// samples k = 0..n-1 in lock-step.  The n sample parameters travel as a kernel parameter (constant
// bank -> uniform loads).  Sample positions and the bilinear weights are computed in fp64 with the
// reference's exact operation order (no FMA contraction): the integer decisions (nearest pixel for the
// face-mask test, floor/ceil for the bilinear footprint) are discontinuous, so fp32 positions flip
// ~1e-4 of the pixels.  Everything after `.float()` in the reference (TRAIN:502) is fp32 here too.
// The face mask is 1 bit/pixel in shared memory (8 KB for 256x256).
#include "gfr_common.cuh"
#include "shade_device.cuh"

#include <math.h>
#include <stdlib.h>

namespace {

struct SampleTable { double t[GFR_MAX_SAMPLES]; };

struct MarchArgs {
  const float* depth;        // [B,H,W]
  const uint32_t* mask_bits; // [1|B, H*W/32]
  const float* light;        // [B,3]
  float* dmin;               // [B,H,W]
  uint8_t* argmin;           // [B,H,W] or null
  float* shadow;             // [B,H,W] or null
  int mask_stride;           // words
  int B, H, W, n;
  int lpf;                   // lights per face: (face, light) pair b reads depth / mask of face b / lpf
  float t0, inv_dt;          // uniform sample table t_k = t0 + k*dt (inv_dt = 0: not uniform, no sample-range culling)
  int order;                 // fast kernel: block order, 0 = tile-major (round 1), 1 = pairs interleaved, far-from-light tiles first
  int cut;                   // 1: fast kernel, sample groups: exact early cut-off of a ray's sample range (needs `range`)
  const int* range;          // [faces][2] ordered-int keys written by the widening pre-pass: max depth, max -depth over the dilated mask
  int coarse;                // 1: the fast kernel builds the 8x8-block occupancy map and skips empty groups of 4 samples
  int fuse_shade;            // 1: normals + Lambert + blend + render of the pixel follow in the same thread (shade)
  gfr_shade::ShadeArgs shade;
  float bonus;
  double hWd, hHd, neg_eps;   // W/2, H/2 and -1e-4 as fp64 kernel parameters: constant-bank operands of the per-sample DADDs
  float bx0, bx1, by0, by1;   // the light must project into [bx0,bx1] x [by0,by1] for the bonus (TEST1:495 / TEST_LT:503)
};

constexpr int TILE_W = 32, TILE_H = 8;
#ifndef GFR_MARCH_DEFAULT_ILP
#define GFR_MARCH_DEFAULT_ILP 2
#endif
#ifndef GFR_MARCH_ILP2_BLOCKS
#define GFR_MARCH_ILP2_BLOCKS 6      // resident 128-thread CTAs per SM the paired-sample kernel is compiled for (80 registers)
#endif

// End point of the 2-D ray (pixel -> projected light) on the image rectangle, fp32, reference op order.
// TRAIN:378-465.  Rectangle: x in [-W/2, W/2-1], y in [1-H/2, H/2].
__device__ __forceinline__ void ray_end(float x, float y, float Lx, float Ly, float xmin, float xmax,
                                        float ymin, float ymax, float& ex, float& ey) {
  const float m = __fdiv_rn(__fsub_rn(Ly, y), __fadd_rn(__fsub_rn(Lx, x), 1e-4f));   // TRAIN:378
  const float b = __fsub_rn(Ly, __fmul_rn(m, Lx));                                    // TRAIN:379
  const int sx = Lx < xmin ? -1 : (Lx <= xmax ? 0 : 1);
  const int sy = Ly < ymin ? -1 : (Ly <= ymax ? 0 : 1);
  const float xe = sx < 0 ? xmin : xmax;
  const float ye = sy < 0 ? ymin : ymax;
  const float exy = __fadd_rn(__fmul_rn(m, xe), b);                                   // y on the x edge
  const float eyx = __fdiv_rn(__fsub_rn(ye, b), __fadd_rn(m, 1e-4f));                 // x on the y edge
  if (sx != 0 && sy != 0) {
    const bool hit = (eyx >= xmin) && (eyx <= xmax);                                  // TRAIN:397-398
    ex = hit ? eyx : xe;
    ey = hit ? ye : exy;
  } else if (sx != 0) {
    ex = xe; ey = exy;
  } else if (sy != 0) {
    ex = eyx; ey = ye;
  } else {
    ex = Lx; ey = Ly;                                                                 // TRAIN:423-425
  }
  ex = ex < xmin ? xmin : ex;  ex = ex > xmax ? xmax : ex;                            // TRAIN:462-465
  ey = ey < ymin ? ymin : ey;  ey = ey > ymax ? ymax : ey;
}

__device__ __forceinline__ float shadow_weight(float d) {   // TRAIN:517
  const float e = expf(-d);
  const float op = 1.0f + e;
  return __fadd_rn(__fdiv_rn(__fmul_rn(-4.0f, e), __fmul_rn(op, op)), 1.0f);
}

// ---------------------------------------------------------------------------------------------------
// Variant 1: depth gathers straight from global memory through L1 (read-only path).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE_W * TILE_H)
shadow_march_fwd_l1(const MarchArgs a, const __grid_constant__ SampleTable tab) {
  extern __shared__ uint32_t s_mask[];
  const int b = blockIdx.z, f = b / a.lpf;
  const int H = a.H, W = a.W;
  const int words = (H * W) >> 5;
  {
    const uint32_t* src = a.mask_bits + (size_t)f * a.mask_stride;
    for (int i = threadIdx.y * TILE_W + threadIdx.x; i < words; i += TILE_W * TILE_H) s_mask[i] = __ldg(src + i);
  }
  __syncthreads();

  const int col = blockIdx.x * TILE_W + threadIdx.x;
  const int row = blockIdx.y * TILE_H + threadIdx.y;
  const float* __restrict__ D = a.depth + (size_t)f * H * W;
  const float halfW = 0.5f * W, halfH = 0.5f * H;
  const float xmin = -halfW, xmax = W - halfW - 1.0f, ymin = 1.0f - halfH, ymax = halfH;
  const float x = (float)col - halfW;            // TRAIN:52
  const float y = halfH - (float)row;            // TRAIN:53
  const float z = __ldg(D + row * W + col);
  const float Lx = __ldg(a.light + 3 * b), Ly = __ldg(a.light + 3 * b + 1), Lz = __ldg(a.light + 3 * b + 2);

  float ex, ey;
  ray_end(x, y, Lx, Ly, xmin, xmax, ymin, ymax, ex, ey);
  const double dx = (double)__fsub_rn(ex, x), dy = (double)__fsub_rn(ey, y);          // TRAIN:467
  const double xd = (double)x, yd = (double)y;
  const double hW = (double)halfW, hH = (double)halfH, neg_eps = -0.0001;
  const float bcx = __fsub_rn(Lx, x), bcy = __fsub_rn(Ly, y), bcz = __fsub_rn(Lz, z); // BC, TRAIN:507

  float qmin = __int_as_float(0x7f800000);   // +inf == "outside the face"
  int kmin = 255;
#pragma unroll 2
  for (int k = 0; k < a.n; ++k) {
    const double t = tab.t[k];
    const double px = __dadd_rn(xd, __dmul_rn(t, dx));                                // TRAIN:472,480
    const double py = __dadd_rn(yd, __dmul_rn(t, dy));
    const int ci = __double2int_rn(px) + (W >> 1);                                    // TRAIN:472-475
    const int ri = (H >> 1) - __double2int_rn(py);
    const int mi = ri * W + ci;
    const bool inside = (s_mask[mi >> 5] >> (mi & 31)) & 1u;                          // TRAIN:510
    if (!inside) continue;        // the reference computes the sample and then overwrites it with 1e6 (TRAIN:510-512)
    const double u = __dadd_rn(__dadd_rn(px, hW), neg_eps);                           // TRAIN:481,483
    const double v = __dadd_rn(__dsub_rn(hH, py), neg_eps);                           // TRAIN:482,483
    const int uf = __double2int_rd(u), uc = __double2int_ru(u);                       // TRAIN:486-487
    const int vf = __double2int_rd(v), vc = __double2int_ru(v);
    const int ufi = uf < 0 ? uf + W : uf, vfi = vf < 0 ? vf + H : vf;                 // python negative index
    const double wu0 = __dsub_rn((double)uc, u), wu1 = __dsub_rn(u, (double)uf);
    const double wv0 = __dsub_rn((double)vc, v), wv1 = __dsub_rn(v, (double)vf);
    const double ul = (double)__ldg(D + vfi * W + ufi), ur = (double)__ldg(D + vfi * W + uc);
    const double ll = (double)__ldg(D + vc * W + ufi), lr = (double)__ldg(D + vc * W + uc);
    const double up = __dadd_rn(__dmul_rn(ul, wu0), __dmul_rn(ur, wu1));              // TRAIN:492
    const double lo = __dadd_rn(__dmul_rn(ll, wu0), __dmul_rn(lr, wu1));              // TRAIN:493
    const double zi = __dadd_rn(__dmul_rn(up, wv0), __dmul_rn(lo, wv1));              // TRAIN:494
    const float ax = (float)__dsub_rn(u, hW), ay = (float)__dsub_rn(hH, v), az = (float)zi;   // TRAIN:498-502
    const float bax = __fsub_rn(ax, x), bay = __fsub_rn(ay, y), baz = __fsub_rn(az, z);
    const float c0 = __fsub_rn(__fmul_rn(bay, bcz), __fmul_rn(baz, bcy));             // TRAIN:508
    const float c1 = __fsub_rn(__fmul_rn(baz, bcx), __fmul_rn(bax, bcz));
    const float c2 = __fsub_rn(__fmul_rn(bax, bcy), __fmul_rn(bay, bcx));
    const float q = __fadd_rn(__fadd_rn(__fmul_rn(c0, c0), __fmul_rn(c1, c1)), __fmul_rn(c2, c2));
    if (q < qmin) { qmin = q; kmin = k; }
  }
  // min_k sqrt(q_k + eps)/den == sqrt(min_k q_k + eps)/den (monotone), TRAIN:509,514
  float d;
  if (kmin == 255) {
    d = 1000000.0f;                                                                   // TRAIN:512
  } else {
    const float den = sqrtf(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(bcx, bcx), __fmul_rn(bcy, bcy)), __fmul_rn(bcz, bcz)), 1e-4f));
    d = __fdiv_rn(sqrtf(__fadd_rn(qmin, 1e-4f)), den);
  }
  if (a.bonus != 0.0f && Lx >= a.bx0 && Lx <= a.bx1 && Ly >= a.by0 && Ly <= a.by1) d = __fadd_rn(d, a.bonus);   // TEST1:495-496
  const size_t o = (size_t)b * H * W + row * W + col;
  a.dmin[o] = d;
  if (a.argmin) a.argmin[o] = (uint8_t)kmin;
  if (a.shadow) a.shadow[o] = shadow_weight(d);
}

// ---------------------------------------------------------------------------------------------------
// Variant 0 (default): the same arithmetic with the int <-> fp64 and fp32 -> fp64 conversions taken off the XU pipe.
// ncu on variant 1: XU (conversion) pipe 81 % busy, fp64 pipe 22 % — the 17 F2I/I2F/F2F per sample were the bound.
//   * round-half-even, floor and ceil come from one fp64 add of 1.5 * 2^52: the sum's low word IS the integer and
//     (sum - magic) is its fp64 value, so no F2I and no I2F is issued (exact for |v| < 2^51; same ties as np.round);
//   * the depth map is widened to fp64 once per launch by a pre-pass (depth64), so the four gathers per sample need
//     no F2F; the only conversions left are the three fp64 -> fp32 casts of TRAIN:502.
// ---------------------------------------------------------------------------------------------------
constexpr double kMagic = 6755399441055744.0;   // 2^52 + 2^51

__device__ __forceinline__ void round_parts(double v, int& r, double& rd) {
  const double s = __dadd_rn(v, kMagic);
  r = __double2loint(s);
  rd = __dsub_rn(s, kMagic);
}

// The in-mask part of one sample (TRAIN:481-509): bilinear depth at the sample in fp64, then the fp32 squared distance of the
// surface point from the pixel -> light line (times |BC|^2).  Every address is inside the image for every walked sample (the
// ray is clipped to the image rectangle), so callers may evaluate it speculatively for out-of-mask samples and discard q.
struct RayConst {
  double hW, hH, neg_eps;
  float x, y, z, bcx, bcy, bcz;
  int W, H;
};

__device__ __forceinline__ float sample_q(const double* D, const RayConst& r, double px, double py) {
  const double u = __dadd_rn(__dadd_rn(px, r.hW), r.neg_eps);                         // TRAIN:481,483
  const double v = __dadd_rn(__dsub_rn(r.hH, py), r.neg_eps);                         // TRAIN:482,483
  // floor / ceil: the magic add in round-down / round-up mode (TRAIN:486-487)
  const double sfu = __dadd_rd(u, kMagic), scu = __dadd_ru(u, kMagic);
  const double sfv = __dadd_rd(v, kMagic), scv = __dadd_ru(v, kMagic);
  const int uf = __double2loint(sfu), uc = __double2loint(scu);
  const int vf = __double2loint(sfv), vc = __double2loint(scv);
  const double ufd = __dsub_rn(sfu, kMagic), ucd = __dsub_rn(scu, kMagic);
  const double vfd = __dsub_rn(sfv, kMagic), vcd = __dsub_rn(scv, kMagic);
  const unsigned ufi = uf < 0 ? uf + r.W : uf, vfi = vf < 0 ? vf + r.H : vf;          // python negative index
  const double wu0 = __dsub_rn(ucd, u), wu1 = __dsub_rn(u, ufd);
  const double wv0 = __dsub_rn(vcd, v), wv1 = __dsub_rn(v, vfd);
  // four corners from ONE 64-bit address: the other three are 32-bit element offsets (one IMAD.WIDE each)
  const double* p00 = D + (vfi * (unsigned)r.W + ufi);
  const int du = uc - (int)ufi, dv = (vc - (int)vfi) * r.W;                           // +1 / 0, or -(W-1) / -(H-1)*W on a wrap
  const double* p10 = p00 + dv;
  const double ul = __ldg(p00), ur = __ldg(p00 + du);
  const double ll = __ldg(p10), lr = __ldg(p10 + du);
  const double up = __dadd_rn(__dmul_rn(ul, wu0), __dmul_rn(ur, wu1));                // TRAIN:492
  const double lo = __dadd_rn(__dmul_rn(ll, wu0), __dmul_rn(lr, wu1));                // TRAIN:493
  const double zi = __dadd_rn(__dmul_rn(up, wv0), __dmul_rn(lo, wv1));                // TRAIN:494
  const float ax = (float)__dsub_rn(u, r.hW), ay = (float)__dsub_rn(r.hH, v), az = (float)zi;   // TRAIN:498-502
  const float bax = __fsub_rn(ax, r.x), bay = __fsub_rn(ay, r.y), baz = __fsub_rn(az, r.z);
  const float c0 = __fsub_rn(__fmul_rn(bay, r.bcz), __fmul_rn(baz, r.bcy));           // TRAIN:508
  const float c1 = __fsub_rn(__fmul_rn(baz, r.bcx), __fmul_rn(bax, r.bcz));
  const float c2 = __fsub_rn(__fmul_rn(bax, r.bcy), __fmul_rn(bay, r.bcx));
  return __fadd_rn(__fadd_rn(__fmul_rn(c0, c0), __fmul_rn(c1, c1)), __fmul_rn(c2, c2));
}

// TH = rows of the CTA tile (32 x TH pixels, TH warps).  TH = 8: 256-thread CTAs, 4 per SM.  TH = 4: 128-thread CTAs (8 per
// SM) that also fit beside two resident tcgen05 conv CTAs (2 x 320 threads x 84 registers + 175 KB shared memory leave
// room for 128 x 62 registers + 8 KB), so the issue-bound march of one runner lane can share SMs with the shared-memory /
// tensor-bound convolutions of another.
//
// WS = warp shape.  0: a warp is 32 x 1 pixels of the tile (round 1).  1: a warp is 8 x 4 pixels (warp w of the CTA owns columns
// 8w .. 8w+7 of the 32 x 4 tile; TH = 4 only): the 32 rays of a warp start closer together, so their sample-range union is
// shorter and they cross the face-mask boundary at more nearly the same k (tools/sim_march_warp_shape.py on the bench masks:
// 85.8 -> 79.1 walked and 68.6 -> 62.3 in-mask warp-iterations per warp, -9 % instructions), and a gather touches 4 x 64 B
// instead of 1 x 256 B of each of its two rows.
// ILP = 2 walks the samples in pairs: both positions / mask tests first, then both in-mask bodies back to back in one basic
// block (the second is speculative where only one of the pair is inside the face; every address is valid), so the scheduler has
// two independent fp64 dependency chains per warp to hide the fixed-latency stalls that dominated round 1's stall sampling
// (`wait` 2.9 and `not selected` 2.2 warps per issue at 6.6 resident warps per scheduler).  Same arithmetic per sample, same
// first-minimum rule: bit-identical results (tests/test_gpu_march_edges.py compares every variant with the literal kernel).
template <int TH, int WS, int ILP>
__global__ void __launch_bounds__(TILE_W * TH, ILP == 2 ? GFR_MARCH_ILP2_BLOCKS : (ILP == 3 ? 5 : (ILP == 4 ? 4 : 1024 / (TILE_W * TH))))
shadow_march_fwd_fast(const MarchArgs a, const double* __restrict__ depth64, const __grid_constant__ SampleTable tab) {
  constexpr int TILE_H = TH;
  static_assert(WS == 0 || TH == 4, "the 8 x 4 warp shape is laid out for 32 x 4 CTA tiles");
  extern __shared__ uint32_t s_mask[];
  const int H = a.H, W = a.W;
  // Block order (1-D grid).  a.order = 0: tile column fastest, then tile row, then the (face, light) pair — round 1's 3-D grid.
  // a.order = 1: the pair index runs fastest and every pair walks ITS tiles starting on the side of the image that is far from
  // its light, along the axis the light direction is closer to.  A ray marches towards the light: the far side's rays cross
  // the whole face (160 in-mask samples), the near side's leave the image after a few, so a CTA's cost varies 1 : 700 across
  // the image and round 1's order ended with one face's 512 tiles, the expensive ones included, on a half-empty GPU
  // (tools/sim_march_order.py: makespan 1.23x the ideal; far-side-first with the pairs interleaved: 1.06x).
  int b, bx, by;
  {
    const int tiles_x = W / TILE_W, tiles_y = H / TILE_H;
    const int id = blockIdx.x;
    if (a.order == 0) {
      bx = id % tiles_x; by = (id / tiles_x) % tiles_y; b = id / (tiles_x * tiles_y);
    } else {
      b = id % a.B;
      const int i = id / a.B;
      const float lx = __ldg(a.light + 3 * b), ly = __ldg(a.light + 3 * b + 1);
      if (fabsf(lx) >= fabsf(ly)) {                 // light to the left / right: tile columns slowest, far column first
        bx = i / tiles_y; by = i % tiles_y;
        if (!(lx > 0.f)) bx = tiles_x - 1 - bx;
      } else {                                      // light above / below (+y is up = small rows): far row first
        by = i / tiles_x; bx = i % tiles_x;
        if (ly > 0.f) by = tiles_y - 1 - by;
      }
    }
  }
  const int f = b / a.lpf;
  const int words = (H * W) >> 5;
  {
    const uint32_t* src = a.mask_bits + (size_t)f * a.mask_stride;
    for (int i = threadIdx.y * TILE_W + threadIdx.x; i < words; i += TILE_W * TILE_H) s_mask[i] = __ldg(src + i);
  }
  __syncthreads();
  // Coarse occupancy (a.coarse): one bit per 8x8-pixel block, set iff the block OR ANY OF ITS 8 NEIGHBOURS holds a mask pixel.
  // A group of 4 consecutive samples stays within 3.2 pixels of its middle (1.5 dt x ray length <= 724 px at 512^2, + 0.5 for the
  // nearest-pixel rounding), i.e. inside the 3x3 block neighbourhood of the middle's block: a clear bit proves that all 4
  // nearest pixels are outside the face, so the warp skips the group (their distance is 1e6 anyway, TRAIN:510-512).
  uint32_t* s_occ = s_mask + words;              // (H/8) x (W/8) bits, rows of W/256-rounded-up words... one word per 32 blocks
  const int cbw = W >> 3, cbh = H >> 3, cwords = (cbw * cbh + 31) >> 5;
  if (a.coarse) {
    uint32_t* s_raw = s_occ + cwords;
    const int tid = threadIdx.y * TILE_W + threadIdx.x;
    for (int i = tid; i < cwords; i += TILE_W * TILE_H) { s_occ[i] = 0u; s_raw[i] = 0u; }
    __syncthreads();
    for (int i = tid; i < cbw * cbh; i += TILE_W * TILE_H) {
      const int by = i / cbw, bx = i % cbw;
      uint32_t any = 0u;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int bit0 = (by * 8 + r) * W + bx * 8;             // W % 32 == 0: the 8 bits sit in one word
        any |= (s_mask[bit0 >> 5] >> (bit0 & 31)) & 0xFFu;
      }
      if (any) atomicOr(&s_raw[i >> 5], 1u << (i & 31));
    }
    __syncthreads();
    for (int i = tid; i < cbw * cbh; i += TILE_W * TILE_H) {
      const int by = i / cbw, bx = i % cbw;
      uint32_t any = 0u;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx2 = -1; dx2 <= 1; ++dx2) {
          const int yy = by + dy, xx = bx + dx2;
          if (yy >= 0 && yy < cbh && xx >= 0 && xx < cbw) { const int j = yy * cbw + xx; any |= (s_raw[j >> 5] >> (j & 31)) & 1u; }
        }
      if (any) atomicOr(&s_occ[i >> 5], 1u << (i & 31));
    }
    __syncthreads();
  }

  const int col = WS == 0 ? bx * TILE_W + threadIdx.x : bx * TILE_W + threadIdx.y * 8 + (threadIdx.x & 7);
  const int row = WS == 0 ? by * TILE_H + threadIdx.y : by * TILE_H + (threadIdx.x >> 3);
  const double* D = depth64 + (size_t)f * H * W;
  asm volatile("" : "+l"(D));               // one 64-bit base register pair: every gather address is then a single IMAD.WIDE
  const float halfW = 0.5f * W, halfH = 0.5f * H;
  const float xmin = -halfW, xmax = W - halfW - 1.0f, ymin = 1.0f - halfH, ymax = halfH;
  const float x = (float)col - halfW;            // TRAIN:52
  const float y = halfH - (float)row;            // TRAIN:53
  const float z = __ldg(a.depth + (size_t)f * H * W + row * W + col);
  const float Lx = __ldg(a.light + 3 * b), Ly = __ldg(a.light + 3 * b + 1), Lz = __ldg(a.light + 3 * b + 2);

  float ex, ey;
  ray_end(x, y, Lx, Ly, xmin, xmax, ymin, ymax, ex, ey);
  const double dx = (double)__fsub_rn(ex, x), dy = (double)__fsub_rn(ey, y);          // TRAIN:467
  double xd = (double)x, yd = (double)y;
  asm volatile("" : "+d"(xd), "+d"(yd));     // keep them in registers: ptxas otherwise re-converts x, y every sample (2 XU ops)
  const double hW = a.hWd, hH = a.hHd, neg_eps = a.neg_eps;   // (ptxas re-derived (double)(0.5f*W) per sample: 2 I2F + 2 F2F)
  const float bcx = __fsub_rn(Lx, x), bcy = __fsub_rn(Ly, y), bcz = __fsub_rn(Lz, z); // BC, TRAIN:507
  const int cW = W >> 1, cH = H >> 1;

  // Sample-range culling (exact): a sample can only pass the face-mask test if its nearest pixel lies in the bounding box
  // of the mask, i.e. for t in an interval that is solved here in fp32 with a 0.01-pixel / one-sample safety margin; the
  // warp walks the union of its 32 intervals (uniform bounds, uniform table loads), everything outside is 1e6 anyway.
  int k_begin = 0, k_end = a.n - 1;
  int k_hi = a.n - 1;                            // this lane's own last useful sample (bounding box, then the cut-off)
  if (a.inv_dt != 0.f) {
    const int* bb = reinterpret_cast<const int*>(a.mask_bits + (size_t)f * a.mask_stride + words);
    const int c_lo = __ldg(bb), c_hi = -__ldg(bb + 1), r_lo = __ldg(bb + 2), r_hi = -__ldg(bb + 3);
    float tlo = 0.f, thi = 1.f;
    bool none = c_lo > c_hi;
    const float xa = (float)(c_lo - cW) - 0.51f, xb = (float)(c_hi - cW) + 0.51f;
    const float ya = (float)(cH - r_hi) - 0.51f, yb = (float)(cH - r_lo) + 0.51f;
    const float dxf = __fsub_rn(ex, x), dyf = __fsub_rn(ey, y);
    if (fabsf(dxf) > 1e-6f) {
      const float t1 = (xa - x) / dxf, t2 = (xb - x) / dxf;
      tlo = fmaxf(tlo, fminf(t1, t2)); thi = fminf(thi, fmaxf(t1, t2));
    } else if (x < xa || x > xb) none = true;
    if (fabsf(dyf) > 1e-6f) {
      const float t1 = (ya - y) / dyf, t2 = (yb - y) / dyf;
      tlo = fmaxf(tlo, fminf(t1, t2)); thi = fminf(thi, fmaxf(t1, t2));
    } else if (y < ya || y > yb) none = true;
    int kl = (int)floorf((tlo - a.t0) * a.inv_dt) - 1, kh = (int)ceilf((thi - a.t0) * a.inv_dt) + 1;
    if (none || tlo > thi) { kl = a.n; kh = -1; }
    k_begin = max(__reduce_min_sync(0xffffffffu, kl), 0);
    k_end = min(__reduce_max_sync(0xffffffffu, kh), a.n - 1);
    k_hi = kh;
  }

  // Exact early cut-off (a.cut; round 2).  With BA = A_k - P and BC = P_L - P (TRAIN:507-508),
  //     |BA x BC|^2 >= (BCx^2 + BCy^2) * (BAz - h_k)^2,   h_k = BCz * (BAx BCx + BAy BCy) / (BCx^2 + BCy^2)
  // (h_k = height of the pixel -> light line above the pixel at the sample's (x, y); Cauchy-Schwarz on the rest).  h_k grows
  // linearly with t_k (the 2-D ray points at the projected light) while BAz = az - z can never exceed zmax - z, the largest
  // depth any in-mask sample's bilinear footprint can return (`range`, reduced over the 1-pixel dilation of the mask by the
  // widening pre-pass, 0 included for the reference's zero-weight case at integral u / v).  Once the line has risen above
  // that by more than the current minimum distance, no later sample can beat it: the lane's last useful sample index is
  //     k_cut = ceil((ca + cb * sqrt(qmin) - t0) / dt) + 1
  // and the warp stops at the maximum over its lanes.  Margins: 4 E on sqrt(q) (E bounds the fp32 rounding of the reference's
  // cross product: products of magnitude <= (|BCz| + |BCx| + |BCy|) (W + H) + (|z range| + |z|) (|BCx| + |BCy|), 2^-21 relative, > 4x the worst case),
  // 2e-4 (|BCx| + |BCy|) on the scalar product (the -1e-4 / +1e-4 index offsets and the fp32 rounding of A_k), 1e-5 relative
  // and one whole sample on t.  Rays that descend (BCz < 0) use zmin the same way.  Skipped samples can only be >= the
  // running minimum, so d_min AND the arg-min (first minimum) are unchanged: bit-identical to the literal kernel.
  float cut_ka = 0.f, cut_kb = 0.f;
  bool cut_ok = false;
  if (ILP >= 2 && a.cut && a.inv_dt != 0.f) {
    const int* rg = a.range + 2 * f;
    const int kmax = __ldg(rg), kmin = __ldg(rg + 1);
    const float zmax = fmaxf(__int_as_float(kmax >= 0 ? kmax : kmax ^ 0x7fffffff), 0.f);
    const float zmin = fminf(-__int_as_float(kmin >= 0 ? kmin : kmin ^ 0x7fffffff), 0.f);
    const float dxf = __fsub_rn(ex, x), dyf = __fsub_rn(ey, y);
    const float A2 = bcx * bcx + bcy * bcy;
    const float p1 = dxf * bcx, p2 = dyf * bcy, S1 = p1 + p2;
    const float nxy = fabsf(bcx) + fabsf(bcy);
    const float G = bcz > 0.f ? zmax - z : z - zmin;                        // room between the pixel and the extreme depth ahead of the line
    if (S1 > 1e-3f * (fabsf(p1) + fabsf(p2)) && S1 > 0.f && A2 > 0.f && bcz != 0.f && G >= 0.f && zmax >= zmin) {
      const float E = 4.76837158e-7f * ((fabsf(bcz) + nxy) * (float)(W + H) + 2.f * ((zmax - zmin) + fmaxf(fmaxf(fabsf(zmax), fabsf(zmin)), fabsf(z))) * nxy);
      const float inv = 1.f / (fabsf(bcz) * (S1 * 0.99999f));
      const float ca = (G * 1.000001f * A2 + 2e-4f * nxy * fabsf(bcz) + 4.f * E * sqrtf(A2)) * inv;
      const float cb = sqrtf(A2) * inv;
      cut_ka = (ca * 1.00001f - a.t0) * a.inv_dt + 2.f;
      cut_kb = cb * 1.00001f * a.inv_dt;
      cut_ok = true;
    }
  }

  float qmin = __int_as_float(0x7f800000);   // +inf == "outside the face"
  int kmin = 255;
  RayConst rc;
  rc.hW = hW; rc.hH = hH; rc.neg_eps = neg_eps; rc.x = x; rc.y = y; rc.z = z; rc.bcx = bcx; rc.bcy = bcy; rc.bcz = bcz; rc.W = W; rc.H = H;
  if (ILP >= 2) {
    bool dirty = false;
    for (int k = k_begin; k <= k_end; k += ILP) {
      if (__any_sync(0xffffffffu, dirty)) {       // every lane is here once per iteration (k, k_end are warp-uniform)
        k_end = min(k_end, __reduce_max_sync(0xffffffffu, k_hi));
        dirty = false;
        if (k > k_end) break;
      }
      double px[ILP], py[ILP];
      bool in[ILP], any = false;
#pragma unroll
      for (int j = 0; j < ILP; ++j) {
        const int kj = min(k + j, k_end);                                             // tail: the group's last samples repeat k_end
        const double t = tab.t[kj];
        px[j] = __dadd_rn(xd, __dmul_rn(t, dx));                                      // TRAIN:472,480
        py[j] = __dadd_rn(yd, __dmul_rn(t, dy));
        const int mi = (cH - __double2loint(__dadd_rn(py[j], kMagic))) * W + __double2loint(__dadd_rn(px[j], kMagic)) + cW;
        in[j] = ((s_mask[mi >> 5] >> (mi & 31)) & 1u) && (j == 0 || k + j <= k_end);  // TRAIN:510-512
        any = any || in[j];
      }
      if (!any) continue;
      float q[ILP];
#pragma unroll
      for (int j = 0; j < ILP; ++j) q[j] = sample_q(D, rc, px[j], py[j]);
      const float q_before = qmin;
#pragma unroll
      for (int j = 0; j < ILP; ++j)
        if (in[j] && q[j] < qmin) { qmin = q[j]; kmin = k + j; }
      if (cut_ok && qmin < q_before) {
        const float kc = fmaf(cut_kb, sqrtf(qmin), cut_ka);
        if (kc < (float)k_hi) { k_hi = (int)kc; dirty = true; }       // kc >= 1 here (ca > 0): the truncation rounds towards the safe side of the +2
      }
    }
  } else {
  const float dxf32 = __fsub_rn(ex, x), dyf32 = __fsub_rn(ey, y);
  const float dt32 = a.inv_dt != 0.f ? 1.0f / a.inv_dt : 0.f;
  int k_skip_checked = k_begin - 1;          // the last group start that has been tested
#pragma unroll 2
  for (int k = k_begin; k <= k_end; ++k) {
    if (a.coarse && ((k - k_begin) & 3) == 0 && k + 3 <= k_end && k > k_skip_checked) {
      // warp-uniform: k, k_begin, k_end are; every lane tests the block of ITS ray's group middle
      const float tm = fmaf((float)k + 1.5f, dt32, a.t0);
      const float pxm = fmaf(tm, dxf32, x), pym = fmaf(tm, dyf32, y);
      int bxm = (int)floorf((pxm + halfW) * 0.125f), bym = (int)floorf((halfH - pym) * 0.125f);
      bxm = min(max(bxm, 0), cbw - 1); bym = min(max(bym, 0), cbh - 1);
      const int j = bym * cbw + bxm;
      const bool occ = (s_occ[j >> 5] >> (j & 31)) & 1u;
      k_skip_checked = k;
      if (!__any_sync(0xffffffffu, occ)) { k += 3; continue; }                        // the four samples k .. k+3 are outside the face
    }
    const double t = tab.t[k];
    const double px = __dadd_rn(xd, __dmul_rn(t, dx));                                // TRAIN:472,480
    const double py = __dadd_rn(yd, __dmul_rn(t, dy));
    const int ci = __double2loint(__dadd_rn(px, kMagic)) + cW;                        // TRAIN:472-475 (np.round: half to even)
    const int ri = cH - __double2loint(__dadd_rn(py, kMagic));
    const int mi = ri * W + ci;
    if (!((s_mask[mi >> 5] >> (mi & 31)) & 1u)) continue;                             // TRAIN:510-512
    const float q = sample_q(D, rc, px, py);
    if (q < qmin) { qmin = q; kmin = k; }
  }
  }
  float d;
  if (kmin == 255) {
    d = 1000000.0f;                                                                   // TRAIN:512
  } else {
    const float den = sqrtf(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(bcx, bcx), __fmul_rn(bcy, bcy)), __fmul_rn(bcz, bcz)), 1e-4f));
    d = __fdiv_rn(sqrtf(__fadd_rn(qmin, 1e-4f)), den);
  }
  if (a.bonus != 0.0f && Lx >= a.bx0 && Lx <= a.bx1 && Ly >= a.by0 && Ly <= a.by1) d = __fadd_rn(d, a.bonus);   // TEST1:495-496
  const size_t o = (size_t)b * H * W + row * W + col;
  if (a.dmin) a.dmin[o] = d;
  if (a.argmin) a.argmin[o] = (uint8_t)kmin;
  if (a.fuse_shade) gfr_shade::shade_pixel(a.shade, b, row, col, d);       // TRAIN:353-369, 517-522: d_min never leaves the SM
  else if (a.shadow) a.shadow[o] = shadow_weight(d);
}

// ---------------------------------------------------------------------------------------------------
// Variant 2 (A/B only): the mapping the north star sketched — ONE WARP PER PIXEL-RAY, lanes = samples (k = lane, lane + 32, ...),
// warp-shuffle arg-min over the 160 samples.  Same arithmetic as variant 0 (fp64 depth scratch, magic-add rounding), so the
// results are bit-identical; a CTA owns the same 32 x 8 pixel tile, warp w walks the 32 pixels of tile row w one after the
// other.  The depth map cannot be staged in shared memory (a ray crosses the whole image: 256 KB fp32 / 512 KB fp64 per face
// against 227 KB), so the gathers go through L1 like in variant 0.  Measured slower (profiles/r02_summary.md): the per-ray
// setup (two fp32 divisions, the 9-way end-point select) is replicated in 32 lanes, the lanes of one ray walk ALONG the ray
// (up to 32 cache lines per gather), in-mask / out-of-mask samples of one ray diverge, and every pixel pays a 10-shuffle
// reduction.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE_W * TILE_H)
shadow_march_fwd_warp_ray(const MarchArgs a, const double* __restrict__ depth64, const __grid_constant__ SampleTable tab) {
  extern __shared__ uint32_t s_mask[];
  const int b = blockIdx.z, f = b / a.lpf;
  const int H = a.H, W = a.W;
  const int words = (H * W) >> 5;
  {
    const uint32_t* src = a.mask_bits + (size_t)f * a.mask_stride;
    for (int i = threadIdx.y * TILE_W + threadIdx.x; i < words; i += TILE_W * TILE_H) s_mask[i] = __ldg(src + i);
  }
  __syncthreads();
  const int lane = threadIdx.x;                   // blockDim = (32, 8): warp = threadIdx.y
  const int row = blockIdx.y * TILE_H + threadIdx.y;
  const double* D = depth64 + (size_t)f * H * W;
  const float halfW = 0.5f * W, halfH = 0.5f * H;
  const float xmin = -halfW, xmax = W - halfW - 1.0f, ymin = 1.0f - halfH, ymax = halfH;
  const float Lx = __ldg(a.light + 3 * b), Ly = __ldg(a.light + 3 * b + 1), Lz = __ldg(a.light + 3 * b + 2);
  const double hW = a.hWd, hH = a.hHd, neg_eps = a.neg_eps;
  const int cW = W >> 1, cH = H >> 1;
  for (int pc = 0; pc < TILE_W; ++pc) {
    const int col = blockIdx.x * TILE_W + pc;
    const float x = (float)col - halfW, y = halfH - (float)row;
    const float z = __ldg(a.depth + (size_t)f * H * W + row * W + col);
    float ex, ey;
    ray_end(x, y, Lx, Ly, xmin, xmax, ymin, ymax, ex, ey);
    const double dx = (double)__fsub_rn(ex, x), dy = (double)__fsub_rn(ey, y);
    const double xd = (double)x, yd = (double)y;
    const float bcx = __fsub_rn(Lx, x), bcy = __fsub_rn(Ly, y), bcz = __fsub_rn(Lz, z);
    float qmin = __int_as_float(0x7f800000);
    int kmin = 255;
    for (int k = lane; k < a.n; k += 32) {
      const double t = tab.t[k];
      const double px = __dadd_rn(xd, __dmul_rn(t, dx));
      const double py = __dadd_rn(yd, __dmul_rn(t, dy));
      const int ci = __double2loint(__dadd_rn(px, kMagic)) + cW;
      const int ri = cH - __double2loint(__dadd_rn(py, kMagic));
      const int mi = ri * W + ci;
      if (!((s_mask[mi >> 5] >> (mi & 31)) & 1u)) continue;
      const double u = __dadd_rn(__dadd_rn(px, hW), neg_eps);
      const double v = __dadd_rn(__dsub_rn(hH, py), neg_eps);
      const double sfu = __dadd_rd(u, kMagic), scu = __dadd_ru(u, kMagic);
      const double sfv = __dadd_rd(v, kMagic), scv = __dadd_ru(v, kMagic);
      const int uf = __double2loint(sfu), uc = __double2loint(scu);
      const int vf = __double2loint(sfv), vc = __double2loint(scv);
      const double ufd = __dsub_rn(sfu, kMagic), ucd = __dsub_rn(scu, kMagic);
      const double vfd = __dsub_rn(sfv, kMagic), vcd = __dsub_rn(scv, kMagic);
      const unsigned ufi = uf < 0 ? uf + W : uf, vfi = vf < 0 ? vf + H : vf;
      const double wu0 = __dsub_rn(ucd, u), wu1 = __dsub_rn(u, ufd);
      const double wv0 = __dsub_rn(vcd, v), wv1 = __dsub_rn(v, vfd);
      const double* p00 = D + (vfi * (unsigned)W + ufi);
      const int du = uc - (int)ufi, dv = (vc - (int)vfi) * W;
      const double* p10 = p00 + dv;
      const double ul = __ldg(p00), ur = __ldg(p00 + du);
      const double ll = __ldg(p10), lr = __ldg(p10 + du);
      const double up = __dadd_rn(__dmul_rn(ul, wu0), __dmul_rn(ur, wu1));
      const double lo = __dadd_rn(__dmul_rn(ll, wu0), __dmul_rn(lr, wu1));
      const double zi = __dadd_rn(__dmul_rn(up, wv0), __dmul_rn(lo, wv1));
      const float ax = (float)__dsub_rn(u, hW), ay = (float)__dsub_rn(hH, v), az = (float)zi;
      const float bax = __fsub_rn(ax, x), bay = __fsub_rn(ay, y), baz = __fsub_rn(az, z);
      const float c0 = __fsub_rn(__fmul_rn(bay, bcz), __fmul_rn(baz, bcy));
      const float c1 = __fsub_rn(__fmul_rn(baz, bcx), __fmul_rn(bax, bcz));
      const float c2 = __fsub_rn(__fmul_rn(bax, bcy), __fmul_rn(bay, bcx));
      const float q = __fadd_rn(__fadd_rn(__fmul_rn(c0, c0), __fmul_rn(c1, c1)), __fmul_rn(c2, c2));
      if (q < qmin) { qmin = q; kmin = k; }          // k ascends within a lane: the first minimum is kept
    }
    // warp arg-min; among equal q the smallest k wins (what the sequential loop keeps)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float q2 = __shfl_xor_sync(0xffffffffu, qmin, o);
      const int k2 = __shfl_xor_sync(0xffffffffu, kmin, o);
      if (q2 < qmin || (q2 == qmin && k2 < kmin)) { qmin = q2; kmin = k2; }
    }
    if (lane == 0) {
      float d;
      if (kmin == 255) {
        d = 1000000.0f;
      } else {
        const float den = sqrtf(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(bcx, bcx), __fmul_rn(bcy, bcy)), __fmul_rn(bcz, bcz)), 1e-4f));
        d = __fdiv_rn(sqrtf(__fadd_rn(qmin, 1e-4f)), den);
      }
      if (a.bonus != 0.0f && Lx >= a.bx0 && Lx <= a.bx1 && Ly >= a.by0 && Ly <= a.by1) d = __fadd_rn(d, a.bonus);
      const size_t o = (size_t)b * H * W + row * W + col;
      if (a.dmin) a.dmin[o] = d;
      if (a.argmin) a.argmin[o] = (uint8_t)kmin;
      if (a.shadow) a.shadow[o] = shadow_weight(d);
    }
  }
}

__global__ void widen_depth_kernel(const float4* __restrict__ in, double* __restrict__ out, size_t n4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = __ldg(in + i);
  double2* o = reinterpret_cast<double2*>(out + 4 * i);
  o[0] = make_double2((double)v.x, (double)v.y);
  o[1] = make_double2((double)v.z, (double)v.w);
}

// The same pass + the depth range of every face over the pixels an in-mask sample can read: the mask dilated by one pixel
// (the bilinear footprint of a sample lies within one pixel of its nearest pixel) plus the last row and column (python's
// negative-index wrap, TRAIN:486-494).  range[f] = {key(max depth), key(max -depth)}, key = the order-preserving int image
// of a float, reduced with atomicMax; pre-set to 0x80808080 (memset) = "nothing yet".  Grid: (chunks, faces); W % 4 == 0.
__device__ __forceinline__ int float_key(float v) { const int b = __float_as_int(v); return b >= 0 ? b : b ^ 0x7fffffff; }

__global__ void __launch_bounds__(256) widen_depth_range_kernel(const float4* __restrict__ in, double* __restrict__ out,
                                                                const uint32_t* __restrict__ mask_bits, int mask_stride, int* __restrict__ range,
                                                                int H, int W) {
  const int f = blockIdx.y;
  const int n4 = (H * W) >> 2;
  const uint32_t* mb = mask_bits + (size_t)f * mask_stride;
  float vmax = -3.0e38f, vneg = -3.0e38f;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n4; i += gridDim.x * 256) {
    const float4 v = __ldg(in + (size_t)f * n4 + i);
    double2* o = reinterpret_cast<double2*>(out + ((size_t)f * n4 + i) * 4);
    o[0] = make_double2((double)v.x, (double)v.y);
    o[1] = make_double2((double)v.z, (double)v.w);
    const int p0 = i * 4, r = p0 / W, c0 = p0 % W;
    const float vv[4] = {v.x, v.y, v.z, v.w};
    // bits of columns c0 - 1 .. c0 + 4 of rows r - 1 .. r + 1, OR-ed over the rows (bit j = column c0 - 1 + j)
    uint32_t near = 0u;
    for (int dr = -1; dr <= 1; ++dr) {
      const int rr = r + dr;
      if (rr < 0 || rr >= H) continue;
      for (int j = 0; j < 6; ++j) {
        const int cc = c0 - 1 + j;
        if (cc < 0 || cc >= W) continue;
        const int bit = rr * W + cc;
        near |= ((__ldg(mb + (bit >> 5)) >> (bit & 31)) & 1u) << j;
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool use = ((near >> e) & 7u) != 0u || r == H - 1 || c0 + e == W - 1;
      if (use) { vmax = fmaxf(vmax, vv[e]); vneg = fmaxf(vneg, -vv[e]); }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    vneg = fmaxf(vneg, __shfl_xor_sync(0xffffffffu, vneg, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (vmax > -3.0e38f) atomicMax(range + 2 * f, float_key(vmax));
    if (vneg > -3.0e38f) atomicMax(range + 2 * f + 1, float_key(vneg));
  }
}

// ---------------------------------------------------------------------------------------------------
// mask packing
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void mask_pack_kernel(const T* __restrict__ mask, uint32_t* __restrict__ bits, size_t n_pixels, int H, int W) {
  // row of mask m in `bits`: H*W/32 bitmap words, then GFR_MASK_EXTRA_WORDS = 4 words {c_lo, -c_hi, r_lo, -r_hi}: the bounding
  // box of the non-zero pixels, kept with atomicMin (the buffer is pre-filled with 0x7f bytes)
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool nz = i < n_pixels ? (mask[i] != (T)0) : false;
  const uint32_t w = __ballot_sync(0xffffffffu, nz);
  const size_t hw = (size_t)H * W;
  const size_t m = i / hw, p = i % hw;
  const size_t row_words = hw / 32 + GFR_MASK_EXTRA_WORDS;
  if ((threadIdx.x & 31) == 0 && i < n_pixels) bits[m * row_words + (p >> 5)] = w;
  if (w != 0u) {                     // hw % 32 == 0: the 32 pixels of a warp belong to one mask
    const int r = (int)(p / W), c = (int)(p % W);
    const int big = 0x7f7f7f7f;
    const int c_lo = __reduce_min_sync(0xffffffffu, nz ? c : big), c_hi_n = __reduce_min_sync(0xffffffffu, nz ? -c : big);
    const int r_lo = __reduce_min_sync(0xffffffffu, nz ? r : big), r_hi_n = __reduce_min_sync(0xffffffffu, nz ? -r : big);
    if ((threadIdx.x & 31) == 0) {
      int* bb = reinterpret_cast<int*>(bits + m * row_words + hw / 32);
      atomicMin(bb + 0, c_lo); atomicMin(bb + 1, c_hi_n); atomicMin(bb + 2, r_lo); atomicMin(bb + 3, r_hi_n);
    }
  }
}

}  // namespace

extern "C" int gfr_mask_pack(const void* mask, int mask_dtype, int n_masks, int H, int W, uint32_t* bits, void* stream) {
  GFR_RETURN_IF_NULL(mask); GFR_RETURN_IF_NULL(bits);
  if (n_masks <= 0 || H <= 0 || W <= 0 || ((H * W) & 31)) return GFR_E_SHAPE;
  const size_t n = (size_t)n_masks * H * W;
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  cudaStream_t s = (cudaStream_t)stream;
  const cudaError_t e = cudaMemsetAsync(bits, 0x7f, (size_t)n_masks * ((size_t)H * W / 32 + GFR_MASK_EXTRA_WORDS) * 4, s);
  if (e != cudaSuccess) return (int)e;
  switch (mask_dtype) {
    case GFR_MASK_U8: mask_pack_kernel<uint8_t><<<blocks, threads, 0, s>>>((const uint8_t*)mask, bits, n, H, W); break;
    case GFR_MASK_F32: mask_pack_kernel<float><<<blocks, threads, 0, s>>>((const float*)mask, bits, n, H, W); break;
    case GFR_MASK_F64: mask_pack_kernel<double><<<blocks, threads, 0, s>>>((const double*)mask, bits, n, H, W); break;
    default: return GFR_E_ARG;
  }
  return gfr_launch_status();
}

// A/B configuration of the fast march kernel: -1 / 0 = the default (environment, else the built-in choice)
static int g_march_warp_shape = -1, g_march_ilp = 0, g_march_order = -1, g_march_cut = -1;

extern "C" int gfr_march_config(int warp_shape, int ilp, int block_order, int early_cut) {
  if (warp_shape < -1 || warp_shape > 1 || ilp < 0 || ilp > 4 || block_order < -1 || block_order > 1 || early_cut < -1 || early_cut > 1)
    return GFR_E_ARG;
  g_march_warp_shape = warp_shape; g_march_ilp = ilp; g_march_order = block_order; g_march_cut = early_cut;
  return GFR_OK;
}

static int march_impl(const float* depth, const uint32_t* mask_bits, int mask_batch_stride, const float* light_pt,
                      const double* t_host, int n, float inside_bonus, const float* bonus_rect_host, float* d_min, uint8_t* argmin, float* shadow,
                      double* depth64_scratch, int B, int H, int W, int lights_per_face, int variant,
                      const gfr_shade::ShadeArgs* fuse, void* stream) {
  GFR_RETURN_IF_NULL(depth); GFR_RETURN_IF_NULL(mask_bits); GFR_RETURN_IF_NULL(light_pt);
  GFR_RETURN_IF_NULL(t_host);
  if (fuse == nullptr) GFR_RETURN_IF_NULL(d_min);
  if (fuse != nullptr && (variant != 0 || depth64_scratch == nullptr)) return GFR_E_ARG;
  if (B <= 0 || H <= 0 || W <= 0 || (W % TILE_W) || (H % TILE_H) || H > 512 || W > 512 || B > 65535) return GFR_E_SHAPE;   // (B: the 3-D grids of variants 1, 2)
  if (n <= 0 || n > 255) return GFR_E_ARG;     // 255 is the "no sample inside the face" argmin code
  if (mask_batch_stride != 0 && mask_batch_stride != (H * W) / 32 + GFR_MASK_EXTRA_WORDS) return GFR_E_ARG;
  if (variant < 0 || variant > 2) return GFR_E_ARG;
  if (variant == 2 && depth64_scratch == nullptr) return GFR_E_NULL;
  if (lights_per_face < 1 || B % lights_per_face) return GFR_E_ARG;
  const int faces = B / lights_per_face;
  SampleTable tab;
  for (int k = 0; k < GFR_MAX_SAMPLES; ++k) tab.t[k] = k < n ? t_host[k] : 0.0;
    // culling needs t_k = t0 + k*dt (true for the reference's np.arange table); any other table marches every sample
  float t0 = (float)t_host[0], inv_dt = 0.f;
  if (n >= 2) {
    const double dt = (t_host[n - 1] - t_host[0]) / (n - 1);
    bool uniform = dt > 0.0;
    for (int k = 0; k < n && uniform; ++k) uniform = fabs(t_host[k] - (t_host[0] + k * dt)) <= 1e-9;
    static const bool no_cull = getenv("GFR_MARCH_NO_CULL") != nullptr;          // read once, not per launch
    if (uniform && !no_cull) inv_dt = (float)(1.0 / dt);
  }
  MarchArgs a = {};
  a.depth = depth; a.mask_bits = mask_bits; a.light = light_pt; a.dmin = d_min; a.argmin = argmin; a.shadow = shadow;
  a.mask_stride = mask_batch_stride; a.B = B; a.H = H; a.W = W; a.n = n; a.lpf = lights_per_face; a.t0 = t0; a.inv_dt = inv_dt;
  a.bonus = inside_bonus;
  a.hWd = (double)(0.5f * W); a.hHd = (double)(0.5f * H); a.neg_eps = -0.0001;
  if (bonus_rect_host != nullptr) {
    a.bx0 = bonus_rect_host[0]; a.bx1 = bonus_rect_host[1]; a.by0 = bonus_rect_host[2]; a.by1 = bonus_rect_host[3];
  } else {                                   // the image rectangle, TEST1:495
    a.bx0 = -0.5f * W; a.bx1 = W - 0.5f * W - 1.0f; a.by0 = 1.0f - 0.5f * H; a.by1 = 0.5f * H;
  }
  if (fuse != nullptr) { a.fuse_shade = 1; a.shade = *fuse; }
  dim3 grid(W / TILE_W, H / TILE_H, B), block(TILE_W, TILE_H);
  // A/B result (tools/time_march.py, B = 8, 256^2): 231.5 us with the coarse group skip vs 189.6 us without — the per-group test
  // (~15 instructions per 4 samples, two more registers) costs more than the skipped samples save on face-like masks, where the
  // bounding-box culling has already removed most out-of-mask samples.  Kept as an opt-in (GFR_MARCH_COARSE=1) for sparse masks.
  static const bool want_coarse = getenv("GFR_MARCH_COARSE") != nullptr;
  a.coarse = (inv_dt != 0.f && want_coarse && (W % 32) == 0 && (H % 8) == 0) ? 1 : 0;
  size_t smem = (size_t)(H * W / 32 + 2 * (((H / 8) * (W / 8) + 31) / 32 + 1)) * sizeof(uint32_t);
#ifdef GFR_MARCH_SMEM_PAD_KB        // A/B builds only: pad the dynamic shared memory so that fewer march CTAs share an SM
  if (smem < (size_t)GFR_MARCH_SMEM_PAD_KB * 1024) smem = (size_t)GFR_MARCH_SMEM_PAD_KB * 1024;
#endif
  static const int fast_th = [] { const char* e = getenv("GFR_MARCH_TILE_H"); return (e && atoi(e) == 8) ? 8 : 4; }();
  const bool fast_ok = ((size_t)faces * H * W) % 4 == 0 && (reinterpret_cast<uintptr_t>(depth) & 15) == 0;
  if (fuse != nullptr && !fast_ok) return GFR_E_SHAPE;
  if (variant == 2) {
    if (!fast_ok) return GFR_E_SHAPE;
    const size_t n4 = (size_t)faces * H * W / 4;
    widen_depth_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(depth), depth64_scratch, n4);
    shadow_march_fwd_warp_ray<<<grid, block, smem, (cudaStream_t)stream>>>(a, depth64_scratch, tab);
  } else if (variant == 0 && depth64_scratch != nullptr && fast_ok) {
    const size_t n4 = (size_t)faces * H * W / 4;
    static const int env_cut = [] { const char* e = getenv("GFR_MARCH_CUT"); return (e && atoi(e) == 0) ? 0 : 1; }();
    const int want_cut = g_march_cut >= 0 ? g_march_cut : env_cut;
    a.cut = (want_cut && inv_dt != 0.f && (W % 4) == 0 && faces <= 65535) ? 1 : 0;
    if (a.cut) {
      // the range words live behind the widened depth map (the scratch holds faces * H * W + faces doubles)
      int* range = reinterpret_cast<int*>(depth64_scratch + (size_t)faces * H * W);
      a.range = range;
      const cudaError_t e = cudaMemsetAsync(range, 0x80, (size_t)faces * 2 * sizeof(int), (cudaStream_t)stream);
      if (e != cudaSuccess) return (int)e;
      int chunks = (2 * 148) / faces;
      if (chunks < 1) chunks = 1;
      const int maxc = (H * W / 4 + 255) / 256;
      if (chunks > maxc) chunks = maxc;
      widen_depth_range_kernel<<<dim3(chunks, faces), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(depth), depth64_scratch,
                                                                                    mask_bits, mask_batch_stride, range, H, W);
    } else {
      widen_depth_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(depth), depth64_scratch, n4);
    }
    // A/B switches (read once): GFR_MARCH_WARP = 0 -> 32 x 1 warps (round 1), GFR_MARCH_ILP = 1 / 2 -> samples one by one / in pairs
    static const int env_warp_shape = [] { const char* e = getenv("GFR_MARCH_WARP"); return (e && atoi(e) == 0) ? 0 : 1; }();
    static const int env_ilp = [] { const char* e = getenv("GFR_MARCH_ILP"); const int v = e ? atoi(e) : 0; return v >= 1 && v <= 4 ? v : GFR_MARCH_DEFAULT_ILP; }();
    const int warp_shape = g_march_warp_shape >= 0 ? g_march_warp_shape : env_warp_shape;
    const int ilp = g_march_ilp > 0 ? g_march_ilp : env_ilp;
    static const int env_order = [] { const char* e = getenv("GFR_MARCH_ORDER"); return (e && atoi(e) == 0) ? 0 : 1; }();
    a.order = g_march_order >= 0 ? g_march_order : env_order;
    cudaStream_t st = (cudaStream_t)stream;
    if (fast_th == 4 && H % 4 == 0) {
      grid = dim3((unsigned)((W / TILE_W) * (H / 4)) * (unsigned)B, 1, 1); block.y = 4;
      if (warp_shape == 1 && !a.coarse) {
        if (ilp == 4) shadow_march_fwd_fast<4, 1, 4><<<grid, block, smem, st>>>(a, depth64_scratch, tab);
        else if (ilp == 3) shadow_march_fwd_fast<4, 1, 3><<<grid, block, smem, st>>>(a, depth64_scratch, tab);
        else if (ilp == 2) shadow_march_fwd_fast<4, 1, 2><<<grid, block, smem, st>>>(a, depth64_scratch, tab);
        else shadow_march_fwd_fast<4, 1, 1><<<grid, block, smem, st>>>(a, depth64_scratch, tab);
      } else {
        if (ilp >= 2 && !a.coarse) shadow_march_fwd_fast<4, 0, 2><<<grid, block, smem, st>>>(a, depth64_scratch, tab);
        else shadow_march_fwd_fast<4, 0, 1><<<grid, block, smem, st>>>(a, depth64_scratch, tab);
      }
    } else {
      grid = dim3((unsigned)((W / TILE_W) * (H / TILE_H)) * (unsigned)B, 1, 1);
      shadow_march_fwd_fast<8, 0, 1><<<grid, block, smem, st>>>(a, depth64_scratch, tab);
    }
  } else {
    shadow_march_fwd_l1<<<grid, block, smem, (cudaStream_t)stream>>>(a, tab);
  }
  return gfr_launch_status();
}

extern "C" int gfr_shadow_march_fwd(const float* depth, const uint32_t* mask_bits, int mask_batch_stride,
                                    const float* light_pt, const double* t_host, int n, float inside_bonus,
                                    const float* bonus_rect_host, float* d_min, uint8_t* argmin, float* shadow,
                                    double* depth64_scratch, int B, int H, int W, int lights_per_face, int variant,
                                    void* stream) {
  return march_impl(depth, mask_bits, mask_batch_stride, light_pt, t_host, n, inside_bonus, bonus_rect_host, d_min, argmin, shadow, depth64_scratch,
                    B, H, W, lights_per_face, variant, nullptr, stream);
}

extern "C" int gfr_march_shade_fwd(const float* albedo, const float* depth, const uint32_t* mask_bits, int mask_batch_stride,
                                   const float* light_pt, const float* ambient, const double* t_host, int n, float inside_bonus,
                                   const float* bonus_rect_host, const float* intr_host, double* depth64_scratch, float* d_min, uint8_t* argmin, float* shadow,
                                   float* full, float* final_shading, float* rendered, float* normals, int B, int H, int W,
                                   int lights_per_face, void* stream) {
  GFR_RETURN_IF_NULL(ambient); GFR_RETURN_IF_NULL(intr_host); GFR_RETURN_IF_NULL(depth64_scratch);
  if (rendered != nullptr && albedo == nullptr) return GFR_E_NULL;
  gfr_shade::ShadeArgs sh{albedo, depth, nullptr, light_pt, ambient, shadow, full, final_shading, rendered, normals, B, H, W,
                          lights_per_face, intr_host[0], intr_host[1], intr_host[2], intr_host[3], intr_host[4], intr_host[5]};
  return march_impl(depth, mask_bits, mask_batch_stride, light_pt, t_host, n, inside_bonus, bonus_rect_host, d_min, argmin, nullptr, depth64_scratch,
                    B, H, W, lights_per_face, 0, &sh, stream);
}
