// Per-pixel normals (kornia depth_to_normals) + Lambertian shading + shadow blend + render — TRAIN:353-369, 517-522 —
// shared by the stand-alone shade/render kernel and the fused epilogue of the ray-march kernel.
#pragma once
#include "gfr_common.cuh"

namespace gfr_shade {

struct ShadeArgs {
  const float* albedo;  // [B,3,H,W]
  const float* depth;   // [B,H,W]
  const float* dmin;    // [B,H,W]
  const float* light;   // [B,3]
  const float* ambient; // [B]
  float* shadow; float* full; float* final_shading; float* rendered; float* normals;
  int B, H, W;
  int lpf;              // lights per face: pair b reads albedo / depth / ambient of face b / lpf
  float fx, fy, cx, cy, depth_offset, intensity;
};

__device__ __forceinline__ void normalize3(float& a, float& b, float& c) {   // F.normalize(p=2, eps=1e-12)
  const float n = sqrtf(a * a + b * b + c * c);
  const float d = fmaxf(n, 1e-12f);
  a = __fdiv_rn(a, d); b = __fdiv_rn(b, d); c = __fdiv_rn(c, d);
}

// pixel (row, col) of (face, light) pair b, with its minimum ray distance d (TRAIN:515)
__device__ __forceinline__ void shade_pixel(const ShadeArgs& a, int b, int row, int col, float d) {
  const int f = b / a.lpf;
  const int H = a.H, W = a.W;
  const float* __restrict__ D = a.depth + (size_t)f * H * W;
  const size_t pix = (size_t)row * W + col;
  const size_t o = (size_t)b * H * W + pix;

  // --- kornia 0.4.1 depth_to_normals(depth + offset, K): xyz = ((u-cx)/fx, (v-cy)/fy, 1) * Z,
  //     Sobel/8 with replicate padding on each of x,y,z, n = normalize(cross(d/du, d/dv))
  float gu[3] = {0.f, 0.f, 0.f}, gv[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int i = -1; i <= 1; ++i) {
    const int r = min(max(row + i, 0), H - 1);
    const float py = __fdiv_rn((float)r - a.cy, a.fy);
#pragma unroll
    for (int j = -1; j <= 1; ++j) {
      const int c = min(max(col + j, 0), W - 1);
      const float Z = __ldg(D + r * W + c) + a.depth_offset;
      const float px = __fdiv_rn((float)c - a.cx, a.fx);
      const float X = px * Z, Y = py * Z;
      const float wu = 0.125f * (float)(j * (i == 0 ? 2 : 1));   // d/du kernel [[-1,0,1],[-2,0,2],[-1,0,1]]/8
      const float wv = 0.125f * (float)(i * (j == 0 ? 2 : 1));   // d/dv = transpose
      gu[0] += wu * X; gu[1] += wu * Y; gu[2] += wu * Z;
      gv[0] += wv * X; gv[1] += wv * Y; gv[2] += wv * Z;
    }
  }
  float nx = gu[1] * gv[2] - gu[2] * gv[1];
  float ny = gu[2] * gv[0] - gu[0] * gv[2];
  float nz = gu[0] * gv[1] - gu[1] * gv[0];
  normalize3(nx, ny, nz);
  ny = -ny;                       // TRAIN:354
  normalize3(nx, ny, nz);         // TRAIN:365

  // --- Lambert, TRAIN:356-369
  const float x = (float)col - 0.5f * W, y = 0.5f * H - (float)row, z = __ldg(D + pix);
  float lx = __ldg(a.light + 3 * b) - x, ly = __ldg(a.light + 3 * b + 1) - y, lz = __ldg(a.light + 3 * b + 2) - z;
  normalize3(lx, ly, lz);
  const float ndotl = (nx * lx + ny * ly) + nz * lz;
  const float directional = a.intensity * fmaxf(ndotl, 0.0f);
  const float amb = __ldg(a.ambient + f);
  const float full = amb + directional;

  // --- shadow weight + blend + render, TRAIN:517-522
  
  const float e = expf(-d);
  const float op = 1.0f + e;
  const float s = __fadd_rn(__fdiv_rn(__fmul_rn(-4.0f, e), __fmul_rn(op, op)), 1.0f);
  const float fin = __fadd_rn(__fmul_rn(s, full), __fmul_rn(__fsub_rn(1.0f, s), amb));
  if (a.shadow) a.shadow[o] = s;
  if (a.full) a.full[o] = full;
  if (a.final_shading) a.final_shading[o] = fin;
  const size_t plane = (size_t)H * W;
  if (a.rendered) {
    const float* A = a.albedo + (size_t)f * 3 * plane + pix;
    float* R = a.rendered + (size_t)b * 3 * plane + pix;
    R[0] = __ldg(A) * fin; R[plane] = __ldg(A + plane) * fin; R[2 * plane] = __ldg(A + 2 * plane) * fin;
  }
  if (a.normals) {
    float* N = a.normals + (size_t)b * 3 * plane + pix;
    N[0] = nx; N[plane] = ny; N[2 * plane] = nz;
  }
}

}  // namespace gfr_shade
