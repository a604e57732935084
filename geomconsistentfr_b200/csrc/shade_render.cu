// Normals (kornia depth_to_normals) + Lambertian shading + shadow blend + render, forward, sm_100a.
// Replaces TRAIN:353-369 and TRAIN:517-522 (TEST1:326-346, 498-503).
//
// One thread per pixel, 32x8 tiles; the 3x3 depth neighbourhood comes through L1 (9 loads of which 6
// are shared with the neighbouring threads).  HBM-bound: 16 B read + up to 36 B written per pixel.
#include "gfr_common.cuh"
#include "shade_device.cuh"

namespace {

using namespace gfr_shade;

__global__ void __launch_bounds__(256) shade_render_fwd_kernel(const ShadeArgs a) {
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int row = blockIdx.y * 8 + threadIdx.y;
  const int b = blockIdx.z;
  if (col >= a.W || row >= a.H) return;
  shade_pixel(a, b, row, col, __ldg(a.dmin + (size_t)b * a.H * a.W + (size_t)row * a.W + col));
}

}  // namespace

extern "C" int gfr_shade_render_fwd(const float* albedo, const float* depth, const float* d_min, const float* light_pt,
                                    const float* ambient, const float* intr_host, float* shadow, float* full,
                                    float* final_shading, float* rendered, float* normals, int B, int H, int W,
                                    int lights_per_face, void* stream) {
  GFR_RETURN_IF_NULL(depth); GFR_RETURN_IF_NULL(d_min); GFR_RETURN_IF_NULL(light_pt);
  GFR_RETURN_IF_NULL(ambient); GFR_RETURN_IF_NULL(intr_host);
  if (rendered != nullptr && albedo == nullptr) return GFR_E_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || B > 65535) return GFR_E_SHAPE;
  if (lights_per_face < 1 || B % lights_per_face) return GFR_E_ARG;
  ShadeArgs a{albedo, depth, d_min, light_pt, ambient, shadow, full, final_shading, rendered, normals, B, H, W, lights_per_face,
              intr_host[0], intr_host[1], intr_host[2], intr_host[3], intr_host[4], intr_host[5]};
  const dim3 grid(gfr_ceil_div(W, 32), gfr_ceil_div(H, 8), B), block(32, 8);
  shade_render_fwd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(a);
  return gfr_launch_status();
}
