// Backward of the differentiable hard-shadow ray tracer: K1b (ray-march) and K2b (normals + Lambert + render).
//
// K1b replaces what autograd does for TRAIN:374-517 (TRAIN = train_raytracing_relighting_CelebAHQ_DSSIM_8x.py):
// the min over the 160 samples routes the gradient to the arg-min sample only (TRAIN:514), so one thread per pixel
// re-evaluates that single sample with the forward's exact arithmetic and applies the closed-form chain rule:
//   d = sqrt(|BA x BC|^2 + eps) / sqrt(|BC|^2 + eps),  BA = A - P,  BC = P_L - P
//   -> depth:  the pixel's own z (through BA and BC) and the four bilinear corners of the sample (TRAIN:488-494)
//   -> light:  directly through BC, and through the sample POSITION: A moves with the ray end point, which depends on
//              the projected light through the slope / intercept of TRAIN:378-379 and the edge formulas of
//              TRAIN:385-458 (no gradient where the end point is clamped, TRAIN:462-465, or where the light
//              projects inside the image — the reference uses detached python floats there, TRAIN:423-425).
// K2b replaces autograd through kornia depth_to_normals (Sobel/8, replicate padding, cross, normalise), the y flip,
// the second normalise, l = normalize(P_L - P), the clamped n.l, the shadow blend and the albedo product
// (TRAIN:353-369, 517-522).
// Scatter-adds into grad_depth use atomicAdd (like the reference's index_put backward, the summation order — and so
// the last fp32 bit — is not deterministic).  grad buffers are ACCUMULATED into: the caller zero-fills them.
#include "gfr_common.cuh"

namespace {

struct SampleTable { double t[GFR_MAX_SAMPLES]; };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block (256 threads) reduction of up to 4 values, then one atomicAdd per value
template <int NV>
__device__ __forceinline__ void block_atomic_add(float (&v)[NV], float* const (&dst)[NV]) {
  __shared__ float s_red[NV][8];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float w = warp_sum(v[i]);
    if (lane == 0) s_red[i][warp] = w;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float w = lane < 8 ? s_red[i][lane] : 0.f;
      w = warp_sum(w);
      if (lane == 0 && w != 0.f) atomicAdd(dst[i], w);
    }
  }
}

// ------------------------------------------------------------------------------------------------- K1b
struct MarchBwdArgs {
  const float* depth;     // [B,H,W]
  const float* light;     // [B,3]
  const uint8_t* argmin;  // [B,H,W]
  const float* g_dmin;    // [B,H,W]   dL/d(d_min)
  float* g_depth;         // [B,H,W]   +=
  float* g_light;         // [B,3]     +=
  int B, H, W, n;
};

__global__ void __launch_bounds__(256) shadow_march_bwd_kernel(const MarchBwdArgs a, const __grid_constant__ SampleTable tab) {
  const int b = blockIdx.z, H = a.H, W = a.W;
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int row = blockIdx.y * 8 + threadIdx.y;
  float gl[3] = {0.f, 0.f, 0.f};
  if (col < W && row < H) {
    const size_t o = (size_t)b * H * W + (size_t)row * W + col;
    const int k = a.argmin[o];
    const float gd = __ldg(a.g_dmin + o);
    if (k != 255 && k < a.n && gd != 0.f) {
      const float* __restrict__ D = a.depth + (size_t)b * H * W;
      float* __restrict__ GD = a.g_depth + (size_t)b * H * W;
      const float halfW = 0.5f * W, halfH = 0.5f * H;
      const float xmin = -halfW, xmax = W - halfW - 1.0f, ymin = 1.0f - halfH, ymax = halfH;
      const float x = (float)col - halfW, y = halfH - (float)row, z = __ldg(D + row * W + col);
      const float Lx = __ldg(a.light + 3 * b), Ly = __ldg(a.light + 3 * b + 1), Lz = __ldg(a.light + 3 * b + 2);
      // ---- forward of the end point (TRAIN:378-465), keeping the branch taken
      const float den_m = __fadd_rn(__fsub_rn(Lx, x), 1e-4f);
      const float m = __fdiv_rn(__fsub_rn(Ly, y), den_m);
      const float bI = __fsub_rn(Ly, __fmul_rn(m, Lx));
      const int sx = Lx < xmin ? -1 : (Lx <= xmax ? 0 : 1);
      const int sy = Ly < ymin ? -1 : (Ly <= ymax ? 0 : 1);
      const float xe = sx < 0 ? xmin : xmax, ye = sy < 0 ? ymin : ymax;
      const float dm = __fadd_rn(m, 1e-4f);
      const float exy = __fadd_rn(__fmul_rn(m, xe), bI);          // y on the x edge
      const float eyx = __fdiv_rn(__fsub_rn(ye, bI), dm);         // x on the y edge
      const bool inside_img = sx == 0 && sy == 0;
      bool use_y = false;
      float ex, ey;
      if (sx != 0 && sy != 0) { use_y = (eyx >= xmin) && (eyx <= xmax); ex = use_y ? eyx : xe; ey = use_y ? ye : exy; }
      else if (sx != 0) { ex = xe; ey = exy; }
      else if (sy != 0) { use_y = true; ex = eyx; ey = ye; }
      else { ex = Lx; ey = Ly; }
      const bool ex_free = ex >= xmin && ex <= xmax, ey_free = ey >= ymin && ey <= ymax;     // not clamped
      ex = fminf(fmaxf(ex, xmin), xmax);
      ey = fminf(fmaxf(ey, ymin), ymax);
      // ---- forward of the arg-min sample (TRAIN:467-509)
      const double t = tab.t[k];
      const double px = __dadd_rn((double)x, __dmul_rn(t, (double)__fsub_rn(ex, x)));
      const double py = __dadd_rn((double)y, __dmul_rn(t, (double)__fsub_rn(ey, y)));
      const double u = __dadd_rn(__dadd_rn(px, (double)halfW), -0.0001);
      const double v = __dadd_rn(__dsub_rn((double)halfH, py), -0.0001);
      const int uf = __double2int_rd(u), uc = __double2int_ru(u), vf = __double2int_rd(v), vc = __double2int_ru(v);
      const int ufi = uf < 0 ? uf + W : uf, vfi = vf < 0 ? vf + H : vf;
      const double wu0 = __dsub_rn((double)uc, u), wu1 = __dsub_rn(u, (double)uf);
      const double wv0 = __dsub_rn((double)vc, v), wv1 = __dsub_rn(v, (double)vf);
      const double ul = (double)__ldg(D + vfi * W + ufi), ur = (double)__ldg(D + vfi * W + uc);
      const double ll = (double)__ldg(D + vc * W + ufi), lr = (double)__ldg(D + vc * W + uc);
      const double up = __dadd_rn(__dmul_rn(ul, wu0), __dmul_rn(ur, wu1));
      const double lo = __dadd_rn(__dmul_rn(ll, wu0), __dmul_rn(lr, wu1));
      const double zi = __dadd_rn(__dmul_rn(up, wv0), __dmul_rn(lo, wv1));
      const float ax = (float)__dsub_rn(u, (double)halfW), ay = (float)__dsub_rn((double)halfH, v), az = (float)zi;
      const float bax = ax - x, bay = ay - y, baz = az - z;
      const float bcx = Lx - x, bcy = Ly - y, bcz = Lz - z;
      const float c0 = bay * bcz - baz * bcy, c1 = baz * bcx - bax * bcz, c2 = bax * bcy - bay * bcx;
      const float q = c0 * c0 + c1 * c1 + c2 * c2;
      const float num = sqrtf(q + 1e-4f);
      const float bc2 = bcx * bcx + bcy * bcy + bcz * bcz;
      const float den = sqrtf(bc2 + 1e-4f);
      // ---- backward
      const float g_num = gd / den;
      const float g_den = -gd * num / (den * den);
      const float s = g_num / num;                                  // g_c = s * c
      const float gc0 = s * c0, gc1 = s * c1, gc2 = s * c2;
      // c = BA x BC:  g_BA = BC x g_c,  g_BC = g_c x BA  (+ the denominator's share)
      const float gba0 = bcy * gc2 - bcz * gc1, gba1 = bcz * gc0 - bcx * gc2, gba2 = bcx * gc1 - bcy * gc0;
      const float r = g_den / den;
      const float gbc0 = gc1 * baz - gc2 * bay + r * bcx;
      const float gbc1 = gc2 * bax - gc0 * baz + r * bcy;
      const float gbc2 = gc0 * bay - gc1 * bax + r * bcz;
      gl[0] = gbc0; gl[1] = gbc1; gl[2] = gbc2;                     // P_L through BC
      atomicAdd(GD + row * W + col, -(gba2 + gbc2));                // the pixel's own depth: BA_z = az - z, BC_z = Lz - z
      // bilinear corners (TRAIN:488-494)
      const float fwu0 = (float)wu0, fwu1 = (float)wu1, fwv0 = (float)wv0, fwv1 = (float)wv1;
      atomicAdd(GD + vfi * W + ufi, gba2 * fwu0 * fwv0);
      atomicAdd(GD + vfi * W + uc, gba2 * fwu1 * fwv0);
      atomicAdd(GD + vc * W + ufi, gba2 * fwu0 * fwv1);
      atomicAdd(GD + vc * W + uc, gba2 * fwu1 * fwv1);
      // sample position -> light (only when the end point depends on the light tensor)
      if (!inside_img) {
        const float daz_du = (float)((ur - ul) * wv0 + (lr - ll) * wv1);
        const float daz_dv = (float)(lo - up);
        const float g_px = gba0 + gba2 * daz_du;                    // ax = u - W/2, u = px + W/2 - 1e-4
        const float g_py = gba1 - gba2 * daz_dv;                    // ay = H/2 - v, v = H/2 - py - 1e-4
        const float tf = (float)t;
        const float g_ex = ex_free ? tf * g_px : 0.f;
        const float g_ey = ey_free ? tf * g_py : 0.f;
        float g_m = 0.f, g_b = 0.f;
        if (use_y) { g_b = -g_ex / dm; g_m = -g_ex * (ye - bI) / (dm * dm); }      // ex = (ye - b)/(m + 1e-4)
        else       { g_m = g_ey * xe; g_b = g_ey; }                                // ey = m*xe + b
        // b = Ly - m*Lx ; m = (Ly - y)/(Lx - x + 1e-4)
        g_m -= g_b * Lx;
        gl[1] += g_b + g_m / den_m;
        gl[0] += -g_b * m - g_m * m / den_m;
      }
    }
  }
  float* const dst[3] = {a.g_light + 3 * b, a.g_light + 3 * b + 1, a.g_light + 3 * b + 2};
  block_atomic_add<3>(gl, dst);
}

// ------------------------------------------------------------------------------------------------- K2b
struct ShadeBwdArgs {
  const float* albedo; const float* depth; const float* dmin; const float* light; const float* ambient;
  const float* g_shadow; const float* g_full; const float* g_final; const float* g_rendered; const float* g_normals;  // any may be null
  float* g_albedo;   // [B,3,H,W]  =  (written, not accumulated) or null
  float* g_depth;    // [B,H,W]    +=
  float* g_dmin;     // [B,H,W]    =
  float* g_light;    // [B,3]      +=
  float* g_ambient;  // [B]        +=
  int B, H, W;
  float fx, fy, cx, cy, depth_offset, intensity;
};

__device__ __forceinline__ float norm3(float a, float b, float c) { return sqrtf(a * a + b * b + c * c); }

// y = x / max(|x|, eps):  g_x = (g_y - y (y.g_y)) / max(|x|, eps)   (|x| > eps)
__device__ __forceinline__ void normalize_bwd(const float (&yv)[3], float nrm, float (&g)[3]) {
  const float d = fmaxf(nrm, 1e-12f);
  const float dot = yv[0] * g[0] + yv[1] * g[1] + yv[2] * g[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) g[i] = (g[i] - yv[i] * dot) / d;
}

__global__ void __launch_bounds__(256) shade_render_bwd_kernel(const ShadeBwdArgs a) {
  const int b = blockIdx.z, H = a.H, W = a.W;
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int row = blockIdx.y * 8 + threadIdx.y;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};       // g_light xyz, g_ambient
  if (col < W && row < H) {
    const float* __restrict__ D = a.depth + (size_t)b * H * W;
    float* __restrict__ GD = a.g_depth + (size_t)b * H * W;
    const size_t pix = (size_t)row * W + col, o = (size_t)b * H * W + pix, plane = (size_t)H * W;
    // ---- forward recompute (identical to shade_render_fwd_kernel)
    float gu[3] = {0.f, 0.f, 0.f}, gv[3] = {0.f, 0.f, 0.f};
    float pxs[3], pys[3];
    int rs[3], cs[3];
#pragma unroll
    for (int i = -1; i <= 1; ++i) {
      rs[i + 1] = min(max(row + i, 0), H - 1);
      cs[i + 1] = min(max(col + i, 0), W - 1);
      pys[i + 1] = __fdiv_rn((float)rs[i + 1] - a.cy, a.fy);
      pxs[i + 1] = __fdiv_rn((float)cs[i + 1] - a.cx, a.fx);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float Z = __ldg(D + rs[i] * W + cs[j]) + a.depth_offset;
        const float X = pxs[j] * Z, Y = pys[i] * Z;
        const float wu = 0.125f * (float)((j - 1) * (i == 1 ? 2 : 1));
        const float wv = 0.125f * (float)((i - 1) * (j == 1 ? 2 : 1));
        gu[0] += wu * X; gu[1] += wu * Y; gu[2] += wu * Z;
        gv[0] += wv * X; gv[1] += wv * Y; gv[2] += wv * Z;
      }
    const float cr[3] = {gu[1] * gv[2] - gu[2] * gv[1], gu[2] * gv[0] - gu[0] * gv[2], gu[0] * gv[1] - gu[1] * gv[0]};
    const float ncr = norm3(cr[0], cr[1], cr[2]);
    const float dcr = fmaxf(ncr, 1e-12f);
    const float n1[3] = {cr[0] / dcr, cr[1] / dcr, cr[2] / dcr};
    const float n1f[3] = {n1[0], -n1[1], n1[2]};                                   // TRAIN:354
    const float nn1f = norm3(n1f[0], n1f[1], n1f[2]);
    const float dn1f = fmaxf(nn1f, 1e-12f);
    const float n2[3] = {n1f[0] / dn1f, n1f[1] / dn1f, n1f[2] / dn1f};             // TRAIN:365
    const float x = (float)col - 0.5f * W, y = 0.5f * H - (float)row, z = __ldg(D + pix);
    const float w[3] = {__ldg(a.light + 3 * b) - x, __ldg(a.light + 3 * b + 1) - y, __ldg(a.light + 3 * b + 2) - z};
    const float nw = norm3(w[0], w[1], w[2]);
    const float dw = fmaxf(nw, 1e-12f);
    const float l[3] = {w[0] / dw, w[1] / dw, w[2] / dw};
    const float ndotl = (n2[0] * l[0] + n2[1] * l[1]) + n2[2] * l[2];
    const float directional = a.intensity * fmaxf(ndotl, 0.f);
    const float amb = __ldg(a.ambient + b);
    const float full = amb + directional;
    const float d = __ldg(a.dmin + o);
    const float e = expf(-d), op = 1.0f + e;
    const float s = 1.0f - 4.0f * e / (op * op);
    const float fin = s * full + (1.0f - s) * amb;
    // ---- backward
    float g_fin = a.g_final ? __ldg(a.g_final + o) : 0.f;
    if (a.g_rendered) {
      const float* A = a.albedo + (size_t)b * 3 * plane + pix;
      const float* G = a.g_rendered + (size_t)b * 3 * plane + pix;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float g = __ldg(G + c * plane);
        g_fin += g * __ldg(A + c * plane);
        if (a.g_albedo) a.g_albedo[(size_t)b * 3 * plane + c * plane + pix] = g * fin;
      }
    } else if (a.g_albedo) {
#pragma unroll
      for (int c = 0; c < 3; ++c) a.g_albedo[(size_t)b * 3 * plane + c * plane + pix] = 0.f;
    }
    const float g_s = (a.g_shadow ? __ldg(a.g_shadow + o) : 0.f) + g_fin * (full - amb);
    const float g_full = (a.g_full ? __ldg(a.g_full + o) : 0.f) + g_fin * s;
    acc[3] = g_fin * (1.0f - s) + g_full;                                          // ambient
    if (a.g_dmin) a.g_dmin[o] = g_s * (4.0f * e * (1.0f - e) / (op * op * op));    // ds/dd of TRAIN:517
    const float g_nl = ndotl >= 0.f ? a.intensity * g_full : 0.f;                  // clamp(min=0) backward
    float g_n2[3], g_l[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      g_n2[i] = g_nl * l[i] + (a.g_normals ? __ldg(a.g_normals + (size_t)b * 3 * plane + i * plane + pix) : 0.f);
      g_l[i] = g_nl * n2[i];
    }
    normalize_bwd(l, nw, g_l);                     // g_l is now g_w, w = P_L - P
    acc[0] = g_l[0]; acc[1] = g_l[1]; acc[2] = g_l[2];
    float g_z_self = -g_l[2];                      // P = (x, y, depth): only z carries a gradient
    normalize_bwd(n2, nn1f, g_n2);                 // -> g_n1f
    g_n2[1] = -g_n2[1];                            // -> g_n1
    normalize_bwd(n1, ncr, g_n2);                  // -> g_cr
    const float g_gu[3] = {gv[1] * g_n2[2] - gv[2] * g_n2[1], gv[2] * g_n2[0] - gv[0] * g_n2[2], gv[0] * g_n2[1] - gv[1] * g_n2[0]};   // gv x g_cr
    const float g_gv[3] = {g_n2[1] * gu[2] - g_n2[2] * gu[1], g_n2[2] * gu[0] - g_n2[0] * gu[2], g_n2[0] * gu[1] - g_n2[1] * gu[0]};   // g_cr x gu
    const bool any = (g_gu[0] != 0.f) | (g_gu[1] != 0.f) | (g_gu[2] != 0.f) | (g_gv[0] != 0.f) | (g_gv[1] != 0.f) | (g_gv[2] != 0.f);
    if (any) {
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (i == 1 && j == 1) continue;
          const float wu = 0.125f * (float)((j - 1) * (i == 1 ? 2 : 1));
          const float wv = 0.125f * (float)((i - 1) * (j == 1 ? 2 : 1));
          const float gz = wu * (g_gu[0] * pxs[j] + g_gu[1] * pys[i] + g_gu[2]) + wv * (g_gv[0] * pxs[j] + g_gv[1] * pys[i] + g_gv[2]);
          if (rs[i] == row && cs[j] == col) g_z_self += gz;        // replicate padding folds a neighbour onto the pixel
          else atomicAdd(GD + rs[i] * W + cs[j], gz);
        }
    }
    if (g_z_self != 0.f) atomicAdd(GD + pix, g_z_self);
  }
  float* const dst[4] = {a.g_light + 3 * b, a.g_light + 3 * b + 1, a.g_light + 3 * b + 2, a.g_ambient + b};
  block_atomic_add<4>(acc, dst);
}

}  // namespace

extern "C" int gfr_shadow_march_bwd(const float* depth, const float* light_pt, const uint8_t* argmin, const float* g_dmin,
                                    const double* t_host, int n, float* g_depth, float* g_light, int B, int H, int W,
                                    void* stream) {
  GFR_RETURN_IF_NULL(depth); GFR_RETURN_IF_NULL(light_pt); GFR_RETURN_IF_NULL(argmin); GFR_RETURN_IF_NULL(g_dmin);
  GFR_RETURN_IF_NULL(t_host); GFR_RETURN_IF_NULL(g_depth); GFR_RETURN_IF_NULL(g_light);
  if (B <= 0 || H <= 0 || W <= 0 || B > 65535) return GFR_E_SHAPE;
  if (n <= 0 || n > 255) return GFR_E_ARG;
  SampleTable tab;
  for (int k = 0; k < GFR_MAX_SAMPLES; ++k) tab.t[k] = k < n ? t_host[k] : 0.0;
  MarchBwdArgs a{depth, light_pt, argmin, g_dmin, g_depth, g_light, B, H, W, n};
  const dim3 grid(gfr_ceil_div(W, 32), gfr_ceil_div(H, 8), B), block(32, 8);
  shadow_march_bwd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(a, tab);
  return gfr_launch_status();
}

extern "C" int gfr_shade_render_bwd(const float* albedo, const float* depth, const float* d_min, const float* light_pt,
                                    const float* ambient, const float* intr_host, const float* g_shadow, const float* g_full,
                                    const float* g_final, const float* g_rendered, const float* g_normals, float* g_albedo,
                                    float* g_depth, float* g_dmin, float* g_light, float* g_ambient, int B, int H, int W,
                                    void* stream) {
  GFR_RETURN_IF_NULL(depth); GFR_RETURN_IF_NULL(d_min); GFR_RETURN_IF_NULL(light_pt); GFR_RETURN_IF_NULL(ambient);
  GFR_RETURN_IF_NULL(intr_host); GFR_RETURN_IF_NULL(g_depth); GFR_RETURN_IF_NULL(g_light); GFR_RETURN_IF_NULL(g_ambient);
  if (g_rendered != nullptr && albedo == nullptr) return GFR_E_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || B > 65535) return GFR_E_SHAPE;
  ShadeBwdArgs a{albedo, depth, d_min, light_pt, ambient, g_shadow, g_full, g_final, g_rendered, g_normals, g_albedo, g_depth,
                 g_dmin, g_light, g_ambient, B, H, W, intr_host[0], intr_host[1], intr_host[2], intr_host[3], intr_host[4],
                 intr_host[5]};
  const dim3 grid(gfr_ceil_div(W, 32), gfr_ceil_div(H, 8), B), block(32, 8);
  shade_render_bwd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(a);
  return gfr_launch_status();
}
