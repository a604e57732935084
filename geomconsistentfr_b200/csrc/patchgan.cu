// PatchGAN discriminator support (TRAIN:15-35; TRAIN = train_raytracing_relighting_CelebAHQ_DSSIM_8x.py).
//
// conv1..conv4 are 4x4 / stride 2 / pad 1.  With a space-to-depth of the input (channel c -> the four phase channels
// 4c + 2dy + dx, i.e. one C4 group per input channel), out[y,x] = sum w[ky,kx] X[2y+ky-1, 2x+kx-1] becomes a 3x3 / stride 1 /
// pad 1 convolution over the half-resolution tensor: row 2y+ky-1 = (row y-1, phase 1), (y, 0), (y, 1), (y+1, 0) for
// ky = 0..3, so the 3x3 kernel holds the 4x4 taps at (tap a, phase dy) in {(-1,1), (0,0), (0,1), (+1,0)} and zeros at the
// other two — and the layer runs on the tcgen05 3x3 kernel (gfr_conv3x3_tc_fwd) forward and backward.  This file holds
// the space-to-depth / depth-to-space passes, the LeakyReLU backward of the BN-less first layer, and conv5
// (4x4 / stride 1 / pad 1, 512 -> 1 on a 16x16 map: 1.8 MMAC per image, CUDA cores).
#include "gfr_common.cuh"

namespace {

// out[n][c][y'][x'][2dy+dx] = in[n, c, 2y'+dy, 2x'+dx];  in: NCHW planes (planar = 1) or C4
__global__ void s2d_kernel(const float* __restrict__ in, float4* __restrict__ out, int C, int H, int W, int planar, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over [N][C][H/2][W/2]
  if (i >= total) return;
  const int Wo = W >> 1, Ho = H >> 1;
  const int xo = (int)(i % Wo);
  long long t = i / Wo;
  const int yo = (int)(t % Ho); t /= Ho;
  const int c = (int)(t % C);
  const long long n = t / C;
  float v[4];
  if (planar) {
    const float* p = in + ((n * C + c) * H + 2 * yo) * (long long)W + 2 * xo;
    const float2 a = *reinterpret_cast<const float2*>(p), b = *reinterpret_cast<const float2*>(p + W);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else {
    const int C4 = (C + 3) >> 2;
    const float* p = in + (((n * C4 + (c >> 2)) * H + 2 * yo) * (long long)W + 2 * xo) * 4 + (c & 3);
    v[0] = __ldg(p); v[1] = __ldg(p + 4); v[2] = __ldg(p + (long long)W * 4); v[3] = __ldg(p + (long long)W * 4 + 4);
  }
  out[i] = make_float4(v[0], v[1], v[2], v[3]);
}

// inverse (the backward of s2d): g_in[n, c, 2y'+dy, 2x'+dx] = g_out[n][c][y'][x'][2dy+dx]
__global__ void d2s_kernel(const float4* __restrict__ g, float* __restrict__ out, int C, int H, int W, int planar, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int Wo = W >> 1, Ho = H >> 1;
  const int xo = (int)(i % Wo);
  long long t = i / Wo;
  const int yo = (int)(t % Ho); t /= Ho;
  const int c = (int)(t % C);
  const long long n = t / C;
  const float4 v = __ldg(g + i);
  if (planar) {
    float* p = out + ((n * C + c) * H + 2 * yo) * (long long)W + 2 * xo;
    *reinterpret_cast<float2*>(p) = make_float2(v.x, v.y);
    *reinterpret_cast<float2*>(p + W) = make_float2(v.z, v.w);
  } else {
    const int C4 = (C + 3) >> 2;
    float* p = out + (((n * C4 + (c >> 2)) * H + 2 * yo) * (long long)W + 2 * xo) * 4 + (c & 3);
    p[0] = v.x; p[4] = v.y; p[(long long)W * 4] = v.z; p[(long long)W * 4 + 4] = v.w;
  }
}

// Space-to-depth of the 1-PADDED input: out[n][c][y'][x'][2dy+dx] = Xpad[2y'+dy][2x'+dx], Xpad[r][s] = X[r-1][s-1] (0 outside),
// y' in [0, H/2], x' in [0, W/2].  A 4x4 / stride 2 / pad 1 convolution of X is then a 2x2-tap, stride-1 convolution of `out`
// (blocks y, y+1 hold the rows 2y-1 .. 2y+2) whose weights W'[co][4c+2dy+dx][a][b] = w[co][c][2a+dy][2b+dx] have no zeros.
__global__ void s2d_pad_kernel(const float* __restrict__ in, float4* __restrict__ out, int C, int H, int W, int planar, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over [N][C][H/2+1][W/2+1]
  if (i >= total) return;
  const int Wo = (W >> 1) + 1, Ho = (H >> 1) + 1;
  const int xo = (int)(i % Wo);
  long long t = i / Wo;
  const int yo = (int)(t % Ho); t /= Ho;
  const int c = (int)(t % C);
  const long long n = t / C;
  float v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = 2 * yo + (k >> 1) - 1, q = 2 * xo + (k & 1) - 1;
    float x = 0.f;
    if (r >= 0 && r < H && q >= 0 && q < W) {
      if (planar) x = __ldg(in + ((n * C + c) * H + r) * (long long)W + q);
      else x = __ldg(in + (((n * ((C + 3) >> 2) + (c >> 2)) * H + r) * (long long)W + q) * 4 + (c & 3));
    }
    v[k] = x;
  }
  out[i] = make_float4(v[0], v[1], v[2], v[3]);
}

// its backward: every input pixel belongs to exactly one (block, phase)
__global__ void d2s_pad_kernel(const float4* __restrict__ g, float* __restrict__ out, int C, int H, int W, int planar, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over [N][C][H][W]
  if (i >= total) return;
  const int q = (int)(i % W);
  long long t = i / W;
  const int r = (int)(t % H); t /= H;
  const int c = (int)(t % C);
  const long long n = t / C;
  const int Wo = (W >> 1) + 1, Ho = (H >> 1) + 1;
  const float4 v = __ldg(g + ((n * C + c) * Ho + ((r + 1) >> 1)) * (long long)Wo + ((q + 1) >> 1));
  const int k = 2 * ((r + 1) & 1) + ((q + 1) & 1);
  const float x = k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w;
  if (planar) out[i] = x;
  else out[(((n * ((C + 3) >> 2) + (c >> 2)) * H + r) * (long long)W + q) * 4 + (c & 3)] = x;
}

// C4 input with C % 4 == 0: one thread = one 4-channel group x one 2x2 pixel block — four 16-byte loads (the block's pixels, 4 channels
// each), a 4x4 transpose in registers, four 16-byte stores (output group 4*cg + e holds the four phases of channel 4*cg + e).  The
// scalar kernels above move 4 bytes per access at a 16-byte stride (1/4 of every sector used).
__global__ void __launch_bounds__(256) s2d_pad_c4_kernel(const float4* __restrict__ in, float4* __restrict__ out, int C4, int H, int W, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over [N][C4][H/2+1][W/2+1]
  if (i >= total) return;
  const int Wo = (W >> 1) + 1, Ho = (H >> 1) + 1;
  const int xo = (int)(i % Wo);
  long long t = i / Wo;
  const int yo = (int)(t % Ho); t /= Ho;
  const int cg = (int)(t % C4);
  const long long n = t / C4;
  float v[4][4];                                                              // [phase k][channel e]
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = 2 * yo + (k >> 1) - 1, q = 2 * xo + (k & 1) - 1;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= 0 && r < H && q >= 0 && q < W) x = __ldg(in + ((n * C4 + cg) * H + r) * (long long)W + q);
    v[k][0] = x.x; v[k][1] = x.y; v[k][2] = x.z; v[k][3] = x.w;
  }
  const long long plane = (long long)Ho * Wo;
  float4* o = out + ((n * (4LL * C4) + 4 * cg) * Ho + yo) * (long long)Wo + xo;
#pragma unroll
  for (int e = 0; e < 4; ++e) o[e * plane] = make_float4(v[0][e], v[1][e], v[2][e], v[3][e]);
}

__global__ void __launch_bounds__(256) d2s_pad_c4_kernel(const float4* __restrict__ g, float4* __restrict__ out, int C4, int H, int W, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over [N][C4][H/2+1][W/2+1]
  if (i >= total) return;
  const int Wo = (W >> 1) + 1, Ho = (H >> 1) + 1;
  const int xo = (int)(i % Wo);
  long long t = i / Wo;
  const int yo = (int)(t % Ho); t /= Ho;
  const int cg = (int)(t % C4);
  const long long n = t / C4;
  const long long plane = (long long)Ho * Wo;
  const float4* gp = g + ((n * (4LL * C4) + 4 * cg) * Ho + yo) * (long long)Wo + xo;
  float v[4][4];                                                              // [channel e][phase k]
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float4 x = __ldg(gp + e * plane);
    v[e][0] = x.x; v[e][1] = x.y; v[e][2] = x.z; v[e][3] = x.w;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = 2 * yo + (k >> 1) - 1, q = 2 * xo + (k & 1) - 1;
    if (r >= 0 && r < H && q >= 0 && q < W)
      out[((n * C4 + cg) * H + r) * (long long)W + q] = make_float4(v[0][k], v[1][k], v[2][k], v[3][k]);
  }
}

// g_pre = g_y * (y > 0 ? 1 : 0.2)   (y = LeakyReLU(pre): sign(y) = sign(pre))
__global__ void lrelu_bwd_kernel(const float4* __restrict__ y, const float4* __restrict__ gy, float4* __restrict__ out, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 a = __ldg(y + i), g = __ldg(gy + i);
  out[i] = make_float4(a.x > 0.f ? g.x : 0.2f * g.x, a.y > 0.f ? g.y : 0.2f * g.y, a.z > 0.f ? g.z : 0.2f * g.z,
                       a.w > 0.f ? g.w : 0.2f * g.w);
}

// ---- conv5: 4x4 / stride 1 / pad 1, C -> 1.  in C4 [N][C4][H][W][4]; w [C][4][4]; out [N][H-1][W-1]
__global__ void __launch_bounds__(128) conv4x4s1_fwd_kernel(const float* __restrict__ in, const float* __restrict__ w, float bias_dummy,
                                                             const float* __restrict__ bias, float* __restrict__ out, int C, int H, int W) {
  // one CTA per output pixel, threads over (channel group, tap)
  const int Ho = H - 1, Wo = W - 1;
  const int o = blockIdx.x, n = blockIdx.y;
  const int y = o / Wo, x = o % Wo;
  const int C4 = C >> 2;
  float s = 0.f;
  for (int i = threadIdx.x; i < C4 * 16; i += 128) {
    const int g = i >> 4, tap = i & 15, ky = tap >> 2, kx = tap & 3;
    const int iy = y + ky - 1, ix = x + kx - 1;
    if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
    const float4 v = __ldg(reinterpret_cast<const float4*>(in) + (((size_t)n * C4 + g) * H + iy) * W + ix);
    const float* wp = w + (size_t)(g * 4) * 16 + tap;
    s += v.x * __ldg(wp) + v.y * __ldg(wp + 16) + v.z * __ldg(wp + 32) + v.w * __ldg(wp + 48);
  }
  __shared__ float red[4];
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) out[(size_t)n * Ho * Wo + o] = red[0] + red[1] + red[2] + red[3] + __ldg(bias);
  (void)bias_dummy;
}

// g_in[n, c, iy, ix] = sum_{ky,kx} w[c][ky][kx] * g[n, iy-ky+1, ix-kx+1]
__global__ void conv4x4s1_dgrad_kernel(const float* __restrict__ g, const float* __restrict__ w, float4* __restrict__ gin, int C, int H,
                                       int W, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over [N][C4][H][W]
  if (i >= total) return;
  const int Ho = H - 1, Wo = W - 1, C4 = C >> 2;
  const int ix = (int)(i % W);
  long long t = i / W;
  const int iy = (int)(t % H); t /= H;
  const int grp = (int)(t % C4);
  const long long n = t / C4;
  float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int ky = 0; ky < 4; ++ky) {
    const int y = iy - ky + 1;
    if (y < 0 || y >= Ho) continue;
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) {
      const int x = ix - kx + 1;
      if (x < 0 || x >= Wo) continue;
      const float gv = __ldg(g + (size_t)n * Ho * Wo + y * Wo + x);
      const float* wp = w + (size_t)(grp * 4) * 16 + ky * 4 + kx;
      a[0] += gv * __ldg(wp); a[1] += gv * __ldg(wp + 16); a[2] += gv * __ldg(wp + 32); a[3] += gv * __ldg(wp + 48);
    }
  }
  gin[i] = make_float4(a[0], a[1], a[2], a[3]);
}

// dw[c][ky][kx] = sum_{n,y,x} g[n,y,x] * in[n, c, y+ky-1, x+kx-1];  db = sum g.   One warp per (c, tap).
__global__ void __launch_bounds__(128) conv4x4s1_wgrad_kernel(const float* __restrict__ in, const float* __restrict__ g, float* __restrict__ dw,
                                                               float* __restrict__ db, int N, int C, int H, int W) {
  const int Ho = H - 1, Wo = W - 1, C4 = C >> 2;
  const int item = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (item > C * 16) return;
  float s = 0.f;
  if (item == C * 16) {                          // bias
    for (int i = lane; i < N * Ho * Wo; i += 32) s += __ldg(g + i);
  } else {
    const int c = item >> 4, tap = item & 15, ky = tap >> 2, kx = tap & 3;
    for (int i = lane; i < N * Ho * Wo; i += 32) {
      const int n = i / (Ho * Wo), r = i % (Ho * Wo), y = r / Wo, x = r % Wo;
      const int iy = y + ky - 1, ix = x + kx - 1;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      s += __ldg(g + i) * __ldg(in + ((((size_t)n * C4 + (c >> 2)) * H + iy) * W + ix) * 4 + (c & 3));
    }
  }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
  if (lane == 0) { if (item == C * 16) atomicAdd(db, s); else atomicAdd(dw + item, s); }
}

}  // namespace

extern "C" int gfr_space_to_depth(const float* in, float* out, int N, int C, int H, int W, int in_is_nchw, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || ((H | W) & 1)) return GFR_E_SHAPE;
  const long long total = (long long)N * C * (H / 2) * (W / 2);
  s2d_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, reinterpret_cast<float4*>(out), C, H, W, in_is_nchw, total);
  return gfr_launch_status();
}

extern "C" int gfr_depth_to_space(const float* g, float* out, int N, int C, int H, int W, int out_is_nchw, void* stream) {
  GFR_RETURN_IF_NULL(g); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || ((H | W) & 1)) return GFR_E_SHAPE;
  const long long total = (long long)N * C * (H / 2) * (W / 2);
  d2s_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(g), out, C, H, W, out_is_nchw, total);
  return gfr_launch_status();
}

extern "C" int gfr_lrelu_bwd_c4(const float* y, const float* g_y, float* g_pre, long long n_floats, void* stream) {
  GFR_RETURN_IF_NULL(y); GFR_RETURN_IF_NULL(g_y); GFR_RETURN_IF_NULL(g_pre);
  if (n_floats <= 0 || (n_floats & 3)) return GFR_E_SHAPE;
  const long long n4 = n_floats / 4;
  lrelu_bwd_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(y), reinterpret_cast<const float4*>(g_y),
                                                                                  reinterpret_cast<float4*>(g_pre), n4);
  return gfr_launch_status();
}

extern "C" int gfr_conv4x4s1_to1_fwd(const float* in, const float* w, const float* bias, float* out, int N, int C, int H, int W, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(w); GFR_RETURN_IF_NULL(bias); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || N > 65535 || C <= 0 || (C & 3) || H < 2 || W < 2) return GFR_E_SHAPE;
  conv4x4s1_fwd_kernel<<<dim3((H - 1) * (W - 1), N), 128, 0, (cudaStream_t)stream>>>(in, w, 0.f, bias, out, C, H, W);
  return gfr_launch_status();
}

extern "C" int gfr_conv4x4s1_to1_bwd(const float* in, const float* w, const float* g_out, float* g_in, float* g_w, float* g_bias, int N, int C,
                                     int H, int W, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(w); GFR_RETURN_IF_NULL(g_out);
  if (N <= 0 || C <= 0 || (C & 3) || H < 2 || W < 2) return GFR_E_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  if (g_in) {
    const long long total = (long long)N * (C / 4) * H * W;
    conv4x4s1_dgrad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(g_out, w, reinterpret_cast<float4*>(g_in), C, H, W, total);
  }
  if (g_w) {
    if (!g_bias) return GFR_E_NULL;
    conv4x4s1_wgrad_kernel<<<(C * 16 + 1 + 3) / 4, 128, 0, s>>>(in, g_out, g_w, g_bias, N, C, H, W);
  }
  return gfr_launch_status();
}

extern "C" int gfr_space_to_depth_pad(const float* in, float* out, int N, int C, int H, int W, int in_is_nchw, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || ((H | W) & 1)) return GFR_E_SHAPE;
  if (!in_is_nchw && (C & 3) == 0) {
    const long long total4 = (long long)N * (C / 4) * ((H >> 1) + 1) * ((W >> 1) + 1);
    s2d_pad_c4_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out),
                                                                                         C / 4, H, W, total4);
    return gfr_launch_status();
  }
  const long long total = (long long)N * C * ((H >> 1) + 1) * ((W >> 1) + 1);
  s2d_pad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, reinterpret_cast<float4*>(out), C, H, W, in_is_nchw, total);
  return gfr_launch_status();
}

extern "C" int gfr_depth_to_space_pad(const float* g, float* out, int N, int C, int H, int W, int out_is_nchw, void* stream) {
  GFR_RETURN_IF_NULL(g); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || ((H | W) & 1)) return GFR_E_SHAPE;
  if (!out_is_nchw && (C & 3) == 0) {
    const long long total4 = (long long)N * (C / 4) * ((H >> 1) + 1) * ((W >> 1) + 1);
    d2s_pad_c4_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(out),
                                                                                         C / 4, H, W, total4);
    return gfr_launch_status();
  }
  const long long total = (long long)N * C * H * W;
  d2s_pad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(g), out, C, H, W, out_is_nchw, total);
  return gfr_launch_status();
}
