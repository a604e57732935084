// LPIPS (PerceptualSimilarity/lpips/lpips.py:112-144, the reference's third evaluation metric, test_network.py:41-48): the
// metric's own arithmetic on the device, forward and backward —
//   layer distance   d[n,p] = sum_c w_c (f0[n,c,p] / (|f0[n,:,p]| + eps) - f1[n,c,p] / (|f1[n,:,p]| + eps))^2
//                    (lpips.normalize_tensor + squared difference + the learned 1x1 `lin` head, lpips.py:125-131)
//   spatial map      out[n,Y,X] += bilinear_upsample(d)[n,Y,X]   (nn.Upsample(size, 'bilinear', align_corners=False), lpips.py:16-18)
//   masked mean      sum(mask * map) / count(mask * map > 0)       (test_network.py:41-45)
// The AlexNet trunk in front of it is five library convolutions (cuDNN through torch): its ImageNet weights do not exist in this
// image, so only the structure can be pinned (a seeded random trunk against the vendored lpips module).
#include "gfr_common.cuh"

namespace {

constexpr float LP_EPS = 1e-10f;

__global__ void __launch_bounds__(256) lpips_layer_fwd_kernel(const float* __restrict__ f0, const float* __restrict__ f1,
                                                              const float* __restrict__ w, float* __restrict__ out, int C, int HW,
                                                              long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over [N][HW]
  if (i >= total) return;
  const long long n = i / HW; const int p = (int)(i % HW);
  const float* a = f0 + n * C * (long long)HW + p;
  const float* b = f1 + n * C * (long long)HW + p;
  float sa = 0.f, sb = 0.f;
  for (int c = 0; c < C; ++c) { const float x = __ldg(a + (size_t)c * HW), y = __ldg(b + (size_t)c * HW); sa += x * x; sb += y * y; }
  const float ia = 1.0f / (sqrtf(sa) + LP_EPS), ib = 1.0f / (sqrtf(sb) + LP_EPS);
  float d = 0.f;
  for (int c = 0; c < C; ++c) {
    const float t = __ldg(a + (size_t)c * HW) * ia - __ldg(b + (size_t)c * HW) * ib;
    d = fmaf(__ldg(w + c) * t, t, d);
  }
  out[i] = d;
}

// with a = f0 / (n0 + eps):  d d / d f0_c = g * [ 2 w_c (a_c - b_c) / (n0 + eps) - f0_c / (n0 (n0 + eps)^2) * sum_k 2 w_k (a_k - b_k) f0_k ]
__global__ void __launch_bounds__(256) lpips_layer_bwd_kernel(const float* __restrict__ f0, const float* __restrict__ f1,
                                                              const float* __restrict__ w, const float* __restrict__ g_out,
                                                              float* __restrict__ g0, float* __restrict__ g1, int C, int HW, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long n = i / HW; const int p = (int)(i % HW);
  const size_t base = (size_t)n * C * HW + p;
  const float* a = f0 + base;
  const float* b = f1 + base;
  float sa = 0.f, sb = 0.f;
  for (int c = 0; c < C; ++c) { const float x = __ldg(a + (size_t)c * HW), y = __ldg(b + (size_t)c * HW); sa += x * x; sb += y * y; }
  const float na = sqrtf(sa), nb = sqrtf(sb);
  const float ia = 1.0f / (na + LP_EPS), ib = 1.0f / (nb + LP_EPS);
  float da = 0.f, db = 0.f;                    // sum_k 2 w_k (a_k - b_k) f0_k  and  ... f1_k
  for (int c = 0; c < C; ++c) {
    const float x = __ldg(a + (size_t)c * HW), y = __ldg(b + (size_t)c * HW);
    const float t = 2.0f * __ldg(w + c) * (x * ia - y * ib);
    da = fmaf(t, x, da); db = fmaf(t, y, db);
  }
  const float g = __ldg(g_out + i);
  const float ka = na > 0.f ? da * ia * ia / na : 0.f, kb = nb > 0.f ? db * ib * ib / nb : 0.f;
  for (int c = 0; c < C; ++c) {
    const float x = __ldg(a + (size_t)c * HW), y = __ldg(b + (size_t)c * HW);
    const float t = 2.0f * __ldg(w + c) * (x * ia - y * ib);
    if (g0) g0[base + (size_t)c * HW] = g * (t * ia - x * ka);
    if (g1) g1[base + (size_t)c * HW] = -g * (t * ib - y * kb);
  }
}

__device__ __forceinline__ void bilinear_src(int dst, float scale, int in, int& i0, int& i1, float& l1) {
  float r = scale * ((float)dst + 0.5f) - 0.5f;           // align_corners = False (ATen area_pixel_compute_source_index)
  if (r < 0.f) r = 0.f;
  i0 = (int)r;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = r - (float)i0;
}

__global__ void __launch_bounds__(256) bilinear_up_add_kernel(const float* __restrict__ m, float* __restrict__ out, int h, int w, int H, int W,
                                                              float sy, float sx, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over [N][H][W]
  if (i >= total) return;
  const int X = (int)(i % W); const long long t = i / W; const int Y = (int)(t % H); const long long n = t / H;
  int y0, y1, x0, x1; float ly, lx;
  bilinear_src(Y, sy, h, y0, y1, ly);
  bilinear_src(X, sx, w, x0, x1, lx);
  const float* p = m + n * (long long)h * w;
  const float v = (1.f - ly) * ((1.f - lx) * __ldg(p + y0 * w + x0) + lx * __ldg(p + y0 * w + x1)) +
                  ly * ((1.f - lx) * __ldg(p + y1 * w + x0) + lx * __ldg(p + y1 * w + x1));
  out[i] += v;
}

__global__ void __launch_bounds__(256) bilinear_up_bwd_kernel(const float* __restrict__ g, float* __restrict__ gm, int h, int w, int H, int W,
                                                              float sy, float sx, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int X = (int)(i % W); const long long t = i / W; const int Y = (int)(t % H); const long long n = t / H;
  int y0, y1, x0, x1; float ly, lx;
  bilinear_src(Y, sy, h, y0, y1, ly);
  bilinear_src(X, sx, w, x0, x1, lx);
  float* p = gm + n * (long long)h * w;
  const float v = __ldg(g + i);
  atomicAdd(p + y0 * w + x0, v * (1.f - ly) * (1.f - lx));
  atomicAdd(p + y0 * w + x1, v * (1.f - ly) * lx);
  atomicAdd(p + y1 * w + x0, v * ly * (1.f - lx));
  atomicAdd(p + y1 * w + x1, v * ly * lx);
}

__global__ void __launch_bounds__(256) lpips_masked_sums_kernel(const float* __restrict__ map, const float* __restrict__ mask,
                                                                long long mask_stride, double* __restrict__ sums, int HW) {
  const int n = blockIdx.y;
  double s = 0.0, cnt = 0.0;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    const float v = __ldg(mask + (long long)n * mask_stride + p) * __ldg(map + (long long)n * HW + p);
    s += (double)v;
    cnt += v > 0.f ? 1.0 : 0.0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
  __shared__ double r0[8], r1[8];
  if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = s; r1[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int k = 0; k < 8; ++k) { a += r0[k]; c += r1[k]; }
    atomicAdd(sums + 2 * n, a);
    atomicAdd(sums + 2 * n + 1, c);
  }
}

}  // namespace

extern "C" int gfr_lpips_layer_fwd(const float* f0, const float* f1, const float* w, float* out, int N, int C, int HW, void* stream) {
  GFR_RETURN_IF_NULL(f0); GFR_RETURN_IF_NULL(f1); GFR_RETURN_IF_NULL(w); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || C <= 0 || HW <= 0) return GFR_E_SHAPE;
  const long long total = (long long)N * HW;
  lpips_layer_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(f0, f1, w, out, C, HW, total);
  return gfr_launch_status();
}

extern "C" int gfr_lpips_layer_bwd(const float* f0, const float* f1, const float* w, const float* g_out, float* g_f0, float* g_f1,
                                   int N, int C, int HW, void* stream) {
  GFR_RETURN_IF_NULL(f0); GFR_RETURN_IF_NULL(f1); GFR_RETURN_IF_NULL(w); GFR_RETURN_IF_NULL(g_out);
  if (N <= 0 || C <= 0 || HW <= 0) return GFR_E_SHAPE;
  if (!g_f0 && !g_f1) return GFR_OK;
  const long long total = (long long)N * HW;
  lpips_layer_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(f0, f1, w, g_out, g_f0, g_f1, C, HW, total);
  return gfr_launch_status();
}

extern "C" int gfr_bilinear_up_add(const float* m, float* out, int N, int h, int w, int H, int W, void* stream) {
  GFR_RETURN_IF_NULL(m); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  const long long total = (long long)N * H * W;
  bilinear_up_add_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(m, out, h, w, H, W, (float)h / (float)H,
                                                                                            (float)w / (float)W, total);
  return gfr_launch_status();
}

extern "C" int gfr_bilinear_up_add_bwd(const float* g_out, float* g_m, int N, int h, int w, int H, int W, void* stream) {
  GFR_RETURN_IF_NULL(g_out); GFR_RETURN_IF_NULL(g_m);
  if (N <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  const long long total = (long long)N * H * W;
  bilinear_up_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g_out, g_m, h, w, H, W, (float)h / (float)H,
                                                                                            (float)w / (float)W, total);
  return gfr_launch_status();
}

extern "C" int gfr_lpips_masked_sums(const float* map, const float* mask, int mask_batch_stride, double* sums, int N, int H, int W,
                                     void* stream) {
  GFR_RETURN_IF_NULL(map); GFR_RETURN_IF_NULL(mask); GFR_RETURN_IF_NULL(sums);
  if (N <= 0 || N > 65535 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  if (mask_batch_stride != 0 && mask_batch_stride != H * W) return GFR_E_ARG;
  const cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)N * 2 * sizeof(double), (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  lpips_masked_sums_kernel<<<dim3((unsigned)min(64, gfr_ceil_div(H * W, 256)), N), 256, 0, (cudaStream_t)stream>>>(map, mask, mask_batch_stride,
                                                                                                                 sums, H * W);
  return gfr_launch_status();
}
