// Evaluation metrics of the reference's MATLAB scripts on the device.
//
// DSSIM_MP_RGB.m:15-27: for one relit / ground-truth image pair (uint8 RGB) and a face mask,
//     [~, ssimmap] = ssim(recon / 255, gt / 255);   average = sum(ssimmap .* mask3) / sum(mask3);   DSSIM = (1 - average) / 2.
// MATLAB's `ssim` on an M x N x 3 double array treats it as a VOLUME: the Gaussian window (sigma 1.5, radius ceil(3 * 1.5) = 5)
// is 11 x 11 x 11, applied with imfilter(..., 'replicate'), so the map is full size and the three colour planes are mixed by the
// window's third axis — with replicate padding over 3 planes that axis is the fixed 3 x 3 matrix CH below.  Dynamic range of a
// double image is 1: C1 = 0.01^2, C2 = 0.03^2, default exponents, so
//     map = ((2 mu_x mu_y + C1) (2 s_xy + C2)) / ((mu_x^2 + mu_y^2 + C1) (s_x^2 + s_y^2 + C2)),
// with mu = filt(x), s_x^2 = filt(x^2) - mu_x^2, s_xy = filt(x y) - mu_x mu_y.  No MATLAB / Octave exists in this image: the
// statement above is checked against oracle/metrics_oracle.py (the same published definition in numpy / scipy), everything in
// fp64 like MATLAB.  `window_3d = 0` gives the per-plane 2-D window instead (what a channel-wise caller would get).
#include "gfr_common.cuh"

namespace {

constexpr int R = 5, WIN = 11, TS = 16, IN = TS + 2 * R;       // 16 x 16 output tile, 26 x 26 replicate-padded input tile

struct SsimMetricArgs {
  const uint8_t* recon;   // [B,H,W,3]
  const uint8_t* gt;      // [B,H,W,3]
  const uint8_t* mask;    // [H,W] or [B,H,W], values 0..255 (used as mask / 255 like the .m file)
  long long mask_stride;
  double* sums;           // [B,2]: sum(map * mask3), sum(mask3)
  int H, W, window_3d;
  double g[WIN];          // normalised 1-D Gaussian
  double ch[3][3];        // channel mixing of the window's third axis under replicate padding
};

__global__ void __launch_bounds__(TS * TS) masked_ssim_map_kernel(const SsimMetricArgs a) {
  __shared__ double s_x[3][IN][IN], s_y[3][IN][IN];            // 2 x 3 x 26 x 26 x 8 B = 32.4 KB
  __shared__ double s_v[TS][IN];                                // one moment after the vertical pass
  __shared__ double s_red[2][TS * TS / 32];
  const int b = blockIdx.z, tid = threadIdx.x, tx = tid % TS, ty = tid / TS;
  const int x0 = blockIdx.x * TS, y0 = blockIdx.y * TS;
  const uint8_t* rp = a.recon + (size_t)b * a.H * a.W * 3;
  const uint8_t* gp = a.gt + (size_t)b * a.H * a.W * 3;
  for (int i = tid; i < IN * IN; i += TS * TS) {
    const int r = i / IN, c = i % IN;
    const int gy = min(max(y0 + r - R, 0), a.H - 1), gx = min(max(x0 + c - R, 0), a.W - 1);      // imfilter 'replicate'
    const size_t o = ((size_t)gy * a.W + gx) * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      s_x[k][r][c] = (double)rp[o + k] / 255.0;
      s_y[k][r][c] = (double)gp[o + k] / 255.0;
    }
  }
  __syncthreads();
  // five moments x three planes, each by a vertical then a horizontal 11-tap pass through s_v
  double mom[5][3];
#pragma unroll 1
  for (int q = 0; q < 5; ++q) {
#pragma unroll 1
    for (int k = 0; k < 3; ++k) {
      for (int i = tid; i < TS * IN; i += TS * TS) {
        const int r = i / IN, c = i % IN;
        double acc = 0.0;
#pragma unroll
        for (int t = 0; t < WIN; ++t) {
          const double xv = s_x[k][r + t][c], yv = s_y[k][r + t][c];
          const double v = q == 0 ? xv : q == 1 ? yv : q == 2 ? xv * xv : q == 3 ? yv * yv : xv * yv;
          acc += a.g[t] * v;
        }
        s_v[r][c] = acc;
      }
      __syncthreads();
      double acc = 0.0;
#pragma unroll
      for (int t = 0; t < WIN; ++t) acc += a.g[t] * s_v[ty][tx + t];
      mom[q][k] = acc;
      __syncthreads();
    }
  }
  const int y = y0 + ty, x = x0 + tx;
  double num = 0.0, den = 0.0;
  if (y < a.H && x < a.W) {
    const double m = (double)a.mask[(size_t)b * a.mask_stride + (size_t)y * a.W + x] / 255.0;
    const double C1 = 0.01 * 0.01, C2 = 0.03 * 0.03;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double f[5];
#pragma unroll
      for (int q = 0; q < 5; ++q)
        f[q] = a.window_3d ? a.ch[k][0] * mom[q][0] + a.ch[k][1] * mom[q][1] + a.ch[k][2] * mom[q][2] : mom[q][k];
      const double mux = f[0], muy = f[1];
      const double sx = f[2] - mux * mux, sy = f[3] - muy * muy, sxy = f[4] - mux * muy;
      const double v = ((2.0 * mux * muy + C1) * (2.0 * sxy + C2)) / ((mux * mux + muy * muy + C1) * (sx + sy + C2));
      num += v * m;
      den += m;
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    num += __shfl_xor_sync(0xFFFFFFFFu, num, s);
    den += __shfl_xor_sync(0xFFFFFFFFu, den, s);
  }
  if ((tid & 31) == 0) { s_red[0][tid >> 5] = num; s_red[1][tid >> 5] = den; }
  __syncthreads();
  if (tid == 0) {
    double n2 = 0.0, d2 = 0.0;
    for (int w = 0; w < TS * TS / 32; ++w) { n2 += s_red[0][w]; d2 += s_red[1][w]; }
    atomicAdd(a.sums + 2 * b, n2);
    atomicAdd(a.sums + 2 * b + 1, d2);
  }
}

}  // namespace

extern "C" int gfr_masked_ssim_u8(const uint8_t* recon, const uint8_t* gt, const uint8_t* mask, int mask_batch_stride, double* sums,
                                  int B, int H, int W, int window_3d, void* stream) {
  GFR_RETURN_IF_NULL(recon); GFR_RETURN_IF_NULL(gt); GFR_RETURN_IF_NULL(mask); GFR_RETURN_IF_NULL(sums);
  if (B <= 0 || B > 65535 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  if (mask_batch_stride != 0 && mask_batch_stride != H * W) return GFR_E_ARG;
  SsimMetricArgs a;
  a.recon = recon; a.gt = gt; a.mask = mask; a.mask_stride = mask_batch_stride; a.sums = sums;
  a.H = H; a.W = W; a.window_3d = window_3d ? 1 : 0;
  double s = 0.0;
  for (int i = 0; i < WIN; ++i) { const double c = (double)(i - R); a.g[i] = exp(-(c * c) / (2.0 * 1.5 * 1.5)); s += a.g[i]; }
  for (int i = 0; i < WIN; ++i) a.g[i] /= s;
  for (int k = 0; k < 3; ++k) {                 // window tap t reads plane clamp(k + t - R, 0, 2)
    a.ch[k][0] = a.ch[k][1] = a.ch[k][2] = 0.0;
    for (int t = 0; t < WIN; ++t) {
      int p = k + t - R;
      p = p < 0 ? 0 : (p > 2 ? 2 : p);
      a.ch[k][p] += a.g[t];
    }
  }
  const cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)B * 2 * sizeof(double), (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  const dim3 grid(gfr_ceil_div(W, TS), gfr_ceil_div(H, TS), B);
  masked_ssim_map_kernel<<<grid, TS * TS, 0, (cudaStream_t)stream>>>(a);
  return gfr_launch_status();
}
