// Training-loss kernels of the relight path (TRAIN:633-645; TRAIN = train_raytracing_relighting_CelebAHQ_DSSIM_8x.py).
//
// K4  SSIM forward + backward — replaces pytorch_msssim.ssim(X, Y, data_range=1, size_average=True,
//     nonnegative_ssim=True) at TRAIN:643 (5 separable VALID 11-tap Gaussian filters + ~15 elementwise launches, and
//     autograd through them).  Forward: one CTA per 16x16 tile of the (H-10)x(W-10) SSIM map and (n,c) plane; the
//     26x26 input tiles of X and Y are staged once, the five filtered moments (mu_x, mu_y, E[xx], E[yy], E[xy]) are
//     produced by a vertical then a horizontal pass in shared memory (pytorch_msssim's order), the per-plane sum goes
//     out through one atomicAdd per CTA and the three partial derivatives dS/dmu_x, dS/dE[xx], dS/dE[xy] are written
//     for the backward, which applies the transposed filter: g_X = G^T*g1 + 2X (G^T*g2) + Y (G^T*g3).
// K5  masked reconstruction / depth / albedo losses with their gradients in one pass (TRAIN:633-639): sums in fp64
//     like the reference (its float64 masks promote the loss arithmetic).
// Adam (TRAIN:589-590, 631, 656): torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8) over one flat parameter buffer.
#include "gfr_common.cuh"

namespace {

constexpr int WIN = 11, TS = 16, IT = TS + WIN - 1;       // window, output tile, input tile (26)

struct SsimWeights { float w[WIN]; };

struct SsimFwdArgs {
  const float* X; const float* Y;   // [P,H,W]   P = N*C planes
  double* sums;                     // [P]  += sum of the SSIM map
  float* gmaps;                     // [3,P,H-10,W-10] or null: dS/dmu_x, dS/dE[xx], dS/dE[xy]
  int P, H, W;
  float C1, C2;
};

__global__ void __launch_bounds__(TS * TS) ssim_fwd_kernel(const SsimFwdArgs a, const __grid_constant__ SsimWeights g) {
  __shared__ float sx[IT][IT + 1], sy[IT][IT + 1];
  __shared__ float sv[5][TS][IT + 1];
  __shared__ double s_red[TS * TS / 32];
  const int Ho = a.H - (WIN - 1), Wo = a.W - (WIN - 1);
  const int tiles_x = gfr_ceil_div(Wo, TS);
  const int x0 = (blockIdx.x % tiles_x) * TS, y0 = (blockIdx.x / tiles_x) * TS;
  const int p = blockIdx.y, tid = threadIdx.x;
  const float* X = a.X + (size_t)p * a.H * a.W;
  const float* Y = a.Y + (size_t)p * a.H * a.W;
  for (int i = tid; i < IT * IT; i += TS * TS) {
    const int r = i / IT, c = i % IT;
    const int gy = y0 + r, gx = x0 + c;
    const bool ok = gy < a.H && gx < a.W;
    sx[r][c] = ok ? __ldg(X + (size_t)gy * a.W + gx) : 0.f;
    sy[r][c] = ok ? __ldg(Y + (size_t)gy * a.W + gx) : 0.f;
  }
  __syncthreads();
  // vertical pass (along rows): 16 x 26 outputs per moment
  for (int i = tid; i < TS * IT; i += TS * TS) {
    const int r = i / IT, c = i % IT;
    float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < WIN; ++k) {
      const float xv = sx[r + k][c], yv = sy[r + k][c], w = g.w[k];
      m[0] = fmaf(w, xv, m[0]); m[1] = fmaf(w, yv, m[1]);
      m[2] = fmaf(w, xv * xv, m[2]); m[3] = fmaf(w, yv * yv, m[3]); m[4] = fmaf(w, xv * yv, m[4]);
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) sv[q][r][c] = m[q];
  }
  __syncthreads();
  const int r = tid / TS, c = tid % TS;
  float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < WIN; ++k) {
    const float w = g.w[k];
#pragma unroll
    for (int q = 0; q < 5; ++q) m[q] = fmaf(w, sv[q][r][c + k], m[q]);
  }
  const int oy = y0 + r, ox = x0 + c;
  const bool ok = oy < Ho && ox < Wo;
  const float mu1 = m[0], mu2 = m[1];
  const float s1 = m[2] - mu1 * mu1, s2 = m[3] - mu2 * mu2, s12 = m[4] - mu1 * mu2;
  const float dA = mu1 * mu1 + mu2 * mu2 + a.C1, nA = 2.f * mu1 * mu2 + a.C1;
  const float dB = s1 + s2 + a.C2, nB = 2.f * s12 + a.C2;
  const float A = nA / dA, Bv = nB / dB;
  double S = ok ? (double)(A * Bv) : 0.0;
  if (ok && a.gmaps) {
    const float dA_dmu1 = (2.f * mu2 * dA - nA * 2.f * mu1) / (dA * dA);
    const float dB_ds12 = 2.f / dB, dB_ds1 = -nB / (dB * dB);
    const size_t plane = (size_t)a.P * Ho * Wo, o = (size_t)p * Ho * Wo + (size_t)oy * Wo + ox;
    a.gmaps[o] = Bv * dA_dmu1 + A * (dB_ds12 * (-mu2) + dB_ds1 * (-2.f * mu1));
    a.gmaps[plane + o] = A * dB_ds1;
    a.gmaps[2 * plane + o] = A * dB_ds12;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) S += __shfl_xor_sync(0xffffffffu, S, o);
  if ((tid & 31) == 0) s_red[tid >> 5] = S;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int i = 0; i < TS * TS / 32; ++i) t += s_red[i];
    atomicAdd(a.sums + p, t);
  }
}

struct SsimBwdArgs {
  const float* X; const float* Y; const float* gmaps;   // gmaps [3,P,Ho,Wo]
  const float* scale;                                   // [P]: upstream dL/d(mean SSIM of plane p) / (Ho*Wo)
  float* gX;                                            // [P,H,W] written
  int P, H, W;
};

__global__ void __launch_bounds__(TS * TS) ssim_bwd_kernel(const SsimBwdArgs a, const __grid_constant__ SsimWeights g) {
  __shared__ float sg[3][IT][IT + 1];
  __shared__ float sv[3][TS][IT + 1];
  const int Ho = a.H - (WIN - 1), Wo = a.W - (WIN - 1);
  const int tiles_x = gfr_ceil_div(a.W, TS);
  const int x0 = (blockIdx.x % tiles_x) * TS, y0 = (blockIdx.x / tiles_x) * TS;
  const int p = blockIdx.y, tid = threadIdx.x;
  const size_t plane = (size_t)a.P * Ho * Wo;
  // g_X(y,x) = sum_{i,j} w_i w_j g(y-i, x-j): stage g over [y0-10, y0+15] x [x0-10, x0+15], zero outside the map
  for (int i = tid; i < IT * IT; i += TS * TS) {
    const int r = i / IT, c = i % IT;
    const int gy = y0 + r - (WIN - 1), gx = x0 + c - (WIN - 1);
    const bool ok = gy >= 0 && gy < Ho && gx >= 0 && gx < Wo;
    const size_t o = (size_t)p * Ho * Wo + (size_t)(ok ? gy : 0) * Wo + (ok ? gx : 0);
#pragma unroll
    for (int q = 0; q < 3; ++q) sg[q][r][c] = ok ? __ldg(a.gmaps + q * plane + o) : 0.f;
  }
  __syncthreads();
  for (int i = tid; i < TS * IT; i += TS * TS) {
    const int r = i / IT, c = i % IT;             // output row r needs staged rows r .. r+10 with w reversed
    float m[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < WIN; ++k) {
      const float w = g.w[WIN - 1 - k];
#pragma unroll
      for (int q = 0; q < 3; ++q) m[q] = fmaf(w, sg[q][r + k][c], m[q]);
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) sv[q][r][c] = m[q];
  }
  __syncthreads();
  const int r = tid / TS, c = tid % TS;
  float m[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < WIN; ++k) {
    const float w = g.w[WIN - 1 - k];
#pragma unroll
    for (int q = 0; q < 3; ++q) m[q] = fmaf(w, sv[q][r][c + k], m[q]);
  }
  const int oy = y0 + r, ox = x0 + c;
  if (oy < a.H && ox < a.W) {
    const size_t o = (size_t)p * a.H * a.W + (size_t)oy * a.W + ox;
    const float xv = __ldg(a.X + o), yv = __ldg(a.Y + o);
    a.gX[o] = __ldg(a.scale + p) * (m[0] + 2.f * xv * m[1] + yv * m[2]);
  }
}

// ------------------------------------------------------------------------------------------------- K5
struct MaskedLossArgs {
  const float* rendered; const float* img;      // [N,3,H,W] (img already NCHW)
  const float* depth; const float* depth_gt;    // [N,H,W]
  const float* albedo; const float* albedo_gt;  // [N,3,H,W], [N,H,W]
  const float* mask_fill; const float* mask;    // [N,H,W]: recon/albedo mask (TRAIN:612), depth mask (TRAIN:610)
  double* sums;        // [5]: 0 sum(mask_fill) ; 1 sum(mask) ; 2 recon SSE ; 3 depth SAE ; 4 albedo SAE
  float* g_rendered; float* g_depth; float* g_albedo;   // written when phase == 1 (any may be null)
  long long n_pix;     // N*H*W
  long long hw;
  int phase;           // 0: mask sums only ; 1: losses + gradients (reads sums[0..1])
};

__global__ void __launch_bounds__(256) masked_losses_kernel(const MaskedLossArgs a) {
  __shared__ double s_red[3][8];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double v[3] = {0.0, 0.0, 0.0};
  if (i < a.n_pix) {
    const long long n = i / a.hw, p = i % a.hw;
    const float mf = __ldg(a.mask_fill + i), m = __ldg(a.mask + i);
    if (a.phase == 0) {
      v[0] = (double)mf; v[1] = (double)m;
    } else {
      const double s_mf3 = 3.0 * a.sums[0], s_mf = a.sums[0], s_m = a.sums[1];      // sum over the 3-channel mask = 3 * sum(mask)
      float alb_mean = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const long long o = (n * 3 + c) * a.hw + p;
        const float d = mf * __ldg(a.rendered + o) - mf * __ldg(a.img + o);         // TRAIN:633
        v[0] += (double)d * (double)d;
        if (a.g_rendered) a.g_rendered[o] = (float)(20.0 * 2.0 * (double)d * (double)mf / s_mf3);
        alb_mean += __ldg(a.albedo + o);
      }
      alb_mean *= (1.0f / 3.0f);
      const float dd = __ldg(a.depth + i) * m - __ldg(a.depth_gt + i) * m;          // TRAIN:634
      v[1] = fabs((double)dd);
      if (a.g_depth) a.g_depth[i] = (float)((dd > 0.f ? 1.0 : (dd < 0.f ? -1.0 : 0.0)) * (double)m / s_m);
      const float da = alb_mean * mf - __ldg(a.albedo_gt + i) * mf;                 // TRAIN:637-639
      v[2] = fabs((double)da);
      if (a.g_albedo) {
        const float ga = (float)(5.0 * (da > 0.f ? 1.0 : (da < 0.f ? -1.0 : 0.0)) * (double)mf / s_mf / 3.0);
#pragma unroll
        for (int c = 0; c < 3; ++c) a.g_albedo[(n * 3 + c) * a.hw + p] = ga;
      }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    double w = v[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if (lane == 0) s_red[q][warp] = w;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_red[threadIdx.x][w];
    const int dst = a.phase == 0 ? threadIdx.x : 2 + threadIdx.x;
    if (!(a.phase == 0 && threadIdx.x == 2) && t != 0.0) atomicAdd(a.sums + dst, t);
  }
}

// ------------------------------------------------------------------------------------------------- Adam
// step counter on the device (so the whole training step can live in a CUDA graph): state = {step, 1-beta1^step, sqrt(1-beta2^step)}
__global__ void adam_tick_kernel(float* __restrict__ state, float beta1, float beta2) {
  const float t = state[0] + 1.0f;
  state[0] = t;
  state[1] = 1.0f - powf(beta1, t);
  state[2] = sqrtf(1.0f - powf(beta2, t));
}

__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                 long long n, float lr, float beta1, float beta2, float eps, const float* __restrict__ state,
                                 float grad_scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float bc1 = __ldg(state + 1), bc2_sqrt = __ldg(state + 2);
  const float gi = g[i] * grad_scale;
  const float mi = beta1 * m[i] + (1.f - beta1) * gi;          // torch.optim.Adam (no amsgrad, no weight decay)
  const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

// Per-parameter step counts (torch.optim.Adam keeps `state[p]['step']` per parameter and skips parameters whose .grad is None):
// seg_state[s] = {step, 1-beta1^step, sqrt(1-beta2^step), active}; only ACTIVE segments tick and are updated.
__global__ void adam_tick_segments_kernel(float* __restrict__ seg_state, int n_seg, float beta1, float beta2) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  float* st = seg_state + 4 * s;
  if (st[3] == 0.f) return;
  const double t = (double)st[0] + 1.0;
  st[0] = (float)t;
  st[1] = (float)(1.0 - pow((double)beta1, t));
  st[2] = (float)sqrt(1.0 - pow((double)beta2, t));
}

__global__ void adam_step_segments_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                          float* __restrict__ v, long long n, const long long* __restrict__ seg_start,
                                          const float* __restrict__ seg_state, int n_seg, float lr, float beta1, float beta2,
                                          float eps, float grad_scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lo = 0, hi = n_seg;                       // the segment with seg_start[lo] <= i < seg_start[lo + 1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(seg_start + mid) <= i) lo = mid; else hi = mid;
  }
  const float4 st = __ldg(reinterpret_cast<const float4*>(seg_state) + lo);
  if (st.w == 0.f) return;                      // no gradient this step (torch: p.grad is None -> skipped, state untouched)
  const float gi = g[i] * grad_scale;
  const float mi = beta1 * m[i] + (1.f - beta1) * gi;
  const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / st.z + eps;
  p[i] -= (lr / st.y) * (mi / denom);
}

SsimWeights gaussian_window() {
  SsimWeights g;
  float s = 0.f;
  for (int i = 0; i < WIN; ++i) {
    const float c = (float)(i - WIN / 2);
    g.w[i] = expf(-(c * c) / (2.0f * 1.5f * 1.5f));
    s += g.w[i];
  }
  for (int i = 0; i < WIN; ++i) g.w[i] /= s;
  return g;
}

}  // namespace

extern "C" int gfr_ssim_fwd(const float* X, const float* Y, double* plane_sums, float* grad_maps, int P, int H, int W,
                            float data_range, void* stream) {
  GFR_RETURN_IF_NULL(X); GFR_RETURN_IF_NULL(Y); GFR_RETURN_IF_NULL(plane_sums);
  if (P <= 0 || P > 65535 || H < WIN || W < WIN) return GFR_E_SHAPE;
  const float C1 = (0.01f * data_range) * (0.01f * data_range), C2 = (0.03f * data_range) * (0.03f * data_range);
  SsimFwdArgs a{X, Y, plane_sums, grad_maps, P, H, W, C1, C2};
  const dim3 grid(gfr_ceil_div(W - WIN + 1, TS) * gfr_ceil_div(H - WIN + 1, TS), P);
  ssim_fwd_kernel<<<grid, TS * TS, 0, (cudaStream_t)stream>>>(a, gaussian_window());
  return gfr_launch_status();
}

extern "C" int gfr_ssim_bwd(const float* X, const float* Y, const float* grad_maps, const float* plane_scale, float* g_X, int P,
                            int H, int W, void* stream) {
  GFR_RETURN_IF_NULL(X); GFR_RETURN_IF_NULL(Y); GFR_RETURN_IF_NULL(grad_maps); GFR_RETURN_IF_NULL(plane_scale);
  GFR_RETURN_IF_NULL(g_X);
  if (P <= 0 || P > 65535 || H < WIN || W < WIN) return GFR_E_SHAPE;
  SsimBwdArgs a{X, Y, grad_maps, plane_scale, g_X, P, H, W};
  const dim3 grid(gfr_ceil_div(W, TS) * gfr_ceil_div(H, TS), P);
  ssim_bwd_kernel<<<grid, TS * TS, 0, (cudaStream_t)stream>>>(a, gaussian_window());
  return gfr_launch_status();
}

extern "C" int gfr_masked_losses(const float* rendered, const float* img_nchw, const float* depth, const float* depth_gt,
                                 const float* albedo, const float* albedo_gt, const float* mask_fill, const float* mask,
                                 double* sums5, float* g_rendered, float* g_depth, float* g_albedo, int N, int H, int W,
                                 void* stream) {
  GFR_RETURN_IF_NULL(rendered); GFR_RETURN_IF_NULL(img_nchw); GFR_RETURN_IF_NULL(depth); GFR_RETURN_IF_NULL(depth_gt);
  GFR_RETURN_IF_NULL(albedo); GFR_RETURN_IF_NULL(albedo_gt); GFR_RETURN_IF_NULL(mask_fill); GFR_RETURN_IF_NULL(mask);
  GFR_RETURN_IF_NULL(sums5);
  if (N <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  MaskedLossArgs a{rendered, img_nchw, depth, depth_gt, albedo, albedo_gt, mask_fill, mask, sums5, g_rendered, g_depth, g_albedo,
                   (long long)N * H * W, (long long)H * W, 0};
  const unsigned blocks = (unsigned)((a.n_pix + 255) / 256);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(sums5, 0, 5 * sizeof(double), s);
  if (e != cudaSuccess) return (int)e;
  masked_losses_kernel<<<blocks, 256, 0, s>>>(a);
  a.phase = 1;
  masked_losses_kernel<<<blocks, 256, 0, s>>>(a);
  return gfr_launch_status();
}

extern "C" int gfr_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float* state3,
                             float lr, float beta1, float beta2, float eps, float grad_scale, void* stream) {
  GFR_RETURN_IF_NULL(params); GFR_RETURN_IF_NULL(grads); GFR_RETURN_IF_NULL(exp_avg); GFR_RETURN_IF_NULL(exp_avg_sq);
  GFR_RETURN_IF_NULL(state3);
  if (n <= 0) return GFR_E_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  adam_tick_kernel<<<1, 1, 0, s>>>(state3, beta1, beta2);
  adam_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, state3,
                                                              grad_scale);
  return gfr_launch_status();
}

extern "C" int gfr_adam_step_segments(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                                      const long long* seg_start, float* seg_state, int n_seg, float lr, float beta1, float beta2,
                                      float eps, float grad_scale, void* stream) {
  GFR_RETURN_IF_NULL(params); GFR_RETURN_IF_NULL(grads); GFR_RETURN_IF_NULL(exp_avg); GFR_RETURN_IF_NULL(exp_avg_sq);
  GFR_RETURN_IF_NULL(seg_start); GFR_RETURN_IF_NULL(seg_state);
  if (n <= 0 || n_seg <= 0) return GFR_E_SHAPE;
  if (reinterpret_cast<uintptr_t>(seg_state) & 15) return GFR_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  adam_tick_segments_kernel<<<(unsigned)((n_seg + 127) / 128), 128, 0, s>>>(seg_state, n_seg, beta1, beta2);
  adam_step_segments_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(params, grads, exp_avg, exp_avg_sq, n, seg_start, seg_state,
                                                                       n_seg, lr, beta1, beta2, eps, grad_scale);
  return gfr_launch_status();
}
