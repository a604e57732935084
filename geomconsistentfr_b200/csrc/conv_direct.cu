// fp32 direct convolution (CUDA cores) with fused epilogues — the exact-fp32 CNN path of RelightNet.
// Replaces the cuDNN conv + BatchNorm(eval) + LeakyReLU + residual/skip adds + nearest upsample of
// TRAIN:197-350 / TEST1:170-323.  BatchNorm (eval) is folded into weight/bias by the host.
//
// Tiling: a CTA computes TW x TH output pixels x CO_T output channels; each thread owns a 1x4 pixel strip
// x CO_T channels in registers.  Input channels are streamed through shared memory CI_T at a time
// (halo tile + the matching weight slice).  Inner loop per (ci, ky): one LDS.128 (+ K-1 scalars) of input,
// K*CO_T/4 broadcast LDS.128 of weights, 4*K*CO_T FFMA.
#include "gfr_common.cuh"

namespace {

struct ConvArgs {
  const float* in;   long long in_sN, in_sC, in_sH, in_sW;   // element strides (NHWC input / channel slices)
  const float* w;    // [Cout, Cin, K, K]  (BN folded)
  const float* bias; // [Cout]             (BN folded)
  const float* res;  // pre-activation residual [N,Cout,H,W] or null
  const float* post; // post-activation add [N,Cout,H>>post_shift,W>>post_shift] or null
  float* out;        // [N,Cout,H,W]
  int N, Cin, Cout, H, W;
  int ups_in;        // 1: `in` is [.., H/2, W/2] and is nearest-upsampled x2 on the fly
  int post_shift;    // 0 or 1
  int act;           // 0 none, 1 LeakyReLU(0.2), 2 sigmoid
  float out_scale;   // applied last
};

constexpr int PX = 4;     // pixels per thread along x
constexpr int CI_T = 8;   // input channels per shared-memory stage

template <int K, int CO_T, int TW, int TH>
__global__ void __launch_bounds__((TW / PX) * TH) conv2d_fwd_kernel(const ConvArgs a) {
  constexpr int PAD = K / 2;
  constexpr int IH = TH + K - 1;
  constexpr int IW = TW + K - 1;
  constexpr int IWP = (IW + 3) & ~3;                 // row pitch (floats), keeps float4 alignment
  constexpr int TX = TW / PX;
  constexpr int NT = TX * TH;
  constexpr int KK = K * K;
  __shared__ __align__(16) float s_in[CI_T][IH][IWP];
  __shared__ __align__(16) float s_w[CI_T][KK][CO_T];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int tiles_x = (a.W + TW - 1) / TW;
  const int x0 = (blockIdx.x % tiles_x) * TW, y0 = (blockIdx.x / tiles_x) * TH;
  const int co0 = blockIdx.y * CO_T;
  const int n = blockIdx.z;
  const float* __restrict__ in = a.in + (long long)n * a.in_sN;

  float acc[CO_T][PX];
#pragma unroll
  for (int c = 0; c < CO_T; ++c)
#pragma unroll
    for (int p = 0; p < PX; ++p) acc[c][p] = 0.f;

  for (int ci0 = 0; ci0 < a.Cin; ci0 += CI_T) {
    __syncthreads();
    // ---- stage the input halo tile (zero padding outside the image / past Cin)
    for (int r = tid / 32; r < CI_T * IH; r += NT / 32) {
      const int ci = r / IH, iy = r % IH;
      const int gy = y0 + iy - PAD;
      const bool rowok = (ci0 + ci < a.Cin) && gy >= 0 && gy < a.H;
      const float* src = in + (long long)(ci0 + ci) * a.in_sC + (long long)((rowok ? gy : 0) >> a.ups_in) * a.in_sH;
      for (int ix = tid % 32; ix < IW; ix += 32) {
        const int gx = x0 + ix - PAD;
        float v = 0.f;
        if (rowok && gx >= 0 && gx < a.W) v = __ldg(src + (long long)(gx >> a.ups_in) * a.in_sW);
        s_in[ci][iy][ix] = v;
      }
    }
    // ---- stage the weight slice  w[co0+co][ci0+ci][tap] -> s_w[ci][tap][co]
    for (int i = tid; i < CI_T * KK * CO_T; i += NT) {
      const int co = i / (CI_T * KK), rem = i % (CI_T * KK);
      const int ci = rem / KK, tap = rem % KK;
      float v = 0.f;
      if (co0 + co < a.Cout && ci0 + ci < a.Cin) v = __ldg(a.w + ((long long)(co0 + co) * a.Cin + ci0 + ci) * KK + tap);
      s_w[ci][tap][co] = v;
    }
    __syncthreads();
#pragma unroll 2
    for (int ci = 0; ci < CI_T; ++ci) {
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        float r[PX + K - 1];
        const float* row = &s_in[ci][ty + ky][tx * PX];
        const float4 r4 = *reinterpret_cast<const float4*>(row);
        r[0] = r4.x; r[1] = r4.y; r[2] = r4.z; r[3] = r4.w;
#pragma unroll
        for (int j = 0; j < K - 1; ++j) r[PX + j] = row[PX + j];
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const float* wp = &s_w[ci][ky * K + kx][0];
#pragma unroll
          for (int c4 = 0; c4 < CO_T; c4 += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wp + c4);
#pragma unroll
            for (int p = 0; p < PX; ++p) {
              acc[c4 + 0][p] = fmaf(w4.x, r[p + kx], acc[c4 + 0][p]);
              acc[c4 + 1][p] = fmaf(w4.y, r[p + kx], acc[c4 + 1][p]);
              acc[c4 + 2][p] = fmaf(w4.z, r[p + kx], acc[c4 + 2][p]);
              acc[c4 + 3][p] = fmaf(w4.w, r[p + kx], acc[c4 + 3][p]);
            }
          }
        }
      }
    }
  }

  // ---- epilogue: bias, residual, activation, post add, scale
  const int oy = y0 + ty;
  if (oy >= a.H) return;
  const long long plane = (long long)a.H * a.W;
  const int pH = a.H >> a.post_shift, pW = a.W >> a.post_shift;
#pragma unroll
  for (int c = 0; c < CO_T; ++c) {
    const int co = co0 + c;
    if (co >= a.Cout) break;
    const float bia = __ldg(a.bias + co);
    const long long obase = ((long long)n * a.Cout + co) * plane + (long long)oy * a.W;
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      const int ox = x0 + tx * PX + p;
      if (ox >= a.W) continue;
      float v = acc[c][p] + bia;
      if (a.res) v += __ldg(a.res + obase + ox);
      if (a.act == 1) v = v > 0.f ? v : 0.2f * v;
      else if (a.act == 2) v = 1.0f / (1.0f + expf(-v));
      if (a.post) v += __ldg(a.post + ((long long)n * a.Cout + co) * pH * pW + (long long)(oy >> a.post_shift) * pW + (ox >> a.post_shift));
      a.out[obase + ox] = v * a.out_scale;
    }
  }
}

template <int K, int CO_T, int TW, int TH>
int launch_conv(const ConvArgs& a, cudaStream_t s) {
  const dim3 grid(gfr_ceil_div(a.W, TW) * gfr_ceil_div(a.H, TH), gfr_ceil_div(a.Cout, CO_T), a.N);
  conv2d_fwd_kernel<K, CO_T, TW, TH><<<grid, (TW / PX) * TH, 0, s>>>(a);
  return gfr_launch_status();
}

template <int K>
int dispatch_conv(const ConvArgs& a, cudaStream_t s) {
  // Pick the tile so that the grid has at least ~2 CTAs per SM where the layer allows it.
  const long long px = (long long)a.N * a.H * a.W;
  if (a.Cout <= 4) return launch_conv<K, 4, 64, 8>(a, s);
  if (a.W >= 64 && px * gfr_ceil_div(a.Cout, 16) >= 148LL * 2 * 64 * 8) {
    if (K == 5) return launch_conv<K, 16, 64, 8>(a, s);
    return launch_conv<K, 16, 64, 8>(a, s);
  }
  if (a.W >= 32) return launch_conv<K, 8, 32, 8>(a, s);
  return launch_conv<K, 4, 16, 8>(a, s);
}

// ---- 2x2 max pool (TRAIN:201,206,212,218) ---------------------------------------------------------
__global__ void maxpool2_kernel(const float* __restrict__ in, float* __restrict__ out, long long n_out, int Ho, int Wo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const int x = (int)(i % Wo);
  const long long t = i / Wo;
  const int y = (int)(t % Ho);
  const long long nc = t / Ho;
  const float* p = in + (nc * (2 * Ho) + 2 * y) * (2LL * Wo) + 2 * x;
  const float2 a = *reinterpret_cast<const float2*>(p);
  const float2 b = *reinterpret_cast<const float2*>(p + 2 * Wo);
  out[i] = fmaxf(fmaxf(a.x, a.y), fmaxf(b.x, b.y));
}

// ---- nearest x2 upsample (+ optional add), TRAIN:240-246 when the epoch gate keeps the skip branch off
__global__ void upsample2_kernel(const float* __restrict__ in, const float* __restrict__ add, float* __restrict__ out,
                                 long long n_out, int Ho, int Wo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const int x = (int)(i % Wo);
  const long long t = i / Wo;
  const int y = (int)(t % Ho);
  const long long nc = t / Ho;
  float v = __ldg(in + (nc * (Ho >> 1) + (y >> 1)) * (long long)(Wo >> 1) + (x >> 1));
  if (add) v += __ldg(add + i);
  out[i] = v;
}

// ---- light head: global average pool of the lighting features + Linear 27->128 + LReLU + Linear 128->4
//      (TRAIN:225-232).  One CTA (128 threads) per image.
__global__ void __launch_bounds__(128) light_head_kernel(const float* __restrict__ feat, long long sN, int c_first, int HW,
                                                          const float* __restrict__ w1, const float* __restrict__ b1,
                                                          const float* __restrict__ w2, const float* __restrict__ b2,
                                                          float* __restrict__ out) {
  __shared__ float s_pool[27];
  __shared__ float s_h[128];
  const int n = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* f = feat + (long long)n * sN + (long long)c_first * HW;
  for (int c = warp; c < 27; c += 4) {
    float s = 0.f;
    for (int i = lane; i < HW; i += 32) s += __ldg(f + (long long)c * HW + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_pool[c] = s / (float)HW;
  }
  __syncthreads();
  float h = __ldg(b1 + tid);
  for (int c = 0; c < 27; ++c) h = fmaf(__ldg(w1 + tid * 27 + c), s_pool[c], h);
  s_h[tid] = h > 0.f ? h : 0.2f * h;
  __syncthreads();
  if (warp < 4) {
    float s = 0.f;
    for (int i = lane; i < 128; i += 32) s = fmaf(__ldg(w2 + warp * 128 + i), s_h[i], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[n * 4 + warp] = s + __ldg(b2 + warp);
  }
}

}  // namespace

extern "C" int gfr_conv2d_fwd(const float* in, const long long* in_strides_host, const float* w, const float* bias,
                              const float* res, const float* post, float* out, int N, int Cin, int Cout, int H, int W,
                              int K, int ups_in, int post_shift, int act, float out_scale, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(in_strides_host); GFR_RETURN_IF_NULL(w); GFR_RETURN_IF_NULL(bias);
  GFR_RETURN_IF_NULL(out);
  if (N <= 0 || N > 65535 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  if (ups_in < 0 || ups_in > 1 || post_shift < 0 || post_shift > 1 || act < 0 || act > 2) return GFR_E_ARG;
  if (ups_in && ((H | W) & 1)) return GFR_E_SHAPE;
  ConvArgs a{in, in_strides_host[0], in_strides_host[1], in_strides_host[2], in_strides_host[3], w, bias, res, post, out,
             N, Cin, Cout, H, W, ups_in, post_shift, act, out_scale};
  cudaStream_t s = (cudaStream_t)stream;
  switch (K) {
    case 1: return dispatch_conv<1>(a, s);
    case 3: return dispatch_conv<3>(a, s);
    case 5: return dispatch_conv<5>(a, s);
    default: return GFR_E_UNSUPPORTED;
  }
}

extern "C" int gfr_maxpool2_fwd(const float* in, float* out, int NC, int Ho, int Wo, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(out);
  if (NC <= 0 || Ho <= 0 || Wo <= 0) return GFR_E_SHAPE;
  const long long n = (long long)NC * Ho * Wo;
  maxpool2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, out, n, Ho, Wo);
  return gfr_launch_status();
}

extern "C" int gfr_upsample2_fwd(const float* in, const float* add, float* out, int NC, int Ho, int Wo, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(out);
  if (NC <= 0 || Ho <= 0 || Wo <= 0 || ((Ho | Wo) & 1)) return GFR_E_SHAPE;
  const long long n = (long long)NC * Ho * Wo;
  upsample2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, add, out, n, Ho, Wo);
  return gfr_launch_status();
}

extern "C" int gfr_light_head_fwd(const float* feat, long long feat_batch_stride, int c_first, int HW, const float* w1,
                                  const float* b1, const float* w2, const float* b2, float* out, int N, void* stream) {
  GFR_RETURN_IF_NULL(feat); GFR_RETURN_IF_NULL(w1); GFR_RETURN_IF_NULL(b1); GFR_RETURN_IF_NULL(w2);
  GFR_RETURN_IF_NULL(b2); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || HW <= 0) return GFR_E_SHAPE;
  light_head_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(feat, feat_batch_stride, c_first, HW, w1, b1, w2, b2, out);
  return gfr_launch_status();
}
