// CUDA-core companions of the tensor-core CNN path (C4 activation layout, see conv_tc.cu):
//   * stem   — conv_c1_og 5x5, 3 -> 16 channels on the NHWC input image + BN(eval) + LeakyReLU, with the first
//              2x2 max pool fused (TRAIN:197-201; TRAIN = train_raytracing_relighting_CelebAHQ_DSSIM_8x.py).  K = 75 is
//              too small and too misaligned (3 channels) for UMMA operands; this is an FFMA-bound register-tile kernel
//              whose weights are FFMA constant-bank operands (they travel as a kernel parameter).
//   * head   — the 1x1 tail of each decoder: c2_2, c2_3 (16 -> 16, BN + LeakyReLU) and c2_o (16 -> 3 sigmoid | 16 -> 1
//              x100) fused per pixel, C4 in, NCHW planes out (TRAIN:285-290, 345-350).
//   * light  — global average pool of the 27 lighting channels + the 2-layer MLP (TRAIN:225-232) on a C4 feature map.
#include "gfr_common.cuh"
#include "p16.cuh"
#include "head_device.cuh"

#include <stdlib.h>

namespace {

// ------------------------------------------------------------------------------------------------- stem
struct StemWeights { float w[25][3][16]; float b[16]; };      // [tap][ci][co], BN folded

struct StemArgs {
  const float* img;   // [N,H,W,3]
  float* out;         // C4 [N,4,H,W,4]
  float* pooled;      // C4 [N,4,H/2,W/2,4] or null
  int N, H, W;
  int p16;            // 1: out / pooled are P16 tensors [N,2,2,H,W,8] halfs (pre-split fp16 pairs, p16.cuh) instead of C4
};

constexpr int ST_TW = 32, ST_TH = 16;                          // CTA tile (pixels); thread = 2x2 pixels
constexpr int ST_IW = ST_TW + 4, ST_IH = ST_TH + 4;

// ROLLED = false: everything unrolled, weights as FFMA uniform-register operands (LDCU from the constant bank): 27k SASS
// instructions (430 KB) - ncu: the top stall is `no_instruction` (instruction-cache misses, 2.8 per issue), FMA pipe 42 %.
// ROLLED = true: the (ci, ky) loops stay rolled (a ~400-instruction body), weights come from shared memory as LDS.128
// broadcasts; same accumulation order (ci, ky, kx), so the result is bit-identical.
template <bool ROLLED>
__global__ void __launch_bounds__(128) stem_conv_kernel(const StemArgs a, const __grid_constant__ StemWeights wt) {
  __shared__ __align__(16) float s_in[3][ST_IH][ST_IW];
  __shared__ __align__(16) float s_w[ROLLED ? 75 : 1][16];
  const int tid = threadIdx.x;
  if (ROLLED) {
    const float* wp = &wt.w[0][0][0];
    for (int i = tid; i < 75 * 16; i += 128) s_w[i >> 4][i & 15] = wp[i];
  }
  const int tiles_x = gfr_ceil_div(a.W, ST_TW);
  const int x0 = (blockIdx.x % tiles_x) * ST_TW, y0 = (blockIdx.x / tiles_x) * ST_TH;
  const int n = blockIdx.y;
  const float* __restrict__ img = a.img + (size_t)n * a.H * a.W * 3;

  // stage the (TH+4) x (TW+4) x 3 halo tile; NHWC rows are contiguous runs of 3*(TW+4) floats
  for (int i = tid; i < ST_IH * ST_IW * 3; i += 128) {
    const int r = i / (ST_IW * 3), rem = i % (ST_IW * 3);
    const int c = rem / 3, ch = rem % 3;
    const int gy = y0 + r - 2, gx = x0 + c - 2;
    float v = 0.f;
    if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) v = __ldg(img + ((size_t)gy * a.W + gx) * 3 + ch);
    s_in[ch][r][c] = v;
  }
  __syncthreads();

  const int tx = tid % (ST_TW / 2), ty = tid / (ST_TW / 2);    // 16 x 8 threads
  float acc[4][16];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[p][c] = wt.b[c];

  if constexpr (ROLLED) {
#pragma unroll 1
    for (int ci = 0; ci < 3; ++ci) {
#pragma unroll 1
      for (int ky = 0; ky < 5; ++ky) {
        float r0[6], r1[6];
#pragma unroll
        for (int c = 0; c < 6; c += 2) {
          const float2 u = *reinterpret_cast<const float2*>(&s_in[ci][2 * ty + ky][2 * tx + c]);
          const float2 v = *reinterpret_cast<const float2*>(&s_in[ci][2 * ty + ky + 1][2 * tx + c]);
          r0[c] = u.x; r0[c + 1] = u.y; r1[c] = v.x; r1[c + 1] = v.y;
        }
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) {
          const float4* w4 = reinterpret_cast<const float4*>(&s_w[(ky * 5 + kx) * 3 + ci][0]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 w = w4[q];
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              acc[0][4 * q + e] = fmaf(wv[e], r0[kx], acc[0][4 * q + e]);
              acc[1][4 * q + e] = fmaf(wv[e], r0[kx + 1], acc[1][4 * q + e]);
              acc[2][4 * q + e] = fmaf(wv[e], r1[kx], acc[2][4 * q + e]);
              acc[3][4 * q + e] = fmaf(wv[e], r1[kx + 1], acc[3][4 * q + e]);
            }
          }
        }
      }
    }
  } else {
#pragma unroll
  for (int ci = 0; ci < 3; ++ci) {
    float win[6][6];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < 6; c += 2) {
        const float2 v = *reinterpret_cast<const float2*>(&s_in[ci][2 * ty + r][2 * tx + c]);
        win[r][c] = v.x; win[r][c + 1] = v.y;
      }
#pragma unroll
    for (int ky = 0; ky < 5; ++ky)
#pragma unroll
      for (int kx = 0; kx < 5; ++kx)
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float w = wt.w[ky * 5 + kx][ci][c];
          acc[0][c] = fmaf(w, win[ky][kx], acc[0][c]);
          acc[1][c] = fmaf(w, win[ky][kx + 1], acc[1][c]);
          acc[2][c] = fmaf(w, win[ky + 1][kx], acc[2][c]);
          acc[3][c] = fmaf(w, win[ky + 1][kx + 1], acc[3][c]);
        }
  }
  }

  const int oy = y0 + 2 * ty, ox = x0 + 2 * tx;
  if (oy >= a.H || ox >= a.W) return;
  const size_t plane = (size_t)a.H * a.W;
  if (a.p16) {
    // P16: 8-channel chunks, hi and lo units one plane apart
    __half* out = reinterpret_cast<__half*>(a.out);
    __half* pooled = reinterpret_cast<__half*>(a.pooled);
#pragma unroll
    for (int c8 = 0; c8 < 2; ++c8) {
      float v[4][8], mx[8];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float z = acc[p][c8 * 8 + e]; v[p][e] = z > 0.f ? z : 0.2f * z; }
#pragma unroll
      for (int e = 0; e < 8; ++e) mx[e] = fmaxf(fmaxf(v[0][e], v[1][e]), fmaxf(v[2][e], v[3][e]));
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int yy = oy + (p >> 1), xx = ox + (p & 1);
        if (yy >= a.H || xx >= a.W) continue;
        uint4 hi, lo;
        gfr_p16::split8(v[p], hi, lo);
        __half* o = out + gfr_p16::unit_offset(n, 2, c8, a.H, a.W, yy, xx);
        *reinterpret_cast<uint4*>(o) = hi;
        *reinterpret_cast<uint4*>(o + plane * 8) = lo;
      }
      if (pooled) {
        uint4 hi, lo;
        gfr_p16::split8(mx, hi, lo);
        __half* o = pooled + gfr_p16::unit_offset(n, 2, c8, a.H >> 1, a.W >> 1, oy >> 1, ox >> 1);
        *reinterpret_cast<uint4*>(o) = hi;
        *reinterpret_cast<uint4*>(o + (plane >> 2) * 8) = lo;
      }
    }
    return;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float4 v[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float t[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float z = acc[p][q * 4 + e]; t[e] = z > 0.f ? z : 0.2f * z; }
      v[p] = make_float4(t[0], t[1], t[2], t[3]);
    }
    float* o = a.out + (((size_t)n * 4 + q) * plane + (size_t)oy * a.W + ox) * 4;
    *reinterpret_cast<float4*>(o) = v[0];
    if (ox + 1 < a.W) *reinterpret_cast<float4*>(o + 4) = v[1];
    if (oy + 1 < a.H) {
      *reinterpret_cast<float4*>(o + (size_t)a.W * 4) = v[2];
      if (ox + 1 < a.W) *reinterpret_cast<float4*>(o + (size_t)a.W * 4 + 4) = v[3];
    }
    if (a.pooled) {      // H, W even: the 2x2 quad is complete
      const float4 m = make_float4(fmaxf(fmaxf(v[0].x, v[1].x), fmaxf(v[2].x, v[3].x)), fmaxf(fmaxf(v[0].y, v[1].y), fmaxf(v[2].y, v[3].y)),
                                   fmaxf(fmaxf(v[0].z, v[1].z), fmaxf(v[2].z, v[3].z)), fmaxf(fmaxf(v[0].w, v[1].w), fmaxf(v[2].w, v[3].w)));
      *reinterpret_cast<float4*>(a.pooled + (((size_t)n * 4 + q) * (plane >> 2) + (size_t)(oy >> 1) * (a.W >> 1) + (ox >> 1)) * 4) = m;
    }
  }
}

// ------------------------------------------------------------------------------------------------- head
using gfr_head::HeadWeights;

struct HeadArgs {
  const float* in;    // C4 [N,4,H,W,4]
  float* out;         // [N,n_out,H,W]
  long long hw;       // H*W
  long long total;    // N*H*W
  int n_out;          // 1 or 3
  int act;            // 0 none, 2 sigmoid
  float scale;
  int p16;            // 1: `in` is a P16 tensor [N,2,2,H,W,8] halfs
};

__global__ void __launch_bounds__(256) head_1x1_kernel(const HeadArgs a, const __grid_constant__ HeadWeights wt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.total) return;
  const long long n = i / a.hw, p = i % a.hw;
  const float4* src = reinterpret_cast<const float4*>(a.in) + n * 4 * a.hw + p;
  float x[16];
  if (a.p16) {
    const __half* hp = reinterpret_cast<const __half*>(a.in) + ((size_t)n * 4 * a.hw + p) * 8;      // [n][c8][part][hw][8]
#pragma unroll
    for (int c8 = 0; c8 < 2; ++c8) {
      float v[8];
      gfr_p16::join8(__ldg(reinterpret_cast<const uint4*>(hp + (size_t)(2 * c8) * a.hw * 8)),
                     __ldg(reinterpret_cast<const uint4*>(hp + (size_t)(2 * c8 + 1) * a.hw * 8)), v);
#pragma unroll
      for (int e = 0; e < 8; ++e) x[c8 * 8 + e] = v[e];
    }
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 v = __ldg(src + q * a.hw);
      x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
    }
  }
  float o3[3];
  gfr_head::apply(wt, x, a.n_out, a.act, a.scale, o3);
#pragma unroll
  for (int o = 0; o < 3; ++o)
    if (o < a.n_out) a.out[(n * a.n_out + o) * a.hw + p] = o3[o];
}

// ------------------------------------------------------------------------------------------------- light head (C4)
template <bool P16>
__global__ void __launch_bounds__(128) light_head_c4_kernel(const float* __restrict__ feat, int C4, int c_first, int HW,
                                                             const float* __restrict__ w1, const float* __restrict__ b1,
                                                             const float* __restrict__ w2, const float* __restrict__ b2,
                                                             float* __restrict__ out) {
  __shared__ float s_pool[27];
  __shared__ float s_h[128];
  const int n = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int c = warp; c < 27; c += 4) {
    const int ch = c_first + c;
    float s = 0.f;
    if (P16) {      // C4 counts 8-channel chunks here; feat is [N][C8][hi|lo][HW][8] halfs
      const __half* f = reinterpret_cast<const __half*>(feat) + (((size_t)n * C4 + (ch >> 3)) * 2 * HW) * 8 + (ch & 7);
      for (int i = lane; i < HW; i += 32)
        s += (__half2float(f[(size_t)i * 8]) + __half2float(f[((size_t)HW + i) * 8])) * gfr_p16::X_INV;
    } else {
      const float* f = feat + (((size_t)n * C4 + (ch >> 2)) * HW) * 4 + (ch & 3);
      for (int i = lane; i < HW; i += 32) s += __ldg(f + (size_t)i * 4);
    }
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

static int stem_launch(const float* img, const float* w_host, const float* bias_host, float* out, float* pooled, int N, int H, int W,
                       int p16, void* stream) {
  GFR_RETURN_IF_NULL(img); GFR_RETURN_IF_NULL(w_host); GFR_RETURN_IF_NULL(bias_host); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || N > 65535 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  if (pooled && ((H | W) & 1)) return GFR_E_SHAPE;
  StemWeights wt;
  for (int co = 0; co < 16; ++co) {
    wt.b[co] = bias_host[co];
    for (int ci = 0; ci < 3; ++ci)
      for (int t = 0; t < 25; ++t) wt.w[t][ci][co] = w_host[(co * 3 + ci) * 25 + t];
  }
  StemArgs a{img, out, pooled, N, H, W, p16};
  const dim3 grid(gfr_ceil_div(W, ST_TW) * gfr_ceil_div(H, ST_TH), N);
  // rolled is the default: 60.1 vs 62.8 us alone, +1.1 % value / +2 % e2e on the whole forward (GFR_STEM_ROLLED=0: the unrolled one)
  static const bool rolled = [] { const char* e = getenv("GFR_STEM_ROLLED"); return !(e != nullptr && e[0] == '0'); }();
  if (rolled) stem_conv_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(a, wt);
  else stem_conv_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(a, wt);
  return gfr_launch_status();
}

extern "C" int gfr_stem_conv_fwd(const float* img, const float* w_host, const float* bias_host, float* out, float* pooled,
                                 int N, int H, int W, void* stream) {
  return stem_launch(img, w_host, bias_host, out, pooled, N, H, W, 0, stream);
}

extern "C" int gfr_stem_conv_p16_fwd(const float* img, const float* w_host, const float* bias_host, void* out, void* pooled, int N,
                                     int H, int W, void* stream) {
  return stem_launch(img, w_host, bias_host, reinterpret_cast<float*>(out), reinterpret_cast<float*>(pooled), N, H, W, 1, stream);
}

static int head_launch(const float* in, const float* w2_host, const float* b2_host, const float* w3_host, const float* b3_host,
                       const float* wo_host, const float* bo_host, float* out, int N, int H, int W, int n_out, int act, float out_scale,
                       int p16, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(w2_host); GFR_RETURN_IF_NULL(b2_host); GFR_RETURN_IF_NULL(w3_host);
  GFR_RETURN_IF_NULL(b3_host); GFR_RETURN_IF_NULL(wo_host); GFR_RETURN_IF_NULL(bo_host); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  if (n_out < 1 || n_out > 3 || (act != 0 && act != 2)) return GFR_E_ARG;
  HeadWeights wt;
  gfr_head::fill(wt, w2_host, b2_host, w3_host, b3_host, wo_host, bo_host, n_out);
  HeadArgs a{in, out, (long long)H * W, (long long)N * H * W, n_out, act, out_scale, p16};
  head_1x1_kernel<<<(unsigned)((a.total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, wt);
  return gfr_launch_status();
}

extern "C" int gfr_head_1x1_fwd(const float* in, const float* w2_host, const float* b2_host, const float* w3_host,
                                const float* b3_host, const float* wo_host, const float* bo_host, float* out, int N, int H,
                                int W, int n_out, int act, float out_scale, void* stream) {
  return head_launch(in, w2_host, b2_host, w3_host, b3_host, wo_host, bo_host, out, N, H, W, n_out, act, out_scale, 0, stream);
}

extern "C" int gfr_head_1x1_p16_fwd(const void* in, const float* w2_host, const float* b2_host, const float* w3_host,
                                    const float* b3_host, const float* wo_host, const float* bo_host, float* out, int N, int H,
                                    int W, int n_out, int act, float out_scale, void* stream) {
  return head_launch(reinterpret_cast<const float*>(in), w2_host, b2_host, w3_host, b3_host, wo_host, bo_host, out, N, H, W, n_out, act,
                     out_scale, 1, stream);
}

extern "C" int gfr_light_head_c4_fwd(const float* feat, int C, int c_first, int HW, const float* w1, const float* b1,
                                     const float* w2, const float* b2, float* out, int N, void* stream) {
  GFR_RETURN_IF_NULL(feat); GFR_RETURN_IF_NULL(w1); GFR_RETURN_IF_NULL(b1); GFR_RETURN_IF_NULL(w2);
  GFR_RETURN_IF_NULL(b2); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || HW <= 0 || C <= 0 || c_first < 0 || c_first + 27 > C) return GFR_E_SHAPE;
  light_head_c4_kernel<false><<<N, 128, 0, (cudaStream_t)stream>>>(feat, (C + 3) / 4, c_first, HW, w1, b1, w2, b2, out);
  return gfr_launch_status();
}

extern "C" int gfr_light_head_p16_fwd(const void* feat, int groups, int c_first, int HW, const float* w1, const float* b1,
                                      const float* w2, const float* b2, float* out, int N, void* stream) {
  GFR_RETURN_IF_NULL(feat); GFR_RETURN_IF_NULL(w1); GFR_RETURN_IF_NULL(b1); GFR_RETURN_IF_NULL(w2);
  GFR_RETURN_IF_NULL(b2); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || HW <= 0 || groups <= 0 || c_first < 0 || c_first + 27 > groups * 8) return GFR_E_SHAPE;
  light_head_c4_kernel<true><<<N, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(feat), groups, c_first, HW, w1, b1, w2,
                                                                 b2, out);
  return gfr_launch_status();
}
