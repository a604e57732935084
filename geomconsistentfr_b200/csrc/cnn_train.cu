// Train-mode building blocks of the RelightNet CNN (the reference never calls .eval() while training, TRAIN:561-563,
// so every BatchNorm uses BATCH statistics and cannot be folded into the convolution).  C4 activation layout
// [N][C/4][H][W][4] throughout (see conv_tc.cu).  TRAIN = train_raytracing_relighting_CelebAHQ_DSSIM_8x.py.
//
//   pack      — device-side weight packing for gfr_conv3x3_tc_fwd (weights change every optimiser step): tf32 hi/lo
//               split + UMMA core-matrix layout, for Conv2d and ConvTranspose2d parameters, forward or data-gradient
//               (dgrad of a stride-1 3x3 conv is the same convolution with the transposed, flipped kernel)
//   bn_stats  — per-channel sum / sum of squares over (N,H,W), fp64 accumulation
//   bn_final  — mean, 1/sqrt(var+eps) -> per-channel scale/shift; running-statistics update (momentum 0.1, unbiased var)
//   bn_apply  — y = act(scale*x + shift + res) + up2(post)   (BatchNorm + residual + LeakyReLU + skip/upsample add)
//   bn_bwd    — reduction (sum g, sum g*xhat) and elementwise input gradient of the same block
//   wgrad     — weight gradient of the 3x3 convolution (CUDA cores, fp32)
//   pw        — 1x1 convolutions of the decoder tails (16 -> 16 raw, 16 -> 1|3 with sigmoid / scale), fwd + bwd
//   pools     — 2x2 max-pool backward, 2x2 sum (backward of the nearest x2 upsample), 27-channel global average pool
#include "gfr_common.cuh"

#include <cuda_bf16.h>
#include <string.h>

namespace {

__device__ __forceinline__ float tf32_rna_dev(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// ------------------------------------------------------------------------------------------------- pack
// TF32 (kind 0): packed[n_tile][cin_step][tap][group][hi|lo][n][4 floats];  BF16 (kind 1): packed[n_tile][cin_step][tap][8-ch chunk][n][8 bf16]
// value(o, i, tap) = w[o*so + i*si + (flip ? taps - 1 - tap : tap)]
__global__ void pack_weights_kernel(const float* __restrict__ w, float* __restrict__ packed, long long total, int O, int I,
                                    int NT, long long so, long long si, int flip, int taps) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int e = (int)(idx & 3);
  long long t = idx >> 2;
  const int n = (int)(t % NT); t /= NT;
  const int part = (int)(t & 1); t >>= 1;
  const int kc = (int)(t & 3); t >>= 2;
  const int tap = (int)(t % taps); t /= taps;
  const int ncb = (I + 15) / 16;
  const int cb = (int)(t % ncb);
  const int nt = (int)(t / ncb);
  const int o = nt * NT + n, i = cb * 16 + kc * 4 + e;
  float v = 0.f;
  if (o < O && i < I) {
    const float x = __ldg(w + o * so + i * si + (flip ? taps - 1 - tap : tap));
    const float hi = tf32_rna_dev(x);
    v = part == 0 ? hi : tf32_rna_dev(x - hi);
  }
  packed[idx] = v;
}

__global__ void pack_weights_bf16_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ packed, long long total, int O, int I,
                                         int NT, long long so, long long si, int flip, int taps) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over [n_tile][cin_step][tap][chunk 2][n][8]
  if (idx >= total) return;
  const int e = (int)(idx & 7);
  long long t = idx >> 3;
  const int n = (int)(t % NT); t /= NT;
  const int ch = (int)(t & 1); t >>= 1;
  const int tap = (int)(t % taps); t /= taps;
  const int ncb = (I + 15) / 16;
  const int cb = (int)(t % ncb);
  const int nt = (int)(t / ncb);
  const int o = nt * NT + n, i = cb * 16 + ch * 8 + e;
  float v = 0.f;
  if (o < O && i < I) v = __ldg(w + o * so + i * si + (flip ? taps - 1 - tap : tap));
  packed[idx] = __float2bfloat16_rn(v);
}

// All packed operands of a training step in ONE launch (gfr_conv_tc_pack_weights_batch): a job = one layer's forward or
// data-gradient operand; a 256-thread block belongs to exactly one job (the jobs' block ranges are prefix sums).
struct PackJob {
  const float* w; void* packed;
  long long first_block, total, so, si;
  int O, I, NT, taps, flip, bf16;
};

__global__ void __launch_bounds__(256) pack_weights_batch_kernel(const PackJob* __restrict__ jobs, int n_jobs) {
  __shared__ PackJob job;
  if (threadIdx.x == 0) {
    int lo = 0, hi = n_jobs - 1;                     // last job whose first_block <= blockIdx.x
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].first_block <= (long long)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    job = jobs[lo];
  }
  __syncthreads();
  const long long idx = ((long long)blockIdx.x - job.first_block) * 256 + threadIdx.x;
  if (idx >= job.total) return;
  const int taps = job.taps, NT = job.NT, ncb = (job.I + 15) / 16;
  if (job.bf16) {
    const int e = (int)(idx & 7);
    long long t = idx >> 3;
    const int n = (int)(t % NT); t /= NT;
    const int ch = (int)(t & 1); t >>= 1;
    const int tap = (int)(t % taps); t /= taps;
    const int cb = (int)(t % ncb), nt = (int)(t / ncb);
    const int o = nt * NT + n, i = cb * 16 + ch * 8 + e;
    float v = 0.f;
    if (o < job.O && i < job.I) v = __ldg(job.w + o * job.so + i * job.si + (job.flip ? taps - 1 - tap : tap));
    reinterpret_cast<__nv_bfloat16*>(job.packed)[idx] = __float2bfloat16_rn(v);
  } else {
    const int e = (int)(idx & 3);
    long long t = idx >> 2;
    const int n = (int)(t % NT); t /= NT;
    const int part = (int)(t & 1); t >>= 1;
    const int kc = (int)(t & 3); t >>= 2;
    const int tap = (int)(t % taps); t /= taps;
    const int cb = (int)(t % ncb), nt = (int)(t / ncb);
    const int o = nt * NT + n, i = cb * 16 + kc * 4 + e;
    float v = 0.f;
    if (o < job.O && i < job.I) {
      const float x = __ldg(job.w + o * job.so + i * job.si + (job.flip ? taps - 1 - tap : tap));
      const float hi = tf32_rna_dev(x);
      v = part == 0 ? hi : tf32_rna_dev(x - hi);
    }
    reinterpret_cast<float*>(job.packed)[idx] = v;
  }
}

// ------------------------------------------------------------------------------------------------- BN statistics
struct BnFinalizeArgs {
  const float* gamma; const float* beta; float* running_mean; float* running_var;
  float* mean_out; float* rstd_out; float* scale; float* shift;
  int C, Cpad; double count; float eps, momentum;
  long long* num_batches_tracked;     // BatchNorm2d.num_batches_tracked (+= 1 by the finalising CTA) or null
};

// running = (1 - momentum) * running + momentum * batch  (nn.BatchNorm2d, TRAIN:59 ff.; unbiased batch variance), with the rounding
// steps spelled out so that every kernel that applies it produces the same bits
__device__ __forceinline__ void bn_running_update(float* running_mean, float* running_var, int c, float mean, float var_unbiased, float momentum) {
  running_mean[c] = __fmaf_rn(momentum, mean, __fmul_rn(1.f - momentum, running_mean[c]));
  running_var[c] = __fmaf_rn(momentum, var_unbiased, __fmul_rn(1.f - momentum, running_var[c]));
}

__device__ __forceinline__ void bn_finalize_channel(const double* __restrict__ sums, const BnFinalizeArgs& f, int c) {
  if (c >= f.C) { f.scale[c] = 0.f; f.shift[c] = 0.f; f.mean_out[c] = 0.f; f.rstd_out[c] = 0.f; return; }   // padded channel slots stay 0
  const double m = sums[c] / f.count;
  double var = sums[f.Cpad + c] / f.count - m * m;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)f.eps));
  const float g = f.gamma[c];
  f.mean_out[c] = (float)m; f.rstd_out[c] = rstd;
  f.scale[c] = g * rstd;
  f.shift[c] = f.beta[c] - (float)m * g * rstd;
  if (f.running_mean) bn_running_update(f.running_mean, f.running_var, c, (float)m, (float)(var * f.count / (f.count - 1.0)), f.momentum);
}

// One more momentum update of the running statistics from the SAME batch sums (gfr_bn_running_update): what a second forward pass
// over the same input with the same weights would do to the buffers, without the pass.
__global__ void bn_running_replay_kernel(const double* __restrict__ sums, float* __restrict__ running_mean, float* __restrict__ running_var,
                                         long long* __restrict__ num_batches_tracked, int C, int Cpad, double count, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches_tracked != nullptr) *num_batches_tracked += 1;
  if (c >= C) return;
  const double m = sums[c] / count;
  double var = sums[Cpad + c] / count - m * m;
  if (var < 0.0) var = 0.0;
  bn_running_update(running_mean, running_var, c, (float)m, (float)(var * count / (count - 1.0)), momentum);
}

// The LAST CTA to finish (a ticket counter behind the sums) turns the sums into mean / rstd / scale / shift and updates the running
// statistics: one launch instead of two per BatchNorm (65 BatchNorms per generator forward).
__global__ void __launch_bounds__(256) bn_stats_kernel(const float4* __restrict__ x, double* __restrict__ sums, int N, int C4,
                                                        int HW, int chunks, const BnFinalizeArgs fin) {
  // grid (chunks, C4, N): a CTA reduces a contiguous chunk of one (n, group) plane; sums = [2][C4*4] (+ the ticket counter)
  __shared__ double s_red[8][8];
  __shared__ bool s_last;
  const int g = blockIdx.y, n = blockIdx.z;
  const int per = (HW + chunks - 1) / chunks;
  const int lo = blockIdx.x * per, hi = min(lo + per, HW);
  const float4* p = x + ((size_t)n * C4 + g) * HW;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  double ds[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int cnt = 0;
  for (int i = lo + threadIdx.x; i < hi; i += 256) {
    const float4 v = __ldg(p + i);
    s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    q[0] += v.x * v.x; q[1] += v.y * v.y; q[2] += v.z * v.z; q[3] += v.w * v.w;
    if (++cnt == 32) {
#pragma unroll
      for (int e = 0; e < 4; ++e) { ds[e] += s[e]; ds[4 + e] += q[e]; s[e] = 0.f; q[e] = 0.f; }
      cnt = 0;
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) { ds[e] += s[e]; ds[4 + e] += q[e]; }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    double v = ds[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[e][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_red[threadIdx.x][w];
    const int e = threadIdx.x & 3, which = threadIdx.x >> 2;
    atomicAdd(sums + (size_t)which * C4 * 4 + g * 4 + e, t);
  }
  // ticket: the sums of every CTA are visible to the one that draws the last number
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* ticket = reinterpret_cast<unsigned int*>(sums + (size_t)2 * C4 * 4);
    const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
    s_last = atomicAdd(ticket, 1u) == total - 1;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int c = threadIdx.x; c < fin.Cpad; c += 256) bn_finalize_channel(sums, fin, c);
    if (threadIdx.x == 0 && fin.num_batches_tracked != nullptr) *fin.num_batches_tracked += 1;     // torch.nn.BatchNorm2d does this per forward
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ mean_out,
                                   float* __restrict__ rstd_out, float* __restrict__ scale, float* __restrict__ shift, int C, int Cpad,
                                   double count, float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cpad) return;
  if (c >= C) { scale[c] = 0.f; shift[c] = 0.f; mean_out[c] = 0.f; rstd_out[c] = 0.f; return; }   // padded channel slots stay 0
  const double m = sums[c] / count;
  double var = sums[Cpad + c] / count - m * m;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma[c];
  mean_out[c] = (float)m; rstd_out[c] = rstd;
  scale[c] = g * rstd;
  shift[c] = beta[c] - (float)m * g * rstd;
  if (running_mean) bn_running_update(running_mean, running_var, c, (float)m, (float)(var * count / (count - 1.0)), momentum);
}

struct BnApplyArgs {
  const float4* x; const float4* res; const float4* post; float4* y;
  const float* scale; const float* shift;      // [C4*4]
  int C4, H, W, post_shift, act;
  long long total;                              // N*C4*H*W
  int bpp;                                      // 256-element blocks per (n, group) plane when H*W % 256 == 0, else 0
};

__global__ void __launch_bounds__(256) bn_apply_kernel(const BnApplyArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.total) return;
  const int HW = a.H * a.W;
  // plane = (n, group) index, p = pixel inside it.  H*W % 256 == 0 (every layer of the two networks): a CTA sits in ONE plane and
  // the split is a 32-bit division per CTA; the generic form costs two 64-bit divisions per thread (~150 instructions against
  // the ~40 of the rest of the kernel — measured issue-bound next to its HBM time)
  unsigned plane; int p;
  if (a.bpp > 0) { plane = blockIdx.x / (unsigned)a.bpp; p = (int)(blockIdx.x - plane * (unsigned)a.bpp) * 256 + (int)threadIdx.x; }
  else { plane = (unsigned)(i / HW); p = (int)(i - (long long)plane * HW); }
  const int g = (int)(plane % (unsigned)a.C4);
  const float4 sc = __ldg(reinterpret_cast<const float4*>(a.scale) + g), sh = __ldg(reinterpret_cast<const float4*>(a.shift) + g);
  const float4 v = __ldg(a.x + i);
  float r[4] = {fmaf(sc.x, v.x, sh.x), fmaf(sc.y, v.y, sh.y), fmaf(sc.z, v.z, sh.z), fmaf(sc.w, v.w, sh.w)};
  if (a.res) { const float4 t = __ldg(a.res + i); r[0] += t.x; r[1] += t.y; r[2] += t.z; r[3] += t.w; }
  if (a.act == 1) {
#pragma unroll
    for (int e = 0; e < 4; ++e) r[e] = r[e] > 0.f ? r[e] : 0.2f * r[e];
  }
  if (a.post) {
    const int y = p / a.W, x = p - y * a.W;
    const long long nc = plane;
    const int pW = a.W >> a.post_shift, pH = a.H >> a.post_shift;
    const float4 t = __ldg(a.post + (nc * pH + (y >> a.post_shift)) * pW + (x >> a.post_shift));
    r[0] += t.x; r[1] += t.y; r[2] += t.z; r[3] += t.w;
  }
  a.y[i] = make_float4(r[0], r[1], r[2], r[3]);
}

// ------------------------------------------------------------------------------------------------- BN backward
// pre = scale*x + shift + res ; y = act(pre) (+ post).  g_pre = g_y * act'(pre).
// reduce: sums[0][c] = sum g_pre, sums[1][c] = sum g_pre * xhat,  xhat = (x - mean) * rstd
struct BnBwdArgs {
  const float4* x; const float4* res; const float4* gy;
  const float* scale; const float* shift; const float* mean; const float* rstd; const float* gamma_pad;   // [C4*4]
  double* sums;                // [2][C4*4]
  float4* gx; float4* gres;    // apply phase outputs (gres may be null)
  int N, C4, HW, act, chunks;
  double count;
  int C;                       // gamma_pad holds C floats when C > 0 (unpadded parameter), else C4*4
  float* g_gamma; float* g_beta;   // += (either may be null): the BatchNorm parameter gradients, written by CTA 0 of the apply pass
  float* g_bias;                   // += (may be null): sum of g_x per channel = the gradient of the conv bias in front of the BatchNorm
  int g0;                          // >= 0: this launch covers the planes of channel group g0 only (HW % 256 == 0); -1: every group
};

__device__ __forceinline__ void bn_gpre(const BnBwdArgs& a, size_t i, int g, float (&gp)[4], float (&xh)[4]) {
  const float4 v = __ldg(a.x + i), gy = __ldg(a.gy + i);
  const float4 sc = __ldg(reinterpret_cast<const float4*>(a.scale) + g), sh = __ldg(reinterpret_cast<const float4*>(a.shift) + g);
  const float4 mu = __ldg(reinterpret_cast<const float4*>(a.mean) + g), rs = __ldg(reinterpret_cast<const float4*>(a.rstd) + g);
  float pre[4] = {fmaf(sc.x, v.x, sh.x), fmaf(sc.y, v.y, sh.y), fmaf(sc.z, v.z, sh.z), fmaf(sc.w, v.w, sh.w)};
  if (a.res) { const float4 t = __ldg(a.res + i); pre[0] += t.x; pre[1] += t.y; pre[2] += t.z; pre[3] += t.w; }
  const float gyv[4] = {gy.x, gy.y, gy.z, gy.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) gp[e] = (a.act == 1 && pre[e] <= 0.f) ? 0.2f * gyv[e] : gyv[e];
  xh[0] = (v.x - mu.x) * rs.x; xh[1] = (v.y - mu.y) * rs.y; xh[2] = (v.z - mu.z) * rs.z; xh[3] = (v.w - mu.w) * rs.w;
}

__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const BnBwdArgs a) {
  __shared__ double s_red[8][8];
  const int g = a.g0 >= 0 ? a.g0 : blockIdx.y, n = blockIdx.z;
  const int per = (a.HW + a.chunks - 1) / a.chunks;
  const int lo = blockIdx.x * per, hi = min(lo + per, a.HW);
  const size_t base = ((size_t)n * a.C4 + g) * a.HW;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double ds[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int cnt = 0;
  // (measured: two or four elements per iteration and twice the CTAs make both reduction passes SLOWER — 78 -> 95 us backward,
  // 16.8 -> 17.1 us statistics on a 16-channel 256^2 layer at B = 16; the passes are not short of loads in flight)
  for (int i = lo + threadIdx.x; i < hi; i += 256) {
    float gp[4], xh[4];
    bn_gpre(a, base + i, g, gp, xh);
#pragma unroll
    for (int e = 0; e < 4; ++e) { s[e] += gp[e]; s[4 + e] += gp[e] * xh[e]; }
    if (++cnt == 32) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { ds[e] += s[e]; s[e] = 0.f; }
      cnt = 0;
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) ds[e] += s[e];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    double v = ds[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[e][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_red[threadIdx.x][w];
    const int e = threadIdx.x & 3, which = threadIdx.x >> 2;
    atomicAdd(a.sums + (size_t)which * a.C4 * 4 + g * 4 + e, t);
  }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BnBwdArgs a) {
  const long long total = (long long)a.N * a.C4 * a.HW;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // per-channel constants: mean of g_pre, mean of g_pre * xhat, gamma * rstd.  HW % 256 == 0 (every layer of the network): the
  // 256 elements of a CTA share their channel group, so 4 threads compute them once per CTA (the first version divided two
  // fp64 sums per channel in EVERY thread: 8 fp64 divisions per element, 75 us for a 16-channel 256^2 layer at B = 16)
  __shared__ float s_c[3][4];
  const bool uniform = (a.HW & 255) == 0;
  unsigned plane_u = uniform ? blockIdx.x / ((unsigned)a.HW >> 8) : 0u;          // the CTA's (n, group) plane: one 32-bit division
  if (a.g0 >= 0) {                  // one channel group per launch: the grid covers its N planes
    const unsigned bpp = (unsigned)a.HW >> 8, n = blockIdx.x / bpp;
    plane_u = n * (unsigned)a.C4 + (unsigned)a.g0;
    i = (long long)plane_u * a.HW + (long long)(blockIdx.x - n * bpp) * 256 + threadIdx.x;
  }
  if (uniform) {
    if (threadIdx.x < 4) {
      const int c = (int)(plane_u % (unsigned)a.C4) * 4 + threadIdx.x;
      s_c[0][threadIdx.x] = (float)(a.sums[c] / a.count);
      s_c[1][threadIdx.x] = (float)(a.sums[a.C4 * 4 + c] / a.count);
      s_c[2][threadIdx.x] = ((a.C > 0 && c >= a.C) ? 0.f : __ldg(a.gamma_pad + c)) * __ldg(a.rstd + c);
    }
    __syncthreads();
  }
  if (i >= total) return;
  const int g = uniform ? (int)(plane_u % (unsigned)a.C4) : (int)((i / a.HW) % a.C4);
  float gp[4], xh[4];
  bn_gpre(a, (size_t)i, g, gp, xh);
  if (a.gres) a.gres[i] = make_float4(gp[0], gp[1], gp[2], gp[3]);
  float out[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = g * 4 + e;
    float sg, sgx, gr;
    if (uniform) { sg = s_c[0][e]; sgx = s_c[1][e]; gr = s_c[2][e]; }
    else {
      sg = (float)(a.sums[c] / a.count); sgx = (float)(a.sums[a.C4 * 4 + c] / a.count);
      gr = ((a.C > 0 && c >= a.C) ? 0.f : __ldg(a.gamma_pad + c)) * __ldg(a.rstd + c);
    }
    out[e] = gr * (gp[e] - sg - xh[e] * sgx);
  }
  a.gx[i] = make_float4(out[0], out[1], out[2], out[3]);
  if (a.g_bias != nullptr) {
    // HW % 256 == 0 (host-checked): the 256 elements of a CTA share their channel group -> ONE atomic per CTA and channel
    // (a per-warp atomic serialises 32 768 same-address updates per 256^2 layer in L2: measured +5 ms per training iteration)
    __shared__ float s_b[8][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v = out[e];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) s_b[threadIdx.x >> 5][e] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4 && g * 4 + (int)threadIdx.x < a.C) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += s_b[w][threadIdx.x];
      atomicAdd(a.g_bias + g * 4 + threadIdx.x, t);
    }
  }
  if (blockIdx.x == 0 && (a.g_gamma != nullptr || a.g_beta != nullptr)) {
    const int c_lo = a.g0 >= 0 ? 4 * a.g0 : 0, c_hi = a.g0 >= 0 ? min(a.C, 4 * a.g0 + 4) : a.C;
    for (int c = c_lo + threadIdx.x; c < c_hi; c += 256) {
      if (a.g_beta) a.g_beta[c] += (float)a.sums[c];
      if (a.g_gamma) a.g_gamma[c] += (float)a.sums[a.C4 * 4 + c];
    }
  }
}

// ------------------------------------------------------------------------------------------------- wgrad 3x3
// dW_eff[co][ci][tap] += sum_{n,y,x} g[n,co,y,x] * in[n,ci,y+ky-1,x+kx-1].  A CTA owns a 16(co) x 16(ci) block and a
// strided subset of the 64x8-pixel tiles; thread (co, ci) keeps the 9 taps in registers; one atomicAdd round at the end.
constexpr int WG_TW = 32, WG_TH = 8;
constexpr int WG_IP = WG_TW + 4;                          // input row pitch (floats), keeps float4 alignment
constexpr int WG_IPLANE = (WG_TH + 2) * WG_IP + 4;        // 364: channel-plane pitch, 364 % 32 = 12 spreads the 16 channels over the banks
constexpr int WG_GPLANE = WG_TH * WG_TW;

struct WgradArgs {
  const float* in; const float* g;     // C4 [N][Cin4_alloc][Hin][Win][4], C4 [N][Cout4][H][W][4]
  float* dw;                           // element (co, ci, tap) at dw[co*so + ci*si + (flip ? KT*KT - 1 - tap : tap)]  +=
  int N, Cin, Cout, in_groups, H, W;   // H, W: size of g (the conv OUTPUT)
  int Hin, Win;                        // size of `in`
  long long so, si; int flip;
  int tiles_x, tiles_y, n_tiles;
};

// KT = 3: 3x3 / pad 1 (input pixel (y + ky - 1, x + kx - 1));  KT = 2: the 2x2-tap layers (input pixel (y + ky, x + kx))
template <int KT>
__global__ void __launch_bounds__(256) wgrad_kernel(const WgradArgs a) {
  constexpr int ORG = KT == 3 ? 1 : 0, NTAP = KT * KT;
  __shared__ __align__(16) float s_in[16 * WG_IPLANE];
  __shared__ __align__(16) float s_g[16 * WG_GPLANE];
  const int tid = threadIdx.x;
  const int co_l = tid >> 4, ci_l = tid & 15;
  const int cob = blockIdx.y * 16, cib = blockIdx.z * 16;
  const int C4out = (a.Cout + 3) >> 2;
  const size_t plane4 = (size_t)a.H * a.W * 4, iplane4 = (size_t)a.Hin * a.Win * 4;
  float acc[NTAP];
#pragma unroll
  for (int t = 0; t < NTAP; ++t) acc[t] = 0.f;
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int tx = tile % a.tiles_x, t2 = tile / a.tiles_x;
    const int ty = t2 % a.tiles_y, n = t2 / a.tiles_y;
    const int x0 = tx * WG_TW, y0 = ty * WG_TH;
    __syncthreads();
    for (int i = tid; i < 4 * (WG_TH + 2) * (WG_TW + 2); i += 256) {          // input halo tile: 4 groups = 16 channels
      const int c = i % (WG_TW + 2), r = (i / (WG_TW + 2)) % (WG_TH + 2), q = i / ((WG_TW + 2) * (WG_TH + 2));
      const int gy = y0 + r - ORG, gx = x0 + c - ORG, grp = (cib >> 2) + q;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < a.Hin && gx >= 0 && gx < a.Win && grp * 4 < a.Cin)
        v = __ldg(reinterpret_cast<const float4*>(a.in + ((size_t)n * a.in_groups + grp) * iplane4 + ((size_t)gy * a.Win + gx) * 4));
      float* d = s_in + (q * 4) * WG_IPLANE + r * WG_IP + c;
      d[0] = v.x; d[WG_IPLANE] = v.y; d[2 * WG_IPLANE] = v.z; d[3 * WG_IPLANE] = v.w;
    }
    for (int i = tid; i < 4 * WG_TH * WG_TW; i += 256) {
      const int c = i % WG_TW, r = (i / WG_TW) % WG_TH, q = i / (WG_TW * WG_TH);
      const int gy = y0 + r, gx = x0 + c, grp = (cob >> 2) + q;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy < a.H && gx < a.W && grp < C4out)
        v = __ldg(reinterpret_cast<const float4*>(a.g + ((size_t)n * C4out + grp) * plane4 + ((size_t)gy * a.W + gx) * 4));
      float* d = s_g + (q * 4) * WG_GPLANE + r * WG_TW + c;
      d[0] = v.x; d[WG_GPLANE] = v.y; d[2 * WG_GPLANE] = v.z; d[3 * WG_GPLANE] = v.w;
    }
    __syncthreads();
#pragma unroll 1
    for (int r = 0; r < WG_TH; ++r) {
#pragma unroll 4
      for (int c = 0; c < WG_TW; c += 4) {
        const float4 gv = *reinterpret_cast<const float4*>(s_g + co_l * WG_GPLANE + r * WG_TW + c);
#pragma unroll
        for (int ky = 0; ky < KT; ++ky) {
          const float* row = s_in + ci_l * WG_IPLANE + (r + ky) * WG_IP + c;
          const float4 i4 = *reinterpret_cast<const float4*>(row);
          const float i5 = row[4], i6 = row[5];
          acc[ky * KT + 0] += gv.x * i4.x + gv.y * i4.y + gv.z * i4.z + gv.w * i4.w;
          acc[ky * KT + 1] += gv.x * i4.y + gv.y * i4.z + gv.z * i4.w + gv.w * i5;
          if (KT == 3) acc[ky * KT + KT - 1] += gv.x * i4.z + gv.y * i4.w + gv.z * i5 + gv.w * i6;
        }
      }
    }
  }
  const int co = cob + co_l, ci = cib + ci_l;
  if (co < a.Cout && ci < a.Cin) {
#pragma unroll
    for (int t = 0; t < NTAP; ++t) atomicAdd(a.dw + co * a.so + ci * a.si + (a.flip ? NTAP - 1 - t : t), acc[t]);
  }
}

// per-channel sum of a C4 tensor (bias gradient): sums[C4*4] += sum over (N,H,W)
__global__ void __launch_bounds__(256) channel_sum_kernel(const float4* __restrict__ x, float* __restrict__ out, int C4, int HW, int chunks,
                                                          int C) {
  __shared__ float s_red[4][8];
  const int g = blockIdx.y, n = blockIdx.z;
  const int per = (HW + chunks - 1) / chunks;
  const int lo = blockIdx.x * per, hi = min(lo + per, HW);
  const float4* p = x + ((size_t)n * C4 + g) * HW;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = lo + threadIdx.x; i < hi; i += 256) { const float4 v = __ldg(p + i); s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w; }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float v = s[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[e][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_red[threadIdx.x][w];
    if (g * 4 + (int)threadIdx.x < C) atomicAdd(out + g * 4 + threadIdx.x, t);       // `out` may be exactly C floats long
  }
}

// gamma / beta gradients of a train-mode BatchNorm from the backward's fp64 sums: g_beta[c] += sums[0][c], g_gamma[c] += sums[1][c]
__global__ void bn_param_grads_kernel(const double* __restrict__ sums, float* __restrict__ g_gamma, float* __restrict__ g_beta, int C, int Cpad) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (g_beta) g_beta[c] += (float)sums[c];
  if (g_gamma) g_gamma[c] += (float)sums[Cpad + c];
}

// ------------------------------------------------------------------------------------------------- pools
__global__ void maxpool2_c4_bwd_kernel(const float4* __restrict__ x, const float4* __restrict__ gy, float4* __restrict__ gx,
                                       long long n_out, int Ho, int Wo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over pooled positions [NC4][Ho][Wo]
  if (i >= n_out) return;
  const int xo = (int)(i % Wo);
  const long long t = i / Wo;
  const int yo = (int)(t % Ho);
  const long long nc = t / Ho;
  const long long b = (nc * (2 * Ho) + 2 * yo) * (2LL * Wo) + 2 * xo;
  const long long off[4] = {b, b + 1, b + 2 * Wo, b + 2 * Wo + 1};
  float v[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { const float4 q = __ldg(x + off[k]); v[k][0] = q.x; v[k][1] = q.y; v[k][2] = q.z; v[k][3] = q.w; }
  const float4 g4 = __ldg(gy + i);
  const float g[4] = {g4.x, g4.y, g4.z, g4.w};
  float o[4][4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    int best = 0;                                   // first maximum in scan order, like torch's max_pool2d backward
#pragma unroll
    for (int k = 1; k < 4; ++k) if (v[k][e] > v[best][e]) best = k;
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k][e] = k == best ? g[e] : 0.f;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) gx[off[k]] = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
}

__global__ void sumpool2_c4_kernel(const float4* __restrict__ x, float4* __restrict__ out, long long n_out, int Ho, int Wo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const int xo = (int)(i % Wo);
  const long long t = i / Wo;
  const int yo = (int)(t % Ho);
  const long long nc = t / Ho;
  const float4* p = x + (nc * (2 * Ho) + 2 * yo) * (2LL * Wo) + 2 * xo;
  const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2 * Wo), d = __ldg(p + 2 * Wo + 1);
  out[i] = make_float4(a.x + b.x + c.x + d.x, a.y + b.y + c.y + d.y, a.z + b.z + c.z + d.z, a.w + b.w + c.w + d.w);
}

// channels [c_first, c_first + n_ch) of a C4 map: out[n][c] = mean over HW (forward) / gx[n, c, :] = g[n][c] / HW (backward, +=0 elsewhere untouched)
__global__ void __launch_bounds__(128) avgpool_c4_kernel(const float* __restrict__ feat, float* __restrict__ out, int C4, int c_first,
                                                          int n_ch, int HW) {
  const int n = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < n_ch; c += 4) {
    const int ch = c_first + c;
    const float* f = feat + (((size_t)n * C4 + (ch >> 2)) * HW) * 4 + (ch & 3);
    float s = 0.f;
    for (int i = lane; i < HW; i += 32) s += __ldg(f + (size_t)i * 4);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[n * n_ch + c] = s / (float)HW;
  }
}

__global__ void avgpool_c4_bwd_kernel(const float* __restrict__ g, float* __restrict__ gfeat, int C4, int c_first, int n_ch, int HW,
                                      long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over [N][n_ch][HW]
  if (i >= total) return;
  const int p = (int)(i % HW);
  const long long t = i / HW;
  const int c = (int)(t % n_ch);
  const long long n = t / n_ch;
  const int ch = c_first + c;
  gfeat[(((size_t)n * C4 + (ch >> 2)) * HW + p) * 4 + (ch & 3)] += __ldg(g + n * n_ch + c) / (float)HW;
}

// ------------------------------------------------------------------------------------------------- 1x1 convolutions (Cin = 16)
struct PwArgs {
  const float* in;    // C4 [N,16,H,W]
  const float* w;     // [Cout,16] device
  const float* bias;  // [Cout] device
  float* out;         // C4 [N,16,H,W] (planar = 0) or NCHW [N,Cout,H,W] (planar = 1)
  long long hw, total;
  int Cout, planar, act; float scale;
};

__global__ void __launch_bounds__(256) pw_conv16_fwd_kernel(const PwArgs a) {
  __shared__ __align__(16) float s_w[16][16];
  __shared__ float s_b[16];
  for (int i = threadIdx.x; i < 16 * 16; i += 256) s_w[i >> 4][i & 15] = (i >> 4) < a.Cout ? __ldg(a.w + i) : 0.f;
  if (threadIdx.x < 16) s_b[threadIdx.x] = threadIdx.x < a.Cout ? __ldg(a.bias + threadIdx.x) : 0.f;
  __syncthreads();
  // persistent over 256-pixel chunks; hw % 256 == 0 (every layer here): the image index is one 32-bit division per chunk
  const long long n_chunks = (a.total + 255) >> 8;
  const unsigned cpi = (a.hw & 255) == 0 ? (unsigned)(a.hw >> 8) : 0u;       // chunks per image
  for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const long long i = chunk * 256 + threadIdx.x;
    if (i >= a.total) continue;
    long long n, p;
    if (cpi) { const unsigned nn = (unsigned)chunk / cpi; n = nn; p = (long long)((unsigned)chunk - nn * cpi) * 256 + threadIdx.x; }
    else { n = i / a.hw; p = i - n * a.hw; }
    const float4* src = reinterpret_cast<const float4*>(a.in) + n * 4 * a.hw + p;
    float x[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float4 v = __ldg(src + q * a.hw); x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w; }
    float y[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) {
      float s = s_b[o];
#pragma unroll
      for (int q = 0; q < 4; ++q) {          // same order as before (c ascending): bit-identical, 64 LDS.128 instead of 256 LDS
        const float4 w4 = *reinterpret_cast<const float4*>(&s_w[o][4 * q]);
        s = fmaf(w4.x, x[4 * q], s); s = fmaf(w4.y, x[4 * q + 1], s); s = fmaf(w4.z, x[4 * q + 2], s); s = fmaf(w4.w, x[4 * q + 3], s);
      }
      if (a.act == 2) s = 1.0f / (1.0f + expf(-s));
      y[o] = s * a.scale;
    }
    if (a.planar) {
      for (int o = 0; o < a.Cout; ++o) a.out[(n * a.Cout + o) * a.hw + p] = y[o];
    } else {
      float4* dst = reinterpret_cast<float4*>(a.out) + n * 4 * a.hw + p;
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[q * a.hw] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
    }
  }
}

struct PwBwdArgs {
  const float* in;    // C4 [N,16,H,W]   forward input
  const float* w;     // [Cout,16]
  const float* gout;  // C4 [N,16,H,W] (planar 0) or NCHW [N,Cout,H,W] (planar 1): gradient w.r.t. the forward OUTPUT
  const float* out;   // forward output (needed for act = 2 sigmoid), same layout as gout; may be null for act 0
  float* gin;         // C4 [N,16,H,W]
  float* gw;          // [Cout,16] +=
  float* gb;          // [Cout]    +=
  long long hw, total;
  int Cout, planar, act; float scale;
};

// Round 2: every shared-memory access is a 16-byte one.  The data gradient reads the weight rows as float4 broadcasts (64 LDS.128
// per pixel instead of 256 LDS.32); for the weight gradient thread t owns the 4x4 block (o4, c4) = (t & 15) >> 2, t & 3 and the
// pixels p = 16 j + (t >> 4) of the CTA's 256: two LDS.128 feed 16 multiply-adds (the first version: two LDS.32 per multiply-add),
// the 16 pixel groups are then summed through shared memory, one global atomicAdd per (o, c) and CTA as before.
constexpr int PW_PITCH = 20;      // floats per pixel row in shared memory: 16-byte aligned, rows 16 apart land on different banks

__global__ void __launch_bounds__(256) pw_conv16_bwd_kernel(const PwBwdArgs a) {
  __shared__ __align__(16) float s_w[16][16];
  __shared__ __align__(16) float s_x[256 * PW_PITCH], s_g[256 * PW_PITCH];     // this CTA's 256 pixels: forward input and output gradient
  for (int i = threadIdx.x; i < 16 * 16; i += 256) s_w[i >> 4][i & 15] = (i >> 4) < a.Cout ? __ldg(a.w + i) : 0.f;
  __syncthreads();
  // Persistent over 256-pixel chunks: the weight / bias gradient partials stay in registers across a CTA's chunks and leave with ONE
  // round of shared-memory sums + atomics per CTA (one CTA per chunk meant 4 096 CTAs x 272 atomics on the same 272 addresses for a
  // 256^2 layer at B = 16, and two 64-bit divisions per thread).
  const long long n_chunks = (a.total + 255) >> 8;
  const unsigned cpi = (a.hw & 255) == 0 ? (unsigned)(a.hw >> 8) : 0u;       // chunks per image
  const int blk = threadIdx.x & 15, o4 = blk >> 2, c4 = blk & 3, pg = threadIdx.x >> 4;
  float acc[4][4];
  float4 accb = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
  const long long i = chunk * 256 + threadIdx.x;
  float x[16], g[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) { x[c] = 0.f; g[c] = 0.f; }
  if (i < a.total) {
    long long n, p;
    if (cpi) { const unsigned nn = (unsigned)chunk / cpi; n = nn; p = (long long)((unsigned)chunk - nn * cpi) * 256 + threadIdx.x; }
    else { n = i / a.hw; p = i - n * a.hw; }
    const float4* src = reinterpret_cast<const float4*>(a.in) + n * 4 * a.hw + p;
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float4 v = __ldg(src + q * a.hw); x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w; }
    if (a.planar) {
      for (int o = 0; o < a.Cout; ++o) {
        float go = __ldg(a.gout + (n * a.Cout + o) * a.hw + p) * a.scale;
        if (a.act == 2) { const float yv = __ldg(a.out + (n * a.Cout + o) * a.hw + p) / a.scale; go *= yv * (1.f - yv); }
        g[o] = go;
      }
    } else {
      const float4* gs = reinterpret_cast<const float4*>(a.gout) + n * 4 * a.hw + p;
#pragma unroll
      for (int q = 0; q < 4; ++q) { const float4 v = __ldg(gs + q * a.hw); g[4 * q] = v.x * a.scale; g[4 * q + 1] = v.y * a.scale; g[4 * q + 2] = v.z * a.scale; g[4 * q + 3] = v.w * a.scale; }
    }
    float gi[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) gi[c] = 0.f;
#pragma unroll
    for (int o = 0; o < 16; ++o) {
      const float go = g[o];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 w4 = *reinterpret_cast<const float4*>(&s_w[o][4 * q]);
        gi[4 * q] = fmaf(w4.x, go, gi[4 * q]); gi[4 * q + 1] = fmaf(w4.y, go, gi[4 * q + 1]);
        gi[4 * q + 2] = fmaf(w4.z, go, gi[4 * q + 2]); gi[4 * q + 3] = fmaf(w4.w, go, gi[4 * q + 3]);
      }
    }
    float4* dst = reinterpret_cast<float4*>(a.gin) + n * 4 * a.hw + p;
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q * a.hw] = make_float4(gi[4 * q], gi[4 * q + 1], gi[4 * q + 2], gi[4 * q + 3]);
  }
  // weight / bias gradient partials of this chunk
  __syncthreads();                                   // (the previous chunk's readers are done with s_x / s_g)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    *reinterpret_cast<float4*>(&s_x[threadIdx.x * PW_PITCH + 4 * q]) = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
    *reinterpret_cast<float4*>(&s_g[threadIdx.x * PW_PITCH + 4 * q]) = make_float4(g[4 * q], g[4 * q + 1], g[4 * q + 2], g[4 * q + 3]);
  }
  __syncthreads();
#pragma unroll 4
  for (int j = 0; j < 16; ++j) {
    const int px = 16 * j + pg;
    const float4 gv = *reinterpret_cast<const float4*>(&s_g[px * PW_PITCH + 4 * o4]);
    const float4 xv = *reinterpret_cast<const float4*>(&s_x[px * PW_PITCH + 4 * c4]);
    const float gg[4] = {gv.x, gv.y, gv.z, gv.w}, xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(gg[r], xx[c], acc[r][c]);
    accb.x += gv.x; accb.y += gv.y; accb.z += gv.z; accb.w += gv.w;
  }
  }
  __syncthreads();                                 // s_x / s_g are free: reuse them for the cross-group sums
  float* s_part = s_x;                             // [16 groups][16 blocks][16] = 4096 floats (s_x holds 5120)
  float* s_pb = s_g;                               // [16 groups][4 o4][4]
#pragma unroll
  for (int r = 0; r < 4; ++r)
    *reinterpret_cast<float4*>(&s_part[(pg * 16 + blk) * 16 + 4 * r]) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  if (c4 == 0) *reinterpret_cast<float4*>(&s_pb[(pg * 4 + o4) * 4]) = accb;
  __syncthreads();
  {
    const int o = threadIdx.x >> 4, c = threadIdx.x & 15;                          // one (o, c) per thread
    if (o < a.Cout) {
      const int b2 = (o >> 2) * 4 + (c >> 2), e = (o & 3) * 4 + (c & 3);
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) t += s_part[(k * 16 + b2) * 16 + e];
      if (t != 0.f) atomicAdd(a.gw + o * 16 + c, t);
      if (c == 0) {
        float tb = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) tb += s_pb[(k * 4 + (o >> 2)) * 4 + (o & 3)];
        if (tb != 0.f) atomicAdd(a.gb + o, tb);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------- stem, train mode
// conv_c1_og with DEVICE weights (they change every step): raw output (bias added, no BN / activation / pool).
constexpr int SD_TW = 32, SD_TH = 16, SD_IW = SD_TW + 4, SD_IH = SD_TH + 4;

__global__ void __launch_bounds__(128) stem_conv_dev_kernel(const float* __restrict__ img_all, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ out, int H, int W) {
  __shared__ __align__(16) float s_in[3][SD_IH][SD_IW];
  __shared__ __align__(16) float s_w[25 * 3][16];       // [tap*3 + ci][co]
  __shared__ float s_b[16];
  const int tid = threadIdx.x;
  const int tiles_x = gfr_ceil_div(W, SD_TW);
  const int x0 = (blockIdx.x % tiles_x) * SD_TW, y0 = (blockIdx.x / tiles_x) * SD_TH;
  const int n = blockIdx.y;
  const float* __restrict__ img = img_all + (size_t)n * H * W * 3;
  for (int i = tid; i < 16 * 75; i += 128) {             // w [16][3][25]
    const int co = i / 75, rem = i % 75, ci = rem / 25, tap = rem % 25;
    s_w[tap * 3 + ci][co] = __ldg(w + i);
  }
  if (tid < 16) s_b[tid] = __ldg(bias + tid);
  for (int i = tid; i < SD_IH * SD_IW * 3; i += 128) {
    const int r = i / (SD_IW * 3), rem = i % (SD_IW * 3);
    const int c = rem / 3, ch = rem % 3;
    const int gy = y0 + r - 2, gx = x0 + c - 2;
    float v = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(img + ((size_t)gy * W + gx) * 3 + ch);
    s_in[ch][r][c] = v;
  }
  __syncthreads();
  const int tx = tid % (SD_TW / 2), ty = tid / (SD_TW / 2);
  float acc[4][16];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[p][c] = s_b[c];
#pragma unroll 1
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
      for (int kx = 0; kx < 5; ++kx) {
        const float4* wp = reinterpret_cast<const float4*>(&s_w[(ky * 5 + kx) * 3 + ci][0]);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 w4 = wp[c4];
          const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = c4 * 4 + e;
            acc[0][c] = fmaf(wv[e], win[ky][kx], acc[0][c]);
            acc[1][c] = fmaf(wv[e], win[ky][kx + 1], acc[1][c]);
            acc[2][c] = fmaf(wv[e], win[ky + 1][kx], acc[2][c]);
            acc[3][c] = fmaf(wv[e], win[ky + 1][kx + 1], acc[3][c]);
          }
        }
      }
  }
  const int oy = y0 + 2 * ty, ox = x0 + 2 * tx;
  if (oy >= H || ox >= W) return;
  const size_t plane = (size_t)H * W;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float* o = out + (((size_t)n * 4 + q) * plane + (size_t)oy * W + ox) * 4;
    *reinterpret_cast<float4*>(o) = make_float4(acc[0][q * 4], acc[0][q * 4 + 1], acc[0][q * 4 + 2], acc[0][q * 4 + 3]);
    if (ox + 1 < W) *reinterpret_cast<float4*>(o + 4) = make_float4(acc[1][q * 4], acc[1][q * 4 + 1], acc[1][q * 4 + 2], acc[1][q * 4 + 3]);
    if (oy + 1 < H) {
      *reinterpret_cast<float4*>(o + (size_t)W * 4) = make_float4(acc[2][q * 4], acc[2][q * 4 + 1], acc[2][q * 4 + 2], acc[2][q * 4 + 3]);
      if (ox + 1 < W) *reinterpret_cast<float4*>(o + (size_t)W * 4 + 4) = make_float4(acc[3][q * 4], acc[3][q * 4 + 1], acc[3][q * 4 + 2], acc[3][q * 4 + 3]);
    }
  }
}

// dW[co][ci][ky][kx] += sum g[n,co,y,x] * img[n,y+ky-2,x+kx-2,ci]; bias gradient alongside.  Persistent CTAs over 32x16 tiles.
// Thread t < 240 owns (co, ci, ky) = (t / 15, (t % 15) / 5, t % 5) and its five kx taps: it walks the tile row by row with a
// five-wide register window over the input row, so a pixel costs two shared-memory loads for five multiply-adds (the first
// version read both operands from shared memory for every multiply-add: 700 us for B = 16).  Threads 240..255 sum g for the
// bias gradient of channel t - 240.
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ img_all, const float* __restrict__ g, float* __restrict__ dw,
                                                          float* __restrict__ db, int N, int H, int W) {
  __shared__ float s_in[3][SD_IH][SD_IW];
  __shared__ float s_g[16][SD_TH][SD_TW + 1];
  const int tid = threadIdx.x;
  const int tiles_x = gfr_ceil_div(W, SD_TW), tiles_y = gfr_ceil_div(H, SD_TH), n_tiles = N * tiles_x * tiles_y;
  const bool is_w = tid < 240;
  const int co = is_w ? tid / 15 : tid - 240, ci = is_w ? (tid % 15) / 5 : 0, ky = is_w ? tid % 5 : 0;
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  const size_t plane = (size_t)H * W;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tx = tile % tiles_x, t2 = tile / tiles_x, ty = t2 % tiles_y, n = t2 / tiles_y;
    const int x0 = tx * SD_TW, y0 = ty * SD_TH;
    const float* img = img_all + (size_t)n * H * W * 3;
    __syncthreads();
    for (int i = tid; i < SD_IH * SD_IW * 3; i += 256) {
      const int r = i / (SD_IW * 3), rem = i % (SD_IW * 3), c = rem / 3, ch = rem % 3;
      const int gy = y0 + r - 2, gx = x0 + c - 2;
      s_in[ch][r][c] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(img + ((size_t)gy * W + gx) * 3 + ch) : 0.f;
    }
    for (int i = tid; i < 4 * SD_TH * SD_TW; i += 256) {
      const int c = i % SD_TW, r = (i / SD_TW) % SD_TH, q = i / (SD_TW * SD_TH);
      const int gy = y0 + r, gx = x0 + c;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy < H && gx < W) v = __ldg(reinterpret_cast<const float4*>(g + (((size_t)n * 4 + q) * plane + (size_t)gy * W + gx) * 4));
      s_g[q * 4][r][c] = v.x; s_g[q * 4 + 1][r][c] = v.y; s_g[q * 4 + 2][r][c] = v.z; s_g[q * 4 + 3][r][c] = v.w;
    }
    __syncthreads();
    if (is_w) {
      for (int r = 0; r < SD_TH; ++r) {
        const float* in_row = &s_in[ci][r + ky][0];
        const float* g_row = &s_g[co][r][0];
        float w0 = in_row[0], w1 = in_row[1], w2 = in_row[2], w3 = in_row[3];
#pragma unroll
        for (int c = 0; c < SD_TW; ++c) {
          const float w4 = in_row[c + 4], gv = g_row[c];
          acc[0] = fmaf(gv, w0, acc[0]); acc[1] = fmaf(gv, w1, acc[1]); acc[2] = fmaf(gv, w2, acc[2]);
          acc[3] = fmaf(gv, w3, acc[3]); acc[4] = fmaf(gv, w4, acc[4]);
          w0 = w1; w1 = w2; w2 = w3; w3 = w4;
        }
      }
    } else {
      float a = 0.f;
      for (int r = 0; r < SD_TH; ++r)
#pragma unroll 8
        for (int c = 0; c < SD_TW; ++c) a += s_g[co][r][c];
      acc[0] += a;
    }
  }
  if (is_w) {
#pragma unroll
    for (int kx = 0; kx < 5; ++kx) atomicAdd(dw + ((co * 3 + ci) * 5 + ky) * 5 + kx, acc[kx]);
  } else {
    atomicAdd(db + co, acc[0]);
  }
}

int chunks_for(int N, int C4, int HW) {
  int chunks = (148 * 4) / (N * C4 > 0 ? N * C4 : 1);
  if (chunks < 1) chunks = 1;
  const int maxc = (HW + 255) / 256;
  return chunks > maxc ? maxc : chunks;
}

}  // namespace

extern "C" int gfr_conv_tc_pack_weights_dev(const float* w, int is_transposed_conv, int for_dgrad, int Cin, int Cout, int NT,
                                            float* packed, void* stream) {
  return gfr_conv_tc_pack_weights_dev_ex(w, is_transposed_conv, for_dgrad, Cin, Cout, NT, 9, 3, packed, stream);
}

extern "C" long long gfr_conv_tc_pack_size_ex(int Cin, int Cout, int NT, int taps, int precision) {
  if (Cin <= 0 || Cout <= 0 || (NT != 16 && NT != 32 && NT != 64 && NT != 128) || (taps != 9 && taps != 4)) return GFR_E_ARG;
  const long long steps = (long long)gfr_ceil_div(Cout, NT) * gfr_ceil_div(Cin, 16) * taps;
  return precision == 4 ? steps * 2 * NT * 4 : steps * 4 * 2 * NT * 4;      // in floats (8 bf16 = 4 floats)
}

namespace {
// Cin / Cout are those of the LAYER (forward direction).  The packed operand computes O outputs from I inputs.
int pack_geometry(int is_transposed_conv, int for_dgrad, int Cin, int Cout, int NT, int taps, int precision, int* O, int* I, long long* so,
                  long long* si, int* flip, long long* steps) {
  if (Cin <= 0 || Cout <= 0 || (NT != 16 && NT != 32 && NT != 64 && NT != 128) || (taps != 9 && taps != 4)) return GFR_E_ARG;
  if (precision != 1 && precision != 3 && precision != 4) return GFR_E_ARG;
  *O = for_dgrad ? Cin : Cout; *I = for_dgrad ? Cout : Cin;
  if (!is_transposed_conv) {        // Conv2d parameter [Cout][Cin][k][k]
    if (!for_dgrad) { *so = (long long)Cin * taps; *si = taps; *flip = 0; } else { *so = taps; *si = (long long)Cin * taps; *flip = 1; }
  } else {                          // ConvTranspose2d parameter [Cin][Cout][k][k]; forward = conv with w.transpose(0,1).flip(2,3)
    if (!for_dgrad) { *so = taps; *si = (long long)Cout * taps; *flip = 1; } else { *so = (long long)Cout * taps; *si = taps; *flip = 0; }
  }
  *steps = (long long)gfr_ceil_div(*O, NT) * gfr_ceil_div(*I, 16) * taps;
  return GFR_OK;
}
}  // namespace

extern "C" int gfr_conv_tc_pack_job_size(void) { return (int)sizeof(PackJob); }

// Fills one record of the job table of gfr_conv_tc_pack_weights_batch (host memory, gfr_conv_tc_pack_job_size() bytes) and returns
// the number of 256-thread blocks the job takes (< 0: error).  first_block = the sum of the earlier jobs' block counts.
extern "C" long long gfr_conv_tc_pack_job_fill(void* job_host, const float* w, int is_transposed_conv, int for_dgrad, int Cin, int Cout,
                                               int NT, int taps, int precision, float* packed, long long first_block) {
  if (job_host == nullptr || w == nullptr || packed == nullptr || first_block < 0) return GFR_E_NULL;
  PackJob j;
  long long steps;
  const int rc = pack_geometry(is_transposed_conv, for_dgrad, Cin, Cout, NT, taps, precision, &j.O, &j.I, &j.so, &j.si, &j.flip, &steps);
  if (rc != GFR_OK) return rc;
  j.w = w; j.packed = packed; j.first_block = first_block; j.NT = NT; j.taps = taps; j.bf16 = precision == 4 ? 1 : 0;
  j.total = precision == 4 ? steps * 2 * NT * 8 : steps * 4 * 2 * NT * 4;
  memcpy(job_host, &j, sizeof(j));
  return (j.total + 255) / 256;
}

// jobs_dev: n_jobs records (device copy of the table), n_blocks = the sum of the jobs' block counts
extern "C" int gfr_conv_tc_pack_weights_batch(const void* jobs_dev, int n_jobs, long long n_blocks, void* stream) {
  GFR_RETURN_IF_NULL(jobs_dev);
  if (n_jobs <= 0 || n_blocks <= 0 || n_blocks > 0x7fffffffLL) return GFR_E_ARG;
  pack_weights_batch_kernel<<<(unsigned)n_blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const PackJob*>(jobs_dev), n_jobs);
  return gfr_launch_status();
}

extern "C" int gfr_conv_tc_pack_weights_dev_ex(const float* w, int is_transposed_conv, int for_dgrad, int Cin, int Cout, int NT,
                                               int taps, int precision, float* packed, void* stream) {
  GFR_RETURN_IF_NULL(w); GFR_RETURN_IF_NULL(packed);
  int O, I, flip; long long so, si, steps;
  const int rc = pack_geometry(is_transposed_conv, for_dgrad, Cin, Cout, NT, taps, precision, &O, &I, &so, &si, &flip, &steps);
  if (rc != GFR_OK) return rc;
  if (precision == 4) {
    const long long total = steps * 2 * NT * 8;
    pack_weights_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, reinterpret_cast<__nv_bfloat16*>(packed), total,
                                                                                                O, I, NT, so, si, flip, taps);
  } else {
    const long long total = steps * 4 * 2 * NT * 4;
    pack_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, packed, total, O, I, NT, so, si, flip, taps);
  }
  return gfr_launch_status();
}

extern "C" int gfr_bn_train_stats_ex(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                                     long long* num_batches_tracked, double* sums_scratch, float* mean, float* rstd, float* scale,
                                     float* shift, int N, int C, int H, int W, float eps, float momentum, void* stream);

// 1: the caller hands every BatchNorm entry point an all-zero `sums_scratch` (one memset for a whole training step instead of one
// memset node in front of each of its ~150 BatchNorm passes); 0 (default): the entry points zero it themselves
static int g_bn_scratch_prezeroed = 0;

extern "C" int gfr_bn_config(int scratch_prezeroed) {
  const int old = g_bn_scratch_prezeroed;
  if (scratch_prezeroed == 0 || scratch_prezeroed == 1) g_bn_scratch_prezeroed = scratch_prezeroed;
  return old;
}

extern "C" int gfr_bn_train_stats(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                                  double* sums_scratch, float* mean, float* rstd, float* scale, float* shift, int N, int C, int H,
                                  int W, float eps, float momentum, void* stream) {
  return gfr_bn_train_stats_ex(x, gamma, beta, running_mean, running_var, nullptr, sums_scratch, mean, rstd, scale, shift, N, C, H, W, eps,
                               momentum, stream);
}

extern "C" int gfr_bn_train_stats_ex(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                                     long long* num_batches_tracked, double* sums_scratch, float* mean, float* rstd, float* scale,
                                     float* shift, int N, int C, int H, int W, float eps, float momentum, void* stream) {
  GFR_RETURN_IF_NULL(x); GFR_RETURN_IF_NULL(gamma); GFR_RETURN_IF_NULL(beta); GFR_RETURN_IF_NULL(sums_scratch);
  GFR_RETURN_IF_NULL(mean); GFR_RETURN_IF_NULL(rstd); GFR_RETURN_IF_NULL(scale); GFR_RETURN_IF_NULL(shift);
  if (N <= 0 || N > 65535 || C <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  const int C4 = (C + 3) / 4, HW = H * W;
  cudaStream_t s = (cudaStream_t)stream;
  if (!g_bn_scratch_prezeroed) {
    const cudaError_t e = cudaMemsetAsync(sums_scratch, 0, ((size_t)2 * C4 * 4 + 1) * sizeof(double), s);      // + the ticket counter
    if (e != cudaSuccess) return (int)e;
  }
  const int chunks = chunks_for(N, C4, HW);
  const BnFinalizeArgs fin{gamma, beta, running_mean, running_var, mean, rstd, scale, shift, C, C4 * 4, (double)N * HW, eps, momentum,
                           num_batches_tracked};
  bn_stats_kernel<<<dim3(chunks, C4, N), 256, 0, s>>>(reinterpret_cast<const float4*>(x), sums_scratch, N, C4, HW, chunks, fin);
  return gfr_launch_status();
}

extern "C" int gfr_bn_running_update(const double* sums, float* running_mean, float* running_var, long long* num_batches_tracked,
                                     int N, int C, int H, int W, float momentum, void* stream) {
  GFR_RETURN_IF_NULL(sums); GFR_RETURN_IF_NULL(running_mean); GFR_RETURN_IF_NULL(running_var);
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  const int Cpad = (C + 3) / 4 * 4;
  bn_running_replay_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, running_mean, running_var, num_batches_tracked, C, Cpad,
                                                                             (double)N * H * W, momentum);
  return gfr_launch_status();
}

extern "C" int gfr_bn_apply_fwd(const float* x, const float* scale, const float* shift, const float* res, const float* post,
                                float* y, int N, int C, int H, int W, int post_shift, int act, void* stream) {
  GFR_RETURN_IF_NULL(x); GFR_RETURN_IF_NULL(scale); GFR_RETURN_IF_NULL(shift); GFR_RETURN_IF_NULL(y);
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  if (post_shift < 0 || post_shift > 1 || act < 0 || act > 1) return GFR_E_ARG;
  const int C4 = (C + 3) / 4;
  BnApplyArgs a{reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(res), reinterpret_cast<const float4*>(post),
                reinterpret_cast<float4*>(y), scale, shift, C4, H, W, post_shift, act, (long long)N * C4 * H * W,
                ((H * W) % 256) == 0 ? (H * W) / 256 : 0};
  bn_apply_kernel<<<(unsigned)((a.total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  return gfr_launch_status();
}

extern "C" int gfr_bn_apply_bwd(const float* x, const float* res, const float* g_y, const float* scale, const float* shift,
                                const float* mean, const float* rstd, const float* gamma_pad, double* sums_scratch, float* g_x,
                                float* g_res, int N, int C, int H, int W, int act, void* stream) {
  GFR_RETURN_IF_NULL(x); GFR_RETURN_IF_NULL(g_y); GFR_RETURN_IF_NULL(scale); GFR_RETURN_IF_NULL(shift); GFR_RETURN_IF_NULL(mean);
  GFR_RETURN_IF_NULL(rstd); GFR_RETURN_IF_NULL(gamma_pad); GFR_RETURN_IF_NULL(sums_scratch); GFR_RETURN_IF_NULL(g_x);
  if (N <= 0 || N > 65535 || C <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  const int C4 = (C + 3) / 4, HW = H * W;
  cudaStream_t s = (cudaStream_t)stream;
  if (!g_bn_scratch_prezeroed) {
    const cudaError_t e = cudaMemsetAsync(sums_scratch, 0, (size_t)2 * C4 * 4 * sizeof(double), s);
    if (e != cudaSuccess) return (int)e;
  }
  BnBwdArgs a{reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(res), reinterpret_cast<const float4*>(g_y), scale,
              shift, mean, rstd, gamma_pad, sums_scratch, reinterpret_cast<float4*>(g_x), reinterpret_cast<float4*>(g_res), N, C4, HW, act,
              chunks_for(N, C4, HW), (double)N * HW, 0, nullptr, nullptr, nullptr, -1};
  bn_bwd_reduce_kernel<<<dim3(a.chunks, C4, N), 256, 0, s>>>(a);
  const long long total = (long long)N * C4 * HW;
  bn_bwd_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a);
  return gfr_launch_status();
}

extern "C" int gfr_bn_apply_bwd_ex(const float* x, const float* res, const float* g_y, const float* scale, const float* shift,
                                   const float* mean, const float* rstd, const float* gamma, double* sums_scratch, float* g_x,
                                   float* g_res, float* g_gamma, float* g_beta, float* g_bias, int N, int C, int H, int W, int act,
                                   void* stream) {
  GFR_RETURN_IF_NULL(x); GFR_RETURN_IF_NULL(g_y); GFR_RETURN_IF_NULL(scale); GFR_RETURN_IF_NULL(shift); GFR_RETURN_IF_NULL(mean);
  GFR_RETURN_IF_NULL(rstd); GFR_RETURN_IF_NULL(gamma); GFR_RETURN_IF_NULL(sums_scratch); GFR_RETURN_IF_NULL(g_x);
  if (N <= 0 || N > 65535 || C <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  const int C4 = (C + 3) / 4, HW = H * W;
  cudaStream_t s = (cudaStream_t)stream;
  if (!g_bn_scratch_prezeroed) {
    const cudaError_t e = cudaMemsetAsync(sums_scratch, 0, (size_t)2 * C4 * 4 * sizeof(double), s);
    if (e != cudaSuccess) return (int)e;
  }
  BnBwdArgs a{reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(res), reinterpret_cast<const float4*>(g_y), scale,
              shift, mean, rstd, gamma, sums_scratch, reinterpret_cast<float4*>(g_x), reinterpret_cast<float4*>(g_res), N, C4, HW, act,
              chunks_for(N, C4, HW), (double)N * HW, C, g_gamma, g_beta, (HW % 256) == 0 ? g_bias : nullptr, -1};
  const long long total = (long long)N * C4 * HW;
  // A/B (GFR_BN_BWD_SPLIT=1, off): one channel group at a time, so that a group's x / g_y planes (a quarter of a 16-channel tensor)
  // could still be in L2 when the apply pass reads them again.  Measured SLOWER on the 67 MB tensors of the 256^2 layers at B = 16:
  // 101 vs 78 us per layer, 9.52 vs 9.08 ms per iteration — eight short launches over strided planes lose more than the re-read costs.
  static const bool no_split = [] { const char* e = getenv("GFR_BN_BWD_SPLIT"); return !(e != nullptr && e[0] == '1'); }();
  if (!no_split && (HW % 256) == 0 && C4 > 1 && total * 16 > (24ll << 20) && N * (HW / 256) >= 148) {
    a.chunks = chunks_for(N, 1, HW);
    for (int g = 0; g < C4; ++g) {
      a.g0 = g;
      bn_bwd_reduce_kernel<<<dim3(a.chunks, 1, N), 256, 0, s>>>(a);
      bn_bwd_apply_kernel<<<(unsigned)(N * (HW / 256)), 256, 0, s>>>(a);
    }
    return gfr_launch_status();
  }
  bn_bwd_reduce_kernel<<<dim3(a.chunks, C4, N), 256, 0, s>>>(a);
  bn_bwd_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a);
  if (g_bias != nullptr && (HW % 256) != 0) {         // ragged planes: the separate per-channel sum
    const int chunks = chunks_for(N, C4, HW);
    channel_sum_kernel<<<dim3(chunks, C4, N), 256, 0, s>>>(reinterpret_cast<const float4*>(g_x), g_bias, C4, HW, chunks, C);
  }
  return gfr_launch_status();
}

static int wgrad_launch(const float* in, const float* g_out, float* g_w, float* g_bias, int is_transposed_conv, int N, int Cin,
                        int in_groups, int Cout, int Hin, int Win, int H, int W, int taps, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(g_out); GFR_RETURN_IF_NULL(g_w);
  if (N <= 0 || N > 65535 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  if (in_groups == 0) in_groups = (Cin + 3) / 4;
  if (in_groups < (Cin + 3) / 4) return GFR_E_ARG;
  WgradArgs a;
  a.in = in; a.g = g_out; a.dw = g_w; a.N = N; a.Cin = Cin; a.Cout = Cout; a.in_groups = in_groups; a.H = H; a.W = W;
  a.Hin = Hin; a.Win = Win;
  if (!is_transposed_conv) { a.so = (long long)Cin * taps; a.si = taps; a.flip = 0; }     // dW_param[co][ci][tap]
  else { a.so = taps; a.si = (long long)Cout * taps; a.flip = 1; }                         // dW_param[ci][co][taps - 1 - tap]
  a.tiles_x = gfr_ceil_div(W, WG_TW); a.tiles_y = gfr_ceil_div(H, WG_TH); a.n_tiles = N * a.tiles_x * a.tiles_y;
  const int gy = gfr_ceil_div(Cout, 16), gz = gfr_ceil_div(Cin, 16);
  int gx = (148 * 2) / (gy * gz);
  if (gx < 1) gx = 1;
  if (gx > a.n_tiles) gx = a.n_tiles;
  cudaStream_t s = (cudaStream_t)stream;
  if (taps == 9) wgrad_kernel<3><<<dim3(gx, gy, gz), 256, 0, s>>>(a);
  else wgrad_kernel<2><<<dim3(gx, gy, gz), 256, 0, s>>>(a);
  if (g_bias) {
    const int C4 = (Cout + 3) / 4, chunks = chunks_for(N, C4, H * W);
    channel_sum_kernel<<<dim3(chunks, C4, N), 256, 0, s>>>(reinterpret_cast<const float4*>(g_out), g_bias, C4, H * W, chunks, Cout);
  }
  return gfr_launch_status();
}

extern "C" int gfr_channel_sum_c4(const float* x, float* out, int N, int C, int H, int W, void* stream) {
  GFR_RETURN_IF_NULL(x); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || N > 65535 || C <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  const int C4 = (C + 3) / 4, chunks = chunks_for(N, C4, H * W);
  channel_sum_kernel<<<dim3(chunks, C4, N), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(x), out, C4, H * W, chunks, C);
  return gfr_launch_status();
}

extern "C" int gfr_conv3x3_wgrad(const float* in, const float* g_out, float* g_w, float* g_bias, int is_transposed_conv, int N,
                                 int Cin, int in_groups, int Cout, int H, int W, void* stream) {
  return wgrad_launch(in, g_out, g_w, g_bias, is_transposed_conv, N, Cin, in_groups, Cout, H, W, H, W, 9, stream);
}

extern "C" int gfr_conv2x2_wgrad(const float* in, const float* g_out, float* g_w, float* g_bias, int N, int Cin, int in_groups,
                                 int Cout, int H, int W, void* stream) {
  return wgrad_launch(in, g_out, g_w, g_bias, 0, N, Cin, in_groups, Cout, H + 1, W + 1, H, W, 4, stream);
}

extern "C" int gfr_maxpool2_c4_bwd(const float* x, const float* g_y, float* g_x, int NC4, int Ho, int Wo, void* stream) {
  GFR_RETURN_IF_NULL(x); GFR_RETURN_IF_NULL(g_y); GFR_RETURN_IF_NULL(g_x);
  if (NC4 <= 0 || Ho <= 0 || Wo <= 0) return GFR_E_SHAPE;
  const long long n = (long long)NC4 * Ho * Wo;
  maxpool2_c4_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(g_y), reinterpret_cast<float4*>(g_x), n, Ho, Wo);
  return gfr_launch_status();
}

extern "C" int gfr_sumpool2_c4(const float* x, float* out, int NC4, int Ho, int Wo, void* stream) {
  GFR_RETURN_IF_NULL(x); GFR_RETURN_IF_NULL(out);
  if (NC4 <= 0 || Ho <= 0 || Wo <= 0) return GFR_E_SHAPE;
  const long long n = (long long)NC4 * Ho * Wo;
  sumpool2_c4_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(out), n, Ho, Wo);
  return gfr_launch_status();
}

extern "C" int gfr_avgpool_c4_fwd(const float* feat, float* out, int N, int C, int c_first, int n_ch, int HW, void* stream) {
  GFR_RETURN_IF_NULL(feat); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || C <= 0 || HW <= 0 || c_first < 0 || n_ch <= 0 || c_first + n_ch > C) return GFR_E_SHAPE;
  avgpool_c4_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(feat, out, (C + 3) / 4, c_first, n_ch, HW);
  return gfr_launch_status();
}

extern "C" int gfr_avgpool_c4_bwd(const float* g, float* g_feat, int N, int C, int c_first, int n_ch, int HW, void* stream) {
  GFR_RETURN_IF_NULL(g); GFR_RETURN_IF_NULL(g_feat);
  if (N <= 0 || C <= 0 || HW <= 0 || c_first < 0 || n_ch <= 0 || c_first + n_ch > C) return GFR_E_SHAPE;
  const long long total = (long long)N * n_ch * HW;
  avgpool_c4_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, g_feat, (C + 3) / 4, c_first, n_ch, HW, total);
  return gfr_launch_status();
}

extern "C" int gfr_pw_conv16_fwd(const float* in, const float* w, const float* bias, float* out, int N, int Cout, int H, int W,
                                 int planar_out, int act, float out_scale, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(w); GFR_RETURN_IF_NULL(bias); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || H <= 0 || W <= 0 || Cout < 1 || Cout > 16) return GFR_E_SHAPE;
  if ((act != 0 && act != 2) || (!planar_out && Cout != 16)) return GFR_E_ARG;
  PwArgs a{in, w, bias, out, (long long)H * W, (long long)N * H * W, Cout, planar_out, act, out_scale};
  const long long chunks = (a.total + 255) / 256;
  pw_conv16_fwd_kernel<<<(unsigned)(chunks < 148 * 8 ? chunks : 148 * 8), 256, 0, (cudaStream_t)stream>>>(a);
  return gfr_launch_status();
}

extern "C" int gfr_pw_conv16_bwd(const float* in, const float* w, const float* g_out, const float* out, float* g_in, float* g_w,
                                 float* g_bias, int N, int Cout, int H, int W, int planar_out, int act, float out_scale, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(w); GFR_RETURN_IF_NULL(g_out); GFR_RETURN_IF_NULL(g_in); GFR_RETURN_IF_NULL(g_w);
  GFR_RETURN_IF_NULL(g_bias);
  if (N <= 0 || H <= 0 || W <= 0 || Cout < 1 || Cout > 16) return GFR_E_SHAPE;
  if ((act != 0 && act != 2) || (!planar_out && Cout != 16) || (act == 2 && (out == nullptr || !planar_out))) return GFR_E_ARG;
  PwBwdArgs a{in, w, g_out, out, g_in, g_w, g_bias, (long long)H * W, (long long)N * H * W, Cout, planar_out, act, out_scale};
  const long long chunks = (a.total + 255) / 256;
  pw_conv16_bwd_kernel<<<(unsigned)(chunks < 148 * 5 ? chunks : 148 * 5), 256, 0, (cudaStream_t)stream>>>(a);      // 5 CTAs of 41 KB per SM
  return gfr_launch_status();
}

extern "C" int gfr_stem_conv_train_fwd(const float* img, const float* w, const float* bias, float* out_raw, int N, int H, int W,
                                       void* stream) {
  GFR_RETURN_IF_NULL(img); GFR_RETURN_IF_NULL(w); GFR_RETURN_IF_NULL(bias); GFR_RETURN_IF_NULL(out_raw);
  if (N <= 0 || N > 65535 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  const dim3 grid(gfr_ceil_div(W, SD_TW) * gfr_ceil_div(H, SD_TH), N);
  stem_conv_dev_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(img, w, bias, out_raw, H, W);
  return gfr_launch_status();
}

extern "C" int gfr_stem_conv_wgrad(const float* img, const float* g_out, float* g_w, float* g_bias, int N, int H, int W, void* stream) {
  GFR_RETURN_IF_NULL(img); GFR_RETURN_IF_NULL(g_out); GFR_RETURN_IF_NULL(g_w); GFR_RETURN_IF_NULL(g_bias);
  if (N <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  const int n_tiles = N * gfr_ceil_div(W, SD_TW) * gfr_ceil_div(H, SD_TH);
  const int grid = n_tiles < 148 * 2 ? n_tiles : 148 * 2;
  stem_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, g_out, g_w, g_bias, N, H, W);
  return gfr_launch_status();
}
