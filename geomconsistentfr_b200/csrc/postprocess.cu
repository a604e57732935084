// Output stage of the inference drivers, sm_100a: what the reference does to the forward's outputs between the model
// call and cv2.imwrite (TEST1:590-620, TESTB:589-608), and the MATLAB border post-fix the shipped result images went
// through (fix_border_artifacts_CVPR2022.m:1-18).  The reference does all of this on the host in numpy / MATLAB after
// a device->host copy of every fp32 plane (52 B per pixel); here the forward's outputs are quantised on the device
// and only the 8-bit images cross PCIe (3-11 B per pixel).
//
// Arithmetic follows numpy's promotion in the reference expressions exactly: `255.0 * f32_array` is an fp32 product,
// the product with the f64 mask (`mask / 255.0`) is fp64, and cv2.imwrite's f64 -> u8 conversion is
// saturate_cast<uchar>(cvRound(v)) = round-half-to-even, clamped to [0, 255].  All kernels are HBM-bound streaming
// passes, one thread per pixel, grid-stride free (H*W*B threads).
#include "gfr_common.cuh"

namespace {

__device__ __forceinline__ uint8_t to_u8(double v) {
  const int r = __double2int_rn(v);                      // cvRound (lrint, half to even); NaN -> 0
  return (uint8_t)min(max(r, 0), 255);
}

__device__ __forceinline__ double mask01(uint8_t m) { return __ddiv_rn((double)m, 255.0); }

// ---- TEST1:613-620 / TESTB:596-601: rendered * mask/255 where mask > 0, the input image elsewhere, RGB -> BGR, u8
template <typename ImgT>
__global__ void __launch_bounds__(256) composite_bgr_u8_kernel(const ImgT* __restrict__ image, const float* __restrict__ rendered,
                                                               const uint8_t* __restrict__ mask, long long mask_stride,
                                                               uint8_t* __restrict__ out, int B, int HW) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * HW) return;
  const int b = (int)(i / HW), p = (int)(i - (long long)b * HW);
  const uint8_t m = __ldg(mask + (long long)b * mask_stride + p);
  uint8_t* o = out + i * 3;
  if (m > 0) {
    const double mf = mask01(m);
    const float* r = rendered + (long long)b * 3 * HW + p;
#pragma unroll
    for (int c = 0; c < 3; ++c)                           // out channel c (B,G,R) <- rendered channel 2-c
      o[c] = to_u8(__dmul_rn((double)__fmul_rn(255.0f, __ldg(r + (long long)(2 - c) * HW)), mf));
  } else {
    const ImgT* x = image + i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = to_u8(__dmul_rn((double)x[2 - c], 255.0));
  }
}

// ---- range of -depth over the whole batch (TESTB:595-597: np.amin / np.amax of the negated array)
__device__ __forceinline__ uint32_t f32_key(float f) {   // order-preserving map float -> uint32
  const uint32_t u = __float_as_uint(f);
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float key_f32(uint32_t k) {
  return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}

__global__ void range_init_kernel(uint32_t* keys) { keys[0] = 0xFFFFFFFFu; keys[1] = 0u; }

__global__ void __launch_bounds__(256) neg_range_kernel(const float* __restrict__ x, long long n, uint32_t* keys) {
  uint32_t lo = 0xFFFFFFFFu, hi = 0u;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t k = f32_key(-__ldg(x + i));
    lo = min(lo, k); hi = max(hi, k);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, s));
    hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, s));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(keys, lo); atomicMax(keys + 1, hi); }
}

// ---- TESTB:590-608: the five auxiliary images
struct PlaneArgs {
  const float *albedo, *depth, *shadow, *final_shading, *normals;
  const uint8_t* mask; long long mask_stride; const uint32_t* range_keys;
  uint8_t *o_shadow, *o_albedo, *o_depth, *o_shading, *o_normals;
  int B, HW;
};

__global__ void __launch_bounds__(256) export_planes_u8_kernel(const PlaneArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)a.B * a.HW) return;
  const int b = (int)(i / a.HW), p = (int)(i - (long long)b * a.HW);
  const double mf = mask01(__ldg(a.mask + (long long)b * a.mask_stride + p));
  if (a.o_shadow) a.o_shadow[i] = to_u8(__dmul_rn((double)__fmul_rn(255.0f, __ldg(a.shadow + i)), mf));
  if (a.o_shading) a.o_shading[i] = to_u8(__dmul_rn((double)__fmul_rn(255.0f, __ldg(a.final_shading + i)), mf));
  if (a.o_depth) {
    const float lo = key_f32(a.range_keys[0]), hi = key_f32(a.range_keys[1]);
    const float d = __fdiv_rn(__fsub_rn(-__ldg(a.depth + i), lo), __fsub_rn(hi, lo));
    a.o_depth[i] = to_u8(__dmul_rn((double)__fmul_rn(255.0f, d), mf));
  }
  if (a.o_albedo) {
    const float* s = a.albedo + (long long)b * 3 * a.HW + p;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      a.o_albedo[i * 3 + c] = to_u8(__dmul_rn((double)__fmul_rn(255.0f, __ldg(s + (long long)(2 - c) * a.HW)), mf));
  }
  if (a.o_normals) {
    const float* s = a.normals + (long long)b * 3 * a.HW + p;
#pragma unroll
    for (int c = 0; c < 3; ++c) {                          // 255.0 * (n + 1.0) / 2.0 in fp32, then * mask (f64)
      const float v = __fmul_rn(__fmul_rn(255.0f, __fadd_rn(__ldg(s + (long long)(2 - c) * a.HW), 1.0f)), 0.5f);
      a.o_normals[i * 3 + c] = to_u8(__dmul_rn((double)v, mf));
    }
  }
}

// ---- fix_border_artifacts_CVPR2022.m: pixels whose 7x7 box sum of the binarised mask is in (0, max_sum] take the
// 3x3 median (zero padded, per channel) of the unfixed image.
__device__ __forceinline__ void cswap(int& a, int& b) { const int lo = min(a, b); b = max(a, b); a = lo; }

__device__ __forceinline__ int median9(int* v) {          // 19-exchange median network
  cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]); cswap(v[0], v[1]); cswap(v[3], v[4]); cswap(v[6], v[7]);
  cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]); cswap(v[0], v[3]); cswap(v[5], v[8]); cswap(v[4], v[7]);
  cswap(v[3], v[6]); cswap(v[1], v[4]); cswap(v[2], v[5]); cswap(v[4], v[7]); cswap(v[4], v[2]); cswap(v[6], v[4]);
  cswap(v[4], v[2]);
  return v[4];
}

constexpr int BT_W = 32, BT_H = 8, BT_R = 3;

__global__ void __launch_bounds__(BT_W * BT_H) border_median_fix_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ mask,
                                                                        long long mask_stride, uint8_t* __restrict__ out,
                                                                        int H, int W, int C, int max_sum) {
  __shared__ uint8_t face[BT_H + 2 * BT_R][BT_W + 2 * BT_R + 2];
  const int b = blockIdx.z, c0 = blockIdx.x * BT_W, r0 = blockIdx.y * BT_H;
  const uint8_t* mk = mask + (long long)b * mask_stride;
  for (int t = threadIdx.y * BT_W + threadIdx.x; t < (BT_H + 2 * BT_R) * (BT_W + 2 * BT_R); t += BT_W * BT_H) {
    const int ty = t / (BT_W + 2 * BT_R), tx = t - ty * (BT_W + 2 * BT_R);
    const int r = r0 + ty - BT_R, c = c0 + tx - BT_R;
    // MATLAB: imread(mask) / 255.0 is uint8 arithmetic -> round(m / 255): 1 iff m >= 128
    face[ty][tx] = (r >= 0 && r < H && c >= 0 && c < W && __ldg(mk + (long long)r * W + c) >= 128) ? 1 : 0;
  }
  __syncthreads();
  const int col = c0 + threadIdx.x, row = r0 + threadIdx.y;
  if (col >= W || row >= H) return;
  int s = 0;
#pragma unroll
  for (int dy = 0; dy < 2 * BT_R + 1; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2 * BT_R + 1; ++dx) s += face[threadIdx.y + dy][threadIdx.x + dx];
  const uint8_t* im = img + (long long)b * H * W * C;
  uint8_t* o = out + (long long)b * H * W * C + ((long long)row * W + col) * C;
  const bool border = s > 0 && s <= max_sum;
  for (int ch = 0; ch < C; ++ch) {
    if (!border) { o[ch] = im[((long long)row * W + col) * C + ch]; continue; }
    int v[9];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int r = row + dy, c = col + dx;
        v[(dy + 1) * 3 + dx + 1] = (r >= 0 && r < H && c >= 0 && c < W) ? (int)im[((long long)r * W + c) * C + ch] : 0;
      }
    o[ch] = (uint8_t)median9(v);
  }
}

// ---- MSE_MP.m:15-25: per-image masked MSE between two 8-bit images, sums in fp64
__global__ void __launch_bounds__(256) masked_mse_kernel(const uint8_t* __restrict__ recon, const uint8_t* __restrict__ gt,
                                                         const uint8_t* __restrict__ mask, long long mask_stride,
                                                         double* __restrict__ sums, int HW, int C) {
  const int b = blockIdx.y;
  double num = 0.0, den = 0.0;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    const double m = __ddiv_rn((double)__ldg(mask + (long long)b * mask_stride + p), 255.0);
    den += m;
    const uint8_t* r = recon + ((long long)b * HW + p) * C;
    const uint8_t* g = gt + ((long long)b * HW + p) * C;
    for (int c = 0; c < C; ++c) {
      const double d = __dsub_rn(__dmul_rn(__ddiv_rn((double)r[c], 255.0), m), __dmul_rn(__ddiv_rn((double)g[c], 255.0), m));
      num += d * d;
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    num += __shfl_xor_sync(0xFFFFFFFFu, num, s);
    den += __shfl_xor_sync(0xFFFFFFFFu, den, s);
  }
  __shared__ double sn[8], sd[8];
  if ((threadIdx.x & 31) == 0) { sn[threadIdx.x >> 5] = num; sd[threadIdx.x >> 5] = den; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int w = 0; w < 8; ++w) { a += sn[w]; c += sd[w]; }
    atomicAdd(sums + 2 * b, a);
    atomicAdd(sums + 2 * b + 1, c);
  }
}

}  // namespace

extern "C" int gfr_masked_mse_u8(const uint8_t* recon, const uint8_t* gt, const uint8_t* mask, int mask_batch_stride,
                                 double* sums, int B, int H, int W, int C, void* stream) {
  GFR_RETURN_IF_NULL(recon); GFR_RETURN_IF_NULL(gt); GFR_RETURN_IF_NULL(mask); GFR_RETURN_IF_NULL(sums);
  if (B <= 0 || B > 65535 || H <= 0 || W <= 0 || C < 1 || C > 4) return GFR_E_SHAPE;
  if (mask_batch_stride != 0 && mask_batch_stride != H * W) return GFR_E_ARG;
  const cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)B * 2 * sizeof(double), (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  const dim3 grid((unsigned)min(64, gfr_ceil_div(H * W, 256)), B);
  masked_mse_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(recon, gt, mask, mask_batch_stride, sums, H * W, C);
  return gfr_launch_status();
}

extern "C" int gfr_composite_bgr_u8(const void* image, int image_is_f64, const float* rendered, const uint8_t* mask,
                                    int mask_batch_stride, uint8_t* out_bgr, int B, int H, int W, void* stream) {
  GFR_RETURN_IF_NULL(image); GFR_RETURN_IF_NULL(rendered); GFR_RETURN_IF_NULL(mask); GFR_RETURN_IF_NULL(out_bgr);
  if (B <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  if (mask_batch_stride != 0 && mask_batch_stride != H * W) return GFR_E_ARG;
  const long long n = (long long)B * H * W;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (image_is_f64)
    composite_bgr_u8_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)image, rendered, mask, mask_batch_stride, out_bgr, B, H * W);
  else
    composite_bgr_u8_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)image, rendered, mask, mask_batch_stride, out_bgr, B, H * W);
  return gfr_launch_status();
}

extern "C" int gfr_neg_depth_range(const float* depth, long long n, uint32_t* range_keys, void* stream) {
  GFR_RETURN_IF_NULL(depth); GFR_RETURN_IF_NULL(range_keys);
  if (n <= 0) return GFR_E_SHAPE;
  range_init_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(range_keys);
  const unsigned grid = (unsigned)min((long long)148 * 8, (n + 255) / 256);
  neg_range_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(depth, n, range_keys);
  return gfr_launch_status();
}

extern "C" int gfr_export_planes_u8(const float* albedo, const float* depth, const float* shadow, const float* final_shading,
                                    const float* normals, const uint8_t* mask, int mask_batch_stride,
                                    const uint32_t* range_keys, uint8_t* out_shadow, uint8_t* out_albedo, uint8_t* out_depth,
                                    uint8_t* out_shading, uint8_t* out_normals, int B, int H, int W, void* stream) {
  GFR_RETURN_IF_NULL(mask);
  if ((out_shadow && !shadow) || (out_albedo && !albedo) || (out_depth && (!depth || !range_keys)) ||
      (out_shading && !final_shading) || (out_normals && !normals))
    return GFR_E_NULL;
  if (B <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  if (mask_batch_stride != 0 && mask_batch_stride != H * W) return GFR_E_ARG;
  PlaneArgs a{albedo, depth, shadow, final_shading, normals, mask, mask_batch_stride, range_keys,
              out_shadow, out_albedo, out_depth, out_shading, out_normals, B, H * W};
  const long long n = (long long)B * H * W;
  export_planes_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  return gfr_launch_status();
}

extern "C" int gfr_border_median_fix_u8(const uint8_t* img, const uint8_t* mask, int mask_batch_stride, uint8_t* out,
                                        int B, int H, int W, int C, int max_sum, void* stream) {
  GFR_RETURN_IF_NULL(img); GFR_RETURN_IF_NULL(mask); GFR_RETURN_IF_NULL(out);
  if (B <= 0 || H <= 0 || W <= 0 || B > 65535 || C < 1 || C > 4) return GFR_E_SHAPE;
  if (img == out) return GFR_E_ARG;                        // the median reads unfixed neighbours: not in place
  if (mask_batch_stride != 0 && mask_batch_stride != H * W) return GFR_E_ARG;
  const dim3 grid(gfr_ceil_div(W, BT_W), gfr_ceil_div(H, BT_H), B), block(BT_W, BT_H);
  border_median_fix_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(img, mask, mask_batch_stride, out, H, W, C, max_sum);
  return gfr_launch_status();
}
