// 3x3 convolution on the 5th-generation tensor cores (tcgen05, kind::tf32) with fp32-grade accuracy (3xTF32
// operand splitting), TMA-staged activations, TMEM accumulators and the fused RelightNet epilogues.
// Replaces the cuDNN Conv2d / ConvTranspose2d(stride 1) + BatchNorm2d(eval) + LeakyReLU + residual / skip adds +
// nearest x2 upsample of TRAIN:197-350 / TEST1:170-323 (TRAIN = train_raytracing_relighting_CelebAHQ_DSSIM_8x.py,
// TEST1 = test_relight_single_image.py) for every 3x3 layer whose input has >= 16 channels.
//
// Activation layout ("C4"): [N][C/4][H][W][4] fp32 — 4-channel groups innermost, so one pixel x one channel group is
// a 16-byte unit.  That is exactly the 16-byte row of a K-major, un-swizzled UMMA core matrix (8 rows x 16 B), so a
// TMA box {4ch, 10 px, 18 rows, 4 groups} lands in shared memory as [group][row][px][4] and the A operand of every
// one of the nine filter taps is the SAME tile addressed through a descriptor whose start address is shifted by
// (ky*10 + kx)*16 bytes: M = 128 output pixels = 16 rows x 8 px (row pitch 160 B = SBO), K = 8 channels = two core
// matrices 2880 B apart (LBO).  Image borders and channel padding come from TMA out-of-bounds zero fill.
//
// 3xTF32: every fp32 operand is split x = hi + lo with hi = tf32(x); D += lo*Whi + hi*Wlo + hi*Whi (the dropped
// lo*lo term is ~2^-22 relative).  Activations are split in shared memory by the CTA after the TMA lands, weights are
// split and packed once on the host (gfr_conv_tc_pack_weights).
//
// One CTA (128 threads) owns a 128-pixel x NT-channel output tile at a time and loops over its tiles; per 16 input
// channels: TMA -> split -> 54 MMAs (9 taps x 2 K-steps x 3 products, issued by one thread) -> commit.  Epilogue:
// tcgen05.ld (one thread per pixel, NT accumulators), bias + residual + activation + up2(post) + scale, 128-byte
// coalesced float4 stores.  Several CTAs are resident per SM so load / split / MMA / epilogue phases of different
// tiles overlap.
#include "gfr_common.cuh"
#include "tc_common.cuh"

#include <mutex>
#include <string.h>

using namespace gfr_tc;

namespace {

constexpr int TILE_PX_W = 8, TILE_PX_H = 16;           // 128 output pixels = UMMA M
constexpr int HALO_W = TILE_PX_W + 2, HALO_H = TILE_PX_H + 2;
constexpr int CB = 16;                                  // input channels per pipeline step (4 groups of 4)
constexpr uint32_t A_BYTES = (CB / 4) * HALO_H * HALO_W * 16;     // 11520
constexpr uint32_t A_LBO = HALO_H * HALO_W * 16;        // 2880: next 4-channel group
constexpr uint32_t A_SBO = HALO_W * 16;                 // 160: next tile row (8 pixels further in M)

struct ConvTcArgs {
  const float* wpk;    // packed weights, see gfr_conv_tc_pack_weights
  const float* bias;   // [Cout]
  const float* res;    // C4 [N][C4out][H][W][4] or null   (added before the activation)
  const float* post;   // C4 [N][C4out][H>>ps][W>>ps][4] or null (added after the activation, nearest x2 when ps=1)
  float* out;          // C4 [N][C4out][H][W][4]
  int N, Cin, Cout, H, W;
  int ncb, tiles_x, tiles_y, m_tiles;
  int post_shift, act;
  int products;        // bit 0: hi*Whi, bit 1: lo*Whi, bit 2: hi*Wlo  (7 = 3xTF32, 1 = single-pass TF32)
  float out_scale;
};

template <int NT>
struct Smem {
  static constexpr uint32_t W_HALF = 9 * (CB / 4) * NT * 16;       // hi (or lo) weights of one 16-channel step
  static constexpr uint32_t OFF_A_HI = 0, OFF_A_LO = A_BYTES, OFF_W = 2 * A_BYTES;
  static constexpr uint32_t OFF_BAR = OFF_W + 2 * W_HALF;
  static constexpr uint32_t BYTES = OFF_BAR + 32;
  static constexpr uint32_t TMEM_COLS = NT <= 32 ? 32 : (NT <= 64 ? 64 : 128);
};

template <int NT>
__global__ void __launch_bounds__(128) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const ConvTcArgs a) {
  using S = Smem<NT>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sA_hi = smem_u32(smem + S::OFF_A_HI), sA_lo = smem_u32(smem + S::OFF_A_LO);
  const uint32_t sW = smem_u32(smem + S::OFF_W);
  const uint32_t bar_full = smem_u32(smem + S::OFF_BAR), bar_done = bar_full + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + S::OFF_BAR + 16);

  if (tid == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_done, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_in);
  }
  if (warp == 0) tmem_alloc(smem_u32(smem + S::OFF_BAR + 16), S::TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  const int n0 = blockIdx.y * NT;                       // first output channel of this CTA
  const int C4out = (a.Cout + 3) >> 2;
  const float* wsrc = a.wpk + (size_t)blockIdx.y * a.ncb * (2 * S::W_HALF / 4);
  constexpr uint32_t IDESC = umma_idesc_tf32(128, NT);
  uint32_t ph_full = 0, ph_done = 0;
  bool w_loaded = false;

  for (int mt = blockIdx.x; mt < a.m_tiles; mt += gridDim.x) {
    const int tx = mt % a.tiles_x;
    const int t2 = mt / a.tiles_x;
    const int ty = t2 % a.tiles_y;
    const int n = t2 / a.tiles_y;
    const int x0 = tx * TILE_PX_W, y0 = ty * TILE_PX_H;

    for (int cb = 0; cb < a.ncb; ++cb) {
      if (tid == 0) {
        const bool need_w = (a.ncb > 1) || !w_loaded;
        mbar_expect_tx(bar_full, A_BYTES + (need_w ? 2 * S::W_HALF : 0));
        tma_load_5d(sA_hi, &tm_in, bar_full, 0, x0 - 1, y0 - 1, cb * (CB / 4), n);
        if (need_w) bulk_load(sW, wsrc + (size_t)cb * (2 * S::W_HALF / 4), 2 * S::W_HALF, bar_full);
      }
      w_loaded = true;
      mbar_wait(bar_full, ph_full);
      ph_full ^= 1;

      // ---- split the fp32 tile into tf32 hi (in place) and lo
      {
        float4* hi4 = reinterpret_cast<float4*>(smem + S::OFF_A_HI);
        float4* lo4 = reinterpret_cast<float4*>(smem + S::OFF_A_LO);
#pragma unroll
        for (int i = tid; i < (int)(A_BYTES / 16); i += 128) {
          const float4 v = hi4[i];
          float4 h, l;
          h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
          l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
          hi4[i] = h;
          lo4[i] = l;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncthreads();

      if (tid == 0) {
        tc_fence_after_sync();
        uint32_t acc = cb > 0 ? 1u : 0u;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t aoff = (uint32_t)((tap / 3) * HALO_W + (tap % 3)) * 16u;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint32_t ao = aoff + (uint32_t)j * 2u * A_LBO;
            const uint32_t bo = (uint32_t)tap * ((CB / 4) * NT * 16u) + (uint32_t)j * 2u * (NT * 16u);
            const uint64_t dA_hi = umma_desc_kmajor_noswz(sA_hi + ao, A_LBO, A_SBO);
            const uint64_t dA_lo = umma_desc_kmajor_noswz(sA_lo + ao, A_LBO, A_SBO);
            const uint64_t dB_hi = umma_desc_kmajor_noswz(sW + bo, NT * 16u, 128u);
            const uint64_t dB_lo = umma_desc_kmajor_noswz(sW + S::W_HALF + bo, NT * 16u, 128u);
            if (a.products & 2) { umma_tf32(tmem, dA_lo, dB_hi, IDESC, acc); acc = 1u; }
            if (a.products & 4) { umma_tf32(tmem, dA_hi, dB_lo, IDESC, acc); acc = 1u; }
            if (a.products & 1) { umma_tf32(tmem, dA_hi, dB_hi, IDESC, acc); acc = 1u; }
          }
        }
        umma_commit(bar_done);
      }
      mbar_wait(bar_done, ph_done);
      ph_done ^= 1;
    }

    // ---- epilogue: thread = pixel (TMEM lane), NT accumulators
    tc_fence_after_sync();
    {
      const int m = warp * 32 + lane;
      const int y = y0 + (m >> 3), x = x0 + (m & 7);
      const bool ok = y < a.H && x < a.W;
      const int pH = a.H >> a.post_shift, pW = a.W >> a.post_shift;
#pragma unroll
      for (int g = 0; g < NT / 16; ++g) {
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(g * 16), r);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int co = n0 + g * 16 + q * 4;
          const int cq = co >> 2;
          if (!ok || cq >= C4out) continue;
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = __uint_as_float(r[q * 4 + e]) + (co + e < a.Cout ? __ldg(a.bias + co + e) : 0.f);
          const size_t o = ((((size_t)n * C4out + cq) * a.H + y) * a.W + x) * 4;
          if (a.res) {
            const float4 rr = __ldg(reinterpret_cast<const float4*>(a.res + o));
            v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
          }
          if (a.act == 1) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = v[e] > 0.f ? v[e] : 0.2f * v[e];
          } else if (a.act == 2) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = 1.0f / (1.0f + expf(-v[e]));
          }
          if (a.post) {
            const size_t po = ((((size_t)n * C4out + cq) * pH + (y >> a.post_shift)) * pW + (x >> a.post_shift)) * 4;
            const float4 pp = __ldg(reinterpret_cast<const float4*>(a.post + po));
            v[0] += pp.x; v[1] += pp.y; v[2] += pp.z; v[3] += pp.w;
          }
          float4 ov;
          ov.x = v[0] * a.out_scale; ov.y = v[1] * a.out_scale; ov.z = v[2] * a.out_scale; ov.w = v[3] * a.out_scale;
          *reinterpret_cast<float4*>(a.out + o) = ov;
        }
      }
    }
    tc_fence_before_sync();
    __syncthreads();          // every TMEM read of this tile is done before the next tile's first MMA overwrites it
  }
  if (warp == 0) tmem_dealloc(tmem, S::TMEM_COLS);
}

// ---- TMA descriptor for a C4 activation tensor ---------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

// tensor [N][C4][H][W][4] fp32; box {4, bw, bh, 4 groups, 1}
int make_c4_map(CUtensorMap* tm, const float* base, int N, int C4, int H, int W, int bw, int bh) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return GFR_E_UNSUPPORTED;
  const cuuint64_t dims[5] = {4, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C4, (cuuint64_t)N};
  const cuuint64_t strides[4] = {16, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)C4 * H * W * 16};
  const cuuint32_t box[5] = {4, (cuuint32_t)bw, (cuuint32_t)bh, CB / 4, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? GFR_OK : GFR_E_ARG;
}

template <int NT>
int launch_tc(const CUtensorMap& tm, const ConvTcArgs& a, cudaStream_t s) {
  using S = Smem<NT>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  const int n_tiles = gfr_ceil_div(a.Cout, NT);
  int occ = (int)(220u * 1024u / S::BYTES);
  if (occ > 4) occ = 4;
  if (occ < 1) occ = 1;
  int gx = (sm_count() * occ) / n_tiles;
  if (gx < 1) gx = 1;
  if (gx > a.m_tiles) gx = a.m_tiles;
  conv3x3_tc_kernel<NT><<<dim3(gx, n_tiles), 128, S::BYTES, s>>>(tm, a);
  return gfr_launch_status();
}

inline float tf32_rn_host(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;
  u += 0xFFFu + ((u >> 13) & 1u);
  u &= ~0x1FFFu;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

// ---- layout helpers ------------------------------------------------------------------------------------
__global__ void nchw_to_c4_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int HW, long long total4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over [N][C4][HW]
  if (i >= total4) return;
  const int C4 = (C + 3) >> 2;
  const int p = (int)(i % HW);
  const long long t = i / HW;
  const int cq = (int)(t % C4);
  const long long n = t / C4;
  float v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = cq * 4 + e;
    v[e] = c < C ? __ldg(in + (n * C + c) * HW + p) : 0.f;
  }
  reinterpret_cast<float4*>(out)[i] = make_float4(v[0], v[1], v[2], v[3]);
}

__global__ void c4_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int HW, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over [N][C][HW]
  if (i >= total) return;
  const int C4 = (C + 3) >> 2;
  const int p = (int)(i % HW);
  const long long t = i / HW;
  const int c = (int)(t % C);
  const long long n = t / C;
  out[i] = __ldg(in + (((n * C4 + (c >> 2)) * HW + p) << 2) + (c & 3));
}

__global__ void maxpool2_c4_kernel(const float4* __restrict__ in, float4* __restrict__ out, long long n_out, int Ho, int Wo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over [NC4][Ho][Wo]
  if (i >= n_out) return;
  const int x = (int)(i % Wo);
  const long long t = i / Wo;
  const int y = (int)(t % Ho);
  const long long nc = t / Ho;
  const float4* p = in + (nc * (2 * Ho) + 2 * y) * (2LL * Wo) + 2 * x;
  const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2 * Wo), d = __ldg(p + 2 * Wo + 1);
  out[i] = make_float4(fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y)),
                       fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z)), fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w)));
}

__global__ void upsample2_c4_kernel(const float4* __restrict__ in, const float4* __restrict__ add, float4* __restrict__ out,
                                    long long n_out, int Ho, int Wo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const int x = (int)(i % Wo);
  const long long t = i / Wo;
  const int y = (int)(t % Ho);
  const long long nc = t / Ho;
  float4 v = __ldg(in + (nc * (Ho >> 1) + (y >> 1)) * (long long)(Wo >> 1) + (x >> 1));
  if (add) {
    const float4 w = __ldg(add + i);
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
  }
  out[i] = v;
}

}  // namespace

extern "C" long long gfr_conv_tc_pack_size(int Cin, int Cout, int NT) {
  if (Cin <= 0 || Cout <= 0 || (NT != 16 && NT != 32 && NT != 64)) return GFR_E_ARG;
  return (long long)gfr_ceil_div(Cout, NT) * gfr_ceil_div(Cin, CB) * 2 * 9 * (CB / 4) * NT * 4;
}

extern "C" int gfr_conv_tc_pack_weights(const float* w_host, int Cin, int Cout, int NT, float* packed_host) {
  GFR_RETURN_IF_NULL(w_host); GFR_RETURN_IF_NULL(packed_host);
  if (Cin <= 0 || Cout <= 0 || (NT != 16 && NT != 32 && NT != 64)) return GFR_E_ARG;
  const int n_tiles = gfr_ceil_div(Cout, NT), ncb = gfr_ceil_div(Cin, CB);
  size_t o = 0;
  for (int nt = 0; nt < n_tiles; ++nt)
    for (int cb = 0; cb < ncb; ++cb)
      for (int part = 0; part < 2; ++part)
        for (int tap = 0; tap < 9; ++tap)
          for (int kc = 0; kc < CB / 4; ++kc)
            for (int n = 0; n < NT; ++n)
              for (int e = 0; e < 4; ++e, ++o) {
                const int co = nt * NT + n, ci = cb * CB + kc * 4 + e;
                float v = 0.f;
                if (co < Cout && ci < Cin) {
                  const float w = w_host[((size_t)co * Cin + ci) * 9 + tap];
                  const float hi = tf32_rn_host(w);
                  v = part == 0 ? hi : tf32_rn_host(w - hi);
                }
                packed_host[o] = v;
              }
  return GFR_OK;
}

extern "C" int gfr_conv3x3_tc_fwd(const float* in, const float* w_packed, const float* bias, const float* res,
                                  const float* post, float* out, int N, int Cin, int Cout, int H, int W, int NT,
                                  int post_shift, int act, float out_scale, int precision, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(w_packed); GFR_RETURN_IF_NULL(bias); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  if (post_shift < 0 || post_shift > 1 || act < 0 || act > 2 || precision < 1 || precision > 7) return GFR_E_ARG;
  if (post && post_shift && ((H | W) & 1)) return GFR_E_SHAPE;
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(w_packed) | reinterpret_cast<uintptr_t>(out) |
       reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(post)) & 15)
    return GFR_E_ARG;
  ConvTcArgs a;
  a.wpk = w_packed; a.bias = bias; a.res = res; a.post = post; a.out = out;
  a.N = N; a.Cin = Cin; a.Cout = Cout; a.H = H; a.W = W;
  a.ncb = gfr_ceil_div(Cin, CB);
  a.tiles_x = gfr_ceil_div(W, TILE_PX_W); a.tiles_y = gfr_ceil_div(H, TILE_PX_H);
  a.m_tiles = N * a.tiles_x * a.tiles_y;
  a.post_shift = post_shift; a.act = act; a.out_scale = out_scale;
  a.products = precision == 3 ? 7 : precision;      // 3 is the documented alias of 'all three products'
  CUtensorMap tm;
  const int rc = make_c4_map(&tm, in, N, (Cin + 3) / 4, H, W, HALO_W, HALO_H);
  if (rc != GFR_OK) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  switch (NT) {
    case 16: return launch_tc<16>(tm, a, s);
    case 32: return launch_tc<32>(tm, a, s);
    case 64: return launch_tc<64>(tm, a, s);
    default: return GFR_E_ARG;
  }
}

extern "C" int gfr_nchw_to_c4(const float* in, float* out, int N, int C, int H, int W, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  const long long total = (long long)N * ((C + 3) / 4) * H * W;
  nchw_to_c4_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, out, C, H * W, total);
  return gfr_launch_status();
}

extern "C" int gfr_c4_to_nchw(const float* in, float* out, int N, int C, int H, int W, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  const long long total = (long long)N * C * H * W;
  c4_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, out, C, H * W, total);
  return gfr_launch_status();
}

extern "C" int gfr_maxpool2_c4_fwd(const float* in, float* out, int NC4, int Ho, int Wo, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(out);
  if (NC4 <= 0 || Ho <= 0 || Wo <= 0) return GFR_E_SHAPE;
  const long long n = (long long)NC4 * Ho * Wo;
  maxpool2_c4_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), n, Ho, Wo);
  return gfr_launch_status();
}

extern "C" int gfr_upsample2_c4_fwd(const float* in, const float* add, float* out, int NC4, int Ho, int Wo, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(out);
  if (NC4 <= 0 || Ho <= 0 || Wo <= 0 || ((Ho | Wo) & 1)) return GFR_E_SHAPE;
  const long long n = (long long)NC4 * Ho * Wo;
  upsample2_c4_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(in), reinterpret_cast<const float4*>(add), reinterpret_cast<float4*>(out), n, Ho, Wo);
  return gfr_launch_status();
}
