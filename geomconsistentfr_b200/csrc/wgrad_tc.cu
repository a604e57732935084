// Weight gradient of the 3x3 (and PatchGAN's 2x2-tap) convolutions on the 5th-generation tensor cores, bf16 operands, fp32
// accumulation in TMEM — the training path's `train_precision = 4` mode (BASELINE configs[2] "bf16 CNN"); replaces what
// autograd + cuDNN's backward-filter do for TRAIN:197-350 / TRAIN:15-35 under `total_loss.backward()` (TRAIN:655).
//
//     dW[co][ci][tap] += sum_{n,y,x} g[n,co,y,x] * in[n,ci,y+ky-org,x+kx-org]
//
// is, per filter tap, a GEMM D[M = co][N = ci] = A[co][k] * B[ci][k]^T whose contraction index k runs over PIXELS.  In the C4
// activation layout a pixel's channels are contiguous and consecutive pixels of a tile row are 16 bytes apart, i.e. both
// operands are "MN-major" UMMA matrices: a 16-byte unit holds 8 consecutive bf16 channels of one pixel, 8 consecutive pixels
// form the 128-byte core matrix (canonical no-swizzle layout ((1,n),(8,k)):((X,SBO),(1,LBO)) in 16-byte units: SBO = stride
// between channel chunks, LBO = stride between groups of 8 pixels) — and a filter tap is again just a shifted start address
// into the same halo tile ((ky*10 + kx) * 16 bytes), like in the forward kernel.
//
//   warps 0-7  producers: coalesced float4 loads of the fp32 C4 tensors (zero fill at the borders) -> bf16 -> 16-byte shared-
//              memory stores in the operand layout [chunk of 8 ch][row][px][8], fence.proxy.async, mbarrier arrive
//   warp 8     MMA issuer (elect.sync): per 128-pixel tile and tap, 8 x tcgen05.mma (M = 128 co, N = NB ci, K = 16 pixels =
//              two tile rows); the TAPS accumulators [128 x NB] live side by side in TMEM for the whole CTA lifetime
//   warps 0-3  epilogue (after the last tile): tcgen05.ld -> one atomicAdd per (co, ci, tap) into the parameter-layout gradient
//
// A CTA owns (a 128-channel slice of co) x (an NB-channel slice of ci) x (a strided subset of the pixel tiles).  M is always
// 128: for the 16-output-channel layers 7/8 of the A operand is zero padding — the tensor-core time that costs is far below
// what the CUDA-core kernel (cnn_train.cu, kept for the fp32-grade 3xTF32 mode) needs for the same layer.
#include "gfr_common.cuh"
#include "tc_common.cuh"

#include <cuda_bf16.h>
#include <mutex>
#include <stdlib.h>

using namespace gfr_tc;

namespace {

constexpr int TW = 8, TH = 16;                     // pixel tile (128 pixels = 8 K-steps of 16)
constexpr int HW_ = TW + 2, HH = TH + 2;           // halo tile of the layer input
#ifndef GFR_WGRAD_GROUPS
#define GFR_WGRAD_GROUPS 2
#endif
constexpr int N_PROD = 128 * GFR_WGRAD_GROUPS, N_THREADS = N_PROD + 32, PROD_WARPS = N_PROD / 32;
constexpr uint32_t G_CHUNK = TH * TW * 16;         // 2048: one 8-channel chunk of the g tile
constexpr uint32_t G_BYTES = 16 * G_CHUNK;         // M = 128 rows = 16 chunks (M = 64: the first 8), always addressed by the MMA
constexpr int N_GROUPS = GFR_WGRAD_GROUPS, GROUP_THREADS = N_PROD / N_GROUPS;   // producer groups of 128 threads take alternate tiles (one tile's loads in flight per group; -DGFR_WGRAD_GROUPS=n, n <= the ring depth)
static_assert(GROUP_THREADS == 128 && N_GROUPS >= 1 && N_GROUPS <= 4, "groups of four warps, at most one per ring slot");
constexpr uint32_t X_CHUNK = HH * HW_ * 16;        // 2880
constexpr int MAX_STAGES = 4;
#ifndef GFR_WGRAD_XSHIFT_DEFAULT
#define GFR_WGRAD_XSHIFT_DEFAULT 0
#endif

struct WgradTcArgs {
  const float* in; const float* g;     // C4 [N][in_groups][Hin][Win][4], C4 [N][ceil(Cout/4)][H][W][4]
  float* dw;                           // element (co, ci, tap) at dw[co*so + ci*si + (flip ? TAPS - 1 - tap : tap)]  +=
  int N, Cin, Cout, in_groups, H, W, Hin, Win;
  long long so, si; int flip;
  int tiles_x, tiles_y, n_tiles;
  int NB, stages, org;
  int M;                               // UMMA M: 128, or 64 when Cout <= 64 (half the A-operand reads; D rows 16 q + i live in TMEM lanes 32 q + i)
  int xshift;                          // 1 (3x3, Cin <= 16): the INPUT is the M operand and its row blocks are pixel shifts — see the MMA issuer
  int Ng;                              // xshift: N = output channels of the CTA's slice, rounded up to 16
};

__host__ __device__ constexpr uint32_t idesc_bf16_mnmajor(int M, int N) {
  // kind::f16, fp32 accumulate, bf16 A and B, both MN-major (bits 15, 16)
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// no-swizzle shared-memory descriptor, MN-major: SBO = stride between 16-byte units along M/N (channel chunks), LBO = stride
// between groups of 8 along K (pixel groups)
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t start, uint32_t lbo, uint32_t sbo) {
  const uint32_t lo = ((start >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
  const uint32_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);
  return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ uint4 pack_bf16x8(const float4& a, const float4& b) {
  const __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
  const __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
  return make_uint4(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1),
                    *reinterpret_cast<const uint32_t*>(&p2), *reinterpret_cast<const uint32_t*>(&p3));
}

template <int TAPS>
__global__ void __launch_bounds__(N_THREADS, 1) wgrad_tc_kernel(const WgradTcArgs a) {
  constexpr int KT = TAPS == 9 ? 3 : 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NB = a.NB, STAGES = a.stages;
  const uint32_t x_bytes = (uint32_t)(NB >> 3) * X_CHUNK;
  const uint32_t stage_bytes = G_BYTES + x_bytes;
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t bars = smem0 + STAGES * stage_bytes;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * MAX_STAGES, bar_done = bars + 16 * MAX_STAGES;
  constexpr uint32_t TMEM_SLOT = 16 * MAX_STAGES + 16;
  uint8_t* gen_bars = smem + (bars - smem0);

  // the g operand is always addressed as 16 chunks: the chunks this layer does not have stay zero for the CTA's lifetime
  for (uint32_t i = tid; i < (uint32_t)STAGES * stage_bytes / 16; i += N_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, GROUP_THREADS);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  if (warp == PROD_WARPS) tmem_alloc(bars + TMEM_SLOT, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen_bars + TMEM_SLOT);

  const int cob = blockIdx.y * a.M, cib = blockIdx.z * NB;
  const int C4out = (a.Cout + 3) >> 2;
  const int co_chunks = min(a.M >> 3, (a.Cout - cob + 7) >> 3);
  const size_t gplane = (size_t)a.H * a.W, iplane = (size_t)a.Hin * a.Win;
  const int n_my = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp < PROD_WARPS) {
    // =============================== producers ===============================
    const float4* gsrc = reinterpret_cast<const float4*>(a.g);
    const float4* isrc = reinterpret_cast<const float4*>(a.in);
    const int grp_id = tid / GROUP_THREADS, gtid = tid % GROUP_THREADS;
    const int n_groups = N_GROUPS < STAGES ? N_GROUPS : STAGES;        // groups beyond that stay idle
    int s = 0, ph = 0, it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      // never more groups than ring slots: a group that ran two phases ahead of its slot's `empty` barrier would pass the parity wait
      const bool mine = (it % n_groups) == grp_id;
      if (!mine) { if (++s == STAGES) { s = 0; ph ^= 1; } continue; }
      const int tx = tile % a.tiles_x, t2 = tile / a.tiles_x;
      const int ty = t2 % a.tiles_y, n = t2 / a.tiles_y;
      const int x0 = tx * TW, y0 = ty * TH;
      if (it >= STAGES) mbar_wait(bar_empty + 8 * s, ph ^ 1);
      uint8_t* st = smem + (size_t)s * stage_bytes;
      // One task = one 16-byte operand unit (8 bf16 channels of one pixel) = two float4 loads.  Tasks [0, n_g) fill the g tile
      // (chunk c8 holds channels cob + 8 c8 .. + 7 = C4 groups 2 c8', 2 c8' + 1), tasks [n_g, n_g + n_x) the halo tile of the
      // layer input (chunk c8 holds input channels cib + 8 c8 .. + 7).  A thread takes LD_BATCH tasks at a time and issues all
      // of their loads before the first conversion: the loop used to run one task per iteration (two loads in flight per
      // thread), which made every tile cost five L2 / DRAM latencies per producer group — the 16-channel layers at 256^2 ran
      // at 2.2 us per tile against 0.7 us of MMA time.
      constexpr int LD_BATCH = 5;
      const int n_g = co_chunks * (TH * TW), n_tasks = n_g + (NB >> 3) * (HH * HW_);
      uint8_t* xt = st + G_BYTES;
      for (int base = gtid; base < n_tasks; base += GROUP_THREADS * LD_BATCH) {
        float4 v0[LD_BATCH], v1[LD_BATCH];
        uint8_t* dst[LD_BATCH];
#pragma unroll
        for (int u = 0; u < LD_BATCH; ++u) {
          const int i = base + u * GROUP_THREADS;
          v0[u] = make_float4(0.f, 0.f, 0.f, 0.f); v1[u] = v0[u];
          dst[u] = nullptr;
          if (i < n_g) {
            const int c8 = i / (TH * TW), pix = i % (TH * TW);
            const int gy = y0 + (pix >> 3), gx = x0 + (pix & 7);
            const int grp = ((cob >> 3) + c8) * 2;
            dst[u] = st + (size_t)c8 * G_CHUNK + (size_t)pix * 16;
            if (gy < a.H && gx < a.W) {
              const size_t o = (size_t)gy * a.W + gx;
              if (grp < C4out) v0[u] = __ldg(gsrc + ((size_t)n * C4out + grp) * gplane + o);
              if (grp + 1 < C4out) v1[u] = __ldg(gsrc + ((size_t)n * C4out + grp + 1) * gplane + o);
            }
          } else if (i < n_tasks) {
            const int j = i - n_g;
            const int c8 = j / (HH * HW_), pix = j % (HH * HW_);
            const int gy = y0 + pix / HW_ - a.org, gx = x0 + pix % HW_ - a.org;
            const int grp = ((cib >> 3) + c8) * 2;
            dst[u] = xt + (size_t)c8 * X_CHUNK + (size_t)pix * 16;
            if (gy >= 0 && gy < a.Hin && gx >= 0 && gx < a.Win) {
              const size_t o = (size_t)gy * a.Win + gx;
              if (grp * 4 < a.Cin) v0[u] = __ldg(isrc + ((size_t)n * a.in_groups + grp) * iplane + o);
              if ((grp + 1) * 4 < a.Cin) v1[u] = __ldg(isrc + ((size_t)n * a.in_groups + grp + 1) * iplane + o);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < LD_BATCH; ++u)
          if (dst[u] != nullptr) *reinterpret_cast<uint4*>(dst[u]) = pack_bf16x8(v0[u], v1[u]);
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_full + 8 * s);
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else {
    // =============================== MMA issuer ===============================
    const uint32_t idesc = idesc_bf16_mnmajor(a.M, NB);
    int s = 0, ph = 0;
    for (int it = 0; it < n_my; ++it) {
      mbar_wait(bar_full + 8 * s, ph);
      tc_fence_after_sync();
      if (elect_one_sync()) {
        const uint32_t gt = smem0 + s * stage_bytes, xt = gt + G_BYTES;
        if (a.xshift) {
          // 16-input-channel layers (the 128^2 / 256^2 ones): A = the input halo tile with SBO = 16 bytes, i.e. M row block j is
          // the SAME 8-channel chunk one pixel further along the row — the three horizontal taps of a filter row come out of
          // one MMA as D rows (j, c) = dW[co][c][ky][kx = j] (rows j >= 3 are never read), B = g (N = output channels).  Six
          // MMAs per K block (3 filter rows x 2 input chunks) instead of nine with 7/8 of the M operand zero padding.
          const uint32_t idx = idesc_bf16_mnmajor(64, a.Ng);
#pragma unroll
          for (int r2 = 0; r2 < TH / 2; ++r2) {
            const uint64_t dG = desc_mnmajor(gt + (uint32_t)(2 * r2) * (TW * 16), TW * 16, G_CHUNK);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
              for (int ch = 0; ch < 2; ++ch) {
                const uint64_t dX = desc_mnmajor(xt + (uint32_t)ch * X_CHUNK + (uint32_t)(ky * HW_) * 16u + (uint32_t)(2 * r2) * (HW_ * 16), HW_ * 16, 16u);
                umma_f16(tmem + (uint32_t)((ky * 2 + ch) * a.Ng), dX, dG, idx, (it == 0 && r2 == 0) ? 0u : 1u);
              }
            }
          }
          umma_commit(bar_empty + 8 * s);
          if (it == n_my - 1) umma_commit(bar_done);
        } else {
        // K block outermost, taps innermost: consecutive MMAs accumulate into DIFFERENT TMEM accumulators (one per tap), so
        // they pipeline; with the taps outermost each accumulator received its 8 K blocks back to back, a dependent chain
        // that ran at the MMA latency (1.5 us per tile of 72 tiny N = 16 MMAs)
#pragma unroll
        for (int r2 = 0; r2 < TH / 2; ++r2) {            // K = 16 pixels = tile rows 2 r2, 2 r2 + 1
          const uint64_t dA = desc_mnmajor(gt + (uint32_t)(2 * r2) * (TW * 16), TW * 16, G_CHUNK);
#pragma unroll
          for (int tap = 0; tap < TAPS; ++tap) {
            const uint32_t xoff = (uint32_t)((tap / KT) * HW_ + (tap % KT)) * 16u;
            const uint64_t dB = desc_mnmajor(xt + xoff + (uint32_t)(2 * r2) * (HW_ * 16), HW_ * 16, X_CHUNK);
            umma_f16(tmem + (uint32_t)(tap * NB), dA, dB, idesc, (it == 0 && r2 == 0) ? 0u : 1u);
          }
        }
        umma_commit(bar_empty + 8 * s);
        if (it == n_my - 1) umma_commit(bar_done);
        }
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  }
  // =============================== epilogue (xshift): D row 8 j + c of accumulator (ky, chunk) = dW[.][ci = 8 chunk + c][ky][kx = j] ===============
  if (a.xshift) {
    if (warp < 2 && n_my > 0) {
      mbar_wait(bar_done, 0);
      tc_fence_after_sync();
      // M = 64: rows 0..15 are TMEM lanes 0..15 (warp 0), rows 16..23 lanes 32..39 (warp 1)
      const int row = warp * 16 + lane, j = row >> 3, c = row & 7;
      const bool useful = lane < 16 && row < 24;
      for (int acc = 0; acc < 6; ++acc) {
        const int ky = acc >> 1, ch = acc & 1;
        const int ci = cib + ch * 8 + c;
        const int tap = ky * 3 + j, t_out = a.flip ? 8 - tap : tap;
        for (int c0 = 0; c0 < a.Ng; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * a.Ng + c0), r);
          tmem_ld_wait();
          if (useful && ci < a.Cin) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const int co = cob + c0 + k;
              if (co < a.Cout) atomicAdd(a.dw + co * a.so + ci * a.si + t_out, __uint_as_float(r[k]));
            }
          }
        }
      }
    }
  } else
  // =============================== epilogue: warps 0-3, TMEM lane = co ===============================
  // The accumulators leave TMEM with one output channel per LANE, but in the parameter layout consecutive addresses run over
  // (ci, tap) [Conv2d] or (co, tap) [ConvTranspose2d]: a direct atomicAdd per lane scatters every warp instruction over 32 cache
  // lines (measured: the wide layers spent ~30 us in this epilogue whatever their pixel count).  So a warp stages 16 input
  // channels x TAPS x its rows through shared memory (the operand ring is free once bar_done has fired) and adds them to global
  // memory with consecutive lanes on consecutive addresses.
  if (warp < 4 && n_my > 0) {
    mbar_wait(bar_done, 0);
    tc_fence_after_sync();
    // M = 128: D row r is TMEM lane r.  M = 64: rows 16 q .. 16 q + 15 are lanes 32 q .. 32 q + 15 (the other lanes are unused)
    const int rows = a.M == 128 ? 32 : 16;
    const int co0 = cob + warp * rows;
    const int n_co = min(rows, a.Cout - co0);
    float* st = reinterpret_cast<float*>(smem) + warp * (16 * TAPS * 33);        // [ci_l * TAPS + tap_out][33]: co_l fastest, padded
    for (int c0 = 0; c0 < NB; c0 += 16) {
      for (int tap = 0; tap < TAPS; ++tap) {
        const int t_out = a.flip ? TAPS - 1 - tap : tap;
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(tap * NB + c0), r);
        tmem_ld_wait();
        if (lane < rows) {
#pragma unroll
          for (int k = 0; k < 16; ++k) st[(k * TAPS + t_out) * 33 + lane] = __uint_as_float(r[k]);
        }
      }
      __syncwarp();
      const int n_ci = min(16, a.Cin - (cib + c0));
      if (n_ci > 0 && n_co > 0) {
        if (!a.flip) {            // Conv2d parameter [co][ci][tap]: per output channel a run of n_ci * TAPS consecutive floats
          for (int co_l = 0; co_l < n_co; ++co_l) {
            float* dst = a.dw + (long long)(co0 + co_l) * a.so + (long long)(cib + c0) * a.si;
            for (int L = lane; L < n_ci * TAPS; L += 32) atomicAdd(dst + L, st[L * 33 + co_l]);
          }
        } else {                  // ConvTranspose2d parameter [ci][co][tap]: per input channel a run of n_co * TAPS consecutive floats
          for (int ci_l = 0; ci_l < n_ci; ++ci_l) {
            float* dst = a.dw + (long long)(cib + c0 + ci_l) * a.si + (long long)co0 * a.so;
            for (int L = lane; L < n_co * TAPS; L += 32) {
              const int co_l = L / TAPS, t = L - co_l * TAPS;
              atomicAdd(dst + L, st[(ci_l * TAPS + t) * 33 + co_l]);
            }
          }
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == PROD_WARPS) tmem_dealloc(tmem, 512);
}

int g_wgrad_xshift = -1;      // -1 default (environment GFR_WGRAD_XSHIFT, else the built-in choice), 0 / 1 (gfr_wgrad_tc_config)

template <int TAPS>
int launch_wgrad_tc(WgradTcArgs a, cudaStream_t s) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(wgrad_tc_kernel<TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (attr_err != cudaSuccess) return (int)attr_err;
  // ci slice: as wide as the TMEM accumulators allow (TAPS * NB <= 512 columns), a multiple of 16
  int NB = (512 / TAPS) & ~15;
  if (NB > 128) NB = 128;
  const int cin16 = (a.Cin + 15) & ~15;
  if (NB > cin16) NB = cin16;
  a.NB = NB;
  a.M = a.Cout <= 64 ? 64 : 128;
  // the pixel-shift form for the 16-input-channel 3x3 layers (one CTA covers every output channel: Cout <= 64)
  static const bool env_xshift = [] { const char* e = getenv("GFR_WGRAD_XSHIFT"); return e ? atoi(e) != 0 : GFR_WGRAD_XSHIFT_DEFAULT != 0; }();
  const bool want_xshift = g_wgrad_xshift >= 0 ? g_wgrad_xshift != 0 : env_xshift;
  a.xshift = (want_xshift && TAPS == 9 && a.Cin <= 16 && a.Cout <= 64) ? 1 : 0;
  a.Ng = (a.Cout + 15) & ~15;
  const uint32_t stage = G_BYTES + (uint32_t)(NB >> 3) * X_CHUNK;
  int stages = (int)((220u * 1024u) / stage);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return GFR_E_UNSUPPORTED;
  a.stages = stages;
  const int gy = gfr_ceil_div(a.Cout, a.M), gz = gfr_ceil_div(a.Cin, NB);
  int gx = 148 / (gy * gz);                // one CTA per SM (all 512 TMEM columns); pixel splits: every split adds one round of atomics on the whole (co, ci, tap) block
  if (gx < 1) gx = 1;
  if (gx > a.n_tiles) gx = a.n_tiles;
  if (4u * 16u * TAPS * 33u * 4u > (uint32_t)stages * stage) return GFR_E_UNSUPPORTED;      // the epilogue's staging tiles reuse the ring
  wgrad_tc_kernel<TAPS><<<dim3(gx, gy, gz), N_THREADS, (size_t)stages * stage + 256, s>>>(a);
  return gfr_launch_status();
}

}  // namespace

extern "C" int gfr_wgrad_tc_config(int pixel_shift_form) {
  if (pixel_shift_form < -1 || pixel_shift_form > 1) return GFR_E_ARG;
  g_wgrad_xshift = pixel_shift_form;
  return GFR_OK;
}

extern "C" int gfr_conv_wgrad_tc_bf16(const float* in, const float* g_out, float* g_w, int is_transposed_conv, int N, int Cin,
                                      int in_groups, int Cout, int Hin, int Win, int H, int W, int taps, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(g_out); GFR_RETURN_IF_NULL(g_w);
  if (N <= 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0 || Hin <= 0 || Win <= 0) return GFR_E_SHAPE;
  if (taps != 9 && taps != 4) return GFR_E_ARG;
  if (taps == 9 && (Hin != H || Win != W)) return GFR_E_SHAPE;
  if (taps == 4 && (Hin != H + 1 || Win != W + 1)) return GFR_E_SHAPE;
  if (in_groups == 0) in_groups = (Cin + 3) / 4;
  if (in_groups < (Cin + 3) / 4) return GFR_E_ARG;
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(g_out)) & 15) return GFR_E_ARG;
  WgradTcArgs a;
  a.in = in; a.g = g_out; a.dw = g_w;
  a.N = N; a.Cin = Cin; a.Cout = Cout; a.in_groups = in_groups; a.H = H; a.W = W; a.Hin = Hin; a.Win = Win;
  if (!is_transposed_conv) { a.so = (long long)Cin * taps; a.si = taps; a.flip = 0; }     // dW_param[co][ci][tap]
  else { a.so = taps; a.si = (long long)Cout * taps; a.flip = 1; }                         // dW_param[ci][co][taps - 1 - tap]
  a.tiles_x = gfr_ceil_div(W, TW); a.tiles_y = gfr_ceil_div(H, TH); a.n_tiles = N * a.tiles_x * a.tiles_y;
  a.org = taps == 9 ? 1 : 0;
  a.NB = 0; a.stages = 0; a.M = 128;
  return taps == 9 ? launch_wgrad_tc<9>(a, (cudaStream_t)stream) : launch_wgrad_tc<4>(a, (cudaStream_t)stream);
}
