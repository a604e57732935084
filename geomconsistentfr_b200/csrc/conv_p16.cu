// 3x3 convolution of the EVAL-mode CNN on the 5th-generation tensor cores (tcgen05, kind::f16) over PRE-SPLIT fp16-pair
// activations (the P16 layout, p16.cuh).  Replaces the cuDNN Conv2d / ConvTranspose2d(stride 1) + BatchNorm2d(eval, folded)
// + LeakyReLU + residual / skip adds + nearest x2 upsample of TRAIN:197-350 / TEST1:170-323
// (TRAIN = train_raytracing_relighting_CelebAHQ_DSSIM_8x.py, TEST1 = test_relight_single_image.py).
//
// Second generation of conv_tc.cu's fp16-pair-split path (that file keeps the fp32 "C4" kernels the TRAIN path uses).
// What changed, and why (profiles/r01_ncu_conv_tc_full.csv: tensor pipe 11 %, four splitter warps waiting 46 % of their time
// for the TMA box, every operand rewritten in shared memory once per tile):
//   * activations are stored already split (hi, lo fp16, 4 bytes per element like fp32), so the TMA box
//     {(TW+2) px x 8 ch, 18 rows, hi|lo, KS chunks} lands as the K-major un-swizzled UMMA operand of BOTH products
//     (hi*[W1|W2] and lo*W1): the splitter warps, their shared-memory pass and the generic->async proxy fence are gone;
//     the producing layer's epilogue splits each output once;
//   * a 16x16-pixel tile is computed as two M = 128 MMAs over ONE 18x18 box (template MH = 2: 288-byte TMA runs and a
//     1.27x halo instead of 160-byte runs and 1.41x); MH = 1 keeps the 8x16 tile for small images;
//   * KS = 4 moves 32 input channels per pipeline step (half the steps of the latency-bound low-resolution layers) and the
//     ring is as deep as shared memory allows (up to 6 stages instead of 2);
//   * an output-channel range may skip the activation (act_channels), so a residual block's first conv and its shortcut
//     conv — same input — run as ONE launch with concatenated output channels; the second conv then reads the leading
//     channels as its input and the trailing ones as its residual operand (res_c8 / res_groups), both in place;
//   * every output is range-checked against the fp16 split's limit (|x| < 4094): an overflow sets a device flag that the
//     module reads at the end of the forward to re-run in 3xTF32 instead of returning inf / NaN.
//
// As before: all nine filter taps read the same shared-memory tile through descriptors whose start address is shifted by
// (ky*HALO_W + kx)*16 bytes; image borders and channel padding are TMA out-of-bounds zero fill; the accumulator chain is
// cut at every pipeline step (fresh TMEM accumulator per step, fp32 register accumulation with round-to-nearest in the
// epilogue warps: the tensor core's accumulate truncates); warp 0 = TMA producer, warp 1 = MMA issuer, 4*MH epilogue warps.
#include "gfr_common.cuh"
#include "p16.cuh"
#include "head_device.cuh"
#include "tc_common.cuh"

#include <cuda_fp16.h>
#include <math.h>
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include <type_traits>

using namespace gfr_tc;

namespace {

constexpr int TILE_H = 16;
constexpr int MAX_STAGES = 8;
#ifndef GFR_P16_MINBLOCKS
#define GFR_P16_MINBLOCKS 2     // CTAs per SM the 16-channel kernels are compiled for (register cap 65536 / (320 * n))
#endif
#ifndef GFR_P16_EARLY_LOADS
#define GFR_P16_EARLY_LOADS 1   // residual / skip operands loaded before the accumulator wait (32 registers more)
#endif
// filter geometry: GEO 0 = 3x3 / pad 1 (halo 1 px all round, 9 taps); GEO 1 = 5x1 VERTICAL / pad (2, 0) (halo 2 rows above and
// below, none sideways, 5 taps) — the stem's 5x5 convolution after its five horizontal taps have been unrolled into channels
// (gfr_stem_unroll_p16): conv5x5(img)[co] = sum_ky sum_{c' = kx*3 + c} W[co][c][ky][kx] * U[c'][y + ky - 2][x]

struct ConvP16Args {
  const __half* wpk;     // packed weights, see gfr_conv_p16_pack_weights
  const float* bias;     // [Cout]
  const __half* res;     // P16 or null: added before the activation; chunk res_c8 + c of a tensor with res_groups chunks
  const __half* post;    // P16 or null: added after the activation, nearest x2 when post_shift = 1
  __half* out;           // P16 [N][out_groups][2][H][W][8]
  int* flags;            // device int or null: bit 0 is set when an output leaves the fp16 split's range
  int N, Cin, Cout, H, W;
  int nsteps, stages;    // K steps per tile (8*KS channels each); ring depth
  int tiles_x, tiles_y, m_tiles;
  int out_groups, res_groups, res_c8, post_groups, post_shift;
  int act, act_channels; // act: 0 none, 1 LeakyReLU(0.2), 2 sigmoid — applied to output channels < act_channels
  int static_w;          // 1: the weights were not written by the preceding kernel (fetch them before griddepcontrol.wait)
  float inv_scale;       // 1 / (X_SCALE * w_scale)
  float out_scale;
  __half* pool;          // P16 [N][out_groups][2][H/2][W/2][8] or null: the 2x2 / stride 2 max pool of `out`, from the same epilogue
};

// ES = epilogue column split: every (M-half, TMEM lane quarter) is served by ES warps that take NT / ES output channels each
template <int NT, int MH, int KS, int GEO = 0, int ES = 1>
struct Cfg {
  static constexpr int TAPS = GEO == 0 ? 9 : 5;
  static constexpr int PAD_X = GEO == 0 ? 1 : 0, PAD_Y = GEO == 0 ? 1 : 2;
  static constexpr int TILE_W = 8 * MH, HALO_W = TILE_W + 2 * PAD_X, HALO_H = TILE_H + 2 * PAD_Y;
  static constexpr uint32_t PART = HALO_H * HALO_W * 16;        // one part (hi or lo) of one 8-channel chunk of the halo tile
  static constexpr uint32_t A_LBO = 2 * PART;                   // K direction: the next 8-channel chunk
  static constexpr uint32_t A_SBO = HALO_W * 16;                // M direction: the next tile row (8 pixels further in M)
  static constexpr uint32_t A_BYTES = KS * A_LBO;
  static constexpr uint32_t B_LBO = 2 * NT * 16;                // [W1 rows | W2 rows] of one 8-channel chunk
  static constexpr uint32_t B_TAP = KS * B_LBO;
  static constexpr uint32_t W_STEP = TAPS * B_TAP;
  // halo-tile offset (in pixels) of filter tap t
  __host__ __device__ static constexpr uint32_t tap_px(int t) { return GEO == 0 ? (uint32_t)((t / 3) * HALO_W + (t % 3)) : (uint32_t)(t * HALO_W); }
  static constexpr int EPI_WARPS = 4 * MH * ES, THREADS = 64 + 32 * EPI_WARPS;
  static_assert(ES == 1 || (ES == 2 && NT == 16), "the column split hands each warp one 8-channel chunk of a 16-channel tile");
  static constexpr uint32_t ACC_COLS = MH * 2 * NT;             // one accumulator buffer: per M-half [main NT | correction NT]
  static constexpr uint32_t TMEM_COLS = 2 * ACC_COLS <= 32 ? 32 : (2 * ACC_COLS <= 64 ? 64 : (2 * ACC_COLS <= 128 ? 128 : (2 * ACC_COLS <= 256 ? 256 : 512)));
  static constexpr uint32_t BAR_BYTES = 512;                    // barriers + TMEM slot (176 B), then bias*16 (NT floats) at +256
  static_assert(A_BYTES % 128 == 0 && W_STEP % 128 == 0, "TMA destinations must stay 128-byte aligned");
};

// HEAD = true: the layer is a decoder's last 3x3 convolution (conv_*_c2_1 + BN + LeakyReLU, 16 channels) and its epilogue runs
// the decoder's 1x1 tail (head_device.cuh: two 16 -> 16 layers + the output layer) on the pixel it holds, then stores the
// n_out fp32 planes: the 16-channel activation (4 B x 16 per pixel written, then read again by a second kernel) never exists.
struct HeadParams {
  gfr_head::HeadWeights wt;
  float* out;            // [N][n_out][H][W] fp32
  int n_out, act;        // act: 0 none, 2 sigmoid
  float scale;
};
struct NoHead { int unused; };

template <int NT, int MH, int KS, int GEO = 0, bool HEAD = false, int ES = 1>
__global__ void __launch_bounds__(Cfg<NT, MH, KS, GEO, ES>::THREADS, (NT * MH <= 32 && ES == 1) ? GFR_P16_MINBLOCKS : 1)
conv3x3_p16_kernel(const __grid_constant__ CUtensorMap tm_in, const ConvP16Args a,
                   const __grid_constant__ typename std::conditional<HEAD, HeadParams, NoHead>::type hp) {
  using C = Cfg<NT, MH, KS, GEO, ES>;
  static_assert(!HEAD || NT == 16, "the fused 1x1 tail works on 16 channels");
  static_assert(!HEAD || ES == 1, "the fused 1x1 tail needs the pixel's 16 channels in one thread");
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool resident = a.nsteps == 1;
  const int STAGES = a.stages;
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t slot_bytes = resident ? C::A_BYTES : C::A_BYTES + C::W_STEP;
  const uint32_t slots0 = smem0 + (resident ? C::W_STEP : 0u);
  const uint32_t bars = slots0 + STAGES * slot_bytes;
  // barrier map: full[s] (MAX_STAGES), empty[s] (MAX_STAGES), accfull[2], accempty[2], TMEM slot
  const uint32_t bar_full = bars, bar_empty = bars + 8 * MAX_STAGES, bar_accfull = bars + 16 * MAX_STAGES,
                 bar_accempty = bar_accfull + 16;
  constexpr uint32_t TMEM_SLOT = 16 * MAX_STAGES + 32;
  uint8_t* gen_bars = smem + (bars - smem0);
  float* s_bias16 = reinterpret_cast<float*>(gen_bars + 256);      // NT floats behind the barriers

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int p = 0; p < 2; ++p) {
      mbar_init(bar_accfull + 8 * p, 1);
      mbar_init(bar_accempty + 8 * p, 32 * C::EPI_WARPS);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tm_in);
  }
  if (warp == 1) tmem_alloc(bars + TMEM_SLOT, C::TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen_bars + TMEM_SLOT);

  const int n_my_tiles = (a.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_steps = n_my_tiles * a.nsteps;
  griddep_launch_dependents();        // PDL: the next layer may start its prologue (barriers, TMEM, weight fetch) now

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (elect_one_sync()) {
      const __half* wsrc = a.wpk + (size_t)blockIdx.y * a.nsteps * (C::W_STEP / 2);
      const int n_pre = n_steps < STAGES ? n_steps : STAGES;
      for (int g = 0; g < n_pre; ++g) {
        const bool load_w = !resident || g == 0;
        mbar_expect_tx(bar_full + 8 * g, C::A_BYTES + (load_w ? C::W_STEP : 0u));
        if (load_w && a.static_w)
          bulk_load(resident ? smem0 : slots0 + g * slot_bytes + C::A_BYTES, wsrc + (size_t)(g % a.nsteps) * (C::W_STEP / 2), C::W_STEP,
                    bar_full + 8 * g);
      }
      griddep_wait();
      int g = 0, s = 0, ph = 0;
      for (int mt = blockIdx.x; mt < a.m_tiles; mt += gridDim.x) {
        const int tx = mt % a.tiles_x, t2 = mt / a.tiles_x;
        const int ty = t2 % a.tiles_y, n = t2 / a.tiles_y;
        for (int st = 0; st < a.nsteps; ++st, ++g) {
          const uint32_t slot = slots0 + s * slot_bytes;
          const bool load_w = !resident || g == 0;
          if (g >= n_pre) {
            mbar_wait(bar_empty + 8 * s, ph ^ 1);
            mbar_expect_tx(bar_full + 8 * s, C::A_BYTES + (load_w ? C::W_STEP : 0u));
          }
          tma_load_5d(slot, &tm_in, bar_full + 8 * s, (tx * C::TILE_W - C::PAD_X) * 8, ty * TILE_H - C::PAD_Y, 0, st * KS, n);
          if (load_w && (g >= n_pre || !a.static_w))
            bulk_load(resident ? smem0 : slot + C::A_BYTES, wsrc + (size_t)st * (C::W_STEP / 2), C::W_STEP, bar_full + 8 * s);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // The whole warp walks the pipeline (uniform waits), ONE elected lane issues the step's MMAs and commits.
    constexpr uint32_t IDESC_2N = umma_idesc_f16(128, 2 * NT), IDESC_N = umma_idesc_f16(128, NT);
    int s = 0, ph = 0;
    for (int g = 0; g < n_steps; ++g) {
      const int p = g & 1;
      mbar_wait(bar_full + 8 * s, ph);
      mbar_wait(bar_accempty + 8 * p, ((g >> 1) & 1) ^ 1);
      tc_fence_after_sync();
      if (elect_one_sync()) {
        const uint32_t slot = slots0 + s * slot_bytes;
        const uint32_t wbase = resident ? smem0 : slot + C::A_BYTES;
        const uint64_t dB0 = umma_desc_kmajor_noswz(wbase, C::B_LBO, 128u);
        const uint64_t dA_hi = umma_desc_kmajor_noswz(slot, C::A_LBO, C::A_SBO);
        const uint64_t dA_lo = umma_desc_kmajor_noswz(slot + C::PART, C::A_LBO, C::A_SBO);
#pragma unroll
        for (int h = 0; h < MH; ++h) {
          const uint32_t d_main = tmem + (uint32_t)(p * C::ACC_COLS + h * 2 * NT), d_corr = d_main + NT;
#pragma unroll
          for (int tap = 0; tap < C::TAPS; ++tap) {
#pragma unroll
            for (int j = 0; j < KS / 2; ++j) {          // one K = 16 MMA covers two 8-channel chunks
              const uint32_t ao = (C::tap_px(tap) * 16u + (uint32_t)h * 128u + (uint32_t)j * 2u * C::A_LBO) >> 4;
              const uint32_t bo = ((uint32_t)tap * C::B_TAP + (uint32_t)j * 2u * C::B_LBO) >> 4;
              const uint32_t acc = (tap == 0 && j == 0) ? 0u : 1u;
              umma_f16(d_main, dA_hi + ao, dB0 + bo, IDESC_2N, acc);     // main += hi*W1 ; corr += hi*W2
              umma_f16(d_corr, dA_lo + ao, dB0 + bo, IDESC_N, 1u);       // corr += lo*W1
            }
          }
        }
        umma_commit(bar_empty + 8 * s);       // the slot may be refilled once these MMAs have read it
        umma_commit(bar_accfull + 8 * p);     // and the accumulators of this step are complete
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else {
    // =============================== epilogue (thread = pixel = TMEM lane of its M-half) ===============================
    // Everything is computed in the x16 domain of the stored pair (16 v = hi + lo): bias and 1/(x_scale*w_scale) are
    // pre-multiplied, LeakyReLU is homogeneous, residual / post operands are added as hi + lo without rescaling.  The
    // instruction count per tile matters: the epilogue warps share their schedulers with the MMA-issuing warp.
    // q: the TMEM lane quarter this warp may access (warp % 4, a hardware rule); h: its M-half; cs: its share of the output channels.
    // ES = 2 (A/B, see launch_p16): 16 epilogue warps, four per scheduler, 8 channels each.  The epilogue warps execute 95 % of the
    // kernel's instructions (~570 per tile and warp, ncu source page) and wait for accumulators only 7 % of their time, yet doubling
    // them buys 7 % on the layer alone and costs more than that in the overlapped forward.
    const int e = warp - 2, h = (e >> 2) % MH, cs = e / (4 * MH), q = warp & 3;
    constexpr int NC = NT / ES;                             // output channels (TMEM columns of main / correction) of this warp
    const int col0 = cs * NC, ch0 = col0 >> 3;
    const int m = q * 32 + lane;
    const int n0 = blockIdx.y * NT;
    const int pH = a.H >> a.post_shift, pW = a.W >> a.post_shift;
    const size_t plane = (size_t)a.H * a.W * 8;            // halfs between the hi and the lo unit of a pixel
    const size_t pplane = (size_t)pH * pW * 8;
    const float inv16 = a.inv_scale * gfr_p16::X_SCALE;
    const float lim16 = gfr_p16::X_LIMIT * gfr_p16::X_SCALE;
    const int n_chunks = min(NT / 8, a.out_groups - (n0 >> 3));
    bool overflow = false;
    for (int c = tid - 64; c < NT; c += 32 * C::EPI_WARPS)                       // bias * 16 (0 for padding channels), once per CTA
      s_bias16[c] = n0 + c < a.Cout ? __ldg(a.bias + n0 + c) * gfr_p16::X_SCALE : 0.f;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * C::EPI_WARPS) : "memory");
    griddep_wait();                                        // residual / skip operands and the output buffer belong to earlier kernels
    int g = 0;
    for (int mt = blockIdx.x; mt < a.m_tiles; mt += gridDim.x) {
      const int tx = mt % a.tiles_x, t2 = mt / a.tiles_x;
      const int ty = t2 % a.tiles_y, n = t2 / a.tiles_y;
      const int y = ty * TILE_H + (m >> 3), x = tx * C::TILE_W + h * 8 + (m & 7);
      const bool ok = y < a.H && x < a.W;
      const __half* res_p = a.res ? a.res + gfr_p16::unit_offset(n, a.res_groups, a.res_c8 + (n0 >> 3), a.H, a.W, y, x) : nullptr;
      const __half* post_p = a.post ? a.post + gfr_p16::unit_offset(n, a.post_groups, n0 >> 3, pH, pW, y >> a.post_shift, x >> a.post_shift) : nullptr;
      __half* out_p = a.out + gfr_p16::unit_offset(n, a.out_groups, n0 >> 3, a.H, a.W, y, x);
      // residual / skip operands.  NT = 16 (the 128^2 / 256^2 layers, where these reads were 12 of 33 us): the loads are issued
      // NOW, before the wait for the accumulators, so their L2 / DRAM latency overlaps the MMAs of this tile; wider layers
      // (32 accumulator registers more per thread) keep the L2 prefetch and load after the MMAs.
      constexpr bool EARLY = (NT == 16) && !HEAD && GFR_P16_EARLY_LOADS;
      uint4 e_res[EARLY ? NC / 4 : 1], e_post[EARLY ? NC / 4 : 1];
      if (EARLY) {
#pragma unroll
        for (int l = 0; l < NC / 8; ++l) {
          const int c = ch0 + l;
          const bool lc = ok && c < n_chunks;
          e_res[2 * l] = e_res[2 * l + 1] = e_post[2 * l] = e_post[2 * l + 1] = make_uint4(0u, 0u, 0u, 0u);
          if (res_p && lc) {
            e_res[2 * l] = __ldg(reinterpret_cast<const uint4*>(res_p + 2 * c * plane));
            e_res[2 * l + 1] = __ldg(reinterpret_cast<const uint4*>(res_p + (2 * c + 1) * plane));
          }
          if (post_p && lc) {
            e_post[2 * l] = __ldg(reinterpret_cast<const uint4*>(post_p + 2 * c * pplane));
            e_post[2 * l + 1] = __ldg(reinterpret_cast<const uint4*>(post_p + (2 * c + 1) * pplane));
          }
        }
      } else if (ok && (res_p || post_p)) {
        for (int c = ch0; c < min(n_chunks, ch0 + NC / 8); ++c) {
          if (res_p) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(res_p + 2 * c * plane));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(res_p + (2 * c + 1) * plane));
          }
          if (post_p) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(post_p + 2 * c * pplane));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(post_p + (2 * c + 1) * pplane));
          }
        }
      }
      float sum[NC];
      for (int st = 0; st < a.nsteps; ++st, ++g) {
        const int p = g & 1;
        mbar_wait(bar_accfull + 8 * p, (g >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t t_main = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(p * C::ACC_COLS + h * 2 * NT + col0);
        if constexpr (NC == 8) {
          uint32_t rm[8], rc[8];
          tmem_ld8(t_main, rm);
          tmem_ld8(t_main + NT, rc);
          tmem_ld_wait();
          if (st == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) sum[k] = __uint_as_float(rm[k]) + __uint_as_float(rc[k]);
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) sum[k] += __uint_as_float(rm[k]) + __uint_as_float(rc[k]);
          }
        } else {
#pragma unroll
          for (int c0 = 0; c0 < NC; c0 += 16) {
            uint32_t rm[16], rc[16];
            tmem_ld16(t_main + c0, rm);
            tmem_ld16(t_main + NT + c0, rc);
            tmem_ld_wait();
            if (st == 0) {
#pragma unroll
              for (int k = 0; k < 16; ++k) sum[c0 + k] = __uint_as_float(rm[k]) + __uint_as_float(rc[k]);
            } else {
#pragma unroll
              for (int k = 0; k < 16; ++k) sum[c0 + k] += __uint_as_float(rm[k]) + __uint_as_float(rc[k]);
            }
          }
        }
        tc_fence_before_sync();
        mbar_arrive(bar_accempty + 8 * p);
      }
      if constexpr (HEAD) {
        // conv + bias + LeakyReLU of the 16 channels (no residual / skip operand on this layer: host-checked), back from the
        // x16 domain, then the 1x1 tail; a warp stores 4 rows x 8 pixels = four full 32-byte sectors per output plane
        float xv[16], o3[3];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float v = fmaf(sum[k], inv16, s_bias16[k]);
          xv[k] = fmaxf(v, 0.2f * v) * gfr_p16::X_INV;
        }
        gfr_head::apply(hp.wt, xv, hp.n_out, hp.act, hp.scale, o3);
        if (ok) {
          float* o = hp.out + ((size_t)n * hp.n_out * a.H + y) * a.W + x;
#pragma unroll
          for (int k = 0; k < 3; ++k)
            if (k < hp.n_out) o[(size_t)k * a.H * a.W] = o3[k];
        }
        continue;
      }
      // every lane walks the chunks (warp-uniform control flow: the fused pool shuffles); loads / stores are predicated by `ok`
      const size_t qplane = (size_t)(a.H >> 1) * (a.W >> 1) * 8;
      __half* pool_p = a.pool ? a.pool + gfr_p16::unit_offset(n, a.out_groups, n0 >> 3, a.H >> 1, a.W >> 1, y >> 1, x >> 1) : nullptr;
      const bool pool_writer = ok && !(lane & 1) && !(lane & 8);
#pragma unroll
      for (int l = 0; l < NC / 8; ++l) {
        const int c = ch0 + l;                                   // chunk of the N tile; l: this warp's
        if (c >= n_chunks) break;
        float v[8];
        const float4 b0 = *reinterpret_cast<const float4*>(s_bias16 + 8 * c), b1 = *reinterpret_cast<const float4*>(s_bias16 + 8 * c + 4);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaf(sum[8 * l + k], inv16, bb[k]);
        if (res_p && ok) {
          float rv[8];
          if (EARLY) gfr_p16::join8_x16(e_res[(2 * l) & 3], e_res[(2 * l + 1) & 3], rv);
          else gfr_p16::join8_x16(__ldg(reinterpret_cast<const uint4*>(res_p + 2 * c * plane)),
                                  __ldg(reinterpret_cast<const uint4*>(res_p + (2 * c + 1) * plane)), rv);
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] += rv[k];
        }
        const bool do_act = n0 + 8 * c < a.act_channels;       // act_channels is a multiple of 8 (checked by the host)
        if (a.act == 1 && do_act) {
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.2f * v[k]);
        } else if (a.act == 2 && do_act) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            v[k] = n0 + 8 * c + k < a.Cout ? gfr_p16::X_SCALE / (1.0f + expf(-v[k] * gfr_p16::X_INV)) : 0.f;
        }
        if (post_p && ok) {
          float pv[8];
          if (EARLY) gfr_p16::join8_x16(e_post[(2 * l) & 3], e_post[(2 * l + 1) & 3], pv);
          else gfr_p16::join8_x16(__ldg(reinterpret_cast<const uint4*>(post_p + 2 * c * pplane)),
                                  __ldg(reinterpret_cast<const uint4*>(post_p + (2 * c + 1) * pplane)), pv);
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] += pv[k];
        }
        if (a.out_scale != 1.0f) {
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] *= a.out_scale;
        }
        if (ok) {
#pragma unroll
          for (int k = 0; k < 8; ++k) overflow = overflow || !(fabsf(v[k]) < lim16);     // also catches NaN
          uint4 hi, lo;
          gfr_p16::split8_x16(v, hi, lo);
          *reinterpret_cast<uint4*>(out_p + 2 * c * plane) = hi;
          *reinterpret_cast<uint4*>(out_p + (2 * c + 1) * plane) = lo;
        }
        if (pool_p != nullptr) {
          // 2x2 / stride 2 max pool of the finished values: lane = (tile row & 3) * 8 + px, so the quad partners of a pixel are
          // lanes ^1 (x) and ^8 (y) of its own warp; the even/even lane writes.  H, W even (host-checked) and tile origins even,
          // so a quad never straddles tiles or the image edge.
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            v[k] = fmaxf(v[k], __shfl_xor_sync(0xffffffffu, v[k], 1));
            v[k] = fmaxf(v[k], __shfl_xor_sync(0xffffffffu, v[k], 8));
          }
          if (pool_writer) {
            uint4 hi, lo;
            gfr_p16::split8_x16(v, hi, lo);
            *reinterpret_cast<uint4*>(pool_p + 2 * c * qplane) = hi;
            *reinterpret_cast<uint4*>(pool_p + (2 * c + 1) * qplane) = lo;
          }
        }
      }
    }
    if (a.flags != nullptr && __any_sync(0xffffffffu, overflow) && lane == 0) atomicOr(a.flags, 1);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, C::TMEM_COLS);
}

// ---- TMA descriptor for a P16 activation tensor ---------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

int sm_count() {
  static int n = 0;
  static std::once_flag once;
  std::call_once(once, [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n = v;
    else n = 148;
  });
  return n;
}

// tensor [N][groups][2][H][W][8] fp16 described as 5-D {W*8, H, 2, C8, N} (pixel and channel slot merged: one contiguous
// 16*bw-byte run per box row); box {8*bw, 18, 2, KS, 1}
int make_p16_map(CUtensorMap* tm, const void* base, int N, int C8, int groups, int H, int W, int bw, int bh, int ks) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return GFR_E_UNSUPPORTED;
  const cuuint64_t dims[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, 2, (cuuint64_t)C8, (cuuint64_t)N};
  const cuuint64_t hw16 = (cuuint64_t)H * W * 16;
  const cuuint64_t strides[4] = {(cuuint64_t)W * 16, hw16, 2 * hw16, (cuuint64_t)groups * 2 * hw16};
  const cuuint32_t box[5] = {(cuuint32_t)bw * 8, (cuuint32_t)bh, 2, (cuuint32_t)ks, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? GFR_OK : GFR_E_ARG;
}

int g_p16_grid_occ = 0;      // 0 = default (environment, else 1), 1 | 2 = persistent CTAs per SM (gfr_conv_p16_config)

template <int NT, int MH, int KS, int GEO = 0, bool HEAD = false, int ES = 1>
int launch_p16(const CUtensorMap& tm, ConvP16Args a, cudaStream_t s, const HeadParams* head = nullptr) {
  if constexpr (ES == 1 && NT == 16 && !HEAD) {
    // A/B (GFR_P16_EPI_SPLIT=1, off): the 16-channel layers with 16 epilogue warps of 8 channels instead of 8 of 16.  Measured: the
    // 256^2 layer alone 28.7 -> 26.7 us, but the forward 17.0k vs 19.1k faces/s and 0.535 vs 0.501 ms latency — the 576-thread CTAs
    // leave less room beside them and start slower on the latency-bound small layers.
    static const bool split = [] { const char* e = getenv("GFR_P16_EPI_SPLIT"); return e != nullptr && e[0] == '1'; }();
    if (split) return launch_p16<NT, MH, KS, GEO, HEAD, 2>(tm, a, s, head);
  }
  using C = Cfg<NT, MH, KS, GEO, ES>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv3x3_p16_kernel<NT, MH, KS, GEO, HEAD, ES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (attr_err != cudaSuccess) return (int)attr_err;
  const bool resident = a.nsteps == 1;
  const uint32_t fixed = (resident ? C::W_STEP : 0u) + C::BAR_BYTES;
  const uint32_t per_stage = C::A_BYTES + (resident ? 0u : C::W_STEP);
  // ring depth: two CTAs per SM where the stage size allows (the epilogue of one CTA overlaps the MMAs of the other); layers
  // whose stage is too large for that run one CTA per SM.  The ring itself is kept at TWO stages (GFR_P16_STAGES raises the
  // cap, up to what shared memory holds): measured on the forward (profiles/r02_summary.md), deeper rings change neither
  // the latency of one forward nor any layer's duration, but the shared memory they pin keeps the CTAs of another runner
  // lane's layer off the SM — 2 stages: 15.6k faces/s, 6 stages: 15.2k.
  static const int max_stages = [] { const char* e = getenv("GFR_P16_STAGES"); const int v = e ? atoi(e) : 0; return v >= 2 && v <= MAX_STAGES ? v : 2; }();
  const bool two_ok = (NT * MH <= 32);
  int stages = two_ok ? (int)((110u * 1024u - fixed) / per_stage) : 0;
  int occ = 2;
  if (stages < 3) {
    stages = (int)((220u * 1024u - fixed) / per_stage);
    occ = 1;
  }
  if (stages > max_stages) stages = max_stages;
  if (stages < 2) return GFR_E_UNSUPPORTED;
  a.stages = stages;
  const uint32_t bytes = fixed + (uint32_t)stages * per_stage;
  const int n_tiles = gfr_ceil_div(a.Cout, NT);
  // Persistent grid: ONE CTA per SM even where two fit.  The second slot then goes to whatever else is running — the other
  // decoder's layer, another runner lane's layer or its ray march — so the ramp (first TMA box, ~2 us with an idle tensor
  // pipe) and the tail of one kernel overlap another kernel's MMAs; two CTAs of the SAME kernel per SM ramp and drain
  // together.  Measured on the forward (3 lanes): 18.7k vs 17.7k faces/s, e2e 19.2k vs 18.6k; the latency of ONE forward is
  // 3 % worse (0.510 vs 0.493 ms) — GFR_P16_GRID_OCC=2 restores the latency-optimal grid; GFR_P16_GRID_CTAS caps the grid.
  static const int env_grid_occ = [] { const char* e = getenv("GFR_P16_GRID_OCC"); const int v = e ? atoi(e) : 0; return v == 2 ? 2 : 1; }();
  const int grid_occ = g_p16_grid_occ > 0 ? g_p16_grid_occ : env_grid_occ;
  static const int grid_cap = [] { const char* e = getenv("GFR_P16_GRID_CTAS"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 0; }();
  int gx = (sm_count() * (occ < grid_occ ? occ : grid_occ)) / n_tiles;
  if (grid_cap > 0 && gx * n_tiles > grid_cap) gx = grid_cap / n_tiles > 0 ? grid_cap / n_tiles : 1;
  if (gx < 1) gx = 1;
  if (gx > a.m_tiles) gx = a.m_tiles;
  static const bool no_pdl = [] { const char* e = getenv("GFR_PDL"); return e != nullptr && e[0] == '0'; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gx, n_tiles);
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  cudaError_t e;
  if constexpr (HEAD) e = cudaLaunchKernelEx(&cfg, conv3x3_p16_kernel<NT, MH, KS, GEO, true, ES>, tm, a, *head);
  else e = cudaLaunchKernelEx(&cfg, conv3x3_p16_kernel<NT, MH, KS, GEO, false, ES>, tm, a, NoHead{0});
  return e == cudaSuccess ? gfr_launch_status() : (int)e;
}

// ---- layout helpers ------------------------------------------------------------------------------------
__global__ void nchw_to_p16_kernel(const float* __restrict__ in, __half* __restrict__ out, int C, int HW, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over [N][C8][HW]
  if (i >= total) return;
  const int C8 = (C + 7) >> 3;
  const int p = (int)(i % HW);
  const long long t = i / HW;
  const int c8 = (int)(t % C8);
  const long long n = t / C8;
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = c8 * 8 + k;
    v[k] = c < C ? __ldg(in + (n * C + c) * HW + p) : 0.f;
  }
  uint4 hi, lo;
  gfr_p16::split8(v, hi, lo);
  __half* o = out + ((n * C8 + c8) * 2 * (long long)HW + p) * 8;
  *reinterpret_cast<uint4*>(o) = hi;
  *reinterpret_cast<uint4*>(o + (size_t)HW * 8) = lo;
}

__global__ void p16_to_nchw_kernel(const __half* __restrict__ in, float* __restrict__ out, int C, int groups, int HW, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over [N][C8][HW]
  if (i >= total) return;
  const int C8 = (C + 7) >> 3;
  const int p = (int)(i % HW);
  const long long t = i / HW;
  const int c8 = (int)(t % C8);
  const long long n = t / C8;
  const __half* s = in + ((n * groups + c8) * 2 * (long long)HW + p) * 8;
  float v[8];
  gfr_p16::join8(__ldg(reinterpret_cast<const uint4*>(s)), __ldg(reinterpret_cast<const uint4*>(s + (size_t)HW * 8)), v);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = c8 * 8 + k;
    if (c < C) out[(n * C + c) * HW + p] = v[k];
  }
}

// 2x2/2 max pool: compares the joined fp32 values and re-splits the winner
__global__ void maxpool2_p16_kernel(const __half* __restrict__ in, __half* __restrict__ out, long long n_out, int Ho, int Wo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over [N*C8][Ho][Wo]
  if (i >= n_out) return;
  const int x = (int)(i % Wo);
  const long long t = i / Wo;
  const int y = (int)(t % Ho);
  const long long nc = t / Ho;
  const size_t iplane = (size_t)4 * Ho * Wo * 8, oplane = (size_t)Ho * Wo * 8;
  const __half* p = in + nc * 2 * iplane + ((size_t)(2 * y) * (2 * Wo) + 2 * x) * 8;
  float m[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half* q = p + ((size_t)(j >> 1) * (2 * Wo) + (j & 1)) * 8;
    float v[8];
    gfr_p16::join8(__ldg(reinterpret_cast<const uint4*>(q)), __ldg(reinterpret_cast<const uint4*>(q + iplane)), v);
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = j == 0 ? v[k] : fmaxf(m[k], v[k]);
  }
  uint4 hi, lo;
  gfr_p16::split8(m, hi, lo);
  __half* o = out + nc * 2 * oplane + ((size_t)y * Wo + x) * 8;
  *reinterpret_cast<uint4*>(o) = hi;
  *reinterpret_cast<uint4*>(o + oplane) = lo;
}

// U[n][y][x][kx*3 + c] = img[n][y][x + kx - 2][c] (zero outside), channel 15 = 0: the five horizontal taps of the stem's 5x5
// convolution unrolled into 15 (+1) channels of a P16 tensor, so the 5x5 layer becomes a 5x1 vertical-tap layer on the tensor cores
__global__ void __launch_bounds__(256) stem_unroll_p16_kernel(const float* __restrict__ img, __half* __restrict__ out, int H, int W, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over [N][H][W]
  if (i >= total) return;
  const int x = (int)(i % W);
  const long long t = i / W;
  const int y = (int)(t % H);
  const long long n = t / H;
  const float* row = img + (n * H + y) * (long long)W * 3;
  float v[16];
#pragma unroll
  for (int kx = 0; kx < 5; ++kx) {
    const int xx = x + kx - 2;
    const bool in = xx >= 0 && xx < W;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[kx * 3 + c] = in ? __ldg(row + (size_t)xx * 3 + c) : 0.f;
  }
  v[15] = 0.f;
  const size_t plane = (size_t)H * W * 8;
  __half* o = out + ((size_t)n * 2 * 2 * H + y) * (size_t)W * 8 + (size_t)x * 8;      // [n][chunk][part][y][x][8]
#pragma unroll
  for (int c8 = 0; c8 < 2; ++c8) {
    float w8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) w8[k] = v[c8 * 8 + k];
    uint4 hi, lo;
    gfr_p16::split8(w8, hi, lo);
    *reinterpret_cast<uint4*>(o + (size_t)(2 * c8) * plane) = hi;
    *reinterpret_cast<uint4*>(o + (size_t)(2 * c8 + 1) * plane) = lo;
  }
}

}  // namespace

extern "C" int gfr_conv_p16_config(int grid_ctas_per_sm) {
  if (grid_ctas_per_sm < 0 || grid_ctas_per_sm > 2) return GFR_E_ARG;
  g_p16_grid_occ = grid_ctas_per_sm;
  return GFR_OK;
}

extern "C" int gfr_stem_unroll_p16(const float* img, void* out, int N, int H, int W, void* stream) {
  GFR_RETURN_IF_NULL(img); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  const long long total = (long long)N * H * W;
  stem_unroll_p16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(img, reinterpret_cast<__half*>(out), H, W, total);
  return gfr_launch_status();
}

static long long p16_pack_size(int Cin, int Cout, int NT, int KS, int taps) {
  if (Cin <= 0 || Cout <= 0 || (NT != 16 && NT != 32) || (KS != 2 && KS != 4) || (taps != 9 && taps != 5)) return GFR_E_ARG;
  return (long long)gfr_ceil_div(Cout, NT) * gfr_ceil_div(Cin, 8 * KS) * taps * KS * 2 * NT * 8;     // in halfs
}

extern "C" long long gfr_conv_p16_pack_size(int Cin, int Cout, int NT, int KS) { return p16_pack_size(Cin, Cout, NT, KS, 9); }
extern "C" long long gfr_conv_p16_pack_size_taps(int Cin, int Cout, int NT, int KS, int taps) { return p16_pack_size(Cin, Cout, NT, KS, taps); }

extern "C" int gfr_conv_p16_pack_weights_taps(const float* w_host, int Cin, int Cout, int NT, int KS, int taps, float w_scale, void* packed_host);

extern "C" int gfr_conv_p16_pack_weights(const float* w_host, int Cin, int Cout, int NT, int KS, float w_scale, void* packed_host) {
  return gfr_conv_p16_pack_weights_taps(w_host, Cin, Cout, NT, KS, 9, w_scale, packed_host);
}

// w_host [Cout][Cin][taps]
extern "C" int gfr_conv_p16_pack_weights_taps(const float* w_host, int Cin, int Cout, int NT, int KS, int taps, float w_scale, void* packed_host) {
  GFR_RETURN_IF_NULL(w_host); GFR_RETURN_IF_NULL(packed_host);
  if (Cin <= 0 || Cout <= 0 || (NT != 16 && NT != 32) || (KS != 2 && KS != 4) || (taps != 9 && taps != 5) || !(w_scale > 0.f)) return GFR_E_ARG;
  const int n_tiles = gfr_ceil_div(Cout, NT), nsteps = gfr_ceil_div(Cin, 8 * KS);
  __half* out = reinterpret_cast<__half*>(packed_host);
  size_t o = 0;
  for (int nt = 0; nt < n_tiles; ++nt)
    for (int st = 0; st < nsteps; ++st)
      for (int tap = 0; tap < taps; ++tap)
        for (int kc = 0; kc < KS; ++kc)
          for (int part = 0; part < 2; ++part)
            for (int n = 0; n < NT; ++n)
              for (int e = 0; e < 8; ++e, ++o) {
                const int co = nt * NT + n, ci = (st * KS + kc) * 8 + e;
                float v = 0.f;
                if (co < Cout && ci < Cin) {
                  const float w = w_host[((size_t)co * Cin + ci) * taps + tap] * w_scale;
                  if (!(fabsf(w) < 65000.f)) return GFR_E_ARG;            // w_scale too large for fp16
                  const float w1 = __half2float(__float2half_rn(w));
                  v = part == 0 ? w1 : w - w1;
                }
                out[o] = __float2half_rn(v);
              }
  return GFR_OK;
}

static int conv_p16_launch(const void* in, const void* w_packed, const float* bias, const void* res, int res_c8,
                           int res_groups, const void* post, int post_groups, void* out, int out_groups, void* pool, int* flags,
                           int N, int Cin, int in_groups, int Cout, int H, int W, int NT, int MH, int KS, int geo, int post_shift,
                           int act, int act_channels, float out_scale, float w_scale, int weights_static, void* stream,
                           const HeadParams* head = nullptr);

extern "C" int gfr_conv3x3_p16_fwd(const void* in, const void* w_packed, const float* bias, const void* res, int res_c8,
                                   int res_groups, const void* post, int post_groups, void* out, int out_groups, int* flags,
                                   int N, int Cin, int in_groups, int Cout, int H, int W, int NT, int MH, int KS, int post_shift,
                                   int act, int act_channels, float out_scale, float w_scale, int weights_static, void* stream) {
  return conv_p16_launch(in, w_packed, bias, res, res_c8, res_groups, post, post_groups, out, out_groups, nullptr, flags, N, Cin, in_groups,
                         Cout, H, W, NT, MH, KS, 0, post_shift, act, act_channels, out_scale, w_scale, weights_static, stream);
}

extern "C" int gfr_conv_p16_fwd_ex(const void* in, const void* w_packed, const float* bias, const void* res, int res_c8,
                                   int res_groups, const void* post, int post_groups, void* out, int out_groups, void* pool_out, int* flags,
                                   int N, int Cin, int in_groups, int Cout, int H, int W, int NT, int MH, int KS, int geometry,
                                   int post_shift, int act, int act_channels, float out_scale, float w_scale, int weights_static,
                                   void* stream) {
  return conv_p16_launch(in, w_packed, bias, res, res_c8, res_groups, post, post_groups, out, out_groups, pool_out, flags, N, Cin, in_groups,
                         Cout, H, W, NT, MH, KS, geometry, post_shift, act, act_channels, out_scale, w_scale, weights_static, stream);
}

static int conv_p16_launch(const void* in, const void* w_packed, const float* bias, const void* res, int res_c8,
                           int res_groups, const void* post, int post_groups, void* out, int out_groups, void* pool, int* flags,
                           int N, int Cin, int in_groups, int Cout, int H, int W, int NT, int MH, int KS, int geo, int post_shift,
                           int act, int act_channels, float out_scale, float w_scale, int weights_static, void* stream,
                           const HeadParams* head) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(w_packed); GFR_RETURN_IF_NULL(bias);
  if (head == nullptr) GFR_RETURN_IF_NULL(out);
  if (head != nullptr && (Cout != 16 || NT != 16 || KS != 2 || geo != 0 || res || post || pool || act != 1 || out_scale != 1.0f || head->out == nullptr))
    return GFR_E_ARG;
  if (geo < 0 || geo > 1) return GFR_E_ARG;
  if (pool != nullptr && ((H | W) & 1)) return GFR_E_SHAPE;
  if (N <= 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  if (post_shift < 0 || post_shift > 1 || act < 0 || act > 2 || !(w_scale > 0.f)) return GFR_E_ARG;
  if (post && post_shift && ((H | W) & 1)) return GFR_E_SHAPE;
  const int C8in = (Cin + 7) / 8, C8out = (Cout + 7) / 8;
  if (in_groups == 0) in_groups = C8in;
  if (out_groups == 0) out_groups = C8out;
  if (in_groups < C8in || out_groups < C8out) return GFR_E_ARG;
  if (res && (res_groups < res_c8 + C8out || res_c8 < 0)) return GFR_E_ARG;
  if (post && post_groups < C8out) return GFR_E_ARG;
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(w_packed) | reinterpret_cast<uintptr_t>(out) |
       reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(post)) & 15)
    return GFR_E_ARG;
  ConvP16Args a;
  a.wpk = reinterpret_cast<const __half*>(w_packed); a.bias = bias;
  a.res = reinterpret_cast<const __half*>(res); a.post = reinterpret_cast<const __half*>(post);
  a.out = reinterpret_cast<__half*>(out); a.flags = flags;
  a.N = N; a.Cin = Cin; a.Cout = Cout; a.H = H; a.W = W;
  a.nsteps = gfr_ceil_div(Cin, 8 * KS); a.stages = 0;
  a.tiles_x = gfr_ceil_div(W, 8 * MH); a.tiles_y = gfr_ceil_div(H, TILE_H);
  a.m_tiles = N * a.tiles_x * a.tiles_y;
  a.out_groups = out_groups; a.res_groups = res_groups; a.res_c8 = res_c8; a.post_groups = post_groups; a.post_shift = post_shift;
  if (act_channels > 0 && act_channels < Cout && (act_channels & 7)) return GFR_E_ARG;      // the activation boundary is a chunk boundary
  a.act = act; a.act_channels = act_channels <= 0 ? 8 * out_groups : act_channels;
  a.static_w = weights_static ? 1 : 0;
  a.inv_scale = 1.0f / (gfr_p16::X_SCALE * w_scale); a.out_scale = out_scale;
  a.pool = reinterpret_cast<__half*>(pool);
  CUtensorMap tm;
  const int rc = geo == 0 ? make_p16_map(&tm, in, N, C8in, in_groups, H, W, 8 * MH + 2, TILE_H + 2, KS)
                          : make_p16_map(&tm, in, N, C8in, in_groups, H, W, 8 * MH, TILE_H + 4, KS);
  if (rc != GFR_OK) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (geo == 1) {
    if (NT == 16 && MH == 2 && KS == 2) return launch_p16<16, 2, 2, 1>(tm, a, s);
    return GFR_E_ARG;
  }
  if (head != nullptr) {
    if (MH == 2) return launch_p16<16, 2, 2, 0, true>(tm, a, s, head);
    if (MH == 1) return launch_p16<16, 1, 2, 0, true>(tm, a, s, head);
    return GFR_E_ARG;
  }
#define GFR_P16_CASE(nt, mh, ks) if (NT == nt && MH == mh && KS == ks) return launch_p16<nt, mh, ks>(tm, a, s)
  GFR_P16_CASE(16, 1, 2); GFR_P16_CASE(16, 2, 2); GFR_P16_CASE(16, 1, 4); GFR_P16_CASE(16, 2, 4);
  GFR_P16_CASE(32, 1, 2); GFR_P16_CASE(32, 2, 2); GFR_P16_CASE(32, 1, 4); GFR_P16_CASE(32, 2, 4);
#undef GFR_P16_CASE
  return GFR_E_ARG;
}

// conv_*_c2_1 (3x3, Cin -> 16, + BN + LeakyReLU) with the decoder's 1x1 tail in its epilogue -> out [N][n_out][H][W] fp32
extern "C" int gfr_conv3x3_p16_head_fwd(const void* in, const void* w_packed, const float* bias, int N, int Cin, int in_groups, int H, int W,
                                        int MH, float w_scale, int weights_static, const float* w2_host, const float* b2_host,
                                        const float* w3_host, const float* b3_host, const float* wo_host, const float* bo_host,
                                        float* out, int n_out, int act, float out_scale, void* stream) {
  GFR_RETURN_IF_NULL(w2_host); GFR_RETURN_IF_NULL(b2_host); GFR_RETURN_IF_NULL(w3_host); GFR_RETURN_IF_NULL(b3_host);
  GFR_RETURN_IF_NULL(wo_host); GFR_RETURN_IF_NULL(bo_host); GFR_RETURN_IF_NULL(out);
  if (n_out < 1 || n_out > 3 || (act != 0 && act != 2)) return GFR_E_ARG;
  HeadParams hp;
  gfr_head::fill(hp.wt, w2_host, b2_host, w3_host, b3_host, wo_host, bo_host, n_out);
  hp.out = out; hp.n_out = n_out; hp.act = act; hp.scale = out_scale;
  return conv_p16_launch(in, w_packed, bias, nullptr, 0, 0, nullptr, 0, nullptr, 2, nullptr, nullptr, N, Cin, in_groups, 16, H, W, 16, MH, 2, 0, 0,
                         1, 0, 1.0f, w_scale, weights_static, stream, &hp);
}

extern "C" int gfr_nchw_to_p16(const float* in, void* out, int N, int C, int H, int W, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  const long long total = (long long)N * ((C + 7) / 8) * H * W;
  nchw_to_p16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, reinterpret_cast<__half*>(out), C, H * W, total);
  return gfr_launch_status();
}

extern "C" int gfr_p16_to_nchw(const void* in, float* out, int N, int C, int groups, int H, int W, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return GFR_E_SHAPE;
  if (groups == 0) groups = (C + 7) / 8;
  if (groups < (C + 7) / 8) return GFR_E_ARG;
  const long long total = (long long)N * ((C + 7) / 8) * H * W;
  p16_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __half*>(in), out, C, groups,
                                                                                      H * W, total);
  return gfr_launch_status();
}

extern "C" int gfr_maxpool2_p16_fwd(const void* in, void* out, int NC8, int Ho, int Wo, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(out);
  if (NC8 <= 0 || Ho <= 0 || Wo <= 0) return GFR_E_SHAPE;
  const long long n = (long long)NC8 * Ho * Wo;
  maxpool2_p16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __half*>(in),
                                                                                    reinterpret_cast<__half*>(out), n, Ho, Wo);
  return gfr_launch_status();
}
