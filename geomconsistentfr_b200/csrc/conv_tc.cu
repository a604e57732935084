// 3x3 convolution on the 5th-generation tensor cores (tcgen05, kind::tf32) with fp32-grade accuracy (3xTF32
// operand splitting), TMA-staged activations, TMEM accumulators and the fused RelightNet epilogues.
// Replaces the cuDNN Conv2d / ConvTranspose2d(stride 1) + BatchNorm2d(eval) + LeakyReLU + residual / skip adds +
// nearest x2 upsample of TRAIN:197-350 / TEST1:170-323 (TRAIN = train_raytracing_relighting_CelebAHQ_DSSIM_8x.py,
// TEST1 = test_relight_single_image.py) for every 3x3 layer whose input has >= 16 channels.
//
// Activation layout ("C4"): [N][C/4][H][W][4] fp32 — 4-channel groups innermost, so one pixel x one channel group is
// a 16-byte unit.  That is exactly the 16-byte row of a K-major, un-swizzled UMMA core matrix (8 rows x 16 B), so a
// TMA box {10 px x 4ch, 18 rows, 4 groups} lands in shared memory as [group][row][px][4] and the A operand of every
// one of the nine filter taps is the SAME tile addressed through a descriptor whose start address is shifted by
// (ky*10 + kx)*16 bytes: M = 128 output pixels = 16 rows x 8 px (row pitch 160 B = SBO), K = 8 channels = two core
// matrices 2880 B apart (LBO).  Image borders and channel padding come from TMA out-of-bounds zero fill.
//
// 3xTF32: every fp32 operand is split x = hi + lo with hi = tf32(x); conv = hi*Whi + (hi*Wlo + lo*Whi) (the dropped
// lo*lo term is ~2^-22 relative).  Activations are split in shared memory after the TMA lands, weights are split and
// packed once on the host (gfr_conv_tc_pack_weights) as [Whi | Wlo] side by side in N, so one MMA of width 2*NT
// computes hi*Whi into the "main" accumulator columns and hi*Wlo into the "correction" columns, and a second MMA of
// width NT adds lo*Whi to the correction columns: 2 instead of 3 A-operand reads per K step.
//
// Accumulator rounding.  The tensor core accumulates into TMEM with truncation, so a long chain of MMAs into one
// accumulator drifts towards zero (measured: mean error grows linearly with Cin, anti-correlated with the sign of the
// result).  The chain is therefore cut every 16 input channels: each pipeline step starts a fresh accumulator, and the
// epilogue warps add (main + correction) into fp32 registers with round-to-nearest.
//
// Warp-specialised CTA (320 threads), S-slot shared-memory ring, double-buffered accumulators:
//   warp 0      TMA producer   : per step, box load of the fp32 tile (+ bulk copy of the step's weights when Cin > 16)
//   warps 2-5   splitters      : fp32 tile -> tf32 hi (in place) + lo, fence.proxy.async, arrive
//   warp 1      MMA issuer     : 36 tcgen05.mma per step (9 taps x 2 K-steps x 2), commit -> frees the slot + signals
//   warps 6-9   epilogue       : tcgen05.ld (thread = pixel), register accumulation; after the last step bias +
//                                residual + activation + up2(post) + scale, 128-byte coalesced float4 stores
#include "gfr_common.cuh"
#include "tc_common.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <mutex>
#include <math.h>
#include <stdlib.h>
#include <string.h>

using namespace gfr_tc;

namespace {

// halo-tile offset (in pixels) of filter tap `tap`: 3x3 taps (ky, kx) = (tap / 3, tap % 3), 2x2 taps (tap / 2, tap % 2)
template <int TAPS>
__host__ __device__ constexpr int tap_offset(int tap) { return TAPS == 9 ? (tap / 3) * 10 + (tap % 3) : (tap / 2) * 10 + (tap % 2); }

constexpr int TILE_PX_W = 8, TILE_PX_H = 16;           // 128 output pixels = UMMA M
constexpr int HALO_W = TILE_PX_W + 2, HALO_H = TILE_PX_H + 2;
constexpr int CB = 16;                                  // input channels per pipeline step (4 groups of 4)
constexpr uint32_t A_BYTES = (CB / 4) * HALO_H * HALO_W * 16;     // 11520
constexpr uint32_t A_LBO = HALO_H * HALO_W * 16;        // 2880: next 4-channel group
constexpr uint32_t A_SBO = HALO_W * 16;                 // 160: next tile row (8 pixels further in M)
constexpr int NUM_THREADS = 320;

struct ConvTcArgs {
  const float* wpk;    // packed weights, see gfr_conv_tc_pack_weights
  const float* bias;   // [Cout]
  const float* res;    // C4 [N][C4out][H][W][4] or null   (added before the activation)
  const float* post;   // C4 [N][C4out][H>>ps][W>>ps][4] or null (added after the activation, nearest x2 when ps=1)
  float* out;          // C4 [N][C4out][H][W][4]
  int N, Cin, Cout, H, W;   // H, W: OUTPUT size (= input size for the 3x3 / pad 1 layers)
  int Hin, Win;             // input size (the 2x2-tap layers map (H+1)x(W+1) -> HxW or HxW -> (H+1)x(W+1))
  int org;                  // the tile's TMA box starts at (x0 - org, y0 - org): 1 for 3x3 / pad 1, 0 | 1 for the 2x2 taps
  int ncb, tiles_x, tiles_y, m_tiles;
  int post_shift, act;
  int single_pass;     // 1: hi*Whi only (plain TF32)
  int static_w;        // 1: the packed weights were not written by the preceding kernel: fetch them before griddepcontrol.wait
  float out_scale;
  float x_scale, inv_scale;   // F16 kind: activations are multiplied by x_scale before the fp16 split, the sum by 1/(x_scale*w_scale)
  // HEAD variant (decoder tail fused into the epilogue of c2_1, TRAIN:284-290 / 344-350): the 16 activated channels of a pixel
  // go through c2_2, c2_3 (1x1, 16 -> 16, BN folded, LeakyReLU) and c2_o (1x1, 16 -> head_n) in the pixel's own thread
  const float* head;   // [w2 16x16 | b2 16 | w3 16x16 | b3 16 | wo 3x16 | bo 4] floats ([co][ci] rows), device
  float* head_out;     // NCHW [N][head_n][H][W]
  int head_n, head_act;
  float head_scale;
};

constexpr int KIND_TF32 = 0, KIND_F16 = 1, KIND_BF16 = 2;      // BF16: one bf16 operand per side, no correction MMAs

template <int NT, int KIND, int TAPS = 9>
struct Smem {
  // ring depth, measured on the whole forward (B = 8, 3 lanes): 3/3 slots (weights resident / streamed) 11 190 faces/s,
  // 3/2 11 800, 4/2 11 810, 2/2 11 990.  With streamed weights (Cin > 16) a third slot makes the CTA 124 KB and it runs
  // alone on its SM, with two slots (83 KB) two CTAs share it - the low-resolution layers are latency-bound and gain most.
#ifndef GFR_CONV_STAGES_RES
#define GFR_CONV_STAGES_RES 2
#endif
#ifndef GFR_CONV_STAGES_STR
#define GFR_CONV_STAGES_STR 2
#endif
#ifndef GFR_CONV_STAGES_WIDE
#define GFR_CONV_STAGES_WIDE 2
#endif
  // NT >= 64 (one CTA per SM: PatchGAN's layers and the 64 / 128 / 155-channel generator layers in bf16 training): ring depth
  // GFR_CONV_STAGES_WIDE (a step streams 11.5 KB of pixels + up to 16 KB of weights out of L2)
  // TF32: [tap][4-ch group (4)][hi|lo][n][4 floats];  F16: [tap][8-ch chunk (2)][w1|w2][n][8 halfs]  — of one 16-channel step
  //   BF16: [tap][8-ch chunk (2)][n][8 bf16].  TAPS = 9 (3x3) or 4 (2x2: PatchGAN's 4x4 / stride 2 layers over a space-to-depth)
  static constexpr uint32_t W_STEP = TAPS * (KIND == KIND_TF32 ? CB / 4 : CB / 8) * (KIND == KIND_BF16 ? 1 : 2) * NT * 16;
  // resident-weights mode (Cin <= 16): [W][slot: A_hi, A_lo] ; streaming mode: [slot: A_hi, A_lo, W]
  static constexpr uint32_t SLOT_RES = 2 * A_BYTES, SLOT_STR = 2 * A_BYTES + W_STEP;
  static constexpr int FIT_RES = (int)((226u * 1024u - 256u - W_STEP) / SLOT_RES), FIT_STR = (int)((226u * 1024u - 256u) / SLOT_STR);   // what shared memory holds
  static constexpr int WIDE_RES = GFR_CONV_STAGES_WIDE < FIT_RES ? GFR_CONV_STAGES_WIDE : FIT_RES, WIDE_STR = GFR_CONV_STAGES_WIDE < FIT_STR ? GFR_CONV_STAGES_WIDE : FIT_STR;
  static constexpr int STAGES_RES = NT <= 32 ? GFR_CONV_STAGES_RES : WIDE_RES, STAGES_STR = NT <= 32 ? GFR_CONV_STAGES_STR : WIDE_STR;
  static constexpr int MAX_STAGES = STAGES_RES > STAGES_STR ? STAGES_RES : STAGES_STR;
  static_assert(STAGES_RES >= 2 && STAGES_STR >= 2, "the ring needs two slots");
  static constexpr uint32_t BYTES_RES = W_STEP + STAGES_RES * SLOT_RES + 256;
  static constexpr uint32_t BYTES_STR = STAGES_STR * SLOT_STR + 256;
  static constexpr uint32_t ACC = (KIND == KIND_BF16 ? 1 : 2) * NT;          // columns of one accumulator buffer
  static constexpr uint32_t TMEM_COLS = 2 * ACC <= 32 ? 32 : (2 * ACC <= 64 ? 64 : (2 * ACC <= 128 ? 128 : (2 * ACC <= 256 ? 256 : 512)));
};

// y[o] = lrelu(b[o] + sum_ci w[o][ci] x[ci]) for 16 -> 16, weights in shared memory ([co][ci] rows)
__device__ __forceinline__ void head_dense16(const float* w, const float* b, const float (&x)[16], float (&y)[16]) {
#pragma unroll
  for (int o = 0; o < 16; o += 4) {
    float acc[4] = {b[o], b[o + 1], b[o + 2], b[o + 3]};
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(w + (o + r) * 16 + 4 * c4);
        acc[r] = fmaf(v.x, x[4 * c4], acc[r]); acc[r] = fmaf(v.y, x[4 * c4 + 1], acc[r]);
        acc[r] = fmaf(v.z, x[4 * c4 + 2], acc[r]); acc[r] = fmaf(v.w, x[4 * c4 + 3], acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) y[o + r] = acc[r] > 0.f ? acc[r] : 0.2f * acc[r];
  }
}

template <int NT, int KIND, bool HEAD = false, int TAPS = 9>
// CTAs per SM of the NT = 16 kernel: 3 would fit shared memory and TMEM but caps the kernel at 64 registers (256 B of
// spills in the epilogue): measured 10 550 vs 12 020 faces/s, so 2
#ifndef GFR_CONV_OCC
#define GFR_CONV_OCC 2
#endif
__global__ void __launch_bounds__(NUM_THREADS, NT <= 32 ? (NT <= 16 ? GFR_CONV_OCC : 2) : 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const ConvTcArgs a) {
  static_assert(!HEAD || NT == 16, "the fused decoder tail needs the pixel's 16 channels in one thread");
  using S = Smem<NT, KIND, TAPS>;
  constexpr bool CUT = KIND != KIND_BF16;      // accumulator chain cut at every pipeline step (fp32-grade kinds)
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool resident = a.ncb == 1;
  const int STAGES = resident ? S::STAGES_RES : S::STAGES_STR;
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t slot_bytes = resident ? S::SLOT_RES : S::SLOT_STR;
  const uint32_t slots0 = smem0 + (resident ? S::W_STEP : 0u);
  const uint32_t bars = slots0 + STAGES * slot_bytes;               // 8-byte aligned (all sizes are multiples of 128)
  // barrier map: full[s] = bars + 8 s, ready[s], empty[s] (STAGES each), accfull[p], accempty[p] (2 each), then the TMEM slot
  constexpr uint32_t SB = 8 * S::MAX_STAGES, TMEM_SLOT = 3 * SB + 32;
  const uint32_t bar_full = bars, bar_ready = bars + SB, bar_empty = bars + 2 * SB, bar_accfull = bars + 3 * SB, bar_accempty = bars + 3 * SB + 16;
  uint8_t* gen_bars = smem + (bars - smem0);

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_ready + 8 * s, 128);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int p = 0; p < 2; ++p) {
      mbar_init(bar_accfull + 8 * p, 1);
      mbar_init(bar_accempty + 8 * p, 128);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tm_in);
  }
  if (warp == 1) tmem_alloc(bars + TMEM_SLOT, S::TMEM_COLS);
  [[maybe_unused]] float* s_head = nullptr;
  if constexpr (HEAD) {                                       // static operand: not written by the preceding kernel
    __shared__ __align__(16) float s_head_buf[600];
    s_head = s_head_buf;
    for (int i = tid; i < 600; i += NUM_THREADS) s_head[i] = __ldg(a.head + i);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen_bars + TMEM_SLOT);

  const int n_my_tiles = (a.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_steps = n_my_tiles * a.ncb;
  griddep_launch_dependents();        // PDL: the next layer may start its prologue (barriers, TMEM, weight fetch) now

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (elect_one_sync()) {
      const float* wsrc = a.wpk + (size_t)blockIdx.y * a.ncb * (S::W_STEP / 4);
      // PDL prologue: the first ring slots are armed and (static weights) their weight copies are in flight before the
      // previous layer has finished; the activation loads follow griddepcontrol.wait.
      const int n_pre = n_steps < STAGES ? n_steps : STAGES;
      for (int g = 0; g < n_pre; ++g) {
        const bool load_w = !resident || g == 0;
        mbar_expect_tx(bar_full + 8 * g, A_BYTES + (load_w ? S::W_STEP : 0u));
        if (load_w && a.static_w)
          bulk_load(resident ? smem0 : slots0 + g * slot_bytes + 2 * A_BYTES, wsrc + (size_t)(g % a.ncb) * (S::W_STEP / 4), S::W_STEP, bar_full + 8 * g);
      }
      griddep_wait();
      int g = 0;
      for (int mt = blockIdx.x; mt < a.m_tiles; mt += gridDim.x) {
        const int tx = mt % a.tiles_x, t2 = mt / a.tiles_x;
        const int ty = t2 % a.tiles_y, n = t2 / a.tiles_y;
        for (int cb = 0; cb < a.ncb; ++cb, ++g) {
          const int s = g % STAGES;
          const uint32_t slot = slots0 + s * slot_bytes;
          const bool load_w = !resident || g == 0;
          if (g >= n_pre) {
            mbar_wait(bar_empty + 8 * s, ((g / STAGES) & 1) ^ 1);
            mbar_expect_tx(bar_full + 8 * s, A_BYTES + (load_w ? S::W_STEP : 0u));
          }
          tma_load_4d(slot, &tm_in, bar_full + 8 * s, (tx * TILE_PX_W - a.org) * 4, ty * TILE_PX_H - a.org, cb * (CB / 4), n);
          if (load_w && (g >= n_pre || !a.static_w))
            bulk_load(resident ? smem0 : slot + 2 * A_BYTES, wsrc + (size_t)cb * (S::W_STEP / 4), S::W_STEP, bar_full + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // The whole warp walks the pipeline (uniform waits); ONE elected lane issues the step's MMAs and commits: under
    // `elect.sync` every tcgen05.mma is one uniform-datapath instruction (+ one 64-bit add), under `if (lane == 0)` ptxas
    // wraps each of them in an ELECT / BRA.U.ANY loop (~10 instructions) and the issuing thread sets the pace.
    {
      constexpr uint32_t IDESC_2N = KIND == KIND_TF32 ? umma_idesc_tf32(128, 2 * NT) : umma_idesc_f16(128, 2 * NT);
      constexpr uint32_t IDESC_N = KIND == KIND_TF32 ? umma_idesc_tf32(128, NT) : (KIND == KIND_BF16 ? umma_idesc_bf16(128, NT) : umma_idesc_f16(128, NT));
      constexpr uint32_t B_LBO = (KIND == KIND_BF16 ? 1 : 2) * NT * 16, B_TAP = (KIND == KIND_TF32 ? CB / 4 : CB / 8) * B_LBO;
      for (int g = 0; g < n_steps; ++g) {
        // CUT kinds: accumulator buffer / parity per pipeline step; BF16: per tile (its steps accumulate in TMEM)
        const int s = g % STAGES, cb = g % a.ncb, e_acc = CUT ? g : g / a.ncb, p = e_acc & 1;
        const bool first_of_acc = CUT || cb == 0, last_of_acc = CUT || cb == a.ncb - 1;
        mbar_wait(bar_ready + 8 * s, (g / STAGES) & 1);
        if (first_of_acc) mbar_wait(bar_accempty + 8 * p, ((e_acc >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        if (elect_one_sync()) {
          const uint32_t slot = slots0 + s * slot_bytes;
          const uint32_t wbase = resident ? smem0 : slot + 2 * A_BYTES;
          const uint32_t d_main = tmem + (uint32_t)(p * S::ACC), d_corr = d_main + NT;
          const uint64_t dB0 = umma_desc_kmajor_noswz(wbase, B_LBO, 128u);
          if (KIND == KIND_TF32) {
            const uint64_t dA_hi0 = umma_desc_kmajor_noswz(slot, A_LBO, A_SBO);
            const uint64_t dA_lo0 = umma_desc_kmajor_noswz(slot + A_BYTES, A_LBO, A_SBO);
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const uint32_t ao = ((uint32_t)tap_offset<TAPS>(tap) * 16u + (uint32_t)j * 2u * A_LBO) >> 4;
                const uint32_t bo = ((uint32_t)tap * B_TAP + (uint32_t)j * 2u * B_LBO) >> 4;
                const uint32_t first = (tap == 0 && j == 0) ? 0u : 1u;
                if (a.single_pass) {
                  umma_tf32(d_main, dA_hi0 + ao, dB0 + bo, IDESC_N, first);
                } else {
                  umma_tf32(d_main, dA_hi0 + ao, dB0 + bo, IDESC_2N, first);      // main += hi*Whi ; corr += hi*Wlo
                  umma_tf32(d_corr, dA_lo0 + ao, dB0 + bo, IDESC_N, 1u);          // corr += lo*Whi
                }
              }
            }
          } else if (KIND == KIND_F16) {
            // fp16 pair split: the two operand tiles [2 chunks of 8 ch][18][10][8 halfs] sit behind the raw fp32 tile;
            // one K = 16 MMA covers the whole 16-channel step of a tap
            const uint64_t dA_1 = umma_desc_kmajor_noswz(slot + A_BYTES, A_LBO, A_SBO);
            const uint64_t dA_2 = umma_desc_kmajor_noswz(slot + A_BYTES + A_BYTES / 2, A_LBO, A_SBO);
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
              const uint32_t ao = ((uint32_t)tap_offset<TAPS>(tap) * 16u) >> 4;
              const uint32_t bo = ((uint32_t)tap * B_TAP) >> 4;
              umma_f16(d_main, dA_1 + ao, dB0 + bo, IDESC_2N, tap == 0 ? 0u : 1u);   // main += h1*W1 ; corr += h1*W2
              umma_f16(d_corr, dA_2 + ao, dB0 + bo, IDESC_N, 1u);                    // corr += h2*W1
            }
          } else {
            // bf16: one operand tile [2 chunks of 8 ch][18][10][8 bf16] behind the raw fp32 tile, one MMA per tap
            const uint64_t dA = umma_desc_kmajor_noswz(slot + A_BYTES, A_LBO, A_SBO);
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
              const uint32_t ao = ((uint32_t)tap_offset<TAPS>(tap) * 16u) >> 4;
              const uint32_t bo = ((uint32_t)tap * B_TAP) >> 4;
              umma_f16(d_main, dA + ao, dB0 + bo, IDESC_N, (tap == 0 && first_of_acc) ? 0u : 1u);
            }
          }
          umma_commit(bar_empty + 8 * s);       // the slot may be refilled once these MMAs have read it
          if (last_of_acc) umma_commit(bar_accfull + 8 * p);     // and the accumulators of this step (BF16: tile) are complete
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // =============================== splitters (128 threads) ===============================
    const int st = tid - 64;
    for (int g = 0; g < n_steps; ++g) {
      const int s = g % STAGES;
      mbar_wait(bar_full + 8 * s, (g / STAGES) & 1);
      uint8_t* slot_g = smem + (slots0 - smem0) + s * slot_bytes;
      if (KIND == KIND_TF32) {
        float4* hi4 = reinterpret_cast<float4*>(slot_g);
        float4* lo4 = reinterpret_cast<float4*>(slot_g + A_BYTES);
#pragma unroll
        for (int i = st; i < (int)(A_BYTES / 16); i += 128) {
          const float4 v = hi4[i];
          float4 h, l;
          h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
          l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
          hi4[i] = h;
          lo4[i] = l;
        }
      } else if (KIND == KIND_BF16) {
        constexpr int PIX = HALO_H * HALO_W;
        const float4* raw = reinterpret_cast<const float4*>(slot_g);
        uint4* t1 = reinterpret_cast<uint4*>(slot_g + A_BYTES);
        for (int i = st; i < 2 * PIX; i += 128) {
          const int c = i >= PIX ? 1 : 0, pix = i - c * PIX;
          const float4 va = raw[(2 * c) * PIX + pix], vb = raw[(2 * c + 1) * PIX + pix];
          const __nv_bfloat162 b0 = __floats2bfloat162_rn(va.x, va.y), b1 = __floats2bfloat162_rn(va.z, va.w);
          const __nv_bfloat162 b2 = __floats2bfloat162_rn(vb.x, vb.y), b3 = __floats2bfloat162_rn(vb.z, vb.w);
          t1[i] = make_uint4(*reinterpret_cast<const uint32_t*>(&b0), *reinterpret_cast<const uint32_t*>(&b1),
                             *reinterpret_cast<const uint32_t*>(&b2), *reinterpret_cast<const uint32_t*>(&b3));
        }
      } else {
        // x*x_scale = h1 + h2 (+ ~2^-22 relative), both fp16: 8 channels of a pixel -> one 16-byte row of each operand tile
        constexpr int PIX = HALO_H * HALO_W;                        // 180
        const float4* raw = reinterpret_cast<const float4*>(slot_g);
        uint4* t1 = reinterpret_cast<uint4*>(slot_g + A_BYTES);
        uint4* t2 = reinterpret_cast<uint4*>(slot_g + A_BYTES + A_BYTES / 2);
        const float xs = a.x_scale;
        for (int i = st; i < 2 * PIX; i += 128) {
          const int c = i >= PIX ? 1 : 0, pix = i - c * PIX;
          const float4 va = raw[(2 * c) * PIX + pix], vb = raw[(2 * c + 1) * PIX + pix];
          const float v[8] = {va.x * xs, va.y * xs, va.z * xs, va.w * xs, vb.x * xs, vb.y * xs, vb.z * xs, vb.w * xs};
          uint32_t w1[4], w2[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const __half2 h = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
            const float2 hf = __half22float2(h);
            const __half2 l = __floats2half2_rn(v[2 * k] - hf.x, v[2 * k + 1] - hf.y);
            w1[k] = *reinterpret_cast<const uint32_t*>(&h);
            w2[k] = *reinterpret_cast<const uint32_t*>(&l);
          }
          t1[i] = make_uint4(w1[0], w1[1], w1[2], w1[3]);
          t2[i] = make_uint4(w2[0], w2[1], w2[2], w2[3]);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_ready + 8 * s);
    }
  } else {
    // =============================== epilogue (128 threads, thread = pixel = TMEM lane) ===============================
    const int q = warp & 3;                               // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;
    const int n0 = blockIdx.y * NT;
    const int C4out = (a.Cout + 3) >> 2;
    const int pH = a.H >> a.post_shift, pW = a.W >> a.post_shift;
    griddep_wait();                                       // residual / skip operands and the output buffer belong to earlier kernels
    // BF16 kind (the training convolutions): this N tile's bias once per CTA in shared memory, zero beyond Cout
    __shared__ __align__(16) float s_bias[CUT ? 4 : NT];
    if constexpr (!CUT) {
      for (int i = tid - 192; i < NT; i += 128) s_bias[i] = (n0 + i < a.Cout) ? __ldg(a.bias + n0 + i) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    int g = 0, ti = 0;
    for (int mt = blockIdx.x; mt < a.m_tiles; mt += gridDim.x, ++ti) {
      const int tx = mt % a.tiles_x, t2 = mt / a.tiles_x;
      const int ty = t2 % a.tiles_y, n = t2 / a.tiles_y;
      // residual / skip operands are fetched before the accumulators are waited for: their L2 latency overlaps the MMAs
      const int y = ty * TILE_PX_H + (m >> 3), x = tx * TILE_PX_W + (m & 7);
      const bool ok = y < a.H && x < a.W;
      const size_t o_base = (((size_t)n * C4out * a.H + y) * a.W + x) * 4;            // + cq * H*W*4
      const size_t o_plane = (size_t)a.H * a.W * 4;
      // NT <= 16: into registers; wider tiles: an L2 prefetch (no registers), the loads themselves come after the MMAs
      constexpr bool PREF_REGS = NT <= 16 && !HEAD;      // (the HEAD variant has no residual / post operands)
      float4 rres[PREF_REGS ? NT / 4 : 1], rpost[PREF_REGS ? NT / 4 : 1];
#pragma unroll
      for (int c0 = 0; c0 < NT; c0 += 4) {
        const int cq = (n0 + c0) >> 2;
        if (PREF_REGS) {
          rres[c0 / 4] = make_float4(0.f, 0.f, 0.f, 0.f);
          rpost[c0 / 4] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (!HEAD && ok && cq < C4out) {
          const size_t po = ((((size_t)n * C4out + cq) * pH + (y >> a.post_shift)) * pW + (x >> a.post_shift)) * 4;
          if (PREF_REGS) {
            if (a.res) rres[c0 / 4] = __ldg(reinterpret_cast<const float4*>(a.res + o_base + (size_t)cq * o_plane));
            if (a.post) rpost[c0 / 4] = __ldg(reinterpret_cast<const float4*>(a.post + po));
          } else {
            if (a.res) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.res + o_base + (size_t)cq * o_plane));
            if (a.post) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.post + po));
          }
        }
      }
      // CUT kinds: every pipeline step has its own accumulator, summed here in fp32 registers (round-to-nearest);
      // BF16: the steps of a tile accumulate in TMEM, read once per tile in 16-column groups (no register array: NT up to 128)
      [[maybe_unused]] float sum[CUT ? NT : 1];
      if constexpr (CUT) {
#pragma unroll
        for (int c = 0; c < NT; ++c) sum[c] = 0.f;
        for (int cb = 0; cb < a.ncb; ++cb, ++g) {
          const int p = g & 1;
          mbar_wait(bar_accfull + 8 * p, (g >> 1) & 1);
          tc_fence_after_sync();
          const uint32_t t_main = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(p * S::ACC);
#pragma unroll
          for (int c0 = 0; c0 < NT; c0 += 16) {
            uint32_t rm[16], rc[16];
            tmem_ld16(t_main + c0, rm);
            if (!a.single_pass) tmem_ld16(t_main + NT + c0, rc);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float part = a.single_pass ? __uint_as_float(rm[e]) : __uint_as_float(rm[e]) + __uint_as_float(rc[e]);
              sum[c0 + e] += part;
            }
          }
          tc_fence_before_sync();
          mbar_arrive(bar_accempty + 8 * p);
        }
      } else {
        mbar_wait(bar_accfull + 8 * (ti & 1), (ti >> 1) & 1);
        tc_fence_after_sync();
      }
      if constexpr (!CUT && !HEAD) {
        if (a.res == nullptr && a.post == nullptr) {
          // The training convolutions (raw = acc + bias, PatchGAN's conv1 + LeakyReLU): nothing but the bias, the activation and the
          // store per 4-channel group.  The generic code below spends ~77 instructions per group (operand pointers re-read from the
          // constant bank, per-element bounds) on ONE warp per scheduler — 2.5 us per 128 x 64 tile, more than the tile's MMAs.
          float* outp = a.out + o_base + (size_t)(n0 >> 2) * o_plane;
          const int c4_lim = C4out - (n0 >> 2);
          const int act = a.act;
          const float scale = a.out_scale;
          const uint32_t t_base = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((ti & 1) * S::ACC);
          constexpr int GW = NT >= 32 ? 32 : 16;            // columns per TMEM round trip
#pragma unroll
          for (int cg = 0; cg < NT; cg += GW) {
            uint32_t rm[GW];
            tmem_ld16(t_base + cg, *reinterpret_cast<uint32_t(*)[16]>(rm));
            if constexpr (GW == 32) tmem_ld16(t_base + cg + 16, *reinterpret_cast<uint32_t(*)[16]>(rm + 16));
            tmem_ld_wait();
            if (ok) {
#pragma unroll
              for (int c0 = 0; c0 < GW; c0 += 4) {
                const int gi = (cg + c0) >> 2;
                if (gi < c4_lim) {
                  const float4 b = *reinterpret_cast<const float4*>(s_bias + cg + c0);
                  float v[4] = {__uint_as_float(rm[c0]) + b.x, __uint_as_float(rm[c0 + 1]) + b.y, __uint_as_float(rm[c0 + 2]) + b.z,
                                __uint_as_float(rm[c0 + 3]) + b.w};
                  if (act == 1) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = v[e] > 0.f ? v[e] : 0.2f * v[e];
                  }
                  *reinterpret_cast<float4*>(outp + (size_t)gi * o_plane) = make_float4(v[0] * scale, v[1] * scale, v[2] * scale, v[3] * scale);
                }
              }
            }
          }
          tc_fence_before_sync();
          mbar_arrive(bar_accempty + 8 * (ti & 1));
          continue;
        }
      }
      [[maybe_unused]] float yv[HEAD ? NT : 1];
#pragma unroll
      for (int cg = 0; cg < NT; cg += 16) {
        float acc16[16];
        if constexpr (CUT) {
#pragma unroll
          for (int e = 0; e < 16; ++e) acc16[e] = sum[cg + e];
        } else {
          uint32_t rm[16];
          tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((ti & 1) * S::ACC + cg), rm);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) acc16[e] = __uint_as_float(rm[e]);
        }
        if (ok) {
#pragma unroll
          for (int c0 = cg; c0 < cg + 16; c0 += 4) {
            const int co = n0 + c0, cq = co >> 2;
            if (cq >= C4out) continue;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              v[e] = (KIND == KIND_F16 ? acc16[c0 - cg + e] * a.inv_scale : acc16[c0 - cg + e]) + (co + e < a.Cout ? __ldg(a.bias + co + e) : 0.f);
            float4 rr = make_float4(0.f, 0.f, 0.f, 0.f), pp = rr;
            if (PREF_REGS) {
              rr = rres[c0 / 4]; pp = rpost[c0 / 4];
            } else if (!HEAD) {
              if (a.res) rr = __ldg(reinterpret_cast<const float4*>(a.res + o_base + (size_t)cq * o_plane));
              if (a.post) pp = __ldg(reinterpret_cast<const float4*>(a.post + ((((size_t)n * C4out + cq) * pH + (y >> a.post_shift)) * pW + (x >> a.post_shift)) * 4));
            }
            v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
            if (a.act == 1) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = v[e] > 0.f ? v[e] : 0.2f * v[e];
            } else if (a.act == 2) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = 1.0f / (1.0f + expf(-v[e]));
            }
            v[0] += pp.x; v[1] += pp.y; v[2] += pp.z; v[3] += pp.w;
            if constexpr (HEAD) {
#pragma unroll
              for (int e = 0; e < 4; ++e) yv[c0 + e] = v[e] * a.out_scale;
            } else {
              *reinterpret_cast<float4*>(a.out + o_base + (size_t)cq * o_plane) =
                  make_float4(v[0] * a.out_scale, v[1] * a.out_scale, v[2] * a.out_scale, v[3] * a.out_scale);
            }
          }
        }
      }
      if constexpr (!CUT) {
        tc_fence_before_sync();
        mbar_arrive(bar_accempty + 8 * (ti & 1));
      }
      if constexpr (HEAD) {
        if (ok) {
          // c2_2 -> c2_3 -> c2_o per pixel.  Weights come from shared memory as LDS.128 broadcasts; four outputs are
          // accumulated side by side (independent FMA chains); per accumulator the order is the stand-alone
          // head_1x1_kernel's (bias first, ci ascending), so the result is bit-identical to the two-launch tail.
          float h2[16];
          head_dense16(s_head, s_head + 256, yv, h2);
          head_dense16(s_head + 272, s_head + 528, h2, yv);
          const size_t hw_px = (size_t)a.H * a.W;
          float* op = a.head_out + (size_t)n * a.head_n * hw_px + (size_t)y * a.W + x;
#pragma unroll
          for (int o = 0; o < 3; ++o) {
            if (o >= a.head_n) break;
            float acc = s_head[592 + o];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const float4 w = *reinterpret_cast<const float4*>(s_head + 544 + o * 16 + 4 * c4);
              acc = fmaf(w.x, yv[4 * c4], acc); acc = fmaf(w.y, yv[4 * c4 + 1], acc);
              acc = fmaf(w.z, yv[4 * c4 + 2], acc); acc = fmaf(w.w, yv[4 * c4 + 3], acc);
            }
            if (a.head_act == 2) acc = 1.0f / (1.0f + expf(-acc));
            op[(size_t)o * hw_px] = acc * a.head_scale;
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, S::TMEM_COLS);
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

// tensor [N][C4][H][W][4] fp32 described as 4-D {W*4, H, C4, N} (pixel and 4-channel slot merged: one contiguous
// 16*bw-byte run per box row — a 16-byte innermost TMA dimension costs one request per pixel); box {4*bw, bh, 4 groups, 1}
int make_c4_map(CUtensorMap* tm, const float* base, int N, int C4, int C4_alloc, int H, int W, int bw, int bh) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return GFR_E_UNSUPPORTED;
  const cuuint64_t dims[4] = {(cuuint64_t)W * 4, (cuuint64_t)H, (cuuint64_t)C4, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)C4_alloc * H * W * 16};
  const cuuint32_t box[4] = {(cuuint32_t)bw * 4, (cuuint32_t)bh, CB / 4, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? GFR_OK : GFR_E_ARG;
}

template <int NT, int KIND, bool HEAD = false, int TAPS = 9>
int launch_tc(const CUtensorMap& tm, const ConvTcArgs& a, cudaStream_t s) {
  using S = Smem<NT, KIND, TAPS>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    const uint32_t mx = S::BYTES_RES > S::BYTES_STR ? S::BYTES_RES : S::BYTES_STR;
    attr_err = cudaFuncSetAttribute(conv3x3_tc_kernel<NT, KIND, HEAD, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx);
  });
  if (attr_err != cudaSuccess) return (int)attr_err;
  const uint32_t bytes = a.ncb == 1 ? S::BYTES_RES : S::BYTES_STR;
  const int n_tiles = gfr_ceil_div(a.Cout, NT);
  int occ = (int)(225u * 1024u / (bytes + 1024u));
  const int occ_max = NT <= 32 ? (NT <= 16 ? GFR_CONV_OCC : 2) : 1;
  if (occ > occ_max) occ = occ_max;
  if (occ < 1) occ = 1;
  // GFR_TC_GRID_OCC=1: one persistent CTA per SM even where two fit (see conv_p16.cu: the free slot overlaps another kernel)
  static const int grid_occ = [] { const char* e = getenv("GFR_TC_GRID_OCC"); const int v = e ? atoi(e) : 0; return v == 1 ? 1 : 2; }();
  int gx = (sm_count() * (occ < grid_occ ? occ : grid_occ)) / n_tiles;
  if (gx < 1) gx = 1;
  if (gx > a.m_tiles) gx = a.m_tiles;
  // Programmatic dependent launch between consecutive conv layers (the next layer's prologue - barriers, TMEM, static weight
  // fetch - overlaps this layer's tail): -3.3 % forward latency (0.833 -> 0.806 ms for 8 faces), throughput unchanged; the
  // whole GPU test-suite (eval, train, graphs, two streams) runs green with it.  GFR_PDL=0 switches it off.
  static const bool no_pdl = [] { const char* e = getenv("GFR_PDL"); return e != nullptr && e[0] == '0'; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gx, n_tiles);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<NT, KIND, HEAD, TAPS>, tm, a);
  return e == cudaSuccess ? gfr_launch_status() : (int)e;
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
      for (int tap = 0; tap < 9; ++tap)
        for (int kc = 0; kc < CB / 4; ++kc)
          for (int part = 0; part < 2; ++part)
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

extern "C" long long gfr_conv_tc_pack_size_f16(int Cin, int Cout, int NT) {
  if (Cin <= 0 || Cout <= 0 || (NT != 16 && NT != 32 && NT != 64)) return GFR_E_ARG;
  return (long long)gfr_ceil_div(Cout, NT) * gfr_ceil_div(Cin, CB) * 9 * (CB / 8) * 2 * NT * 4;     // in floats (8 halfs = 4 floats)
}

extern "C" int gfr_conv_tc_pack_weights_f16(const float* w_host, int Cin, int Cout, int NT, float w_scale, float* packed_host) {
  GFR_RETURN_IF_NULL(w_host); GFR_RETURN_IF_NULL(packed_host);
  if (Cin <= 0 || Cout <= 0 || (NT != 16 && NT != 32 && NT != 64) || !(w_scale > 0.f)) return GFR_E_ARG;
  const int n_tiles = gfr_ceil_div(Cout, NT), ncb = gfr_ceil_div(Cin, CB);
  __half* out = reinterpret_cast<__half*>(packed_host);
  size_t o = 0;
  for (int nt = 0; nt < n_tiles; ++nt)
    for (int cb = 0; cb < ncb; ++cb)
      for (int tap = 0; tap < 9; ++tap)
        for (int ch = 0; ch < CB / 8; ++ch)
          for (int part = 0; part < 2; ++part)
            for (int n = 0; n < NT; ++n)
              for (int e = 0; e < 8; ++e, ++o) {
                const int co = nt * NT + n, ci = cb * CB + ch * 8 + e;
                float v = 0.f;
                if (co < Cout && ci < Cin) {
                  const float w = w_host[((size_t)co * Cin + ci) * 9 + tap] * w_scale;
                  if (!(fabsf(w) < 65000.f)) return GFR_E_ARG;            // w_scale too large for fp16
                  const float w1 = __half2float(__float2half_rn(w));
                  v = part == 0 ? w1 : w - w1;
                }
                out[o] = __float2half_rn(v);
              }
  return GFR_OK;
}

// taps 9: 3x3 / pad 1 (Hin = H, Win = W, org 1).  taps 4: 2x2 taps, out[y][x] = sum_{ky,kx in {0,1}} w[ky][kx] in[y - org + ky][x - org + kx]:
//   org 0 maps an (H+1) x (W+1) input to H x W (forward of PatchGAN's 4x4 / stride 2 layers over the space-to-depth of the padded
//   input), org 1 maps Hin x Win to (Hin+1) x (Win+1) (its data gradient).  precision 1 TF32, 2 fp16 pair split, 3 3xTF32, 4 bf16.
static int conv_tc_launch(const float* in, const float* w_packed, const float* bias, const float* res, const float* post, float* out,
                          int N, int Cin, int in_groups, int Cout, int Hin, int Win, int H, int W, int NT, int taps, int org,
                          int post_shift, int act, float out_scale, int precision, int weights_static, float x_scale, float w_scale,
                          void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(w_packed); GFR_RETURN_IF_NULL(bias); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0 || Hin <= 0 || Win <= 0) return GFR_E_SHAPE;
  if (post_shift < 0 || post_shift > 1 || act < 0 || act > 2 || precision < 1 || precision > 4) return GFR_E_ARG;
  if ((taps != 9 && taps != 4) || org < 0 || org > 1) return GFR_E_ARG;
  if (precision == 2 && !(x_scale > 0.f && w_scale > 0.f)) return GFR_E_ARG;
  if (post && post_shift && ((H | W) & 1)) return GFR_E_SHAPE;
  if (in_groups == 0) in_groups = (Cin + 3) / 4;
  if (in_groups < (Cin + 3) / 4) return GFR_E_ARG;
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(w_packed) | reinterpret_cast<uintptr_t>(out) |
       reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(post)) & 15)
    return GFR_E_ARG;
  ConvTcArgs a;
  a.wpk = w_packed; a.bias = bias; a.res = res; a.post = post; a.out = out;
  a.N = N; a.Cin = Cin; a.Cout = Cout; a.H = H; a.W = W; a.Hin = Hin; a.Win = Win; a.org = org;
  a.ncb = gfr_ceil_div(Cin, CB);
  a.tiles_x = gfr_ceil_div(W, TILE_PX_W); a.tiles_y = gfr_ceil_div(H, TILE_PX_H);
  a.m_tiles = N * a.tiles_x * a.tiles_y;
  a.post_shift = post_shift; a.act = act; a.out_scale = out_scale;
  a.single_pass = precision == 1;
  a.x_scale = x_scale; a.inv_scale = precision == 2 ? 1.0f / (x_scale * w_scale) : 1.0f;
  a.static_w = weights_static ? 1 : 0;
  a.head = nullptr; a.head_out = nullptr; a.head_n = 0; a.head_act = 0; a.head_scale = 1.0f;
  CUtensorMap tm;
  const int rc = make_c4_map(&tm, in, N, (Cin + 3) / 4, in_groups, Hin, Win, HALO_W, HALO_H);
  if (rc != GFR_OK) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int kind = precision == 2 ? KIND_F16 : (precision == 4 ? KIND_BF16 : KIND_TF32);
#define GFR_TC_CASE(nt, k, t) if (NT == nt && kind == k && taps == t) return launch_tc<nt, k, false, t>(tm, a, s)
  GFR_TC_CASE(16, KIND_TF32, 9); GFR_TC_CASE(32, KIND_TF32, 9); GFR_TC_CASE(64, KIND_TF32, 9);
  GFR_TC_CASE(16, KIND_F16, 9); GFR_TC_CASE(32, KIND_F16, 9); GFR_TC_CASE(64, KIND_F16, 9);
  GFR_TC_CASE(16, KIND_BF16, 9); GFR_TC_CASE(32, KIND_BF16, 9); GFR_TC_CASE(64, KIND_BF16, 9); GFR_TC_CASE(128, KIND_BF16, 9);
  GFR_TC_CASE(16, KIND_TF32, 4); GFR_TC_CASE(64, KIND_TF32, 4);
  GFR_TC_CASE(16, KIND_BF16, 4); GFR_TC_CASE(64, KIND_BF16, 4); GFR_TC_CASE(128, KIND_BF16, 4);
#undef GFR_TC_CASE
  return GFR_E_ARG;
}

extern "C" int gfr_conv3x3_tc_fwd(const float* in, const float* w_packed, const float* bias, const float* res,
                                  const float* post, float* out, int N, int Cin, int in_groups, int Cout, int H, int W,
                                  int NT, int post_shift, int act, float out_scale, int precision, int weights_static, float x_scale,
                                  float w_scale, void* stream) {
  return conv_tc_launch(in, w_packed, bias, res, post, out, N, Cin, in_groups, Cout, H, W, H, W, NT, 9, 1, post_shift, act, out_scale,
                        precision, weights_static, x_scale, w_scale, stream);
}

extern "C" int gfr_conv_tc_fwd_ex(const float* in, const float* w_packed, const float* bias, const float* res, const float* post,
                                  float* out, int N, int Cin, int in_groups, int Cout, int Hin, int Win, int H, int W, int NT,
                                  int taps, int org, int post_shift, int act, float out_scale, int precision, int weights_static,
                                  void* stream) {
  if (precision == 2) return GFR_E_ARG;          // the fp16 pair split is the eval path's (gfr_conv3x3_tc_fwd / gfr_conv3x3_p16_fwd)
  if (taps == 9 && (Hin != H || Win != W)) return GFR_E_SHAPE;
  if (taps == 4 && !((org == 0 && Hin == H + 1 && Win == W + 1) || (org == 1 && Hin + 1 == H && Win + 1 == W))) return GFR_E_SHAPE;
  return conv_tc_launch(in, w_packed, bias, res, post, out, N, Cin, in_groups, Cout, Hin, Win, H, W, NT, taps, org, post_shift, act,
                        out_scale, precision, weights_static, 1.0f, 1.0f, stream);
}

extern "C" int gfr_conv3x3_tc_head_fwd(const float* in, const float* w_packed, const float* bias, const float* head, float* out,
                                       int N, int Cin, int in_groups, int H, int W, int n_out, int head_act, float head_scale,
                                       int precision, int weights_static, float x_scale, float w_scale, void* stream) {
  GFR_RETURN_IF_NULL(in); GFR_RETURN_IF_NULL(w_packed); GFR_RETURN_IF_NULL(bias); GFR_RETURN_IF_NULL(head); GFR_RETURN_IF_NULL(out);
  if (N <= 0 || Cin <= 0 || H <= 0 || W <= 0 || n_out < 1 || n_out > 3) return GFR_E_SHAPE;
  if ((head_act != 0 && head_act != 2) || precision < 2 || precision > 3) return GFR_E_ARG;
  if (precision == 2 && !(x_scale > 0.f && w_scale > 0.f)) return GFR_E_ARG;
  if (in_groups == 0) in_groups = (Cin + 3) / 4;
  if (in_groups < (Cin + 3) / 4) return GFR_E_ARG;
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(w_packed) | reinterpret_cast<uintptr_t>(head)) & 15) return GFR_E_ARG;
  ConvTcArgs a;
  a.wpk = w_packed; a.bias = bias; a.res = nullptr; a.post = nullptr; a.out = nullptr;
  a.N = N; a.Cin = Cin; a.Cout = 16; a.H = H; a.W = W; a.Hin = H; a.Win = W; a.org = 1;
  a.ncb = gfr_ceil_div(Cin, CB);
  a.tiles_x = gfr_ceil_div(W, TILE_PX_W); a.tiles_y = gfr_ceil_div(H, TILE_PX_H);
  a.m_tiles = N * a.tiles_x * a.tiles_y;
  a.post_shift = 0; a.act = 1; a.out_scale = 1.0f;
  a.single_pass = 0;
  a.x_scale = x_scale; a.inv_scale = precision == 2 ? 1.0f / (x_scale * w_scale) : 1.0f;
  a.static_w = weights_static ? 1 : 0;
  a.head = head; a.head_out = out; a.head_n = n_out; a.head_act = head_act; a.head_scale = head_scale;
  CUtensorMap tm;
  const int rc = make_c4_map(&tm, in, N, (Cin + 3) / 4, in_groups, H, W, HALO_W, HALO_H);
  if (rc != GFR_OK) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  return precision == 2 ? launch_tc<16, KIND_F16, true>(tm, a, s) : launch_tc<16, KIND_TF32, true>(tm, a, s);
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
