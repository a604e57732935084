/* gfr_b200.h — C ABI of libgfr_b200.so: the B200 (sm_100a) implementation of the relight hot path
 * of andrewhou1/GeomConsistentFR.
 *
 * The reference has no FFI of its own (the path is inlined in RelightNet.forward); every entry point
 * below therefore cites the reference LINES it replaces.  Citations are relative to the reference root:
 *   TRAIN = train_raytracing_relighting_CelebAHQ_DSSIM_8x.py      TEST1 = test_relight_single_image.py
 *
 * Conventions (all entry points)
 *   - plain pointers and sizes; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns every buffer; the library never allocates device memory and never synchronises;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - tensors are contiguous, row-major, fp32 unless stated; images are NCHW;
 *   - return value: 0 on success, a negative GFR_E_* argument error, or a positive cudaError_t;
 *     gfr_error_string() describes either;
 *   - re-entrant, no global state except the lazily resolved driver entry point for TMA descriptors.
 */
#ifndef GFR_B200_H_
#define GFR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GFR_OK 0
#define GFR_E_NULL (-1)      /* a required pointer is NULL            */
#define GFR_E_SHAPE (-2)     /* unsupported shape / size              */
#define GFR_E_ARG (-3)       /* bad scalar argument                   */
#define GFR_E_UNSUPPORTED (-4)

#define GFR_MAX_SAMPLES 256  /* the sample table travels as a kernel parameter */

int gfr_version(void);
const char* gfr_error_string(int code);

/* mask dtypes accepted by gfr_mask_pack */
#define GFR_MASK_U8 0
#define GFR_MASK_F32 1
#define GFR_MASK_F64 2

#define GFR_MASK_EXTRA_WORDS 4

/* Face-mask bit packing.  The march only ever tests `mask[r, c] == 0` (TRAIN:510, TEST1:488), so the
 * mask travels as 1 bit per pixel: bit (r*W + c) & 31 of word (r*W + c) >> 5 is set iff mask != 0.
 * mask: [n_masks, H, W] of the given dtype; bits: [n_masks, H*W/32 + GFR_MASK_EXTRA_WORDS] uint32 — each row is the
 * bitmap followed by the bounding box of the non-zero pixels as int32 {c_lo, -c_hi, r_lo, -r_hi} (the march uses it to
 * skip samples that cannot lie on the face).  H*W must be a multiple of 32. */
int gfr_mask_pack(const void* mask, int mask_dtype, int n_masks, int H, int W, uint32_t* bits, void* stream);

/* Ray-march forward: minimum point-to-ray distance over the samples, per pixel.
 * Replaces TRAIN:374-515 / TEST1:351-496 (end points, 160 fp64 sample positions, bilinear depth,
 * point-to-line distance, face-mask reject, min, optional "+5 when the light projects inside the image").
 *   depth      [B,1,H,W]   raw depth (100 x the depth head), fp32
 *   mask_bits  [1|B, H*W/32 + 4] from gfr_mask_pack; mask_batch_stride = 0 (one mask for the batch,
 *              TEST1:488) or H*W/32 + GFR_MASK_EXTRA_WORDS (one per image, TRAIN:510), in uint32 words
 *   light_pt   [B,3]       light_distance * unit(L)   (TRAIN:360-363)
 *   t_host     [n] HOST doubles: the sample parameters, exactly np.arange(0.025, 0.825, 0.005) for the
 *              reference configuration (TRAIN:468); n <= GFR_MAX_SAMPLES
 *   inside_bonus  0 (train) or 5 (test, TEST1:495-496)
 *   bonus_rect_host  NULL: the bonus applies when the light projects inside the image rectangle (TEST1:495); else HOST
 *              floats {x_min, x_max, y_min, y_max} the projected light point must lie in (the lighting-transfer script
 *              uses +-4 image sizes: {-4W, 4W, 4(1-H), 4H}, TEST_LT:503)
 *   d_min      [B,H,W] out
 *   argmin     [B,H,W] out, uint8 index of the minimising sample (255 = every sample was outside the
 *              face); may be NULL.  Needed by the backward.
 *   shadow     [B,H,W] out, 1 - 4e^-d/(1+e^-d)^2 (TRAIN:517); may be NULL.
 *   depth64_scratch  (B/L)*H*W + B/L doubles of caller-owned scratch: the kernel widens the depth map into it once so
 *              the per-sample gathers need no fp32->fp64 conversion, and keeps every face's depth range (two ints) behind
 *              it for the early cut-off; NULL selects variant 1.
 *   lights_per_face  L >= 1: B counts (face, light) pairs, pair b uses the depth map and mask of face b / L (depth,
 *              depth64_scratch and per-image masks then hold B / L entries) — one CNN pass relit under L lights
 *              (the reference's 18-light Multi-PIE sweep, TESTB:565-583).  1 = one light per face.
 *   variant    0 = default (conversion-free rounding + fp64 depth scratch, 2 launches); 1 = reference-literal
 *              conversions (cvt.rni/floor/ceil per sample, 1 launch); both are bit-identical, used by tests/bench.
 */
int gfr_shadow_march_fwd(const float* depth, const uint32_t* mask_bits, int mask_batch_stride,
                         const float* light_pt, const double* t_host, int n, float inside_bonus,
                         const float* bonus_rect_host, float* d_min, uint8_t* argmin, float* shadow, double* depth64_scratch, int B, int H, int W,
                         int lights_per_face, int variant, void* stream);

/* A/B configuration of variant 0's kernel (process-wide; results are bit-identical in every setting):
 *   warp_shape  -1 default (environment GFR_MARCH_WARP, else 1), 0 = a warp marches 32 x 1 pixels, 1 = 8 x 4 pixels
 *   ilp          0 default (environment GFR_MARCH_ILP, else 2), 1 = samples one by one, 2..4 = in groups of 2..4
 *   block_order -1 default (environment GFR_MARCH_ORDER, else 1), 0 = tile-major CTA order, 1 = (face, light) pairs interleaved,
 *               each pair's tiles far-from-its-light first (load balance: a CTA's cost varies 1 : 700 across the image)
 *   early_cut   -1 default (environment GFR_MARCH_CUT, else 1), 0 = every sample of the culled range is walked, 1 = a ray stops
 *               at the sample index beyond which the pixel -> light line is provably farther above (below) every depth an
 *               in-mask sample can return than the ray's current minimum distance (exact: csrc/shadow_march.cu)
 * Same loop as TRAIN:467-515 either way; no reference counterpart (a tuning knob for tests / tools). */
int gfr_march_config(int warp_shape, int ilp, int block_order, int early_cut);

/* Normals + Lambertian shading + shadow blend + albedo render.  Replaces TRAIN:353-369 and 517-522
 * (kornia depth_to_normals(depth + depth_offset, K), y flip, double normalise, l = normalize(P_L - P),
 * dir = intensity * max(n.l, 0), full = ambient + dir, s = shadow(d_min),
 * final = s*full + (1-s)*ambient, rendered = albedo*final).
 *   albedo [B,3,H,W]; depth [B,1,H,W]; d_min [B,H,W]; light_pt [B,3]; ambient [B]
 *   intr_host: HOST floats {fx, fy, cx, cy, depth_offset(1610), directional_intensity(0.5)}
 *   outputs (any may be NULL): shadow [B,H,W], full [B,H,W], final [B,H,W], rendered [B,3,H,W],
 *   normals [B,3,H,W]
 *   lights_per_face L: as in gfr_shadow_march_fwd — albedo, depth and ambient hold B / L faces, d_min / light_pt /
 *   the outputs B (face, light) pairs.
 */
int gfr_shade_render_fwd(const float* albedo, const float* depth, const float* d_min, const float* light_pt,
                         const float* ambient, const float* intr_host, float* shadow, float* full,
                         float* final_shading, float* rendered, float* normals, int B, int H, int W,
                         int lights_per_face, void* stream);

/* gfr_shadow_march_fwd and gfr_shade_render_fwd fused into one launch (+ the depth-widening pre-pass): every thread
 * shades its pixel right after its ray march, d_min never makes the round trip through memory.  Arguments as in the
 * two functions; d_min / argmin / shadow / full / final_shading / rendered / normals may each be NULL;
 * depth64_scratch is required.  Bit-identical to the two-launch sequence. */
int gfr_march_shade_fwd(const float* albedo, const float* depth, const uint32_t* mask_bits, int mask_batch_stride,
                        const float* light_pt, const float* ambient, const double* t_host, int n, float inside_bonus,
                        const float* bonus_rect_host,
                        const float* intr_host, double* depth64_scratch, float* d_min, uint8_t* argmin, float* shadow,
                        float* full, float* final_shading, float* rendered, float* normals, int B, int H, int W,
                        int lights_per_face, void* stream);

/* Ray-march backward (what autograd does for TRAIN:374-517): the min over samples routes the gradient to the arg-min
 * sample, which is re-evaluated with the forward's arithmetic; gradients flow to the pixel's depth, the four bilinear
 * corners of the sample, and the light point — directly (BC = P_L - P) and through the sample position (end point <-
 * slope/intercept <- projected light, TRAIN:378-458; none where the end point is clamped or the light projects inside
 * the image).  g_dmin [B,H,W] = dL/d(d_min); argmin from gfr_shadow_march_fwd; t_host / n as in the forward.
 * g_depth [B,H,W] and g_light [B,3] are ACCUMULATED into (atomicAdd): zero-fill them first. */
int gfr_shadow_march_bwd(const float* depth, const float* light_pt, const uint8_t* argmin, const float* g_dmin,
                         const double* t_host, int n, float* g_depth, float* g_light, int B, int H, int W,
                         void* stream);

/* Backward of gfr_shade_render_fwd (autograd through TRAIN:353-369, 517-522).  Upstream gradients g_shadow, g_full,
 * g_final [B,H,W], g_rendered, g_normals [B,3,H,W] may each be NULL (= zero).  Outputs: g_albedo [B,3,H,W] (written;
 * may be NULL), g_dmin [B,H,W] (written; may be NULL), and ACCUMULATED (zero-fill first): g_depth [B,H,W],
 * g_light [B,3], g_ambient [B]. */
int gfr_shade_render_bwd(const float* albedo, const float* depth, const float* d_min, const float* light_pt,
                         const float* ambient, const float* intr_host, const float* g_shadow, const float* g_full,
                         const float* g_final, const float* g_rendered, const float* g_normals, float* g_albedo,
                         float* g_depth, float* g_dmin, float* g_light, float* g_ambient, int B, int H, int W,
                         void* stream);

/* ---- training losses and optimiser (TRAIN:589-590, 633-656) ---------------------------------------------------
 * SSIM (pytorch_msssim.ssim at TRAIN:643: 11-tap Gaussian sigma 1.5, VALID, per (n,c) plane).  X, Y [P,H,W] with
 * P = N*C planes.  Forward: plane_sums [P] doubles are ACCUMULATED (zero-fill first) with the sum of the
 * (H-10)x(W-10) SSIM map of each plane (mean = sum / ((H-10)(W-10))); grad_maps [3,P,H-10,W-10] (may be NULL) receives
 * dS/dmu_x, dS/dE[xx], dS/dE[xy] for the backward.  Backward: g_X [P,H,W] = plane_scale[p] * (transposed filter of the
 * three maps), plane_scale[p] = dL/d(mean SSIM of plane p) / ((H-10)(W-10)). */
int gfr_ssim_fwd(const float* X, const float* Y, double* plane_sums, float* grad_maps, int P, int H, int W,
                 float data_range, void* stream);
int gfr_ssim_bwd(const float* X, const float* Y, const float* grad_maps, const float* plane_scale, float* g_X, int P,
                 int H, int W, void* stream);

/* Masked reconstruction / depth / albedo losses with their gradients (TRAIN:633-639).  rendered, img_nchw, albedo
 * [N,3,H,W]; depth, depth_gt, albedo_gt, mask_fill, mask [N,H,W].  sums5 (5 doubles, overwritten):
 *   {sum(mask_fill), sum(mask), sum((mask_fill*(rendered-img))^2), sum|mask*(depth-depth_gt)|,
 *    sum|mask_fill*(mean_c albedo - albedo_gt)|}
 * so recon = 20*s[2]/(3*s[0]), depth = s[3]/s[1], albedo = 5*s[4]/s[0].  g_rendered / g_depth / g_albedo (each may be
 * NULL) receive the gradients of those three loss terms. */
int gfr_masked_losses(const float* rendered, const float* img_nchw, const float* depth, const float* depth_gt,
                      const float* albedo, const float* albedo_gt, const float* mask_fill, const float* mask,
                      double* sums5, float* g_rendered, float* g_depth, float* g_albedo, int N, int H, int W,
                      void* stream);

/* One torch.optim.Adam step (no weight decay, no amsgrad; TRAIN:589-590, 656) over a flat fp32 parameter buffer.
 * state3: 3 device floats {step count, 1-beta1^step, sqrt(1-beta2^step)}, zero-initialised by the caller and advanced
 * on the device by every call (no host-side counter, so the whole training step can be replayed from a CUDA graph);
 * grads are multiplied by grad_scale first (1/world_size after an all-reduce). */
int gfr_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float* state3, float lr,
                  float beta1, float beta2, float eps, float grad_scale, void* stream);

/* The same step with torch.optim.Adam's PER-PARAMETER state (TRAIN:589-590, 656): the flat buffer is cut into n_seg
 * segments (seg_start[0..n_seg], device int64, seg_start[n_seg] = n), one per parameter tensor.  seg_state: n_seg x 4 device
 * floats {step, 1-beta1^step, sqrt(1-beta2^step), active}, 16-byte aligned; only segments with active != 0 tick and are
 * updated — torch skips parameters whose .grad is None (the epoch-gated skip blocks before their gate opens,
 * TRAIN:245,258,271,283) and starts their bias correction at their first gradient. */
int gfr_adam_step_segments(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                           const long long* seg_start, float* seg_state, int n_seg, float lr, float beta1, float beta2,
                           float eps, float grad_scale, void* stream);

/* ---- train-mode CNN building blocks (the reference trains with BATCH-statistics BatchNorm: it never calls .eval(),
 * TRAIN:561-563) — all activations C4 -------------------------------------------------------------------------- */

/* Device-side packing for gfr_conv3x3_tc_fwd (weights change every optimiser step).  w is the layer PARAMETER on the
 * device: Conv2d [Cout,Cin,3,3] (is_transposed_conv 0) or ConvTranspose2d [Cin,Cout,3,3] (1; its forward is a conv
 * with w.transpose(0,1).flip(2,3)).  for_dgrad 0: operand of the forward (Cin -> Cout); 1: operand of the
 * data-gradient, itself a 3x3 conv Cout -> Cin with the transposed + flipped kernel.  Same layout / size as
 * gfr_conv_tc_pack_weights (gfr_conv_tc_pack_size(I, O, NT) floats, O/I = outputs/inputs of the packed operand). */
int gfr_conv_tc_pack_weights_dev(const float* w, int is_transposed_conv, int for_dgrad, int Cin, int Cout, int NT,
                                 float* packed, void* stream);

/* The general forms (training path).  taps 9 = 3x3 / pad 1; taps 4 = 2x2 taps over k x k = 2 x 2 parameters [O][I][2][2]:
 * PatchGAN's 4x4 / stride 2 / pad 1 layers (TRAIN:18-27) are 2x2-tap convolutions over the space-to-depth of the 1-padded
 * input (gfr_space_to_depth_pad) with NO structurally-zero weights.  precision 1 TF32 | 3 3xTF32 | 4 bf16 (one bf16 operand
 * per side, fp32 accumulation in TMEM — BASELINE configs[2] "bf16 CNN").  NT 16|32|64|128 (128: bf16 only; 2x2 taps: 16|64|128).
 * gfr_conv_tc_pack_size_ex returns FLOATS.  gfr_conv_tc_fwd_ex: out[y][x] = sum_taps w . in[y - org + ky][x - org + kx];
 * taps 9: Hin = H, Win = W, org 1; taps 4: org 0 with (Hin, Win) = (H+1, W+1), or org 1 with (H, W) = (Hin+1, Win+1) (the
 * data gradient of the former).  Epilogue as gfr_conv3x3_tc_fwd. */
long long gfr_conv_tc_pack_size_ex(int Cin, int Cout, int NT, int taps, int precision);
int gfr_conv_tc_pack_weights_dev_ex(const float* w, int is_transposed_conv, int for_dgrad, int Cin, int Cout, int NT, int taps,
                                    int precision, float* packed, void* stream);

/* All packed operands of a training step in ONE launch.  gfr_conv_tc_pack_job_fill writes one record of the job table (host memory,
 * gfr_conv_tc_pack_job_size() bytes per record; arguments as gfr_conv_tc_pack_weights_dev_ex; first_block = the sum of the block
 * counts of the earlier jobs) and returns the job's block count (< 0: error); gfr_conv_tc_pack_weights_batch runs the device copy
 * of the table.  Replaces ~120 per-layer pack launches of a training iteration (TRAIN:617-656: every conv layer's forward and
 * data-gradient operand is re-packed after each Adam step, TRAIN:656). */
int gfr_conv_tc_pack_job_size(void);
long long gfr_conv_tc_pack_job_fill(void* job_host, const float* w, int is_transposed_conv, int for_dgrad, int Cin, int Cout, int NT,
                                    int taps, int precision, float* packed, long long first_block);
int gfr_conv_tc_pack_weights_batch(const void* jobs_dev, int n_jobs, long long n_blocks, void* stream);
int gfr_conv_tc_fwd_ex(const float* in, const float* w_packed, const float* bias, const float* res, const float* post, float* out,
                       int N, int Cin, int in_groups, int Cout, int Hin, int Win, int H, int W, int NT, int taps, int org,
                       int post_shift, int act, float out_scale, int precision, int weights_static, void* stream);

/* BatchNorm2d, training mode, part 1: batch statistics of x [N,C,H,W] (C4) -> mean, rstd = 1/sqrt(var_biased + eps),
 * scale = gamma*rstd, shift = beta - mean*scale (all [4*ceil(C/4)] floats, padded slots 0); running_mean/var (may be
 * NULL) are updated like torch (momentum, unbiased variance).  sums_scratch: 2*4*ceil(C/4) + 1 doubles (the last one is the
 * ticket counter with which the last CTA of the statistics pass finalises: one launch). */
int gfr_bn_train_stats(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                       double* sums_scratch, float* mean, float* rstd, float* scale, float* shift, int N, int C, int H,
                       int W, float eps, float momentum, void* stream);
/* the same + BatchNorm2d's `num_batches_tracked += 1` (device int64, may be NULL) from the finalising CTA: the 65 BatchNorms of
 * a generator forward otherwise cost 65 one-element launches per iteration (TRAIN:197-350 in train mode) */
int gfr_bn_train_stats_ex(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                          long long* num_batches_tracked, double* sums_scratch, float* mean, float* rstd, float* scale,
                          float* shift, int N, int C, int H, int W, float eps, float momentum, void* stream);

/* gfr_bn_config(1): the caller guarantees that every `sums_scratch` it hands to gfr_bn_train_stats[_ex] / gfr_bn_apply_bwd[_ex] is
 * all zero (e.g. slices of one arena cleared once per training step), so the entry points skip their own memset; 0 (default) restores
 * it.  Other values only query.  Returns the previous setting.  Process-wide. */
int gfr_bn_config(int scratch_prezeroed);

/* One more momentum update of running_mean / running_var (+ num_batches_tracked += 1, may be NULL) from the batch sums that a
 * gfr_bn_train_stats[_ex] call left in `sums_scratch`: exactly what a second train-mode forward over the same input with the same
 * weights does to the BatchNorm buffers.  The reference runs the discriminator twice on the same composite (TRAIN:619 and 641); on
 * the iterations without a discriminator update (TRAIN:624) both passes are the same function, so the second is replaced by this. */
int gfr_bn_running_update(const double* sums_scratch, float* running_mean, float* running_var, long long* num_batches_tracked,
                          int N, int C, int H, int W, float momentum, void* stream);

/* part 2: y = act(scale*x + shift + res) + up(post)  — BatchNorm + residual add + LeakyReLU(0.2) (act 1) + skip /
 * nearest-x2-upsample add, the epilogue of gfr_conv3x3_tc_fwd as its own pass.  res/post may be NULL. */
int gfr_bn_apply_fwd(const float* x, const float* scale, const float* shift, const float* res, const float* post,
                     float* y, int N, int C, int H, int W, int post_shift, int act, void* stream);

/* Backward of part 1+2 w.r.t. x and res given g_y = dL/dy (the `post` branch's gradient is g_y itself, or its 2x2 sum
 * — gfr_sumpool2_c4 — when it was upsampled): g_res (may be NULL) = g_y*act'(pre); g_x = gamma*rstd*(g_pre - mean(g_pre)
 * - xhat*mean(g_pre*xhat)).  After the call sums_scratch holds [sum g_pre | sum g_pre*xhat] = [d beta | d gamma]
 * (2 x 4*ceil(C/4) doubles).  gamma_pad: gamma padded with zeros to 4*ceil(C/4). */
int gfr_bn_apply_bwd(const float* x, const float* res, const float* g_y, const float* scale, const float* shift,
                     const float* mean, const float* rstd, const float* gamma_pad, double* sums_scratch, float* g_x,
                     float* g_res, int N, int C, int H, int W, int act, void* stream);

/* gfr_bn_apply_bwd with the UNPADDED parameter gamma [C] and the parameter gradients accumulated in place:
 * g_gamma[c] += sum g_pre*xhat, g_beta[c] += sum g_pre (either may be NULL) — the optimiser's flat gradient buffer can be
 * passed directly (no temporaries, no separate add).  g_bias (may be NULL): g_bias[c] += sum over (N,H,W) of g_x[:,c] — the
 * gradient of the bias of the convolution in front of the BatchNorm, from the same pass. */
int gfr_bn_apply_bwd_ex(const float* x, const float* res, const float* g_y, const float* scale, const float* shift,
                        const float* mean, const float* rstd, const float* gamma, double* sums_scratch, float* g_x,
                        float* g_res, float* g_gamma, float* g_beta, float* g_bias, int N, int C, int H, int W, int act,
                        void* stream);

/* Weight (and bias) gradient of a 3x3 stride-1 convolution: g_w (+=, parameter layout: Conv2d [Cout,Cin,3,3] or
 * ConvTranspose2d [Cin,Cout,3,3]) and g_bias [Cout] (+=, may be NULL) from the layer input `in`
 * (in_groups as in gfr_conv3x3_tc_fwd) and g_out = dL/d(conv output).  fp32 on CUDA cores. */
int gfr_conv3x3_wgrad(const float* in, const float* g_out, float* g_w, float* g_bias, int is_transposed_conv, int N,
                      int Cin, int in_groups, int Cout, int H, int W, void* stream);

/* The same weight gradient on the tensor cores with bf16 operands and fp32 accumulation in TMEM (train_precision 4,
 * BASELINE configs[2] "bf16 CNN"): per filter tap a GEMM over PIXELS whose operands are the C4 tensors themselves as
 * MN-major UMMA matrices (csrc/wgrad_tc.cu).  taps 9: 3x3 / pad 1 (Hin = H, Win = W); taps 4: the 2x2-tap layers
 * (Hin = H+1, Win = W+1).  g_w (+=) in the parameter layout as gfr_conv3x3_wgrad; the bias gradient is
 * gfr_channel_sum_c4(g_out): out[c] += sum over (N,H,W) of a C4 tensor, out exactly C floats. */
int gfr_conv_wgrad_tc_bf16(const float* in, const float* g_out, float* g_w, int is_transposed_conv, int N, int Cin,
                           int in_groups, int Cout, int Hin, int Win, int H, int W, int taps, void* stream);
/* A/B switch of gfr_conv_wgrad_tc_bf16 for 3x3 layers with <= 16 input channels (process-wide; same sums in a different MMA
 * arrangement): -1 default, 0 = one MMA per filter tap, 1 = the input as the M operand with pixel-shifted row blocks (the three
 * taps of a filter row per MMA: 6 instead of 9 MMAs per K block). */
int gfr_wgrad_tc_config(int pixel_shift_form);
int gfr_channel_sum_c4(const float* x, float* out, int N, int C, int H, int W, void* stream);

/* 2x2 max-pool backward (gradient to the first maximum, like torch), 2x2 sum (backward of the nearest x2 upsample),
 * global average pool of channels [c_first, c_first+n_ch) of a C4 map -> [N,n_ch] and its backward (+= into g_feat). */
int gfr_maxpool2_c4_bwd(const float* x, const float* g_y, float* g_x, int NC4, int Ho, int Wo, void* stream);
int gfr_sumpool2_c4(const float* x, float* out, int NC4, int Ho, int Wo, void* stream);
int gfr_avgpool_c4_fwd(const float* feat, float* out, int N, int C, int c_first, int n_ch, int HW, void* stream);
int gfr_avgpool_c4_bwd(const float* g, float* g_feat, int N, int C, int c_first, int n_ch, int HW, void* stream);

/* 1x1 convolution with 16 input channels and DEVICE weights w [Cout,16], bias [Cout] (the decoder tails in train mode,
 * TRAIN:285-290, 345-350): out = out_scale * act(w x + b); planar_out 0: C4 [N,16,H,W] (Cout = 16), 1: NCHW
 * [N,Cout,H,W]; act 0 | 2 (sigmoid).  Backward: g_in (written), g_w / g_bias (+=) from g_out = dL/d(out). */
int gfr_pw_conv16_fwd(const float* in, const float* w, const float* bias, float* out, int N, int Cout, int H, int W,
                      int planar_out, int act, float out_scale, void* stream);
int gfr_pw_conv16_bwd(const float* in, const float* w, const float* g_out, const float* out, float* g_in, float* g_w,
                      float* g_bias, int N, int Cout, int H, int W, int planar_out, int act, float out_scale, void* stream);

/* Stem in train mode: conv_c1_og (5x5, 3 -> 16) with DEVICE weights w [16,3,5,5], bias [16] on the NHWC image -> raw
 * conv output C4 [N,16,H,W] (BatchNorm / LeakyReLU / pool follow as separate passes); and its weight/bias gradient
 * (+=) from g_out = dL/d(raw). */
int gfr_stem_conv_train_fwd(const float* img, const float* w, const float* bias, float* out_raw, int N, int H, int W,
                            void* stream);
int gfr_stem_conv_wgrad(const float* img, const float* g_out, float* g_w, float* g_bias, int N, int H, int W, void* stream);

/* ---- PatchGAN discriminator support (TRAIN:15-35) ---------------------------------------------------------------
 * conv1..conv4 (4x4, stride 2, pad 1) run as 3x3 / stride 1 convolutions on the tensor cores over a space-to-depth of
 * their input: out[n][c][y][x][2dy+dx] = in[n, c, 2y+dy, 2x+dx] (channel c becomes one C4 group holding its four
 * phases; H, W even).  in: NCHW planes (in_is_nchw 1, e.g. the [N,3,H,W] image) or C4; out: C4 [N,4C,H/2,W/2].
 * gfr_depth_to_space is the inverse (= the backward). */
int gfr_space_to_depth(const float* in, float* out, int N, int C, int H, int W, int in_is_nchw, void* stream);
int gfr_depth_to_space(const float* g, float* out, int N, int C, int H, int W, int out_is_nchw, void* stream);

/* The same over the 1-PADDED input: out [N,4C,H/2+1,W/2+1] with out[n][c][y'][x'][2dy+dx] = Xpad[2y'+dy][2x'+dx],
 * Xpad[r][s] = X[r-1][s-1] (0 outside).  A 4x4 / stride 2 / pad 1 convolution of X (TRAIN:18-27) is a 2x2-tap / stride 1
 * convolution of `out` (gfr_conv_tc_fwd_ex, taps 4) with weights W'[co][4c+2dy+dx][a][b] = w[co][c][2a+dy][2b+dx]: no
 * structurally-zero weights (the 3x3 form above multiplies 5/9 zeros).  gfr_depth_to_space_pad is its backward
 * (g [N,4C,H/2+1,W/2+1] -> [N,C,H,W]); gfr_conv2x2_wgrad the weight gradient of the 2x2-tap layer (in [N,Cin,H+1,W+1],
 * g_out [N,Cout,H,W], g_w [Cout,Cin,2,2] +=). */
int gfr_space_to_depth_pad(const float* in, float* out, int N, int C, int H, int W, int in_is_nchw, void* stream);
int gfr_depth_to_space_pad(const float* g, float* out, int N, int C, int H, int W, int out_is_nchw, void* stream);
int gfr_conv2x2_wgrad(const float* in, const float* g_out, float* g_w, float* g_bias, int N, int Cin, int in_groups, int Cout,
                      int H, int W, void* stream);

/* g_pre = g_y * (y > 0 ? 1 : 0.2): backward of y = LeakyReLU(pre, 0.2) from the forward OUTPUT (n_floats % 4 == 0). */
int gfr_lrelu_bwd_c4(const float* y, const float* g_y, float* g_pre, long long n_floats, void* stream);

/* conv5 of the PatchGAN: 4x4, stride 1, pad 1, C -> 1 (TRAIN:33).  in C4 [N,C,H,W] (C % 4 == 0); w [1,C,4,4]; bias [1];
 * out [N,1,H-1,W-1].  Backward: g_in C4 (written, may be NULL), g_w / g_bias (+=, may both be NULL). */
int gfr_conv4x4s1_to1_fwd(const float* in, const float* w, const float* bias, float* out, int N, int C, int H, int W, void* stream);
int gfr_conv4x4s1_to1_bwd(const float* in, const float* w, const float* g_out, float* g_in, float* g_w, float* g_bias, int N, int C,
                          int H, int W, void* stream);

/* fp32 convolution with fused epilogue (exact-fp32 CNN path).  Replaces one
 * Conv2d / ConvTranspose2d(stride 1) + BatchNorm2d(eval, folded into w/bias by the caller) + residual add +
 * LeakyReLU(0.2) / sigmoid + skip add + nearest x2 upsample step of RelightNet (TRAIN:197-350, TEST1:170-323):
 *     out = out_scale * ( act( conv(in') + bias + res ) + up(post) )
 *   in      activations; element strides {sN, sC, sH, sW} in in_strides_host (HOST, 4 x int64) so that an NHWC
 *           image (TRAIN:197 permute) or a channel slice (TRAIN:225) is read in place; in' = in, or its nearest
 *           x2 upsampling when ups_in = 1 (TRAIN:240 etc.; `in` is then [N,Cin,H/2,W/2])
 *   w       [Cout,Cin,K,K], K in {1,3,5}, stride 1, padding K/2 (a ConvTranspose2d(k=3,s=1,p=1) weight must be
 *           passed as w.transpose(0,1).flip(2,3))
 *   res     [N,Cout,H,W] or NULL; post [N,Cout,H>>post_shift,W>>post_shift] or NULL
 *   act     0 none, 1 LeakyReLU(0.2), 2 sigmoid
 */
int gfr_conv2d_fwd(const float* in, const long long* in_strides_host, const float* w, const float* bias,
                   const float* res, const float* post, float* out, int N, int Cin, int Cout, int H, int W, int K,
                   int ups_in, int post_shift, int act, float out_scale, void* stream);

/* ---- tensor-core (tcgen05) CNN path ------------------------------------------------------------------------
 * Activations of this path use the "C4" layout [N][ceil(C/4)][H][W][4] fp32 (channel c of pixel (y,x) lives in
 * group c/4, slot c%4; the slots past C in the last group hold zeros).  All C4 pointers must be 16-byte aligned. */

/* NCHW <-> C4 layout conversion. */
int gfr_nchw_to_c4(const float* in, float* out, int N, int C, int H, int W, void* stream);
int gfr_c4_to_nchw(const float* in, float* out, int N, int C, int H, int W, void* stream);

/* Weight packing for gfr_conv3x3_tc_fwd (HOST function, host pointers).  w_host is [Cout,Cin,3,3] (BN folded; a
 * ConvTranspose2d(k=3,s=1,p=1) weight as w.transpose(0,1).flip(2,3)).  The packed buffer holds, for every
 * (NT-channel output tile, 16-channel input step), the tf32 "hi" and "lo" halves of the weights as K-major UMMA
 * core matrices: [n_tile][cin_step][tap 0..8][4-channel group 0..3][hi|lo][n 0..NT-1][4].  NT in {16,32,64}.
 * gfr_conv_tc_pack_size returns the number of floats (negative on a bad argument). */
long long gfr_conv_tc_pack_size(int Cin, int Cout, int NT);
int gfr_conv_tc_pack_weights(const float* w_host, int Cin, int Cout, int NT, float* packed_host);

/* The same for precision = 2 (fp16 pair split): weights are multiplied by w_scale (a power of two that keeps
 * |w|*w_scale < 65000, chosen by the caller from max|w|) and stored as fp16 pairs w1 = fp16(w s), w2 = fp16(w s - w1):
 * [n_tile][cin_step][tap][8-channel chunk 0..1][w1|w2][n 0..NT-1][8 halfs].  Size in floats. */
long long gfr_conv_tc_pack_size_f16(int Cin, int Cout, int NT);
int gfr_conv_tc_pack_weights_f16(const float* w_host, int Cin, int Cout, int NT, float w_scale, float* packed_host);

/* 3x3 convolution (stride 1, padding 1) on the tcgen05 tensor cores, 3xTF32 (fp32-grade accuracy), fused epilogue
 *     out = out_scale * ( act( conv(in) + bias + res ) + up(post) )
 * Replaces one Conv2d / ConvTranspose2d(s=1) + BatchNorm2d(eval) + residual add + LeakyReLU + skip add + nearest x2
 * upsample step of RelightNet (TRAIN:197-350, TEST1:170-323) for layers with Cin >= 16.
 *   in [N,Cin,H,W] C4, read in place from a buffer that holds in_groups >= ceil(Cin/4) channel groups per image
 *   (0 = dense; > ceil(Cin/4) reads the leading channels of a wider tensor, TRAIN:225); w_packed (device) from gfr_conv_tc_pack_weights with the same NT; bias [Cout] (device);
 *   res C4 [N,Cout,H,W] or NULL; post C4 [N,Cout,H>>post_shift,W>>post_shift] or NULL; out C4 [N,Cout,H,W];
 *   act 0 none, 1 LeakyReLU(0.2), 2 sigmoid;
 *   precision 3 = 3xTF32 (hi*hi + lo*hi + hi*lo; full fp32 exponent range: training, gradients), 1 = single-pass TF32
 *   (what cuDNN does under torch.backends.cudnn.allow_tf32, the reference's default on Ampere+; not parity-grade for
 *   the depth head), 2 = fp16 pair split (kind::f16, K = 16 per MMA: half the shared-memory operand bytes and MMA count
 *   of 3xTF32 at the same ~22-bit products; operands are x*x_scale and w*w_scale as fp16 pairs, so |x|*x_scale and
 *   |w|*w_scale must stay below 65504 — inference on bounded activations; w_packed from gfr_conv_tc_pack_weights_f16);
 *   weights_static 1: w_packed was NOT written by the kernel that precedes this call in the stream (inference) — the
 *   kernel is launched with programmatic dependent launch and fetches its weights while the previous layer is still
 *   running; 0 (training: the pack kernel precedes the conv) fetches them after the dependency has resolved. */
int gfr_conv3x3_tc_fwd(const float* in, const float* w_packed, const float* bias, const float* res, const float* post,
                       float* out, int N, int Cin, int in_groups, int Cout, int H, int W, int NT, int post_shift,
                       int act, float out_scale, int precision, int weights_static, float x_scale, float w_scale,
                       void* stream);

/* gfr_conv3x3_tc_fwd for a 16-output-channel layer with LeakyReLU, with the decoder tail fused into its epilogue: the
 * 16 channels of every pixel go straight through c2_2, c2_3 (1x1, 16 -> 16, BN folded, LeakyReLU) and c2_o (1x1,
 * 16 -> n_out; head_act 0 = none | 2 = sigmoid; x head_scale) — TRAIN:284-290 (albedo) / 344-350 (depth) — without a
 * round trip of the 16-channel map through memory.  head: DEVICE floats [w2 16x16 | b2 16 | w3 16x16 | b3 16 |
 * wo 3x16 | bo 4] ([co][ci] rows, 600 floats, 16-byte aligned); out: NCHW [N, n_out, H, W].  precision 2 | 3. */
int gfr_conv3x3_tc_head_fwd(const float* in, const float* w_packed, const float* bias, const float* head, float* out,
                            int N, int Cin, int in_groups, int H, int W, int n_out, int head_act, float head_scale,
                            int precision, int weights_static, float x_scale, float w_scale, void* stream);

/* ---- eval-mode tensor-core CNN, second generation: PRE-SPLIT fp16-pair activations ("P16", csrc/p16.cuh) -------------
 * Every activation x is stored as the fp16 pair the tensor cores consume: 16 x = hi + lo, layout
 * [N][groups = ceil(C/8)][hi|lo][H][W][8 halfs] (4 bytes per element, |x| < 4094).  Replaces the same reference lines as
 * gfr_conv3x3_tc_fwd (TRAIN:197-350 / TEST1:170-323: Conv2d / ConvTranspose2d(stride 1) + BatchNorm(eval, folded) +
 * LeakyReLU + residual / skip adds + nearest x2 upsample). */

/* HOST packing for gfr_conv3x3_p16_fwd.  w_host [Cout,Cin,3,3] (BN folded; ConvTranspose2d as w.transpose(0,1).flip(2,3));
 * the buffer holds [n_tile][cin_step][tap][chunk 0..KS-1][w1|w2][n 0..NT-1][8 halfs] with w*w_scale = w1 + w2 as fp16.
 * NT in {16,32}, KS in {2,4} (8*KS input channels per pipeline step).  gfr_conv_p16_pack_size returns HALFS. */
long long gfr_conv_p16_pack_size(int Cin, int Cout, int NT, int KS);
int gfr_conv_p16_pack_weights(const float* w_host, int Cin, int Cout, int NT, int KS, float w_scale, void* packed_host);

/* out = out_scale * (act(conv3x3(in[:, :Cin]) + bias + res) + up(post)), all tensors P16.
 *   in_groups / out_groups / res_groups / post_groups: 8-channel chunks ALLOCATED per image in that tensor (0 = exactly
 *   ceil(C/8)); res_c8: first chunk of the residual operand inside its tensor.
 *   act 0 none | 1 LeakyReLU(0.2) | 2 sigmoid, applied to output channels < act_channels only (0 = all): a residual block's
 *   first conv and its shortcut conv (same input, TRAIN:203-223 / 235-239) run as ONE launch with concatenated output
 *   channels, and the block's second conv reads the leading channels as input and the trailing ones as `res`, in place.
 *   MH 1 | 2: 8x16- or 16x16-pixel CTA tile (one or two M = 128 MMAs over one TMA box).
 *   flags (may be NULL): device int, bit 0 is OR-ed in when an output leaves the split's range (|x| >= 4094 or NaN).
 *   Padding channels (>= Cout, up to 8*out_groups) are written as zeros. */
int gfr_conv3x3_p16_fwd(const void* in, const void* w_packed, const float* bias, const void* res, int res_c8, int res_groups,
                        const void* post, int post_groups, void* out, int out_groups, int* flags, int N, int Cin,
                        int in_groups, int Cout, int H, int W, int NT, int MH, int KS, int post_shift, int act,
                        int act_channels, float out_scale, float w_scale, int weights_static, void* stream);

/* The general form.  geometry 0: 3x3 / pad 1 (gfr_conv3x3_p16_fwd).  geometry 1: 5x1 VERTICAL taps / pad (2, 0), weights
 * [Cout][Cin][5] packed with gfr_conv_p16_pack_weights_taps(taps = 5) — the stem's 5x5 / pad 2 convolution (conv_c1_og,
 * TRAIN:197-200) after gfr_stem_unroll_p16 has unrolled its five horizontal taps into channels:
 *     U[n][kx*3 + c][y][x] = img[n][y][x + kx - 2][c]  (P16, 16 channels, channel 15 = 0; img NHWC fp32)
 *     conv5x5(img)[co][y][x] = sum_ky sum_c' W5[co][c'][ky] U[c'][y + ky - 2][x],  W5[co][kx*3 + c][ky] = w[co][c][ky][kx]
 * (NT 16, MH 2, KS 2 only).  pool_out (may be NULL): P16 [N][out_groups][2][H/2][W/2][8], the 2x2 / stride 2 max pool of the
 * layer's output (TRAIN:201,206,212,218) written by the same epilogue (warp shuffles across the 2x2 quad; H, W even). */
int gfr_conv_p16_fwd_ex(const void* in, const void* w_packed, const float* bias, const void* res, int res_c8, int res_groups,
                        const void* post, int post_groups, void* out, int out_groups, void* pool_out, int* flags, int N, int Cin,
                        int in_groups, int Cout, int H, int W, int NT, int MH, int KS, int geometry, int post_shift, int act,
                        int act_channels, float out_scale, float w_scale, int weights_static, void* stream);
long long gfr_conv_p16_pack_size_taps(int Cin, int Cout, int NT, int KS, int taps);
int gfr_conv_p16_pack_weights_taps(const float* w_host, int Cin, int Cout, int NT, int KS, int taps, float w_scale, void* packed_host);
int gfr_stem_unroll_p16(const float* img, void* out, int N, int H, int W, void* stream);
/* Persistent grid of the P16 convolutions (process-wide A/B knob, results identical): 0 = default (environment
 * GFR_P16_GRID_OCC, else 1), 1 = one CTA per SM (throughput: the free slot overlaps another kernel), 2 = two where they fit
 * (lowest latency of a single launch). */
int gfr_conv_p16_config(int grid_ctas_per_sm);

/* NCHW fp32 <-> P16, 2x2 max pool on P16 (compares the joined fp32 values), and the P16 forms of the stem (out / pooled
 * P16 [N,16,H,W] / [N,16,H/2,W/2]), the fused 1x1 decoder tail (in P16 [N,16,H,W]) and the light head (feat P16 with
 * `groups` chunks). */
int gfr_nchw_to_p16(const float* in, void* out, int N, int C, int H, int W, void* stream);
int gfr_p16_to_nchw(const void* in, float* out, int N, int C, int groups, int H, int W, void* stream);
int gfr_maxpool2_p16_fwd(const void* in, void* out, int NC8, int Ho, int Wo, void* stream);
int gfr_stem_conv_p16_fwd(const float* img, const float* w_host, const float* bias_host, void* out, void* pooled, int N,
                          int H, int W, void* stream);
int gfr_head_1x1_p16_fwd(const void* in, const float* w2_host, const float* b2_host, const float* w3_host,
                         const float* b3_host, const float* wo_host, const float* bo_host, float* out, int N, int H, int W,
                         int n_out, int act, float out_scale, void* stream);
int gfr_light_head_p16_fwd(const void* feat, int groups, int c_first, int HW, const float* w1, const float* b1,
                           const float* w2, const float* b2, float* out, int N, void* stream);

/* A decoder's last 3x3 layer with its 1x1 tail in the epilogue (round 2): conv_*_c2_1 + BN(folded) + LeakyReLU (Cin -> 16,
 * TRAIN:284 / 344) followed, on the pixel each epilogue thread holds, by conv_*_c2_2 / c2_3 (+ BN + LeakyReLU) and
 * conv_*_c2_o (+ sigmoid for the albedo, TRAIN:285-290; x out_scale = 100 for the depth, TRAIN:345-350) — the arguments of
 * gfr_conv3x3_p16_fwd (NT = 16, KS = 2) and of gfr_head_1x1_p16_fwd in one launch; the 16-channel activation between them is
 * never stored.  out [N,n_out,H,W] fp32. */
int gfr_conv3x3_p16_head_fwd(const void* in, const void* w_packed, const float* bias, int N, int Cin, int in_groups, int H, int W,
                             int MH, float w_scale, int weights_static, const float* w2_host, const float* b2_host,
                             const float* w3_host, const float* b3_host, const float* wo_host, const float* bo_host,
                             float* out, int n_out, int act, float out_scale, void* stream);

/* Stem: conv_c1_og (5x5, 3 -> 16, padding 2) + BatchNorm(eval, folded) + LeakyReLU(0.2) on the NHWC image, with the
 * first 2x2 max pool fused (TRAIN:197-201).  img [N,H,W,3]; w_host [16,3,5,5] and bias_host [16] are HOST pointers
 * (they travel as kernel parameters); out C4 [N,16,H,W]; pooled C4 [N,16,H/2,W/2] or NULL. */
int gfr_stem_conv_fwd(const float* img, const float* w_host, const float* bias_host, float* out, float* pooled, int N,
                      int H, int W, void* stream);

/* Decoder tail: c2_2 and c2_3 (1x1, 16 -> 16, BN folded, LeakyReLU) and c2_o (1x1, 16 -> n_out) fused per pixel
 * (TRAIN:285-290 albedo: n_out 3, act 2 = sigmoid; TRAIN:345-350 depth: n_out 1, act 0, out_scale 100).
 * in C4 [N,16,H,W]; all weights/biases are HOST pointers ([16,16],[16],[16,16],[16],[n_out,16],[n_out]);
 * out NCHW [N,n_out,H,W]. */
int gfr_head_1x1_fwd(const float* in, const float* w2_host, const float* b2_host, const float* w3_host,
                     const float* b3_host, const float* wo_host, const float* bo_host, float* out, int N, int H, int W,
                     int n_out, int act, float out_scale, void* stream);

/* gfr_light_head_fwd on a C4 feature map feat [N,C,HW]: channels [c_first, c_first+27). */
int gfr_light_head_c4_fwd(const float* feat, int C, int c_first, int HW, const float* w1, const float* b1,
                          const float* w2, const float* b2, float* out, int N, void* stream);

/* 2x2/2 max pool and nearest x2 upsample (+ optional add) in the C4 layout; NC4 = N * ceil(C/4). */
int gfr_maxpool2_c4_fwd(const float* in, float* out, int NC4, int Ho, int Wo, void* stream);
int gfr_upsample2_c4_fwd(const float* in, const float* add, float* out, int NC4, int Ho, int Wo, void* stream);

/* 2x2/2 max pool, NCHW: in [NC, 2*Ho, 2*Wo] -> out [NC, Ho, Wo]  (TRAIN:201,206,212,218). */
int gfr_maxpool2_fwd(const float* in, float* out, int NC, int Ho, int Wo, void* stream);

/* nearest x2 upsample with optional add: out[NC,Ho,Wo] = up2(in[NC,Ho/2,Wo/2]) (+ add[NC,Ho,Wo])
 * (nn.Upsample(scale_factor=2, mode='nearest'), TRAIN:240,253,266,278 when the epoch gate is off). */
int gfr_upsample2_fwd(const float* in, const float* add, float* out, int NC, int Ho, int Wo, void* stream);

/* Light head (TRAIN:225-232): global average pool of channels [c_first, c_first+27) of feat [N,C,HW]
 * (feat_batch_stride = C*HW elements), Linear 27->128 (w1 [128,27], b1), LeakyReLU(0.2), Linear 128->4
 * (w2 [4,128], b2) -> out [N,4] = {ambient, lx, ly, lz}. */
int gfr_light_head_fwd(const float* feat, long long feat_batch_stride, int c_first, int HW, const float* w1,
                       const float* b1, const float* w2, const float* b2, float* out, int N, void* stream);

/* ---- output stage of the inference drivers (csrc/postprocess.cu) --------------------------------------------------
 * The reference copies every fp32 plane to the host and quantises there in numpy; these entry points produce the
 * 8-bit images cv2.imwrite would store, on the device.  Arithmetic = numpy's promotion in the reference expressions
 * (fp32 product with 255, fp64 product with mask/255, round-half-to-even, saturate).  masks are the skin masks as
 * stored (u8, values {0,64,128,255}); mask_batch_stride = 0 (one mask for the batch) or H*W. */

/* TEST1:613-620 / TESTB:596-601: out[b,r,c,:] (BGR) = 255 * rendered[b,2-ch,r,c] * mask/255 where mask > 0, else
 * 255 * image[b,r,c,2-ch].  image [B,H,W,3] RGB in [0,1], fp64 (the reference's dtype; image_is_f64 = 1) or fp32;
 * rendered [B,3,H,W] fp32; out_bgr [B,H,W,3] u8. */
int gfr_composite_bgr_u8(const void* image, int image_is_f64, const float* rendered, const uint8_t* mask,
                         int mask_batch_stride, uint8_t* out_bgr, int B, int H, int W, void* stream);

/* TESTB:595-597: range of -depth over all n elements, as two order-preserving uint32 keys {min, max} that
 * gfr_export_planes_u8 decodes (2 launches: init + reduce). */
int gfr_neg_depth_range(const float* depth, long long n, uint32_t* range_keys, void* stream);

/* TESTB:590-608: the five auxiliary images.  albedo/normals [B,3,H,W], depth [B,1,H,W], shadow/final_shading [B,H,W];
 * outputs (any may be NULL, its source may then be NULL too): out_shadow/out_depth/out_shading [B,H,W] u8,
 * out_albedo/out_normals [B,H,W,3] u8 BGR.  depth is written as 255 * (-d - min)/(max - min) * mask/255 with the
 * range from gfr_neg_depth_range; normals as 255 * (n + 1) / 2 * mask/255. */
int gfr_export_planes_u8(const float* albedo, const float* depth, const float* shadow, const float* final_shading,
                         const float* normals, const uint8_t* mask, int mask_batch_stride, const uint32_t* range_keys,
                         uint8_t* out_shadow, uint8_t* out_albedo, uint8_t* out_depth, uint8_t* out_shading,
                         uint8_t* out_normals, int B, int H, int W, void* stream);

/* fix_border_artifacts_CVPR2022.m:1-18: face = (mask >= 128) (MATLAB's uint8 `imread(mask)/255.0`), s = 7x7 box sum of
 * face with zero padding; pixels with 0 < s <= max_sum take the 3x3 median (zero padded, per channel) of img.
 * max_sum = 30 reproduces the shipped FFHQ_relighting_results/ PNGs on every pixel; 29 is the .m file as written
 * (`convolved < 30`).  img/out [B,H,W,C] u8, C <= 4, out != img. */
int gfr_border_median_fix_u8(const uint8_t* img, const uint8_t* mask, int mask_batch_stride, uint8_t* out, int B, int H,
                             int W, int C, int max_sum, void* stream);

/* MSE_MP.m:15-25 (the reference's Multi-PIE evaluation metric): per image b, sums[2b] = sum over pixels and channels of
 * (recon/255 * m - gt/255 * m)^2 and sums[2b+1] = sum of m, with m = mask/255, in fp64; the metric is
 * sums[2b] / (C * sums[2b+1]).  recon, gt [B,H,W,C] u8; mask [1|B,H,W] u8; sums [B,2] doubles (zeroed by the call). */
int gfr_masked_mse_u8(const uint8_t* recon, const uint8_t* gt, const uint8_t* mask, int mask_batch_stride, double* sums,
                      int B, int H, int W, int C, void* stream);

/* DSSIM_MP_RGB.m:15-27 on the device: sums[b] = {sum(ssimmap .* mask3), sum(mask3)} of MATLAB's `ssim(recon/255, gt/255)` map
 * (3-D 11x11x11 Gaussian window with replicate padding when window_3d != 0 — what MATLAB does for an M x N x 3 array — or the
 * per-plane 11x11 window), fp64 like MATLAB; DSSIM_b = (1 - sums[b][0] / sums[b][1]) / 2.  recon / gt uint8 [B,H,W,3]; mask
 * uint8 [H,W] (mask_batch_stride 0) or [B,H,W] (H*W), used as mask / 255. */
int gfr_masked_ssim_u8(const uint8_t* recon, const uint8_t* gt, const uint8_t* mask, int mask_batch_stride, double* sums,
                       int B, int H, int W, int window_3d, void* stream);

/* LPIPS (PerceptualSimilarity/lpips/lpips.py:112-144; test_network.py:41-48), the metric's own arithmetic, forward and backward.
 * gfr_lpips_layer_fwd: out[n,p] = sum_c w[c] (f0[n,c,p]/(|f0[n,:,p]|+1e-10) - f1[n,c,p]/(|f1[n,:,p]|+1e-10))^2 for NCHW feature
 * maps f0, f1 [N,C,HW] (normalize_tensor + squared difference + the learned 1x1 `lin` head).  gfr_lpips_layer_bwd: g_f0 / g_f1
 * (either may be NULL) from g_out [N,HW].  gfr_bilinear_up_add: out[N,H,W] += nn.Upsample(size=(H,W), mode='bilinear',
 * align_corners=False)(m[N,h,w]); _bwd: g_m (+=, zero it first) from g_out.  gfr_lpips_masked_sums: sums[n] = {sum(mask*map),
 * count(mask*map > 0)} (mask float [H,W] with stride 0, or [N,H,W]); the metric of test_network.py:45 is sums[0]/sums[1]. */
int gfr_lpips_layer_fwd(const float* f0, const float* f1, const float* w, float* out, int N, int C, int HW, void* stream);
int gfr_lpips_layer_bwd(const float* f0, const float* f1, const float* w, const float* g_out, float* g_f0, float* g_f1, int N,
                        int C, int HW, void* stream);
int gfr_bilinear_up_add(const float* m, float* out, int N, int h, int w, int H, int W, void* stream);
int gfr_bilinear_up_add_bwd(const float* g_out, float* g_m, int N, int h, int w, int H, int W, void* stream);
int gfr_lpips_masked_sums(const float* map, const float* mask, int mask_batch_stride, double* sums, int N, int H, int W,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GFR_B200_H_ */
