/* ipdm_b200 -- C ABI of the B200 (sm_100a) kernels behind the ALD MRI-reconstruction hot path of
 * 10258392511/InverseProblemWithDiffusionModel.
 *
 * The reference has no FFI layer of its own (it is pure Python/PyTorch on this path, SURVEY.md
 * section 8b); each entry point below names the reference call it replaces (file:line relative to
 * the reference root).  The Python host code in `inverseproblemwithdiffusionmodel_b200/` binds these
 * with ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *  - plain pointers, sizes and scalars only; every pointer is a DEVICE pointer unless it says host;
 *  - complex64 tensors are interleaved (re,im) float pairs ("c64"); `planar` state tensors are two
 *    float planes [2][...] (plane 0 = real part, plane 1 = imaginary part);
 *  - activations of the score network are NHWC; "f16" = IEEE half; accumulation is always fp32;
 *  - all calls are asynchronous on `stream` (a cudaStream_t passed as void*), never allocate, never
 *    synchronise, and are capturable in a CUDA graph;
 *  - return value 0 = ok, >0 = cudaError_t, <0 = IPDM_E_*; `ipdm_last_error()` gives the text.
 */
#ifndef IPDM_B200_H
#define IPDM_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPDM_E_BADARG (-1)      /* unsupported shape / null pointer / bad flag          */
#define IPDM_E_UNSUPPORTED (-2) /* transform length not a power of two in [8,512], ... */
#define IPDM_E_DRIVER (-3)      /* driver entry point (tensor-map encode) unavailable  */

int ipdm_abi_version(void);   /* 3 */
const char* ipdm_last_error(void);
/* number of kernels launched by this library in this process so far (bench.py's gpu_launches) */
unsigned long long ipdm_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Centred orthonormal 2-D DFT, SENSE forward / adjoint
 * ---------------------------------------------------------------------------------------------- */

/* Bytes of scratch the FFT/SENSE calls need for (ncoils, batch, H, W): one c64 image per coil image. */
size_t ipdm_sense_workspace_bytes(int ncoils, int batch, int H, int W);

/* out[c,b] = mask[b % mask_frames] * i2k(maps[c] * x[b]).
 *   x     c64 [batch][H][W]
 *   maps_re / maps_im  f32 [ncoils][H][W]; maps_im may be NULL (real maps); maps_re NULL => ncoils
 *         must be 1 and no coil multiply is done (plain `i2k_complex`)
 *   mask  u8 [mask_frames][W] (column mask, broadcast over H), NULL => no mask; mask_frames is 1 or
 *         any value with image b using row b % mask_frames (the reference's (24,1,1,W) mask)
 *   out   c64 [ncoils][batch][H][W]
 * Replaces: SENSE.__call__ (ncsn/linear_transforms/undersampling_fourier.py:140-150),
 *   RandomUndersamplingFourier.__call__ (:77-82), i2k_complex (ncsn/linear_transforms/__init__.py:36-45). */
int ipdm_sense_forward(const void* x, const float* maps_re, const float* maps_im, const uint8_t* mask,
                       int mask_frames, void* out, int ncoils, int batch, int H, int W, void* workspace,
                       void* stream);

/* out[b] = sum_c conj(maps[c]) * k2i(S[c,b]).  `mask` (may be NULL) is only a promise that S is zero
 * on unmasked columns so they can be skipped; the public conj_op passes NULL (quirk Q3: no mask).
 *   ssos != 0: out is f32 [batch][H][W] = sqrt(sum_c |k2i(S[c,b])|^2) instead (maps unused).
 * Replaces: SENSE.conj_op (:152-160), SENSE.SSOS (:162-170), RandomUndersamplingFourier.conj_op
 *   (:84-87), k2i_complex (ncsn/linear_transforms/__init__.py:48-57). */
int ipdm_sense_adjoint(const void* S, const float* maps_re, const float* maps_im, const uint8_t* mask,
                       int mask_frames, void* out, int ncoils, int batch, int H, int W, int ssos,
                       void* workspace, void* stream);

/* ---- mask plans: the column mask compiled once, on the host, into the tables the fast kernels read ------------
 * A plan belongs to one (mask, H, W) and to the device that was current at creation; it is immutable afterwards and
 * may be used from any number of streams / threads at once.  plan_create allocates device memory and copies
 * synchronously (call it outside graph capture); every *_plan operation below is asynchronous like the rest.
 *   mask_host  u8 [mask_frames][W] in HOST memory (non-zero = sampled column)
 * When every frame keeps few columns (<= 32 of W in {256, 512}, <= 16 of 128) and H is in {64,128,256,512} the plan
 * is "pruned": row transforms compute / consume the sampled columns only and the scratch is compact; otherwise the
 * *_plan calls run the general masked kernels with the plan's device copy of the mask.  Same results either way.
 * Replaces nothing in the reference (its mask is a tensor multiplied after a full FFT,
 * undersampling_fourier.py:77-82); it is the `*_plan_create/_destroy` cache SURVEY 8(b) asks for. */
int ipdm_sense_plan_create(const uint8_t* mask_host, int mask_frames, int H, int W, void** plan_out);
int ipdm_sense_plan_destroy(void* plan);
/* info[8] = { pruned (rows and columns), max sampled columns per frame, compact scratch row length, max active
 * 32-byte sectors per row, mask_frames, H, W, pruned_rows (the fused step only needs the row transforms) } */
int ipdm_sense_plan_info(const void* plan, int* info);
/* ipdm_sense_forward with the plan's mask (x c64 [batch][H][W] -> out c64 [ncoils][batch][H][W]). */
int ipdm_sense_forward_plan(const void* plan, const void* x, const float* maps_re, const float* maps_im, void* out,
                            int ncoils, int batch, void* workspace, void* stream);
/* ipdm_sense_adjoint on data that is zero off the plan's mask (columns off the mask are not read). */
int ipdm_sense_adjoint_plan(const void* plan, const void* S, const float* maps_re, const float* maps_im, void* out,
                            int ncoils, int batch, int ssos, void* workspace, void* stream);

/* k-space elementwise helpers for SingleCoil / projection (proximal_op.py:72-94,
 * undersampling_fourier.py:89-97):  mode 0: S *= 1/(1 + a*mask);
 * mode 1: S = a*Y + (1-a)*mask*S + (1-mask)*S   (Y = measured k-space, a = lamda);  mode 2: S *= mask.
 * S is c64 [batch][H][W]; image i uses mask row i % mask_frames. */
int ipdm_kspace_combine(void* S, const void* Y, const uint8_t* mask, int mask_frames, float a, int mode,
                        int batch, int H, int W, void* stream);

/* out = a + s * b  on c64 (n complex elements); used for z + alpha*k2i(y) and log_lh_grad. */
int ipdm_caxpy(void* out, const void* a, const void* b, float s, size_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Annealed Langevin update (+ fused SENSE L2-penalty proximal step)
 * ---------------------------------------------------------------------------------------------- */

/* Per-step scalars, either passed by value (sched == NULL) or read on the device from
 * sched[*cursor] so that one captured CUDA graph serves every noise level. */
typedef struct {
  float step;        /* step_lr * (sigma/sigma_L)^2                 (ALD_optimizers.py:101,217,438) */
  float noise_scale; /* sqrt(2*step)                                (:117,239,241)                  */
  float kappa;       /* 0.05 * step_lr*lr_scaled / (Nc*W)           (proximal_op.py:26-49) */
  float sigma;       /* sigma of this level (informational)                                          */
} ipdm_ald_scalars;

/* In-kernel noise: Philox4x32-10 with key = seed (^ *seed_dev) and counter = (position inside the chain, chain id,
 * step, tag), so the stream of a chain depends on its GLOBAL id only -- not on the rank, the number of ranks or its
 * slot in the batch (reference draw site ALD_optimizers.py:238-241, per-sample randn_like).  Passed by host
 * pointer; NULL = all zero. */
typedef struct {
  uint64_t seed;
  const uint64_t* seed_dev; /* device, may be NULL: XORed into `seed` when the kernel runs, so that a captured graph can
                               be replayed with a fresh stream (successive sampler calls must not repeat their noise) */
  uint32_t rng_step;        /* step counter; with a device schedule the kernel adds *cursor                         */
  int32_t chain_base;       /* chain id of sample i = chain_ids ? chain_ids[i] : chain_base + i                      */
  const int32_t* chain_ids; /* device int32 [samples], may be NULL                                                  */
  size_t chain_elems;       /* ipdm_langevin_update only: floats per sample (0: the whole buffer is one chain)       */
} ipdm_rng;

/* x <- x + step*grad + noise_scale*noise, n floats.  noise == NULL => in-kernel N(0,1) (see ipdm_rng; element pairs
 * (2p, 2p+1) of a sample share one Philox call).  step_per_sample (f32 [batch], may be NULL)
 * overrides `step`/`noise_scale` per sample (sde corrector, sde/sampling.py:320-322); per_sample_elems
 * = elements per sample.  x_mean (may be NULL) receives x + step*grad.
 * Replaces: ALDOptimizer.__call__ update (ncsn/models/ALD_optimizers.py:114-117),
 *   anneal_Langevin_dynamics (ncsn/models/__init__.py:58-61), AnnealedLangevinDynamics.update_fn. */
int ipdm_langevin_update(float* x, const float* grad, const float* noise, float* x_mean, size_t n,
                         const ipdm_ald_scalars* scalars_host, const ipdm_ald_scalars* sched, const int* cursor,
                         const float* step_per_sample, size_t per_sample_elems, const ipdm_rng* rng_host, void* stream);

/* One fused data-consistency ALD step on the planar state x f32 [2][batch][H][W]:
 *     z = x + step*grad + noise_scale*noise          (real and imaginary planes, independent noises)
 *     x <- z - kappa * (A^H A z - b)                 b = A^H y, planar f32 [2][batch][H][W]
 * with A the SENSE operator (maps, column mask).  Because the mask acts on W only, the H-axis
 * transform cancels in A^H A and the kernel needs only length-W row FFTs.
 * grad planar f32 [2][batch][H][W] (score of the real plane, score of the imaginary plane);
 * noise planar or NULL (in-kernel: image i of the batch is chain ipdm_rng.chain_ids[i]; the pixels w and w + W/2 of a
 * row share one Philox call).
 * Replaces: the loop body of ALDInvSegProximalRealImag.__call__ + post_processing
 *   (ALD_optimizers.py:238-241,288-327), ALD2DTime.spatial_step update + proximal_step (:442-449,
 *   543-554) with L2Penalty.__call__ (proximal_op.py:19-51). */
int ipdm_ald_sense_step(float* x, const float* grad, const float* noise, const float* b, const float* maps_re,
                        const float* maps_im, const uint8_t* mask, int mask_frames, int ncoils, int batch, int H,
                        int W, const ipdm_ald_scalars* scalars_host, const ipdm_ald_scalars* sched,
                        const int* cursor, const ipdm_rng* rng_host, void* stream);
/* The same step with the mask given as a plan (pruned row transforms when the plan allows); H = rows per image. */
int ipdm_ald_sense_step_plan(const void* plan, float* x, const float* grad, const float* noise, const float* b,
                             const float* maps_re, const float* maps_im, int ncoils, int batch, int H,
                             const ipdm_ald_scalars* scalars_host, const ipdm_ald_scalars* sched, const int* cursor,
                             const ipdm_rng* rng_host, void* stream);

/* labels[i] = *cursor / n_steps_each for i < batch, then (*cursor)++  (one tiny launch per step). */
int ipdm_ald_advance(int* cursor, int64_t* labels, int batch, int n_steps_each, void* stream);

/* Temporal total-variation step of ALD2DTime: x += -lamda * D^T sign(D x) along T (circular), on
 * each plane of planar x f32 [2][B][T][HW].  Replaces FiniteDiff.log_lh_grad
 *   (ncsn/linear_transforms/finite_diff.py:29-35) as used at ALD_optimizers.py:455-462. */
int ipdm_temporal_tv_step(float* x, int B, int T, size_t hw, float lamda, void* stream);

/* planar f32 [2][n]  <->  interleaved c64 [n] */
int ipdm_planar_to_c64(const float* planar, void* c64, size_t n, void* stream);
int ipdm_c64_to_planar(const void* c64, float* planar, size_t n, void* stream);

/* acc f64 [4][hw] += per-pixel (|x|, |x|^2, angle x, angle^2 x) summed over `chains` images of x c64
 * [chains][hw].  The cross-GPU sum is a plain all-reduce of acc.  Replaces the numpy mean/std of
 * helpers/visualizations.py:93-95,117-142. */
int ipdm_chain_stats_accumulate(const void* x, double* acc, int chains, size_t hw, void* stream);

/* sums f64 [images][4] = { sum (img-ref)^2, sum img^2, sum ref^2, sum |img-ref| } per image of n f32 values
 * (ref_images = 1: one reference for all, else = images).  MSE, MAE and the reference's NRMSE
 * (skimage normalized_root_mse(img, img_orig, "euclidean") = sqrt(sums[0] / sums[1]), normalised by its FIRST
 * argument) follow on the host.  Replaces helpers/metrics.py:47-74 (numpy / skimage on the CPU). */
int ipdm_image_sums(const float* img, const float* ref, double* sums, int images, int ref_images, size_t n, void* stream);

/* out f64 [images] = SUM over the (H-6)x(W-6) interior of the skimage-default SSIM map (7x7 uniform window, sample
 * covariance, K1 = 0.01, K2 = 0.03, C = (K * data_range)^2); divide by (H-6)*(W-6) for structural_similarity's
 * value.  data_range must be given (skimage infers it from the dtype; the reference leaves its version unpinned).
 * Replaces helpers/metrics.py:55-68. */
int ipdm_ssim(const float* img, const float* ref, double* out, int images, int ref_images, int H, int W,
              double data_range, void* stream);

/* ------------------------------------------------------------------------------------------------
 * NCSNv2 score network building blocks (NHWC)
 * ---------------------------------------------------------------------------------------------- */

/* Epilogue / operand description of one implicit-GEMM convolution. */
typedef struct {
  const void* in_f16;     /* A operand, f16 NHWC [N][H][W][Cin]                                         */
  const void* w_f16;      /* weights, f16 [Cout][taps][Cin] (taps = 9 row-major ky,kx, or 1)            */
  const float* bias;      /* f32 [Cout] or NULL                                                         */
  const float* residual;  /* f32 NHWC [N][H][W][Cout] or NULL: added before the stores                 */
  float* out_f32;         /* f32 NHWC or NULL: acc + bias + residual                                    */
  void* out_f16;          /* f16 NHWC or NULL: see flags                                                */
  double* stats;          /* f64 [N][Cout][2] or NULL: (zeroed, then) sum / sum-of-squares of the f32 result
                             per (n, channel), un-pivoted -- feeds InstanceNorm++ without a second pass.
                             Tile partials are fp32 in a fixed order; only their combination is an fp64
                             atomic, which keeps results run-to-run stable at fp32 precision             */
  int N, H, W, Cin, Cout;
  int taps;               /* 9 (3x3, zero padding = dilation), 1 (1x1) or 27 (3x3x3 over slice volumes: weights
                             [Cout][kx][kh][kw][Cin], the kx-plane reads slice x + (kx-1)*dilation, zero outside) */
  int dilation;
  int flags;              /* IPDM_CONV_* below                                                          */
  int slices;             /* 0/1: plain 2-D.  X > 1: the N images are N/X volumes of X consecutive slices
                             ([P][X][H][W][C]); `stats` rows are then per VOLUME ([N/X][Cout][2])            */
  int slice_shift;        /* output slice x reads input slice x + slice_shift of the same volume (zero outside
                             [0, X)): one kx-plane of a 3x3x3 convolution (layers3d.py:38-60) = one launch   */
  /* The residual stream in 16 bits (tensor-core kernels only): instead of `residual` / `out_f32`, the residual is read
   * from and the result f16(acc + bias + residual) written to f16 NHWC tensors -- 8 instead of 12 bytes per output
   * element for a residual convolution.  InstanceNorm++ sums still come from the fp32 values.  Exclusive with the f32
   * pair; the flags keep their meaning (IPDM_CONV_POOL2 shapes, IPDM_CONV_RES_ELU, IPDM_CONV_F16_PRE_RES). */
  const void* residual_f16;
  void* out_raw_f16;
  /* Operand exponent shift (block floating point per tensor; 0 is read as 1): the accumulator is multiplied by acc_scale
   * before bias / residual (= 1 / scale of the input operand), the f16 operand output by out_f16_scale after its ELU.
   * With power-of-two scales this is exact; the score network uses it to keep the operands of its un-normalised decoder
   * inside the f16 range when an activation would otherwise be clipped.  Scales of 1 cost nothing: the
   * straight-line epilogue paths are taken as before. */
  float acc_scale;
  float out_f16_scale;
  /* Optional sparsity hint for 3x3 convolutions, per 64-channel chunk of Cin (chunk i = channels [64 i, 64 i + 64), Cin <=
   * 1024): bit ky*3+kx set = the weight block (tap, chunk) may be non-zero; every block whose bit is clear MUST be zero in
   * w_f16.  0 = all nine taps.  The persistent halo kernel skips the cleared blocks (no weight traffic, no MMAs); the other
   * kernels ignore the hint (same result).  Used for ConvMeanPool in its 4x4 stride-2 form on space-to-depth operands:
   * 16 of 36 (tap, parity) blocks. */
  uint16_t tap_mask[16];
} ipdm_conv_desc;

#define IPDM_CONV_F16_ELU 1       /* out_f16 = f16(ELU(v)) instead of f16(v)                               */
#define IPDM_CONV_F16_PRE_RES 2   /* out_f16 is taken from acc+bias (before the residual add)            */
#define IPDM_CONV_RES_ELU 4       /* residual is ELU(residual[...]) (CRP entry, layers.py:77)            */
#define IPDM_CONV_POOL2 8         /* 2x2 mean-pool the result: residual/out tensors are [N][H/2][W/2][Cout]
                                     (ConvMeanPool, layers.py:309-313)                                    */

/* 3x3 (dilated) / 1x1 stride-1 convolution as a tcgen05 implicit GEMM: M = N*H*W pixels (8x16-pixel
 * tiles), N = Cout, K = taps*Cin; TMA-fed, fp32 accumulators in TMEM.  Cin % 64 == 0, Cout % 128 == 0.
 * Replaces nn.Conv2d inside ResidualBlock / RCUBlock / CRPBlock / MSFBlock / ConvMeanPool
 *   (ncsn/models/layers.py:28-60,62-83,112-134,165-184,291-313,401-456). */
int ipdm_conv_igemm(const ipdm_conv_desc* desc_host, void* stream);

/* Diagnostics knob (tests / profiling only; none of the settings changes results): key 1 = convolution kernel choice
 * (0 auto: persistent halo-tile kernel for 3x3 with dilation <= 2, per-tap tile kernel otherwise; 1 = always the per-tap
 * kernel).  key 2: 0 (default) / 1 = the halo kernel's L2 bulk prefetch of residual tiles off / on (A/B timing: on is
 * 5-10 % slower).  key 3: 0 (default) / 1 = launch the halo kernel with programmatic stream serialization (PDL; also env
 * IPDM_CONV_PDL; measured +0.4 %).  key 4: cudaLimitMaxL2FetchGranularity in bytes (32 / 64 / 128; measured: no effect on
 * the strided k-space reads of the masked adjoint).  key 5: 0 (default) / 1 = the pruned SENSE plan entry points split a
 * batch whose k-space is >= 512 MB into image sub-ranges and run row and column kernels of neighbouring sub-ranges side by
 * side on the plan's high-priority side stream (also env IPDM_SENSE_SPLIT; measured 2-7 % slower).  Timing experiments that produce wrong results exist only in builds
 * with -DIPDM_EXPERIMENTS (tools/exp_weights.py) and are refused otherwise. */
int ipdm_debug_option(int key, int value);

/* Same contract on CUDA cores, any Cin/Cout (used for narrow test nets and as the on-device
 * cross-check of the tensor-core kernel; not used by the product path when the igemm applies). */
int ipdm_conv_direct(const ipdm_conv_desc* desc_host, void* stream);

/* begin_conv: out f32 NHWC [N][H][W][Cout] = conv3x3(affine ? 2x-1 : x) + bias, x f32 [N][H][W]
 * (Cin == 1), w f32 [Cout][9].  Also fills `stats` like ipdm_conv_desc.stats when non-NULL.
 * Replaces ncsnv2.py:270-275. */
int ipdm_conv_first(const float* x, const float* w, const float* bias, float* out, double* stats, int N, int H,
                    int W, int Cout, int affine, void* stream);

/* end_conv: out f32 [N][H][W] = (conv3x3(in f16 NHWC [N][H][W][Cin], w f32 [9][Cin]) + bias) / sigmas[labels[n]].
 * workspace: f32 [N*H*W*9] scratch (per-pixel tap dot products).  Replaces ncsnv2.py:293-297. */
int ipdm_conv_last(const void* in_f16, const float* w, const float* bias, const float* sigmas,
                   const int64_t* labels, float* out, float* workspace, int N, int H, int W, int Cin, void* stream);

/* InstanceNorm++ (normalization.py:163-176).  stats f64 [N][C][2] = per-(n,c) sum and sum of squares of
 * (x - pivot), pivot = x[n,0,0,c] if `pivoted` else 0 (the conv epilogues produce the un-pivoted form);
 * apply: out_f16 = f16(ELU(gamma*(IN(x) + alpha*m_hat) + beta)), `stats_pivoted` as given to stats. */
int ipdm_instnorm_stats(const float* x, double* stats, int N, int HW, int C, int pivoted, void* stream);
int ipdm_instnorm_apply_elu(const float* x, const double* stats, int stats_pivoted, const float* alpha,
                            const float* gamma, const float* beta, void* out_f16, int N, int HW, int C,
                            void* stream);

/* The same three kernels on a 16-bit residual stream (the tensor-core path of the score network keeps its residual stream in
 * f16, ipdm_conv_desc.residual_f16 / out_raw_f16): begin_conv writing f16, InstanceNorm++ reading f16, the MSF upsample-
 * accumulate on f16 tensors.  Arithmetic is fp32; values are rounded once, when stored. */
int ipdm_conv_first_f16out(const float* x, const float* w, const float* bias, void* out_f16, double* stats, int N, int H,
                           int W, int Cout, int affine, void* stream);
int ipdm_instnorm_apply_elu_f16in(const void* x_f16, const double* stats, int stats_pivoted, const float* alpha,
                                  const float* gamma, const float* beta, void* out_f16, int N, int HW, int C, void* stream);
/* out_f16[i] = f16(scale * (elu ? ELU(x[i]) : x[i])): the cast of ipdm_act_to_f16 with an operand exponent shift. */
int ipdm_act_to_f16_scaled(const float* x, void* out_f16, size_t n, int elu, float scale, void* stream);
/* InstanceNorm++ + ELU with the f16 result written in space-to-depth layout [N][H/2][W/2][(y&1)*2 + (x&1)][C] (x: f32, or the
 * 16-bit stream when x_is_f16): the operand of ConvMeanPool evaluated as one 4x4 stride-2 convolution = a 3x3 convolution over
 * the space-to-depth tensor with 16 of its 36 (tap, parity) weight blocks non-zero (ipdm_conv_desc.tap_mask). */
int ipdm_instnorm_apply_elu_s2d(const void* x, int x_is_f16, const double* stats, int stats_pivoted, const float* alpha,
                                const float* gamma, const float* beta, void* out_f16, int N, int H, int W, int C, void* stream);
int ipdm_bilinear_add_f16(const void* src_f16, void* dst_f16, void* out_elu_f16, int N, int h, int w, int H, int W, int C,
                          int accumulate, void* stream);

/* Range audit of an f16 tensor (n values, 16-byte aligned): *max_abs = max(*max_abs, max |x|) and *n_saturated += the
 * number of values with |x| >= 65504 or not finite.  The f16 stores of this library saturate instead of overflowing
 * (the reference is fp32, ncsn/models/layers.py:37-60: no such limit), so a clipped activation is silent; the score
 * networks run this over every f16 buffer of a forward on request (`range_audit()`), which is how a checkpoint whose
 * activations leave the f16 range is found.  Both outputs are device memory the caller zeroes. */
int ipdm_f16_range_audit(const void* x_f16, size_t n, float* max_abs, unsigned long long* n_saturated, void* stream);

/* out_f16 = f16(elu ? ELU(x) : x), n elements (layers.py:12-13). */
int ipdm_act_to_f16(const float* x, void* out_f16, size_t n, int elu, void* stream);
/* 5x5 stride-1 max pool with -inf padding on f16 NHWC (CRPBlock, layers.py:69-70,80). */
int ipdm_maxpool5_f16(const void* in_f16, void* out_f16, int N, int H, int W, int C, void* stream);
/* dst f32 [N][H][W][C] (+)= bilinear_align_corners(src f32 [N][h][w][C]) (MSFBlock, layers.py:182-183);
 * out_elu_f16 (may be NULL) also receives f16(ELU(dst)) -- the CRP block that follows wants it. */
int ipdm_bilinear_add(const float* src, float* dst, void* out_elu_f16, int N, int h, int w, int H, int W, int C,
                      int accumulate, void* stream);
/* out f32 [N][H/2][W/2][C] = mean of the four stride-2 phases of in (+ add if non-NULL). */
int ipdm_meanpool2(const float* in, const float* add, float* out, int N, int H, int W, int C, void* stream);
/* 2x2 mean-pool of an f16 NHWC tensor (fp32 arithmetic): the operand of a pooled 1x1 shortcut, which the score network
 * evaluates as conv1x1(meanpool(x)) instead of meanpool(conv1x1(x)) (ConvMeanPool with kernel 1, layers.py:291-313). */
int ipdm_meanpool2_f16(const void* in_f16, void* out_f16, int N, int H, int W, int C, void* stream);
/* f32 [Cout][Cin][kh][kw] (PyTorch OIHW) -> f16 [Cout][kh*kw][Cin] weight repack for the igemm. */
int ipdm_pack_weights_f16(const float* w_oihw, void* w_f16, int Cout, int Cin, int taps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 3-D (patch x time) score network NCSN3DShallow: everything around its 3x3x3 convolutions, which are three
 * ipdm_conv_igemm launches (one kx-plane each, ipdm_conv_desc.slices / slice_shift).
 * Volume layout [P][X][T][Y][C]: P patches, X slices, each slice an NHWC image of H = T rows, W = Y columns.
 * Reference: ncsn/models/ncsn3d.py:123-224, layers3d.py.
 * ---------------------------------------------------------------------------------------------- */
/* slice axis of MaxPool3d(5, stride 1, padding 2) (layers3d.py:71); the (T, Y) plane is ipdm_maxpool5_f16. */
int ipdm_maxpool5_slices_f16(const void* in_f16, void* out_f16, int P, int X, size_t plane_elems, void* stream);
/* begin_conv (ncsn3d.py:137): out f32 [P][X][T][Y][Cout] = Conv3d(1 -> Cout, 3, padding 1)(affine ? 2x-1 : x) + bias;
 * x f32 [P][X][T][Y], w f32 [Cout][27] with taps ordered (kx, kt, ky). */
int ipdm_conv3d_first(const float* x, const float* w, const float* bias, float* out, int P, int X, int T, int Y, int Cout,
                      int affine, void* stream);
/* end_conv + noise-level division (ncsn3d.py:141,211-215): out f32 [P][X][T][Y] = (Conv3d(C -> 1, 3)(in) + bias) /
 * sigmas[labels[p]]; w f32 [27][C], taps ordered (kx, kt, ky). */
int ipdm_conv3d_last(const void* in_f16, const float* w, const float* bias, const float* sigmas, const int64_t* labels,
                     float* out, int P, int X, int T, int Y, int C, void* stream);
/* out f16 [NS][T2][Y][K*C]: out[..][t2][y][k*C + c] = in[..][stride*t2 + offset0 + k][y][c] (0 outside [0, T)): lays the
 * taps of conv_temporal_down / conv_temporal_up (ncsn3d.py:181-182) side by side for ONE 1x1 implicit GEMM. */
int ipdm_gather_t_f16(const void* in_f16, void* out_f16, size_t NS, int T, int T2, int Y, int C, int stride, int offset0, int K,
                      void* stream);
/* out_f32 [NS][2T][Y][C] (and f16(ELU(.)) if non-NULL): out[..][2m+ph][y][c] = in[..][m][y][ph*C + c] -- the two output
 * phases of the stride-2 ConvTranspose3d back onto the time axis. */
int ipdm_interleave_t(const float* in, float* out_f32, void* out_elu_f16, size_t NS, int T, int Y, int C, void* stream);
/* Patch fold of the 2D+time sampler: planar state f32 [2][B][T][H][W] <-> volumes f32 [2*B*(H/k)*(W/k)][k][T][k]
 * (`reshape_temporal_dim`, helpers/utils.py:330-359, in the device layout of NCSN3DShallow), with the optional random
 * roll (ALD_optimizers.py:466-470,495-499) as (shift_h, shift_w).  unfold = 0: state -> vol; 1: vol -> state. */
int ipdm_patch_fold(float* state, float* vol, int B, int T, int H, int W, int k, int shift_h, int shift_w, int unfold,
                    void* stream);

/* Same fold / unfold with the shifts of the current step read on the device: shifts = int32 [steps][2] (shift_h, shift_w per
 * ALD step, drawn by the host in the reference's np.random order, ALD_optimizers.py:466-470), cursor = the step counter
 * that ipdm_ald_advance increments -- so that a captured step graph can run the random-roll variant (if_random_shift). */
int ipdm_patch_fold_sched(float* state, float* vol, int B, int T, int H, int W, int k, const int* shifts, const int* cursor,
                          int unfold, void* stream);
/* out = a + (elu_b ? ELU(b) : b), n f32 elements (n % 4 == 0). */
int ipdm_add_act(const float* a, const float* b, float* out, size_t n, int elu_b, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IPDM_B200_H */
