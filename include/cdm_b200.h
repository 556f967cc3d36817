/*
 * cdm_b200.h -- C ABI of libcdm_b200.so, the sm_100a (B200) implementation of the
 * composed-score diffusion sampler hot path of mo-rsa24/composable_diffusion_models.
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; the boundary this
 * library replaces is the *loop body* of its sampler functions.
 * Every entry point below cites the reference lines whose work it performs.  The Python
 * shims in composable_diffusion_models_b200/ keep the reference's call signatures and
 * bind these symbols with ctypes (see INTEGRATION.md for the stub a maintainer adds).
 *
 * Conventions
 *  - plain C types only; every pointer named x / eps / z / logq / ... is a DEVICE pointer
 *    into caller-owned storage (torch tensors: tensor.data_ptr()); arrays documented as
 *    "host array" are read on the host during the call;
 *  - images are NCHW fp32, contiguous, exactly as the reference's tensors are;
 *  - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *    all work is enqueued on it, nothing synchronises;
 *  - return value: CDM_OK or a negative cdm_status; cdm_last_error() gives the message
 *    (thread-local);
 *  - there is no CPU fallback anywhere: without an sm_100 device every compute call fails
 *    with CDM_ERR_CUDA / CDM_ERR_UNSUPPORTED.
 */
#ifndef CDM_B200_H
#define CDM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDM_MAX_EXPERTS 8
#define CDM_ABI_VERSION 1

typedef enum {
  CDM_OK = 0,
  CDM_ERR_INVALID = -1,      /* bad argument (raises ValueError in the shims)            */
  CDM_ERR_CUDA = -2,         /* CUDA runtime / driver error                              */
  CDM_ERR_UNSUPPORTED = -3,  /* shape or mode this build does not implement              */
  CDM_ERR_NOT_READY = -4,    /* expert used before all parameters were set + finalized   */
  CDM_ERR_WORKSPACE = -5,    /* workspace too small                                      */
  CDM_ERR_KEY = -6           /* unknown / mis-sized state_dict key (KeyError)            */
} cdm_status;

/* CDM_PREC_FP32:  fp32 arithmetic on the CUDA cores (parity <= 1e-5; the slow path).
 * CDM_PREC_F16:   fp16 operands, fp32 accumulation on tcgen05 -- the throughput mode (TF32-class accuracy: both have an
 *                 11-bit significand).
 * CDM_PREC_F16X3: fp32-CLASS accuracy on tcgen05: every operand is split into high + low fp16 parts and each K step issues
 *                 three MMAs (lo*hi + hi*lo + hi*hi) into the fp32 accumulator; activations stay fp32 in HBM.  Parity <= 1e-5
 *                 like CDM_PREC_FP32, at tensor-core speed (the small-UNet expert; other experts: CDM_ERR_INVALID). */
typedef enum { CDM_PREC_FP32 = 0, CDM_PREC_F16 = 1, CDM_PREC_F16X3 = 4 } cdm_precision;

int cdm_abi_version(void);
/* First 32 bits of sha256(this header) as compiled into the library: the ctypes shim refuses a .so built from another header. */
unsigned cdm_abi_stamp(void);
/* Number of kernels this library has launched in this process (every launch is counted). */
long long cdm_launch_count(void);
/* Per-launch timing for measurement runs: cdm_prof_enable(1) starts bracketing every launch with CUDA
 * events on its own stream (and clears earlier records); cdm_prof_summary() synchronises, sums them by
 * kernel class and returns the number of classes written.  flops / bytes are the ALGORITHMIC work of the
 * launches (DESIGN.md states the formulas), so flops/ms and bytes/ms are the roofline numerators. */
typedef struct {
  char name[24];
  long long launches;
  double ms;
  double flops;
  double bytes;
} cdm_prof_entry;
int cdm_prof_enable(int on);
int cdm_prof_summary(cdm_prof_entry* out, int max_entries);
/* Print one line per recorded launch (class, shape tag, ms, TFLOP/s, GB/s) to stderr; returns the record count. */
int cdm_prof_dump(void);
const char* cdm_last_error(void);
/* 0 when `device` is an sm_100 GPU this library can run on. */
int cdm_device_check(int device);

/* Counter-based noise source used when a step's `z` pointer is NULL: the kernel draws
 * N(0,1) itself (Philox4x32-10 + Box-Muller) from (seed, step) so no noise tensor ever
 * touches HBM.  With z != NULL the injected tensor is used (parity mode). */
typedef struct {
  uint64_t seed;
  uint64_t step;
} cdm_rng;

/* ------------------------------------------------------------------------------------
 * Fused combine + update step kernels.
 * x, x_out: [B, C, HW] fp32 (x_out may alias x).  eps: host array of K device pointers,
 * expert k's prediction is [B, eps_channels[k], HW] with eps_channels[k] in {1, C}
 * (1 = broadcast over channels, the reference's `.repeat(1, 3, 1, 1)`).
 * ------------------------------------------------------------------------------------ */

/* Euler-Maruyama reverse-SDE step on a weighted SUM of K experts.
 * reference: mnist/compose_scores.py:37-46 (K=2), mnist/sample_image.py:33-39 (K=1),
 *            mnist/visualize_composition_latent.py:76-84 (2-D latents: C=1, HW=2).
 *   e  = sum_k w[k]*eps[k]                      (NOT normalised, as written there)
 *   x' = x + ( -(a*x - c*e)*dt + g*z )
 * a = dlog_alphadt(t), c = beta(t)/sigma(t), g = sqrt(2*xi*beta(t))*sqrt(dt); the shims
 * evaluate them in fp32 in the reference's operation order. */
int cdm_step_sde(const float* x, const float* const* eps, const int* eps_channels, const float* w, int K,
                 const float* z, const cdm_rng* rng, float a, float c, float dt, float g,
                 float* x_out, int B, int C, int HW, void* stream);

/* Deterministic DDIM step on a weighted MEAN of K experts, x0 clamped to [-1, 1].
 * reference: shapes/compose_images_ddim.py:52-68 (== shapes/compose_scores.py:52-74);
 *            K=1: shapes/train_image.py:43-85.
 *   e  = (sum_k w[k]*eps[k]) / wsum
 *   x0 = clamp((x - sigma_now*e)/alpha_now, -1, 1);  x' = alpha_next*x0 + sigma_next*e
 * gray_out (optional, [B,1,HW], C must be 3): 0.2989 R + 0.587 G + 0.114 B of x', i.e. the
 * Grayscale(x) the NEXT step's shape expert consumes (compose_images_ddim.py:47). */
int cdm_step_ddim(const float* x, const float* const* eps, const int* eps_channels, const float* w, int K,
                  float wsum, float alpha_now, float sigma_now, float alpha_next, float sigma_next,
                  float* x_out, float* gray_out, int B, int C, int HW, void* stream);

/* SuperDiff step: per-sample kappa from the running log-densities, DDPM ancestral update,
 * Ito log-density accumulation.  reference: src/diffusion/samplers.py:20-58.
 *   s_k   = -noise_k / sqrt_one_minus_ab
 *   kappa = softmax_k(temp*logq + bias) [operation 0 = OR] | softmax_k(-logq) [1 = AND] | 0.5 [2]
 *   mean  = (1/sqrt_alpha)*(x + beta*sum_k kappa_k s_k);  x' = mean + sqrt_post_var*z   (z NULL and
 *           rng NULL: last step, x' = mean)
 *   logq_k += sum(dx*s_k) + (-0.5*beta*D + sum((-0.5*beta*x - 0.5*beta*s_k)*s_k))*dtau
 * logq: [B, K] fp32, updated in place.  kappa_out: optional [B, K]. */
int cdm_step_ddpm_logq(const float* x, const float* const* noise_pred, int K, const float* z, const cdm_rng* rng,
                       float* logq, int operation, float temp, float bias, float sqrt_one_minus_ab, float beta,
                       float sqrt_alpha, float sqrt_post_var, float dtau, float* x_out, float* kappa_out,
                       int B, int C, int HW, void* stream);

/* Ito / kappa composition of two experts on the probability-flow ODE.
 * reference: shapes/compose_images_ito.py:66-85,119-135; shapes/compose_images_ito_2.py:127-149;
 *            2-D latents: shapes/visualize_composition_latent_ito.py:60-78,125-144 (mode 0),
 *            shapes/visualize_composition_latent_ito_2.py:39-52,99-116 (mode 1).
 * mode 0 (score space):  s_k = -eps_k/sigma, kappa = (-div1/sigma + div2/sigma + sum s1(s1-s2)) /
 *                        (sum (s1-s2)^2 + den_eps);  x' = x - (a*x - coef*(s2 + kappa(s1-s2)))*dt
 * mode 1 (eps space, clipped): kappa = clip((-sigma(div1-div2) + sum e1(e1-e2))/(sum (e1-e2)^2 + den_eps),
 *                        lo, hi);  x' = x - (a*x + coef*(e2 + kappa(e1-e2)))*dt
 * mode 2: kappa as mode 0, update in eps space with the sign of the "stable" latent script:
 *                        x' = x - (a*x - coef*(-(e2) + kappa(-(e1) + e2)))*dt
 * div1 is multiplied by div1_scale first (3.0 in compose_images_ito.py:113). */
int cdm_step_ode_kappa(const float* x, const float* eps1, int eps1_channels, const float* eps2,
                       const float* div1, const float* div2, float div1_scale, int mode, float sigma,
                       float a, float coef, float dt, float den_eps, float clip_lo, float clip_hi,
                       float* x_out, float* kappa_out, int B, int C, int HW, void* stream);

/* Ito / kappa composition of K = 2 .. 4 experts on the probability-flow ODE (BASELINE config 4 names four experts; the
 * reference writes the closed form for two).  K-expert semantics: equal d log q_k / dt for every expert + sum(kappa) = 1,
 * the linear system of SuperDiff Prop. 6 that src/composing_conditional_diffusion_on_shape_and_color_6_1.py:374-396 solves
 * for the SDE.  With s_k = -eps_k/sigma, div s_k = -div_scale[k]*div[k]/sigma, d_r = s_r - s_{r+1}, e_j = s_j - s_{K-1}:
 *   sum_{j<K-1} (<d_r, e_j> + [r == j] den_eps) kappa_j = div s_r - div s_{r+1} + <d_r, s_r + s_{r+1} - s_{K-1}>,  r < K-1
 *   kappa_{K-1} = 1 - sum_j kappa_j;   x' = x - (a*x - coef*(s_{K-1} + sum_j kappa_j e_j))*dt
 * solved per sample in the kernel (partial pivoting; singular -> 1/K each).  K = 2 IS get_kappa
 * (shapes/compose_images_ito.py:66-85) and is routed to cdm_step_ode_kappa mode 0, bit for bit.
 * eps: host array of K device pointers, eps_channels[k] in {1, C}; div: host array of K device pointers [B]; div_scale:
 * host array of K floats (3 for a 1-channel expert whose output is repeated over RGB, compose_images_ito.py:113) or NULL.
 * kappa_out: optional [B, K] (written for K >= 3). */
int cdm_step_ode_kappa_k(const float* x, const float* const* eps, const int* eps_channels, const float* const* div,
                         const float* div_scale, int K, float sigma, float a, float coef, float dt, float den_eps, float* x_out,
                         float* kappa_out, int B, int C, int HW, void* stream);

/* Guidance-sum / weighted-mean composition with a discrete-time update.
 * combine 0: e = eps[0] + sum_{k>=1} w[k]*(eps[k] - eps[0])          (eps[0] = unconditional)
 *            reference: src/compositional_diffusion_with_cross_attention.py:294-299,
 *                       src/composing_conditional_diffusion_on_shape_and_color_5.py:313-339
 * combine 1: e = (sum_k w[k]*eps[k]) / wsum
 *            reference: src/composing_conditional_diffusion_on_shape_and_color.py:347-352, _4.py:391-392
 * update 0 (x0 form, cross_attention.py:302-313): x' = c0*e + c1*e, c0 = sqrt(ab_prev), c1 = sqrt(1-ab_prev)
 * update 1 (ancestral, shape_and_color.py:354-366): x' = c0*(x - c1*e/c2) + c3*z,
 *            c0 = sqrt_recip_alpha, c1 = beta, c2 = sqrt_one_minus_ab, c3 = sqrt(post_var) (z NULL & rng NULL: no noise) */
int cdm_step_cfg(const float* x, const float* const* eps, const float* w, int K, float wsum, int combine,
                 int update, float c0, float c1, float c2, float c3, const float* z, const cdm_rng* rng,
                 float* x_out, int B, int C, int HW, void* stream);

/* LayoutDiff spatial-mask composition with the clamped-x0 posterior-mean DDPM step.
 * reference: src/composing_colored_digit_to_simulate_overlaying.py:84-119 (LayoutDiff.sample loop body)
 *   e = sum_k eps[k] * masks[k][pixel]        masks: [K][HW] float64 DEVICE array, broadcast over batch and channels
 *   x0 = clamp((x - s1m*e)/sab, -1, 1);  x' = c0*x0 + c1*x + spv*z         (z NULL & rng NULL: last step, no noise)
 *   s1m = sqrt(1-ab_t), sab = sqrt(ab_t), c0 = sqrt(ab_prev)*beta_t/(1-ab_t), c1 = sqrt(alpha_t)*(1-ab_prev)/(1-ab_t),
 *   spv = sqrt(posterior_variance_t).  masks_f64 != 0: accumulate e in double and round to float after every expert,
 *   exactly what torch does for the reference's float64 masks; 0: float arithmetic (float32 masks). */
int cdm_step_layout(const float* x, const float* const* eps, int K, const double* masks, int masks_f64, float s1m,
                    float sab, float c0, float c1, float spv, const float* z, const cdm_rng* rng, float* x_out, int B,
                    int C, int HW, void* stream);

/* Grayscale(num_output_channels=1) of an RGB batch; reference: shapes/compose_images_ddim.py:47. */
/* SuperDiff step with kappa from the K x K linear system ("stochastic AND") or the softmax
 * (OR), K <= 4 experts, one launch, no host syncs.
 * reference (K = 2, batch 1, with .item() round trips): src/composing_conditional_diffusion_on_shape_and_color_6_1.py:352-428
 *   s_k = -noise_pred_k / som;  f = f_coef * x;  div_f = f_coef * D
 *   mode 0 (OR):  kappa = softmax(temp * logq + bias)
 *   mode 1 (AND): a[r][c] = d_tau <-f + g_sq/2 s_c, s_r>,  b[r] = d_tau (div_f + <f - g_sq/2 s_r, s_r>) + <sqrt(g_sq) dW, s_r>,
 *                 dW = dw * sqrt(d_tau) (dw: unit-normal draws, [B, C, HW]); rows r < K-1: (a[r] - a[r+1]) kappa = b[r+1] - b[r]
 *                 (+ bias on row 0), last row sum(kappa) = 1; clamp to [0, 1], renormalise; singular system -> 1/K each
 *   x' = sqrt_recip_alpha * (x - beta * (-(sum_k kappa_k s_k) * som) / som) + sqrt_post_var * z   (z NULL & rng NULL: none)
 *   logq_k += <x' - x, s_k> + d_tau (div_f + <f - g_sq/2 s_k, s_k>)
 * f_coef, g_sq: get_forward_process_params (:296-327), computed on the host.  kappa_out: optional [B, K]. */
int cdm_step_superdiff_solve(const float* x, const float* const* noise_pred, int K, int mode, float temp, float bias,
                             float som, float beta, float sqrt_recip_alpha, float sqrt_post_var, float d_tau, float f_coef,
                             float g_sq, const float* dw, const float* z, const cdm_rng* rng, float* logq, float* x_out,
                             float* kappa_out, int B, int C, int HW, void* stream);

/* PCA inverse transform of sampled latents back to pixel space:
 * out[b, :] = z[b, :] @ components + mean.   z [B, L] (L <= 8), components [L, D], mean [D], out [B, D]; D % 4 == 0.
 * reference: mnist/sample_latent.py:88-89 (np.dot(final_latents, pca_components) + pca_mean),
 *            shapes/visualize_composition_latent_ito.py:188 (pca.inverse_transform). */
int cdm_latent_decode(const float* z, const float* components, const float* mean, float* out, int B, int L, int D,
                      void* stream);
int cdm_grayscale(const float* x, float* gray, int B, int HW, void* stream);

/* Fill z[n] with the N(0,1) stream the step kernels would draw for (rng, n) -- lets tests
 * replay in-kernel noise into a checker. */
int cdm_fill_normal(float* z, int64_t n, const cdm_rng* rng, void* stream);

/* ------------------------------------------------------------------------------------
 * Expert: the small GroupNorm ResBlock UNet (rows a3-a5).
 * reference: mnist/models/unet_small.py:47-92, shapes/models/unet_small.py:53-120.
 * Parameters are set by their state_dict key (row a14), fp32, torch's native layouts.
 * ------------------------------------------------------------------------------------ */
typedef struct cdm_unet cdm_unet;

typedef struct {
  int in_channels;   /* 1 or 3 */
  int base_dim;      /* 64 */
  int time_emb_dim;  /* 256 */
  int num_classes;   /* 0 = unconditional (mnist), >0 = label_emb rows (shapes) */
} cdm_unet_config;

int cdm_unet_create(const cdm_unet_config* cfg, int device, cdm_unet** out);
void cdm_unet_destroy(cdm_unet* m);
/* host_data: HOST fp32, numel elements, in the tensor's natural (contiguous) order. */
int cdm_unet_set_param(cdm_unet* m, const char* key, const float* host_data, int64_t numel);
/* Verifies every key of the reference's state_dict was provided, packs weights for both precisions. */
int cdm_unet_finalize(cdm_unet* m);
/* Number of state_dict keys this config expects, and the i-th key / its element count. */
int cdm_unet_num_params(const cdm_unet* m);
const char* cdm_unet_param_key(const cdm_unet* m, int i, int64_t* numel);
/* Experts process the batch in micro-batches of this many samples (workspace is sized for one micro-batch;
 * default 4096, env CDM_MICROBATCH; <= 0 restores the default).  Results do not depend on it. */
int cdm_set_microbatch(int samples);
/* Test hook: kernel-path selection, so the parity tests can drive EVERY convolution kernel through the same expert graph
 * (also env CDM_MICROBATCH / CDM_CONV_HALO / CDM_FUSE_GN / CDM_CONV_STACK / CDM_FUSE_PROJ): "microbatch" (samples),
 * "conv_halo" (1 = halo-tile tcgen05 kernel where it applies, 0 = shifted-box kernel everywhere), "fuse_gn" (GroupNorm+SiLU
 * fused into the halo kernel's prologue or a separate pass), "conv_stack" (0/1/2: stacked-tap kernel never / where supported /
 * where it wins), "fuse_proj" (out_conv fused into the last conv's epilogue), "grouped" (K-expert grouped conv launches in the
 * chain entries / cdm_unet_forward_grouped, or back-to-back forwards), "init_conv_tc" (the fp16 graphs' init conv on tcgen05 or on
 * CUDA cores; also env CDM_INIT_CONV_TC), "conv_pair" / "conv_pair64" / "stack_pair" (CTA-pair instances), "conv_scheme_c", "pdl".
 * -1 = default.  Every setting computes the same
 * function; none of them is a debug mode (role-wait timers and ablation switches exist only in -DCDM_INSTRUMENT builds). */
int cdm_set_option(const char* name, int value);
size_t cdm_unet_workspace_bytes(const cdm_unet* m, int B, int img_size, int precision);
/* eps = UNet(x, t, y).  x: [B, in_channels, S, S]; t: [B] fp32; y: [B] int64 or NULL (must be non-NULL
 * when num_classes > 0: CDM_ERR_INVALID, the reference's ValueError); eps: [B, in_channels, S, S].
 * precision CDM_PREC_FP32: fp32 CUDA-core path (parity <= 1e-5); CDM_PREC_F16: tcgen05/TMA implicit-GEMM
 * convolutions, fp16 operands, fp32 accumulation; CDM_PREC_F16X3: three-term split-fp16 tcgen05 convolutions
 * (parity <= 1e-5 at tensor-core speed). */
int cdm_unet_forward(cdm_unet* m, const float* x, const float* t, const int64_t* y, float* eps, int B,
                     int img_size, int precision, void* workspace, size_t workspace_bytes, void* stream);
/* K expert forwards with ONE grouped launch per convolution (north_star: "all K experts are batched into one grouped launch
 * per timestep"; reference call sites: the back-to-back model calls of mnist/compose_scores.py:33-34 and
 * shapes/compose_images_ddim.py:49-50).  gridDim.y of every tcgen05 convolution launch indexes the expert: each expert keeps
 * its own weights, tensor maps and workspace slice and gets num_sms / K persistent CTAs, so small per-GPU batches (strong
 * scaling) fill the machine with K x the tiles and half the launches; results are bit-identical to K cdm_unet_forward calls.
 * Grouping applies to 2 .. 4 fp16 experts of one architecture (channel counts of the image may differ, e.g. the 1-channel
 * shape expert next to the 3-channel colour expert); anything else runs back to back.  x / y / eps: HOST arrays of K device
 * pointers (x[k] may alias).  The sampler chain entries (cdm_unet_sample_sde / _ddim) use it automatically. */
size_t cdm_unet_forward_grouped_workspace_bytes(cdm_unet* const* experts, int K, int B, int img_size, int precision);
int cdm_unet_forward_grouped(cdm_unet* const* experts, int K, const float* const* x, const float* t, const int64_t* const* y,
                             float* const* eps, int B, int img_size, int precision, void* workspace, size_t workspace_bytes,
                             void* stream);

/* Whole reverse-SDE chain for K UNet experts in ONE host call: per step K expert forwards + the fused combine/update
 * launch, all enqueued on `stream` without returning to the caller.  reference: the loop of mnist/compose_scores.py:26-46
 * (K = 2) and mnist/sample_image.py:24-39 (K = 1).  Inside the chain every sample shares t, so the time embedding is
 * computed as ONE row per step and expert; results are bit-identical to calling
 * cdm_unet_forward + cdm_step_sde per step.
 *   x: [B, C, S, S] in/out.  y: NULL, or K device label arrays [B] (entry NULL = unconditional expert); y_uniform != 0
 *   promises that each array holds one repeated label.  z: [n_steps, B, C, S, S] injected noise, or NULL with rng (step
 *   i draws from (rng->seed, rng->step + i)).  step_coef_host: HOST [n_steps, 4] rows {t, a, c, g} as cdm_step_sde. */
size_t cdm_unet_sample_workspace_bytes(cdm_unet* const* experts, int K, int B, int img_size, int precision);
int cdm_unet_sample_sde(cdm_unet* const* experts, const float* w, int K, float* x, const int64_t* const* y, int y_uniform,
                        const float* z, const cdm_rng* rng, const float* step_coef_host, int n_steps, float dt, int B,
                        int img_size, int precision, void* workspace, size_t workspace_bytes, void* stream);

/* Whole DDIM chain for K UNet experts in ONE host call (reference: the loop of shapes/compose_images_ddim.py:39-68; K = 1:
 * shapes/train_image.py:60-85).  x: [B, C, S, S] in/out.  An expert with ONE input channel under an RGB state (C = 3) reads
 * Grayscale(x) (:47) -- taken once before the loop, afterwards emitted by the step kernel -- and its prediction is broadcast
 * over the channels.  w / wsum: the weighted MEAN of cdm_step_ddim.  step_coef_host: HOST [n_steps + 1, 3] rows
 * {t, alpha(t), sigma(t)} at the grid points (fp32, evaluated by the shim in the reference's operation order).  y, y_uniform:
 * as cdm_unet_sample_sde.  Bit-identical to calling cdm_unet_forward + cdm_step_ddim per step. */
size_t cdm_unet_sample_ddim_workspace_bytes(cdm_unet* const* experts, int K, int B, int C, int img_size, int precision);
int cdm_unet_sample_ddim(cdm_unet* const* experts, const float* w, int K, float wsum, float* x, const int64_t* const* y,
                         int y_uniform, const float* step_coef_host, int n_steps, int B, int C, int img_size, int precision,
                         void* workspace, size_t workspace_bytes, void* stream);

/* As cdm_unet_forward (precision CDM_PREC_FP32: CUDA-core convs; CDM_PREC_F16: primal and tangent convs on tcgen05), plus the bilinear form vjv[b] = <v_out_b, (d eps_b / d x_b) v_in_b> by
 * forward-mode differentiation of the same kernels: the tangent v_in is pushed through the network next to
 * the primal.  v_out == NULL means v_out = v_in, which is the Hutchinson estimator v^T J v of
 * shapes/compose_images_ito.py:46-63 (the reference takes the VJP with autograd and dots it with v: the same
 * scalar).  Separate v_in / v_out express the divergence through Grayscale of compose_images_ito_2.py:46-69:
 * v_in = Grayscale(v), v_out = sum over channels of v.  Workspace: cdm_unet_jvp_workspace_bytes. */
size_t cdm_unet_jvp_workspace_bytes(const cdm_unet* m, int B, int img_size, int precision);
int cdm_unet_forward_jvp(cdm_unet* m, const float* x, const float* t, const int64_t* y, const float* v_in,
                         const float* v_out, float* eps, float* vjv, int B, int img_size, int precision, void* workspace,
                         size_t workspace_bytes, void* stream);
/* Whole Ito / kappa probability-flow chain for K = 2 .. 4 UNet experts on an RGB state in ONE host call (reference: the
 * loop of shapes/compose_images_ito.py:101-135 and compose_images_ito_2.py:118-149, two experts; K > 2: cdm_step_ode_kappa_k).
 * Per step: Grayscale(x); per expert a primal + tangent forward (cdm_unet_forward_jvp) with its Hutchinson probe; the fused
 * kappa + Euler step.  1-channel experts read Grayscale(x); variant 0 ("beta"): their divergence is taken w.r.t. the grayscale
 * input (1-channel probe) and scaled by 3 (:113); variant 1 ("g2"): through Grayscale w.r.t. the RGB input (3-channel probe,
 * v_in = Grayscale(v), v_out = sum over channels).  With K = 2 the last expert must be the 3-channel one (the reference's
 * order).  probes: HOST array of K device pointers [n_steps, B, 1 or 3, S, S] (NULL entries / NULL array: drawn in the
 * library from rng, step i expert k uses (seed, step + i*K + k)).  step_coef_host: HOST [n_steps, 4] rows
 * {t, sigma(t), dlog_alphadt(t), coef} (coef = beta/2 or g2/2). */
size_t cdm_unet_sample_ito_workspace_bytes(cdm_unet* const* experts, int K, int B, int img_size, int precision);
int cdm_unet_sample_ito(cdm_unet* const* experts, int K, float* x, const int64_t* const* y, int variant,
                        const float* const* probes, const cdm_rng* rng, const float* step_coef_host, int n_steps, float dt, int B,
                        int img_size, int precision, void* workspace, size_t workspace_bytes, void* stream);

/* Debug/test hook: copy a named intermediate of the LAST forward ("x0","d1","d2","b1","u1","u2") to `out`
 * as NCHW fp32. */
int cdm_unet_debug_read(cdm_unet* m, const char* name, float* out, int B, int img_size, void* stream);

/* ------------------------------------------------------------------------------------
 * Expert: the 2-D latent MLP (row a6).  reference: mnist/models/mlp_2d.py:5-20.
 * ------------------------------------------------------------------------------------ */
typedef struct cdm_mlp cdm_mlp;
int cdm_mlp_create(int num_hid, int num_out, int device, cdm_mlp** out);
void cdm_mlp_destroy(cdm_mlp* m);
int cdm_mlp_set_param(cdm_mlp* m, const char* key, const float* host_data, int64_t numel);
int cdm_mlp_finalize(cdm_mlp* m);
/* eps[B, num_out] = MLP(t[B], x[B, num_out]) */
int cdm_mlp_forward(cdm_mlp* m, const float* t, const float* x, float* eps, int B, void* stream);
/* eps and vjv[b] = v_b^T (d eps_b / d x_b) v_b by forward-mode differentiation.
 * reference: vector_field, shapes/visualize_composition_latent_ito.py:47-60 (autograd VJP dotted with v). */
int cdm_mlp_forward_jvp(cdm_mlp* m, const float* t, const float* x, const float* v, float* eps, float* vjv, int B,
                        void* stream);
/* Whole reverse-SDE chain for K latent experts in ONE persistent launch: every sample's n_steps-step
 * chain runs in registers/shared memory.  reference: mnist/visualize_composition_latent.py:63-87.
 * x: [B, num_out] in/out.  z: [n_steps, B, num_out] injected noise or NULL (in-kernel rng).
 * step_coef: DEVICE [n_steps, 4] fp32 rows {t, a, c, g} (same meaning as cdm_step_sde), evaluated by the
 * shim in the reference's fp32 operation order. */
int cdm_mlp_sample_sde(cdm_mlp* const* experts, const float* w, int K, float* x, const float* z,
                       const cdm_rng* rng, const float* step_coef, int n_steps, float dt, int B, void* stream);
/* The same chain with the two 256x256 hidden layers on tcgen05 (fp16 operands, fp32 accumulation; everything else fp32):
 * K <= 2 experts of width 256 with 2 outputs (the reference's MLP(num_hid=256, num_out=2)), else CDM_ERR_UNSUPPORTED --
 * callers then use cdm_mlp_sample_sde.  256 samples per CTA, weights streamed from L2 by TMA. */
int cdm_mlp_sample_sde_tc(cdm_mlp* const* experts, const float* w, int K, float* x, const float* z,
                          const cdm_rng* rng, const float* step_coef, int n_steps, float dt, int B, void* stream);

/* ------------------------------------------------------------------------------------
 * Expert: ColoredMNISTScoreModel / ScoreModel, the BatchNorm UNet of the SuperDiff scripts (row a8).
 * reference: src/models/compose_grayscale_object_and_color.py:35-112 (== src/models/composing_colored_digit_...py).
 * Keys are the module's state_dict keys incl. the BatchNorm buffers (num_batches_tracked is accepted and
 * ignored); eval-mode semantics (running statistics).  fp32 path (cdm_score_forward) and fp16 tensor-core path (cdm_score_forward_prec).
 * ------------------------------------------------------------------------------------ */
typedef struct cdm_score cdm_score;
int cdm_score_create(int in_channels, int time_emb_dim, int device, cdm_score** out);
void cdm_score_destroy(cdm_score* m);
int cdm_score_set_param(cdm_score* m, const char* key, const float* host_data, int64_t numel);
int cdm_score_finalize(cdm_score* m);
size_t cdm_score_workspace_bytes(const cdm_score* m, int B, int img_size);
/* eps = model(x, t): x [B, in_channels, S, S], t [B] fp32 (the reference passes timestep indices as floats). */
int cdm_score_forward(cdm_score* m, const float* x, const float* t, float* eps, int B, int img_size, void* workspace,
                      size_t workspace_bytes, void* stream);
/* The same with a precision: CDM_PREC_FP32 = cdm_score_forward; CDM_PREC_F16 = every 3x3 conv, the k4-s2 strided "transform"
 * convs (read through the four input-parity views, no im2col) and the k4-s2 transposed up-convs (four output-parity classes
 * of 2x2 taps) on tcgen05, with bias -> ReLU -> BatchNorm affine -> time bias fused into the conv epilogues; 32-channel
 * tensors are kept zero-padded to 64 channels. */
size_t cdm_score_workspace_bytes_prec(const cdm_score* m, int B, int img_size, int precision);
int cdm_score_forward_prec(cdm_score* m, const float* x, const float* t, float* eps, int B, int img_size, int precision,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Whole SuperDiff chain over K score UNets in ONE host call (reference: the loop of src/diffusion/samplers.py:19-58): per
 * step K forwards + cdm_step_ddpm_logq.  x [B, C, S, S] and logq [B, K] in/out (logq starts at zero).  The call runs n_steps
 * consecutive steps of the chain (callers that stage injected noise in chunks call it once per chunk); ends_chain != 0 says
 * that its last step is the chain's last, which adds no noise.  z: injected noise [n_steps (- 1 when ends_chain), B, C, S, S] or
 * NULL with rng (step i draws (seed, step + i)).  step_coef_host: HOST [n_steps, 5] rows
 * {t_idx, sqrt(1 - alphas_cumprod), beta, sqrt(alpha), sqrt(posterior_variance)} in sampling order. */
size_t cdm_score_sample_superdiff_workspace_bytes(cdm_score* const* experts, int K, int B, int img_size, int precision);
int cdm_score_sample_superdiff(cdm_score* const* experts, int K, float* x, float* logq, int operation, float temp, float bias,
                               const float* z, const cdm_rng* rng, const float* step_coef_host, int n_steps, int ends_chain,
                               float dtau, int B, int img_size, int precision, void* workspace, size_t workspace_bytes,
                               void* stream);

/* ------------------------------------------------------------------------------------
 * BetaVAE decoder: the image-space epilogue of the latent samplers.
 * reference: src/4.3 best_of_both_worlds_3.py:95-126 (BetaVAE.decoder_input, .decoder, .decode), called as
 * `vae_decoder(z)` at the end of sample_composed_latent (:262-293).  Keys are the BetaVAE state_dict keys
 * (decoder_input.*, decoder.0.*, decoder.3.*, decoder.5.*, decoder.7.*); encoder.*, fc_mu.*, fc_log_var.* are accepted
 * and ignored.  fp32 path.
 * ------------------------------------------------------------------------------------ */
typedef struct cdm_vae_decoder cdm_vae_decoder;
int cdm_vae_decoder_create(int latent_dims, int device, cdm_vae_decoder** out);
void cdm_vae_decoder_destroy(cdm_vae_decoder* m);
int cdm_vae_decoder_set_param(cdm_vae_decoder* m, const char* key, const float* host_data, int64_t numel);
int cdm_vae_decoder_finalize(cdm_vae_decoder* m);
size_t cdm_vae_decoder_workspace_bytes(const cdm_vae_decoder* m, int B);
/* images [B,3,32,32] in (0,1) = decode(z [B,latent_dims]) */
int cdm_vae_decode(cdm_vae_decoder* m, const float* z, float* images, int B, void* workspace, size_t workspace_bytes,
                   void* stream);
/* torchvision.utils.save_image's quantisation of a [0,1] image: out = uint8(clamp(x * 255 + 0.5, 0, 255)), bit-exact
 * (reference call sites: mnist/viz.py, shapes/viz.py, src/4.3 best_of_both_worlds_3.py:340-352 via save_image). */
int cdm_quantize_u8(const float* x, uint8_t* out, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------
 * Expert: SimpleUnet, the 62 M-parameter GroupNorm UNet of the classifier-free-guidance SuperDiff scripts (SURVEY.md
 * section 8(f) row 4).  reference: src/composing_conditional_diffusion_on_shape_and_color_6.py:145-221 (same classes in
 * _6_1 / _7).  Keys are the module's state_dict keys (time_mlp.1.*, label_emb.weight, conv0.*, downs.i.*, ups.i.*,
 * output.*).  fp32 path; img_size % 16 == 0.
 * ------------------------------------------------------------------------------------ */
typedef struct cdm_simple_unet cdm_simple_unet;
int cdm_simple_unet_create(int num_classes, int device, cdm_simple_unet** out);
void cdm_simple_unet_destroy(cdm_simple_unet* m);
int cdm_simple_unet_set_param(cdm_simple_unet* m, const char* key, const float* host_data, int64_t numel);
int cdm_simple_unet_finalize(cdm_simple_unet* m);
size_t cdm_simple_unet_workspace_bytes(const cdm_simple_unet* m, int B, int img_size);
/* eps = model(x, timestep, y): x [B,3,S,S]; t [B] fp32 (timestep indices as floats); y [B] int64 in [0, num_classes]
 * (num_classes = the null token of classifier-free guidance). */
int cdm_simple_unet_forward(cdm_simple_unet* m, const float* x, const float* t, const int64_t* y, float* eps, int B,
                            int img_size, void* workspace, size_t workspace_bytes, void* stream);
/* The same with a precision: CDM_PREC_F16 runs every 3x3, k4-s2 strided and k4-s2 transposed conv on tcgen05 (bias + ReLU +
 * GroupNorm statistics in the conv epilogues; one elementwise pass per GroupNorm for its affine and the time bias). */
size_t cdm_simple_unet_workspace_bytes_prec(const cdm_simple_unet* m, int B, int img_size, int precision);
int cdm_simple_unet_forward_prec(cdm_simple_unet* m, const float* x, const float* t, const int64_t* y, float* eps, int B,
                                 int img_size, int precision, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Expert: GuidedUNet, the cross-attention UNet (row a7).
 * reference: src/compositional_diffusion_with_cross_attention.py:86-208.  Each block attends to ONE context
 * token, so softmax == 1 and the attention output is out_proj(v_proj(context)) for every pixel; that product is
 * folded into one matrix per block at finalize (q/k projections do not influence the output).  fp32 path.
 * ------------------------------------------------------------------------------------ */
typedef struct cdm_guided cdm_guided;
int cdm_guided_create(int num_digits, int num_colors, int embed_dim, int device, cdm_guided** out);
void cdm_guided_destroy(cdm_guided* m);
int cdm_guided_set_param(cdm_guided* m, const char* key, const float* host_data, int64_t numel);
int cdm_guided_finalize(cdm_guided* m);
size_t cdm_guided_workspace_bytes(const cdm_guided* m, int B, int img_size, int precision);
/* eps = model(x, t, digit_labels, color_labels): x [B,3,S,S]; t [B] fp32; labels [B] int64 (null index = num_*).
 * precision CDM_PREC_FP32: CUDA-core path (parity <= 1e-5); CDM_PREC_F16: every 3x3 conv and both ConvTranspose2d on
 * tcgen05 (fp16 operands, fp32 accumulation; img_size % 8 == 0). */
int cdm_guided_forward(cdm_guided* m, const float* x, const float* t, const int64_t* digits, const int64_t* colors,
                       float* eps, int B, int img_size, int precision, void* workspace, size_t workspace_bytes, void* stream);

/* Whole two-condition classifier-free-guidance chain in ONE host call (reference: the loop of
 * src/compositional_diffusion_with_cross_attention.py:279-313): per step the three forwards that enter the update
 * ((null, null), (digit, null), (null, colour)) + cdm_step_cfg (combine 0, x0-form update).  step_coef_host: HOST [n_steps, 3]
 * rows {t, sqrt(ab_prev), sqrt(1 - ab_prev)} in sampling order (t = T-1 .. 0). */
size_t cdm_guided_sample_cfg_workspace_bytes(const cdm_guided* m, int B, int img_size, int precision);
int cdm_guided_sample_cfg(cdm_guided* m, float* x, int digit, int color, float w_shape, float w_color,
                          const float* step_coef_host, int n_steps, int B, int img_size, int precision, void* workspace,
                          size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Test hook: ONE convolution layer through the chosen path, torch layouts in and out, so the
 * parity tests can check the implicit-GEMM kernels in isolation against conv2d.
 *   x [B,Cin,H,W] fp32 device; w_host [Cout,Cin,k,k] fp32 HOST (k = 3 when taps == 9, 1 when taps == 1);
 *   bias [bias_rows, Cout] fp32 device (bias_rows = 1 or B); res / wres_host: optional 1x1 residual conv
 *   input [B,Cres,H,W] device / weights [Cout,Cres] HOST; identity: optional [B,Cout,H,W] device;
 *   out [B,Cout,H,W] fp32 device; stats_out: optional [B,8,2] {sum, sumsq} per GroupNorm group.
 * precision: CDM_PREC_FP32, CDM_PREC_F16 (shifted-box tcgen05 kernel), 2 (halo-tile tcgen05 kernel, 3x3 only),
 * 3 (stacked halo-tile kernel, 3x3, Cout = 64, full-width strips: CDM_ERR_UNSUPPORTED otherwise) or CDM_PREC_F16X3
 * (three-term split-fp16 tcgen05 kernel, fp32 in / out).
 * Allocates and frees its own temporaries and synchronises the stream (debug only). */
int cdm_debug_conv(const float* x, const float* w_host, const float* bias, int bias_rows, const float* res,
                   const float* wres_host, const float* identity, float* out, float* stats_out, int B, int Cin,
                   int Cres, int Cout, int H, int W, int taps, int precision, void* stream);

/* Test hook: ONE general fp16 tensor-core convolution (conv_x3.cu, one product per MAC), torch layouts in and out.
 *   kind 0: 3x3 stride 1 pad 1; 1: 4x4 stride 2 pad 1 (Conv2d); 2: 4x4 stride 2 pad 1 TRANSPOSED (ConvTranspose2d, w_host in its
 *   [Cin, Cout, 4, 4] layout); 3: 1x1.  x1 [B,C1,H,W] (+ x2 [B,C2,H,W] concatenated along channels for kinds 0 / 3) fp32 device;
 *   w_host HOST fp32; bias [Cout] device or NULL; out [B,Cout,Ho,Wo] fp32 device.  Channel counts need not be multiples of
 *   64: the hook zero-pads them the way the expert graphs do.  Allocates its own temporaries and synchronises (debug only). */
int cdm_debug_conv_t16(const float* x1, const float* x2, const float* w_host, const float* bias, float* out, int B, int C1, int C2,
                       int Cout, int H, int W, int kind, int relu, void* stream);

/* Test hook: the init conv of the fp16 graphs alone (reference: UNet.init_conv, mnist/models/unet_small.py:57,78 / shapes/models/unet_small.py:74,106:
 * Conv2d(Cin, 64, 3, padding=1) on the NCHW fp32 image).  x [B,Cin,H,W] fp32 device; w [64,Cin,3,3], bias [64] (or NULL) fp32
 * DEVICE; out [B,64,H,W] fp32 device (the fp16 NHWC result converted back); stats_out: optional [B,8,2] {sum, sumsq}.
 * tensor_core = 1: the tcgen05 kernel (csrc/init_conv_tc.cu; CDM_ERR_UNSUPPORTED when it has no instance), 0: the CUDA-core kernel.
 * Allocates its own temporaries and synchronises the stream (debug only). */
int cdm_debug_init_conv(const float* x, const float* w, const float* bias, float* out, float* stats_out, int B, int Cin, int H, int W,
                        int tensor_core, void* stream);

/* Test hooks: the two elementwise producers of the UNet graphs in isolation, torch layouts in and out.
 *   cdm_debug_maxpool: F.max_pool2d(x, 2) (reference: self.pool = nn.MaxPool2d(2), mnist/models/unet_small.py:62,80,82).  x [B,C,H,W]
 *     fp32 device (H, W even; C/8 a multiple of 8); out [B,C,H/2,W/2]; stats_out / stats_in_out: optional [B,8,2] {sum, sumsq} per
 *     GroupNorm group of the pooled tensor / of the input tensor.
 *   cdm_debug_upcat: cat([F.interpolate(low, scale_factor=2, mode="bilinear", align_corners=True), skip], 1) (reference: self.unpool +
 *     torch.cat, mnist/models/unet_small.py:70,84-85,88-89).  low [B,Ca,h,w], skip [B,Cs,2h,2w]; out [B,Ca+Cs,2h,2w], or [B,Ca,2h,2w] with
 *     virtual_concat = 1 (only the upsampled channels are materialised; the statistics still cover the whole concat);
 *     stats_out: optional [B,8,2] of the concat tensor.
 * precision: CDM_PREC_FP32 or CDM_PREC_F16 (the values are rounded to fp16 on the way in).  Both allocate their own temporaries
 * and synchronise the stream (debug only). */
int cdm_debug_maxpool(const float* x, float* out, float* stats_out, float* stats_in_out, int B, int C, int H, int W, int precision,
                      void* stream);
int cdm_debug_upcat(const float* low, const float* skip, float* out, float* stats_out, int B, int Ca, int Cs, int h, int w,
                    int precision, int virtual_concat, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CDM_B200_H */
