/* du_b200.h — C ABI of the B200-native per-step uncertainty path of diffusion-uncertainty.
 *
 * The reference (Michedev/diffusion-uncertainty) is pure Python/PyTorch and has no FFI of its own
 * (SURVEY.md §8b); this header IS the drop-in boundary.  Every entry point replaces a group of eager
 * ATen expressions of the reference, cited per function as file:line relative to
 * /root/reference/diffusion_uncertainty/ (SU = schedulers_uncertainty, PU = pipeline_uncertainty).
 * The reference-side binding (the ctypes stub a maintainer would add) is in INTEGRATION.md; the
 * Python host layer that mirrors the reference's scheduler/pipeline API on top of this ABI is
 * diffusion-uncertainty_b200/.
 *
 * Conventions
 *  - Plain pointers and sizes only.  All data pointers are DEVICE pointers owned by the caller (torch
 *    allocations); the library never allocates user-visible memory and never synchronises the device.
 *  - A tensor is passed as a "rows view": B rows of n contiguous elements, row b starting
 *    `stride` ELEMENTS after row b-1.  This covers contiguous [B,C,H,W] tensors (stride = n), the
 *    ADM `model(...)[:, :3]` channel-slice view (stride = 6*H*W, n = 3*H*W) and a slot
 *    [:, t] of the [B,T_uc,C,H,W] accumulation buffer (stride = T_uc*n).
 *  - dtype codes: du_dtype.  Arithmetic is always fp32 (fp64 for the few global sums).
 *  - Every function is asynchronous on `stream` (a cudaStream_t), re-entrant, CUDA-graph capturable
 *    (no host reads of device results), and returns DU_OK or a negative du_status;
 *    du_last_error() gives the thread-local message.
 *  - No CPU fallback exists.
 */
#ifndef DU_B200_H
#define DU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* du_stream_t; /* cudaStream_t */

enum du_dtype { DU_F32 = 0, DU_F16 = 1, DU_BF16 = 2 };

enum du_status {
  DU_OK = 0,
  DU_ERR_BAD_ARG = -1,   /* null pointer, negative size, q outside [0,1], ...  -> ValueError      */
  DU_ERR_DTYPE = -2,     /* unsupported dtype code                             -> RuntimeError    */
  DU_ERR_ALIGN = -3,     /* pointer not aligned to its element size            -> RuntimeError    */
  DU_ERR_TOO_LARGE = -4, /* quantile row longer than 2^24 (torch.quantile's own limit)            */
  DU_ERR_CUDA = -5,      /* a CUDA runtime call failed (message in du_last_error)                 */
  DU_ERR_SCRATCH = -6    /* scratch buffer too small                                               */
};

#define DU_MAX_M 64 /* maximum number of score tensors reduced by one du_moments call */

const char* du_last_error(void);
int du_version(void);            /* ABI version, currently 1 */
int du_num_sms(int device);      /* SM count of `device` (148 on B200); <0 on error */
/* Select the device the calling THREAD's subsequent calls run on (thread-local; < 0 = whatever device is current).  The CUDA
 * current device of the caller is never changed: each call switches for its own duration and restores it on return. */
int du_set_device(int device);

/* ------------------------------------------------------------------------------------------------
 * F1 — reduction over the M axis.
 * Replaces torch.stack(scores,0) followed by
 *   DU_MOM_VAR_UNBIASED       torch.var(., dim=0)                SU/scheduling_ddim_mc_dropout.py:506,
 *                             SU/scheduling_ddim_uncertainty_threshold.py:537, generate_samples.py:815
 *   DU_MOM_CENTERED           (. - eps[None]).pow(2).mean(0)     SU/scheduling_ddim_uncertainty_zigzag_centered.py:549
 *   DU_MOM_VAR_WITH_CENTER    torch.var(stack(scores+[eps]),0)   uncertainty_guidance.py:101-106,
 *                             PU/pipeline_sampler_class_conditional_uncertainty_guided_posterior_distribution.py:58-61
 *   DU_MOM_RAW                .pow(2).mean(0)                    uncertainty_guidance.py:48
 *   DU_MOM_STD_UNBIASED       .std(0)                            generate_samples.py:941
 * scores: HOST array of M device pointers (M <= DU_MAX_M), all of dtype score_dtype and row stride
 * score_stride.  center may be NULL unless the mode needs it.  mean_out (fp32, nullable) receives
 * the mean over the samples the mode reduces (M, or M+1 for VAR_WITH_CENTER).  unc_out may be a slot
 * of the accumulation buffer (F8 fused: generate_samples.py:192-201).
 * ---------------------------------------------------------------------------------------------- */
enum du_moments_mode {
  DU_MOM_VAR_UNBIASED = 0,
  DU_MOM_CENTERED = 1,
  DU_MOM_VAR_WITH_CENTER = 2,
  DU_MOM_RAW = 3,
  DU_MOM_STD_UNBIASED = 4,
  /* partial results for M-sharding across GPUs: unc_out = sum of squared deviations
   * about the LOCAL mean (or about `center` when center != NULL), mean_out = local mean */
  DU_MOM_PARTIAL_M2 = 5
};

int du_moments(const void* const* scores, int M, int64_t score_stride, int score_dtype,
               const void* center, int64_t center_stride, int center_dtype, int mode,
               int64_t B, int64_t n,
               void* unc_out, int64_t unc_stride, int unc_dtype,
               float* mean_out, int64_t mean_stride, du_stream_t stream);

/* Chan merge of R per-rank partials (after an all-gather) into the final map.
 * means[r], m2s[r]: device pointers to contiguous [B*n] fp32 (from DU_MOM_PARTIAL_M2); counts[r] =
 * samples reduced on rank r.  mode: DU_MOM_VAR_UNBIASED / DU_MOM_STD_UNBIASED (divide by total-1) or
 * DU_MOM_CENTERED (m2s are sums about the common centre: plain sum / total; means may be NULL). */
int du_moments_merge(const float* const* means, const float* const* m2s, const int* counts, int R,
                     int mode, int64_t N, float* unc_out, float* mean_out, du_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * F2a — per-row linear-interpolated quantile, bit-identical to
 *   torch.quantile(u.flatten(1).to(float32), q, dim=1)
 * PU/...posterior_distribution.py:15, uncertainty_guidance.py:112, generate_samples.py:946.
 * rank = fp32(q)*fp32(n-1); lo = floor, hi = ceil; thr = lerp(sorted[lo], sorted[hi], rank-lo) with
 * torch's two-branch lerp.  lerp_fma = 0 rounds the lerp once per operation (torch's CPU kernel where it
 * is not dispatched to an FMA build), 1 fuses the multiply-add of the selected branch (torch's CUDA
 * kernel, and its CPU kernels on AVX2 / AVX-512 dispatch).  The two differ by at most one unit in the
 * last place of the threshold, and only when the two order statistics are far apart relative to their
 * value (short rows): for rows of thousands of elements they coincide.  A row containing NaN yields NaN.
 * thr_out[B]; rank_out[B][2] (nullable) = {lo, hi}; val_out[B][2] (nullable) = the two order
 * statistics.  scratch: du_quantile_scratch_bytes(B, n) bytes of device memory.
 * ---------------------------------------------------------------------------------------------- */
size_t du_quantile_scratch_bytes(int64_t B, int64_t n);
int du_quantile_threshold(const float* u, int64_t B, int64_t n, int64_t stride, float q, int lerp_fma,
                          float* thr_out, int32_t* rank_out, float* val_out,
                          void* scratch, size_t scratch_bytes, du_stream_t stream);

/* mask = (u > thr[b]) (higher != 0) or (u < thr[b]), strict, as fp32 0/1.
 * PU/...posterior_distribution.py:16-20. */
int du_threshold_mask(const void* u, int64_t u_stride, int u_dtype, const float* thr, int higher,
                      int64_t B, int64_t n, float* mask_out, int64_t mask_stride, du_stream_t stream);

/* F2b — mask = (u > thr_map) with thr_map one row of n elements broadcast over the batch
 * (threshold[i].unsqueeze(0)).  PU/...posterior_distribution.py:21-29, generate_samples.py:819. */
int du_tensor_threshold_mask(const void* u, int64_t u_stride, int u_dtype, const void* thr_map,
                             int thr_dtype, int higher, int64_t B, int64_t n, float* mask_out,
                             int64_t mask_stride, du_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * F2c — whole-batch z-normalisation statistics: stats_out[0] = mean, [1] = unbiased std,
 * [2] = count, [3] = sum of squared deviations (fp32, device).  Deterministic (fixed merge order).
 * SU/scheduling_ddim_uncertainty_threshold.py:539-540.
 * du_znorm_stats_merge rescales stats after the per-rank (count, mean, M2) triples were all-reduced
 * is not needed: pass R gathered stat blocks to du_znorm_stats_combine.
 * ---------------------------------------------------------------------------------------------- */
size_t du_znorm_scratch_bytes(int64_t B, int64_t n);
int du_znorm_stats(const void* u, int64_t u_stride, int u_dtype, int64_t B, int64_t n, float* stats_out,
                   void* scratch, size_t scratch_bytes, du_stream_t stream);
/* Chan merge of R stats blocks ([R][4], device, e.g. after an all-gather over ranks) into one. */
int du_znorm_stats_combine(const float* stats_in, int R, float* stats_out, du_stream_t stream);

enum du_znorm_mode { DU_ZN_BELOW = 0 /* 'max': z < thr */, DU_ZN_ABOVE = 1 /* z > thr */, DU_ZN_MULTISCALE = 2 };
/* z = normalize ? (u-mean)/std : u;  weights per mode (multiscale: 0.8 / 0.9 / 1.0 bands,
 * SU/scheduling_ddim_infer_noise_multiscale_threshold.py:538-548).  z_out / w_out nullable. */
int du_znorm_weights(const void* u, int64_t u_stride, int u_dtype, const float* stats, int normalize,
                     int mode, float thr, int64_t B, int64_t n, float* z_out, int64_t z_stride,
                     float* w_out, int64_t w_stride, du_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * F3 — DDIM / DDPM-variance x_{t-1} update.  SU/scheduling_ddim_uncertainty_zigzag_centered.py:472-525.
 * The host computes the scalars with the reference's own fp32 expressions (:462-468, 294-302, 507).
 * ---------------------------------------------------------------------------------------------- */
enum du_prediction_type { DU_PRED_EPSILON = 0, DU_PRED_SAMPLE = 1, DU_PRED_V = 2 };

typedef struct du_ddim_coeffs {
  float sqrt_alpha_t;     /* alpha_prod_t ** 0.5                       */
  float sqrt_beta_t;      /* (1 - alpha_prod_t) ** 0.5                 */
  float sqrt_alpha_prev;  /* alpha_prod_t_prev ** 0.5                  */
  float dir_coef;         /* (1 - alpha_prod_t_prev - std_dev_t**2) ** 0.5 */
  float sigma;            /* std_dev_t = eta * variance ** 0.5         */
  float clip_range;       /* clip_sample_range                         */
  int32_t prediction_type;
  int32_t clip_sample;
  int32_t use_clipped_model_output;
  int32_t add_noise;      /* eta > 0: prev += sigma * noise            */
} du_ddim_coeffs;

/* prev_out / x0_out / eps_out are nullable (at least one must be given).  noise required iff add_noise. */
int du_ddim_step(const void* model_output, int64_t mo_stride, int mo_dtype,
                 const void* sample, int64_t s_stride, int s_dtype,
                 const void* noise, int64_t noise_stride, int noise_dtype,
                 const du_ddim_coeffs* c, int64_t B, int64_t n,
                 void* prev_out, int64_t prev_stride, int prev_dtype,
                 void* x0_out, int64_t x0_stride, int x0_dtype,
                 void* eps_out, int64_t eps_stride, int eps_dtype, du_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * F2 (mask) + F4/F5/F6 (guided score) + F3 (DDIM) in ONE elementwise pass.
 *   guidance   reference                                                      eps' =
 *   POSTERIOR  uncertainty_guidance.py:115-120; PU/...posterior_distribution.py:63-68,160
 *                                      eps(1-m) + m * [1/(M/u + 1/abar)] * (1/u) * S
 *   GRAD_BLEND PU/...guided_gradient.py:117-118        eps(1-m) + (eps + lam*g) m
 *   GRAD_ADD   uncertainty_guidance.py:129             eps + lam*g*m
 *   WEIGHTS    SU/scheduling_ddim_uncertainty_threshold.py:554-574   eps*w; x0 from the UNMASKED eps
 *   LINCOMB    SU/scheduling_ddim_mc_dropout_gradient.py:514         post_M*eps + lam*g   (post_M carries the eps weight)
 *   SIGN_ADD   PU/..._guided_second_order.py:249                     eps + u * sign(n) * m   (aux = the normal draw n)
 *   MUL_BLEND  generate_samples.py:953                               eps(1-m) + eps*m*g
 *   NONE       plain F3
 * x0_unguided != 0: x0 is computed from the UNGUIDED eps and only the direction term uses eps' — what every in-scheduler
 * guidance of the reference does (SU/scheduling_ddim_uncertainty_grad.py:551-570, ..._mc_dropout_gradient.py:514-515,
 * ..._model_gradient_guided.py:554-562); WEIGHTS always behaves that way.  GRAD_ADD / LINCOMB need no mask (m = 1).
 * mask source: per-row threshold thr[B] compared with u (strict; `higher`), or an explicit fp32
 * mask/weight tensor.
 * ---------------------------------------------------------------------------------------------- */
enum du_guidance { DU_GUIDE_NONE = 0, DU_GUIDE_POSTERIOR = 1, DU_GUIDE_GRAD_BLEND = 2, DU_GUIDE_GRAD_ADD = 3, DU_GUIDE_WEIGHTS = 4,
                   DU_GUIDE_LINCOMB = 5, DU_GUIDE_SIGN_ADD = 6, DU_GUIDE_MUL_BLEND = 7 };

typedef struct du_guided_params {
  /* inputs */
  const void* eps;      int64_t eps_stride;    int32_t eps_dtype;    int32_t guidance;
  const void* sample;   int64_t sample_stride; int32_t sample_dtype; int32_t higher;
  const float* u;       int64_t u_stride;      /* uncertainty map (POSTERIOR, or thr masks)         */
  const float* thr;                            /* [B] per-row thresholds, or NULL                   */
  const float* mask;    int64_t mask_stride;   /* explicit mask / weights, or NULL                  */
  const void* aux;      int64_t aux_stride;    int32_t aux_dtype;    int32_t aux_broadcast;
                                               /* POSTERIOR: S (aux_broadcast: one row for all b);
                                                  GRAD_*: g                                         */
  float lam;                                   /* lambda_update / lr                                */
  float post_M;                                /* float(M) of the posterior formula                 */
  float inv_alpha_hat;                         /* 1 / alpha_hat_t, computed on the host in fp32     */
  int32_t skip_ddim;                           /* only write eps_out / mask_out                     */
  du_ddim_coeffs ddim;
  int64_t B, n;
  /* outputs, all nullable */
  void* prev_out;       int64_t prev_stride;   int32_t prev_dtype;   int32_t _pad0;
  void* x0_out;         int64_t x0_stride;     int32_t x0_dtype;     int32_t x0_unguided;
  void* eps_out;        int64_t eps_out_stride; int32_t eps_out_dtype; int32_t _pad2;
  float* mask_out;      int64_t mask_out_stride;
  int64_t mask_period;                         /* > 0: mask rows hold mask_period elements, element i uses mask[i % mask_period]
                                                  (a [B,1,H,W] mask over [B,C,H,W] rows); 0: same length as the rows */
} du_guided_params;

int du_guided_step(const du_guided_params* p, du_stream_t stream);

/* F5 helper — S[n] = sum over the batch axis of x (the reference's `pred_epsilon.sum(dim=0)`,
 * uncertainty_guidance.py:119), accumulated in fp64, deterministic. */
int du_batch_sum(const void* x, int64_t x_stride, int x_dtype, int64_t B, int64_t n, float* out, du_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * F7 — out = a*x + b*noise (one rounding per operation).
 * SU/scheduling_ddim_uncertainty_zigzag_centered.py:538 (a = sqrt(1-beta_t), b = sqrt(beta_t)),
 * :593-626 add_noise (a = sqrt(abar_t), b = sqrt(1-abar_t)), uncertainty_guidance.py:87.
 * ---------------------------------------------------------------------------------------------- */
int du_perturb(const void* x, int64_t x_stride, int x_dtype, const void* noise, int64_t noise_stride,
               int noise_dtype, float a, float b, int64_t B, int64_t n, void* out, int64_t out_stride,
               int out_dtype, du_stream_t stream);   /* noise == NULL: out = a*x */
/* The same with one (a, b) pair PER ROW, read from device vectors a_rows[B], b_rows[B] — add_noise / get_velocity called with a
 * vector of per-sample timesteps (SU/scheduling_ddim_uncertainty_zigzag_centered.py:606-626, 629-646). */
int du_perturb_rows(const void* x, int64_t x_stride, int x_dtype, const void* noise, int64_t noise_stride,
                    int noise_dtype, const float* a_rows, const float* b_rows, int64_t B, int64_t n, void* out,
                    int64_t out_stride, int out_dtype, du_stream_t stream);

/* Second-order momentum of the map (PU/pipeline_sampler_class_conditional_uncertainty_guided_second_order.py:212-218):
 *   m' = momentum ? beta*momentum + one_minus_beta*u : u;   corrected = m' / denom;   root = sqrt(corrected)
 * over N contiguous elements (momentum fp32, nullable on the first step; corrected_out / sqrt_out nullable).  The host passes
 * denom = 1 - beta**i + 1e-5 and one_minus_beta = 1 - beta, both evaluated in double like the reference's Python scalars. */
int du_ema_update(const float* momentum, const void* u, int u_dtype, float beta, float one_minus_beta, float denom, int64_t N,
                  float* momentum_out, float* corrected_out, float* sqrt_out, du_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * F7 + N1 — the same combination with the noise drawn in the kernel:
 *   n   = the N standard normals `torch.randn_like(x)` yields on a CUDA device for the Philox generator state
 *         (seed, offset) — bit-identical (same Philox4x32-10 words, same Box-Muller, same element mapping as
 *         ATen/native/cuda/DistributionTemplates.h `normal_and_transform`), rounded to noise_dtype;
 *   out = a*x + b*n (one rounding per operation, as du_perturb).
 * x: N contiguous elements (NULL: only the noise is produced, into noise_out).  noise_out: nullable.
 * device_state: nullable device pointer to {seed, offset}; when given its seed replaces the host seed and the host `offset` is
 * ADDED to its offset (draws that replay from a CUDA graph: the k-th draw of a graph passes the increments of the draws before it
 * as `offset`, and ONE du_rng_advance at the end of the graph moves the state past all of them).
 * Replaces `noise = torch.randn_like(pred_x_0)` + the expression of
 * SU/scheduling_ddim_uncertainty_zigzag_centered.py:529-538, SU/scheduling_ddim_uncertainty_centered.py:525-531,
 * uncertainty_guidance.py:86-88.  The caller advances its generator by du_randn_offset_increment(N).
 * ---------------------------------------------------------------------------------------------- */
int du_perturb_randn(const void* x, int x_dtype, int64_t N, uint64_t seed, uint64_t offset,
                     const uint64_t* device_state, float a, float b, void* out, int out_dtype,
                     void* noise_out, int noise_dtype, du_stream_t stream);
/* Philox offset consumed by one N-element normal draw on the current device (torch's calc_execution_policy). */
int du_randn_offset_increment(int64_t N, uint64_t* increment_out);
/* device_state[1] += increment, on the stream (graph-capturable). */
int du_rng_advance(uint64_t* device_state, uint64_t increment, du_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * F8 — copy (and convert) one step's map into its slot of the [B, T_uc, ...] accumulation buffer:
 * dst row b = dst + b*dst_stride.  generate_samples.py:192-201,229-231.
 * ---------------------------------------------------------------------------------------------- */
int du_accumulate_slot(const void* src, int64_t src_stride, int src_dtype, int64_t B, int64_t n,
                       void* dst, int64_t dst_stride, int dst_dtype, du_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Image epilogue of the sampling loops: out = uint8(round((x/2 + 0.5).clamp(0,1) * 255)),
 * round half to even like torch.round.  generate_samples.py:203-215.
 * ---------------------------------------------------------------------------------------------- */
int du_image_uint8(const void* x, int64_t x_stride, int x_dtype, int64_t B, int64_t n, uint8_t* out,
                   int64_t out_stride, du_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * The fused uncertainty step: F1 -> F2a -> F5 -> F3 (+F8) in ONE launch.  One thread-block cluster per
 * image keeps the image's map (and eps) in distributed shared memory, selects the two order
 * statistics there, and applies the guided DDIM update, so HBM sees each input and output once
 * (36 B/element at fp32, M=5).
 * ---------------------------------------------------------------------------------------------- */
typedef struct du_fused_params {
  const void* scores[DU_MAX_M]; int32_t M; int32_t score_dtype; int64_t score_stride;
  const void* eps;      int64_t eps_stride;     /* same dtype as scores                              */
  const void* sample;   int64_t sample_stride;  int32_t sample_dtype; int32_t moments_mode;
  const float* S;       int64_t S_stride;       int32_t S_broadcast;  int32_t higher;
                                                /* posterior sum source (fp32); NULL = eps itself    */
  float q; int32_t lerp_fma; float post_M; float inv_alpha_hat;
  du_ddim_coeffs ddim;
  int64_t B, n;
  float* unc_out;       int64_t unc_stride;     /* map (or its accumulation slot)                    */
  float* thr_out;                               /* [B], nullable                                     */
  void* prev_out;       int64_t prev_stride;    int32_t prev_dtype;   int32_t S_overlap;
                                                /* S_overlap != 0: S is written by the du_batch_sum launch that
                                                 * IMMEDIATELY precedes this call on the stream; the step is then
                                                 * launched as its programmatic dependent (its sampling pilot, which
                                                 * does not read S, overlaps the sum's tail; griddepcontrol.wait
                                                 * orders the first S read).  0: plain stream order.              */
  void* x0_out;         int64_t x0_stride;      /* nullable, prev_dtype                              */
  void* eps_out;        int64_t eps_out_stride; /* nullable, fp32                                    */
  float* mask_out;      int64_t mask_out_stride;/* nullable                                          */
  int32_t skip_ddim;    int32_t _reserved0;     /* skip_ddim != 0: stop after the posterior blend — eps_out (required) receives the
                                                 * guided score, sample / prev_out / x0_out are ignored (may be NULL).  This is the
                                                 * whole body of get_uncertainty_guided_score_with_percentile (uncertainty_guidance.py:
                                                 * 99-120), whose caller applies its own scheduler afterwards.                      */
} du_fused_params;

/* returns DU_ERR_TOO_LARGE when a row does not fit the cluster's shared memory (use the unfused calls) */
int du_fused_uncertainty_step(const du_fused_params* p, du_stream_t stream);
int du_fused_supported(int64_t n, int score_dtype);
/* Which kernel the calling thread's last du_fused_uncertainty_step launched: 0 = none yet, 1 = the three-phase cluster
 * kernel (fused_step_kernel), 2 = the predictive single-pass kernel (fused_pred_kernel; slices of >= 4 trips with the
 * epsilon-prediction fp32 update).  Both give bit-identical results; benchmarks use this to name what they timed. */
int du_fused_last_kernel(void);

/* ------------------------------------------------------------------------------------------------
 * DPM-Solver++ multistep update of the `dpm_2_uncertainty_centered` scheduler
 * (SU/scheduling_dpm_2_uncertainty_centered.py:617-620 first order, :686-700 second order midpoint / heun):
 *   out = (a*sample + b*m0) + c*(k*(m0 - m1))        one rounding per operation; m1 == NULL: out = a*sample + b*m0
 * with the host scalars a = sigma_t/sigma_s0, b = -(alpha_t (exp(-h) - 1)), k = 1/r0 and c = -0.5*(alpha_t (exp(-h) - 1))
 * (midpoint) or alpha_t ((exp(-h) - 1)/h + 1) (heun).  m0 / m1 are the converted model outputs (x0 predictions).
 * ---------------------------------------------------------------------------------------------- */
int du_dpm_solver_update(const void* sample, int64_t s_stride, int s_dtype, const void* m0, int64_t m0_stride, int m0_dtype,
                         const void* m1, int64_t m1_stride, int m1_dtype, float a, float b, float c, float k,
                         int64_t B, int64_t n, void* out, int64_t out_stride, int out_dtype, du_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * N4 — flip-based uncertainty.  Tensors are [B, C, H, W] rows views (row = C*H*W elements).
 * du_flip_h: out[b,c,h,w] = x[b,c,H-1-h,w]  — torch.flip(x0, dims=[2]), the model input of the flipped forward
 *   (SU/scheduling_ddim_flip.py:487).
 * du_flip_sqdiff: u = (eps - flip_h(flipped_output))^2 (SU/scheduling_ddim_flip.py:488-493); channel_amax != 0 also reduces
 *   with amax over C (NaN-propagating) into rows of H*W elements (SU/scheduling_ddim_flip_threshold.py:504-506).
 * ---------------------------------------------------------------------------------------------- */
int du_flip_h(const void* x, int64_t x_stride, int x_dtype, int64_t B, int64_t C, int64_t H, int64_t W,
              void* out, int64_t out_stride, int out_dtype, du_stream_t stream);
int du_flip_sqdiff(const void* eps, int64_t eps_stride, int eps_dtype, const void* flipped, int64_t f_stride, int f_dtype,
                   int64_t B, int64_t C, int64_t H, int64_t W, int channel_amax, float* out, int64_t out_stride,
                   du_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * F6 — backward of du_moments for the schedulers / pipelines that differentiate the map through the score model
 * (`uncertainty.mean(dim=0).sum().backward()`: SU/scheduling_ddim_uncertainty_grad.py:536-538,
 * SU/scheduling_ddim_mc_dropout_gradient.py:499-503, SU/scheduling_ddim_model_gradient_guided.py:546-548,
 * PU/pipeline_sampler_class_conditional_uncertainty_guided_gradient.py:190-194).  grad_scores[m] = grad_u * d u / d s_m for
 * modes VAR_UNBIASED, STD_UNBIASED (generate_samples.py:941-943), CENTERED, VAR_WITH_CENTER; grad_center (nullable) = grad_u * d u / d center.  grad_scores[m] may be
 * NULL for samples that need no gradient.  The score model's own backward stays torch autograd.
 * ---------------------------------------------------------------------------------------------- */
int du_moments_backward(const void* const* scores, int M, int64_t score_stride, int score_dtype, const void* center,
                        int64_t center_stride, int center_dtype, int mode, const void* grad_u, int64_t gu_stride,
                        int gu_dtype, int64_t B, int64_t n, void* const* grad_scores, int64_t grad_stride,
                        int grad_dtype, void* grad_center, int64_t gc_stride, du_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * N2 — per-pixel threshold fitting: out[j] = the k-th smallest (0-based) of x[0..N-1][j] over the sample axis, NaN last,
 * i.e. uncertainties_timestep.gather(0, uncertainties_timestep.argsort(dim=0)[k]) with k = int(num_samples * perc)
 * (scripts/compute_threshold_pixel_wise.py:90-100, 143-152).  x: [N, n] with row stride `row_stride` elements (a [:, i]
 * slice of the [N, T_uc, C, H, W] accumulated maps); out: [n], same dtype.
 * ---------------------------------------------------------------------------------------------- */
int du_column_kth(const void* x, int dtype, int64_t N, int64_t n, int64_t row_stride, int64_t k, void* out, du_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * N3 — per-image reductions of the accumulated map: du_row_sum: out[b] = sum of row b (uncertainty.sum(dim=(1,2,3,4)),
 * scripts/uncertainty_benchmark_imagenet.py:314); du_slot_sum: out[b, :] = sum over the T slots of [B, T, n]
 * (uncertainty.sum(dim=1), scripts/compute_ause.py:128).  fp32 accumulation in a fixed order.
 * ---------------------------------------------------------------------------------------------- */
int du_row_sum(const void* x, int64_t x_stride, int x_dtype, int64_t B, int64_t n, float* out, du_stream_t stream);
int du_slot_sum(const void* x, int64_t x_stride, int64_t slot_stride, int x_dtype, int64_t B, int T, int64_t n,
                float* out, int64_t out_stride, du_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DU_B200_H */
