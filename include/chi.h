/*
 * chi.h -- C ABI of libchi: the B200 (sm_100a) forward-model likelihood path of
 * caribou_hi (spin temperature -> line-profile optical depth -> ordered-cloud
 * radiative transfer -> Gaussian log-likelihood -> reverse-mode gradient).
 *
 * The reference (tvwenger/caribou_hi) has no FFI: its seam is the Python function
 * API of caribou_hi/physics.py plus the add_likelihood() bodies of the six model
 * classes, executed by PyTensor.  Every entry point below names the reference
 * code it replaces (file:line into the reference tree).  The reference-side
 * binding (ctypes stub + PyTensor Op) is shown in INTEGRATION.md.
 *
 * Conventions
 *  - every call returns 0 on success or a negative CHI_E* code; it never throws,
 *    aborts or falls back to a CPU implementation.  chi_last_error() gives text.
 *  - all arrays are C-contiguous float64 (fp64 is the reference's floatX);
 *    "host" entry points take host pointers (pageable or pinned), "_dev" entry
 *    points take device pointers valid on the handle's device and enqueue on the
 *    handle's stream without synchronising.
 *  - the caller owns every buffer passed in; the library owns its device
 *    workspace, stream and events inside the handle.  One handle per
 *    (process, device); calls on one handle must be serialised by the caller.
 *  - parameter blocks are structure-of-arrays: theta[B][CHI_NSLOT][N]
 *    (draw-major, then slot, clouds fastest).
 */
#ifndef CHI_H
#define CHI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHI_VERSION 100

/* error codes */
#define CHI_OK 0
#define CHI_EINVAL (-1)   /* bad argument / unsupported configuration */
#define CHI_ECUDA (-2)    /* CUDA runtime error (text in chi_last_error) */
#define CHI_ENODEV (-3)   /* no sm_100 device: there is NO CPU fallback */
#define CHI_ESTATE (-4)   /* call order violated (e.g. spectra not set) */
#define CHI_ENOMEM (-5)

/* Model kinds.  CHI_MODEL_PHYSICS takes the seven likelihood inputs directly
 * (the boundary a PyTensor Op uses); the other six take the free random variables
 * of the corresponding reference class and apply its add_priors()/
 * add_deterministics() maps on the device:
 *   ABSORPTION                    absorption_model.py:44-72 + hi_model.py:93-136
 *   EMISSION                      emission_model.py:60-132
 *   EMISSION_ABSORPTION           emission_absorption_model.py:45-64
 *   ABSORPTION_PHYSICAL           absorption_physical_model.py:43-77 + hi_physical_model.py:109-197
 *   EMISSION_PHYSICAL             emission_physical_model.py:59-129
 *   EMISSION_ABSORPTION_PHYSICAL  emission_absorption_physical_model.py:41-60 */
enum {
  CHI_MODEL_PHYSICS = 0,
  CHI_MODEL_ABSORPTION = 1,
  CHI_MODEL_EMISSION = 2,
  CHI_MODEL_EMISSION_ABSORPTION = 3,
  CHI_MODEL_ABSORPTION_PHYSICAL = 4,
  CHI_MODEL_EMISSION_PHYSICAL = 5,
  CHI_MODEL_EMISSION_ABSORPTION_PHYSICAL = 6
};

/* Slots of theta[B][CHI_NSLOT][N]. */
#define CHI_NSLOT 10
/* CHI_MODEL_PHYSICS: the likelihood inputs (slots 7..9 ignored) */
enum {
  CHI_PH_VELOCITY = 0,
  CHI_PH_FWHM2 = 1,
  CHI_PH_FWHM_L = 2,
  CHI_PH_TAU_TOTAL = 3,
  CHI_PH_TSPIN = 4,
  CHI_PH_FILLING_FACTOR = 5,
  CHI_PH_ABSORPTION_WEIGHT = 6
};
/* model kinds 1..6: the free RVs, named as the reference names them */
enum {
  CHI_FR_FWHM2_NORM = 0,          /* hi_model.py:95, hi_physical_model.py:111 */
  CHI_FR_VELOCITY_NORM = 1,       /* hi_model.py:109, hi_physical_model.py:115 */
  CHI_FR_LOG10_NHI_NORM = 2,      /* hi_model.py:99 (observational models only) */
  CHI_FR_LOG10_N_ALPHA_NORM = 3,  /* hi_model.py:122, hi_physical_model.py:128 */
  CHI_FR_NTH_FWHM_1PC = 4,        /* hi_physical_model.py:138 (physical models only) */
  CHI_FR_FWHM_L_NORM = 5,         /* hi_model.py:133 per cloud; hi_physical_model.py:148 one
                                     value per draw, replicated over clouds by the caller, whose
                                     gradient is the sum over clouds */
  CHI_FR_AMPLITUDE_NORM = 6,      /* tau_total_norm (absorption_model.py:46) | TB_fwhm_norm
                                     (emission_model.py:62) | NHI_fwhm2_thermal_norm
                                     (absorption_physical_model.py:45) | ff_NHI_norm
                                     (emission_physical_model.py:61) */
  CHI_FR_THERMAL = 7,             /* tkin_factor (absorption_model.py:55) | tkin_factor_norm
                                     (emission_model.py:76) | fwhm2_thermal_fraction
                                     (absorption_physical_model.py:51) |
                                     fwhm2_thermal_fraction_norm (emission_physical_model.py:75) */
  CHI_FR_FILLING_FACTOR_NORM = 8, /* emission_model.py:107, emission_physical_model.py:110 */
  CHI_FR_LOG10_WT = 9             /* log10_wt_ff_tkin (emission_absorption_model.py:53) |
                                     log10_wt_ff_fwhm2_thermal
                                     (emission_absorption_physical_model.py:46) */
};

/* Per-cloud derived quantities written by chi_deterministics (the reference's
 * pm.Deterministic nodes), det[B][CHI_NDET][N]. NaN where a model lacks one. */
#define CHI_NDET 20
enum {
  CHI_DET_VELOCITY = 0,
  CHI_DET_FWHM2 = 1,
  CHI_DET_FWHM_L = 2,
  CHI_DET_TAU_TOTAL = 3,
  CHI_DET_TSPIN = 4,
  CHI_DET_FILLING_FACTOR = 5,
  CHI_DET_ABSORPTION_WEIGHT = 6,
  CHI_DET_TKIN = 7,
  CHI_DET_LOG10_NHI = 8,       /* column density */
  CHI_DET_LOG10_nHI = 9,       /* volume density */
  CHI_DET_LOG10_DEPTH = 10,
  CHI_DET_FWHM2_THERMAL = 11,
  CHI_DET_FWHM2_NONTHERMAL = 12,
  CHI_DET_LOG10_PTH = 13,
  CHI_DET_LOG10_PNTH = 14,
  CHI_DET_LOG10_PTOT = 15,
  CHI_DET_LOG10_N_ALPHA = 16,
  CHI_DET_DEPTH = 17,                 /* pc; physical models (hi_physical_model.py:163) */
  CHI_DET_TB_FWHM = 18,               /* emission models (emission_model.py:66) */
  CHI_DET_FWHM2_THERMAL_FRACTION = 19 /* emission physical models (emission_physical_model.py:81) */
};

typedef struct chi_config {
  int32_t model;            /* CHI_MODEL_* */
  int32_t n_clouds;         /* N, 1..CHI_MAX_CLOUDS */
  int32_t has_emission;     /* CHI_MODEL_PHYSICS only; implied by the kind otherwise */
  int32_t has_absorption;   /* CHI_MODEL_PHYSICS only */
  int32_t lorentzian;       /* 0: prior_fwhm_L=None (pure Gaussian, fwhm_L==0, hi_model.py:136);
                               1: pseudo-Voigt with free fwhm_L */
  int32_t free_weight;      /* EMISSION_ABSORPTION only: 0 -> absorption_weight==1
                               (emission_absorption_model.py:46-47), 1 -> log-normal weight */
  int32_t dtype;            /* 0 = fp64 (only value implemented) */
  int32_t reserved;
  double bg_temp;               /* emission_model.py:24 */
  double depth_nth_fwhm_power;  /* hi_physical_model.py:24 */
  double prior_fwhm2;           /* hi_model.py:96 */
  double prior_velocity[2];     /* hi_model.py:117 */
  double prior_log10_nHI[2];    /* hi_model.py:104 */
  double prior_log10_n_alpha[2];/* hi_model.py:127 */
  double prior_fwhm_L;          /* hi_model.py:134 */
  double prior_amplitude;       /* prior_tau_total | prior_TB_fwhm | prior_NHI_fwhm2_thermal |
                                   prior_ff_NHI, by model kind */
} chi_config;

#define CHI_MAX_CLOUDS 16
#define CHI_MAX_BASELINE 8

/* One dataset ("emission" or "absorption") of a bayes_spec SpecData dict, for
 * n_sightlines lines of sight sharing one spectral axis. */
typedef struct chi_dataset {
  int32_t n_channels;        /* S >= 2; 0 = dataset absent */
  int32_t n_baseline;        /* D+1 baseline terms, 0..CHI_MAX_BASELINE */
  const double* spectral;    /* [S] velocity axis (km/s), SpecData.spectral */
  const double* brightness;  /* [n_sightlines][S] observed spectrum, SpecData.brightness */
  const double* noise;       /* [n_sightlines][S] rms per channel (>0), SpecData.noise */
  const double* basis;       /* [n_baseline][S] baseline basis; mu += sum_i coeff_i*basis_i[s]
                                (BaseModel.predict_baseline is affine in its coefficients) */
  const double* offset;      /* [S] constant part of the baseline, or NULL for 0 */
} chi_dataset;

typedef struct chi_handle chi_handle;

/* ---- lifetime ------------------------------------------------------------- */
int chi_version(void);
/* number of CUDA devices with compute capability 10.x visible to the process */
int chi_device_count(void);
/* Create a handle on `device`.  Fails with CHI_ENODEV if the device is not
 * sm_100: there is no CPU fallback. */
int chi_create(chi_handle** out, int device, const chi_config* cfg);
void chi_destroy(chi_handle* h);
/* Text of the last error on this handle (or of the last failed chi_create when h
 * is NULL).  Valid until the next call on the handle. */
const char* chi_last_error(const chi_handle* h);
/* Upload the observed spectra (host pointers).  Replaces self.data[...] captured by
 * add_likelihood() (emission_model.py:141,164-169 etc.).  Either dataset may be
 * NULL / have n_channels == 0 when the model kind does not use it. */
int chi_set_spectra(chi_handle* h, int n_sightlines, const chi_dataset* emission,
                    const chi_dataset* absorption);
/* The handle's CUDA stream (cudaStream_t as void*) for callers that enqueue
 * their own work around the *_dev entry points. */
void* chi_stream(chi_handle* h);
int chi_synchronize(chi_handle* h);

/* pinned host memory for fast host<->device copies of the host entry points */
int chi_host_alloc(void** out, size_t bytes);
int chi_host_free(void* p);

/* ---- physics-level entry points (handle needs no spectra) ------------------ */
/* physics.calc_spin_temp (physics.py:58-106), elementwise over n values. */
int chi_spin_temp(chi_handle* h, int64_t n, const double* kinetic_temp, const double* density,
                  const double* n_alpha, double* spin_temp_out);
/* Its vector-Jacobian product: given gbar = d(cost)/d(spin_temp), the gradients
 * with respect to the three inputs (what PyTensor's grad of physics.py:58-106
 * produces; the switch passes gradient to the selected branch only). */
int chi_spin_temp_vjp(chi_handle* h, int64_t n, const double* kinetic_temp, const double* density,
                      const double* n_alpha, const double* gbar, double* g_kinetic_temp,
                      double* g_density, double* g_n_alpha);
/* The elementwise helpers of physics.py that the reference evaluates through PyTensor
 * (tests/test_physics.py .eval()s them): out[i] = fn(a[i], b[i], c[i]).
 *   CHI_FN_GAUSSIAN                   a=x, b=center, c=fwhm          physics.py:16-35
 *   CHI_FN_LORENTZIAN                 a=x, b=center, c=fwhm          physics.py:38-55
 *   CHI_FN_LOG10_COLUMN_DENSITY       a=tau_total, b=spin_temp       physics.py:191-207
 *   CHI_FN_LOG10_NONTHERMAL_PRESSURE  a=log10_density, b=fwhm2_nth   physics.py:247-262
 * c may be NULL for the two-argument functions. */
#define CHI_FN_GAUSSIAN 0
#define CHI_FN_LORENTZIAN 1
#define CHI_FN_LOG10_COLUMN_DENSITY 2
#define CHI_FN_LOG10_NONTHERMAL_PRESSURE 3
int chi_elementwise(chi_handle* h, int fn, int64_t n, const double* a, const double* b, const double* c,
                    double* out);
/* physics.calc_pseudo_voigt (physics.py:265-309): profile_out[S][N].  fwhm_L has
 * N entries (replicate a scalar). */
int chi_pseudo_voigt(chi_handle* h, int S, int N, const double* velo_axis, const double* velocity,
                     const double* fwhm2, const double* fwhm_L, double* profile_out);
/* physics.radiative_transfer (physics.py:312-365): tau[S][N] -> tb_out[S];
 * filling_factor has N entries (replicate a scalar). */
int chi_radiative_transfer(chi_handle* h, int S, int N, const double* tau, const double* tspin,
                           const double* filling_factor, double bg_temp, double* tb_out);
/* Their vector-Jacobian products (what PyTensor's grad of physics.py:265-309 and
 * physics.py:312-365 yields): gbar[S][N] -> gradients w.r.t. velocity, fwhm2, fwhm_L (N each);
 * gbar[S] -> g_tau[S][N], g_tspin[N], g_filling_factor[N] (N <= 64). */
int chi_pseudo_voigt_vjp(chi_handle* h, int S, int N, const double* velo_axis, const double* velocity,
                         const double* fwhm2, const double* fwhm_L, const double* gbar,
                         double* g_velocity, double* g_fwhm2, double* g_fwhm_L);
int chi_radiative_transfer_vjp(chi_handle* h, int S, int N, const double* tau, const double* tspin,
                               const double* filling_factor, double bg_temp, const double* gbar,
                               double* g_tau, double* g_tspin, double* g_filling_factor);

/* ---- model-level entry points (need chi_set_spectra) ----------------------- */
/* Likelihood logp and its gradient for B independent evaluations (draws x chains x
 * sightlines).  Replaces one call of PyMC's compiled logp_dlogp_function restricted
 * to the observed Normal nodes: add_likelihood() bodies (absorption_model.py:80-104,
 * emission_model.py:140-169, emission_absorption_model.py:69-119 and the physical
 * equivalents) + pm.Normal logp + PyTensor reverse-mode grad.
 *   theta        [B][CHI_NSLOT][N]
 *   coeff_e/_a   [B][n_baseline] baseline coefficients (NULL = zeros)
 *   sightline    [B] index into the uploaded sightlines (NULL = all 0)
 *   logp_out     [B]
 *   grad_out     [B][CHI_NSLOT][N]   d logp / d theta   (NULL = skip the gradient)
 *   gcoeff_e/_a  [B][n_baseline]     d logp / d coeff   (NULL = skip) */
int chi_logp_grad(chi_handle* h, int64_t B, const double* theta, const double* coeff_e,
                  const double* coeff_a, const int32_t* sightline, double* logp_out,
                  double* grad_out, double* gcoeff_e, double* gcoeff_a);
int chi_logp_grad_dev(chi_handle* h, int64_t B, const double* theta, const double* coeff_e,
                      const double* coeff_a, const int32_t* sightline, double* logp_out,
                      double* grad_out, double* gcoeff_e, double* gcoeff_a);

/* Packed form of chi_logp_grad[_dev]: the caller's blocks carry only the rows it needs,
 * rows[B][n_rows][N] with row r holding slot slot_of_row[r] (distinct slots; a slot that is not
 * listed reads as 0 and its gradient is dropped), grad_rows[B][n_rows][N] likewise.  A model kind
 * has 6-9 free RVs per cloud, so listing exactly those (Engine.row_slots in the Python host
 * side) moves 10-40 % fewer bytes over PCIe than the CHI_NSLOT-row block.  Replaces the same
 * reference call as chi_logp_grad (PyMC's logp_dlogp_function over the free RVs). */
int chi_logp_grad_rows(chi_handle* h, int64_t B, int32_t n_rows, const int32_t* slot_of_row,
                       const double* rows, const double* coeff_e, const double* coeff_a,
                       const int32_t* sightline, double* logp_out, double* grad_rows_out,
                       double* gcoeff_e_out, double* gcoeff_a_out);
/* Asynchronous form of chi_logp_grad_rows: enqueues the uploads, kernels and downloads and returns
 * at once with a ticket; chi_wait(h, ticket) blocks until that call's outputs are in host memory
 * (ticket < 0: everything enqueued on the handle).  All host buffers must be page-locked
 * (chi_host_alloc) and stay untouched until the wait returns.  Calls complete in the order they
 * were issued; keeping two in flight (double-buffered inputs / outputs) hides the pipeline's fill
 * and drain of one call behind the kernels of the other. */
int chi_logp_grad_rows_async(chi_handle* h, int64_t B, int32_t n_rows, const int32_t* slot_of_row,
                             const double* rows, const double* coeff_e, const double* coeff_a,
                             const int32_t* sightline, double* logp_out, double* grad_rows_out,
                             double* gcoeff_e_out, double* gcoeff_a_out, int64_t* ticket_out);
int chi_wait(chi_handle* h, int64_t ticket);
/* chi_logp_grad_rows for consumers of TOTALS (a survey-total or hierarchical logp, a sharded job that
 * gathers per-shard sums): the per-draw gradient is computed and stays on the device; what comes back is
 * logp_out[B] (NULL = not wanted) and sums_out[2 + n_rows * N] = {sum of logp, sum of every gradient
 * entry [n_rows][N], number of draws with a non-finite logp}, both sums over the finite draws only.
 * 28 MB less device->host traffic per 1e5 draws at configs[1]; deterministic (per-piece sums in fixed
 * order, pieces added in order).  The asynchronous form completes its sums in chi_wait. */
int chi_logp_sums_rows(chi_handle* h, int64_t B, int32_t n_rows, const int32_t* slot_of_row,
                       const double* rows, const double* coeff_e, const double* coeff_a,
                       const int32_t* sightline, double* logp_out, double* sums_out);
int chi_logp_sums_rows_async(chi_handle* h, int64_t B, int32_t n_rows, const int32_t* slot_of_row,
                             const double* rows, const double* coeff_e, const double* coeff_a,
                             const int32_t* sightline, double* logp_out, double* sums_out, int64_t* ticket_out);
int chi_logp_grad_rows_dev(chi_handle* h, int64_t B, int32_t n_rows, const int32_t* slot_of_row,
                           const double* rows, const double* coeff_e, const double* coeff_a,
                           const int32_t* sightline, double* logp_out, double* grad_rows_out,
                           double* gcoeff_e_out, double* gcoeff_a_out);

/* Predicted spectra mu_e[B][S_e], mu_a[B][S_a] (baseline included): the `mu` of
 * the pm.Normal nodes, i.e. what sample_posterior_predictive evaluates per draw.
 * Either output may be NULL. */
int chi_spectra(chi_handle* h, int64_t B, const double* theta, const double* coeff_e,
                const double* coeff_a, double* mu_e_out, double* mu_a_out);
int chi_spectra_dev(chi_handle* h, int64_t B, const double* theta, const double* coeff_e,
                    const double* coeff_a, double* mu_e_out, double* mu_a_out);
/* The same spectra in single precision (the optional fp32 path of posterior / prior predictive
 * generation: half the output bytes, channel loop on the FP32 pipe and the SFU).  theta and the
 * parameter maps stay fp64; each spectrum agrees with chi_spectra to 1e-5 of its largest value
 * (compensated v - c, cancellation-free 1 - exp(-tau)).  logp and gradients have no fp32 form:
 * at realistic signal-to-noise the residual y - mu loses 3-4 digits. */
int chi_spectra_f32(chi_handle* h, int64_t B, const double* theta, const double* coeff_e,
                    const double* coeff_a, float* mu_e_out, float* mu_a_out);
int chi_spectra_f32_dev(chi_handle* h, int64_t B, const double* theta, const double* coeff_e,
                        const double* coeff_a, float* mu_e_out, float* mu_a_out);

/* Predictive DRAWS of the observed nodes, y = mu + sigma z with z ~ N(0, 1): what PyMC's
 * sample_posterior_predictive / sample_prior_predictive return for the pm.Normal(mu, sigma =
 * data.noise, observed = ...) nodes of the add_likelihood() bodies (emission_model.py:164-169,
 * absorption_model.py:99-104, ...).  sigma is the noise of sightline[b] (sightline 0 if NULL).
 * z comes from a counter-based generator keyed by (seed, draw_offset + b, dataset, channel):
 * shards and sub-batches that pass their offset reproduce the single-call draws bit for bit.
 * fp64: 53-bit uniforms, Box-Muller in fp64; _f32: 24-bit uniforms, fp32 (chi_spectra_f32 + noise). */
int chi_predictive(chi_handle* h, int64_t B, const double* theta, const double* coeff_e,
                   const double* coeff_a, const int32_t* sightline, uint64_t seed, int64_t draw_offset,
                   double* y_e_out, double* y_a_out);
int chi_predictive_dev(chi_handle* h, int64_t B, const double* theta, const double* coeff_e,
                       const double* coeff_a, const int32_t* sightline, uint64_t seed, int64_t draw_offset,
                       double* y_e_out, double* y_a_out);
int chi_predictive_f32(chi_handle* h, int64_t B, const double* theta, const double* coeff_e,
                       const double* coeff_a, const int32_t* sightline, uint64_t seed, int64_t draw_offset,
                       float* y_e_out, float* y_a_out);
int chi_predictive_f32_dev(chi_handle* h, int64_t B, const double* theta, const double* coeff_e,
                           const double* coeff_a, const int32_t* sightline, uint64_t seed, int64_t draw_offset,
                           float* y_e_out, float* y_a_out);

/* Vector-Jacobian product of chi_spectra: given cotangents gbar_e[B][S_e],
 * gbar_a[B][S_a] (NULL = zeros) returns grad_out[B][CHI_NSLOT][N] and the baseline
 * coefficient gradients.  This is the L_op of the spectra Op. */
int chi_spectra_vjp(chi_handle* h, int64_t B, const double* theta, const double* coeff_e,
                    const double* coeff_a, const double* gbar_e, const double* gbar_a,
                    double* grad_out, double* gcoeff_e, double* gcoeff_a);
int chi_spectra_vjp_dev(chi_handle* h, int64_t B, const double* theta, const double* coeff_e,
                        const double* coeff_a, const double* gbar_e, const double* gbar_a,
                        double* grad_out, double* gcoeff_e, double* gcoeff_a);

/* The reference's pm.Deterministic cloud quantities for each draw:
 * det_out[B][CHI_NDET][N] (hi_model.py:138-162, hi_physical_model.py:153-220 and the
 * subclasses' add_priors). */
int chi_deterministics(chi_handle* h, int64_t B, const double* theta, double* det_out);

/* ---- total model logp in PyMC's unconstrained space ----------
 * What PyMC's compiled logp_dlogp_function returns for the whole model: the likelihood
 * above + the prior density of every free RV + the log-Jacobian of PyMC's default
 * transforms, as a function of the UNCONSTRAINED value variables
 *   fwhm2_norm_log__, *_norm_log__ (HalfNormal), nth_fwhm_1pc_interval__ : x = exp(u)
 *   tkin_factor*_logodds__, fwhm2_thermal_fraction*_logodds__, filling_factor_norm_interval__ :
 *                                                                          x = sigmoid(u)
 *   Normal RVs (velocity_norm, log10_nHI_norm, log10_n_alpha_norm, log10_wt_ff_*) untransformed
 * in the same slots as theta.  Distributions: hi_model.py:95-134, hi_physical_model.py:111-149,
 * absorption_model.py:46-61, emission_model.py:62-110, emission_absorption_model.py:50-55,
 * absorption_physical_model.py:45-56, emission_physical_model.py:61-112,
 * emission_absorption_physical_model.py:43-52; densities/transforms are PyMC's (third-party). */
typedef struct chi_priors {
  double beta[2];               /* prior_tkin_factor | prior_fwhm2_thermal_fraction (alpha, beta) */
  double nth_fwhm_1pc[2];       /* prior_nth_fwhm_1pc (mu, sigma), TruncatedNormal lower = 0 */
  double sigma_log10_NHI;       /* prior_sigma_log10_NHI of the coupled Normal (log10_wt_ff_*) */
  double baseline_sigma_e[8];   /* prior_baseline_coeffs["emission"] (CHI_MAX_BASELINE entries) */
  double baseline_sigma_a[8];   /* prior_baseline_coeffs["absorption"] */
} chi_priors;
int chi_set_priors(chi_handle* h, const chi_priors* priors);
/* u[B][CHI_NSLOT][N] unconstrained; grad_out is d logp / d u (for a per-draw scalar RV replicated
 * over clouds -- fwhm_L of the physical models -- sum the slot over clouds).  Not available for
 * CHI_MODEL_PHYSICS.  Needs chi_set_spectra and chi_set_priors. */
int chi_model_logp_grad(chi_handle* h, int64_t B, const double* u, const double* coeff_e,
                        const double* coeff_a, const int32_t* sightline, double* logp_out,
                        double* grad_out, double* gcoeff_e, double* gcoeff_a);
int chi_model_logp_grad_dev(chi_handle* h, int64_t B, const double* u, const double* coeff_e,
                            const double* coeff_a, const int32_t* sightline, double* logp_out,
                            double* grad_out, double* gcoeff_e, double* gcoeff_a);
/* PyMC's default transforms on a parameter block: direction +1 maps unconstrained -> constrained
 * (theta usable by chi_logp_grad / chi_spectra), -1 the inverse.  in/out: [B][CHI_NSLOT][N] host. */
int chi_transform(chi_handle* h, int64_t B, int direction, const double* in, double* out);

/* ---- chain-batched NUTS -------------------------------------------
 * Replaces PyMC's process-per-chain pm.sample(...) loop (optimization.ipynb cell 12) for the model
 * logp above: all `chains` live on the device and one batched logp+grad evaluation per tick serves
 * the leapfrog step in flight of every chain (each chain runs its own tree / iteration state
 * machine, see chi_nuts_config::lock_step); multinomial NUTS with iterative U-turn checks, dual-averaging step size and
 * windowed diagonal mass adaptation during the n_tune warm-up iterations.  Draws come back in
 * the unconstrained space (chi_transform maps them), host pointers:
 *   init_u[chains][CHI_NSLOT][N], init_ce/_ca[chains][n_baseline]   starting points (finite logp!)
 *   draws_u[n_draws][chains][CHI_NSLOT][N], draws_ce/_ca[n_draws][chains][n_baseline]
 *   stat_*[n_draws][chains]: mean acceptance probability, tree depth, divergence flag, model logp,
 *   step size (any may be NULL). */
typedef struct chi_nuts_config {
  int64_t chains;
  int32_t n_tune, n_draws;
  int32_t max_treedepth;   /* <= 12; 0 -> 10 */
  int32_t pool_mass;       /* 1: the diagonal mass matrix of a window is estimated from all chains'
                              samples (chains must share one posterior: ignored when `sightline`
                              is given); 0: per chain, as PyMC does */
  uint64_t seed;
  double target_accept;    /* 0 -> 0.8 */
  double init_step;        /* 0 -> 0.25 / n_free^(1/4) (PyMC's default scaling) */
  int32_t lock_step;       /* 0: every chain runs its own tree / iteration state machine, one batched
                              evaluation per tick serves them all (default); 1: all chains wait for
                              the deepest tree of every iteration */
  int32_t reserved;
} chi_nuts_config;
int chi_nuts_sample(chi_handle* h, const chi_nuts_config* cfg, const double* init_u, const double* init_ce,
                    const double* init_ca, const int32_t* sightline, double* draws_u, double* draws_ce,
                    double* draws_ca, double* stat_accept, int32_t* stat_depth, int32_t* stat_diverged,
                    double* stat_logp, double* stat_step);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink only to COLLECT ---------------
 * The path has no exchange step (draws, chains and sightlines are independent: SURVEY 8e; the
 * reference parallelises over chains with one process each, optimization.ipynb cell 12).  A sharded
 * job evaluates its shard with the entry points above and then gathers the per-shard sums of logp and
 * gradient (or rows of per-draw results) with the calls below.  The handle owns the communicator
 * and a communication stream: a collective waits for what the handle's stream has enqueued so far and
 * runs next to whatever is enqueued afterwards; chi_comm_join makes the handle's stream wait for it.
 * NCCL is loaded at run time (libnccl.so.2), so single-GPU use has no dependency on it.
 *   rank 0: chi_comm_unique_id(id) -> the caller ships the 128 bytes to every rank (any side
 *   channel: torch.distributed, MPI, a file) -> every rank: chi_comm_init(h, id, rank, n_ranks). */
#define CHI_COMM_ID_BYTES 128
int chi_comm_unique_id(void* id_out);
int chi_comm_init(chi_handle* h, const void* id, int rank, int n_ranks);
int chi_comm_destroy(chi_handle* h);
int chi_comm_rank(const chi_handle* h);
int chi_comm_size(const chi_handle* h);
/* Sums over the FINITE draws of this shard (device pointers): sums_out[0] = sum logp,
 * sums_out[1 .. width] = sum over draws of grad[b][0 .. width) (width = n_rows * N of the block the
 * gradient came in, <= 256), sums_out[1 + width] = number of draws with a non-finite logp (their
 * gradients are left out too).  Deterministic (fixed summation order).  Runs on the communication
 * stream behind everything enqueued on the handle's stream. */
int chi_shard_sums_dev(chi_handle* h, int64_t B, int32_t width, const double* logp, const double* grad,
                       double* sums_out);
/* recv[n_ranks][count] <- every rank's send[count] (ncclAllGather); buf[count] <- sum over ranks
 * (ncclAllReduce).  Device pointers; with one rank / no communicator the gather is a copy. */
int chi_allgather_dev(chi_handle* h, const double* send, double* recv, int64_t count);
int chi_allreduce_sum_dev(chi_handle* h, double* buf, int64_t count);
/* The handle's stream waits for the communication stream / the host waits for it; the stream itself
 * (cudaStream_t as void*) for callers that order their own work with events. */
int chi_comm_join(chi_handle* h);
int chi_comm_synchronize(chi_handle* h);
void* chi_comm_stream(chi_handle* h);

/* ---- measurement helpers --------------------------------------------------- */
/* Milliseconds spent in the kernels of the last model-level call on this handle
 * (CUDA events on the handle's stream; synchronises). */
int chi_last_kernel_ms(chi_handle* h, float* ms_out);
/* Kernel milliseconds of the four stages {cloud map, emission moments, absorption
 * moments, epilogue} of the n most recent logp/VJP batches enqueued on this handle,
 * oldest first: ms_out[n][4].  CUDA events on the handle's stream, recorded on every
 * batch; the library keeps the last 64.  Synchronises on the newest batch asked for. */
int chi_stage_ms(chi_handle* h, int n, float* ms_out);
/* Number of logp / VJP batches (kernel pipelines) enqueued so far; large device batches and the
 * pipelined host path enqueue several per call. */
int64_t chi_stage_count(const chi_handle* h);
/* Number of kernel launches issued by this handle since creation. */
int64_t chi_launch_count(const chi_handle* h);
/* Tuning / test knobs of a handle (no reference counterpart; results stay within the parity
 * tolerance whatever the setting, which is what the tests use them for).
 *   CHI_OPT_SMALL_PATH    0 = by batch size (default), 1 = never, 2 = always: the one-launch
 *                         evaluation of small batches (csrc/small.cuh) instead of the
 *                         map -> channel kernels -> epilogue pipeline
 *   CHI_OPT_SMALL_NL      clouds per thread of that kernel (0 = automatic, 1, 2)
 *   CHI_OPT_SMALL_CHUNKS  its blocks per draw (0 = automatic) */
#define CHI_OPT_SMALL_PATH 0
#define CHI_OPT_SMALL_NL 1
#define CHI_OPT_SMALL_CHUNKS 2
#define CHI_OPT_SMALL_STAMPS 3 /* development aid: device pointer to 16 int64 receiving SM-clock offsets
                                  of block 0's phases (0 = off) */
#define CHI_OPT_NUTS_TICK 4    /* how chi_nuts_sample runs a tick of its free-running chains:
                                  0 = by model size (default): 3 for small models, 2 otherwise;
                                  1 = a CUDA graph of kernels per tick, the tick passing every phase through global
                                      memory (round-1 form; same draws as 2 bit for bit);
                                  2 = that graph with the common tick staged in shared memory;
                                  3 = one resident block per chain runs whole leapfrog steps tick after tick inside
                                      one kernel (csrc/nuts_chain.cuh), 256 threads per chain, whenever the model fits
                                      (<= 16 clouds); same draws as 2 with CHI_OPT_SMALL_PATH 2, CHI_OPT_SMALL_CHUNKS 1;
                                  4 = that kernel with 512 threads per chain where it can (<= 8 clouds, chains <= SMs);
                                  5 = that kernel with the 2..8 blocks of a chain as a thread-block cluster (for models
                                      whose channel sums want more than one SM per chain; needs chains <= SMs);
                                  6 = that kernel with the chain's state machine on a ninth warp: the next point is
                                      published first and the bookkeeping of the finished leaf runs next to its
                                      evaluation (same draws as 3 bit for bit; the default for small models when
                                      chains <= SMs) */
#define CHI_OPT_DEV_OVERLAP 5  /* 1 = consecutive chi_logp_grad[_rows]_dev calls of large batches alternate between
                                  two internal lanes and may overlap (map / epilogue of one call next to the channel
                                  kernel of the other); each call still starts after everything enqueued on the
                                  handle's stream before it, but its outputs are ordered into the handle's stream
                                  by chi_dev_join() or chi_synchronize(), not by the call.  0 (default): every call
                                  runs on the handle's stream. */
#define CHI_OPT_EA_ONE_PASS 6  /* 1 = large Gaussian batches on a shared axis run the fused channel kernel as ONE
                                  launch with the overflow branch of 2^y in it (the form small batches always use);
                                  0 (default) = two launches, see moments_ea.cuh MODE.  Same results bit for bit. */
#define CHI_OPT_SMALL_CLUSTER 7 /* 0 (default) = when the one-launch evaluation of a small batch runs 2..8 blocks per
                                  draw they form a thread-block cluster and meet in distributed shared memory;
                                  1 = they meet in global memory (records + a counter per draw).  Same results
                                  bit for bit. */
#define CHI_OPT_COUNT 8
int chi_set_option(chi_handle* h, int option, int64_t value);
/* CHI_OPT_DEV_OVERLAP: make the handle's stream wait for the device-resident calls issued so far. */
int chi_dev_join(chi_handle* h);
/* Own-microbenchmark peaks of the device the handle is on (the driver's
 * MEASURED_PEAKS.json has no CUDA-core FP64 figure): dependent-free DFMA chains. */
int chi_microbench_fp64(chi_handle* h, double* tflops_out);
/* FP32 FMA peak (TFLOP/s, dependent-free FFMA chains) and SFU peak (1e12 MUFU.EX2 per second): the
 * denominators of the single-precision spectra kernels' roofline. */
int chi_microbench_fp32(chi_handle* h, double* tflops_out, double* tera_sfu_per_s_out);

#ifdef __cplusplus
}
#endif
#endif /* CHI_H */
