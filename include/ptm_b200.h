/* ptm_b200.h -- C ABI of the B200-native PTM log-density hot path.
 *
 * Drop-in boundary for liesel-devs/liesel-ptm's data-parallel hot path (SURVEY.md
 * section 8b).  Plain C: no C++ types, no exceptions across the boundary.  Every
 * per-call pointer is a DEVICE pointer owned by the caller; only
 * ptm_problem_create() takes HOST pointers (it copies the static data into HBM
 * once).  All work is enqueued on the caller's stream (a cudaStream_t passed as
 * void*), with no hidden synchronisation and no allocation after create /
 * workspace sizing.  Return value 0 = OK, non-zero = PtmStatus; the message of
 * the last failure on the calling thread is ptm_last_error().
 *
 * Reference interfaces replaced (paths under /root/reference/liesel_ptm/):
 *   ptm_logp_* / ptm_logp_grad_*   model.log_prob(state) and jax.grad of it as the
 *                                  MCMC kernels call them: logprob.py:36-41,50-69,
 *                                  104-140; iwls_fixed.py:91-116,159-186; the
 *                                  density itself is dist.py:185-188,339-395,
 *                                  718-740 over bspline/util.py:91-103,
 *                                  bspline/ptm.py:412-478, bspline/approx.py:
 *                                  418-522, gam/var.py:157-162, predictor.py:
 *                                  19-21,369-371,583-585, gam/dist.py:56-70,
 *                                  var.py:77-82,106-115.
 *   ptm_basis_*                    Basis.bspline / basis_matrix: bspline/approx.py:
 *                                  74-210 (called at gam/var.py:1001-1002).
 *   ptm_transform_*                TransformationSpline.dot_and_deriv_n_fullbatch:
 *                                  bspline/util.py:91-103.
 *   ptm_quadform_*                 beta' K beta of the Gibbs scale updates:
 *                                  gam/kernel.py:29-39, kernel.py:99-111.
 *   ptm_xtwx_* / ptm_loc_precision_*  GaussianLocCholInfo.working_weights /
 *                                  precision / chol_info: iwls_proposals.py:55-89
 *                                  (scale terms, W == 2: 118-144).
 *   ptm_loc_precision_all_*        the same for every location term of a sampler sweep
 *                                  (the weights depend only on the scale predictor).
 *   ptm_intercept_stats_*          LocationIntercept / LogScaleIntercept
 *                                  .compute_intercept: predictor.py:80-100,126-130.
 *   ptm_pointwise_*                EvaluatePTM._lppdi / .lppdi / .waic:
 *                                  model/evaluate.py:106-133,177-236.
 *
 * Parameter vector of ONE chain (row of the [C, P] row-major `params` array, the
 * layout JAX hands over after ravel_pytree):
 *     [ beta0, gamma0,                     loc / log-scale intercepts
 *       beta_0[p_0], ..., beta_{T-1}[p_{T-1}],   smooth-term coefficients
 *       delta_z[nparam-1],                 latent transformation coefficients
 *       tau_0, ..., tau_{T-1},             smoothing scales of the terms
 *       tau_delta ]                        scale of the delta_z prior
 * `grad` has the same layout.  logp = sum_i loglik_i + sum_t prior(beta_t | tau_t)
 * + prior(delta_z | tau_delta); the O(1)-per-chain hyper-priors on tau^2 stay in
 * host JAX.
 */
#ifndef PTM_B200_H_
#define PTM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  PTM_OK = 0,
  PTM_ERR_NULL = 1,       /* null pointer argument */
  PTM_ERR_SHAPE = 2,      /* invalid / unsupported shape */
  PTM_ERR_DTYPE = 3,      /* entry point does not match the problem's dtype */
  PTM_ERR_RANGE = 4,      /* covariate outside its interior knot range (the
                             reference raises ValueError, approx.py:197-206) */
  PTM_ERR_CUDA = 5,       /* a CUDA call failed; see ptm_last_error() */
  PTM_ERR_WORKSPACE = 6   /* workspace too small */
} PtmStatus;

enum { PTM_F32 = 0, PTM_F64 = 1 };
enum { PTM_LOC = 0, PTM_SCALE = 1 };
enum { PTM_MAX_TERMS = 24, PTM_MAX_NBASES = 32, PTM_MAX_NPARAM = 40 };

/* Static data of one model.  All pointers are HOST pointers; element type is
 * float when dtype == PTM_F32 and double when dtype == PTM_F64. */
typedef struct {
  int32_t dtype;               /* PTM_F32 | PTM_F64 */
  int64_t n;                   /* observations */
  const void* y;               /* [n] response */
  int32_t n_terms;             /* T smooth terms (<= PTM_MAX_TERMS) */
  const void* const* x;        /* [T] -> [n] covariate of term t */
  const int32_t* nbases;       /* [T] nb_t: B-spline bases (4 <= nb_t <= 32) */
  const int32_t* ncoef;        /* [T] p_t: columns of Z_t */
  const int32_t* predictor;    /* [T] PTM_LOC | PTM_SCALE */
  const void* const* knots;    /* [T] -> [nb_t + 4] equidistant knots */
  const void* const* Z;        /* [T] -> [nb_t, p_t] row-major reparam matrix */
  const void* const* K;        /* [T] -> [p_t, p_t] penalty after reparams */
  const int32_t* rank;         /* [T] rank of K_t */
  int32_t nparam;              /* shape parameters of h (J = nparam + 1 bases) */
  const void* trafo_knots;     /* [nparam + 5] PTMKnots(a, b, nparam).knots (see trafo_variant) */
  const void* trafo_Z;         /* [nparam, nparam - 1] delta = Z delta_z */
  double trafo_lambda;         /* PTMSpline eps: transition width / (b - a) */
  const void* trafo_grid;      /* [1002] BSplineApprox.grid, or NULL to rebuild it
                                  with the restated jnp.linspace formula */
  const int32_t* noncentered;  /* [T] or NULL: 1 = term.reparam_noncentered() (gam/var.py:176-197):
                                  the parameter is the LATENT coefficient, coef = tau_t * latent,
                                  prior latent ~ N(0, inv(K)) (scale 1); the IWLS factor is
                                  tau_t * chol (iwls_proposals.py:82-89) */
  int32_t trafo_noncentered;   /* PTMCoef(noncentered=True), var.py:100-107: delta = Z (tau_delta *
                                  delta_z), delta_z ~ N(0, 1) */
  int32_t trafo_variant;       /* PTMDist(bspline=...), model/model.py:98-131:
                                  PTM_TRAFO_PTM (0): PTMSpline, trafo_knots [nparam + 5];
                                  PTM_TRAFO_ONION (1): OnionSpline (bspline/onion.py:13-138), trafo_knots
                                  [nparam + 11] = OnionKnots(a, b, nparam).knots, J = nparam + 7 bases
                                  (nparam <= 34), identity off the core, trafo_lambda unused;
                                  PTM_TRAFO_IDENTITY (2): GaussianPseudoTransformationDist
                                  (dist.py:766-850, LocScalePTM.new_gaussian model/model.py:386-417):
                                  h = identity; the delta_z / tau_delta entries of the parameter
                                  rows are ignored, carry no prior, and get a zero gradient */
} PtmProblemDesc;

enum { PTM_TRAFO_PTM = 0, PTM_TRAFO_ONION = 1, PTM_TRAFO_IDENTITY = 2 };

typedef struct PtmProblem PtmProblem; /* opaque, immutable after create (tuning knobs in the
                                          environment are read once, at create) */

int ptm_problem_create(const PtmProblemDesc* desc, PtmProblem** out);
int ptm_problem_destroy(PtmProblem* prob);

/* P = 2 + sum_t p_t + (nparam - 1) + T + 1. */
int64_t ptm_num_params(const PtmProblem* prob);
/* Offsets into one chain's parameter row; any output pointer may be NULL.
 * beta_off has T entries, tau_off has T entries. */
int ptm_param_layout(const PtmProblem* prob, int64_t* beta0_off, int64_t* gamma0_off,
                     int64_t* beta_off, int64_t* delta_z_off, int64_t* tau_off,
                     int64_t* tau_delta_off);

/* Bytes of caller-owned device scratch needed for C chains (any entry point). */
size_t ptm_workspace_bytes(const PtmProblem* prob, int64_t n_chains);

/* Observation-axis sharding across GPUs: a NEW handle that evaluates the
 * observations [obs_begin, obs_end) of `base` only.  Handles are immutable, so the shard is
 * a create-time property of the view; the view borrows base's device arrays (base must
 * outlive it; destroy the view with ptm_problem_destroy).  With add_priors == 0 the per-chain
 * prior terms are left out of logp / grad so that exactly one rank adds them to the
 * all-reduced sum.  What a view returns is the SHARD-LOCAL, additive part of:
 *   ptm_logp_* / ptm_logp_grad_* / ptm_logp_grad_block_*   likelihood sums (+ priors if asked)
 *   ptm_xtwx_*                                             X' W X over the shard
 *   ptm_intercept_stats_*                                  the five sums
 *   ptm_pointwise_*                                        records of the shard's observations
 * ptm_loc_precision_* / ptm_loc_precision_all_* return PTM_ERR_SHAPE on a view (or with
 * add_priors == 0): penalty, ridge and Cholesky are not additive -- all-reduce ptm_xtwx and
 * finish on the sum. */
int ptm_problem_shard_view(const PtmProblem* base, int64_t obs_begin, int64_t obs_end,
                           int32_t add_priors, PtmProblem** out);
/* The observation range and prior flag of a handle ([0, n), 1 for a created problem). */
int ptm_problem_shard(const PtmProblem* prob, int64_t* obs_begin, int64_t* obs_end,
                      int32_t* add_priors);

/* logp[c] for C chains.  params [C, P] row-major, logp [C]. */
int ptm_logp_f64(const PtmProblem* prob, const double* params, int64_t n_chains,
                 double* logp, void* workspace, size_t workspace_bytes, void* stream);
int ptm_logp_f32(const PtmProblem* prob, const float* params, int64_t n_chains,
                 float* logp, void* workspace, size_t workspace_bytes, void* stream);

/* logp[c] and grad[c, :] in one fused pass over the observations. */
int ptm_logp_grad_f64(const PtmProblem* prob, const double* params, int64_t n_chains,
                      double* logp, double* grad, void* workspace,
                      size_t workspace_bytes, void* stream);
int ptm_logp_grad_f32(const PtmProblem* prob, const float* params, int64_t n_chains,
                      float* logp, float* grad, void* workspace,
                      size_t workspace_bytes, void* stream);

/* The same results as ptm_logp_grad_* written into ONE contiguous block [C, 1 + P] row-major:
 * block[c, 0] = logp[c], block[c, 1:] = grad[c, :].  This is the payload of the
 * observation-axis all-reduce: the epilogue kernel writes it in place, so the
 * collective runs on the buffer without a repacking step. */
int ptm_logp_grad_block_f64(const PtmProblem* prob, const double* params, int64_t n_chains,
                            double* block, void* workspace, size_t workspace_bytes, void* stream);
int ptm_logp_grad_block_f32(const PtmProblem* prob, const float* params, int64_t n_chains,
                            float* block, void* workspace, size_t workspace_bytes, void* stream);

/* Banded design of term `term`: interval[i] = j with knots[j] <= x_i < knots[j+1]
 * (int32, bit-exact with the reference's order-0 indicator) and the four non-zero
 * cubic values of columns j-3..j in values[i, 0..3]. */
int ptm_basis_f64(const PtmProblem* prob, int32_t term, int32_t* interval,
                  double* values, void* stream);
int ptm_basis_f32(const PtmProblem* prob, int32_t term, int32_t* interval,
                  float* values, void* stream);

/* h(r) and h'(r) of the transformation for raw shape coefficients delta
 * [C, nparam] at r [m]: z, dz are [C, m].  cell (optional, [C? no: m]) receives
 * the grid-cell index searchsorted(grid, r, 'right') - 1 for core points and -1
 * elsewhere. */
int ptm_transform_f64(const PtmProblem* prob, const double* delta, int64_t n_chains,
                      const double* r, int64_t m, double* z, double* dz,
                      int32_t* cell, void* stream);
int ptm_transform_f32(const PtmProblem* prob, const float* delta, int64_t n_chains,
                      const float* r, int64_t m, float* z, float* dz,
                      int32_t* cell, void* stream);

/* Chain-batched IWLS information of a LOC term: xtwx [C, p, p] = X' diag(w) X with
 * w_i = 1/max(sigma_i, sqrt(eps))^2 and X = B_t Z_t (iwls_proposals.py:55-74).  For a SCALE
 * term w == 2 (iwls_proposals.py:118-119).  (The score X'r of an IWLS step is the beta_t
 * block of ptm_logp_grad_*; there is no separate output for it.) */
int ptm_xtwx_f64(const PtmProblem* prob, int32_t term, const double* params,
                 int64_t n_chains, double* xtwx, void* workspace,
                 size_t workspace_bytes, void* stream);
int ptm_xtwx_f32(const PtmProblem* prob, int32_t term, const float* params,
                 int64_t n_chains, float* xtwx, void* workspace,
                 size_t workspace_bytes, void* stream);

/* precision [C, p, p] = XtWX + K / max(tau, sqrt(eps))^2 + 1e-6 mean(diag) I and
 * its lower Cholesky factor chol [C, p, p] (iwls_proposals.py:61-89). */
int ptm_loc_precision_f64(const PtmProblem* prob, int32_t term, const double* params,
                          int64_t n_chains, double* precision, double* chol,
                          void* workspace, size_t workspace_bytes, void* stream);
int ptm_loc_precision_f32(const PtmProblem* prob, int32_t term, const float* params,
                          int64_t n_chains, float* precision, float* chol,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Precision and Cholesky factor of ALL location terms from ONE sweep over the
 * observations (the working weights depend only on the scale predictor, so the loc
 * terms share them).  Term blocks follow each other in model order: block t is
 * [C, p_t, p_t] row-major.  chol may be NULL. */
int ptm_loc_precision_all_f64(const PtmProblem* prob, const double* params, int64_t n_chains,
                              double* precision, double* chol, void* workspace,
                              size_t workspace_bytes, void* stream);
int ptm_loc_precision_all_f32(const PtmProblem* prob, const float* params, int64_t n_chains,
                              float* precision, float* chol, void* workspace,
                              size_t workspace_bytes, void* stream);

/* Chain-batched OBSERVED information of one coefficient block,
 *     info[c] = - jacfwd(grad(log_prob))   restricted to the block       (logprob.py:104-113),
 * with the reference's post-processing and Cholesky factor.  Replaces FlatLogProb.hessian as
 * CholInfo / PTMCholInfo / PTMCholInfoFixed / ObservedCholInfoOrIdentity
 * (iwls_proposals.py:168-184, 227-245, 279-298, 333-364), IWLSKernelFixed.init_state
 * (iwls_fixed.py:118-143) and kernel.IWLSKernel._chol_info (kernel.py:240-258) use it.
 *   block: smooth-term index t (coefficients beta_t, d = p_t) or PTM_BLOCK_DELTA_Z (the latent
 *          transformation coefficients, d = nparam - 1)
 *   mode 0: info as computed;  chol = cholesky(sym(info))
 *        1: info = make_pd(info)  [symmetrise, raise the spectrum to >= 1e-6]   (PTMCholInfo(Fixed),
 *           IWLSKernelFixed.init_state);  chol of that
 *        2: info = make_pd(info + 1e-6 I)                           (CholInfo);  chol of that
 *        3: info as computed;  chol = cholesky(info + I)            (kernel.IWLSKernel._chol_info)
 *        4: info + 1e-6 I;  chol of that, the identity when it has NaNs
 *           (ObservedCholInfoOrIdentity.chol_info_or_identity)
 * info, chol: [C, d, d] row-major; chol may be NULL.  A failed factorisation gives NaN like
 * jnp.linalg.cholesky (modes 0-3).  The second-order terms follow jax's semantics for the
 * reference's custom_jvp (inner grad: the declared rule; outer jacfwd: the rule body
 * differentiated as ordinary code, i.e. the true slopes of the lerped tables).
 * On a shard view only mode 0 without chol is allowed (the likelihood part is additive over
 * shards).  Workspace: ptm_hessian_workspace_bytes(prob, n_chains, block). */
#define PTM_BLOCK_DELTA_Z (-1)
size_t ptm_hessian_workspace_bytes(const PtmProblem* prob, int64_t n_chains, int32_t block);
int ptm_hessian_block_f64(const PtmProblem* prob, int32_t block, int32_t mode, const double* params,
                          int64_t n_chains, double* info, double* chol, void* workspace,
                          size_t workspace_bytes, void* stream);
int ptm_hessian_block_f32(const PtmProblem* prob, int32_t block, int32_t mode, const float* params,
                          int64_t n_chains, float* info, float* chol, void* workspace,
                          size_t workspace_bytes, void* stream);

/* quad[c, t] = beta_t' K_t beta_t for every chain and smooth term: the quadratic form of
 * the Gibbs updates of the smoothing variances (gam/kernel.py:29-39, kernel.py:99-111:
 * b + 0.5 * beta' K beta) and of the prior (gam/dist.py:64-70).  quad is [C, T]. */
int ptm_quadform_f64(const PtmProblem* prob, const double* params, int64_t n_chains,
                     double* quad, void* stream);
int ptm_quadform_f32(const PtmProblem* prob, const float* params, int64_t n_chains,
                     float* quad, void* stream);

/* Sufficient statistics of the intercept pseudo-Gibbs updates (predictor.py:80-100,
 * 126-130, 133-175, 239-281), one fused sweep over the handle's observation range.
 * stats is [C, 5] with w_i = exp(-eta_sigma_i), r_i = (y_i - eta_mu_i) w_i, both intercepts
 * at their current values:
 *   stats[c, 0] = sum_i w_i      stats[c, 1] = sum_i r_i      stats[c, 2] = sum_i r_i^2
 *   stats[c, 3] = sum_i w_i^2    stats[c, 4] = sum_i r_i w_i
 * Then, with n observations and m_k = stats[c, k] / n,
 *   beta0_new  = beta0 + m_1 / m_0                            (LocationIntercept)
 *   gamma0_new = gamma0 + log sqrt(var_d),   d = beta0' - beta0,
 *   var_d = m_2 - 2 d m_4 + d^2 m_3 - (m_1 - d m_0)^2          (LogScaleIntercept)
 * where beta0' is the location intercept LogScaleIntercept sees: the reference passes it the
 * full `loc` predictor (predictor.py:604-609), i.e. d = beta0_new - beta0 when the location
 * intercept was updated first in the same sweep, d = 0 when it is evaluated at the current
 * state.  The sums are additive over observation shards (all-reduce before the formulas). */
int ptm_intercept_stats_f64(const PtmProblem* prob, const double* params, int64_t n_chains,
                            double* stats, void* workspace, size_t workspace_bytes, void* stream);
int ptm_intercept_stats_f32(const PtmProblem* prob, const float* params, int64_t n_chains,
                            float* stats, void* workspace, size_t workspace_bytes, void* stream);

/* Number of location / scale term pairs that evaluate the same design (same covariate
 * values, knots and basis count -- the usual "same covariates in both predictors" model,
 * model/model.py:345-420 adds terms to .loc and .scale independently).  ptm_problem_create
 * detects them; the sweeps then stage and gather every such design once.  0 = general
 * layout (also with PTM_NO_SHARED=1 in the environment at create time). */
int ptm_problem_shared_designs(const PtmProblem* prob);

/* Predictive evaluation on NEW data for n_samples posterior samples (params [S, P] like
 * the chain states): y_new [m] and x_new [T, m] (term-major, the problem's term order) are
 * device arrays; z, logpdf, cdf are [S, m] device outputs, each optional (NULL to skip):
 *   z      = h((y - mu) / sigma)                        (dist.py:385-395, 718-740)
 *   logpdf = -z^2/2 - log(2 pi)/2 + log max(h', tiny) - log sigma     (dist.py:185-188)
 *   cdf    = Phi(z)                                      (dist.py:179-181)
 * This is LocScalePTM.init_dist(samples, newdata=...).log_prob / .cdf and summarise_dist
 * (model/model.py:698-790) as one sweep without materialising loc / scale per sample.
 * Covariates outside a term's interior knots give NaN (the reference raises,
 * approx.py:197-206; the host wrapper raises the same ValueError before the call).
 * n_samples <= 65535 per call.  Workspace: ptm_predict_workspace_bytes(prob, n_samples). */
size_t ptm_predict_workspace_bytes(const PtmProblem* prob, int64_t n_samples);
int ptm_predict_f64(const PtmProblem* prob, const double* params, int64_t n_samples,
                    const double* y_new, const double* x_new, int64_t m, double* z, double* logpdf,
                    double* cdf, void* workspace, size_t workspace_bytes, void* stream);
int ptm_predict_f32(const PtmProblem* prob, const float* params, int64_t n_samples,
                    const float* y_new, const float* x_new, int64_t m, float* z, float* logpdf,
                    float* cdf, void* workspace, size_t workspace_bytes, void* stream);

/* Predictive quantiles: yq[s, i] = mu_si + sigma_si * h_s^{-1}(Phi^{-1}(u_i)) for
 * probabilities u [m] (one per new observation row) -- TransformationDist._quantile /
 * inverse_transformation (dist.py:203-209, 575-650). */
int ptm_predict_quantile_f64(const PtmProblem* prob, const double* params, int64_t n_samples,
                             const double* u, const double* x_new, int64_t m, double* yq,
                             void* workspace, size_t workspace_bytes, void* stream);
int ptm_predict_quantile_f32(const PtmProblem* prob, const float* params, int64_t n_samples,
                             const float* u, const float* x_new, int64_t m, float* yq,
                             void* workspace, size_t workspace_bytes, void* stream);

/* Predictive draws: ydraw[s, i] = mu_si + sigma_si * h_s^{-1}(Phi^{-1}(u[s, i])) with one uniform
 * number per (sample, new row), u [S, m] -- TransformationDist._sample_n (dist.py:193-201:
 * u ~ U(eps, 1), then _quantile).  The uniforms come from the caller (jax.random.uniform with
 * the caller's key on the JAX side), so the draw is reproducible from the caller's seed. */
int ptm_predict_sample_f64(const PtmProblem* prob, const double* params, int64_t n_samples,
                           const double* u, const double* x_new, int64_t m, double* ydraw,
                           void* workspace, size_t workspace_bytes, void* stream);
int ptm_predict_sample_f32(const PtmProblem* prob, const float* params, int64_t n_samples,
                           const float* u, const float* x_new, int64_t m, float* ydraw,
                           void* workspace, size_t workspace_bytes, void* stream);

/* r [C, m] with h_c(r) = z for raw shape coefficients delta [C, nparam] and z [m] --
 * TransformationSpline.dot_inverse_n (bspline/util.py:105-122).  DEVIATION from the
 * reference, on purpose: this is the EXACT inverse of the reference's forward function (lerp
 * cells in the core, quadratic transitions, linear tails); the reference interpolates a 1000-point table with interpax and its tests accept
 * 1e-5 .. 1e-4 (tests/test_bspline/test_ptm.py:167-301). */
int ptm_inverse_f64(const PtmProblem* prob, const double* delta, int64_t n_chains,
                    const double* z, int64_t m, double* r, void* stream);
int ptm_inverse_f32(const PtmProblem* prob, const float* delta, int64_t n_chains,
                    const float* z, int64_t m, float* r, void* stream);

/* Pointwise predictive records over n_samples posterior samples -- the streaming WAIC /
 * lppd of the reference (EvaluatePTM._lppdi, .lppdi, .waic: model/evaluate.py:106-133,
 * 177-236) as one sweep.  params is [n_samples, P] like the chain states (draws x chains
 * flattened).  For every observation i of the shard (device arrays of length n, entries
 * outside the shard are left untouched):
 *   lppd[i]  = logsumexp_s log p(y_i | sample s) - log n_samples
 *   pwaic[i] = Var_s log p(y_i | sample s)                       (ddof = 0)
 * with log p(y_i | s) = -z^2/2 - log(2 pi)/2 + log h'(r) - eta_sigma (dist.py:339-346).
 * The aggregates of the WAIC table are sums over i on the caller's side.
 * Workspace: ptm_pointwise_workspace_bytes(prob, n_samples). */
size_t ptm_pointwise_workspace_bytes(const PtmProblem* prob, int64_t n_samples);
int ptm_pointwise_f64(const PtmProblem* prob, const double* params, int64_t n_samples, double* lppd,
                      double* pwaic, void* workspace, size_t workspace_bytes, void* stream);
int ptm_pointwise_f32(const PtmProblem* prob, const float* params, int64_t n_samples, float* lppd,
                      float* pwaic, void* workspace, size_t workspace_bytes, void* stream);

/* Number of kernels the last logp / logp_grad call on this thread launched. */
int ptm_last_launch_count(void);

/* Route the last logp / logp_grad sweep on this thread took: 0 = CUDA-core sweep (k_main);
 * 1 = tcgen05 sweep (float problems: eta = Gamma B' and the coefficient gradients G' B on the
 * tensor cores with the 3xTF32 split, ptm_main_tc.cuh; one launch when sum of nbases <= 224,
 * otherwise groups of terms of one predictor in eta-only passes whose partial predictors go
 * through the workspace).  PTM_MAIN_TC=0 in the environment at create time forces the CUDA-core
 * sweep.  On the tensor-core route ptm_workspace_bytes() includes the dl/deta exchange buffer
 * (8 bytes per observation and chain, 16 in the several-pass form). */
int ptm_last_main_route(void);

/* Route the last information sweep (ptm_xtwx / ptm_loc_precision / ptm_loc_precision_all)
 * on this thread took: 0 = banded CUDA-core sweep; 2 = tcgen05 tensor-core sweep (float
 * problems, location terms, when the terms fit TMEM / shared memory) with the scale
 * predictor evaluated on the tensor cores as well; 1 = tcgen05 contraction with the weights
 * evaluated on CUDA cores.  By default route 2 is taken when its operand tiles can be
 * double-buffered in shared memory, else route 1, else route 0.  The environment variable
 * PTM_XTWX_TC=0/1/2 forces a route (falling back downwards when it does not fit). */
int ptm_last_xtwx_route(void);

/* Measurement hook (bench.py): when enabled, logp / logp_grad calls on this thread
 * record CUDA events around the prologue, the fused sweep and the epilogue on the
 * caller's stream.  ptm_last_kernel_ms() synchronises on the last event and returns
 * the three durations in milliseconds.  Off by default (no events, no sync). */
int ptm_profile_enable(int on);
int ptm_last_kernel_ms(float* pre_ms, float* main_ms, float* post_ms);

const char* ptm_last_error(void);
const char* ptm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PTM_B200_H_ */
