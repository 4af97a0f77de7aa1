/* bsyn.h -- C ABI of the B200-native gp_grid / nointeraction hot path of bayesynergy.
 *
 * This is the drop-in boundary: every entry point replaces one method of the Rcpp module
 * class `rstantools_model_gp_grid` / `rstantools_model_nointeraction`
 * (= rstan::stan_fit<model, ecuyer1988>, reference src/stanExports_gp_grid.cc:10-30,
 *  src/stanExports_nointeraction.cc:10-30), batched over many experiments ("combinations") and chains.
 * INTEGRATION.md shows the Rcpp stub that binds them.
 *
 * Conventions: plain pointers and sizes, caller owns every buffer passed in or out, the
 * library owns the opaque batch handle.  No function throws; each returns BSYN_OK or an
 * error code and bsyn_last_error() gives the text.  There is no CPU fallback: without a CUDA
 * device every compute entry point returns BSYN_ERR_CUDA.
 * All floating point is IEEE double; all indices int32.
 */
#ifndef BSYN_H
#define BSYN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSYN_VERSION 201

enum {
    BSYN_OK = 0,
    BSYN_ERR_ARG = 1,        /* invalid argument / inconsistent sizes */
    BSYN_ERR_UNSUPPORTED = 2,/* grid too large for the compiled kernel variants, etc. */
    BSYN_ERR_CUDA = 3,       /* CUDA runtime error or no device */
    BSYN_ERR_ALLOC = 4,
    BSYN_ERR_DOMAIN = 5      /* bsyn_prepare: monotherapy (zero-concentration) rows are missing */
};

enum { BSYN_MODEL_GP_GRID = 0, BSYN_MODEL_NOINTERACTION = 1 };

/* One experiment = the Stan data block of inst/stan/gp_grid.stan:10-29 (nointeraction.stan:7-23 reads a
 * subset), i.e. the `stan_data` list built at R/bayesynergy.R:167-200; what the generated constructor
 * reads at src/stanExports_gp_grid.h:510-636. */
typedef struct bsyn_problem {
    int32_t n1, n2;          /* number of non-zero dose levels of drug 1 / drug 2 */
    int32_t nmissing, nrep;  /* only used to check N = (n1+n2+n1*n2+1)*nrep - nmissing */
    int32_t N;               /* number of observations */
    const double *y;         /* [N] */
    const int32_t *ii_obs;   /* [N] 1-based index into the column-major (n2+1) x (n1+1) grid f */
    const double *x1;        /* [n1] log10 doses */
    const double *x2;        /* [n2] */
} bsyn_problem;

/* Model options: uniform over a batch, exactly the scalar data fields of the Stan programs. */
typedef struct bsyn_options {
    int32_t model;           /* BSYN_MODEL_* */
    int32_t est_la, robust, pcprior, heteroscedastic;
    int32_t kernel;          /* 1 RBF, 2 Matern, 3 rational quadratic (gp_grid.stan:112-141) */
    int32_t nu_matern;       /* 1: nu=1/2, 2: 3/2, 3: 5/2 */
    int32_t est_alpha;
    double pcprior_hypers[4];
    double rho;              /* LPTN */
    double lambda;           /* heteroscedastic offset */
} bsyn_options;

typedef struct bsyn_batch bsyn_batch;   /* opaque: problems packed and resident in HBM */

/* ---- lifetime ------------------------------------------------------------------------------------- */
/* Replaces: `new(rstantools_model_gp_grid, data, seed, cxxfun)` = the model ctor, once per experiment
 * (src/stanExports_gp_grid.cc:12, src/stanExports_gp_grid.h:489-801).  `device` = CUDA ordinal. */
int bsyn_batch_create(const bsyn_problem *problems, int32_t n_problems, const bsyn_options *opts,
                      int32_t device, bsyn_batch **out);
void bsyn_batch_destroy(bsyn_batch *b);
const char *bsyn_last_error(void);
int bsyn_version(void);

/* ---- layout queries --------------------------------------------------------------------------------
 * Replace: num_pars_unconstrained, param_names/param_dims/param_fnames_oi (cc:16-21, 29). */
int32_t bsyn_num_params(const bsyn_batch *b, int32_t problem);     /* D, unconstrained */
int32_t bsyn_num_outputs(const bsyn_batch *b, int32_t problem);    /* one write_array row (without lp__) */
/* Flat name ("z[3,2]", "CPO[17]", "VUS_syn", ...) of entry `index` of one write_array row of `problem`: what
 * param_names / param_fnames_oi (cc:16-18) report, in the declaration order of src/stanExports_gp_grid.h:2256-2402.
 * The first bsyn_num_params entries are the parameters.  Host-only (no GPU involved). */
int bsyn_output_name(const bsyn_batch *b, int32_t problem, int32_t index, char *buf, int32_t buflen);
/* Parameter blocks: the declared names and dimensions param_names() / param_dims() report (cc:16, 19;
 * get_param_names / get_dims, src/stanExports_gp_grid.h:1588-1704), zero-length la_1 / la_2 / alpha included, "lp__"
 * last.  block i of `problem`: name, ndims (0 scalar, 1 vector, 2 matrix), dims[2], and the offset of its first element
 * in a write_array row (matrices column-major).  update_param_oi / param_oi_tidx / param_fnames_oi are bookkeeping over
 * these blocks and stay on the R side of the binding. */
int32_t bsyn_num_param_blocks(const bsyn_batch *b);
int bsyn_param_block(const bsyn_batch *b, int32_t problem, int32_t block, char *name, int32_t buflen, int32_t *ndims,
                     int32_t *dims, int32_t *flat_offset);
int32_t bsyn_max_params(const bsyn_batch *b);
int32_t bsyn_max_outputs(const bsyn_batch *b);
int32_t bsyn_num_problems(const bsyn_batch *b);

/* ---- log density + gradient -----------------------------------------------------------------------
 * Replaces: log_prob(upar, jacobian_adjust, gradient) / grad_log_prob (cc:24-25) =
 * model::log_prob<propto=false*, jacobian>(...) + reverse sweep (src/stanExports_gp_grid.h:1076-1577).
 * (* the Stan programs use `target +=`, so no constant is dropped either way.)
 * B evaluation points; point k uses problem prob_idx[k] and u + k*ldu (first D entries).
 * lp[B], grad[B*ldu] (may be NULL), status[B]: 0 ok, 1 = the reference would throw std::domain_error
 * (non-PD kernel matrix, transformed parameter out of range, NaN): lp = -inf, grad = 0.
 * Where the reference's reverse sweep yields NaN although lp is finite (the Bliss product rounds to exactly 0 or 1,
 * so log(p0 / (1 - p0)) in gp_grid.stan:160-161 is infinite, or 10^(slope (x - m)) overflows) the gradient is NaN
 * with status 0 and the finite lp, as from grad_log_prob.  grad == NULL takes a value-only path (no reverse sweep).
 * Host pointers (copies are made inside the call). */
int bsyn_logp_grad(bsyn_batch *b, int32_t B, const int32_t *prob_idx, const double *u, int32_t ldu,
                   int32_t jacobian, double *lp, double *grad, int32_t *status);

/* Same, with every pointer a DEVICE pointer on the batch's device (no copies; asynchronous on `stream`,
 * a cudaStream_t passed as void*; NULL = default stream). */
int bsyn_logp_grad_device(bsyn_batch *b, int32_t B, const int32_t *d_prob_idx, const double *d_u, int32_t ldu,
                          int32_t jacobian, double *d_lp, double *d_grad, int32_t *d_status, void *stream);

/* ---- write_array ----------------------------------------------------------------------------------
 * Replaces: constrain_pars (cc:23) and the per-saved-draw model.write_array (h:1705-2235): constrained
 * parameters, transformed parameters (p0, p01, p02, Delta, GP) and generated quantities
 * (ec50_k, CPO[N], dss_k, rVUS_f, rVUS_p0, VUS_Delta, VUS_syn, VUS_ant), in the order of h:1588-1622.
 * out[B*ldo].  seed drives the U(1e-6,1e-4) replacement of exact-zero VUS values (gp_grid.stan:294-297). */
int bsyn_write_array(bsyn_batch *b, int32_t B, const int32_t *prob_idx, const double *u, int32_t ldu,
                     double *out, int32_t ldo, uint64_t seed, int32_t *status);

/* Replaces: unconstrain_pars (cc:22) / transform_inits (h:803-1075).  theta = the first D entries of a
 * write_array row (constrained parameters in declaration order). Pure host arithmetic. */
int bsyn_unconstrain(const bsyn_batch *b, int32_t problem, const double *theta, double *u);

/* ---- sampling -------------------------------------------------------------------------------------
 * Replaces: call_sampler(args) (cc:15) for method="sampling", algorithm="NUTS": Stan's adaptive NUTS
 * (hmc_nuts_diag_e_adapt / hmc_nuts_dense_e_adapt) run for every problem x chain of the batch in ONE
 * device-resident launch.  Field names and defaults follow rstan's `control` list. */
typedef struct bsyn_sampler_args {
    int32_t chains;          /* 4 */
    int32_t iter;            /* 2000 (warmup + sampling, as in rstan) */
    int32_t warmup;          /* 1000 */
    int32_t thin;            /* 1 */
    uint64_t seed;
    double init_r;           /* 2.0: inits ~ U(-init_r, init_r) on the unconstrained scale */
    double adapt_delta;      /* 0.9 (bayesynergy override of rstan's 0.8, R/bayesynergy.R:107-110) */
    double stepsize;         /* 1.0 */
    double stepsize_jitter;  /* 0 (0.5 when robust, R/bayesynergy.R:112-115) */
    int32_t max_treedepth;   /* 10 */
    int32_t metric;          /* 0 = diag_e, 1 = dense_e */
    double adapt_gamma, adapt_kappa, adapt_t0;            /* 0.05, 0.75, 10 */
    int32_t adapt_init_buffer, adapt_term_buffer, adapt_window;   /* 75, 50, 25 */
    int32_t save_warmup;     /* 0 */
} bsyn_sampler_args;

void bsyn_sampler_defaults(bsyn_sampler_args *a);

/* Number of generated scalars kept per saved draw for the screen path (see BSYN_Q_* below). */
#define BSYN_NQ 12
enum {
    BSYN_Q_EC50_1 = 0, BSYN_Q_EC50_2, BSYN_Q_DSS_1, BSYN_Q_DSS_2, BSYN_Q_S,
    BSYN_Q_RVUS_F, BSYN_Q_RVUS_P0, BSYN_Q_VUS_DELTA, BSYN_Q_VUS_SYN, BSYN_Q_VUS_ANT,
    BSYN_Q_LP, BSYN_Q_LOGLIK_UNUSED
};
#define BSYN_NSP 7   /* accept_stat__, stepsize__, treedepth__, n_leapfrog__, divergent__, energy__, lp__ */

/* Result sinks (host pointers; any may be NULL).  n_save = (iter - warmup)/thin (+ warmup/thin if
 * save_warmup).  Strides are in doubles. */
typedef struct bsyn_sample_sinks {
    /* full write_array rows: [n_problems][chains][n_save][ld_draws]; what rstan::extract() reads
     * (R/bayesynergy.R:244-291).  Large: intended for bayesynergy() on one / a few experiments. */
    double *draws;
    int64_t ld_draws;
    /* unconstrained draws [n_problems][chains][n_save][ld_udraws] (bridge sampling) */
    double *udraws;
    int64_t ld_udraws;
    /* the BSYN_NQ generated scalars per draw: [n_problems][chains][n_save][BSYN_NQ] */
    double *qdraws;
    /* sampler params per draw: [n_problems][chains][n_save][BSYN_NSP] (attr "sampler_params") */
    double *sampler_params;
    /* per-problem posterior summaries over all chains, what synergyscreen() puts in one screenSummary row
     * (R/synergyscreen.R:185-218): [n_problems][BSYN_NQ][2] = mean, sd (sd with n-1 denominator, pooled) */
    double *summary;
    /* per-problem mean over draws of CPO_n: [n_problems][ld_cpo] -> LPML = -sum log(mean CPO) (R/bayesynergy.R:244-246) */
    double *cpo_mean;
    int64_t ld_cpo;
    /* per problem x chain: [n_problems][chains][8] = stepsize, n_divergent, n_max_treedepth, n_leapfrog_total
     * (warmup+sampling), n_leapfrog_sampling, mean accept_stat (sampling), init_failed, reserved */
    double *chain_stats;
    /* adapted inverse metric [n_problems][chains][bsyn_max_params()]: the diagonal M^-1 of diag_e, the diagonal of the
     * dense M^-1 of dense_e (the device keeps the dense matrix rounded to float for the M^-1 p products; its exact Cholesky
     * factor is used for the momentum draws, so the sampler stays a valid HMC: a narrower arithmetic than the reference's) */
    double *inv_metric;
    /* the initial point of every chain on the unconstrained scale [n_problems][chains][bsyn_max_params()] (what rstan
     * reports, constrained, as attr "inits"; constrain it with bsyn_write_array) */
    double *init_u;
    /* posterior mean of the response surface f per cell of the column-major (n2+1) x (n1+1) dose grid, accumulated on
     * the device per chain (no draws kept): [n_problems][ld_f].  What the outlier rule of bayesynergy() needs
     * (R/bayesynergy.R:305-314) and, for non-robust fits, posterior_mean$f (R/bayesynergy.R:251-262). */
    double *f_mean;
    int64_t ld_f;            /* >= (n1+1)(n2+1) of every problem; stride of f_mean and surfaces */
    /* posterior surfaces reduced on the device from the write_array draws (the draws stay in HBM unless `draws` is
     * also requested): [n_problems][3][ld_f] = f (mean; median when robust), p0 (median), Delta (median), exactly
     * R/bayesynergy.R:251-283 (R's median); and LPML = -sum_n log(mean CPO_n) (R/bayesynergy.R:246): [n_problems].
     * Up to 8192 pooled sampling draws per problem. */
    double *surfaces;
    double *lpml;
    /* bridge-sampling log marginal likelihood of every problem (Meng & Wong 1996, normal proposal: what
     * bridgesampling::bridge_sampler(fit) returns as $logml, R/bayesynergy.R:332-335), computed on the device from the
     * unconstrained draws of this call, which stay in HBM: proposal fit on the first half of every chain, ONE batched
     * log_prob launch at the second-half draws and the proposal draws of all problems, iterative scheme.  [n_problems];
     * bridge_niter (may be NULL): iterations of the scheme.  Needs save_warmup = 0, <= 160 parameters, <= 8192 draws in
     * the second halves.  The Bayes factor of R/bayesynergy.R:336 is exp(logml(gp_grid batch) - logml(nointeraction batch)). */
    double *bridge_logml;
    int32_t *bridge_niter;
    /* device time of the sampling launch in ms (attr "elapsed_time" is this, split by the leapfrog counts of chain_stats) */
    float *elapsed_ms;
} bsyn_sample_sinks;

/* Runs the whole screen.  total_grad_evals (may be NULL) receives the number of log-density gradient
 * evaluations performed (all problems, chains, warmup + sampling). */
int bsyn_sample(bsyn_batch *b, const bsyn_sampler_args *args, const bsyn_sample_sinks *sinks,
                int64_t *total_grad_evals);

/* Device-resident variant for throughput measurement and for callers that post-process on the GPU:
 * runs the sampler and leaves qdraws / sampler_params / summary in device buffers owned by the batch
 * (valid until the next sampling call or destroy).  Pointers are returned through *d_qdraws etc. (may be NULL).
 * elapsed_ms: device time of the sampling kernel(s), measured with CUDA events on the launch stream. */
int bsyn_sample_device(bsyn_batch *b, const bsyn_sampler_args *args, int64_t *total_grad_evals,
                       float *elapsed_ms, const double **d_qdraws, const double **d_sampler_params,
                       const double **d_summary);

/* ESS-target control (north_star's target: every combination sampled to an ESS of 400).
 * Phase 1: the sampler runs with `args` (warm-up + a first block of iter - warmup sampling draws); on the device the
 * rank-normalised bulk ESS of the generated scalars q1 and q2 (e.g. BSYN_Q_VUS_SYN, BSYN_Q_VUS_ANT) is computed for every
 * problem, and the sampler is relaunched over the problems whose min(ESS) is still below `target_ess`: their chains RESUME
 * (last position, adapted step size and diagonal metric) and append another block, up to `max_rounds` launches -- the
 * batched form of the reference's rerun of `failed` experiments (R/synergyscreen.R:97-103), nothing refitted from scratch.
 * Phase 2, the reference's remedy for fits with warnings (adapt_delta = 0.99, R/synergyscreen.R:149-157): what is still
 * below the target resumes once more, re-adapts its STEP SIZE to `retry_adapt_delta` for `retry_warmup` iterations
 * (init_stepsize + dual averaging; the adapted metric is kept) and is extended up to `retry_rounds` more times
 * (retry_rounds = 0: no phase 2).  diag_e only; chains * (iter - warmup) * (max_rounds + retry_rounds) <= 8192. */
typedef struct bsyn_ess_target {
    double target_ess;           /* 400 */
    int32_t q1, q2;              /* BSYN_Q_VUS_SYN, BSYN_Q_VUS_ANT */
    int32_t max_rounds;          /* 6: launches of phase 1 */
    double retry_adapt_delta;    /* 0.99 */
    int32_t retry_warmup;        /* 150 */
    int32_t retry_rounds;        /* 6 */
} bsyn_ess_target;
typedef struct bsyn_ess_run {
    int64_t total_grad_evals;
    float elapsed_ms;            /* device time of all sampler and ESS launches */
    float phase1_ms;
    int32_t rounds;              /* sampler launches made */
    int32_t n_retried;           /* problems that entered phase 2 */
} bsyn_ess_run;
void bsyn_ess_target_defaults(bsyn_ess_target *t);
/* Sinks: qdraws / sampler_params use a row stride of (iter - warmup) * (max_rounds + retry_rounds) draws per chain (unused
 * slots 0), summary is over each problem's own draws; every other sink must be NULL.  ess[n_problems] (min over q1, q2 at
 * the problem's last check), draws_per_chain[n_problems], run: may be NULL. */
int bsyn_sample_to_ess(bsyn_batch *b, const bsyn_sampler_args *args, const bsyn_ess_target *target,
                       const bsyn_sample_sinks *sinks, double *ess, int32_t *draws_per_chain, bsyn_ess_run *run);

/* Rank-normalisation-free split-chain bulk ESS (Geyer initial positive sequence on the pooled split chains,
 * as rstan::summary's n_eff) of generated scalar `q` for every problem, computed on the device from the
 * qdraws of the last sampling call.  ess[n_problems], rhat[n_problems] (may be NULL).  Host pointers. */
int bsyn_ess(bsyn_batch *b, int32_t q, double *ess, double *rhat);

/* Rank-normalised split-chain bulk ESS and R-hat, and tail ESS (Vehtari et al. 2021: rstan's Bulk_ESS / Tail_ESS; the
 * estimator SURVEY.md 8(d) defines "ESS per combination" on) of generated scalar `q` for every problem: per problem the
 * pooled draws of the last sampling call are sorted on the device, replaced by normal scores of their average ranks,
 * and scored with the estimator of bsyn_ess; tail = min over the 5 % / 95 % quantile indicators.  Up to 8192 pooled
 * draws per problem.  ess_bulk[n_problems]; rhat, ess_tail may be NULL.  Host pointers. */
int bsyn_bulk_ess(bsyn_batch *b, int32_t q, double *ess_bulk, double *rhat, double *ess_tail);

/* ---- data preparation in native code: R/bayesynergy.R:129-154 (the step before the path; no GPU involved).
 * y: [n x nrep_in] column-major (replicate r of row i at y[i + r n]), NaN = missing.  x: [n x 2] row-major
 * concentrations (0 = monotherapy).  Outputs (caller-allocated): x1_out[>= n], x2_out[>= n] log10 dose levels,
 * y_out / ii_out [>= n nrep_in] the non-missing observations and their 1-based positions in the column-major
 * (n2+1) x (n1+1) grid, and the sizes.  nrep_in == 1: replicates are repeated rows and nrep = the largest
 * multiplicity (R/bayesynergy.R:149-154).  Returns BSYN_ERR_DOMAIN when a drug has no zero concentration. */
int bsyn_prepare(const double *y, const double *x, int32_t n, int32_t nrep_in, double *x1_out, int32_t *n1_out,
                 double *x2_out, int32_t *n2_out, double *y_out, int32_t *ii_out, int32_t *N_out,
                 int32_t *nrep_out, int32_t *nmissing_out);

/* ---- full-rank ADVI: replaces rstan::vb(stanmodels$gp_grid, stan_data, algorithm = "fullrank", ...)
 * (R/bayesynergy.R:225-231; the default method of synergyscreen, R/synergyscreen.R:90-93), i.e. the module's
 * call_sampler with method = "variational" (src/stanExports_gp_grid.cc:15): Stan's advi::run for every problem
 * of the batch in ONE launch.  Field names and defaults follow rstan::vb. */
typedef struct bsyn_advi_args {
    int32_t iter;            /* 10000: maximum number of stochastic-gradient iterations */
    int32_t grad_samples;    /* 1 (the only supported value) */
    int32_t elbo_samples;    /* 100 */
    int32_t eval_elbo;       /* 100 */
    double eta;              /* 1.0; ignored when adapt_engaged */
    int32_t adapt_engaged;   /* 1 */
    int32_t adapt_iter;      /* 50 */
    double tol_rel_obj;      /* 0.01 */
    int32_t output_samples;  /* 1000 draws from the approximate posterior */
    double init_r;           /* 2.0 */
    uint64_t seed;
} bsyn_advi_args;

void bsyn_advi_defaults(bsyn_advi_args *a);

#define BSYN_NVI 8   /* per-problem info: status, eta, iterations, last ELBO, convergence bits, n_grad, n_lp, reserved */
/* info[0] status: 0 ok, 1 initialisation failed, 2 initial ELBO failed, 3 all step sizes failed, 4 gradient
 * failed, 5 ELBO failed (the reference throws an error in each of these cases -> returnCode 2).
 * info[4] convergence bits: 1 mean ELBO converged, 2 median ELBO converged, 4 "may be diverging" was printed,
 * 8 the maximum number of iterations was reached (the reference warns). */
typedef struct bsyn_advi_sinks {
    double *mean;            /* [n_problems][ld_mean] variational mean on the unconstrained scale (Stan order) */
    int64_t ld_mean;
    double *info;            /* [n_problems][BSYN_NVI] */
    double *draws;           /* [n_problems][output_samples][ld_draws] write_array rows of the draws from the approximation */
    int64_t ld_draws;
    double *udraws;          /* [n_problems][output_samples][ld_udraws] */
    int64_t ld_udraws;
    double *qdraws;          /* [n_problems][output_samples][BSYN_NQ] */
    double *summary;         /* [n_problems][BSYN_NQ][2] mean, sd over the output draws */
    double *cpo_mean;        /* [n_problems][ld_cpo] */
    int64_t ld_cpo;
} bsyn_advi_sinks;

/* total_evals (may be NULL): lp(+gradient) evaluations performed; elapsed_ms (may be NULL): device time. */
int bsyn_advi(bsyn_batch *b, const bsyn_advi_args *args, const bsyn_advi_sinks *sinks, int64_t *total_evals,
              float *elapsed_ms);

/* Number of CUDA kernel launches issued by this library so far in this process (for bench accounting). */
int64_t bsyn_launch_count(void);

/* Register-resident DFMA microbenchmark: achieved FP64 FMA lane-operations per second on `device`
 * (the measured denominator of the FP64 roofline; x2 = FLOP/s). */
int bsyn_measure_fp64_peak(int32_t device, double *fma_per_sec);

/* Test hooks (not part of the reference-facing surface): bsyn_logp_grad plus the kernel's intermediates. */
int bsyn_logp_grad_debug(bsyn_batch *b, int32_t B, const int32_t *prob_idx, const double *u, int32_t ldu,
                         int32_t jacobian, double *lp, double *grad, int32_t *status, double *dbg,
                         int32_t dbg_stride);
int bsyn_debug_layout(const bsyn_batch *b, int32_t *nm, int32_t *nn, int32_t *gm);

#ifdef __cplusplus
}
#endif
#endif /* BSYN_H */
