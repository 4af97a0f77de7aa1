/* ORACLE (test infrastructure, not product code) -- PARITY UNPINNED.
 *
 * Plain-C CPU restatement of the reference's gp_grid / nointeraction hot path:
 * log-density + hand-derived reverse-mode gradient, write_array (constrained
 * parameters, transformed parameters, generated quantities) and the NUTS control
 * loop that calls them.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may link or call this.
 *
 * "Parity unpinned": the reference holds no tests or golden vectors for this path and
 * its arithmetic lives in un-vendored dependencies (rstan >= 2.18.1, StanHeaders >=
 * 2.18.0 -- DESCRIPTION:30-44; generated code stamped Stan 2.21.0) that cannot be
 * built here.  This file restates the Stan programs and the published Stan Math /
 * Stan NUTS algorithms and is cross-checked against oracle/torch_oracle.py (autograd)
 * and oracle/mp_oracle.py (50-digit mpmath).
 */
#ifndef GP_ORACLE_H
#define GP_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The Stan data block (inst/stan/gp_grid.stan:10-29, nointeraction.stan:7-23). */
typedef struct {
    int model;            /* 0 = gp_grid, 1 = nointeraction */
    int n1, n2, N;
    int est_la, robust, pcprior, heteroscedastic, kernel, nu_matern, est_alpha;
    double lambda, rho;
    double pcprior_hypers[4];
    const double *y;      /* [N] */
    const int *ii_obs;    /* [N] 1-based, column-major (n2+1)x(n1+1) */
    const double *x1;     /* [n1] */
    const double *x2;     /* [n2] */
} orc_problem;

int orc_dim(const orc_problem *p);           /* length of the unconstrained vector */
int orc_num_outputs(const orc_problem *p);   /* length of one write_array row */

/* log_prob<propto=false, jacobian>(u) and its gradient (grad may be NULL).
 * Returns 0 = ok, 1 = the reference would have thrown std::domain_error (reject). */
int orc_logp_grad(const orc_problem *p, const double *u, int jacobian, double *lp, double *grad);

/* write_array: out[orc_num_outputs].  Exact zeros of rVUS_f/rVUS_p0/VUS_syn/VUS_ant are replaced by
 * U(1e-6,1e-4) draws from *rng (gp_grid.stan:294-297). Returns 0 ok / 1 domain error. */
int orc_write_array(const orc_problem *p, const double *u, double *out, uint64_t *rng);

typedef struct {
    int num_warmup, num_samples, max_treedepth;
    double adapt_delta, stepsize, stepsize_jitter;
    double gamma, kappa, t0;
    int init_buffer, term_buffer, window;
    double init_radius;
    uint64_t seed;
    int chain_id;
    int save_warmup;
    int metric;           /* 0 = diag_e, 1 = dense_e (what robust = TRUE selects, R/bayesynergy.R:112-115) */
} orc_nuts_args;

void orc_nuts_defaults(orc_nuts_args *a);

/* One chain of adaptive NUTS (Stan's hmc_nuts_diag_e_adapt, or hmc_nuts_dense_e_adapt when a->metric == 1).
 * draws_u: [n_save * D] unconstrained draws (may be NULL); outputs: [n_save * n_out] write_array rows
 * (may be NULL); sampler_params: [n_save * 7] = accept_stat, stepsize, treedepth, n_leapfrog, divergent,
 * energy, lp (may be NULL).  n_save = num_samples (+ num_warmup if save_warmup).
 * Returns total number of gradient evaluations (leapfrogs incl. warmup), or -1 on init failure. */
long orc_nuts_chain(const orc_problem *p, const orc_nuts_args *a, double *draws_u, double *outputs,
                    double *sampler_params, double *final_stepsize, double *final_inv_metric);

/* Full-rank ADVI (rstan::vb(algorithm = "fullrank") defaults; gp_oracle_advi.c). */
typedef struct {
    int iter, grad_samples, elbo_samples, eval_elbo;
    double eta;
    int adapt_engaged, adapt_iter;
    double tol_rel_obj;
    int output_samples;
    double init_radius;
    uint64_t seed;
    int chain_id;
} orc_advi_args;

void orc_advi_defaults(orc_advi_args *a);

/* mean_u[D], L_out[D*D] row-major (may be NULL), outputs[output_samples * n_out] write_array rows of the draws
 * from the approximation (may be NULL), draws_u[output_samples * D] (may be NULL), info[8] = status, eta,
 * iterations, last ELBO, convergence bits (1 mean, 2 median, 4 may-be-diverging message seen, 8 max iterations), number of
 * gradient evaluations, number of lp evaluations, 0.
 * status: 0 ok, 1 initialisation failed, 2 initial ELBO failed, 3 all step sizes failed, 4 gradient failed,
 * 5 ELBO failed (the reference throws in each of these cases). */
int orc_advi_fullrank(const orc_problem *p, const orc_advi_args *a, double *mean_u, double *L_out, double *outputs,
                      double *draws_u, double *info);

#ifdef __cplusplus
}
#endif
#endif
