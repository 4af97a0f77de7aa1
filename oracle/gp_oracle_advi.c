/* ORACLE (test infrastructure, not product code) -- PARITY UNPINNED.  See gp_oracle.h.
 *
 * Full-rank ADVI as `rstan::vb(..., algorithm = "fullrank")` runs it for bayesynergy
 * (R/bayesynergy.R:225-231; synergyscreen's default method, R/synergyscreen.R:90-93).
 * The algorithm lives in the un-vendored Stan library (StanHeaders, stamped 2.21.0 in the
 * generated model code): stan/variational/advi.hpp (run, adapt_eta, stochastic_gradient_ascent,
 * calc_ELBO, calc_ELBO_grad, circ_buff_median, rel_difference) and
 * stan/variational/families/normal_fullrank.hpp (calc_grad, entropy, sample).  This file
 * restates that published algorithm on top of orc_logp_grad / orc_write_array.
 */
#include "gp_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

static uint64_t sm64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static double runif(uint64_t *s) { return ((sm64(s) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static double rnorm(uint64_t *s)
{
    double u1 = runif(s), u2 = runif(s);
    return sqrt(-2 * log(u1)) * cos(2 * M_PI * u2);
}

void orc_advi_defaults(orc_advi_args *a)
{
    a->iter = 10000; a->grad_samples = 1; a->elbo_samples = 100; a->eval_elbo = 100;
    a->eta = 1.0; a->adapt_engaged = 1; a->adapt_iter = 50; a->tol_rel_obj = 0.01;
    a->output_samples = 1000; a->init_radius = 2.0; a->seed = 1; a->chain_id = 0;
}

typedef struct {
    const orc_problem *p;
    int D;
    double *mu, *L;        /* L row-major D x D, lower triangular */
    double *eta, *zeta, *g;
    uint64_t *rng;
    long n_grad, n_lp;
} vstate;

/* normal_fullrank::sample: zeta = L * eta + mu, eta ~ N(0, I) */
static void q_sample(vstate *V)
{
    const int D = V->D;
    for (int d = 0; d < D; ++d) V->eta[d] = rnorm(V->rng);
    for (int i = 0; i < D; ++i) {
        double s = V->mu[i];
        for (int j = 0; j <= i; ++j) s += V->L[(size_t)i * D + j] * V->eta[j];
        V->zeta[i] = s;
    }
}

/* advi::calc_ELBO.  Returns 0 ok, 1 = "number of dropped evaluations has reached its maximum amount". */
static int calc_elbo(vstate *V, int n_mc, double *out)
{
    const int D = V->D;
    double elbo = 0;
    int dropped = 0;
    for (int i = 0; i < n_mc;) {
        q_sample(V);
        double lp;
        int st = orc_logp_grad(V->p, V->zeta, 1, &lp, NULL);
        ++V->n_lp;
        if (st == 0 && isfinite(lp)) { elbo += lp; ++i; }
        else if (++dropped >= n_mc) return 1;
    }
    elbo /= n_mc;
    double ent = 0.5 * D * (1.0 + log(2 * M_PI));
    for (int d = 0; d < D; ++d) {
        double t = V->L[(size_t)d * D + d];
        if (t != 0.0) ent += log(fabs(t));
    }
    *out = elbo + ent;
    return 0;
}

/* normal_fullrank::calc_grad with the entropy term.  gm[D], gL[D*D].  Returns 0 ok, 1 failed. */
static int calc_grad(vstate *V, int n_mc, double *gm, double *gL)
{
    const int D = V->D;
    memset(gm, 0, sizeof(double) * D);
    memset(gL, 0, sizeof(double) * (size_t)D * D);
    for (int m = 0; m < n_mc; ++m) {
        q_sample(V);
        double lp;
        int st = orc_logp_grad(V->p, V->zeta, 1, &lp, V->g);
        ++V->n_grad;
        if (st != 0) return 1;
        for (int i = 0; i < D; ++i) if (!isfinite(V->g[i])) return 1;
        for (int i = 0; i < D; ++i) {
            gm[i] += V->g[i];
            for (int j = 0; j <= i; ++j) gL[(size_t)i * D + j] += V->g[i] * V->eta[j];
        }
    }
    for (int i = 0; i < D; ++i) {
        gm[i] /= n_mc;
        for (int j = 0; j <= i; ++j) gL[(size_t)i * D + j] /= n_mc;
    }
    for (int d = 0; d < D; ++d) gL[(size_t)d * D + d] += 1.0 / V->L[(size_t)d * D + d];
    return 0;
}

static void q_reset(vstate *V, const double *init)
{
    const int D = V->D;
    memcpy(V->mu, init, sizeof(double) * D);
    memset(V->L, 0, sizeof(double) * (size_t)D * D);
    for (int d = 0; d < D; ++d) V->L[(size_t)d * D + d] = 1.0;
}

/* one adaGrad-style step shared by adapt_eta and stochastic_gradient_ascent */
static void sga_step(vstate *V, int iter, double eta, const double *gm, const double *gL, double *hm, double *hL)
{
    const int D = V->D;
    const double pre = 0.9, post = 0.1, tau = 1.0;
    const double es = eta / sqrt((double)iter);
    for (int i = 0; i < D; ++i) {
        hm[i] = (iter == 1) ? hm[i] + gm[i] * gm[i] : pre * gm[i] * gm[i] + post * hm[i];
        V->mu[i] += es * gm[i] / (tau + sqrt(hm[i]));
        for (int j = 0; j <= i; ++j) {
            const size_t e = (size_t)i * D + j;
            hL[e] = (iter == 1) ? hL[e] + gL[e] * gL[e] : pre * gL[e] * gL[e] + post * hL[e];
            V->L[e] += es * gL[e] / (tau + sqrt(hL[e]));
        }
    }
}

static int cmp_d(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return (x > y) - (x < y); }

int orc_advi_fullrank(const orc_problem *p, const orc_advi_args *a, double *mean_u, double *L_out, double *outputs,
                      double *draws_u, double *info)
{
    const int D = orc_dim(p), nout = orc_num_outputs(p);
    uint64_t rng = a->seed * 0x9E3779B97F4A7C15ULL + (uint64_t)a->chain_id * 0xD1B54A32D192ED03ULL + 777;
    vstate V;
    V.p = p; V.D = D; V.rng = &rng; V.n_grad = 0; V.n_lp = 0;
    V.mu = (double *)malloc(sizeof(double) * D);
    V.L = (double *)malloc(sizeof(double) * (size_t)D * D);
    V.eta = (double *)malloc(sizeof(double) * D);
    V.zeta = (double *)malloc(sizeof(double) * D);
    V.g = (double *)malloc(sizeof(double) * D);
    double *init = (double *)malloc(sizeof(double) * D);
    double *gm = (double *)malloc(sizeof(double) * D), *hm = (double *)calloc(D, sizeof(double));
    double *gL = (double *)malloc(sizeof(double) * (size_t)D * D), *hL = (double *)calloc((size_t)D * D, sizeof(double));
    int status = 0, conv = 0, iters = 0;
    double eta = a->eta, elbo = 0.0;

    /* services::util::initialize: u ~ U(-R, R) until lp and gradient are finite (<= 100 tries) */
    int ok = 0;
    for (int tr = 0; tr < 100 && !ok; ++tr) {
        for (int i = 0; i < D; ++i) init[i] = a->init_radius * (2 * runif(&rng) - 1);
        double lp;
        if (orc_logp_grad(p, init, 1, &lp, V.g) == 0 && isfinite(lp)) {
            ok = 1;
            for (int i = 0; i < D; ++i) if (!isfinite(V.g[i])) ok = 0;
        }
    }
    if (!ok) { status = 1; goto done; }
    q_reset(&V, init);

    if (a->adapt_engaged) {
        /* advi::adapt_eta */
        static const double seq[5] = {100, 10, 1, 0.1, 0.01};
        double elbo_best = -DBL_MAX, elbo_init, eta_best = 0.0;
        if (calc_elbo(&V, a->elbo_samples, &elbo_init)) { status = 2; goto done; }
        int more = 1, k = 0;
        while (more) {
            const double e = seq[k];
            for (int it = 1; it <= a->adapt_iter; ++it) {
                if (calc_grad(&V, a->grad_samples, gm, gL)) {   /* "(ROBUST) ... It's OK if it diverges" */
                    memset(gm, 0, sizeof(double) * D);
                    memset(gL, 0, sizeof(double) * (size_t)D * D);
                }
                sga_step(&V, it, e, gm, gL, hm, hL);
            }
            double el;
            if (calc_elbo(&V, a->elbo_samples, &el)) el = -DBL_MAX;
            if (el < elbo_best && elbo_best > elbo_init) {
                more = 0;
            } else {
                if (k < 4) { elbo_best = el; eta_best = e; }
                else if (el > elbo_init) { eta_best = e; more = 0; }
                else { status = 3; goto done; }       /* "All proposed step-sizes failed" */
                memset(hm, 0, sizeof(double) * D);
                memset(hL, 0, sizeof(double) * (size_t)D * D);
            }
            ++k;
            q_reset(&V, init);
        }
        eta = eta_best;
        memset(hm, 0, sizeof(double) * D);
        memset(hL, 0, sizeof(double) * (size_t)D * D);
    }

    {   /* advi::stochastic_gradient_ascent */
        int cb_size = (int)fmax(0.1 * a->iter / a->eval_elbo, 2.0);
        double *cb = (double *)malloc(sizeof(double) * cb_size), *tmp = (double *)malloc(sizeof(double) * cb_size);
        int cb_n = 0, cb_head = 0;
        double elbo_prev;
        int more = 1;
        for (int it = 1; more; ++it) {
            if (calc_grad(&V, a->grad_samples, gm, gL)) { status = 4; free(cb); free(tmp); goto done; }
            sga_step(&V, it, eta, gm, gL, hm, hL);
            iters = it;
            if (it % a->eval_elbo == 0) {
                elbo_prev = elbo;
                if (calc_elbo(&V, a->elbo_samples, &elbo)) { status = 5; free(cb); free(tmp); goto done; }
                const double delta = fabs((elbo_prev - elbo) / elbo);   /* rel_difference(elbo, elbo_prev) */
                cb[cb_head] = delta; cb_head = (cb_head + 1) % cb_size; if (cb_n < cb_size) ++cb_n;
                double ave = 0;
                for (int i = 0; i < cb_n; ++i) { ave += cb[i]; tmp[i] = cb[i]; }
                ave /= cb_n;
                qsort(tmp, cb_n, sizeof(double), cmp_d);
                const double med = tmp[cb_n / 2];
                if (ave < a->tol_rel_obj) { conv |= 1; more = 0; }
                if (med < a->tol_rel_obj) { conv |= 2; more = 0; }
                if (it > 10 * a->eval_elbo && (med > 0.5 || ave > 0.5)) conv |= 4;   /* "MAY BE DIVERGING... INSPECT ELBO": a message only */
            }
            if (it == a->iter) { if (more) conv |= 8; more = 0; }
        }
        free(cb); free(tmp);
    }

    if (mean_u) memcpy(mean_u, V.mu, sizeof(double) * D);
    if (L_out) memcpy(L_out, V.L, sizeof(double) * (size_t)D * D);
    for (int n = 0; n < a->output_samples; ++n) {
        q_sample(&V);
        if (draws_u) memcpy(draws_u + (size_t)n * D, V.zeta, sizeof(double) * D);
        if (outputs) orc_write_array(p, V.zeta, outputs + (size_t)n * nout, &rng);
    }
done:
    if (info) {
        info[0] = status; info[1] = eta; info[2] = iters; info[3] = elbo; info[4] = conv;
        info[5] = (double)V.n_grad; info[6] = (double)V.n_lp; info[7] = 0;
    }
    free(V.mu); free(V.L); free(V.eta); free(V.zeta); free(V.g); free(init); free(gm); free(hm); free(gL); free(hL);
    return status;
}
