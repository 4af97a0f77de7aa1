/* ORACLE (test infrastructure, not product code) -- PARITY UNPINNED.  See gp_oracle.h.
 *
 * Follows, statement by statement:
 *   inst/stan/gp_grid.stan:64-299, inst/stan/nointeraction.stan:33-186,
 *   inst/stan/include/{kron_mvprod,VUS,dss,pc_prior,lptn,matrix_min,matrix_max}.stan,
 *   src/stanExports_gp_grid.h:1086-1215 (parameter order / transforms), :1396-1457 (validation),
 *   :1497-1568 (lpdf calls), :1705-2235 (write_array),
 * and for the sampler the published Stan algorithms (stan/mcmc/hmc/nuts/base_nuts.hpp,
 * hamiltonians/diag_e_metric.hpp, integrators/expl_leapfrog.hpp, stepsize_adaptation.hpp,
 * windowed_adaptation.hpp, var_adaptation.hpp, services/util/initialize.hpp) -- SURVEY.md App. B/D.
 */
#include "gp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LOG_2PI_HALF 0.91893853320467274178
#define LN10 2.30258509299404568402
#define LOG_PI 1.14472988584940017414
#define LOG_2 0.69314718055994530942

/* ------------------------------------------------------------------------------------------ */

int orc_dim(const orc_problem *p)
{
    if (p->model == 1) return 9 + 2 * p->est_la;
    return 13 + 2 * p->est_la + p->est_alpha + p->n1 * p->n2;
}

int orc_num_outputs(const orc_problem *p)
{
    int G = p->n1 * p->n2;
    if (p->model == 1) return orc_dim(p) + G + p->n1 + p->n2 + 2 + p->N + 4;
    return orc_dim(p) + 3 * G + p->n1 + p->n2 + 2 + p->N + 7;
}

/* inverse normal cdf (Acklam's rational approximation + one Halley step against erfc): tauLPTN */
static double inv_Phi(double pr)
{
    static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                               1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                               6.680131188771972e+01, -1.328068155288572e+01};
    static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                               -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
    static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                               3.754408661907416e+00};
    double q, r, x;
    if (pr < 0.02425) {
        q = sqrt(-2 * log(pr));
        x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    } else if (pr <= 1 - 0.02425) {
        q = pr - 0.5;
        r = q * q;
        x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
            (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
    } else {
        q = sqrt(-2 * log(1 - pr));
        x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    }
    for (int it = 0; it < 3; ++it) { /* Newton/Halley refinement to full double precision */
        double e = 0.5 * erfc(-x / sqrt(2.0)) - pr;
        double u = e * sqrt(2 * M_PI) * exp(x * x / 2);
        x = x - u / (1 + x * u / 2);
    }
    return x;
}

static void lptn_consts(double rho, double *tau, double *lam)
{
    /* gp_grid.stan:37-38 */
    double t = inv_Phi((1 + rho) / 2);
    *tau = t;
    *lam = 2.0 / (1.0 - rho) * exp(-0.5 * t * t - LOG_2PI_HALF) * t * log(t);
}

/* dense lower Cholesky, row-major n x n; returns 1 if not positive definite */
static int chol_lower(int n, const double *A, double *L)
{
    memset(L, 0, sizeof(double) * n * n);
    for (int j = 0; j < n; ++j) {
        double d = A[j * n + j];
        for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k];
        if (!(d > 0.0)) return 1;
        d = sqrt(d);
        L[j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = A[i * n + j];
            for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = s / d;
        }
    }
    return 0;
}

/* Sbar_sym = 1/2 * L^-T (Phi + Phi^T) L^-1, Phi = tril(L^T Lbar) with halved diagonal (Murray 2016). */
static void chol_adjoint(int n, const double *L, const double *Lbar, double *Sbar, double *w)
{
    double *P = w, *S = w + n * n, *Y = w + 2 * n * n;
    /* P = tril(L^T Lbar), diag halved */
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) {
            double s = 0;
            if (b <= a) {
                for (int i = a; i < n; ++i) s += L[i * n + a] * Lbar[i * n + b];
                if (a == b) s *= 0.5;
            }
            P[a * n + b] = s;
        }
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) S[a * n + b] = P[a * n + b] + P[b * n + a];
    /* Y = L^-T S : solve L^T Y = S (back substitution per column) */
    for (int c = 0; c < n; ++c)
        for (int r = n - 1; r >= 0; --r) {
            double s = S[r * n + c];
            for (int k = r + 1; k < n; ++k) s -= L[k * n + r] * Y[k * n + c];
            Y[r * n + c] = s / L[r * n + r];
        }
    /* X = Y L^-1 : solve X L = Y (per row, columns from the right) */
    for (int r = 0; r < n; ++r)
        for (int c = n - 1; c >= 0; --c) {
            double s = Y[r * n + c];
            for (int k = c + 1; k < n; ++k) s -= Sbar[r * n + k] * L[k * n + c];
            Sbar[r * n + c] = s / L[c * n + c];
        }
    for (int i = 0; i < n * n; ++i) Sbar[i] *= 0.5;
}

/* kernel value k(d) (without sigma_f / jitter) and its partials wrt ell and alpha */
static void kern(const orc_problem *p, double d, double ell, double alpha, double *k, double *dk_dell,
                 double *dk_dalpha)
{
    *dk_dalpha = 0;
    if (p->kernel == 1) {
        double v = exp(-(d * d) / (2 * ell * ell));
        *k = v;
        *dk_dell = v * d * d / (ell * ell * ell);
    } else if (p->kernel == 2) {
        if (p->nu_matern == 1) {
            double v = exp(-d / ell);
            *k = v;
            *dk_dell = v * d / (ell * ell);
        } else if (p->nu_matern == 2) {
            double r = sqrt(3.0) * d / ell, e = exp(-r);
            *k = (1 + r) * e;
            *dk_dell = r * r * e / ell;
        } else {
            double r = sqrt(5.0) * d / ell, e = exp(-r);
            *k = (1 + r + (5.0 / 3.0) * (d * d) / (ell * ell)) * e;
            *dk_dell = (r * r / 3.0) * (1 + r) * e / ell;
        }
    } else {
        double w = d * d / (2 * alpha * ell * ell);
        double v = exp(-alpha * log(1 + w));
        *k = v;
        *dk_dell = 2 * alpha * w * v / ((1 + w) * ell);
        *dk_dalpha = v * (-log(1 + w) + w / (1 + w));
    }
}

typedef struct {
    /* constrained scalars */
    double la1, la2, m1, m2, th1, th2, sl1, sl2, b1, b2, ell, sigf, alpha, s, s21, s22;
    int i_la1, i_m1, i_th1, i_sl1, i_b1, i_ell, i_sigf, i_alpha, i_z, i_s, i_s21;
    double jac;
} pars_t;

/* read + constrain (h:1086-1215).  Stan Math lub_constrain / lb_constrain semantics (App. B). */
static void read_pars(const orc_problem *p, const double *u, pars_t *q)
{
    int k = 0;
    double jac = 0;
    q->la1 = q->la2 = 0;
    q->i_la1 = -1;
    if (p->est_la) {
        q->i_la1 = k;
        for (int t = 0; t < 2; ++t) {
            double x = u[k++], th;
            if (x > 0) {
                double e = exp(-x);
                th = 1.0 / (1.0 + e);
                jac += -x - 2 * log1p(e);
                if (x < INFINITY && th == 1.0) th = 1 - 1e-15;
            } else {
                double e = exp(x);
                th = 1.0 - 1.0 / (1.0 + e);
                jac += x - 2 * log1p(e);
                if (x > -INFINITY && th == 0.0) th = 1e-15;
            }
            if (t == 0) q->la1 = th; else q->la2 = th;
        }
    }
    q->i_m1 = k; q->m1 = u[k++]; q->m2 = u[k++];
    q->i_th1 = k; q->th1 = u[k++]; q->th2 = u[k++];
    q->i_sl1 = k; q->sl1 = exp(u[k]); jac += u[k++]; q->sl2 = exp(u[k]); jac += u[k++];
    q->alpha = 0; q->i_alpha = -1; q->i_b1 = q->i_ell = q->i_sigf = q->i_z = -1;
    q->b1 = q->b2 = q->ell = q->sigf = 0;
    if (p->model == 0) {
        q->i_b1 = k; q->b1 = exp(u[k]); jac += u[k++]; q->b2 = exp(u[k]); jac += u[k++];
        q->i_ell = k; q->ell = exp(u[k]); jac += u[k++];
        q->i_sigf = k; q->sigf = exp(u[k]); jac += u[k++];
        if (p->est_alpha) { q->i_alpha = k; q->alpha = exp(u[k]); jac += u[k++]; }
        q->i_z = k; k += p->n1 * p->n2;
    }
    q->i_s = k; q->s = exp(u[k]); jac += u[k++];
    q->i_s21 = k; q->s21 = exp(u[k]); jac += u[k++]; q->s22 = exp(u[k]); jac += u[k++];
    q->jac = jac;
}

typedef struct {
    int n1, n2;
    double *cov1, *cov2, *L1, *L2, *T, *GP, *p01, *p02, *p0, *Delta, *f;
    double *E1, *E2; /* 10^(slope*(x-m)) */
    double *buf;
} fwd_t;

static size_t fwd_size(int n1, int n2)
{
    int G = n1 * n2;
    return (size_t)(2 * n1 * n1 + 2 * n2 * n2 + 4 * G + 2 * (n1 + n2) + (n1 + 1) * (n2 + 1));
}

static void fwd_bind(fwd_t *F, int n1, int n2, double *buf)
{
    int G = n1 * n2;
    F->n1 = n1; F->n2 = n2; F->buf = buf;
    double *b = buf;
    F->cov1 = b; b += n1 * n1; F->L1 = b; b += n1 * n1;
    F->cov2 = b; b += n2 * n2; F->L2 = b; b += n2 * n2;
    F->T = b; b += G; F->GP = b; b += G; F->p0 = b; b += G; F->Delta = b; b += G;
    F->p01 = b; b += n1; F->p02 = b; b += n2; F->E1 = b; b += n1; F->E2 = b; b += n2;
    F->f = b;
}

/* transformed parameters (gp_grid.stan:91-164).  Matrices n2 x n1 are stored column-major
 * (index i + j*n2) like Stan/Eigen; L1/L2/cov row-major.  Returns 1 on domain error. */
static int forward_tp(const orc_problem *p, const pars_t *q, const double *u, fwd_t *F)
{
    int n1 = p->n1, n2 = p->n2;
    if (p->model == 0) {
        for (int a = 0; a < n1; ++a)
            for (int b = 0; b < n1; ++b) {
                double d = sqrt((p->x1[a] - p->x1[b]) * (p->x1[a] - p->x1[b])), k, dk, da;
                kern(p, d, q->ell, q->alpha, &k, &dk, &da);
                F->cov1[a * n1 + b] = q->sigf * k + (a == b ? 1e-10 : 0.0);
            }
        for (int a = 0; a < n2; ++a)
            for (int b = 0; b < n2; ++b) {
                double d = sqrt((p->x2[a] - p->x2[b]) * (p->x2[a] - p->x2[b])), k, dk, da;
                kern(p, d, q->ell, q->alpha, &k, &dk, &da);
                F->cov2[a * n2 + b] = k + (a == b ? 1e-10 : 0.0);
            }
        if (chol_lower(n1, F->cov1, F->L1)) return 1;
        if (chol_lower(n2, F->cov2, F->L2)) return 1;
        const double *z = u + q->i_z;
        /* kron_mvprod: (L1 * (L2 * z)')'  -- dense multiply as the reference does */
        for (int i = 0; i < n2; ++i)
            for (int j = 0; j < n1; ++j) {
                double s = 0;
                for (int k = 0; k < n2; ++k) s += F->L2[i * n2 + k] * z[k + j * n2];
                F->T[i + j * n2] = s;
            }
        for (int i = 0; i < n2; ++i)
            for (int j = 0; j < n1; ++j) {
                double s = 0;
                for (int l = 0; l < n1; ++l) s += F->L1[j * n1 + l] * F->T[i + l * n2];
                F->GP[i + j * n2] = s;
            }
    }
    for (int j = 0; j < n1; ++j) {
        F->E1[j] = pow(10.0, q->sl1 * (p->x1[j] - q->m1));
        F->p01[j] = q->la1 + (1 - q->la1) / (1 + F->E1[j]);
    }
    for (int i = 0; i < n2; ++i) {
        F->E2[i] = pow(10.0, q->sl2 * (p->x2[i] - q->m2));
        F->p02[i] = q->la2 + (1 - q->la2) / (1 + F->E2[i]);
    }
    for (int j = 0; j < n1; ++j)
        for (int i = 0; i < n2; ++i) {
            double p0 = F->p01[j] * F->p02[i];
            F->p0[i + j * n2] = p0;
            if (p->model == 0) {
                double B = F->GP[i + j * n2];
                double lg = log(p0 / (1 - p0));
                F->Delta[i + j * n2] = -p0 / (1 + exp(q->b1 * B + lg)) + (1 - p0) / (1 + exp(-q->b2 * B - lg));
            }
        }
    /* validation h:1396-1457: range checks are written as !(x >= lo) / !(x <= hi), so NaN fails too */
    for (int j = 0; j < n1; ++j)
        if (!(F->p01[j] >= 0) || !(F->p01[j] <= 1)) return 1;
    for (int i = 0; i < n2; ++i)
        if (!(F->p02[i] >= 0) || !(F->p02[i] <= 1)) return 1;
    for (int c = 0; c < n1 * n2; ++c) {
        if (!(F->p0[c] >= 0) || !(F->p0[c] <= 1)) return 1;
        if (p->model == 0 && (!(F->Delta[c] >= -1) || !(F->Delta[c] <= 1))) return 1;
    }
    /* f: (n2+1) x (n1+1), column-major index b + a*(n2+1) (gp_grid.stan:170-173) */
    int ld = n2 + 1;
    F->f[0] = 1.0;
    for (int a = 1; a <= n1; ++a) F->f[a * ld] = F->p01[a - 1];
    for (int b = 1; b <= n2; ++b) F->f[b] = F->p02[b - 1];
    for (int a = 1; a <= n1; ++a)
        for (int b = 1; b <= n2; ++b) {
            int c = (b - 1) + (a - 1) * n2;
            F->f[b + a * ld] = F->p0[c] + (p->model == 0 ? F->Delta[c] : 0.0);
        }
    return 0;
}

static double inv_gamma_lpdf(double y, double logy, double a, double b)
{
    return a * log(b) - lgamma(a) - (a + 1) * logy - b / y;
}

int orc_logp_grad(const orc_problem *p, const double *u, int jacobian, double *lp_out, double *grad)
{
    int n1 = p->n1, n2 = p->n2, G = n1 * n2, D = orc_dim(p), ld = n2 + 1;
    int gp = (p->model == 0);
    pars_t q;
    read_pars(p, u, &q);
    size_t fs = fwd_size(n1, n2);
    int nmax = n1 > n2 ? n1 : n2;
    size_t ws = fs + (size_t)(n1 + 1) * (n2 + 1) + 4 * G + 2 * (n1 + n2) + 2 * n1 * n1 + 2 * n2 * n2 +
                3 * (size_t)nmax * nmax + 16;
    double *buf = (double *)calloc(ws, sizeof(double));
    fwd_t F;
    fwd_bind(&F, n1, n2, buf);
    if (forward_tp(p, &q, u, &F)) { free(buf); return 1; }

    double lp = 0;
    /* priors gp_grid.stan:177-209 (all normalising constants kept: target += *_lpdf) */
    lp += (-LOG_PI - log1p(q.s * q.s)) - log(0.5);
    lp += inv_gamma_lpdf(q.s21, u[q.i_s21], 3, 2) + inv_gamma_lpdf(q.s22, u[q.i_s21 + 1], 3, 2);
    if (p->est_la) {
        double cb = lgamma(2.25) - lgamma(1.0) - lgamma(1.25);
        lp += cb + 0.25 * log1p(-q.la1);
        lp += cb + 0.25 * log1p(-q.la2);
    }
    lp += -q.sl1 - q.sl2;
    lp += -0.5 * q.th1 * q.th1 - LOG_2PI_HALF - 0.5 * q.th2 * q.th2 - LOG_2PI_HALF;
    {
        double sd1 = sqrt(q.s21), sd2 = sqrt(q.s22);
        double r1 = (q.m1 - q.th1) / sd1, r2 = (q.m2 - q.th2) / sd2;
        lp += -LOG_2PI_HALF - log(sd1) - 0.5 * r1 * r1;
        lp += -LOG_2PI_HALF - log(sd2) - 0.5 * r2 * r2;
    }
    double lam1 = 0, lam2 = 0;
    if (gp) {
        double r1 = (q.b1 - 1) / 0.1, r2 = (q.b2 - 1) / 0.1;
        lp += -LOG_2PI_HALF - log(0.1) - 0.5 * r1 * r1;
        lp += -LOG_2PI_HALF - log(0.1) - 0.5 * r2 * r2;
        if (p->pcprior) {
            /* gp_grid.stan:43-44 + include/pc_prior.stan:3-4 (d_half = 1) */
            lam1 = -log(p->pcprior_hypers[1]) * pow(p->pcprior_hypers[0], 1.0);
            lam2 = -log(p->pcprior_hypers[3]) * pow(p->pcprior_hypers[2], -1.0);
            lp += log(1.0) + log(lam1) + log(lam2) + (-2.0) * log(q.ell) - lam1 * pow(q.ell, -1.0) -
                  lam2 * sqrt(q.sigf) - log(sqrt(q.sigf));
        } else {
            lp += inv_gamma_lpdf(q.ell, log(q.ell), 5, 5);
            double ls = log(q.sigf);
            lp += -LOG_2PI_HALF - ls - 0.5 * (ls - 1) * (ls - 1);
        }
        if (p->est_alpha) lp += -q.alpha;
        for (int c = 0; c < G; ++c) lp += -0.5 * u[q.i_z + c] * u[q.i_z + c] - LOG_2PI_HALF;
    }
    /* likelihood gp_grid.stan:212-232 */
    double *fbar = buf + fs; /* adjoint of f, same layout */
    double sbar = 0;
    double tau = 0, lamL = 0;
    if (p->robust) lptn_consts(p->rho, &tau, &lamL);
    for (int n = 0; n < p->N; ++n) {
        int c = p->ii_obs[n] - 1;
        double f = F.f[c], y = p->y[n];
        double v = f + p->lambda;
        double sd = p->heteroscedastic ? q.s * sqrt(v) : q.s;
        if (p->heteroscedastic && !(sd > 0)) { free(buf); return 1; } /* normal_lpdf: scale must be > 0 */
        double r = (y - f) / sd;
        double dsd_df = p->heteroscedastic ? sd / (2 * v) : 0.0;
        if (!p->robust) {
            lp += -LOG_2PI_HALF - log(sd) - 0.5 * r * r;
            /* d/df = r/sd + (-1/sd + r^2/sd) dsd/df ; d/ds = -1/s + r^2/s */
            fbar[c] += r / sd + (-1.0 / sd + r * r / sd) * dsd_df;
            sbar += -1.0 / q.s + r * r / q.s;
        } else {
            double dl; /* d lptn / dr */
            if (fabs(r) <= tau) {
                lp += -0.5 * r * r - LOG_2PI_HALF;
                dl = -r;
            } else {
                double logt = log(tau);
                double az = fabs(r) > 1.0001 ? fabs(r) : 1.0001;
                double logabsz = log(az);
                lp += (-0.5 * tau * tau - LOG_2PI_HALF) + logt - logabsz + (lamL + 1) * (log(logt) - log(logabsz));
                if (fabs(r) > 1.0001)
                    dl = -(r > 0 ? 1.0 : -1.0) / az * (1 + (lamL + 1) / logabsz);
                else
                    dl = 0;
            }
            lp += -log(sd);
            double dr_df = -1.0 / sd - r / sd * dsd_df;
            fbar[c] += dl * dr_df - dsd_df / sd;
            sbar += dl * (-r / q.s) - 1.0 / q.s;
        }
    }
    if (jacobian) lp += q.jac;
    *lp_out = lp;
    if (!grad) { free(buf); return 0; }

    /* ---------------- reverse sweep ---------------- */
    memset(grad, 0, sizeof(double) * D);
    double *p0bar = fbar + (n1 + 1) * (n2 + 1);
    double *Gbar = p0bar + G, *Zbar = Gbar + G, *W = Zbar + G;
    double *p01bar = W + G, *p02bar = p01bar + n1;
    double *L1bar = p02bar + n2, *L2bar = L1bar + n1 * n1, *S1bar = L2bar + n2 * n2, *S2bar = S1bar + n1 * n1;
    double *cw = S2bar + n2 * n2;
    double b1bar = 0, b2bar = 0;
    for (int a = 1; a <= n1; ++a) p01bar[a - 1] += fbar[a * ld];
    for (int b = 1; b <= n2; ++b) p02bar[b - 1] += fbar[b];
    for (int j = 0; j < n1; ++j)
        for (int i = 0; i < n2; ++i) {
            int c = i + j * n2;
            double fb = fbar[(i + 1) + (j + 1) * ld];
            double p0 = F.p0[c];
            double dp0 = fb; /* f = p0 + Delta */
            if (gp) {
                double B = F.GP[c], lg = log(p0 / (1 - p0));
                double t1 = 1 / (1 + exp(q.b1 * B + lg)), t2 = 1 / (1 + exp(-q.b2 * B - lg));
                double w1 = p0 * t1 * (1 - t1), w2 = (1 - p0) * t2 * (1 - t2);
                Gbar[c] = fb * (w1 * q.b1 + w2 * q.b2);
                b1bar += fb * w1 * B;
                b2bar += fb * w2 * B;
                dp0 += fb * (-t1 - t2 + t1 * (1 - t1) / (1 - p0) + t2 * (1 - t2) / p0);
            }
            p0bar[c] = dp0;
            p01bar[j] += dp0 * F.p02[i];
            p02bar[i] += dp0 * F.p01[j];
        }
    /* monotherapy curves */
    double la1bar = 0, la2bar = 0, m1bar = 0, m2bar = 0, sl1bar = 0, sl2bar = 0;
    for (int j = 0; j < n1; ++j) {
        double E = F.E1[j], o = 1 + E, dx = p->x1[j] - q.m1;
        la1bar += p01bar[j] * (E / o);
        sl1bar += p01bar[j] * (-(1 - q.la1) * E * LN10 * dx / (o * o));
        m1bar += p01bar[j] * ((1 - q.la1) * E * LN10 * q.sl1 / (o * o));
    }
    for (int i = 0; i < n2; ++i) {
        double E = F.E2[i], o = 1 + E, dx = p->x2[i] - q.m2;
        la2bar += p02bar[i] * (E / o);
        sl2bar += p02bar[i] * (-(1 - q.la2) * E * LN10 * dx / (o * o));
        m2bar += p02bar[i] * ((1 - q.la2) * E * LN10 * q.sl2 / (o * o));
    }
    double ellbar = 0, sigfbar = 0, alphabar = 0;
    if (gp) {
        const double *z = u + q.i_z;
        /* W = Gbar * L1 (n2 x n1);  Zbar = L2^T W;  L2bar = tril(W Z^T);  L1bar = tril(Gbar^T T) */
        for (int i = 0; i < n2; ++i)
            for (int l = 0; l < n1; ++l) {
                double s = 0;
                for (int j = 0; j < n1; ++j) s += Gbar[i + j * n2] * F.L1[j * n1 + l];
                W[i + l * n2] = s;
            }
        for (int k = 0; k < n2; ++k)
            for (int l = 0; l < n1; ++l) {
                double s = 0;
                for (int i = 0; i < n2; ++i) s += F.L2[i * n2 + k] * W[i + l * n2];
                Zbar[k + l * n2] = s;
            }
        for (int i = 0; i < n2; ++i)
            for (int k = 0; k < n2; ++k) {
                double s = 0;
                if (k <= i)
                    for (int l = 0; l < n1; ++l) s += W[i + l * n2] * z[k + l * n2];
                L2bar[i * n2 + k] = s;
            }
        for (int j = 0; j < n1; ++j)
            for (int l = 0; l < n1; ++l) {
                double s = 0;
                if (l <= j)
                    for (int i = 0; i < n2; ++i) s += Gbar[i + j * n2] * F.T[i + l * n2];
                L1bar[j * n1 + l] = s;
            }
        chol_adjoint(n1, F.L1, L1bar, S1bar, cw);
        chol_adjoint(n2, F.L2, L2bar, S2bar, cw);
        for (int a = 0; a < n1; ++a)
            for (int b = 0; b < n1; ++b) {
                double d = sqrt((p->x1[a] - p->x1[b]) * (p->x1[a] - p->x1[b])), k, dk, da;
                kern(p, d, q.ell, q.alpha, &k, &dk, &da);
                ellbar += S1bar[a * n1 + b] * q.sigf * dk;
                sigfbar += S1bar[a * n1 + b] * k;
                alphabar += S1bar[a * n1 + b] * q.sigf * da;
            }
        for (int a = 0; a < n2; ++a)
            for (int b = 0; b < n2; ++b) {
                double d = sqrt((p->x2[a] - p->x2[b]) * (p->x2[a] - p->x2[b])), k, dk, da;
                kern(p, d, q.ell, q.alpha, &k, &dk, &da);
                ellbar += S2bar[a * n2 + b] * dk;
                alphabar += S2bar[a * n2 + b] * da;
            }
        for (int c = 0; c < G; ++c) grad[q.i_z + c] = Zbar[c] - z[c];
        /* prior adjoints of the GP scalars */
        b1bar += -(q.b1 - 1) / 0.01;
        b2bar += -(q.b2 - 1) / 0.01;
        if (p->pcprior) {
            ellbar += -2.0 / q.ell + lam1 / (q.ell * q.ell);
            sigfbar += -lam2 * 0.5 / sqrt(q.sigf) - 0.5 / q.sigf;
        } else {
            ellbar += -6.0 / q.ell + 5.0 / (q.ell * q.ell);
            sigfbar += -1.0 / q.sigf - (log(q.sigf) - 1) / q.sigf;
        }
        if (p->est_alpha) alphabar += -1.0;
    }
    /* remaining prior adjoints */
    sbar += -2 * q.s / (1 + q.s * q.s);
    double s21bar = -4.0 / q.s21 + 2.0 / (q.s21 * q.s21);
    double s22bar = -4.0 / q.s22 + 2.0 / (q.s22 * q.s22);
    if (p->est_la) {
        la1bar += -0.25 / (1 - q.la1);
        la2bar += -0.25 / (1 - q.la2);
    }
    sl1bar += -1;
    sl2bar += -1;
    double th1bar = -q.th1, th2bar = -q.th2;
    {
        double d1 = q.m1 - q.th1, d2 = q.m2 - q.th2;
        m1bar += -d1 / q.s21;
        m2bar += -d2 / q.s22;
        th1bar += d1 / q.s21;
        th2bar += d2 / q.s22;
        s21bar += -0.5 / q.s21 + 0.5 * d1 * d1 / (q.s21 * q.s21);
        s22bar += -0.5 / q.s22 + 0.5 * d2 * d2 / (q.s22 * q.s22);
    }
    /* chain through the constraining transforms */
    double J = jacobian ? 1.0 : 0.0;
    if (p->est_la) {
        grad[q.i_la1] = la1bar * q.la1 * (1 - q.la1) + J * (1 - 2 * q.la1);
        grad[q.i_la1 + 1] = la2bar * q.la2 * (1 - q.la2) + J * (1 - 2 * q.la2);
    }
    grad[q.i_m1] = m1bar; grad[q.i_m1 + 1] = m2bar;
    grad[q.i_th1] = th1bar; grad[q.i_th1 + 1] = th2bar;
    grad[q.i_sl1] = sl1bar * q.sl1 + J; grad[q.i_sl1 + 1] = sl2bar * q.sl2 + J;
    if (gp) {
        grad[q.i_b1] = b1bar * q.b1 + J; grad[q.i_b1 + 1] = b2bar * q.b2 + J;
        grad[q.i_ell] = ellbar * q.ell + J;
        grad[q.i_sigf] = sigfbar * q.sigf + J;
        if (p->est_alpha) grad[q.i_alpha] = alphabar * q.alpha + J;
    }
    grad[q.i_s] = sbar * q.s + J;
    grad[q.i_s21] = s21bar * q.s21 + J;
    grad[q.i_s21 + 1] = s22bar * q.s22 + J;
    free(buf);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* RNG for the oracle: splitmix64 -> uniform / normal.  (The reference uses boost::ecuyer1988;
 * only statistical equivalence is required, SURVEY.md App. D.) */
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

/* include/VUS.stan; Z is n2 x n1 column-major */
static double VUS(const double *Z, const double *x1, const double *x2, int n1, int n2)
{
    double r = 0, Bprev = 0;
    for (int i = 0; i < n2; ++i) {
        double b = 0;
        for (int j = 1; j < n1; ++j) b += (x1[j] - x1[j - 1]) * (Z[i + j * n2] + Z[i + (j - 1) * n2]) / 2;
        if (i > 0) r += (x2[i] - x2[i - 1]) * (b + Bprev) / 2;
        Bprev = b;
    }
    double mx1 = x1[0], mn1 = x1[0], mx2 = x2[0], mn2 = x2[0];
    for (int j = 0; j < n1; ++j) { if (x1[j] > mx1) mx1 = x1[j]; if (x1[j] < mn1) mn1 = x1[j]; }
    for (int i = 0; i < n2; ++i) { if (x2[i] > mx2) mx2 = x2[i]; if (x2[i] < mn2) mn2 = x2[i]; }
    return 100 * r / ((mx1 - mn1) * (mx2 - mn2));
}

/* include/dss.stan */
static double dss(double c1, double c2, double la, double slope, double m)
{
    double val = (c2 - c1) + (la - 1) / slope *
                 (log10(1 + pow(10.0, slope * (c2 - m))) - log10(1 + pow(10.0, slope * (c1 - m))));
    return 100 * (1 - val / (c2 - c1));
}

int orc_write_array(const orc_problem *p, const double *u, double *out, uint64_t *rng)
{
    int n1 = p->n1, n2 = p->n2, G = n1 * n2;
    int gp = (p->model == 0);
    pars_t q;
    read_pars(p, u, &q);
    double *buf = (double *)calloc(fwd_size(n1, n2) + 2 * G + 8, sizeof(double));
    fwd_t F;
    fwd_bind(&F, n1, n2, buf);
    int k = 0;
    if (p->est_la) { out[k++] = q.la1; out[k++] = q.la2; }
    out[k++] = q.m1; out[k++] = q.m2; out[k++] = q.th1; out[k++] = q.th2; out[k++] = q.sl1; out[k++] = q.sl2;
    if (gp) {
        out[k++] = q.b1; out[k++] = q.b2; out[k++] = q.ell; out[k++] = q.sigf;
        if (p->est_alpha) out[k++] = q.alpha;
        for (int c = 0; c < G; ++c) out[k++] = u[q.i_z + c];
    }
    out[k++] = q.s; out[k++] = q.s21; out[k++] = q.s22;
    int nout = orc_num_outputs(p);
    if (forward_tp(p, &q, u, &F)) {
        for (; k < nout; ++k) out[k] = NAN;
        free(buf);
        return 1;
    }
    for (int c = 0; c < G; ++c) out[k++] = F.p0[c];
    for (int j = 0; j < n1; ++j) out[k++] = F.p01[j];
    for (int i = 0; i < n2; ++i) out[k++] = F.p02[i];
    if (gp) {
        for (int c = 0; c < G; ++c) out[k++] = F.Delta[c];
        for (int c = 0; c < G; ++c) out[k++] = F.GP[c];
    }
    out[k++] = pow(10.0, q.m1);
    out[k++] = pow(10.0, q.m2);
    for (int n = 0; n < p->N; ++n) {
        double f = F.f[p->ii_obs[n] - 1];
        double sd = p->heteroscedastic ? q.s * sqrt(f + p->lambda) : q.s;
        double r = (p->y[n] - f) / sd;
        out[k++] = exp(-(-LOG_2PI_HALF - log(sd) - 0.5 * r * r));
    }
    double c11 = p->x1[0], c12 = p->x1[0], c21 = p->x2[0], c22 = p->x2[0];
    for (int j = 0; j < n1; ++j) { if (p->x1[j] < c11) c11 = p->x1[j]; if (p->x1[j] > c12) c12 = p->x1[j]; }
    for (int i = 0; i < n2; ++i) { if (p->x2[i] < c21) c21 = p->x2[i]; if (p->x2[i] > c22) c22 = p->x2[i]; }
    out[k++] = dss(c11, c12, q.la1, q.sl1, q.m1);
    out[k++] = dss(c21, c22, q.la2, q.sl2, q.m2);
    double *t = buf + fwd_size(n1, n2), *t2 = t + G;
    for (int c = 0; c < G; ++c) t[c] = 1 - (F.p0[c] + (gp ? F.Delta[c] : 0.0));
    double rVUS_f = VUS(t, p->x1, p->x2, n1, n2);
    for (int c = 0; c < G; ++c) t[c] = 1 - F.p0[c];
    double rVUS_p0 = VUS(t, p->x1, p->x2, n1, n2);
    double VUS_Delta = 0, VUS_syn = 0, VUS_ant = 0;
    if (gp) {
        VUS_Delta = VUS(F.Delta, p->x1, p->x2, n1, n2);
        for (int c = 0; c < G; ++c) { t[c] = fmin(F.Delta[c], 0.0); t2[c] = fmax(F.Delta[c], 0.0); }
        VUS_syn = VUS(t, p->x1, p->x2, n1, n2);
        VUS_ant = VUS(t2, p->x1, p->x2, n1, n2);
    }
    if (rVUS_f == 0) rVUS_f = 1e-6 + (1e-4 - 1e-6) * runif(rng);
    if (rVUS_p0 == 0) rVUS_p0 = 1e-6 + (1e-4 - 1e-6) * runif(rng);
    if (gp) {
        if (VUS_syn == 0) VUS_syn = 1e-6 + (1e-4 - 1e-6) * runif(rng);
        if (VUS_ant == 0) VUS_ant = 1e-6 + (1e-4 - 1e-6) * runif(rng);
    }
    out[k++] = rVUS_f;
    out[k++] = rVUS_p0;
    if (gp) { out[k++] = VUS_Delta; out[k++] = VUS_syn; out[k++] = VUS_ant; }
    free(buf);
    return (k == nout) ? 0 : 2;
}

/* ------------------------------------------------------------------------------------------ */
/* NUTS, diag_e, Stan >= 2.21 semantics (multinomial sampling, extra cross-subtree U-turn checks) */

void orc_nuts_defaults(orc_nuts_args *a)
{
    a->num_warmup = 1000; a->num_samples = 1000; a->max_treedepth = 10;
    a->adapt_delta = 0.9;  /* bayesynergy override, R/bayesynergy.R:107-110 */
    a->stepsize = 1.0; a->stepsize_jitter = 0.0;
    a->gamma = 0.05; a->kappa = 0.75; a->t0 = 10;
    a->init_buffer = 75; a->term_buffer = 50; a->window = 25;
    a->init_radius = 2.0; a->seed = 1; a->chain_id = 1; a->save_warmup = 0; a->metric = 0;
}

typedef struct {
    const orc_problem *p;
    int D;
    const double *minv; /* diagonal inverse metric (diag_e) */
    long n_grad;
    int dense;          /* dense_e_metric: inverse metric Minv [D*D] (symmetric) and its Cholesky factor Lm (lower, row-major) */
    const double *Minv, *Lm;
} ham_t;

/* p_sharp = M^-1 p  (diag_e_metric / dense_e_metric ::dtau_dp) */
static void sharp(const ham_t *H, const double *p, double *out)
{
    int D = H->D;
    if (!H->dense) { for (int i = 0; i < D; ++i) out[i] = H->minv[i] * p[i]; return; }
    for (int i = 0; i < D; ++i) {
        double s = 0;
        for (int j = 0; j < D; ++j) s += H->Minv[(size_t)i * D + j] * p[j];
        out[i] = s;
    }
}

typedef struct { double *q, *p, *g; double V; } ps_t; /* V = -lp */

static void ps_alloc(ps_t *z, int D) { z->q = (double *)malloc(sizeof(double) * 3 * D); z->p = z->q + D; z->g = z->p + D; z->V = 0; }
static void ps_copy(ps_t *d, const ps_t *s, int D) { memcpy(d->q, s->q, sizeof(double) * 3 * D); d->V = s->V; }

/* potential + gradient; domain error => V = +inf, g = 0 (Stan: exception => infinite potential) */
static void update_pg(ham_t *H, ps_t *z)
{
    double lp;
    H->n_grad++;
    if (orc_logp_grad(H->p, z->q, 1, &lp, z->g) || !isfinite(lp)) {
        z->V = INFINITY;
        for (int i = 0; i < H->D; ++i) z->g[i] = 0;
        return;
    }
    z->V = -lp;
    for (int i = 0; i < H->D; ++i) z->g[i] = -z->g[i]; /* gradient of the potential */
}
static double kinetic(const ham_t *H, const ps_t *z)
{
    double s = 0;
    if (!H->dense) {
        for (int i = 0; i < H->D; ++i) s += z->p[i] * z->p[i] * H->minv[i];
    } else {
        double *t = (double *)malloc(sizeof(double) * H->D);
        sharp(H, z->p, t);
        for (int i = 0; i < H->D; ++i) s += z->p[i] * t[i];
        free(t);
    }
    return 0.5 * s;
}
static double hamil(const ham_t *H, const ps_t *z) { return z->V + kinetic(H, z); }
static void leapfrog(ham_t *H, ps_t *z, double eps)
{
    int D = H->D;
    for (int i = 0; i < D; ++i) z->p[i] -= 0.5 * eps * z->g[i];
    if (!H->dense) {
        for (int i = 0; i < D; ++i) z->q[i] += eps * H->minv[i] * z->p[i];
    } else {
        double *t = (double *)malloc(sizeof(double) * D);
        sharp(H, z->p, t);
        for (int i = 0; i < D; ++i) z->q[i] += eps * t[i];
        free(t);
    }
    update_pg(H, z);
    for (int i = 0; i < D; ++i) z->p[i] -= 0.5 * eps * z->g[i];
}
static double log_sum_exp2(double a, double b)
{
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    return a > b ? a + log1p(exp(b - a)) : b + log1p(exp(a - b));
}
static int criterion(const ham_t *H, const double *ps_a, const double *ps_b, const double *rho)
{
    double a = 0, b = 0;
    for (int i = 0; i < H->D; ++i) { a += ps_a[i] * rho[i]; b += ps_b[i] * rho[i]; }
    return a > 0 && b > 0;
}

typedef struct {
    ham_t *H; double eps, H0; int sign; int n_leapfrog; double sum_metro; int divergent; uint64_t *rng;
} tree_ctx;

/* base_nuts::build_tree */
static int build_tree(tree_ctx *T, int depth, ps_t *z, ps_t *z_propose, double *psharp_beg, double *psharp_end,
                      double *rho, double *p_beg, double *p_end, double *log_sum_weight)
{
    ham_t *H = T->H;
    int D = H->D;
    if (depth == 0) {
        leapfrog(H, z, T->sign * T->eps);
        T->n_leapfrog++;
        double h = hamil(H, z);
        if (isnan(h)) h = INFINITY;
        if (h - T->H0 > 1000) T->divergent = 1;
        *log_sum_weight = log_sum_exp2(*log_sum_weight, T->H0 - h);
        T->sum_metro += (T->H0 - h > 0) ? 1.0 : exp(T->H0 - h);
        ps_copy(z_propose, z, D);
        sharp(H, z->p, psharp_beg);
        for (int i = 0; i < D; ++i) {
            psharp_end[i] = psharp_beg[i];
            rho[i] += z->p[i];
            p_beg[i] = p_end[i] = z->p[i];
        }
        return !T->divergent;
    }
    double *w = (double *)calloc(6 * (size_t)D, sizeof(double));
    double *psharp_init_end = w, *psharp_final_beg = w + D, *rho_init = w + 2 * D, *rho_final = w + 3 * D;
    double *p_init_end = w + 4 * D, *p_final_beg = w + 5 * D;
    double lsw_init = -INFINITY, lsw_final = -INFINITY;
    int ok = build_tree(T, depth - 1, z, z_propose, psharp_beg, psharp_init_end, rho_init, p_beg, p_init_end, &lsw_init);
    if (!ok) { free(w); return 0; }
    ps_t zpf;
    ps_alloc(&zpf, D);
    ps_copy(&zpf, z, D);
    ok = build_tree(T, depth - 1, z, &zpf, psharp_final_beg, psharp_end, rho_final, p_final_beg, p_end, &lsw_final);
    if (!ok) { free(zpf.q); free(w); return 0; }
    double lsw_sub = log_sum_exp2(lsw_init, lsw_final);
    *log_sum_weight = log_sum_exp2(*log_sum_weight, lsw_sub);
    if (lsw_final > lsw_sub) {
        ps_copy(z_propose, &zpf, D);
    } else {
        if (runif(T->rng) < exp(lsw_final - lsw_sub)) ps_copy(z_propose, &zpf, D);
    }
    free(zpf.q);
    double *rho_sub = (double *)malloc(sizeof(double) * 2 * D), *rho_ext = rho_sub + D;
    for (int i = 0; i < D; ++i) { rho_sub[i] = rho_init[i] + rho_final[i]; rho[i] += rho_sub[i]; }
    int persist = criterion(H, psharp_beg, psharp_end, rho_sub);
    for (int i = 0; i < D; ++i) rho_ext[i] = rho_init[i] + p_final_beg[i];
    persist &= criterion(H, psharp_beg, psharp_final_beg, rho_ext);
    for (int i = 0; i < D; ++i) rho_ext[i] = rho_final[i] + p_init_end[i];
    persist &= criterion(H, psharp_init_end, psharp_end, rho_ext);
    free(rho_sub);
    free(w);
    return persist;
}

typedef struct { double accept, stepsize; int depth, n_leapfrog, divergent; double energy; } trans_info;

/* diag_e_metric::sample_p: p_i = N(0,1) / sqrt(minv_i);  dense_e_metric::sample_p: p = L^-T z with M^-1 = L L^T */
static void sample_p(const ham_t *H, double *p, uint64_t *rng)
{
    int D = H->D;
    if (!H->dense) { for (int i = 0; i < D; ++i) p[i] = rnorm(rng) / sqrt(H->minv[i]); return; }
    for (int i = 0; i < D; ++i) p[i] = rnorm(rng);
    for (int i = D - 1; i >= 0; --i) {      /* back substitution on L^T */
        double s = p[i];
        for (int j = i + 1; j < D; ++j) s -= H->Lm[(size_t)j * D + i] * p[j];
        p[i] = s / H->Lm[(size_t)i * D + i];
    }
}

/* base_nuts::transition; z holds the current state (q, g, V) and is overwritten with the sample */
static void nuts_transition(ham_t *H, ps_t *z, double eps, int max_depth, uint64_t *rng, trans_info *info)
{
    int D = H->D;
    sample_p(H, z->p, rng);
    ps_t z_fwd, z_bck, z_sample, z_propose;
    ps_alloc(&z_fwd, D); ps_alloc(&z_bck, D); ps_alloc(&z_sample, D); ps_alloc(&z_propose, D);
    ps_copy(&z_fwd, z, D); ps_copy(&z_bck, z, D); ps_copy(&z_sample, z, D); ps_copy(&z_propose, z, D);
    double *w = (double *)calloc(11 * (size_t)D, sizeof(double));
    double *p_fwd_fwd = w, *ps_fwd_fwd = w + D, *p_fwd_bck = w + 2 * D, *ps_fwd_bck = w + 3 * D;
    double *p_bck_fwd = w + 4 * D, *ps_bck_fwd = w + 5 * D, *p_bck_bck = w + 6 * D, *ps_bck_bck = w + 7 * D;
    double *rho = w + 8 * D, *rho_fwd = w + 9 * D, *rho_bck = w + 10 * D;
    sharp(H, z->p, ps_fwd_fwd);
    for (int i = 0; i < D; ++i) {
        p_fwd_fwd[i] = p_fwd_bck[i] = p_bck_fwd[i] = p_bck_bck[i] = rho[i] = z->p[i];
        ps_fwd_bck[i] = ps_bck_fwd[i] = ps_bck_bck[i] = ps_fwd_fwd[i];
    }
    double lsw = 0, H0 = hamil(H, z);
    tree_ctx T = {H, eps, H0, 1, 0, 0.0, 0, rng};
    int depth = 0;
    while (depth < max_depth) {
        for (int i = 0; i < D; ++i) rho_fwd[i] = rho_bck[i] = 0;
        int valid;
        double lsw_sub = -INFINITY;
        if (runif(rng) > 0.5) {
            /* extend forward: the old trajectory becomes the backward subtree */
            memcpy(rho_bck, rho, sizeof(double) * D);
            memcpy(p_bck_fwd, p_fwd_fwd, sizeof(double) * D);
            memcpy(ps_bck_fwd, ps_fwd_fwd, sizeof(double) * D);
            T.sign = 1;
            valid = build_tree(&T, depth, &z_fwd, &z_propose, ps_fwd_bck, ps_fwd_fwd, rho_fwd, p_fwd_bck,
                               p_fwd_fwd, &lsw_sub);
        } else {
            /* extend backward: the old trajectory becomes the forward subtree */
            memcpy(rho_fwd, rho, sizeof(double) * D);
            memcpy(p_fwd_bck, p_bck_bck, sizeof(double) * D);
            memcpy(ps_fwd_bck, ps_bck_bck, sizeof(double) * D);
            T.sign = -1;
            valid = build_tree(&T, depth, &z_bck, &z_propose, ps_bck_fwd, ps_bck_bck, rho_bck, p_bck_fwd,
                               p_bck_bck, &lsw_sub);
        }
        if (!valid) break;
        ++depth;
        if (lsw_sub > lsw) {
            ps_copy(&z_sample, &z_propose, D);
        } else {
            if (runif(rng) < exp(lsw_sub - lsw)) ps_copy(&z_sample, &z_propose, D);
        }
        lsw = log_sum_exp2(lsw, lsw_sub);
        for (int i = 0; i < D; ++i) rho[i] = rho_bck[i] + rho_fwd[i];
        int persist = criterion(H, ps_bck_bck, ps_fwd_fwd, rho);
        double *rho_ext = (double *)malloc(sizeof(double) * D);
        for (int i = 0; i < D; ++i) rho_ext[i] = rho_bck[i] + p_fwd_bck[i];
        persist &= criterion(H, ps_bck_bck, ps_fwd_bck, rho_ext);
        for (int i = 0; i < D; ++i) rho_ext[i] = rho_fwd[i] + p_bck_fwd[i];
        persist &= criterion(H, ps_bck_fwd, ps_fwd_fwd, rho_ext);
        free(rho_ext);
        if (!persist) break;
    }
    info->accept = T.sum_metro / (double)T.n_leapfrog;
    info->depth = depth;
    info->n_leapfrog = T.n_leapfrog;
    info->divergent = T.divergent;
    info->stepsize = eps;
    ps_copy(z, &z_sample, D);
    info->energy = hamil(H, z);
    free(w); free(z_fwd.q); free(z_bck.q); free(z_sample.q); free(z_propose.q);
}

/* base_hmc::init_stepsize */
static double init_stepsize(ham_t *H, ps_t *z, double eps, uint64_t *rng)
{
    int D = H->D;
    ps_t z0;
    ps_alloc(&z0, D);
    ps_copy(&z0, z, D);
    sample_p(H, z->p, rng);
    update_pg(H, z);
    double H0 = hamil(H, z);
    leapfrog(H, z, eps);
    double h = hamil(H, z);
    if (isnan(h)) h = INFINITY;
    double dH = H0 - h;
    int dir = dH > log(0.8) ? 1 : -1;
    while (1) {
        ps_copy(z, &z0, D);
        sample_p(H, z->p, rng);
        update_pg(H, z);
        H0 = hamil(H, z);
        leapfrog(H, z, eps);
        h = hamil(H, z);
        if (isnan(h)) h = INFINITY;
        dH = H0 - h;
        if (dir == 1 && !(dH > log(0.8))) break;
        if (dir == -1 && !(dH < log(0.8))) break;
        eps = dir == 1 ? 2 * eps : 0.5 * eps;
        if (eps > 1e7 || eps == 0) break;
    }
    ps_copy(z, &z0, D);
    free(z0.q);
    return eps;
}

long orc_nuts_chain(const orc_problem *p, const orc_nuts_args *a, double *draws_u, double *outputs,
                    double *sampler_params, double *final_stepsize, double *final_inv_metric)
{
    int D = orc_dim(p), nout = orc_num_outputs(p);
    uint64_t rng = a->seed * 0x9E3779B97F4A7C15ULL + (uint64_t)a->chain_id * 0xD1B54A32D192ED03ULL + 12345;
    double *minv = (double *)malloc(sizeof(double) * D);
    for (int i = 0; i < D; ++i) minv[i] = 1.0;
    const int dense = a->metric == 1;
    double *Minv = NULL, *Lm = NULL, *cm = NULL, *cm2 = NULL;   /* dense_e: M^-1, chol(M^-1), Welford mean / cross products */
    if (dense) {
        Minv = (double *)calloc(2 * (size_t)D * D, sizeof(double)); Lm = Minv + (size_t)D * D;
        cm = (double *)calloc((size_t)D + (size_t)D * D, sizeof(double)); cm2 = cm + D;
        for (int i = 0; i < D; ++i) Minv[(size_t)i * D + i] = Lm[(size_t)i * D + i] = 1.0;
    }
    ham_t H = {p, D, minv, 0, dense, Minv, Lm};
    ps_t z;
    ps_alloc(&z, D);
    /* services::util::initialize: u ~ U(-R, R) until lp and gradient are finite (<= 100 tries) */
    int ok = 0;
    for (int t = 0; t < 100 && !ok; ++t) {
        for (int i = 0; i < D; ++i) z.q[i] = a->init_radius * (2 * runif(&rng) - 1);
        double lp;
        if (orc_logp_grad(p, z.q, 1, &lp, z.g)) continue;
        if (!isfinite(lp)) continue;
        ok = 1;
        for (int i = 0; i < D; ++i) if (!isfinite(z.g[i])) ok = 0;
    }
    if (!ok) { free(minv); free(z.q); return -1; }
    for (int i = 0; i < D; ++i) z.p[i] = 0;
    update_pg(&H, &z);
    double eps = init_stepsize(&H, &z, a->stepsize, &rng);
    /* dual averaging state */
    double mu = log(10 * eps), sbar = 0, xbar = 0, cnt = 0;
    /* windowed variance adaptation */
    int nw = a->num_warmup;
    int init_buffer = a->init_buffer, term_buffer = a->term_buffer, base_window = a->window;
    if (nw < 20) { init_buffer = term_buffer = base_window = 0; }
    else if (init_buffer + base_window + term_buffer > nw) {
        init_buffer = (int)(0.15 * nw); term_buffer = (int)(0.1 * nw);
        base_window = nw - (init_buffer + term_buffer);
    }
    int win_size = base_window, win_end = init_buffer + base_window - 1;
    double *wm = (double *)calloc(2 * (size_t)D, sizeof(double)), *wm2 = wm + D;
    double wn = 0;
    int total = a->num_warmup + a->num_samples, isave = 0;
    for (int it = 0; it < total; ++it) {
        int warm = it < a->num_warmup;
        double e = eps;
        if (a->stepsize_jitter > 0) e = eps * (1 + a->stepsize_jitter * (2 * runif(&rng) - 1));
        trans_info info;
        nuts_transition(&H, &z, e, a->max_treedepth, &rng, &info);
        if (warm) {
            /* stepsize_adaptation::learn_stepsize */
            cnt += 1;
            double adapt_stat = info.accept > 1 ? 1 : info.accept;
            double eta = 1.0 / (cnt + a->t0);
            sbar = (1 - eta) * sbar + eta * (a->adapt_delta - adapt_stat);
            double x = mu - sbar * sqrt(cnt) / a->gamma;
            double xeta = pow(cnt, -a->kappa);
            xbar = (1 - xeta) * xbar + xeta * x;
            eps = exp(x);
            /* windowed_adaptation / var_adaptation::learn_variance */
            int in_window = (it >= init_buffer) && (it < nw - term_buffer) && (it != nw);
            if (in_window && !dense) {
                wn += 1;
                for (int i = 0; i < D; ++i) {
                    double d = z.q[i] - wm[i];
                    wm[i] += d / wn;
                    wm2[i] += (z.q[i] - wm[i]) * d;
                }
            }
            if (in_window && dense) {   /* welford_covar_estimator::add_sample */
                wn += 1;
                double *dl = (double *)malloc(sizeof(double) * D);
                for (int i = 0; i < D; ++i) { dl[i] = z.q[i] - cm[i]; cm[i] += dl[i] / wn; }
                for (int i = 0; i < D; ++i)
                    for (int j = 0; j < D; ++j) cm2[(size_t)i * D + j] += (z.q[i] - cm[i]) * dl[j];
                free(dl);
            }
            if (in_window && it == win_end) {
                if (!dense) {
                    for (int i = 0; i < D; ++i) {
                        double var = wm2[i] / (wn - 1.0);
                        minv[i] = (wn / (wn + 5.0)) * var + 1e-3 * (5.0 / (wn + 5.0));
                    }
                } else {   /* covar_adaptation::learn_covariance, then the factor dense_e_metric::sample_p needs */
                    for (int i = 0; i < D; ++i)
                        for (int j = 0; j < D; ++j)
                            Minv[(size_t)i * D + j] = (wn / (wn + 5.0)) * cm2[(size_t)i * D + j] / (wn - 1.0) +
                                                      (i == j ? 1e-3 * (5.0 / (wn + 5.0)) : 0.0);
                    for (int i = 0; i < D; ++i)           /* symmetrise (the Welford update is not exactly symmetric) */
                        for (int j = 0; j < i; ++j) {
                            double v = 0.5 * (Minv[(size_t)i * D + j] + Minv[(size_t)j * D + i]);
                            Minv[(size_t)i * D + j] = Minv[(size_t)j * D + i] = v;
                        }
                    for (int i = 0; i < D; ++i)
                        for (int j = 0; j <= i; ++j) {
                            double sum = Minv[(size_t)i * D + j];
                            for (int k = 0; k < j; ++k) sum -= Lm[(size_t)i * D + k] * Lm[(size_t)j * D + k];
                            Lm[(size_t)i * D + j] = (i == j) ? sqrt(sum) : sum / Lm[(size_t)j * D + j];
                        }
                    memset(cm, 0, sizeof(double) * ((size_t)D + (size_t)D * D));
                }
                wn = 0;
                memset(wm, 0, sizeof(double) * 2 * D);
                /* compute_next_window */
                if (win_end != nw - term_buffer - 1) {
                    win_size *= 2;
                    win_end = it + win_size;
                    if (win_end != nw - term_buffer - 1) {
                        int next_boundary = win_end + 2 * win_size;
                        if (next_boundary >= nw - term_buffer) win_end = nw - term_buffer - 1;
                    }
                }
                eps = init_stepsize(&H, &z, eps, &rng);
                mu = log(10 * eps); sbar = 0; xbar = 0; cnt = 0;
            }
            if (it == a->num_warmup - 1) eps = exp(xbar); /* complete_adaptation */
        }
        if (!warm || a->save_warmup) {
            if (draws_u) memcpy(draws_u + (size_t)isave * D, z.q, sizeof(double) * D);
            if (outputs) orc_write_array(p, z.q, outputs + (size_t)isave * nout, &rng);
            if (sampler_params) {
                double *sp = sampler_params + (size_t)isave * 7;
                sp[0] = info.accept; sp[1] = info.stepsize; sp[2] = info.depth; sp[3] = info.n_leapfrog;
                sp[4] = info.divergent; sp[5] = info.energy; sp[6] = -z.V;
            }
            ++isave;
        }
    }
    if (final_stepsize) *final_stepsize = eps;
    if (final_inv_metric) {
        if (!dense) memcpy(final_inv_metric, minv, sizeof(double) * D);
        else for (int i = 0; i < D; ++i) final_inv_metric[i] = Minv[(size_t)i * D + i];
    }
    long ng = H.n_grad;
    free(wm); free(minv); free(z.q); free(Minv); free(cm);
    return ng;
}
