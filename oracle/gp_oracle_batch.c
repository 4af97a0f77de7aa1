/* ORACLE (test infrastructure, not product code) -- PARITY UNPINNED.  See gp_oracle.h.
 *
 * Host-core batch drivers around the scalar oracle: the stand-in for
 * synergyscreen(parallel=TRUE)'s one-experiment-per-worker loop (R/synergyscreen.R:117-130),
 * here one chain per worker thread (pthreads, dynamic work queue).  Used by bench.py's
 * cpu_baseline / --impl reference legs and by tests.
 */
#include "gp_oracle.h"

#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

int orc_num_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct {
    atomic_int next;
    int n_items;
    void (*fn)(int item, void *ctx);
    void *ctx;
} pool_t;

static void *pool_worker(void *arg)
{
    pool_t *P = (pool_t *)arg;
    for (;;) {
        int i = atomic_fetch_add(&P->next, 1);
        if (i >= P->n_items) break;
        P->fn(i, P->ctx);
    }
    return NULL;
}

static void pool_run(int n_items, int nthreads, void (*fn)(int, void *), void *ctx)
{
    if (nthreads <= 0) nthreads = orc_num_threads();
    if (nthreads > n_items) nthreads = n_items;
    pool_t P;
    atomic_init(&P.next, 0);
    P.n_items = n_items; P.fn = fn; P.ctx = ctx;
    if (nthreads <= 1) { pool_worker(&P); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, pool_worker, &P);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th);
}

/* ---- lp/grad at B points of one problem; status[b] = 0 ok / 1 domain error ---- */
typedef struct {
    const orc_problem *p; const double *u; int jacobian; double *lp, *grad; int *status; int D;
} lg_ctx;
static void lg_item(int b, void *c)
{
    lg_ctx *C = (lg_ctx *)c;
    C->status[b] = orc_logp_grad(C->p, C->u + (size_t)b * C->D, C->jacobian, C->lp + b,
                                 C->grad ? C->grad + (size_t)b * C->D : NULL);
}
void orc_logp_grad_batch(const orc_problem *p, int B, const double *u, int jacobian, double *lp, double *grad,
                         int *status, int nthreads)
{
    lg_ctx C = {p, u, jacobian, lp, grad, status, orc_dim(p)};
    pool_run(B, nthreads, lg_item, &C);
}

/* ---- n_prob problems x n_chains chains of adaptive NUTS, one chain per worker ----
 * outputs: [n_prob][n_chains][num_samples][n_out_stride] or NULL;
 * sampler_params: [n_prob][n_chains][num_samples][7] or NULL.  Returns total gradient evaluations. */
typedef struct {
    const orc_problem *probs; int n_chains; const orc_nuts_args *a; int n_out_stride;
    double *outputs, *sampler_params; long *ngrad;
} sc_ctx;
static void sc_item(int t, void *c)
{
    sc_ctx *C = (sc_ctx *)c;
    int ip = t / C->n_chains, ic = t % C->n_chains;
    orc_nuts_args aa = *C->a;
    aa.chain_id = ic + 1;
    aa.seed = C->a->seed + (uint64_t)ip * 7919u;
    int nout = orc_num_outputs(&C->probs[ip]);
    int ns = aa.num_samples;
    double *tmp = NULL;
    if (C->outputs) tmp = (double *)malloc(sizeof(double) * (size_t)ns * nout);
    long ng = orc_nuts_chain(&C->probs[ip], &aa, NULL, tmp,
                             C->sampler_params ? C->sampler_params + ((size_t)t * ns) * 7 : NULL, NULL, NULL);
    if (C->outputs) {
        for (int s = 0; s < ns; ++s)
            memcpy(C->outputs + ((size_t)t * ns + s) * C->n_out_stride, tmp + (size_t)s * nout, sizeof(double) * nout);
        free(tmp);
    }
    C->ngrad[t] = ng;
}
long orc_nuts_screen(const orc_problem *probs, int n_prob, int n_chains, const orc_nuts_args *a, int n_out_stride,
                     double *outputs, double *sampler_params, int nthreads)
{
    int n = n_prob * n_chains;
    long *ng = (long *)calloc(n, sizeof(long));
    sc_ctx C = {probs, n_chains, a, n_out_stride, outputs, sampler_params, ng};
    pool_run(n, nthreads, sc_item, &C);
    long total = 0;
    for (int t = 0; t < n; ++t) if (ng[t] > 0) total += ng[t];
    free(ng);
    return total;
}

/* ---- n_prob problems of full-rank ADVI, one problem per worker ----
 * mean_u: [n_prob][ld_u] or NULL; outputs: [n_prob][output_samples][n_out_stride] or NULL; info: [n_prob][8]. */
typedef struct {
    const orc_problem *probs; const orc_advi_args *a; int ld_u, n_out_stride;
    double *mean_u, *outputs, *info;
} vb_ctx;
static void vb_item(int ip, void *c)
{
    vb_ctx *C = (vb_ctx *)c;
    orc_advi_args aa = *C->a;
    aa.seed = C->a->seed + (uint64_t)ip * 7919u;
    const orc_problem *p = &C->probs[ip];
    int nout = orc_num_outputs(p), ns = aa.output_samples;
    double *tmp = C->outputs ? (double *)malloc(sizeof(double) * (size_t)ns * nout) : NULL;
    int st = orc_advi_fullrank(p, &aa, C->mean_u ? C->mean_u + (size_t)ip * C->ld_u : NULL, NULL, tmp, NULL,
                               C->info + (size_t)ip * 8);
    if (C->outputs) {
        if (st == 0)
            for (int s = 0; s < ns; ++s)
                memcpy(C->outputs + ((size_t)ip * ns + s) * C->n_out_stride, tmp + (size_t)s * nout, sizeof(double) * nout);
        free(tmp);
    }
}
void orc_advi_screen(const orc_problem *probs, int n_prob, const orc_advi_args *a, int ld_u, int n_out_stride,
                     double *mean_u, double *outputs, double *info, int nthreads)
{
    vb_ctx C = {probs, a, ld_u, n_out_stride, mean_u, outputs, info};
    pool_run(n_prob, nthreads, vb_item, &C);
}
