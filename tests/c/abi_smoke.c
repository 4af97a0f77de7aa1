/* abi_smoke.c -- plain-C client of include/bsyn.h: compiles with a C compiler (no C++, no Python in between), links
 * libbsyn.so and calls EVERY exported entry point.
 *   abi_smoke         : no GPU needed -- host-only entry points work, compute entry points fail loudly (BSYN_ERR_CUDA
 *                       from bsyn_batch_create, BSYN_ERR_ARG on a NULL handle), nothing crashes
 *   abi_smoke --gpu   : the same calls on a real device with a small 3 x 3 experiment, results sanity-checked
 * Exit code 0 = every check passed; prints "abi_smoke ok (<n> calls)".
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bsyn.h"

static int n_calls = 0, n_fail = 0;
#define CHECK(cond)                                                                   \
    do {                                                                              \
        ++n_calls;                                                                    \
        if (!(cond)) { ++n_fail; fprintf(stderr, "FAILED line %d: %s  [%s]\n", __LINE__, #cond, bsyn_last_error()); } \
    } while (0)

int main(int argc, char **argv)
{
    const int gpu = argc > 1 && strcmp(argv[1], "--gpu") == 0;
    /* ---- a 3 x 3 experiment with 2 replicates, as rows (y, x1, x2): R/bayesynergy.R:129-154 through bsyn_prepare ---- */
    enum { NL = 4, NROW = NL * NL, NREP = 2 };
    const double conc[NL] = {0.0, 0.1, 1.0, 10.0};
    double y[NROW * NREP], x[NROW * 2];
    for (int a = 0; a < NL; ++a)
        for (int c = 0; c < NL; ++c) {
            const int r = a * NL + c;
            x[2 * r] = conc[a]; x[2 * r + 1] = conc[c];
            const double f = 1.0 / (1.0 + conc[a]) / (1.0 + 0.5 * conc[c]);
            y[r] = f + 0.01 * ((r * 7) % 5 - 2);                 /* replicate 0, column-major */
            y[NROW + r] = f - 0.01 * ((r * 3) % 5 - 2);          /* replicate 1 */
        }
    double x1[NROW], x2[NROW], yo[NROW * NREP];
    int32_t ii[NROW * NREP], n1 = 0, n2 = 0, N = 0, nrep = 0, nmiss = 0;
    CHECK(bsyn_prepare(y, x, NROW, NREP, x1, &n1, x2, &n2, yo, ii, &N, &nrep, &nmiss) == BSYN_OK);
    CHECK(n1 == 3 && n2 == 3 && N == NROW * NREP && nrep == NREP && nmiss == 0);
    CHECK(fabs(x1[0] + 1.0) < 1e-12 && fabs(x2[2] - 1.0) < 1e-12);      /* log10 doses */
    double xbad[4] = {1.0, 1.0, 2.0, 1.0}, ybad[2] = {0.5, 0.4};
    CHECK(bsyn_prepare(ybad, xbad, 2, 1, x1, &n1, x2, &n2, yo, ii, &N, &nrep, &nmiss) == BSYN_ERR_DOMAIN);
    CHECK(bsyn_prepare(y, x, NROW, NREP, x1, &n1, x2, &n2, yo, ii, &N, &nrep, &nmiss) == BSYN_OK);

    CHECK(bsyn_version() == BSYN_VERSION);
    bsyn_sampler_args sa;
    bsyn_sampler_defaults(&sa); ++n_calls;
    CHECK(sa.chains == 4 && sa.iter == 2000 && sa.warmup == 1000 && sa.adapt_delta == 0.9 && sa.max_treedepth == 10);
    bsyn_advi_args va;
    bsyn_advi_defaults(&va); ++n_calls;
    CHECK(va.iter == 10000 && va.output_samples == 1000 && va.grad_samples == 1);
    CHECK(bsyn_launch_count() >= 0);

    bsyn_problem pr;
    pr.n1 = n1; pr.n2 = n2; pr.nmissing = nmiss; pr.nrep = nrep; pr.N = N; pr.y = yo; pr.ii_obs = ii; pr.x1 = x1; pr.x2 = x2;
    bsyn_options op;
    memset(&op, 0, sizeof op);
    op.model = BSYN_MODEL_GP_GRID; op.est_la = 1; op.heteroscedastic = 1; op.kernel = 2; op.nu_matern = 2;
    op.pcprior_hypers[0] = 1; op.pcprior_hypers[1] = 0.1; op.pcprior_hypers[2] = 1; op.pcprior_hypers[3] = 0.2;
    op.rho = 0.9; op.lambda = 0.005;
    bsyn_batch *b = NULL;
    const int rc = bsyn_batch_create(&pr, 1, &op, 0, &b);
    ++n_calls;
    if (!gpu) {
        /* no device: loud failure, never a CPU fallback; every other entry point rejects the NULL handle */
        if (rc == BSYN_OK) { fprintf(stderr, "a CUDA device is present: run with --gpu\n"); bsyn_batch_destroy(b); return 2; }
        CHECK(rc == BSYN_ERR_CUDA && b == NULL && strlen(bsyn_last_error()) > 0);
        double d1[4] = {0};
        int32_t i1[4] = {0};
        char buf[32];
        float ms = 0.f;
        int64_t ng = 0;
        bsyn_sample_sinks sk;
        memset(&sk, 0, sizeof sk);
        bsyn_advi_sinks vk;
        memset(&vk, 0, sizeof vk);
        CHECK(bsyn_num_params(NULL, 0) == -1 && bsyn_num_outputs(NULL, 0) == -1 && bsyn_max_params(NULL) == -1);
        CHECK(bsyn_max_outputs(NULL) == -1 && bsyn_num_problems(NULL) == -1 && bsyn_num_param_blocks(NULL) == -1);
        CHECK(bsyn_output_name(NULL, 0, 0, buf, 32) == BSYN_ERR_ARG);
        CHECK(bsyn_param_block(NULL, 0, 0, buf, 32, i1, i1 + 1, i1 + 3) == BSYN_ERR_ARG);
        CHECK(bsyn_logp_grad(NULL, 1, i1, d1, 4, 1, d1, d1, i1) == BSYN_ERR_ARG);
        CHECK(bsyn_logp_grad_device(NULL, 1, i1, d1, 4, 1, d1, d1, i1, NULL) == BSYN_ERR_ARG);
        CHECK(bsyn_logp_grad_debug(NULL, 1, i1, d1, 4, 1, d1, d1, i1, d1, 4) == BSYN_ERR_ARG);
        CHECK(bsyn_debug_layout(NULL, i1, i1 + 1, i1 + 2) == BSYN_ERR_ARG);
        CHECK(bsyn_write_array(NULL, 1, i1, d1, 4, d1, 4, 1, i1) == BSYN_ERR_ARG);
        CHECK(bsyn_unconstrain(NULL, 0, d1, d1) == BSYN_ERR_ARG);
        CHECK(bsyn_sample(NULL, &sa, &sk, &ng) == BSYN_ERR_ARG);
        CHECK(bsyn_sample_device(NULL, &sa, &ng, &ms, NULL, NULL, NULL) == BSYN_ERR_ARG);
        bsyn_ess_target et;
        bsyn_ess_run er;
        bsyn_ess_target_defaults(&et); ++n_calls;
        CHECK(et.target_ess == 400.0 && et.q1 == BSYN_Q_VUS_SYN && et.q2 == BSYN_Q_VUS_ANT && et.retry_adapt_delta == 0.99);
        CHECK(bsyn_sample_to_ess(NULL, &sa, &et, &sk, d1, i1, &er) == BSYN_ERR_ARG);
        CHECK(bsyn_ess(NULL, 0, d1, d1) == BSYN_ERR_ARG);
        CHECK(bsyn_bulk_ess(NULL, 0, d1, d1, d1) == BSYN_ERR_ARG);
        CHECK(bsyn_advi(NULL, &va, &vk, &ng, &ms) == BSYN_ERR_ARG);
        CHECK(bsyn_measure_fp64_peak(0, d1) != BSYN_OK);
        bsyn_batch_destroy(NULL); ++n_calls;
    } else {
        CHECK(rc == BSYN_OK && b != NULL);
        if (!b) return 1;
        const int D = bsyn_num_params(b, 0), NO = bsyn_num_outputs(b, 0);
        CHECK(D == 13 + 2 + 9 && NO == D + 3 * 9 + 6 + 2 + N + 7);
        CHECK(bsyn_max_params(b) == D && bsyn_max_outputs(b) == NO && bsyn_num_problems(b) == 1);
        char buf[48];
        CHECK(bsyn_output_name(b, 0, 12, buf, 48) == BSYN_OK && strcmp(buf, "z[1,1]") == 0);
        CHECK(bsyn_output_name(b, 0, NO - 1, buf, 48) == BSYN_OK && strcmp(buf, "VUS_ant") == 0);
        int32_t nd, dm[2], off;
        CHECK(bsyn_num_param_blocks(b) == 33);
        CHECK(bsyn_param_block(b, 0, 13, buf, 48, &nd, dm, &off) == BSYN_OK && strcmp(buf, "z") == 0 && nd == 2 && dm[0] == 3 && dm[1] == 3 && off == 12);
        CHECK(bsyn_param_block(b, 0, 12, buf, 48, &nd, dm, &off) == BSYN_OK && strcmp(buf, "alpha") == 0 && nd == 1 && dm[0] == 0);
        CHECK(bsyn_param_block(b, 0, 32, buf, 48, &nd, dm, &off) == BSYN_OK && strcmp(buf, "lp__") == 0 && nd == 0 && off == NO);
        /* log_prob / grad_log_prob at two points, host and debug variants */
        double *u = calloc(2 * D, sizeof(double)), *g = calloc(2 * D, sizeof(double)), lp[2], lp2[2];
        for (int k = 0; k < 2 * D; ++k) u[k] = 0.3 * sin(k + 1.0);
        int32_t idx[2] = {0, 0}, st[2];
        CHECK(bsyn_logp_grad(b, 2, idx, u, D, 1, lp, g, st) == BSYN_OK && st[0] == 0 && st[1] == 0 && isfinite(lp[0]) && isfinite(g[3]));
        CHECK(bsyn_logp_grad(b, 2, idx, u, D, 1, lp2, NULL, st) == BSYN_OK && fabs(lp2[0] - lp[0]) < 1e-9 * fabs(lp[0]));
        int32_t nm, nn, gm;
        CHECK(bsyn_debug_layout(b, &nm, &nn, &gm) == BSYN_OK && nn == nm * nm);
        double *dbg = calloc(2 * (6 * nn + 4 * gm), sizeof(double));
        CHECK(bsyn_logp_grad_debug(b, 2, idx, u, D, 1, lp2, g, st, dbg, 6 * nn + 4 * gm) == BSYN_OK && fabs(lp2[1] - lp[1]) < 1e-9 * fabs(lp[1]));
        /* write_array + unconstrain round trip */
        double *row = calloc(2 * NO, sizeof(double)), *u2 = calloc(D, sizeof(double));
        CHECK(bsyn_write_array(b, 2, idx, u, D, row, NO, 7, st) == BSYN_OK && st[0] == 0);
        CHECK(bsyn_unconstrain(b, 0, row, u2) == BSYN_OK);
        double md = 0;
        for (int k = 0; k < D; ++k) md = fmax(md, fabs(u2[k] - u[k]));
        CHECK(md < 1e-9);
        /* sampling with every sink */
        sa.chains = 2; sa.iter = 200; sa.warmup = 100; sa.seed = 3;
        const int ns = 100, nr = 2 * ns;
        bsyn_sample_sinks sk;
        memset(&sk, 0, sizeof sk);
        float ems = 0.f;
        sk.draws = calloc((size_t)nr * NO, sizeof(double)); sk.ld_draws = NO;
        sk.udraws = calloc((size_t)nr * D, sizeof(double)); sk.ld_udraws = D;
        sk.qdraws = calloc((size_t)nr * BSYN_NQ, sizeof(double));
        sk.sampler_params = calloc((size_t)nr * BSYN_NSP, sizeof(double));
        sk.summary = calloc(BSYN_NQ * 2, sizeof(double));
        sk.cpo_mean = calloc(N, sizeof(double)); sk.ld_cpo = N;
        sk.chain_stats = calloc(2 * 8, sizeof(double));
        sk.inv_metric = calloc(2 * D, sizeof(double));
        sk.init_u = calloc(2 * D, sizeof(double));
        sk.elapsed_ms = &ems;
        int64_t ng = 0;
        CHECK(bsyn_sample(b, &sa, &sk, &ng) == BSYN_OK && ng > 0 && ems > 0.f);
        CHECK(sk.summary[2 * BSYN_Q_RVUS_F] > 0 && sk.summary[2 * BSYN_Q_RVUS_F] < 100 && sk.chain_stats[0] > 0 && sk.inv_metric[0] > 0);
        CHECK(fabs(sk.init_u[2]) <= 2.0 && sk.cpo_mean[0] > 0 && sk.sampler_params[3] >= 1);
        double ess[1], rh[1], tl[1];
        CHECK(bsyn_ess(b, BSYN_Q_RVUS_F, ess, rh) == BSYN_OK && ess[0] > 1);
        CHECK(bsyn_bulk_ess(b, BSYN_Q_RVUS_F, ess, rh, tl) == BSYN_OK && ess[0] > 1 && tl[0] > 1);
        float ms = 0.f;
        const double *dq = NULL, *dsp = NULL, *dsum = NULL;
        CHECK(bsyn_sample_device(b, &sa, &ng, &ms, &dq, &dsp, &dsum) == BSYN_OK && dq && dsp && dsum && ms > 0);
        int32_t ndraw[1];
        bsyn_ess_target et;
        bsyn_ess_run er;
        bsyn_ess_target_defaults(&et);
        et.target_ess = 50.0; et.q1 = BSYN_Q_RVUS_F; et.q2 = BSYN_Q_RVUS_P0; et.max_rounds = 4; et.retry_rounds = 2; et.retry_warmup = 50;
        sa.iter = 150; sa.warmup = 100;
        CHECK(bsyn_sample_to_ess(b, &sa, &et, NULL, ess, ndraw, &er) == BSYN_OK && er.rounds >= 1 && ndraw[0] >= 50 && er.elapsed_ms > 0 &&
              er.phase1_ms <= er.elapsed_ms);
        /* device-pointer variant: exercised with a NULL batch only (this client has no CUDA runtime of its own) */
        CHECK(bsyn_logp_grad_device(NULL, 1, idx, u, D, 1, lp, g, st, NULL) == BSYN_ERR_ARG);
        /* ADVI */
        bsyn_advi_sinks vk;
        memset(&vk, 0, sizeof vk);
        double vmean[64], vinfo[BSYN_NVI];
        vk.mean = vmean; vk.ld_mean = 64; vk.info = vinfo;
        va.iter = 400; va.output_samples = 50; va.seed = 5;
        CHECK(bsyn_advi(b, &va, &vk, &ng, &ms) == BSYN_OK && ng > 0);
        double pk = 0;
        CHECK(bsyn_measure_fp64_peak(0, &pk) == BSYN_OK && pk > 1e12);
        CHECK(bsyn_launch_count() > 0);
        bsyn_batch_destroy(b); ++n_calls;
    }
    if (n_fail) { fprintf(stderr, "abi_smoke: %d of %d checks failed\n", n_fail, n_calls); return 1; }
    printf("abi_smoke ok (%d calls, %s)\n", n_calls, gpu ? "gpu" : "no gpu");
    return 0;
}
