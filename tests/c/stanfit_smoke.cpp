// stanfit_smoke.cpp -- the call_sampler / param_* contract of include/bsyn_stanfit.hpp on a real device (GPU only):
// shapes, names and bookkeeping that rstan::extract / summary / get_sampler_params rely on.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "bsyn_stanfit.hpp"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "FAILED line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main()
{
    enum { NL = 4, NROW = NL * NL, NREP = 2 };
    const double conc[NL] = {0.0, 0.1, 1.0, 10.0};
    double y[NROW * NREP], x[NROW * 2];
    for (int a = 0; a < NL; ++a)
        for (int c = 0; c < NL; ++c) {
            const int r = a * NL + c;
            x[2 * r] = conc[a]; x[2 * r + 1] = conc[c];
            const double f = 1.0 / (1.0 + conc[a]) / (1.0 + 0.5 * conc[c]);
            y[r] = f + 0.01 * ((r * 7) % 5 - 2);
            y[NROW + r] = f - 0.01 * ((r * 3) % 5 - 2);
        }
    double x1[NROW], x2[NROW], yo[NROW * NREP];
    std::int32_t ii[NROW * NREP], n1, n2, N, nrep, nmiss;
    CHECK(bsyn_prepare(y, x, NROW, NREP, x1, &n1, x2, &n2, yo, ii, &N, &nrep, &nmiss) == BSYN_OK);
    bsyn_problem pr = {n1, n2, nmiss, nrep, N, yo, ii, x1, x2};
    bsyn_options op;
    std::memset(&op, 0, sizeof op);
    op.model = BSYN_MODEL_GP_GRID; op.est_la = 1; op.heteroscedastic = 1; op.kernel = 2; op.nu_matern = 2;
    op.rho = 0.9; op.lambda = 0.005;
    bsyn_batch *b = nullptr;
    CHECK(bsyn_batch_create(&pr, 1, &op, 0, &b) == BSYN_OK);
    using namespace bsyn_stanfit;
    ParamIndex pi(b, 0);
    // param_names / param_dims (src/stanExports_gp_grid.h:1588-1704)
    auto names = pi.param_names();
    auto dims = pi.param_dims();
    CHECK(names.size() == 33 && names[0] == "la_1" && names[12] == "alpha" && names[13] == "z" && names.back() == "lp__");
    CHECK(dims[0].size() == 1 && dims[0][0] == 1 && dims[12][0] == 0 && dims[13].size() == 2 && dims[13][0] == 3 && dims[13][1] == 3);
    CHECK(dims[2].empty());
    const int NO = bsyn_num_outputs(b, 0);
    CHECK((int)pi.param_fnames_oi().size() == NO + 1 && pi.param_fnames_oi()[12] == "z[1,1]" && pi.param_fnames_oi().back() == "lp__");
    SamplerArgs a;
    a.iter = 300; a.warmup = 150; a.chain_id = 2; a.seed = 11;
    ChainFit f = call_sampler(b, a, pi);
    CHECK(f.n_save == 300 && f.n_warmup_saved == 150 && (int)f.samples.size() == NO + 1 && f.fnames == pi.param_fnames_oi());
    for (auto &v : f.samples) CHECK((int)v.size() == 300);
    CHECK(f.sampler_params.size() == 6 && f.sampler_param_names[3] == "n_leapfrog__" && (int)f.sampler_params[0].size() == 300);
    CHECK((int)f.inits.size() == NO && (int)f.mean_pars.size() == NO && std::isfinite(f.mean_lp));
    // mean_pars = post-warmup mean of the returned draws
    double m = 0;
    for (int i = 150; i < 300; ++i) m += f.samples[2][i];
    CHECK(std::fabs(m / 150 - f.mean_pars[2]) < 1e-12 * (1 + std::fabs(m)));
    CHECK(f.adaptation_info.find("# Step size = ") != std::string::npos && f.adaptation_info.find("inverse mass matrix") != std::string::npos);
    CHECK(f.elapsed_warmup > 0 && f.elapsed_sample > 0);
    // step size is constant after warm-up and equals the adapted one
    CHECK(f.sampler_params[1][299] == f.sampler_params[1][200] && f.sampler_params[1][299] > 0);
    // a different chain_id gives a different chain
    a.chain_id = 3;
    ChainFit g = call_sampler(b, a, pi);
    CHECK(g.samples[2][299] != f.samples[2][299]);
    // update_param_oi / param_oi_tidx (what rstan uses for pars = c(...))
    CHECK(pi.update_param_oi({"VUS_syn", "z"}) == 3);
    auto fn = pi.param_fnames_oi();
    CHECK(fn.size() == 1 + 9 + 1 && fn[0] == "VUS_syn" && fn[1] == "z[1,1]" && fn[10] == "lp__");
    auto tidx = pi.param_oi_tidx({"z", "lp__"});
    CHECK(tidx["z"].size() == 9 && tidx["z"][0] == 1 && tidx["lp__"][0] == 10);
    ChainFit h = call_sampler(b, a, pi);
    CHECK(h.samples.size() == 11 && h.fnames == fn);
    CHECK(h.samples[0][299] == g.samples[NO - 2][299]);        // same seed and chain: VUS_syn column is the same draw
    bsyn_batch_destroy(b);
    std::printf("stanfit_smoke ok\n");
    return 0;
}
