// bsyn_stanfit.hpp -- the `call_sampler` / `param_*` contract of rstan::stan_fit<model, RNG>, assembled from the C ABI
// of bsyn.h in plain C++ (no Rcpp, no R headers), so that the Rcpp module of INTEGRATION.md only has to wrap these
// containers into SEXPs -- and so that the contract can be compiled and tested without R (tests/c/stanfit_smoke.cpp).
//
// What it mirrors (reference src/stanExports_gp_grid.cc:15-30 -> rstan/stan_fit.hpp [un-vendored]):
//   call_sampler(args)  -> one named numeric vector per flat parameter of interest (order: src/stanExports_gp_grid.h
//                          :2256-2402, then "lp__"), each of length n_save, with attributes
//                          sampler_params (accept_stat__, stepsize__, treedepth__, n_leapfrog__, divergent__, energy__),
//                          adaptation_info (step size + diagonal of the inverse metric, Stan's text format),
//                          elapsed_time (warmup, sample), inits, mean_pars, mean_lp__, args
//   param_names / param_dims / param_fnames_oi / update_param_oi / param_oi_tidx  -> ParamIndex
// One call = one chain of one experiment, as rstan drives it (R/bayesynergy.R:221-223 -> rstan::sampling loops over
// chains); the batched screen path (bsyn_sample over thousands of experiments x chains) does not go through here.
#ifndef BSYN_STANFIT_HPP
#define BSYN_STANFIT_HPP

#include <cstdint>
#include <cstdio>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "bsyn.h"

namespace bsyn_stanfit {

// the `args` list rstan passes to call_sampler (method = "sampling", algorithm = "NUTS")
struct SamplerArgs {
    int chain_id = 1, iter = 2000, warmup = 1000, thin = 1;
    std::uint64_t seed = 1;
    double init_r = 2.0;
    bool save_warmup = true;                 // rstan::sampling's default
    // control list
    double adapt_delta = 0.9, stepsize = 1.0, stepsize_jitter = 0.0;
    int max_treedepth = 10;
    std::string metric = "diag_e";           // or "dense_e"
    double adapt_gamma = 0.05, adapt_kappa = 0.75, adapt_t0 = 10.0;
    int adapt_init_buffer = 75, adapt_term_buffer = 50, adapt_window = 25;
};

struct ChainFit {
    std::vector<std::string> fnames;                  // flat names of the parameters of interest, "lp__" last
    std::vector<std::vector<double>> samples;         // samples[i] = the draws of fnames[i], length n_save
    std::vector<std::string> sampler_param_names;     // accept_stat__ ... energy__
    std::vector<std::vector<double>> sampler_params;  // 6 x n_save
    std::string adaptation_info;
    double elapsed_warmup = 0, elapsed_sample = 0;    // seconds
    std::vector<double> inits;                        // write_array row at the chain's initial point (all flat parameters)
    std::vector<double> mean_pars;                    // post-warmup mean of every flat parameter (without lp__)
    double mean_lp = 0;
    int n_save = 0, n_warmup_saved = 0;
    SamplerArgs args;
};

// param_names / param_dims / param_fnames_oi / update_param_oi / param_oi_tidx over the blocks of bsyn_param_block
class ParamIndex {
public:
    struct Block { std::string name; std::vector<int> dims; int offset, len; };
    ParamIndex(const bsyn_batch *b, int problem)
    {
        char buf[64];
        const int nb = bsyn_num_param_blocks(b);
        if (nb < 0) throw std::runtime_error("bsyn_stanfit: bad batch handle");
        for (int i = 0; i < nb; ++i) {
            std::int32_t nd = 0, dm[2] = {0, 0}, off = 0;
            if (bsyn_param_block(b, problem, i, buf, sizeof buf, &nd, dm, &off) != BSYN_OK) throw std::runtime_error(bsyn_last_error());
            Block k;
            k.name = buf; k.offset = off; k.len = 1;
            for (int d = 0; d < nd; ++d) { k.dims.push_back(dm[d]); k.len *= dm[d]; }
            blocks_.push_back(k);
        }
        n_out_ = bsyn_num_outputs(b, problem);
        for (int i = 0; i < n_out_; ++i) {
            if (bsyn_output_name(b, problem, i, buf, sizeof buf) != BSYN_OK) throw std::runtime_error(bsyn_last_error());
            flat_.push_back(buf);
        }
        flat_.push_back("lp__");
        oi_.resize(blocks_.size());
        for (size_t i = 0; i < blocks_.size(); ++i) oi_[i] = (int)i;
    }
    std::vector<std::string> param_names() const { std::vector<std::string> o; for (auto &k : blocks_) o.push_back(k.name); return o; }
    std::vector<std::vector<int>> param_dims() const { std::vector<std::vector<int>> o; for (auto &k : blocks_) o.push_back(k.dims); return o; }
    std::vector<std::string> param_names_oi() const { std::vector<std::string> o; for (int i : oi_) o.push_back(blocks_[i].name); return o; }
    std::vector<std::vector<int>> param_dims_oi() const { std::vector<std::vector<int>> o; for (int i : oi_) o.push_back(blocks_[i].dims); return o; }
    // select the parameters of interest ("lp__" is always kept, as in rstan); returns how many blocks are selected
    int update_param_oi(const std::vector<std::string> &pars)
    {
        std::vector<int> sel;
        for (auto &p : pars) {
            int f = -1;
            for (size_t i = 0; i < blocks_.size(); ++i) if (blocks_[i].name == p) f = (int)i;
            if (f < 0) throw std::invalid_argument("parameter " + p + " not found");
            bool dup = false;
            for (int s : sel) dup = dup || s == f;
            if (!dup) sel.push_back(f);
        }
        const int lp = (int)blocks_.size() - 1;
        bool has = false;
        for (int s : sel) has = has || s == lp;
        if (!has) sel.push_back(lp);
        oi_ = sel;
        return (int)oi_.size();
    }
    // flat indices (into a write_array row followed by lp__) of the selected parameters, in selection order
    std::vector<int> flat_index_oi() const
    {
        std::vector<int> o;
        for (int i : oi_) for (int e = 0; e < blocks_[i].len; ++e) o.push_back(blocks_[i].offset + e);
        return o;
    }
    std::vector<std::string> param_fnames_oi() const { std::vector<std::string> o; for (int i : flat_index_oi()) o.push_back(flat_[i]); return o; }
    // name -> positions among the flat parameters of interest (0-based; the R binding adds 1)
    std::map<std::string, std::vector<int>> param_oi_tidx(const std::vector<std::string> &names) const
    {
        std::map<std::string, std::vector<int>> o;
        for (auto &nm : names) {
            int pos = 0;
            bool found = false;
            for (int i : oi_) {
                if (blocks_[i].name == nm) { for (int e = 0; e < blocks_[i].len; ++e) o[nm].push_back(pos + e); found = true; }
                pos += blocks_[i].len;
            }
            if (!found) throw std::invalid_argument("parameter " + nm + " not found");
        }
        return o;
    }
    int n_outputs() const { return n_out_; }
    const std::vector<std::string> &flat_names() const { return flat_; }

private:
    std::vector<Block> blocks_;
    std::vector<std::string> flat_;
    std::vector<int> oi_;
    int n_out_ = 0;
};

// call_sampler for a batch that holds ONE experiment.  Throws std::runtime_error with bsyn_last_error() on failure
// (the Rcpp wrapper turns it into an R error, which bayesynergy() maps to returnCode 2).
inline ChainFit call_sampler(bsyn_batch *b, const SamplerArgs &a, const ParamIndex &pi)
{
    if (bsyn_num_problems(b) != 1) throw std::invalid_argument("call_sampler: the module object wraps one experiment");
    if (a.metric != "diag_e" && a.metric != "dense_e") throw std::invalid_argument("control$metric must be diag_e or dense_e");
    bsyn_sampler_args sa;
    bsyn_sampler_defaults(&sa);
    sa.chains = 1; sa.iter = a.iter; sa.warmup = a.warmup; sa.thin = a.thin;
    // Stan advances one RNG by chain_id; here the chain id is folded into the Philox key
    sa.seed = a.seed + 0x9E3779B97F4A7C15ull * (std::uint64_t)a.chain_id;
    sa.init_r = a.init_r; sa.adapt_delta = a.adapt_delta; sa.stepsize = a.stepsize; sa.stepsize_jitter = a.stepsize_jitter;
    sa.max_treedepth = a.max_treedepth; sa.metric = a.metric == "dense_e" ? 1 : 0;
    sa.adapt_gamma = a.adapt_gamma; sa.adapt_kappa = a.adapt_kappa; sa.adapt_t0 = a.adapt_t0;
    sa.adapt_init_buffer = a.adapt_init_buffer; sa.adapt_term_buffer = a.adapt_term_buffer; sa.adapt_window = a.adapt_window;
    sa.save_warmup = a.save_warmup ? 1 : 0;
    const int NO = pi.n_outputs(), D = bsyn_num_params(b, 0);
    const int ns_samp = (a.iter - a.warmup + a.thin - 1) / a.thin, ns_warm = a.save_warmup ? (a.warmup + a.thin - 1) / a.thin : 0;
    const int ns = ns_samp + ns_warm;
    std::vector<double> draws((size_t)ns * NO), sp((size_t)ns * BSYN_NSP), cs(8), im(D), iu(D);
    float ems = 0.f;
    bsyn_sample_sinks sk = bsyn_sample_sinks();
    sk.draws = draws.data(); sk.ld_draws = NO;
    sk.sampler_params = sp.data(); sk.chain_stats = cs.data(); sk.inv_metric = im.data(); sk.init_u = iu.data();
    sk.elapsed_ms = &ems;
    std::int64_t ng = 0;
    if (bsyn_sample(b, &sa, &sk, &ng) != BSYN_OK) throw std::runtime_error(bsyn_last_error());
    if (cs[6] != 0.0) throw std::runtime_error("Initialization failed.");   // rstan: error after 100 failed attempts
    ChainFit f;
    f.args = a; f.n_save = ns; f.n_warmup_saved = ns_warm;
    const std::vector<int> idx = pi.flat_index_oi();
    f.fnames = pi.param_fnames_oi();
    f.samples.assign(idx.size(), std::vector<double>(ns));
    for (size_t c = 0; c < idx.size(); ++c)
        for (int i = 0; i < ns; ++i)
            f.samples[c][i] = idx[c] < NO ? draws[(size_t)i * NO + idx[c]] : sp[(size_t)i * BSYN_NSP + 6];   // lp__
    static const char *spn[6] = {"accept_stat__", "stepsize__", "treedepth__", "n_leapfrog__", "divergent__", "energy__"};
    f.sampler_params.assign(6, std::vector<double>(ns));
    for (int k = 0; k < 6; ++k) {
        f.sampler_param_names.push_back(spn[k]);
        for (int i = 0; i < ns; ++i) f.sampler_params[k][i] = sp[(size_t)i * BSYN_NSP + k];
    }
    // mean_pars / mean_lp__: over the post-warmup draws
    f.mean_pars.assign(NO, 0.0);
    for (int i = ns_warm; i < ns; ++i) {
        for (int c = 0; c < NO; ++c) f.mean_pars[c] += draws[(size_t)i * NO + c];
        f.mean_lp += sp[(size_t)i * BSYN_NSP + 6];
    }
    for (int c = 0; c < NO; ++c) f.mean_pars[c] /= ns_samp > 0 ? ns_samp : 1;
    f.mean_lp /= ns_samp > 0 ? ns_samp : 1;
    // inits: the initial point through write_array
    {
        std::vector<double> row(NO);
        std::int32_t zero = 0, st = 0;
        if (bsyn_write_array(b, 1, &zero, iu.data(), D, row.data(), NO, a.seed, &st) != BSYN_OK) throw std::runtime_error(bsyn_last_error());
        f.inits = row;
    }
    // adaptation_info in Stan's text format (stan::mcmc::adapt_diag_e_nuts::write_sampler_state)
    {
        char buf[64];
        std::string s = "# Adaptation terminated\n# Step size = ";
        std::snprintf(buf, sizeof buf, "%g", cs[0]);
        s += buf;
        s += a.metric == "dense_e" ? "\n# Diagonal of the dense inverse mass matrix:\n# " : "\n# Diagonal elements of inverse mass matrix:\n# ";
        for (int d = 0; d < D; ++d) { std::snprintf(buf, sizeof buf, d ? ", %g" : "%g", im[d]); s += buf; }
        s += "\n";
        f.adaptation_info = s;
    }
    // elapsed_time: the device time of the launch, split by the leapfrog counts of the two phases
    {
        const double tot = 1e-3 * ems, lw = cs[3] - cs[4], lsamp = cs[4];
        const double fw = (lw + lsamp) > 0 ? lw / (lw + lsamp) : 0.0;
        f.elapsed_warmup = tot * fw; f.elapsed_sample = tot * (1.0 - fw);
    }
    return f;
}

}  // namespace bsyn_stanfit
#endif
